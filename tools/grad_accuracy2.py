#!/usr/bin/env python3
"""Where does the H=256 gradient error enter?  Gradients w.r.t. the hidden states after every layer, CUDA vs fp64 oracle."""
import os, sys, torch
import torch.nn.functional as F
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [REPO, os.path.join(REPO, "deep-fem-uav-wing_b200")]
from deep_fem_uav_wing.gnn import ops, synth
from deep_fem_uav_wing.gnn.model import GraphSAGEModel, MaskedMSELoss
from oracle.sage_oracle import GraphSAGEModelRef, MaskedMSELossRef
H, L = int(sys.argv[1]), int(sys.argv[2])
mesh = synth.surface_tri_wing(6000, seed=3)
x, ei = torch.from_numpy(mesh["x"]), torch.from_numpy(mesh["edge_index"])
y, m = torch.from_numpy(mesh["y"]), torch.from_numpy(mesh["loss_mask"])
torch.manual_seed(11)
ref = GraphSAGEModelRef(10, H, 1, L, dropout=0.0)
model = GraphSAGEModel(10, H, 1, L, dropout=0.0)
model.load_state_dict(ref.state_dict())
model = model.cuda()
r64 = GraphSAGEModelRef(10, H, 1, L, dropout=0.0).double()
r64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
# oracle, keeping the hidden states
hs64 = []
h = r64.encoder(x.double()); h.retain_grad(); hs64.append(h)
for conv, norm in zip(r64.convs, r64.norms):
    h = h + F.relu(norm(conv(h, ei))); h.retain_grad(); hs64.append(h)
MaskedMSELossRef()(r64.decoder(h), y.double(), m).backward()
# CUDA model, same wiring as GraphSAGEModel.forward
xc, eic = x.cuda(), ei.cuda()
graph = ops.get_graph(eic, xc.shape[0])
hs = []
e0, e2 = model.encoder[0], model.encoder[2]
hc = ops.LinearFn.apply(ops.LinearFn.apply(xc, e0.weight, e0.bias, True, 0.0, 0), e2.weight, e2.bias, True, 0.0, 0); hc.retain_grad(); hs.append(hc)
for conv, norm in zip(model.convs, model.norms):
    hc = ops.SageConvFn.apply(hc, conv.lin_l.weight, conv.lin_l.bias, conv.lin_r.weight, norm.weight, norm.bias, graph, float(norm.eps), 0.0, 0, True)
    hc.retain_grad(); hs.append(hc)
d0, d3 = model.decoder[0], model.decoder[3]
out = ops.DecoderTailFn.apply(hc, d0.weight, d0.bias, d3.weight, d3.bias, 0.0, 0)
MaskedMSELoss()(out, y.cuda(), m.cuda()).backward()
rl2 = lambda a, b: ((a.double().cpu() - b).norm() / b.norm()).item()
for i, (a, b) in enumerate(zip(hs, hs64)):
    ga, gb = a.grad.double().cpu(), b.grad
    print(f"h{i}: value err {rl2(a.detach(), b.detach()):.2e}  grad err {rl2(a.grad, b.grad):.2e}  |grad| {gb.norm().item():.3e}  "
          f"col-mean err {((ga.mean(0) - gb.mean(0)).norm() / gb.mean(0).norm()).item():.2e}  max|grad| {gb.abs().max().item():.2e}")

# ---- layer 0 backward, block by block, each against fp64 arithmetic on ITS OWN (CUDA) inputs ----
conv, norm = model.convs[0], model.norms[0]
h0 = hs[0].detach()
g_out = hs[1].grad.detach()
agg = ops.aggregate(graph.rowptr, graph.col, graph.inv_deg, h0)
wl, wr, bl = conv.lin_l.weight.detach(), conv.lin_r.weight.detach(), conv.lin_l.bias.detach()
gam, bet = norm.weight.detach(), norm.bias.detach()
out, pre, stats, _ = ops.linear_fwd(agg, wl, h0, wr, bias=bl, ln=(gam, bet), eps=float(norm.eps), relu=True, residual=h0, save_pre=True)
g_y, dgamma, dbeta, _, _, dbl = ops.epilogue_bwd(g_out, h0.shape[0], H, g_out, pre=pre, stats=stats, ln=(gam, bet), relu=True, want_bias_grad=True)
d = lambda t: t.double()
pre64 = d(agg) @ d(wl).T + d(h0) @ d(wr).T + d(bl)
mean64 = pre64.mean(1, keepdim=True); rstd64 = 1.0 / torch.sqrt(pre64.var(1, unbiased=False, keepdim=True) + float(norm.eps))
xh = (pre64 - mean64) * rstd64
gg = d(g_out) * ((xh * d(gam) + d(bet)) > 0) * d(gam)
gy64 = rstd64 * (gg - gg.mean(1, keepdim=True) - xh * (gg * xh).mean(1, keepdim=True))
r = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
print(f"layer0: pre {r(pre, pre64):.2e} mean {r(stats[:, :1], mean64):.2e} rstd {r(stats[:, 1:], rstd64):.2e}  g_y {r(g_y, gy64):.2e}  |g_y| {gy64.norm().item():.3e} |gg| {gg.norm().item():.3e} "
      f"rstd range [{rstd64.min().item():.2e}, {rstd64.max().item():.2e}]")
rp_t, col_t = graph.transpose()
g_t = ops.aggregate_scaled(rp_t, col_t, graph.inv_deg, g_y)
src, dst = eic[0], eic[1]
gt64 = torch.zeros_like(gy64).index_add_(0, src, (d(g_y) * d(graph.inv_deg)[:, None])[dst])
g_x, _, _, _ = ops.linear_fwd(g_t, wl.t().contiguous(), g_y, wr.t().contiguous(), residual=g_out)
gx64 = d(g_t) @ d(wl) + d(g_y) @ d(wr) + d(g_out)
print(f"        g_t {r(g_t, gt64):.2e}  g_x (own inputs) {r(g_x, gx64):.2e}  |g_x| {gx64.norm().item():.3e}  |g_t W_l| {(d(g_t) @ d(wl)).norm().item():.3e} |g_y W_r| {(d(g_y) @ d(wr)).norm().item():.3e}")
print(f"        g_y abs max {g_y.abs().max().item():.3e} min nonzero {g_y[g_y != 0].abs().min().item():.3e}; rows with huge rstd: {(rstd64 > 100).sum().item()}")
# split the g_y error: same fp64 formula fed with the CUDA kernel's own saved tensors (pre, stats)
xh_c = (d(pre) - d(stats[:, :1])) * d(stats[:, 1:])
mask_c = (xh_c * d(gam) + d(bet)) > 0
mask_64 = (xh * d(gam) + d(bet)) > 0
gg_c = d(g_out) * mask_c * d(gam)
gy_c = d(stats[:, 1:]) * (gg_c - gg_c.mean(1, keepdim=True) - xh_c * (gg_c * xh_c).mean(1, keepdim=True))
print(f"        fp64 formula on CUDA-saved pre/stats vs pure fp64: {r(gy_c, gy64):.2e}; kernel vs fp64-on-its-own-inputs: {r(g_y, gy_c):.2e}; "
      f"ReLU mask mismatches {(mask_c != mask_64).sum().item()} of {mask_64.numel()}; |mean| max {mean64.abs().max().item():.2e} |pre| max {pre64.abs().max().item():.2e}")
