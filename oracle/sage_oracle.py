"""CPU oracle for the GraphSAGE hot path of Deep-FEM-UAV-Wing.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``deep-fem-uav-wing_b200/``) never imports
anything from ``oracle/`` and has no CPU fallback.

Parity status
-------------
* Model wiring (encoder -> L x [conv, LayerNorm, ReLU, dropout, residual] -> decoder),
  ``MaskedMSELoss`` and ``compute_metrics`` are PINNED: ``tests/golden/make_golden.py``
  imports the reference's own ``src/deep_fem_uav_wing/gnn/model.py`` in the build
  container and records its outputs; ``tests/test_oracle_golden.py`` replays them.
* ``_faces_to_edge_index`` / ``build_graph_data`` are PINNED the same way (the reference
  functions are pure numpy and import here).
* ``SAGEConv`` itself is **parity unpinned**: the arithmetic lives in the third-party
  package ``torch-geometric`` (``requirements-gnn.txt:20-21``: ``torch-geometric>=2.4.0``,
  given as a comment only - no lock file, not vendored, not installed in this image, no
  network).  ``SAGEConvRef`` restates its published algorithm for the configuration the
  reference constructs at ``model.py:63`` (``SAGEConv(H, H)`` -> ``aggr='mean'``,
  ``root_weight=True``, ``bias=True``, ``normalize=False``, ``project=False``):

      out_i = W_l . mean_{j in N_in(i)} x_j + b_l + W_r . x_i

  with ``edge_index[0]`` = source j, ``edge_index[1]`` = target i (flow
  ``source_to_target``), zero for nodes with no incoming edge, duplicates counted with
  multiplicity, self-loops treated as ordinary edges.  If ``torch_geometric`` ever becomes
  importable, ``pyg_selfcheck()`` compares this restatement with the real layer.

Reference anchors: ``src/deep_fem_uav_wing/gnn/model.py`` (call sites ``:15,:63,:90``).
"""

from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# CSR oracle (bit-exact target for kernel (a))
# --------------------------------------------------------------------------------------
def csr_oracle(edge_index: np.ndarray, num_nodes: int, by: str = "dst"):
    """Canonical CSR of ``edge_index`` grouped by destination (``by='dst'``) or source.

    Returns ``(rowptr int32[N+1], col int32[E], perm int32[E], inv_deg float32[N])``.

    ``by='dst'``: row i lists the sources of the edges arriving at i, ordered by
    (src, original edge position) - i.e. ``np.lexsort((src, dst))`` (stable), which is what
    PyG's ``sort_edge_index(..., sort_by_row=False)`` produces.  ``perm[k]`` is the position
    in the input ``edge_index`` of the k-th sorted edge.  ``inv_deg = 1/max(deg, 1)``
    (isolated node -> mean 0, PyG's ``clamp_(min=1)``).
    """
    ei = np.asarray(edge_index)
    assert ei.ndim == 2 and ei.shape[0] == 2
    src = ei[0].astype(np.int64)
    dst = ei[1].astype(np.int64)
    if by == "src":
        src, dst = dst, src
    elif by != "dst":
        raise ValueError(by)
    order = np.lexsort((src, dst))  # primary key dst, secondary src, stable
    col = src[order].astype(np.int32)
    deg = np.bincount(dst, minlength=num_nodes).astype(np.int64)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    inv_deg = (1.0 / np.maximum(deg, 1)).astype(np.float32)
    return rowptr.astype(np.int32), col, order.astype(np.int32), inv_deg


# --------------------------------------------------------------------------------------
# SAGEConv restatement  (PyG, un-vendored - see module docstring)
# --------------------------------------------------------------------------------------
class SAGEConvRef(nn.Module):
    """Stock-torch restatement of ``torch_geometric.nn.SAGEConv(in, out)`` defaults.

    State-dict keys: ``lin_l.weight [out,in]``, ``lin_l.bias [out]``, ``lin_r.weight
    [out,in]`` (``lin_r`` has no bias) - the names checkpoints of the reference carry
    (``scripts/train_gnn.py:224-231``).
    """

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lin_l = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        # PyG Linear: kaiming_uniform(a=sqrt(5)) on the weight == U(-1/sqrt(in), 1/sqrt(in)),
        # bias U(-1/sqrt(in), 1/sqrt(in)).  Draw order differs from PyG's, so identical
        # weights across implementations are obtained by copying a state_dict only.
        bound = 1.0 / math.sqrt(self.in_channels)
        for w in (self.lin_l.weight, self.lin_r.weight, self.lin_l.bias):
            nn.init.uniform_(w, -bound, bound)

    @staticmethod
    def mean_aggregate(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        """agg_i = mean over incoming edges (j -> i) of x_j; 0 if there is none."""
        n = x.shape[0]
        src, dst = edge_index[0], edge_index[1]
        msg = x.index_select(0, src)
        summed = torch.zeros_like(x).index_add_(0, dst, msg)
        cnt = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(
            0, dst, torch.ones(dst.shape[0], dtype=x.dtype, device=x.device)
        )
        return summed / cnt.clamp_(min=1).unsqueeze(-1)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        agg = self.mean_aggregate(x, edge_index)
        return self.lin_l(agg) + self.lin_r(x)


# --------------------------------------------------------------------------------------
# Model restatement  (model.py:24-112)
# --------------------------------------------------------------------------------------
class GraphSAGEModelRef(nn.Module):
    """Restates ``GraphSAGEModel`` (``model.py:35-99``) on top of ``SAGEConvRef``."""

    def __init__(self, in_channels=10, hidden_channels=128, out_channels=1, num_layers=4, dropout=0.1):
        super().__init__()
        self.in_channels = in_channels
        self.hidden_channels = hidden_channels
        self.out_channels = out_channels
        self.num_layers = num_layers
        self.dropout = dropout
        # model.py:52-57
        self.encoder = nn.Sequential(
            nn.Linear(in_channels, 64), nn.ReLU(), nn.Linear(64, hidden_channels), nn.ReLU()
        )
        # model.py:60-64
        self.convs = nn.ModuleList()
        self.norms = nn.ModuleList()
        for _ in range(num_layers):
            self.convs.append(SAGEConvRef(hidden_channels, hidden_channels))
            self.norms.append(nn.LayerNorm(hidden_channels))
        # model.py:67-72
        self.decoder = nn.Sequential(
            nn.Linear(hidden_channels, 64), nn.ReLU(), nn.Dropout(dropout), nn.Linear(64, out_channels)
        )

    def forward(self, x, edge_index, batch=None):  # batch accepted and ignored (model.py:74)
        h = self.encoder(x)
        for conv, norm in zip(self.convs, self.norms):  # model.py:89-95
            h_new = conv(h, edge_index)
            h_new = norm(h_new)
            h_new = F.relu(h_new)
            h_new = F.dropout(h_new, p=self.dropout, training=self.training)
            h = h + h_new
        return self.decoder(h)  # model.py:98

    def predict(self, data):  # model.py:101-112
        self.eval()
        with torch.no_grad():
            return self.forward(data.x, data.edge_index, getattr(data, "batch", None))


# --------------------------------------------------------------------------------------
# Train-mode parity: the reference draws dropout masks from torch's RNG (model.py:70,93), which no other
# implementation can reproduce, so train-mode tests run the oracle with the PRODUCT's masks.  This is a numpy
# restatement of the product's counter RNG (deep-fem-uav-wing_b200/csrc/dfw_common.cuh: mix32 / dropout_row_key /
# dropout_bits) - test infrastructure like the rest of this module.
# --------------------------------------------------------------------------------------
def _mix32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x.astype(np.uint64) * np.uint64(0x21F0AAAD) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x.astype(np.uint64) * np.uint64(0x735A2D97) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    return x


def dropout_keep_scale(seed: int, n_rows: int, n_cols: int, p: float) -> np.ndarray:
    """[n_rows, n_cols] float32: 1/(1-p) where the product keeps the unit, 0 where it drops it."""
    seed &= 0xFFFFFFFFFFFFFFFF
    rows = np.arange(n_rows, dtype=np.uint64)
    k = _mix32(((rows & np.uint64(0xFFFFFFFF)) ^ np.uint64(seed & 0xFFFFFFFF)).astype(np.uint32))
    k = ((k.astype(np.uint64) + (rows >> np.uint64(32)) * np.uint64(0x85EBCA6B) + np.uint64(seed >> 32)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    k = _mix32(k)
    cols = (np.arange(n_cols, dtype=np.uint64) * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF)
    bits = _mix32(((k.astype(np.uint64)[:, None] + cols[None, :]) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
    thr = min(max(int(float(np.float32(p)) * 4294967296.0), 0), 4294967295)  # dropout_threshold()
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(bits >= np.uint32(thr), scale, np.float32(0.0)).astype(np.float32)


def forward_with_masks(model: "GraphSAGEModelRef", x, edge_index, layer_masks, decoder_mask):
    """``GraphSAGEModelRef.forward`` in training mode with the dropout factors given (model.py:86-98)."""
    h = model.encoder(x)
    for conv, norm, m in zip(model.convs, model.norms, layer_masks):
        h_new = F.relu(norm(conv(h, edge_index)))
        h = h + h_new * m
    dec = model.decoder
    hid = F.relu(dec[0](h)) * decoder_mask
    return dec[3](hid)


class MaskedMSELossRef(nn.Module):
    """Restates ``MaskedMSELoss`` (``model.py:115-153``)."""

    def __init__(self, reduction: str = "mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, pred, target, mask=None):
        if mask is None:
            return F.mse_loss(pred, target, reduction=self.reduction)
        if mask.dim() == 1:
            mask = mask.unsqueeze(-1)
        mp = pred[mask.expand_as(pred)].view(-1, pred.size(-1))
        mt = target[mask.expand_as(target)].view(-1, target.size(-1))
        if mp.numel() == 0:
            return torch.tensor(0.0, device=pred.device, requires_grad=True)
        return F.mse_loss(mp, mt, reduction=self.reduction)


def compute_metrics_ref(pred, target, mask=None, log_scale=True):
    """Restates ``compute_metrics`` (``model.py:156-216``)."""
    p = pred.detach().cpu().numpy().flatten()
    t = target.detach().cpu().numpy().flatten()
    m = mask.detach().cpu().numpy().flatten() if mask is not None else None
    if log_scale:
        p, t = np.expm1(p), np.expm1(t)

    def subset(pp, tt, mm):
        if mm is not None:
            pp, tt = pp[mm], tt[mm]
        if len(pp) == 0:
            return {"mae": 0.0, "rmse": 0.0, "max_error": 0.0, "count": 0}
        err = np.abs(pp - tt)
        return {
            "mae": float(np.mean(err)),
            "rmse": float(np.sqrt(np.mean(err**2))),
            "max_error": float(np.max(err)),
            "count": int(len(pp)),
        }

    return {"all_nodes": subset(p, t, None), "masked_nodes": subset(p, t, m)}


# --------------------------------------------------------------------------------------
# Graph construction restatement (dataset.py:26-63) - vectorised; same edge SET as the
# reference, which emits Python-set iteration order (compare after canonical sorting).
# --------------------------------------------------------------------------------------
def faces_to_edge_index_ref(faces, node_id_to_idx) -> np.ndarray:
    edges = set()
    for face in faces:
        try:
            a, b, c = (node_id_to_idx[n] for n in face)
        except KeyError:
            continue  # dataset.py:43-46: faces touching unknown ids are skipped
        for u, v in ((a, b), (b, c), (c, a)):
            edges.add((min(u, v), max(u, v)))
    if not edges:
        return np.zeros((2, 0), dtype=np.int64)
    und = np.array(sorted(edges), dtype=np.int64)
    out = np.empty((2, 2 * len(und)), dtype=np.int64)
    out[0, 0::2], out[1, 0::2] = und[:, 0], und[:, 1]
    out[0, 1::2], out[1, 1::2] = und[:, 1], und[:, 0]
    return out


def pyg_selfcheck(n=200, e=1200, h=16, seed=0):
    """If real PyG is importable, return max|SAGEConvRef - SAGEConv|; else None."""
    try:
        from torch_geometric.nn import SAGEConv  # type: ignore
    except Exception:
        return None
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, generator=g)
    ei = torch.randint(0, n, (2, e), generator=g)
    real = SAGEConv(h, h)
    mine = SAGEConvRef(h, h)
    mine.load_state_dict(real.state_dict())
    return float((real(x, ei) - mine(x, ei)).abs().max())
