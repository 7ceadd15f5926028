"""Partitioned execution of ONE mesh across the ranks of a process group (inference).

The reference caps meshes at 500k nodes (``meshing.py:597-599``) and runs one mesh on one device
(``inference_gnn.py:224-328``); this module is the path for a mesh that should not - or cannot - live on one
GPU: a 1-D partition of the nodes with a halo exchange per SAGE layer.

* Nodes are relabelled once in k-d order (``ops.locality_order``), so a contiguous range of rows is a spatially
  compact piece of the mesh and its halo (the remote sources of its edges) is thin; rank ``r`` owns the rows
  ``bounds[r] .. bounds[r+1]`` (multiples of 128).
* Per SAGE layer (``model.py:89-95``): every rank sends the rows its peers' halos name (one
  ``all_to_all_single`` over NCCL/NVLink - the path's only exchange step), appends the received rows behind
  its own, aggregates over its LOCAL CSR (destinations = own rows, sources = own + halo rows) and runs the
  fused linear on its own rows.  Encoder, decoder and the epilogue are row-local.
* The arithmetic is the single-GPU path's: the same ``dfw_sage_aggregate`` (or, for bf16 rows on pieces of >= 131k rows,
  ``dfw_sage_aggregate_tc`` with a plan of the local graph) / ``dfw_linear_fwd`` launches on the same rows; a row's remote neighbours are numbered behind its local ones, so its fp32 neighbour sum is taken in a
  different order - the result equals the unpartitioned forward up to fp32 rounding
  (``tests/test_partition_nccl.py``: 2 ranks vs 1 GPU, 1e-5).

``build_partition`` is plain torch index arithmetic (CPU-testable, ``tests/test_partition_logic.py``).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import ops


@dataclass
class MeshPartition:
    """Rank-local view of a 1-D node partition (all indices in the RELABELLED numbering)."""

    rank: int
    world: int
    bounds: torch.Tensor        # int64 [world+1] (host): rank p owns rows bounds[p] .. bounds[p+1]
    n_own: int
    halo_ids: torch.Tensor      # int64 [n_halo] ascending global ids of the remote sources (grouped by owner)
    local_edge_index: torch.Tensor  # int64 [2, E_local]: src in [0, n_own + n_halo), dst in [0, n_own)
    recv_counts: list           # rows received from every peer (its share of halo_ids)
    send_index: torch.Tensor | None = None  # int64: own-local rows to send, concatenated per peer
    send_counts: list | None = None


def partition_bounds(num_nodes: int, world: int, align: int = 128) -> torch.Tensor:
    """Equal contiguous row ranges, boundaries on multiples of ``align`` (the aggregation kernels' block size)."""
    per = (num_nodes + world - 1) // world
    per = (per + align - 1) // align * align
    b = [min(p * per, num_nodes) for p in range(world)] + [num_nodes]
    return torch.tensor(b, dtype=torch.int64)


def build_partition(edge_index_relabelled: torch.Tensor, bounds: torch.Tensor, rank: int) -> MeshPartition:
    """Local graph of ``rank``: the edges whose DESTINATION it owns; sources it does not own become halo rows, numbered
    behind the own rows in ascending global order (so they are grouped by owner: one contiguous slice per peer)."""
    ei = edge_index_relabelled
    world = int(bounds.numel()) - 1
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    n_own = hi - lo
    mine = (ei[1] >= lo) & (ei[1] < hi)
    src, dst = ei[0][mine], ei[1][mine] - lo
    own = (src >= lo) & (src < hi)
    halo_ids = torch.unique(src[~own])  # sorted ascending
    src_local = torch.where(own, src - lo, n_own + torch.searchsorted(halo_ids, src))
    owner = torch.searchsorted(bounds.to(halo_ids.device), halo_ids, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world).tolist() if halo_ids.numel() else [0] * world
    return MeshPartition(rank, world, bounds, n_own, halo_ids, torch.stack([src_local, dst]).contiguous(), [int(c) for c in recv_counts])


def send_lists_from_halos(parts: list) -> None:
    """Fill ``send_index`` / ``send_counts`` of every partition from the others' halo lists (single-process helper for tests;
    ``PartitionedMeshInference`` does the same with two all-to-alls)."""
    world = len(parts)
    for p in parts:
        lo, hi = int(p.bounds[p.rank]), int(p.bounds[p.rank + 1])
        chunks = []
        for q in parts:
            h = q.halo_ids
            chunks.append(h[(h >= lo) & (h < hi)] - lo)  # what q wants from p, in q's halo order
        p.send_counts = [int(c.numel()) for c in chunks]
        p.send_index = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int64)
    assert all(parts[q].recv_counts[p] == parts[p].send_counts[q] for p in range(world) for q in range(world))


class PartitionedMeshInference:
    """``PartitionedMeshInference(model, x, edge_index)(…)``: the model's eval forward on one mesh, rows split over the ranks
    of ``process_group`` (NCCL).  Every rank passes the same ``x [N, in_channels]`` and ``edge_index [2, E]`` (on its device)
    for the one-time partitioning; afterwards it holds only its rows.  ``forward()`` returns this rank's predictions (rows in
    relabelled order); ``gather()`` assembles the full ``[N, out]`` prediction in the caller's node order on every rank."""

    def __init__(self, model, x: torch.Tensor, edge_index: torch.Tensor, process_group=None, reorder: bool = True):
        ops._require_cuda(x, "x")
        self.model = model.eval()
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.rank = dist.get_rank(process_group)
        n = int(x.shape[0])
        dev = x.device
        self.num_nodes = n
        if reorder and n > 128:
            new_id = ops.locality_order(x[:, :3], edge_index)
            dist.broadcast(new_id, src=0, group=process_group)  # one numbering for everybody (sort ties are device-deterministic, but be explicit)
        else:
            new_id = torch.arange(n, device=dev)
        self.new_id = new_id
        bounds = partition_bounds(n, self.world)
        self.part = build_partition(new_id[edge_index], bounds, self.rank)
        lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        order = torch.empty_like(new_id)
        order[new_id] = torch.arange(n, device=dev)
        self.x_own = x.index_select(0, order[lo:hi]).contiguous()
        # who wants which of my rows: counts, then the index lists themselves
        recv_counts = torch.tensor(self.part.recv_counts, dtype=torch.int64, device=dev)
        send_counts = torch.empty_like(recv_counts)
        dist.all_to_all_single(send_counts, recv_counts, group=process_group)
        self.part.send_counts = [int(c) for c in send_counts.tolist()]
        want = self.part.halo_ids.contiguous()  # grouped by owner, ascending
        asked = torch.empty(int(send_counts.sum()), dtype=torch.int64, device=dev)
        dist.all_to_all_single(asked, want, output_split_sizes=self.part.send_counts, input_split_sizes=self.part.recv_counts, group=process_group)
        self.part.send_index = (asked - lo).contiguous()
        assert self.part.send_index.numel() == 0 or (int(self.part.send_index.min()) >= 0 and int(self.part.send_index.max()) < hi - lo)
        # local CSR: destinations = own rows; sources index [own rows | halo rows]
        n_ext = self.part.n_own + int(self.part.halo_ids.numel())
        rowptr, col, _, inv_deg, _ = ops.csr_build_raw(self.part.local_edge_index, max(n_ext, 1), want_perm=False)
        self.rowptr, self.col, self.inv_deg = rowptr[: self.part.n_own + 1].contiguous(), col, inv_deg[: self.part.n_own].contiguous()
        self.halo_rows = int(self.part.halo_ids.numel())
        # bf16 rows on a large enough piece: the tensor-core block aggregation on the LOCAL graph (destinations = own rows, which are
        # also the first rows of [own | halo], so a block's own rows are plain tiles and every remote source is one more halo row)
        self.plan = None
        n_own, e_loc = self.part.n_own, int(self.col.numel())
        if (model.compute_dtype == torch.bfloat16 and model.hidden_channels in ops.TC_AGG_WIDTHS and n_own >= ops.TC_AGG_MIN_NODES
                and e_loc > 0):
            plan = ops.build_agg_plan(self.rowptr, self.col, n_own)
            if plan.check() and e_loc / max(n_own, 1) >= ops.TC_AGG_MIN_REUSE * plan.staged_rows_per_row:
                self.plan = plan

    @torch.no_grad()
    def forward(self) -> torch.Tensor:
        m = self.model
        cd = m.compute_dtype
        x = self.x_own
        enc0, enc2 = m.encoder[0], m.encoder[2]
        if x.dtype == torch.float32 and cd == torch.bfloat16 and m.in_channels <= 16 and enc0.out_features % 4 == 0:
            h, _, _, _ = ops.linear_fwd(x, ops._f32(enc0.weight.detach()), bias=ops._f32(enc0.bias.detach()), relu=True, out_bf16=True)  # as model.py
        else:
            h = x if (x.dtype == torch.float32 and cd == torch.bfloat16) else ops.cast(x, cd)
            h, _, _, _ = ops.linear_fwd(h, ops._w(enc0.weight, h.dtype), bias=ops._f32(enc0.bias.detach()), relu=True)
            h = ops.cast(h, cd)
        h, _, _, _ = ops.linear_fwd(h, ops._w(enc2.weight, cd), bias=ops._f32(enc2.bias.detach()), relu=True)
        H = h.shape[1]
        for conv, norm in zip(m.convs, m.norms):
            send = h.index_select(0, self.part.send_index)
            halo = torch.empty(self.halo_rows, H, dtype=h.dtype, device=h.device)
            dist.all_to_all_single(halo, send, output_split_sizes=self.part.recv_counts, input_split_sizes=self.part.send_counts, group=self.pg)
            h_ext = torch.cat([h, halo], dim=0)
            if self.plan is not None:
                agg = ops.aggregate_tc(self.plan, self.inv_deg, h_ext, int(self.col.numel()), n_rows=self.part.n_own)
            else:
                agg = ops.aggregate_rows(self.rowptr, self.col, self.inv_deg, h_ext, self.part.n_own)
            h, _, _, _ = ops.linear_fwd(agg, ops._w(conv.lin_l.weight, cd), h, ops._w(conv.lin_r.weight, cd), bias=ops._f32(conv.lin_l.bias.detach()),
                                        ln=(ops._f32(norm.weight.detach()), ops._f32(norm.bias.detach())), eps=float(norm.eps), relu=True, residual=h)
        dec0, dec3 = m.decoder[0], m.decoder[3]
        b0 = ops._f32(dec0.bias.detach())
        if m.out_channels == 1:  # Linear(H,64) -> ReLU -> Linear(64,1) in one launch: the 64 -> 1 projection is the epilogue's row dot
            _, _, _, rd = ops.linear_fwd(h, ops._w(dec0.weight, cd), bias=b0, relu=True, rowdot=(ops._f32(dec3.weight.detach()).reshape(-1), ops._f32(dec3.bias.detach()).reshape(-1)),
                                         want_out=False)
            return rd.unsqueeze(1)
        hid, _, _, _ = ops.linear_fwd(h, ops._w(dec0.weight, cd), bias=b0, relu=True)
        out, _, _, _ = ops.linear_fwd(hid, ops._w(dec3.weight, cd), bias=ops._f32(dec3.bias.detach()))
        return out.float()

    __call__ = forward

    @torch.no_grad()
    def gather(self, out_own: torch.Tensor) -> torch.Tensor:
        """All ranks' rows -> the full prediction in the CALLER's node order (on every rank)."""
        bounds = self.part.bounds.tolist()
        sizes = [bounds[p + 1] - bounds[p] for p in range(self.world)]
        cap = max(sizes)
        mine = torch.zeros(cap, out_own.shape[1], dtype=out_own.dtype, device=out_own.device)  # equal-sized pieces for all_gather
        mine[: out_own.shape[0]] = out_own
        pieces = [torch.empty_like(mine) for _ in sizes]
        dist.all_gather(pieces, mine, group=self.pg)
        return torch.cat([pc[:s] for pc, s in zip(pieces, sizes)], dim=0).index_select(0, self.new_id)
