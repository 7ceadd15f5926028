"""CUDA-graph execution of the training step / the inference forward.

The reference's training step (``scripts/train_gnn.py:50-60``) is ~80 kernel launches here; issued one by one from
Python they cost ~1.8 ms of host time per step - as much as half of the GPU time on a B200.  ``GraphedTrainStep``
captures forward + loss + backward + AdamW for one batch shape into a CUDA graph (static input buffers, private memory
pool) and replays it: per step the host only copies the batch into the static buffers and launches the graph.

* Shapes: one graph per (num_nodes, num_edges); the first ``eager_steps`` calls of a new shape run eagerly (they are
  ordinary training steps and also warm every lazily initialised state up), then the shape is captured.
* The CSR build of ``edge_index`` is part of the graph (it depends on the batch).
* Dropout: seeds live in a device tensor that the graph bumps before the forward, so every replay draws new masks.
* Optimizer: must be created with ``capturable=True`` (step counter on the device).
* Data parallel: pass the ``MeshDataParallel`` wrapper as ``ddp``; the masked-count all-reduce, the bucketed gradient
  all-reduces (launched from the autograd hooks while the rest of the backward is still being recorded) and the wait
  before the optimizer are captured with the step, so a replay costs one launch per rank, NCCL included.
"""
from __future__ import annotations

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, eager_steps: int = 3, max_graphs: int = 8, ddp=None):
        self.model, self.criterion, self.optimizer, self.ddp = model, criterion, optimizer, ddp
        self.eager_steps, self.max_graphs = eager_steps, max_graphs
        self._seen: dict = {}
        self._graphs: dict = {}
        self.kernels_per_replay = 0
        dev = next(model.parameters()).device
        if model.device_seeds is None:
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            g = torch.Generator().manual_seed((torch.initial_seed() + 0x9E3779B9 * rank) & 0x7FFFFFFFFFFFFFFF)  # per-rank masks
            model.device_seeds = torch.randint(0, 2**62, (model.num_layers + 1,), generator=g, dtype=torch.int64).to(dev)
        self._pool = None

    def _step(self, x, edge_index, y, mask):
        """One training step (``train_gnn.py:50-60``); runs eagerly or under capture."""
        self.model.device_seeds.add_(0x9E3779B97F4A7C15 & 0x3FFFFFFFFFFFFFFF)
        if self.ddp is not None:
            self.ddp.zero_grad()  # gradients are views of one flat buffer: zero in place (also inside the graph)
        loss = self.criterion(self.model(x, edge_index, None), y, mask)
        if self.ddp is not None:
            self.ddp.scale_loss(loss, self.criterion.last_count).backward()
            self.ddp.finish()
        else:
            loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _eager(self, x, edge_index, y, mask):
        if self.ddp is None:
            self.optimizer.zero_grad(set_to_none=True)
        return self._step(x, edge_index, y, mask)

    def __call__(self, x, edge_index, y, mask, graph=None):
        """``graph``: the batch's prebuilt ``ops.CSRGraph`` with its transpose (``DataLoader(build_graph=True)`` builds it on the
        copy stream under the previous step): the captured step then starts at the encoder - the CSR arrays are copied into
        static buffers like ``x`` - instead of rebuilding the CSR from ``edge_index`` inside the graph."""
        if graph is not None and graph.rowptr_t is not None:
            return self._call_prebuilt(x, graph, y, mask)
        key = (int(x.shape[0]), int(edge_index.shape[1]), x.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            n = self._seen.get(key, 0)
            if n < self.eager_steps or len(self._graphs) >= self.max_graphs:
                self._seen[key] = n + 1
                return self._eager(x, edge_index, y, mask)
            entry = self._capture(x, edge_index, y, mask)
            self._graphs[key] = entry
        g, sx, se, sy, sm, sloss = entry
        sx.copy_(x, non_blocking=True)
        se.copy_(edge_index, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        sm.copy_(mask, non_blocking=True)
        g.replay()
        ops.LAUNCH_COUNTER["kernels"] += self.kernels_per_replay
        return sloss

    def _call_prebuilt(self, x, graph, y, mask):
        key = (int(x.shape[0]), int(graph.num_edges), x.dtype, "csr")
        entry = self._graphs.get(key)
        if entry is None:
            n = self._seen.get(key, 0)
            if n < self.eager_steps or len(self._graphs) >= self.max_graphs:
                self._seen[key] = n + 1
                return self._eager(x, graph, y, mask)
            # static copies of the batch and of its CSR (forward and transposed); the captured step reads only these
            sx, sy, sm = (torch.empty_like(t) for t in (x, y, mask))
            sg = ops.CSRGraph(None, graph.num_nodes, graph.num_edges, torch.empty_like(graph.rowptr), torch.empty_like(graph.col),
                              torch.empty_like(graph.inv_deg), None, None)
            sg.rowptr_t, sg.col_t = torch.empty_like(graph.rowptr_t), torch.empty_like(graph.col_t)
            for d, s_ in ((sx, x), (sy, y), (sm, mask), (sg.rowptr, graph.rowptr), (sg.col, graph.col), (sg.inv_deg, graph.inv_deg),
                          (sg.rowptr_t, graph.rowptr_t), (sg.col_t, graph.col_t)):
                d.copy_(s_)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            if self.ddp is None:
                self.optimizer.zero_grad(set_to_none=True)
            k0 = ops.LAUNCH_COUNTER["kernels"]
            with torch.cuda.graph(g, pool=self._pool, capture_error_mode="thread_local"):
                sloss = self._step(sx, sg, sy, sm)
            kernels = ops.LAUNCH_COUNTER["kernels"] - k0
            if self._pool is None:
                self._pool = g.pool()
            entry = (g, sx, sy, sm, sg, sloss, kernels)
            self._graphs[key] = entry
        g, sx, sy, sm, sg, sloss, kernels = entry
        sx.copy_(x, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        sm.copy_(mask, non_blocking=True)
        sg.rowptr.copy_(graph.rowptr, non_blocking=True)
        sg.col.copy_(graph.col, non_blocking=True)
        sg.inv_deg.copy_(graph.inv_deg, non_blocking=True)
        sg.rowptr_t.copy_(graph.rowptr_t, non_blocking=True)
        sg.col_t.copy_(graph.col_t, non_blocking=True)
        g.replay()
        ops.LAUNCH_COUNTER["kernels"] += kernels
        return sloss

    def capture_resident(self, x, edge_index, y, mask):
        """Capture the step ON these (device-resident, never freed) tensors: no staging copies, and a CSR that is already
        cached for ``edge_index`` stays outside the graph.  Returns ``(replay, loss)``: ``replay()`` runs one training
        step on this batch, ``loss`` is the static tensor it fills.  Used for pre-staged data sets (the whole
        data set fits in HBM) where a step should cost one launch and no copies."""
        # the CSR (and its transpose) is built - or fetched from the cache - BEFORE the capture and handed to the model as an
        # object: the CUDA graph holds raw pointers into it, so it must stay alive with the replay closure whatever the
        # LRU cache of ops.get_graph evicts later
        csr = edge_index if isinstance(edge_index, ops.CSRGraph) else ops.get_graph(edge_index, int(x.shape[0]))
        csr.transpose()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        if self.ddp is None:
            self.optimizer.zero_grad(set_to_none=True)
        k0 = ops.LAUNCH_COUNTER["kernels"]
        with torch.cuda.graph(g, pool=self._pool, capture_error_mode="thread_local"):
            sloss = self._step(x, csr, y, mask)
        kernels = ops.LAUNCH_COUNTER["kernels"] - k0
        if self._pool is None:
            self._pool = g.pool()
        keep = (x, edge_index, y, mask, csr)  # the graph holds raw pointers to them

        def replay(_g=g, _keep=keep, _k=kernels):
            _g.replay()
            ops.LAUNCH_COUNTER["kernels"] += _k

        return replay, sloss

    def _capture(self, x, edge_index, y, mask):
        sx, se, sy, sm = (torch.empty_like(t) for t in (x, edge_index, y, mask))
        sx.copy_(x); se.copy_(edge_index); sy.copy_(y); sm.copy_(mask)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        if self.ddp is None:
            self.optimizer.zero_grad(set_to_none=True)
        k0 = ops.LAUNCH_COUNTER["kernels"]
        # thread_local: the NCCL watchdog thread may poll events while this thread captures
        with torch.cuda.graph(g, pool=self._pool, capture_error_mode="thread_local"):
            sloss = self._step(sx, se, sy, sm)
        self.kernels_per_replay = ops.LAUNCH_COUNTER["kernels"] - k0
        if self._pool is None:
            self._pool = g.pool()
        # the capture itself did not run anything: the caller's replay() performs this step
        return g, sx, se, sy, sm, sloss


class GraphedForward:
    """Inference forward (``model.eval()``, no grad) replayed from a CUDA graph per (num_nodes, num_edges)."""

    def __init__(self, model, max_graphs: int = 8):
        self.model, self.max_graphs = model.eval(), max_graphs
        self._graphs: dict = {}

    @torch.no_grad()
    def __call__(self, x, edge_index):
        key = (int(x.shape[0]), int(edge_index.shape[1]), x.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= self.max_graphs:
                return self.model(x, edge_index)
            # warm-up (lazy initialisation outside the capture) on the CALLER's tensors: the static edge_index buffer must
            # reach the capture without a cached CSR, so that the CSR build is part of the graph and every replay rebuilds
            # it from the edge_index of that call (a CSR cached for the buffer would be silently reused for every later
            # graph of the same shape)
            self.model(x, edge_index)
            sx, se = torch.empty_like(x), torch.empty_like(edge_index)
            sx.copy_(x); se.copy_(edge_index)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.model(sx, se)
            entry = (g, sx, se, out)
            self._graphs[key] = entry
        g, sx, se, out = entry
        sx.copy_(x, non_blocking=True)
        se.copy_(edge_index, non_blocking=True)
        g.replay()
        return out

    @torch.no_grad()
    def capture_resident(self, x, graph):
        """Forward captured ON the caller's device-resident ``x`` with the batch's prebuilt ``ops.CSRGraph`` (no staging copies, no
        CSR build in the graph).  Returns ``(replay, out)``: ``replay()`` re-runs the forward into the static ``out``."""
        self.model(x, graph)  # lazy initialisation outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        k0 = ops.LAUNCH_COUNTER["kernels"]
        with torch.cuda.graph(g):
            out = self.model(x, graph)
        kernels = ops.LAUNCH_COUNTER["kernels"] - k0
        keep = (x, graph)  # the graph holds raw pointers into these

        def replay(_g=g, _keep=keep, _k=kernels):
            _g.replay()
            ops.LAUNCH_COUNTER["kernels"] += _k

        return replay, out
