"""PyG-free ``Data`` / ``Batch`` / ``DataLoader`` with the semantics the reference relies on.

The reference batches graphs with ``torch_geometric.loader.DataLoader`` (``scripts/train_gnn.py:30,150``):
a mini-batch is the disjoint union of its graphs - node tensors concatenated, ``edge_index``
shifted by the running node offset, plus ``batch`` (graph id per node), ``ptr`` and ``num_graphs``
(``train_gnn.py:51-60`` uses ``.to(device)``, ``.x``, ``.edge_index``, ``.batch``, ``.y``, ``.loss_mask``,
``.num_graphs``).  This module restates exactly that, and adds what a B200 box wants:

* pinned-memory collation + asynchronous H2D on a side stream with one batch of prefetch
  (``DataLoader(..., device=...)``), so the copy of step i+1 overlaps the compute of step i;
* mesh-level sharding for data-parallel training (``rank`` / ``world_size`` / ``set_epoch``), the
  ``DistributedSampler`` contract: every rank draws the same permutation and takes a strided slice.
"""
from __future__ import annotations

from typing import Any, Iterable, Sequence

import torch

_NODE_KEYS_DEFAULT = ("x", "y", "loss_mask", "pos", "disp", "stress_vm_raw")


class Data:
    """Attribute bag for one graph (stand-in for ``torch_geometric.data.Data``, ``dataset.py:266-277``)."""

    def __init__(self, **kwargs: Any):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to_dict(self) -> dict:
        return {k: getattr(self, k) for k in self.keys()}

    @property
    def num_nodes(self) -> int:
        x = getattr(self, "x", None)
        if x is not None:
            return int(x.shape[0])
        ei = getattr(self, "edge_index", None)
        return int(ei.max()) + 1 if ei is not None and ei.numel() else 0

    @property
    def num_edges(self) -> int:
        ei = getattr(self, "edge_index", None)
        return int(ei.shape[1]) if ei is not None else 0

    def _apply(self, fn):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = fn(v) if isinstance(v, torch.Tensor) else v
        return out

    def to(self, device, non_blocking: bool = False):
        return self._apply(lambda t: t.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def cuda(self, device=None, non_blocking: bool = False):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"), non_blocking)

    def cpu(self):
        return self.to("cpu")

    def __repr__(self) -> str:
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v!r}")
        return f"{self.__class__.__name__}({', '.join(parts)})"


class Batch(Data):
    """Disjoint union of graphs (PyG ``Batch.from_data_list`` semantics)."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data], pin: bool = False) -> "Batch":
        if len(data_list) == 0:
            raise ValueError("empty batch")
        sizes = [d.num_nodes for d in data_list]
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.int64)
        ptr[1:] = torch.tensor(sizes, dtype=torch.int64).cumsum(0)
        out = cls()
        first = data_list[0]
        for k in first.keys():
            v0 = getattr(first, k)
            vals = [getattr(d, k, None) for d in data_list]
            if k == "edge_index":
                dev = v0.device
                shifted = [v + int(ptr[i]) for i, v in enumerate(vals)]
                val = torch.cat(shifted, dim=1)
            elif isinstance(v0, torch.Tensor) and v0.dim() >= 1 and v0.shape[0] == first.num_nodes and k != "global_params" \
                    and k != "global_params_raw":
                val = torch.cat(vals, dim=0)
            elif isinstance(v0, torch.Tensor):
                val = torch.stack(vals, dim=0)
            else:
                val = list(vals)
            if pin and isinstance(val, torch.Tensor) and not val.is_cuda:
                val = val.pin_memory()
            setattr(out, k, val)
        dev = first.x.device if getattr(first, "x", None) is not None else "cpu"
        out.batch = torch.repeat_interleave(torch.arange(len(sizes), device=dev), torch.tensor(sizes, device=dev))
        out.ptr = ptr.to(dev)
        if pin and not out.batch.is_cuda:
            out.batch = out.batch.pin_memory()
        out.num_graphs = len(data_list)
        return out

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def __getattr__(self, name):
        # `batch` (graph id per node) is derived from `ptr` on first use: the model ignores it
        # (model.py:74), so the prefetching loader does not ship it over PCIe.
        if name == "batch" and "ptr" in self.__dict__:
            ptr = self.__dict__["ptr"]
            b = torch.repeat_interleave(torch.arange(ptr.numel() - 1, device=ptr.device), ptr[1:] - ptr[:-1])
            self.__dict__["batch"] = b
            return b
        raise AttributeError(name)


class DataLoader:
    """Mini-batches of graphs.  ``DataLoader(dataset, batch_size=4, shuffle=True)`` behaves like the
    PyG loader the reference uses (``train_gnn.py:150-152``).  Extras: ``device`` (prefetching H2D
    pipeline from pinned memory), ``rank``/``world_size`` (mesh-level data parallel sharding),
    ``drop_last``."""

    def __init__(self, dataset: Sequence[Data], batch_size: int = 1, shuffle: bool = False, drop_last: bool = False,
                 device=None, rank: int = 0, world_size: int = 1, seed: int | None = None, keys: Iterable[str] | None = None,
                 build_graph: bool = False):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = shuffle
        self.drop_last = drop_last
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world_size = int(rank), int(world_size)
        self.seed = seed
        self.epoch = 0
        self.keys = tuple(keys) if keys is not None else None
        # build_graph: the prefetch worker also builds the batch's CSR (+ its transpose) ON THE COPY STREAM, right behind the H2D
        # copies, and attaches it as ``batch.graph`` (registered for ``batch.edge_index``): the graph construction of batch i+1
        # then runs under the compute of batch i instead of at the head of its step.  Off by default: on the config-2 step it
        # measured no gain (2.74 vs 2.75 ms; the small build kernels only find room between the step's full-chip kernels) and the
        # per-batch allocations on a second stream make the caching allocator grow for the first epochs (tools/e2e_probe.py).
        self.build_graph = bool(build_graph)
        self._copy_stream = None
        self._pinned = None

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _indices(self) -> list[int]:
        n = len(self.dataset)
        if self.shuffle:
            if self.world_size > 1 or self.seed is not None:
                g = torch.Generator()
                g.manual_seed((self.seed if self.seed is not None else 0) + self.epoch)
                perm = torch.randperm(n, generator=g).tolist()
            else:
                perm = torch.randperm(n).tolist()  # global CPU generator, like torch's RandomSampler
        else:
            perm = list(range(n))
        if self.world_size > 1:
            total = (n + self.world_size - 1) // self.world_size * self.world_size
            perm = perm + perm[: total - n]  # pad by wrap-around (DistributedSampler contract)
            perm = perm[self.rank:total:self.world_size]
        return perm

    def __len__(self) -> int:
        n = len(self._indices())
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _select(self, d: Data) -> Data:
        if self.keys is None:
            return d
        return Data(**{k: getattr(d, k) for k in self.keys})

    def _host_batches(self):
        idx = self._indices()
        for s in range(0, len(idx), self.batch_size):
            chunk = idx[s:s + self.batch_size]
            if self.drop_last and len(chunk) < self.batch_size:
                break
            yield [self._select(self.dataset[i]) for i in chunk]

    def __iter__(self):
        if self.device is None or self.device.type != "cuda":
            for items in self._host_batches():
                yield Batch.from_data_list(items)
            return
        first = self.dataset[0] if len(self.dataset) else None
        if first is not None and getattr(first, "x", None) is not None and first.x.is_cuda:
            for items in self._host_batches():  # device-resident dataset: collate on the device
                yield Batch.from_data_list(items)
            return
        yield from self._prefetch_iter()

    # -- host -> device pipeline ---------------------------------------------------------------
    # The dataset's tensors are pinned once; a worker thread then assembles every batch ON THE DEVICE:
    # one asynchronous H2D copy per (mesh, tensor) straight into its slice of the batch tensor on a side
    # stream, plus an in-place offset add for edge_index.  No host-side concatenation (a 28 MB memcpy per
    # step on one core was slower than the whole GPU step).  One batch of prefetch: the copy of batch i+1
    # overlaps the compute of batch i; the consumer only waits on the copy's event.
    def pin_dataset(self, indices=None) -> None:
        """Page-lock the (host) graphs this loader will draw - ``indices`` (default: this rank's shard of the current
        epoch, ``_indices()``), never the whole dataset: under torchrun every rank would otherwise pin all N cases although
        it consumes N / world of them.  Done eagerly at the first iteration of an epoch: pinning is slow (milliseconds per
        mesh) and must not land in a step; graphs pinned in earlier epochs stay pinned."""
        if self._pinned is None:
            self._pinned = {}
        for i in (self._indices() if indices is None else indices):
            if i not in self._pinned:
                d = self._select(self.dataset[i])
                self._pinned[i] = d if (getattr(d, "x", None) is not None and d.x.is_pinned()) else d.pin_memory()

    def _pinned_item(self, i):
        if self._pinned is None or i not in self._pinned:
            self.pin_dataset([i])
        return self._pinned[i]

    def _device_collate(self, items, dev):
        first = items[0]
        sizes = [d.num_nodes for d in items]
        out = Batch()
        for k in first.keys():
            v0 = getattr(first, k)
            vals = [getattr(d, k) for d in items]
            if k == "edge_index":
                es = [int(v.shape[1]) for v in vals]
                buf = torch.empty((2, sum(es)), dtype=v0.dtype, device=dev)
                off_n = off_e = 0
                for v, n, e in zip(vals, sizes, es):
                    buf[0, off_e:off_e + e].copy_(v[0], non_blocking=True)
                    buf[1, off_e:off_e + e].copy_(v[1], non_blocking=True)
                    if off_n:
                        buf[:, off_e:off_e + e].add_(off_n)
                    off_n += n
                    off_e += e
                setattr(out, k, buf)
            elif isinstance(v0, torch.Tensor) and v0.dim() >= 1 and v0.shape[0] == first.num_nodes and k not in ("global_params", "global_params_raw"):
                buf = torch.empty((sum(sizes),) + tuple(v0.shape[1:]), dtype=v0.dtype, device=dev)
                off = 0
                for v, n in zip(vals, sizes):
                    buf[off:off + n].copy_(v, non_blocking=True)
                    off += n
                setattr(out, k, buf)
            elif isinstance(v0, torch.Tensor):
                buf = torch.empty((len(vals),) + tuple(v0.shape), dtype=v0.dtype, device=dev)
                for j, v in enumerate(vals):
                    buf[j].copy_(v, non_blocking=True)
                setattr(out, k, buf)
            else:
                setattr(out, k, list(vals))
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.int64)
        ptr[1:] = torch.tensor(sizes, dtype=torch.int64).cumsum(0)
        out.ptr = ptr.to(dev, non_blocking=False)
        out.num_graphs = len(items)
        return out

    def _index_batches(self):
        idx = self._indices()
        for s in range(0, len(idx), self.batch_size):
            chunk = idx[s:s + self.batch_size]
            if self.drop_last and len(chunk) < self.batch_size:
                break
            yield chunk

    def _prefetch_iter(self):
        import queue
        import threading

        dev = self.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        self.pin_dataset()
        q: "queue.Queue" = queue.Queue(maxsize=2)
        stop = threading.Event()

        def worker():
            try:
                torch.cuda.set_device(dev)
                for chunk in self._index_batches():
                    if stop.is_set():
                        return
                    items = [self._pinned_item(i) for i in chunk]
                    with torch.cuda.stream(cs):
                        out = self._device_collate(items, dev)
                        if self.build_graph and getattr(out, "edge_index", None) is not None:
                            from . import ops

                            g = ops.get_graph(out.edge_index, int(out.x.shape[0]))
                            g.transpose()
                            out.graph = g
                        ev = torch.cuda.Event()
                        ev.record(cs)
                    q.put((out, ev))
                q.put(None)
            except BaseException as e:  # surface worker failures in the consumer
                q.put(e)

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                batch, ev = item
                cur = torch.cuda.current_stream(dev)
                cur.wait_event(ev)
                for v in batch.__dict__.values():
                    if isinstance(v, torch.Tensor) and v.is_cuda:
                        v.record_stream(cur)
                g = batch.__dict__.get("graph")
                if g is not None:  # the CSR arrays were allocated on the copy stream too
                    for v in (g.rowptr, g.col, g.inv_deg, g.rowptr_t, g.col_t, g.status, g.status_t):
                        if isinstance(v, torch.Tensor) and v.is_cuda:
                            v.record_stream(cur)
                yield batch
        finally:
            stop.set()
            while th.is_alive():
                try:
                    q.get_nowait()
                except Exception:
                    pass
                th.join(timeout=0.05)


class _StagingSlot:
    """Reusable pinned host buffers for one in-flight batch."""

    def __init__(self):
        self.bufs: dict[str, torch.Tensor] = {}
        self.event = None

    def _buf(self, key, shape, dtype):
        need = 1
        for s_ in shape:
            need *= int(s_)
        b = self.bufs.get(key)
        if b is None or b.numel() < need or b.dtype != dtype:
            b = torch.empty(max(need, 1) * 5 // 4, dtype=dtype).pin_memory()
            self.bufs[key] = b
        return b[:need].view(shape)

    def collate(self, items):
        first = items[0]
        sizes = [d.num_nodes for d in items]
        out = {}
        for k in first.keys():
            v0 = getattr(first, k)
            vals = [getattr(d, k) for d in items]
            if k == "edge_index":
                e_tot = sum(int(v.shape[1]) for v in vals)
                buf = self._buf(k, (2, e_tot), v0.dtype)
                off_n = off_e = 0
                for v, n in zip(vals, sizes):
                    e = int(v.shape[1])
                    torch.add(v, off_n, out=buf[:, off_e:off_e + e])
                    off_n += n
                    off_e += e
                out[k] = buf
            elif isinstance(v0, torch.Tensor) and v0.dim() >= 1 and v0.shape[0] == first.num_nodes and k not in ("global_params", "global_params_raw"):
                shape = (sum(sizes),) + tuple(v0.shape[1:])
                buf = self._buf(k, shape, v0.dtype)
                torch.cat(vals, dim=0, out=buf)
                out[k] = buf
            elif isinstance(v0, torch.Tensor):
                buf = self._buf(k, (len(vals),) + tuple(v0.shape), v0.dtype)
                torch.stack(vals, dim=0, out=buf)
                out[k] = buf
            else:
                out[k] = list(vals)
        ptr = self._buf("ptr", (len(sizes) + 1,), torch.int64)
        ptr[0] = 0
        ptr[1:] = torch.tensor(sizes, dtype=torch.int64).cumsum(0)
        out["ptr"] = ptr
        return out
