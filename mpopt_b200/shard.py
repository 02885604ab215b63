"""Segment sharding over GPUs and the per-evaluation all-gather of the shards' g / Jacobian blocks.

The reference has no parallelism of any kind; the axis the *problem* offers is that, given the replicated
decision vector, every row owned by a segment is independent.  Rank r evaluates the
contiguous segment range ``partition(...)[r]`` with the same kernels, writing its rows straight into a
full-size ``g`` / ``values`` buffer at their final CSR positions; because rows are state-major, a shard's
output is a handful of contiguous runs (``Layout.shard_runs``), not one block.

``Gatherer`` makes every rank's buffers complete with ONE collective per evaluation:

* ``inplace``  (uniform degrees, K divisible by the world size): the runs of all ranks tile each block with
  equal sizes once rank 0's extra first row (global node 0) is set aside, so each block is an in-place
  ``all_gather_into_tensor`` on a view of the final buffer; the calls are coalesced into one NCCL group.
  Row 0 is recomputed locally by every rank (a one-segment plan), which is cheaper than broadcasting it.
* ``packed``   (anything else): runs are packed into one send buffer, gathered, and scattered back.

Works with any ``torch.distributed`` backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def partition(poly_orders, world):
    """Contiguous segment ranges [(begin, end)] * world, balanced by the Jacobian work p(p+1) per segment."""
    po = np.asarray(poly_orders, dtype=np.int64)
    K = len(po)
    if world > K:
        raise ValueError("more ranks than segments")
    if len(set(po.tolist())) == 1 and K % world == 0:
        step = K // world
        return [(r * step, (r + 1) * step) for r in range(world)]
    cost = np.cumsum(po * (po + 1), dtype=np.float64)
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(cost, cost[-1] * r / world)) + 1
        k = min(max(k, cuts[-1] + 1), K - (world - r))
        cuts.append(k)
    cuts.append(K)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class GatherPlan:
    """Pure index logic of the gather (testable without a GPU)."""

    def __init__(self, layout, part):
        self.world = len(part)
        self.layout = layout
        self.runs = {kind: [layout.shard_runs(kind, kb, ke) for kb, ke in part] for kind in (0, 1)}
        self.mode = "inplace" if self._tiles() else "packed"

    def _tiles(self):
        layout = self.layout
        """True if, apart from a leading piece of rank 0, corresponding runs of consecutive ranks are adjacent and
        of equal size (then each block is an in-place all-gather)."""
        self.blocks = {0: [], 1: []}  # (kind) -> [(start, count_per_rank)]
        for kind in (0, 1):
            rr = self.runs[kind]
            n_common = min(len(r) for r in rr)
            # tail-only runs (terminal rows, events) exist on the last rank only: broadcast-free if absent
            if any(len(r) != n_common for r in rr[:-1]):
                return False
            self.extra = getattr(self, "extra", {})
            self.extra[kind] = rr[-1][n_common:]
            node0 = layout.node0_counts(kind)
            if len(node0) < n_common:
                return False
            for i in range(n_common):
                cnt = rr[-1][i][1]
                if any(r[i][1] != cnt for r in rr[1:]):
                    return False
                lead = rr[0][i][1] - cnt
                if lead != node0[i]:  # rank 0 may only exceed the others by global node 0's row
                    return False
                start = rr[0][i][0] + lead
                for r in range(self.world):
                    if rr[r][i][0] + (lead if r == 0 else 0) != start + r * cnt:
                        return False
                self.blocks[kind].append((start, cnt, rr[0][i][0], lead))
        return True


class Gatherer:
    def __init__(self, layout, part, dist, rank, device=None, row0_eval=None):
        """``row0_eval(g, vals)``: evaluates segment 0 locally into the buffers (ranks != 0, inplace mode)."""
        self.plan = GatherPlan(layout, part)
        self.dist, self.rank, self.world, self.device = dist, rank, len(part), device
        self.row0_eval = row0_eval
        self.mode = self.plan.mode
        if self.mode == "inplace" and any(self.plan.extra[k] for k in (0, 1)):
            # terminal / event rows live on the last rank only: tiny, sent with a broadcast
            self.tail_runs = {k: self.plan.extra[k] for k in (0, 1)}
        else:
            self.tail_runs = {0: [], 1: []}
        self._send = self._recv = None

    # ------------------------------------------------------------------ in-place mode
    def _inplace(self, g, vals):
        dist = self.dist
        if self.rank != 0 and self.row0_eval is not None:
            self.row0_eval(g, vals)  # global node 0's rows, recomputed locally
        bufs = {0: g, 1: vals}
        ops = [(bufs[k][start:start + self.world * cnt], bufs[k][start + self.rank * cnt:start + (self.rank + 1) * cnt])
               for k in (0, 1) for (start, cnt, _, _) in self.plan.blocks[k]]
        cm = getattr(dist, "_coalescing_manager", None)
        if cm is not None and dist.get_backend() == "nccl":
            with cm(device=self.device):
                for out, inp in ops:
                    dist.all_gather_into_tensor(out, inp)
        else:
            for out, inp in ops:
                dist.all_gather_into_tensor(out, inp.clone())
        for k in (0, 1):
            for off, cnt in self.tail_runs[k]:
                dist.broadcast(bufs[k][off:off + cnt], src=self.world - 1)

    # ------------------------------------------------------------------ packed mode
    def _packed(self, g, vals):
        import torch

        dist = self.dist
        runs = self.plan.runs
        lens = [sum(c for _, c in runs[0][r]) + sum(c for _, c in runs[1][r]) for r in range(self.world)]
        mx = max(lens)
        if self._send is None:
            self._send = torch.empty(mx, dtype=g.dtype, device=g.device)
            self._recv = torch.empty(mx * self.world, dtype=g.dtype, device=g.device)
        mine = [g[o:o + c] for o, c in runs[0][self.rank]] + [vals[o:o + c] for o, c in runs[1][self.rank]]
        torch.cat(mine, out=self._send[:lens[self.rank]])
        dist.all_gather_into_tensor(self._recv, self._send)
        dst, src = [], []
        for r in range(self.world):
            if r == self.rank:
                continue
            pos = r * mx
            for buf, kind in ((g, 0), (vals, 1)):
                for o, c in runs[kind][r]:
                    dst.append(buf[o:o + c])
                    src.append(self._recv[pos:pos + c])
                    pos += c
        torch._foreach_copy_(dst, src)

    def all_gather(self, g, vals):
        """Complete ``g`` and ``vals`` (full-size tensors holding this rank's shard) on every rank."""
        if self.world == 1:
            return
        if self.mode == "inplace":
            self._inplace(g, vals)
        else:
            self._packed(g, vals)


class ObjectiveGatherer:
    """The objective side of the per-evaluation exchange: ``J`` partials and ``grad_f`` shards
    (BASELINE.json north_star: "allgather of the per-shard CSR blocks and objective partials"; the objective is
    ``J = Mayer + compW . vec(q)``, /root/reference/mpopt/mpopt.py:455).

    A shard's ``f + grad_f`` evaluation (``Transcription.f_grad_dev`` of a plan created with ``segments=``) leaves in
    a full-size gradient buffer (i) the node entries d J / d X(i, s), d J / d U(i, c) of the nodes it owns -- contiguous
    runs, kind 2 of ``Layout.shard_runs`` -- and (ii) its PARTIAL sums of the entries every node contributes to,
    d J / d t0, d J / d tf, d J / d a; the shard that holds the last segment adds the Mayer term, whose x0 entries land
    on node 0's columns (owned by rank 0's shard).  So per evaluation:

    * one all-reduce (sum) of ``[J, dJ/dt0, dJ/dtf, dJ/da (na), dJ/dx0 (nx)]`` per phase -- the partials;
    * one all-gather of the kind-2 runs (packed) -- the shards.
    """

    def __init__(self, layout, part, dist, rank):
        self.layout, self.part, self.dist, self.rank, self.world = layout, part, dist, rank, len(part)
        self.runs = [layout.shard_runs(2, kb, ke) for kb, ke in part]
        P, N, nx, nu, na = layout.P, layout.N, layout.nx, layout.nu, layout.na
        nvar = layout.n_z // P
        self.glob = []  # columns of the entries that are sums over shards, per phase: t0, tf, a.., x0..
        for ph in range(P):
            zo = ph * nvar
            self.glob += [zo + (nx + nu) * N + j for j in range(2 + na)] + [zo + s * N for s in range(nx)]
        self._send = self._recv = self._small = None

    def all_gather(self, f, grad):
        """``f``: 1-element tensor with this shard's partial objective, ``grad``: full-size gradient tensor holding
        this shard's entries.  On return both are complete on every rank."""
        import torch

        if self.world == 1:
            return
        dist, idx = self.dist, torch.as_tensor(self.glob, device=grad.device)
        # ---- partials: J and the entries summed over shards
        small = torch.empty(1 + len(self.glob), dtype=grad.dtype, device=grad.device)
        small[0] = f[0]
        small[1:] = grad[idx]
        nx = self.layout.nx
        per = len(self.glob) // self.layout.P
        own0, tail = self.part[self.rank][0] == 0, self.part[self.rank][1] == self.layout.K
        if not (own0 or tail):  # neither the running-cost part (owner of node 0) nor the Mayer part (last shard)
            for ph in range(self.layout.P):
                small[1 + ph * per + per - nx:1 + (ph + 1) * per] = 0.0
        dist.all_reduce(small)
        # ---- shards: node entries of the gradient
        lens = [sum(c for _, c in r) for r in self.runs]
        mx = max(lens)
        if self._send is None or self._send.device != grad.device:
            self._send = torch.zeros(mx, dtype=grad.dtype, device=grad.device)
            self._recv = torch.empty(mx * self.world, dtype=grad.dtype, device=grad.device)
        torch.cat([grad[o:o + c] for o, c in self.runs[self.rank]], out=self._send[:lens[self.rank]])
        dist.all_gather_into_tensor(self._recv, self._send)
        dst, src = [], []
        for r in range(self.world):
            if r == self.rank:
                continue
            pos = r * mx
            for o, c in self.runs[r]:
                dst.append(grad[o:o + c])
                src.append(self._recv[pos:pos + c])
                pos += c
        torch._foreach_copy_(dst, src)
        f[0] = small[0]
        grad[idx] = small[1:]


class _DevArray:
    """A device allocation seen through ``__cuda_array_interface__`` (lets torch view memory it did not allocate)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerBuffers:
    """Output buffers every rank can write into: the fused evaluation + all-gather of the north-star path.

    Each rank allocates ``n_sets`` pairs (g, values) with ``mpx_peer_alloc`` (cudaMalloc + CUDA IPC handle), the
    handles are exchanged through ``dist.all_gather_object`` and mapped with ``mpx_peer_open``; ``local(k)`` gives this
    rank's tensors of set ``k`` and ``peers(k)`` the device pointers of the same set on the other ranks, in the form
    ``Transcription.g_jac_dev_peers`` takes.  One process per GPU, all GPUs on one node (NVLink / NVSwitch)."""

    def __init__(self, n_g, nnz, dist, rank, device_index, n_sets=1):
        import ctypes as C

        import torch

        from . import _lib

        L = _lib.lib()
        self._L, self.dist, self.rank, self.world = L, dist, rank, dist.get_world_size()
        self._own, self._opened, self._tensors = [], [], []
        handles = []
        for _ in range(n_sets):
            pair = []
            for n in (n_g, nnz):
                ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
                _lib.check(L.mpx_peer_alloc(device_index, 8 * n, C.byref(ptr), h))
                self._own.append(ptr.value)
                pair.append((ptr.value, bytes(h)))
            handles.append([hb for _, hb in pair])
            self._tensors.append(tuple(torch.as_tensor(_DevArray(pv, n), device=torch.device("cuda", device_index))
                                       for (pv, _), n in zip(pair, (n_g, nnz))))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles)
        self._peer_ptrs = []  # [set] -> ([g pointers of the other ranks], [values pointers of the other ranks])
        for k in range(n_sets):
            pg, pv = [], []
            for r in range(self.world):
                if r == rank:
                    continue
                for which, dstl in ((0, pg), (1, pv)):
                    ptr = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(everyone[r][k][which])
                    _lib.check(L.mpx_peer_open(device_index, hb, C.byref(ptr)))
                    self._opened.append(ptr.value)
                    dstl.append(ptr.value)
            self._peer_ptrs.append((pg, pv))

    def local(self, k):
        return self._tensors[k]

    def peers(self, k):
        return self._peer_ptrs[k]

    def close(self):
        import torch

        torch.cuda.synchronize()
        self.dist.barrier()  # nobody is still writing into anybody's buffers
        for p in self._opened:
            self._L.mpx_peer_close(p)
        self._tensors = []
        self.dist.barrier()
        for p in self._own:
            self._L.mpx_peer_free(p)
        self._opened, self._own = [], []
