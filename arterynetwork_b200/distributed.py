"""z-slab multi-GPU driver: one process per GPU, ``torch.distributed`` for the plumbing.

The volume is cut into contiguous slabs along z, the slowest axis (SURVEY.md
section 8(e)).  Each rank owns one slab plus ``HALO`` = 2 halo planes per side
and runs the same kernels as the single-GPU path; between them the host
interleaves the only two exchanges the algorithm has:

* after ``cancel`` -- the two boundary planes of the executed-flip bit-plane go
  to each neighbour, which applies them to its halo copy of the segmented plane
  (the next sweep classifies own planes +-1, i.e. needs state at distance 2);
* after ``flip``   -- one SUM all-reduce of the int64 statistics vector
  ``[hist_in[L], hist_out[L], n_in, n_out, n_excluded, n_flips, ...]``; every
  rank then derives the same decision table.  When the input holds label 4 the
  cancelled-flip and excluded planes' boundary planes are exchanged as well.

Intensities never move after upload (each rank uploads its extended slab).
Integer statistics make the result independent of the number of ranks: labels,
iteration count and trace are bit-identical to the single-GPU run.

The driver is engine-agnostic: ``tests/`` drive it on CPU with ``gloo`` and a
NumPy slab engine (world_size 2); on GPUs it drives ``VRGEngine`` over NCCL.
"""
from __future__ import annotations

import numpy as np

from . import _native as nat

HALO = nat.HALO


def slab_bounds(Z: int, world: int):
    """Balanced contiguous z-slabs: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(Z, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    if min(b[i + 1] - b[i] for i in range(world)) < HALO:
        raise ValueError("each slab needs at least %d planes (Z=%d over %d ranks)" % (HALO, Z, world))
    return b


class _CudaBlob:
    """Minimal ``__cuda_array_interface__`` carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class GpuSlabEngine:
    """Adapter: VRGEngine + torch views of the buffers the driver exchanges."""

    def __init__(self, eng, device):
        import torch
        self.eng = eng
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.own_planes = eng.own_planes

    def bind_buffers(self):
        torch, eng = self.torch, self.eng
        _, wpp, nzl = eng.plane_geometry()

        def view(which, shape, typestr):
            ptr, _ = eng.buffer(which)
            return torch.as_tensor(_CudaBlob(ptr, shape, typestr), device=self.device)
        self.seg = view(nat.BUF_SEG, (nzl, wpp), "<i4")
        self.excl = view(nat.BUF_EXCL, (nzl, wpp), "<i4")
        self.flips = view(nat.BUF_FLIPS, (nzl, wpp), "<i4")
        self.cancelled = view(nat.BUF_CANCELLED, (nzl, wpp), "<i4")
        n = eng.buffer(nat.BUF_LOCAL_STATS)[1] // 8
        self.local_stats = view(nat.BUF_LOCAL_STATS, (n,), "<i8")
        self.global_stats = view(nat.BUF_GLOBAL_STATS, (n,), "<i8")

    # the protocol the driver uses ------------------------------------------------------------
    def local_levels(self):
        return self.eng.scan_levels()

    def set_levels(self, levels):
        self.eng.set_levels(levels)
        self.eng.use_separate_global_stats()

    def init(self):
        self.eng.init()
        self.bind_buffers()

    def decide(self):
        self.eng.enqueue_decide()

    def cancel(self):
        self.eng.enqueue_cancel()

    def flip(self):
        self.eng.enqueue_flip()

    def absorb(self):
        self.eng.enqueue_absorb()

    def advance(self):
        self.eng.enqueue_advance()

    def poll(self):
        return self.eng.poll()

    def trace(self):
        return self.eng.trace()

    def labels(self):
        return self.eng.labels()


class DistributedVRG:
    """Runs the slab protocol over an initialised ``torch.distributed`` process group.

    ``use_graph`` (GPU engines only): after one eager batch (NCCL sets up its channels lazily), a batch of
    ``check_every`` iterations -- kernels, halo sends/receives and the all-reduce -- is captured once into a CUDA
    graph and replayed, so the host costs one launch per batch instead of ~20 calls per iteration.  Kernels read
    the device-side status word, so replays after the exit are no-ops.
    """

    def __init__(self, engine, rank, world, check_every=4, use_graph=False, transport="collective"):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.e, self.rank, self.world = engine, rank, world
        self.check_every = check_every
        self.iters_enqueued = 0
        self.has_excl = True
        self.use_graph = bool(use_graph)
        self._graph, self._graph_key = None, None
        # "collective": halo send/recv + all-reduce through torch.distributed (NCCL on GPUs, gloo on CPU), host-driven.
        # "p2p": the library's own kernels move halos and statistics over NVLink peer memory (vrg_p2p.cuh); the host
        #        only gathers the CUDA IPC handles once, then every rank calls vrg_init / vrg_run as on one GPU.
        self.transport = transport
        self._p2p_connected = False

    def _connect_p2p(self):
        if self._p2p_connected:
            return

        def gather(mine: bytes):
            out = [None] * self.world
            self.dist.all_gather_object(out, mine)
            return out
        self.e.eng.p2p_connect(self.rank, self.world, gather)
        self._p2p_connected = True

    # -- collectives ----------------------------------------------------------------------------
    def _exchange(self, t):
        """Send my boundary planes to the neighbours' halo planes of tensor ``t`` (planes x words)."""
        dist, n = self.dist, self.e.own_planes
        ops = []
        if self.rank > 0:  # lower neighbour: my first own planes -> its upper halo; its last planes -> my lower halo
            ops.append(dist.P2POp(dist.isend, t[HALO:2 * HALO], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, t[0:HALO], self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, t[n:n + HALO], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, t[HALO + n:2 * HALO + n], self.rank + 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _allreduce_stats(self):
        self.e.global_stats.copy_(self.e.local_stats)
        self.dist.all_reduce(self.e.global_stats, op=self.dist.ReduceOp.SUM)

    # -- phases ---------------------------------------------------------------------------------
    def prepare_levels(self):
        """Union of the slabs' intensity levels: every rank builds the same table domain.

        One fixed-size all-gather of float64 tensors ([count, levels..., padding]; 4096 slots, or 65537 when some slab
        holds more levels) instead of a pickled object gather, which costs milliseconds per call on NCCL."""
        torch, dist = self.torch, self.dist
        mine = np.asarray(self.e.local_levels(), dtype=np.float64)
        dev = getattr(self.e, "device", "cpu") if dist.get_backend() == "nccl" else "cpu"
        cap = 4096
        while True:
            buf = torch.zeros(cap, dtype=torch.float64)
            n = min(len(mine), cap - 1)
            buf[0] = float(len(mine))
            buf[1:1 + n] = torch.from_numpy(np.ascontiguousarray(mine[:n]))
            buf = buf.to(dev)
            parts = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(self.world)]
            dist.all_gather(parts, buf)
            g = torch.stack(parts).cpu().numpy()
            counts = g[:, 0].astype(np.int64)
            if int(counts.max()) <= cap - 1:
                break
            cap = nat.MAX_LEVELS + 1  # some slab has more than 4095 levels: once more with room for the maximum
        levels = np.unique(np.concatenate([g[r, 1:1 + counts[r]] for r in range(self.world)]))
        self.e.set_levels(levels)
        return levels

    def init(self):
        if self.transport == "p2p":
            self._connect_p2p()
            # no barrier: the ranks left prepare_levels (a collective) together, and the init exchange below waits for
            # its peers on the device (sequence-numbered flags, 20 s timeout)
            self.e.eng.init()    # seeds, bands, histograms + device-side halo / statistics exchange and the error checks
            self.init_row = tuple(int(x) for x in self.e.eng.trace()[0])
            return
        self.e.init()
        self._exchange(self.e.excl)  # the 4->3 absorption around seeds (VRG:137) is computed for own planes +-1 only
        self._allreduce_stats()
        g = self.e.global_stats.cpu().numpy()
        base = len(g) - nat.ST_EXTRA
        if g[base + nat.ST_BAD_LABEL]:
            raise ValueError("vrg_b200: initial valueMap may only hold labels 0 (seed), 3 (outside) and 4 (excluded)")
        if g[base + nat.ST_N_IN] == 0:
            raise ValueError("vrg_b200: no seed voxel (label 0) in valueMap")
        if g[base + nat.ST_N_BAND] == 0:
            raise ValueError("vrg_b200: seed has no boundary: every voxel is inside")
        self.has_excl = bool(g[base + nat.ST_N_EXCL] > 0)
        self.init_row = (-1, int(g[base + nat.ST_N_IN]), int(g[base + nat.ST_N_OUT]))
        self.iters_enqueued = 0

    def iterate_once(self):
        e = self.e
        e.decide()
        e.cancel()
        self._exchange(e.flips)
        if self.has_excl:
            self._exchange(e.cancelled)
            e.absorb()
        e.flip()
        if self.has_excl:
            self._exchange(e.excl)
        self._allreduce_stats()
        e.advance()
        self.iters_enqueued += 1

    def _batch(self):
        for _ in range(self.check_every):
            self.iterate_once()

    def _graph_for_current_buffers(self):
        e, torch = self.e, self.torch
        key = (self.has_excl, self.check_every, e.eng.params_signature()) + tuple(int(t.data_ptr()) for t in (
            e.seg, e.excl, e.flips, e.cancelled, e.local_stats, e.global_stats))
        if self._graph is None or key != self._graph_key:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream(), capture_error_mode="thread_local"):
                self._batch()
            self._graph, self._graph_key = g, key
        return self._graph

    def run(self):
        if self.transport == "p2p":
            return self.e.eng.run()  # the whole loop, exchanges included, is enqueued by the library
        first = True
        while True:
            if self.use_graph and not first:
                self._graph_for_current_buffers().replay()
                self.iters_enqueued += self.check_every
            else:
                self._batch()
            first = False
            res = self.e.poll()
            if res["exit_reason"] != nat.EXIT_RUNNING:
                # every rank left at the same update; one more all-reduce folds in the counters that are written after the
                # loop's last exchange (the order-dependence counters of the last applied update)
                self._allreduce_stats()
                return self.e.poll()

    def trace(self):
        t = np.array(self.e.trace(), copy=True)
        t[0] = self.init_row
        return t
