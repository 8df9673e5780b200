"""TEST-ONLY NumPy slab engine: the same per-slab protocol as the CUDA kernels (halo planes,
local/global statistics vectors, decide / apply / absorb / advance), built on the oracle's rules,
so the multi-rank driver (arterynetwork_b200/distributed.py) can be exercised on CPU with gloo."""
import numpy as np
import torch

from arterynetwork_b200 import _native as nat
from oracle.vrg_oracle import A, dil3

HALO = nat.HALO


class NumpySlabEngine:
    def __init__(self, shape, z_begin, z_end, data, value_map, H=2.25, max_segment_size=10 ** 15, iter_max=200):
        self.shape = shape
        Z, Y, X = shape
        self.z_begin, self.z_end = z_begin, z_end
        self.own_planes = z_end - z_begin
        self.nzl = self.own_planes + 2 * HALO
        e0, e1 = max(0, z_begin - HALO), min(Z, z_end + HALO)
        self.vlo, self.vhi = e0 - (z_begin - HALO), e1 - (z_begin - HALO)
        self.data = np.zeros((self.nzl, Y, X))
        self.vm = np.full((self.nzl, Y, X), 3, dtype=np.uint8)
        self.data[self.vlo:self.vhi] = data[e0:e1]
        self.vm[self.vlo:self.vhi] = value_map[e0:e1]
        self.inb = np.zeros((self.nzl, Y, X), dtype=bool)
        self.inb[self.vlo:self.vhi] = True
        self.H, self.max_seg, self.iter_max = H, max_segment_size, iter_max
        self.own = slice(HALO, HALO + self.own_planes)
        self.own1 = slice(max(self.vlo, HALO - 1), min(self.vhi, HALO + self.own_planes + 1))

    # planes x voxels byte tensors stand in for the bit-planes
    def _planes(self):
        n = self.shape[1] * self.shape[2]
        self.seg, self.excl, self.flips, self.cancelled = (torch.zeros((self.nzl, n), dtype=torch.uint8) for _ in range(4))

    def _np(self, t):
        return t.numpy().reshape((self.nzl,) + tuple(self.shape[1:])).view(np.bool_)

    def local_levels(self):
        return np.unique(self.data[self.vlo:self.vhi])

    def set_levels(self, levels):
        self.levels = np.asarray(levels, dtype=np.float64)
        self.L = len(self.levels)
        self.idx = np.searchsorted(self.levels, self.data)
        self.idx[~self.inb] = 0
        diff = self.levels[:, None] - self.levels[None, :]
        self.kmat = A * np.exp(-0.5 * self.H * diff ** 2)
        self.local_stats = torch.zeros(2 * self.L + nat.ST_EXTRA, dtype=torch.int64)
        self.global_stats = torch.zeros_like(self.local_stats)

    def init(self):
        self._planes()
        S, E = self._np(self.seg), self._np(self.excl)
        S[:] = (self.vm <= 1) & self.inb  # a returned map can be fed back in: labels 0 / 1 are segmented
        E[:] = (self.vm == 4) & self.inb
        E[self.own1] &= ~dil3(S)[self.own1]  # like the kernel: absorbed around seeds on own planes +-1 only
        st = self.local_stats.numpy()
        st[:] = 0
        so, eo, io = S[self.own], E[self.own], self.idx[self.own]
        st[: self.L] = np.bincount(io[so], minlength=self.L)
        st[self.L: 2 * self.L] = np.bincount(io[~so & ~eo], minlength=self.L)
        b = 2 * self.L
        st[b + nat.ST_N_IN] = so.sum()
        st[b + nat.ST_N_OUT] = (~so & ~eo).sum()
        st[b + nat.ST_N_EXCL] = eo.sum()
        st[b + nat.ST_BAD_LABEL] = int((self.vm[self.vlo:self.vhi] > 4).any())
        nonseg = ~S & self.inb
        band = (S & dil3(nonseg)) | (nonseg & ~E & dil3(S))
        st[b + nat.ST_N_BAND] = band[self.own].sum()
        self.ctrl = {"status": nat.EXIT_RUNNING, "iter": 1, "applied": 0, "sweeps": 0, "apply": 1}
        self.trace_rows = [(-1, 0, 0)]

    def decide(self):
        if self.ctrl["status"] != nat.EXIT_RUNNING:
            return
        g = self.global_stats.numpy()
        b = 2 * self.L
        self.ctrl["apply"] = int(g[b + nat.ST_N_IN] < self.max_seg)
        self.local_stats[b + nat.ST_N_FLIPS] = 0
        with np.errstate(all="ignore"):
            pin = (g[: self.L].astype(np.float64) @ self.kmat) / g[b + nat.ST_N_IN]
            pout = (g[self.L: b].astype(np.float64) @ self.kmat) / g[b + nat.ST_N_OUT]
        d = pin >= pout
        S, E, F = self._np(self.seg), self._np(self.excl), self._np(self.flips)
        nonseg = ~S & self.inb
        band = (S & dil3(nonseg)) | (nonseg & ~E & dil3(S))
        F[self.own1] = (band & (d[self.idx] ^ S))[self.own1]  # flips on own planes +-1; outer halo planes keep stale data
        self.local_stats[b + nat.ST_N_FLIPS] = int(F[self.own].sum())

    def cancel(self):
        if self.ctrl["status"] != nat.EXIT_RUNNING or not self.ctrl["apply"]:
            return
        S, F, C = self._np(self.seg), self._np(self.flips), self._np(self.cancelled)
        Fv = np.zeros_like(F)
        Fv[self.own1] = F[self.own1]  # the cancel rule looks one plane out, never at the stale outer halo
        keep = S & ~Fv
        a0, r = Fv & ~S, Fv & S
        a = a0 & dil3(keep)
        F[self.own] = (r | a)[self.own]
        C[self.own] = (a0 & ~a)[self.own]
        st = self.local_stats.numpy()
        io = self.idx[self.own]
        ra, aa = r[self.own], a[self.own]
        dr, da = np.bincount(io[ra], minlength=self.L), np.bincount(io[aa], minlength=self.L)
        st[: self.L] += da - dr
        st[self.L: 2 * self.L] += dr - da
        b = 2 * self.L
        st[b + nat.ST_N_IN] += int(aa.sum()) - int(ra.sum())
        st[b + nat.ST_N_OUT] -= int(aa.sum()) - int(ra.sum())

    def absorb(self):
        if self.ctrl["status"] != nat.EXIT_RUNNING or not self.ctrl["apply"]:
            return
        F, C, E = self._np(self.flips), self._np(self.cancelled), self._np(self.excl)
        Cv = np.zeros_like(C)
        Cv[self.own1] = C[self.own1]
        hit = dil3(dil3(F & self.inb)) | dil3(Cv & self.inb)
        ab = np.zeros_like(E)
        ab[self.own] = (E & hit)[self.own]
        E[self.own] &= ~ab[self.own]
        st = self.local_stats.numpy()
        st[self.L: 2 * self.L] += np.bincount(self.idx[ab], minlength=self.L)
        b = 2 * self.L
        st[b + nat.ST_N_OUT] += int(ab.sum())
        st[b + nat.ST_N_EXCL] -= int(ab.sum())

    def flip(self):
        if self.ctrl["status"] != nat.EXIT_RUNNING or not self.ctrl["apply"]:
            return
        S, F = self._np(self.seg), self._np(self.flips)
        S ^= F & self.inb  # own planes and the (exchanged) halo planes
        # order-dependence counters of this update (k_quirks): own planes, looking one plane into the exchanged halos
        C = self._np(self.cancelled)
        Fv = F & self.inb
        aex, rem = Fv & S, Fv & ~S
        nonseg = ~S & self.inb
        st = self.local_stats.numpy()
        b = 2 * self.L
        o = self.own
        st[b + nat.ST_Q_CANCELLED] += int(C[o].sum())
        st[b + nat.ST_Q_ADD_INSIDE] += int((aex & ~dil3(nonseg))[o].sum())
        st[b + nat.ST_Q_REM_OUTSIDE] += int((rem & ~dil3(S))[o].sum())
        st[b + nat.ST_Q_REPROMOTED] += int((C & dil3(aex))[o].sum())

    def advance(self):
        c = self.ctrl
        if c["status"] != nat.EXIT_RUNNING:
            return
        g = self.global_stats.numpy()
        b = 2 * self.L
        c["sweeps"] += 1
        if g[b + nat.ST_N_FLIPS] == 0:
            c["status"] = nat.EXIT_CONVERGED
            return
        if not c["apply"]:
            c["status"] = nat.EXIT_MAX_SEGMENT
            return
        self.trace_rows.append((int(g[b + nat.ST_N_FLIPS]), int(g[b + nat.ST_N_IN]), int(g[b + nat.ST_N_OUT])))
        c["applied"] += 1
        c["iter"] += 1
        if c["iter"] > self.iter_max:
            c["status"] = nat.EXIT_MAX_ITER

    def poll(self):
        g = self.global_stats.numpy()
        b = 2 * self.L
        return {"iterations": self.ctrl["iter"], "exit_reason": self.ctrl["status"], "n_in": int(g[b + nat.ST_N_IN]),
                "n_out": int(g[b + nat.ST_N_OUT]), "n_excluded": int(g[b + nat.ST_N_EXCL]), "n_levels": self.L,
                "sweeps": self.ctrl["sweeps"], "kernel_launches": 0,
                "q_cancelled": int(g[b + nat.ST_Q_CANCELLED]), "q_add_to_inside": int(g[b + nat.ST_Q_ADD_INSIDE]),
                "q_remove_to_outside": int(g[b + nat.ST_Q_REM_OUTSIDE]), "q_cancel_repromoted": int(g[b + nat.ST_Q_REPROMOTED])}

    def trace(self):
        return np.asarray(self.trace_rows, dtype=np.int64)

    def labels(self):
        from oracle.vrg_oracle import canonical_labels
        S, E = self._np(self.seg), self._np(self.excl)
        nonseg = ~S & self.inb
        lab = np.full(S.shape, 3, dtype=np.uint8)
        lab[S] = 0
        lab[S & dil3(nonseg)] = 1
        lab[nonseg & dil3(S)] = 2
        lab[E] = 4
        assert canonical_labels is not None
        return lab[self.own]
