// Strict (list-order) mode, SURVEY.md section 8(f) N4: the reference's variational region growing WITH the effects of its
// list processing order -- Code/variationalRegionGrowing.py (VRG:line) update(), incremental branch VRG:156-259.
//
// The order-free path (vrg_b200.cu) restates the band state machine as set rules and is bit-identical to the reference
// wherever the reference's result does not depend on the order of its band lists.  On noisy inputs it does depend on it:
//   * flipped points are processed one after the other in the order of allBnd = innerBndList ++ outerBndList (VRG:163),
//     each tested against the labels as they are AT ITS TURN (VRG:169,198) -- an outer-band voxel whose segmented
//     neighbours left earlier in the same call is in no band any more and is skipped; one that a later neighbour
//     re-promotes is processed after all;
//   * a removed / added voxel is relabelled 2 / 1 without a look at its neighbours (VRG:174,202): stale band labels (Q2);
//   * the Parzen corrections are selected by the labels AFTER the call (VRG:232-233): a flipped voxel relabelled 0 or 3
//     by a later flip is left out (Q3), so the running sums innerProb / outerProb drift from the true ones and later
//     decisions follow the drifted sums (VRG:79-88).
// This file reproduces all of that on the GPU.  What makes a sequential list algorithm parallel:
//   1. list order = append time.  Every voxel is in a list at most once (list membership <=> label), so a list is the set
//      of voxels with that label ordered by a 64-bit key (iteration << 40 | rank of the flipped point being processed
//      << 5 | step within its processing).  allBnd order is a sort of the band by (inner first, key); `segmented` rows
//      come out in the order of a second key written when a voxel is added.
//   2. processing a flipped point reads labels within Chebyshev distance 2 and writes them within distance 1.  Two
//      flipped points further than 3 apart commute, so the call is a wavefront: in every round, each flipped point with
//      no unprocessed lower-ranked flipped point within its 7x7x7 neighbourhood is processed (one warp each: the 26
//      neighbours in turn, the 26 neighbours of each across the lanes).  Rounds = longest chain of order dependences.
//      The only writes with a longer reach, 4 -> 3 absorption at distance 2 (VRG:177-179,205-208), commute (no test
//      distinguishes 3 from 4 at the voxels it reads) and are claimed by a compare-and-swap so that each is counted once.
//   3. the sums go through the intensity levels, as in the order-free path: corrections and full sums are L-entry tables
//      (mat-vec of small integer histograms with the kernel values), added per band voxel in the reference's textual order
//      (VRG:242-247).  The reference adds the same terms voxel by voxel (np.sum); agreement is to rounding (tests: 1e-11).
// Single GPU, whole volume, fewer than 2^31 voxels, at most 65536 distinct intensities.  No CPU fallback.
#include "../../include/vrg_b200.h"

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

void vrg_set_error_internal(const char *msg);  // vrg_b200.cu: text behind vrg_last_error()

namespace {

namespace cg = cooperative_groups;
typedef unsigned long long u64;

constexpr int BLK = 256;
constexpr double VRG_A = 0.3989422804014327;  // (2*pi)**-0.5, VRG:7
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr int HASH_CAP = 1 << 18;
constexpr u64 EMPTY = ~0ull;

enum {  // counters (long long each)
    C_NF = 0, C_N_IN, C_N_OUT, C_N_EXCL, C_BAD_LABEL, C_NONFINITE, C_SKIPPED, C_DROPPED, C_ACT0, C_ACT1, C_READY0, C_READY1,
    C_ROUNDS, C_NLEV, C_N_BAND, C_COLLECT, C_WORDS
};

struct SP {
    int Z, Y, X, L;
    long long N, iter;
    double mhH;
    const uint8_t *vm_in;
    uint8_t *vm, *newf;
    uint16_t *lev;
    u64 *key, *skey;
    double *pin, *pout;
    uint32_t *rank;
    const double *levels;
    long long *cnt;
    long long *hin, *hout, *ha, *hb, *hc;  // L each, contiguous in this order
    double *tin, *tout, *ta, *tb, *tc;     // L each, contiguous in this order
    u64 *fkey;                             // flip / collect list: sort keys
    uint32_t *fvox;                        // ... and voxels
    uint32_t *act0, *act1, *ready, *blocker;
    long long cap;
};

__device__ __forceinline__ uint8_t ld8(const uint8_t *p) { return __ldcg(p); }
__device__ __forceinline__ u64 order_key(long long iter, long long rank, int sub) { return ((u64)iter << 40) | ((u64)rank << 5) | (u64)sub; }

// ---- levels: hash set of the distinct bit patterns, then a sorted table and a per-voxel index ---------------------------
__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(BLK) k_hash_levels(const double *__restrict__ data, long long n, u64 *table, long long *cnt) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = data[i] + 0.0;  // -0.0 and +0.0 are one level (np.unique)
        if (!isfinite(v)) { atomicAdd((u64 *)&cnt[C_NONFINITE], 1ull); continue; }
        const u64 b = (u64)__double_as_longlong(v);
        uint32_t s = (uint32_t)mix64(b) & (HASH_CAP - 1);
        for (int probe = 0; probe < HASH_CAP; ++probe) {
            const u64 cur = table[s];
            if (cur == b) break;
            if (cur == EMPTY) {
                const u64 prev = atomicCAS(&table[s], EMPTY, b);
                if (prev == EMPTY) { atomicAdd((u64 *)&cnt[C_NLEV], 1ull); break; }
                if (prev == b) break;
            }
            if (cnt[C_NLEV] > VRG_MAX_LEVELS) return;  // continuous data: give up early
            s = (s + 1) & (HASH_CAP - 1);
        }
    }
}
// order-preserving map double -> u64
__device__ __forceinline__ u64 sortable(u64 b) { return (b >> 63) ? ~b : (b | 0x8000000000000000ull); }
__device__ __forceinline__ u64 unsortable(u64 s) { return (s >> 63) ? (s & 0x7FFFFFFFFFFFFFFFull) : ~s; }
__global__ void __launch_bounds__(BLK) k_compact_levels(const u64 *__restrict__ table, u64 *out, long long *cnt, int padded) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HASH_CAP; i += gridDim.x * blockDim.x) {
        const u64 b = table[i];
        if (b != EMPTY) {
            const long long w = (long long)atomicAdd((u64 *)&cnt[C_COLLECT], 1ull);
            if (w < padded) out[w] = sortable(b);
        }
    }
}
__global__ void __launch_bounds__(BLK) k_levels_to_double(const u64 *__restrict__ sorted, double *levels, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L) levels[i] = __longlong_as_double((long long)unsortable(sorted[i]));
}
__global__ void __launch_bounds__(BLK) k_level_index(const double *__restrict__ data, long long n, const double *__restrict__ levels, int L,
                                                     uint16_t *lev) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = data[i] + 0.0;
        int lo = 0, hi = L - 1;
        while (lo < hi) { const int m = (lo + hi) >> 1; if (levels[m] < v) lo = m + 1; else hi = m; }
        lev[i] = (uint16_t)lo;
    }
}

// ---- bitonic sort of (key, voxel) pairs; keys are unique, the list is padded to a power of two with ~0 ---------------
__global__ void __launch_bounds__(BLK) k_pad(u64 *key, uint32_t *val, long long n, long long n2) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x + n;
    if (i < n2) { key[i] = EMPTY; if (val) val[i] = NONE; }
}
__global__ void __launch_bounds__(BLK) k_bitonic(u64 *key, uint32_t *val, long long n2, long long j, long long k) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n2) return;
    const long long l = i ^ j;
    if (l > i) {
        const bool up = (i & k) == 0;
        const u64 a = key[i], b = key[l];
        if ((a > b) == up) {
            key[i] = b; key[l] = a;
            if (val) { const uint32_t t = val[i]; val[i] = val[l]; val[l] = t; }
        }
    }
}
// all steps with j < 1024 of one k inside shared memory: 2048 elements per block
constexpr int BT = 2048;
__global__ void __launch_bounds__(BT / 2) k_bitonic_local(u64 *key, uint32_t *val, long long k, long long j0) {
    __shared__ u64 sk[BT];
    __shared__ uint32_t sv[BT];
    const long long base = (long long)blockIdx.x * BT;
    for (int t = threadIdx.x; t < BT; t += BT / 2) { sk[t] = key[base + t]; sv[t] = val ? val[base + t] : 0u; }
    __syncthreads();
    for (long long j = j0; j >= 1; j >>= 1) {
        const int t = threadIdx.x;
        const int i = (int)(((t & ~((int)j - 1)) << 1) | (t & ((int)j - 1)));  // the lower index of pair t
        const int l = i | (int)j;
        const bool up = ((base + i) & k) == 0;
        const u64 a = sk[i], b = sk[l];
        if ((a > b) == up) { sk[i] = b; sk[l] = a; const uint32_t x = sv[i]; sv[i] = sv[l]; sv[l] = x; }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < BT; t += BT / 2) { key[base + t] = sk[t]; if (val) val[base + t] = sv[t]; }
}

// ---- neighbourhood helpers ------------------------------------------------------------------------------------------
struct Pos { int z, y, x; };
__device__ __forceinline__ Pos pos_of(const SP &p, long long v) {
    Pos q;
    q.x = (int)(v % p.X);
    const long long t = v / p.X;
    q.y = (int)(t % p.Y);
    q.z = (int)(t / p.Y);
    return q;
}
// neighbour o (0..26 in C order of the offsets, 13 = centre) of a voxel, or -1 outside the volume: get_neighbours, VRG:263-282
__device__ __forceinline__ long long nbr(const SP &p, const Pos &c, int o) {
    const int z = c.z + o / 9 - 1, y = c.y + (o / 3) % 3 - 1, x = c.x + o % 3 - 1;
    if (z < 0 || z >= p.Z || y < 0 || y >= p.Y || x < 0 || x >= p.X) return -1;
    return ((long long)z * p.Y + y) * p.X + x;
}

// ---- init branch, VRG:129-145, as a gather (the seed set does not change during it) -------------------------------------
__global__ void __launch_bounds__(BLK) k_init(SP p) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < p.N; v += (long long)gridDim.x * blockDim.x) {
        const uint8_t in = p.vm_in[v];
        if (in != 0 && in != 3 && in != 4) { atomicAdd((u64 *)&p.cnt[C_BAD_LABEL], 1ull); continue; }
        const Pos c = pos_of(p, v);
        uint8_t lab;
        u64 key = 0;
        if (in == 0) {  // seed: inner band iff it has an unsegmented neighbour; lists are filled in raster order of the seeds
            bool open = false;
            for (int o = 0; o < 27 && !open; ++o) {
                const long long q = nbr(p, c, o);
                if (o != 13 && q >= 0 && p.vm_in[q] != 0) open = true;
            }
            lab = open ? 1 : 0;
            key = order_key(0, v, 0);
            p.skey[v] = key;
        } else {  // outer band iff a seed touches it: appended while its first seed neighbour (raster order) is processed
            long long first = -1;
            int sub = 0;
            for (int o = 0; o < 27; ++o) {
                const long long q = nbr(p, c, o);
                if (o != 13 && q >= 0 && p.vm_in[q] == 0) { first = q; sub = 26 - o + 1; break; }  // this voxel is neighbour 26 - o of q
            }
            if (first >= 0) { lab = 2; key = order_key(0, first, sub); }
            else lab = in;  // 3 or 4 (VRG:137 turns the 4s next to a seed into 3s; those are all outer band)
        }
        p.vm[v] = lab;
        p.key[v] = key;
        p.pin[v] = 0.0; p.pout[v] = 0.0; p.rank[v] = 0u; p.newf[v] = 0;
        const int b = p.lev[v];
        if (lab <= 1) { atomicAdd((u64 *)&p.hin[b], 1ull); atomicAdd((u64 *)&p.cnt[C_N_IN], 1ull); }
        else if (lab <= 3) { atomicAdd((u64 *)&p.hout[b], 1ull); atomicAdd((u64 *)&p.cnt[C_N_OUT], 1ull); }
        else atomicAdd((u64 *)&p.cnt[C_N_EXCL], 1ull);
        if (lab == 1 || lab == 2) atomicAdd((u64 *)&p.cnt[C_N_BAND], 1ull);
    }
}

// ---- tables: out[t][b] = sum_c hist[t][c] * A * exp(-0.5 H (level_c - level_b)^2), fixed summation order --------------
__global__ void __launch_bounds__(BLK) k_tables(SP p, int first, int count) {
    const int lane = threadIdx.x & 31;
    const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long job = w; job < (long long)count * p.L; job += nw) {
        const int t = first + (int)(job / p.L), b = (int)(job % p.L);
        const long long *hist = p.hin + (long long)t * p.L;
        const double lb = p.levels[b];
        double acc = 0.0;
        for (int c = lane; c < p.L; c += 32) {
            const long long h = hist[c];
            if (h) { const double d = p.levels[c] - lb; acc += (double)h * (VRG_A * exp(p.mhH * d * d)); }
        }
        for (int s = 16; s; s >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s);
        if (lane == 0) p.tin[(long long)t * p.L + b] = acc;
    }
}

// full sums of every band voxel after the init branch, VRG:148-155
__global__ void __launch_bounds__(BLK) k_init_sums(SP p) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < p.N; v += (long long)gridDim.x * blockDim.x) {
        const uint8_t l = p.vm[v];
        if (l == 1 || l == 2) { p.pin[v] = p.tin[p.lev[v]]; p.pout[v] = p.tout[p.lev[v]]; }
    }
}

// ---- decision, VRG:79-88: list every band voxel whose side disagrees with pin/n_in >= pout/n_out (ties inside) ------
// mode 0: flipped points (sort key: inner band first, then list order); 1: the whole band, same key; 2: segmented voxels, key = skey
__global__ void __launch_bounds__(BLK) k_collect(SP p, int mode) {
    const double nin = (double)p.cnt[C_N_IN], nout = (double)p.cnt[C_N_OUT];
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < p.N; v += (long long)gridDim.x * blockDim.x) {
        const uint8_t l = p.vm[v];
        bool take;
        u64 k;
        if (mode == 2) { take = l <= 1; k = take ? p.skey[v] : 0; }
        else {
            take = l == 1 || l == 2;
            if (take && mode == 0) {
                const bool inside = p.pin[v] / nin >= p.pout[v] / nout;
                take = (l == 1) != inside;
            }
            k = take ? (p.key[v] | (l == 2 ? (1ull << 63) : 0ull)) : 0;
        }
        if (take) {
            const long long w = (long long)atomicAdd((u64 *)&p.cnt[C_COLLECT], 1ull);
            if (w < p.cap) { p.fkey[w] = k; p.fvox[w] = (uint32_t)v; }
        }
    }
}
__global__ void __launch_bounds__(BLK) k_scatter_rank(SP p, long long nf) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < nf) { p.rank[p.fvox[i]] = (uint32_t)(i + 1); p.blocker[i] = NONE; }
}

// ---- the wavefront ------------------------------------------------------------------------------------------------------
// 4 -> 3 (the voxel joins includedPoints, VRG:165-168,177-179,205-208): claimed by compare-and-swap on the 32-bit word
__device__ __forceinline__ void absorb(const SP &p, long long v) {
    if (v < 0 || ld8(p.vm + v) != 4) return;
    unsigned int *w = (unsigned int *)(p.vm + (v & ~3ll));
    const int sh = (int)(v & 3) * 8;
    unsigned int old = *(volatile unsigned int *)w;
    while (((old >> sh) & 0xFFu) == 4u) {
        const unsigned int prev = atomicCAS(w, old, old ^ (7u << sh));  // 4 ^ 7 = 3
        if (prev == old) {
            const int b = p.lev[v];
            atomicAdd((u64 *)&p.hc[b], 1ull); atomicAdd((u64 *)&p.hout[b], 1ull);
            atomicAdd((u64 *)&p.cnt[C_N_OUT], 1ull); atomicAdd((u64 *)&p.cnt[C_N_EXCL], (u64)-1ll);
            return;
        }
        old = prev;
    }
}

// one flipped point, one warp: VRG:163-228.  The 5x5x5 labels around the point are fetched once into shared memory (`lab`, 128
// bytes per warp; 0xFF = outside the volume): nothing else writes them while this point is processed (points processed side by
// side are further than 3 apart), so the 26 neighbours are walked on the cached copy -- their order matters (a neighbour
// promoted at its step is seen by the tests of the later ones) -- and each lane then writes its own neighbour's new state back.
__device__ void process_flip(const SP &p, long long idx, int lane, uint8_t *lab) {
    const long long vox = p.fvox[idx];
    const Pos c = pos_of(p, vox);
    long long g[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * t;
        g[t] = -1;
        if (i < 125) {
            const int z = c.z + i / 25 - 2, y = c.y + (i / 5) % 5 - 2, x = c.x + i % 5 - 2;
            if (z >= 0 && z < p.Z && y >= 0 && y < p.Y && x >= 0 && x < p.X) g[t] = ((long long)z * p.Y + y) * p.X + x;
            lab[i] = g[t] >= 0 ? ld8(p.vm + g[t]) : (uint8_t)0xFF;
        }
    }
    __syncwarp();
    const uint8_t lab0 = lab[62];
    const bool exec = lab0 == 1 || lab0 == 2;
    // 4 -> 3: the 3x3x3 of every listed point (VRG:165-168), the 5x5x5 of a processed one (VRG:177-179,205-208: the 3x3x3 of
    // each of its in-bounds neighbours).  No label test below tells 3 from 4, so all of it can come first.
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int i = lane + 32 * t;
        if (i < 125 && g[t] >= 0 && lab[i] == 4) {
            const bool inner = abs(i / 25 - 2) <= 1 && abs((i / 5) % 5 - 2) <= 1 && abs(i % 5 - 2) <= 1;
            if (exec || inner) { absorb(p, g[t]); lab[i] = 3; }
        }
    }
    __syncwarp();
    if (!exec) {  // in no band any more at its turn
        if (lane == 0) { atomicAdd((u64 *)&p.cnt[C_SKIPPED], 1ull); p.rank[vox] = 0u; }
        return;
    }
    const bool removal = lab0 == 1;
    if (lane == 0) {
        const int b = p.lev[vox];
        lab[62] = removal ? 2 : 1;
        p.vm[vox] = removal ? 2 : 1;  // VRG:174,202: no look at the neighbours
        p.key[vox] = order_key(p.iter, idx, 0);
        if (removal) {
            atomicAdd((u64 *)&p.hin[b], (u64)-1ll); atomicAdd((u64 *)&p.hout[b], 1ull);
            atomicAdd((u64 *)&p.cnt[C_N_IN], (u64)-1ll); atomicAdd((u64 *)&p.cnt[C_N_OUT], 1ull);
        } else {
            p.skey[vox] = order_key(p.iter, idx, 0);  // segmentedList.append, VRG:200
            atomicAdd((u64 *)&p.hin[b], 1ull); atomicAdd((u64 *)&p.hout[b], (u64)-1ll);
            atomicAdd((u64 *)&p.cnt[C_N_IN], 1ull); atomicAdd((u64 *)&p.cnt[C_N_OUT], (u64)-1ll);
        }
    }
    __syncwarp();
    // the labels a neighbour can move between: removal: 2 -> 3 unless a 1 is near, 0 -> 1;  addition: 1 -> 0 unless a 2 is near, 3 -> 2
    const uint8_t settle_from = removal ? 2 : 1, needs = removal ? 1 : 2, settle_to = removal ? 3 : 0;
    const uint8_t promote_from = removal ? 0 : 3, promote_to = removal ? 1 : 2;
    const bool nlane = lane < 27 && lane != 13;
    const int myoff = (lane / 9 - 1) * 25 + ((lane / 3) % 3 - 1) * 5 + (lane % 3 - 1);  // neighbour `lane` in the 5^3 cache
    int action = 0;  // what happened to neighbour `lane`: 1 settled, 2 promoted
    for (int o = 0; o < 27; ++o) {
        if (o == 13) continue;
        const int qi = 62 + (o / 9 - 1) * 25 + ((o / 3) % 3 - 1) * 5 + (o % 3 - 1);
        const uint8_t lq = lab[qi];
        if (lq == settle_from) {  // VRG:183-190,218-227
            const unsigned near = __ballot_sync(0xFFFFFFFFu, nlane && lab[qi + myoff] == needs);
            if (!near) {
                if (lane == o) action = 1;
                if (lane == 0) lab[qi] = settle_to;
            }
        } else if (lq == promote_from) {  // VRG:193-196,209-212
            if (lane == o) action = 2;
            if (lane == 0) lab[qi] = promote_to;
        }
        __syncwarp();
    }
    if (action) {
        const long long q = nbr(p, c, lane);
        if (action == 1) { p.vm[q] = settle_to; p.pin[q] = 0.0; p.pout[q] = 0.0; }
        else { p.vm[q] = promote_to; p.newf[q] = 1; p.key[q] = order_key(p.iter, idx, lane + 1); }
    }
    if (lane == 0) p.rank[vox] = 0u;
    __syncwarp();
}

__global__ void __launch_bounds__(BLK) k_wave(SP p, long long nf) {
    cg::grid_group grid = cg::this_grid();
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsize = (long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const long long gwarp = gtid >> 5, nwarps = gsize >> 5;
    volatile long long *cnt = p.cnt;
    __shared__ uint8_t s_lab[BLK / 32][128];
    long long n_act = nf;
    int round = 0;
    while (n_act > 0) {
        const int cur = round & 1;
        const uint32_t *act = cur ? p.act1 : p.act0;
        uint32_t *next = cur ? p.act0 : p.act1;
        // phase A: which of the unprocessed flipped points have no unprocessed predecessor within distance 3?
        for (long long i = gtid; i < n_act; i += gsize) {
            const uint32_t idx = round == 0 ? (uint32_t)i : act[i];
            const uint32_t me = idx + 1;
            bool blocked = false;
            const uint32_t b = p.blocker[idx];
            if (b != NONE && __ldcg(p.rank + b) != 0u) blocked = true;
            else {
                const Pos c = pos_of(p, p.fvox[idx]);
                uint32_t best = 0, bestv = NONE;
                for (int dz = -3; dz <= 3; ++dz) {
                    const int z = c.z + dz;
                    if (z < 0 || z >= p.Z) continue;
                    for (int dy = -3; dy <= 3; ++dy) {
                        const int y = c.y + dy;
                        if (y < 0 || y >= p.Y) continue;
                        const long long row = ((long long)z * p.Y + y) * p.X;
                        for (int dx = -3; dx <= 3; ++dx) {
                            const int x = c.x + dx;
                            if (x < 0 || x >= p.X) continue;
                            const uint32_t r = __ldcg(p.rank + row + x);
                            if (r != 0u && r < me && r > best) { best = r; bestv = (uint32_t)(row + x); }
                        }
                    }
                }
                if (best) { blocked = true; p.blocker[idx] = bestv; }  // the latest predecessor: likely the last one to finish
            }
            if (blocked) next[atomicAdd((u64 *)&p.cnt[C_ACT0 + (cur ^ 1)], 1ull)] = idx;
            else p.ready[atomicAdd((u64 *)&p.cnt[C_READY0 + cur], 1ull)] = idx;
        }
        grid.sync();
        const long long n_ready = cnt[C_READY0 + cur], n_next = cnt[C_ACT0 + (cur ^ 1)];
        // phase B: they are pairwise further than 3 apart: process them side by side, one warp each
        for (long long j = gwarp; j < n_ready; j += nwarps) process_flip(p, __ldcg(p.ready + j), lane, s_lab[threadIdx.x >> 5]);
        if (gtid == 0) { cnt[C_READY0 + (cur ^ 1)] = 0; cnt[C_ACT0 + cur] = 0; cnt[C_ROUNDS] += 1; }
        grid.sync();
        n_act = n_next;
        ++round;
    }
}

// VRG:232-233: the corrections are selected by the labels after the call
__global__ void __launch_bounds__(BLK) k_post(SP p, long long nf) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const long long v = p.fvox[i];
    const uint8_t l = p.vm[v];
    if (l == 1) atomicAdd((u64 *)&p.ha[p.lev[v]], 1ull);
    else if (l == 2) atomicAdd((u64 *)&p.hb[p.lev[v]], 1ull);
    else atomicAdd((u64 *)&p.cnt[C_DROPPED], 1ull);
}
// VRG:236-255: corrections at every band voxel in the reference's textual order, then full sums where a voxel entered a band
__global__ void __launch_bounds__(BLK) k_apply(SP p) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < p.N; v += (long long)gridDim.x * blockDim.x) {
        const uint8_t l = p.vm[v];
        const uint8_t nw = p.newf[v];
        if (nw) { p.newf[v] = 0; p.pin[v] = p.tin[p.lev[v]]; p.pout[v] = p.tout[p.lev[v]]; }
        else if (l == 1 || l == 2) {
            const int b = p.lev[v];
            double a = p.pin[v], o = p.pout[v];
            a += p.ta[b]; a -= p.tb[b];
            o -= p.ta[b]; o += p.tb[b]; o += p.tc[b];
            p.pin[v] = a; p.pout[v] = o;
        }
    }
}
__global__ void __launch_bounds__(BLK) k_gather_sums(SP p, long long n, double *a, double *b) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = p.fvox[i];
    a[i] = p.pin[v] / (double)p.cnt[C_N_IN];
    b[i] = p.pout[v] / (double)p.cnt[C_N_OUT];
}
__global__ void __launch_bounds__(BLK) k_seg_map(const uint8_t *__restrict__ vm, uint8_t *out, long long n) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) out[v] = vm[v] <= 1;
}

}  // namespace

struct vrg_strict {
    int device = 0, sms = 148, wave_blocks = 148;
    SP p;
    double H = 2.25, max_seconds = 0;
    int64_t iter_max = 200, max_seg = 5000, iter_num = 1, exit_code = VRG_EXIT_RUNNING, launches = 0;
    cudaStream_t st = nullptr;
    uint8_t *d_vm_in = nullptr;
    double *d_levels = nullptr;
    long long *d_hist = nullptr;
    double *d_tab = nullptr;
    long long *h_cnt = nullptr;  // pinned
    std::vector<int64_t> trace;
    bool inited = false;
    std::chrono::steady_clock::time_point t0;
};

static int sfail(int code, const std::string &msg) { vrg_set_error_internal(msg.c_str()); return code; }
#define SCK(call)                                                                                                     \
    do {                                                                                                              \
        cudaError_t e_ = (call);                                                                                      \
        if (e_ != cudaSuccess) {                                                                                      \
            char buf_[400];                                                                                           \
            snprintf(buf_, sizeof buf_, "strict mode: %s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return sfail(e_ == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA, buf_);                        \
        }                                                                                                             \
    } while (0)

static void free_lists(vrg_strict *h) {
    SP &p = h->p;
    cudaFree(p.fkey); cudaFree(p.fvox); cudaFree(p.act0); cudaFree(p.act1); cudaFree(p.ready); cudaFree(p.blocker);
    p.fkey = nullptr; p.fvox = p.act0 = p.act1 = p.ready = p.blocker = nullptr; p.cap = 0;
}
static long long pow2_at_least(long long n) { long long c = BT; while (c < n) c <<= 1; return c; }
static int ensure_cap(vrg_strict *h, long long n) {  // list buffers hold the next power of two (bitonic padding)
    SP &p = h->p;
    const long long want = pow2_at_least(n);
    if (want <= p.cap) return VRG_OK;
    SCK(cudaStreamSynchronize(h->st));
    free_lists(h);
    SCK(cudaMalloc((void **)&p.fkey, want * sizeof(u64)));
    SCK(cudaMalloc((void **)&p.fvox, want * sizeof(uint32_t)));
    SCK(cudaMalloc((void **)&p.act0, want * sizeof(uint32_t)));
    SCK(cudaMalloc((void **)&p.act1, want * sizeof(uint32_t)));
    SCK(cudaMalloc((void **)&p.ready, want * sizeof(uint32_t)));
    SCK(cudaMalloc((void **)&p.blocker, want * sizeof(uint32_t)));
    p.cap = want;
    return VRG_OK;
}
static int grid_for(const vrg_strict *h, long long n) { return (int)std::max<long long>(1, std::min<long long>((n + BLK - 1) / BLK, (long long)h->sms * 16)); }

// sort the first n pairs of (fkey, fvox) ascending
static int sort_pairs(vrg_strict *h, u64 *key, uint32_t *val, long long n) {
    if (n <= 1) return VRG_OK;
    const long long n2 = pow2_at_least(n);
    if (n2 > n) k_pad<<<(unsigned)((n2 - n + BLK - 1) / BLK), BLK, 0, h->st>>>(key, val, n, n2);
    const unsigned g = (unsigned)((n2 + BLK - 1) / BLK);
    for (long long k = 2; k <= n2; k <<= 1) {
        long long j = k >> 1;
        for (; j >= BT; j >>= 1) { k_bitonic<<<g, BLK, 0, h->st>>>(key, val, n2, j, k); h->launches++; }
        k_bitonic_local<<<(unsigned)(n2 / BT), BT / 2, 0, h->st>>>(key, val, k, j);
        h->launches++;
    }
    SCK(cudaGetLastError());
    return VRG_OK;
}

// collect + sort: mode 0 flipped points, 1 band (allBnd order), 2 segmented (list order); the count comes back in *n
static int collect_sorted(vrg_strict *h, int mode, long long *n) {
    SP &p = h->p;
    for (int attempt = 0; attempt < 2; ++attempt) {
        SCK(cudaMemsetAsync(p.cnt + C_COLLECT, 0, sizeof(long long), h->st));
        k_collect<<<grid_for(h, p.N), BLK, 0, h->st>>>(p, mode);
        h->launches++;
        SCK(cudaMemcpyAsync(h->h_cnt, p.cnt, C_WORDS * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
        SCK(cudaStreamSynchronize(h->st));
        *n = h->h_cnt[C_COLLECT];
        if (*n <= p.cap) break;
        int rc = ensure_cap(h, *n);
        if (rc != VRG_OK) return rc;
    }
    return sort_pairs(h, p.fkey, p.fvox, *n);
}

extern "C" {

int vrg_strict_create(int device, const int64_t *shape, double H, int64_t iter_max, int64_t max_segment_size, double max_seconds,
                      vrg_strict **out) {
    if (!shape || !out) return sfail(VRG_ERR_ARG, "strict mode: null argument");
    const int64_t Z = shape[0], Y = shape[1], X = shape[2];
    if (Z <= 0 || Y <= 0 || X <= 0 || Z * Y * X >= 0x7FFFFFF0ll) return sfail(VRG_ERR_ARG, "strict mode: bad shape or 2^31 voxels and more");
    if (!(H > 0) || iter_max < 1 || iter_max >= (1 << 22)) return sfail(VRG_ERR_ARG, "strict mode: bad H or iter_max");
    int ndev = 0;
    SCK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return sfail(VRG_ERR_ARG, "strict mode: bad device");
    SCK(cudaSetDevice(device));
    vrg_strict *h = new vrg_strict();
    h->device = device; h->H = H; h->iter_max = iter_max; h->max_seg = max_segment_size; h->max_seconds = max_seconds;
    cudaDeviceProp prop;
    SCK(cudaGetDeviceProperties(&prop, device));
    h->sms = prop.multiProcessorCount;
    int coop = 0, per_sm = 0;
    SCK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) { delete h; return sfail(VRG_ERR_CUDA, "strict mode: the device cannot launch cooperative kernels"); }
    SCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave, BLK, 0));
    h->wave_blocks = h->sms * std::max(1, std::min(per_sm, 4));
    SP &p = h->p;
    memset(&p, 0, sizeof p);
    p.Z = (int)Z; p.Y = (int)Y; p.X = (int)X; p.N = Z * Y * X; p.mhH = -0.5 * H;
    const size_t N4 = ((size_t)p.N + 3) & ~(size_t)3;
    SCK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void **ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
    alloc((void **)&h->d_vm_in, N4);
    alloc((void **)&p.vm, N4);
    alloc((void **)&p.newf, N4);
    alloc((void **)&p.lev, N4 * sizeof(uint16_t));
    alloc((void **)&p.key, N4 * sizeof(u64));
    alloc((void **)&p.skey, N4 * sizeof(u64));
    alloc((void **)&p.pin, N4 * sizeof(double));
    alloc((void **)&p.pout, N4 * sizeof(double));
    alloc((void **)&p.rank, N4 * sizeof(uint32_t));
    alloc((void **)&p.cnt, C_WORDS * sizeof(long long));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_cnt, C_WORDS * sizeof(long long));
    if (e != cudaSuccess) {
        vrg_strict_destroy(h);
        return sfail(e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA, std::string("strict mode: allocation: ") + cudaGetErrorString(e));
    }
    p.vm_in = h->d_vm_in;
    *out = h;
    return VRG_OK;
}

int vrg_strict_destroy(vrg_strict *h) {
    if (!h) return VRG_OK;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    SP &p = h->p;
    cudaFree(h->d_vm_in); cudaFree(p.vm); cudaFree(p.newf); cudaFree(p.lev); cudaFree(p.key); cudaFree(p.skey);
    cudaFree(p.pin); cudaFree(p.pout); cudaFree(p.rank); cudaFree(p.cnt); cudaFree(h->d_levels); cudaFree(h->d_hist); cudaFree(h->d_tab);
    free_lists(h);
    if (h->h_cnt) cudaFreeHost(h->h_cnt);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return VRG_OK;
}

// inputs + the init branch (VRG:40-50,129-155)
int vrg_strict_init(vrg_strict *h, const double *data_host, const uint8_t *value_map_host) {
    if (!h || !data_host || !value_map_host) return sfail(VRG_ERR_ARG, "strict mode: null argument");
    SCK(cudaSetDevice(h->device));
    SP &p = h->p;
    h->t0 = std::chrono::steady_clock::now();
    h->inited = false;
    const long long N = p.N;
    // levels (the temporaries go back before the lists are sized)
    double *d_data = nullptr;
    u64 *d_table = nullptr, *d_sorted = nullptr;
    SCK(cudaMalloc((void **)&d_data, (size_t)N * sizeof(double)));
    SCK(cudaMalloc((void **)&d_table, (size_t)HASH_CAP * sizeof(u64)));
    SCK(cudaMalloc((void **)&d_sorted, (size_t)(2 * VRG_MAX_LEVELS) * sizeof(u64)));
    auto drop = [&]() { cudaFree(d_data); cudaFree(d_table); cudaFree(d_sorted); };
    cudaError_t e = cudaMemcpyAsync(d_data, data_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_vm_in, value_map_host, (size_t)N, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_table, 0xFF, (size_t)HASH_CAP * sizeof(u64), h->st);
    if (e == cudaSuccess) e = cudaMemsetAsync(p.cnt, 0, C_WORDS * sizeof(long long), h->st);
    if (e != cudaSuccess) { drop(); return sfail(VRG_ERR_CUDA, std::string("strict mode: upload: ") + cudaGetErrorString(e)); }
    k_hash_levels<<<grid_for(h, N), BLK, 0, h->st>>>(d_data, N, d_table, p.cnt);
    const int padded = 2 * VRG_MAX_LEVELS;
    k_compact_levels<<<h->sms, BLK, 0, h->st>>>(d_table, d_sorted, p.cnt, padded);
    h->launches += 2;
    e = cudaMemcpyAsync(h->h_cnt, p.cnt, C_WORDS * sizeof(long long), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) { drop(); return sfail(VRG_ERR_CUDA, std::string("strict mode: level scan: ") + cudaGetErrorString(e)); }
    if (h->h_cnt[C_NONFINITE]) { drop(); return sfail(VRG_ERR_NONFINITE, "strict mode: intensity volume holds NaN or Inf"); }
    const long long L = h->h_cnt[C_COLLECT];
    if (L > VRG_MAX_LEVELS || h->h_cnt[C_NLEV] > VRG_MAX_LEVELS) {
        drop();
        return sfail(VRG_ERR_LEVELS, "strict mode: more than 65536 distinct intensity levels");
    }
    {
        int rc = sort_pairs(h, d_sorted, nullptr, L);
        if (rc != VRG_OK) { drop(); return rc; }
    }
    if (p.L != (int)L || !h->d_levels) {
        cudaFree(h->d_levels); cudaFree(h->d_hist); cudaFree(h->d_tab);
        h->d_levels = nullptr; h->d_hist = nullptr; h->d_tab = nullptr;
        e = cudaMalloc((void **)&h->d_levels, (size_t)L * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_hist, (size_t)5 * L * sizeof(long long));
        if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_tab, (size_t)5 * L * sizeof(double));
        if (e != cudaSuccess) { drop(); return sfail(VRG_ERR_NOMEM, "strict mode: table allocation failed"); }
    }
    p.L = (int)L;
    p.levels = h->d_levels;
    p.hin = h->d_hist; p.hout = p.hin + L; p.ha = p.hout + L; p.hb = p.ha + L; p.hc = p.hb + L;
    p.tin = h->d_tab; p.tout = p.tin + L; p.ta = p.tout + L; p.tb = p.ta + L; p.tc = p.tb + L;
    k_levels_to_double<<<(unsigned)((L + BLK - 1) / BLK), BLK, 0, h->st>>>(d_sorted, h->d_levels, (int)L);
    k_level_index<<<grid_for(h, N), BLK, 0, h->st>>>(d_data, N, h->d_levels, (int)L, p.lev);
    h->launches += 2;
    e = cudaStreamSynchronize(h->st);
    drop();
    if (e != cudaSuccess) return sfail(VRG_ERR_CUDA, std::string("strict mode: level index: ") + cudaGetErrorString(e));
    // init branch
    SCK(cudaMemsetAsync(p.cnt, 0, C_WORDS * sizeof(long long), h->st));
    SCK(cudaMemsetAsync(h->d_hist, 0, (size_t)5 * L * sizeof(long long), h->st));
    k_init<<<grid_for(h, N), BLK, 0, h->st>>>(p);
    k_tables<<<grid_for(h, 2 * L * 32), BLK, 0, h->st>>>(p, 0, 2);
    k_init_sums<<<grid_for(h, N), BLK, 0, h->st>>>(p);
    h->launches += 3;
    SCK(cudaMemcpyAsync(h->h_cnt, p.cnt, C_WORDS * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
    SCK(cudaStreamSynchronize(h->st));
    SCK(cudaGetLastError());
    if (h->h_cnt[C_BAD_LABEL]) return sfail(VRG_ERR_LABEL, "strict mode: the initial valueMap may only hold the labels 0, 3 and 4");
    if (h->h_cnt[C_N_IN] == 0) return sfail(VRG_ERR_EMPTY_SEED, "strict mode: empty seed set (reference: IndexError at VRG:88)");
    if (h->h_cnt[C_N_BAND] == 0) return sfail(VRG_ERR_NO_BAND, "strict mode: seed has no boundary (reference: IndexError at VRG:88)");
    {
        int rc = ensure_cap(h, std::max<long long>(4 * h->h_cnt[C_N_BAND], 1 << 16));
        if (rc != VRG_OK) return rc;
    }
    h->iter_num = 1;
    h->exit_code = VRG_EXIT_RUNNING;
    h->trace.assign({-1, (int64_t)h->h_cnt[C_N_IN], (int64_t)h->h_cnt[C_N_OUT]});
    h->inited = true;
    return VRG_OK;
}

static void fill_result(vrg_strict *h, vrg_strict_result *r) {
    if (!r) return;
    r->iterations = h->iter_num; r->exit_reason = h->exit_code;
    r->n_in = h->h_cnt[C_N_IN]; r->n_out = h->h_cnt[C_N_OUT]; r->n_excluded = h->h_cnt[C_N_EXCL]; r->n_levels = h->p.L;
    r->kernel_launches = h->launches; r->skipped = h->h_cnt[C_SKIPPED]; r->dropped = h->h_cnt[C_DROPPED];
    r->rounds = h->h_cnt[C_ROUNDS];
}

// one pass of the while loop, VRG:58-117
int vrg_strict_step(vrg_strict *h, vrg_strict_result *res) {
    if (!h || !h->inited) return sfail(VRG_ERR_ARG, "strict mode: init first");
    SCK(cudaSetDevice(h->device));
    SP &p = h->p;
    if (h->exit_code != VRG_EXIT_RUNNING) { fill_result(h, res); return VRG_OK; }
    if (h->iter_num > h->iter_max) { h->exit_code = VRG_EXIT_MAX_ITER; fill_result(h, res); return VRG_OK; }  // VRG:118-121
    long long nf = 0;
    {
        int rc = collect_sorted(h, 0, &nf);
        if (rc != VRG_OK) return rc;
    }
    if (nf == 0) h->exit_code = VRG_EXIT_CONVERGED;  // VRG:91
    else if (h->max_seconds > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - h->t0).count() >= h->max_seconds)
        h->exit_code = VRG_EXIT_MAX_TIME;            // VRG:97
    else if (h->h_cnt[C_N_IN] >= h->max_seg) h->exit_code = VRG_EXIT_MAX_SEGMENT;  // VRG:101
    if (h->exit_code != VRG_EXIT_RUNNING) { fill_result(h, res); return VRG_OK; }
    p.iter = h->iter_num;
    const unsigned gf = (unsigned)((nf + BLK - 1) / BLK);
    k_scatter_rank<<<gf, BLK, 0, h->st>>>(p, nf);
    SCK(cudaMemsetAsync(p.ha, 0, (size_t)3 * p.L * sizeof(long long), h->st));  // ha, hb, hc
    SCK(cudaMemsetAsync(p.cnt + C_ACT0, 0, 4 * sizeof(long long), h->st));
    {
        void *args[] = {(void *)&p, (void *)&nf};
        const int blocks = (int)std::max<long long>(1, std::min<long long>(h->wave_blocks, (nf * 32 + BLK - 1) / BLK));
        SCK(cudaLaunchCooperativeKernel((void *)k_wave, dim3(blocks), dim3(BLK), args, 0, h->st));
    }
    k_post<<<gf, BLK, 0, h->st>>>(p, nf);
    k_tables<<<grid_for(h, 5ll * p.L * 32), BLK, 0, h->st>>>(p, 0, 5);
    k_apply<<<grid_for(h, p.N), BLK, 0, h->st>>>(p);
    h->launches += 5;
    SCK(cudaMemcpyAsync(h->h_cnt, p.cnt, C_WORDS * sizeof(long long), cudaMemcpyDeviceToHost, h->st));
    SCK(cudaStreamSynchronize(h->st));
    SCK(cudaGetLastError());
    h->trace.insert(h->trace.end(), {(int64_t)nf, (int64_t)h->h_cnt[C_N_IN], (int64_t)h->h_cnt[C_N_OUT]});
    h->iter_num++;
    fill_result(h, res);
    return VRG_OK;
}

int vrg_strict_run(vrg_strict *h, vrg_strict_result *res) {
    if (!h || !h->inited) return sfail(VRG_ERR_ARG, "strict mode: init first");
    while (h->exit_code == VRG_EXIT_RUNNING) {
        int rc = vrg_strict_step(h, res);
        if (rc != VRG_OK) return rc;
    }
    fill_result(h, res);
    return VRG_OK;
}

// outputs: the reference's valueMap as it stands (stale band labels included) and segmentedMap (0/1)
int vrg_strict_download(vrg_strict *h, uint8_t *value_map_out, uint8_t *seg_map_out) {
    if (!h || !h->inited) return sfail(VRG_ERR_ARG, "strict mode: init first");
    SCK(cudaSetDevice(h->device));
    SP &p = h->p;
    if (value_map_out) SCK(cudaMemcpyAsync(value_map_out, p.vm, (size_t)p.N, cudaMemcpyDeviceToHost, h->st));
    if (seg_map_out) {
        k_seg_map<<<grid_for(h, p.N), BLK, 0, h->st>>>(p.vm, p.newf, p.N);  // newf is all zero between steps: scratch
        h->launches++;
        SCK(cudaMemcpyAsync(seg_map_out, p.newf, (size_t)p.N, cudaMemcpyDeviceToHost, h->st));
        SCK(cudaMemsetAsync(p.newf, 0, (size_t)p.N, h->st));
    }
    SCK(cudaStreamSynchronize(h->st));
    return VRG_OK;
}

// which: 0 = the band in allBnd order (VRG:48,111) with the normalised sums the next decision reads; 1 = segmented voxels
// in the reference's list order (VRG:126,172,200).  vox_out: flat voxel indices; cap in entries; *n = how many there are.
int vrg_strict_list(vrg_strict *h, int which, int64_t *vox_out, double *pin_out, double *pout_out, int64_t cap, int64_t *n) {
    if (!h || !h->inited || !n || (which != 0 && which != 1)) return sfail(VRG_ERR_ARG, "strict mode: bad argument");
    SCK(cudaSetDevice(h->device));
    SP &p = h->p;
    long long cnt = 0;
    {
        int rc = collect_sorted(h, which == 0 ? 1 : 2, &cnt);
        if (rc != VRG_OK) return rc;
    }
    *n = cnt;
    if (cnt > cap) return sfail(VRG_ERR_ARG, "strict mode: list buffer too small");
    if (cnt == 0) return VRG_OK;
    std::vector<uint32_t> v((size_t)cnt);
    SCK(cudaMemcpyAsync(v.data(), p.fvox, (size_t)cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->st));
    if (which == 0 && pin_out && pout_out) {
        double *a = (double *)p.fkey, *b = nullptr;  // the sort keys are spent; the second array comes from act0 + act1?  no: allocate
        SCK(cudaMalloc((void **)&b, (size_t)cnt * sizeof(double)));
        k_gather_sums<<<(unsigned)((cnt + BLK - 1) / BLK), BLK, 0, h->st>>>(p, cnt, a, b);
        h->launches++;
        cudaError_t e = cudaMemcpyAsync(pin_out, a, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pout_out, b, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
        cudaFree(b);
        if (e != cudaSuccess) return sfail(VRG_ERR_CUDA, std::string("strict mode: band sums: ") + cudaGetErrorString(e));
    }
    SCK(cudaStreamSynchronize(h->st));
    if (vox_out) for (long long i = 0; i < cnt; ++i) vox_out[i] = (int64_t)v[(size_t)i];
    return VRG_OK;
}

int vrg_strict_get_trace(vrg_strict *h, int64_t *rows_out, int64_t cap_rows, int64_t *n_rows) {
    if (!h || !n_rows) return sfail(VRG_ERR_ARG, "strict mode: null argument");
    const int64_t rows = (int64_t)h->trace.size() / 3;
    *n_rows = rows;
    if (rows_out) {
        if (rows > cap_rows) return sfail(VRG_ERR_ARG, "strict mode: trace buffer too small");
        memcpy(rows_out, h->trace.data(), h->trace.size() * sizeof(int64_t));
    }
    return VRG_OK;
}

// the running sums innerProb / outerProb as they stand (unnormalised, VRG:132-133), whole volume
int vrg_strict_get_sums(vrg_strict *h, double *pin_out, double *pout_out) {
    if (!h || !h->inited || !pin_out || !pout_out) return sfail(VRG_ERR_ARG, "strict mode: bad argument");
    SCK(cudaSetDevice(h->device));
    SCK(cudaMemcpyAsync(pin_out, h->p.pin, (size_t)h->p.N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SCK(cudaMemcpyAsync(pout_out, h->p.pout, (size_t)h->p.N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SCK(cudaStreamSynchronize(h->st));
    return VRG_OK;
}

}  // extern "C"
