// Device code of the B200-native variational region growing path (sm_100a).
//
// State is bit-packed along x: one uint32 word = 32 voxels of one row.  Three
// persistent bit-planes (segmented S ping/pong, excluded E) plus two per-iteration
// flag planes (R = inner-band voxels that leave, A0 = outer-band voxels that want
// to enter).  The reference's label alphabet (VRG:21) is a function of (S, E) and
// the 26-neighbourhood, so it is never stored; it is materialised on download.
//
// Per iteration (reference: Code/variationalRegionGrowing.py, VRG:line):
//   k_table   region histograms -> normalised Parzen sums per level -> decision bit   VRG:79-87,151-155
//   k_decide  bands from S (26-neighbourhood), decision per voxel, R / A0 planes       VRG:87-88,139-145
//   k_apply   cancel rule, S' = (S & ~R) | A, integer histogram deltas                 VRG:165-233
//   k_absorb  label 4 -> 3 around flips (only when the input holds label 4)            VRG:167-168,177-179
//   k_advance exit tests and trace row                                                 VRG:91-117
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrg {

constexpr int HALO = 2;
constexpr int WORDS_PER_WARP = 30;  // a warp covers 30 output words + 1 halo word on each side
constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;

enum { ST_N_IN = 0, ST_N_OUT = 1, ST_N_EXCL = 2, ST_N_FLIPS = 3, ST_N_BAND = 4, ST_BAD_LABEL = 5, ST_NONFINITE = 6, ST_EXTRA = 8 };
enum { C_STATUS = 0, C_ITER = 1, C_ITER_MAX = 2, C_MAX_SEG = 3, C_APPLY = 4, C_APPLIED = 5, C_TRACE_N = 6, C_SWEEPS = 7, C_WORDS = 16 };
constexpr long long RUNNING = -1;

enum { MODE_F64_DENSE = 0, MODE_F64_BAND = 1, MODE_INDEX = 2 };

struct Params {
    // geometry
    int Y, X, XW, WP, nseg;          // rows, voxels per row, words per row, word pitch, warp segments per row
    int nzl;                          // local planes incl. halos
    int own_lo, own_hi;               // local plane range owned
    int valid_lo, valid_hi;           // local planes that lie inside the global volume
    uint32_t tail_mask;               // valid bits of word XW-1
    long long plane_words;            // Y * WP
    long long plane_vox;              // Y * X
    // state
    uint32_t *seg[2];
    uint32_t *excl;                   // nullptr when the run never had label 4
    uint32_t *R, *A0;
    const double *data;               // fp64 intensities, local planes
    const uint16_t *index;            // level index volume (MODE_INDEX)
    // levels / table
    int L;                            // table size
    int LW;                           // ceil(L / 32)
    int lattice;                      // 1: index = rint((v - lev0) * inv_step)
    double lev0, inv_step;
    const double *levels;             // [L]
    uint32_t *dbits;                  // [LW]
    double *pin, *pout;               // [L] normalised Parzen sums of the last table
    double mhH;                       // -0.5 * H
    long long *lstats;                // local  [2L + ST_EXTRA]
    const long long *gstats;          // global [2L + ST_EXTRA] (aliases lstats on one GPU)
    long long *ctrl;                  // [C_WORDS]
    long long *trace;                 // [3 * (iter_max + 2)]
};

__device__ __forceinline__ uint32_t valid_mask(const Params &p, int c) {
    return c < p.XW - 1 ? 0xFFFFFFFFu : (c == p.XW - 1 ? p.tail_mask : 0u);
}

// x-dilation by one voxel of a row of words held one-per-lane
__device__ __forceinline__ uint32_t dilate_x1(uint32_t v) {
    uint32_t l = __shfl_up_sync(0xFFFFFFFFu, v, 1), r = __shfl_down_sync(0xFFFFFFFFu, v, 1);
    return v | (v << 1) | (v >> 1) | (l >> 31) | (r << 31);
}
__device__ __forceinline__ uint32_t dilate_x2(uint32_t v) {
    uint32_t l = __shfl_up_sync(0xFFFFFFFFu, v, 1), r = __shfl_down_sync(0xFFFFFFFFu, v, 1);
    return v | (v << 1) | (v >> 1) | (v << 2) | (v >> 2) | (l >> 31) | (l >> 30) | (r << 31) | (r << 30);
}

__device__ __forceinline__ int level_of(const Params &p, double v) {
    if (p.lattice) {
        // round-to-nearest via the 2^52+2^51 trick: the integer lands in the low word
        double t = __fma_rn(v - p.lev0, p.inv_step, 6755399441055744.0);
        return __double2loint(t);
    }
    int lo = 0, hi = p.L - 1;
    v += 0.0;
    while (lo < hi) {
        int m = (lo + hi) >> 1;
        if (p.levels[m] < v) lo = m + 1; else hi = m;
    }
    return lo;
}

template <int MODE>
__device__ __forceinline__ int level_at(const Params &p, long long vox) {
    if (MODE == MODE_INDEX) return p.index[vox];
    return level_of(p, p.data[vox]);
}

struct RowSeg { int zl, y, c; bool inrange; };
__device__ __forceinline__ RowSeg decode_row(const Params &p, long long r, int zbase, int lane) {
    RowSeg s;
    int sg = (int)(r % p.nseg);
    long long t = r / p.nseg;
    s.y = (int)(t % p.Y);
    s.zl = zbase + (int)(t / p.Y);
    s.c = sg * WORDS_PER_WARP - 1 + lane;
    s.inrange = s.c >= 0 && s.c < p.XW;
    return s;
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// k_table: one block per 32 levels; a warp sums one level's two Parzen sums over all levels.
// Fixed order: lane-strided partial sums, then an xor-shuffle tree -> deterministic.
__global__ void __launch_bounds__(BLOCK) k_table(Params p) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const long long *g = p.gstats;
    const long long n_in = g[2 * p.L + ST_N_IN], n_out = g[2 * p.L + ST_N_OUT];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.ctrl[C_APPLY] = n_in < p.ctrl[C_MAX_SEG];  // cap is tested before the flips are applied, VRG:101
        p.lstats[2 * p.L + ST_N_FLIPS] = 0;
    }
    __shared__ uint32_t s_bits[WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t mybits = 0;
    for (int k = 0; k < 32 / WARPS; ++k) {
        int b = blockIdx.x * 32 + warp * (32 / WARPS) + k;
        if (b >= p.L) break;
        if (g[b] + g[p.L + b] == 0) continue;  // level absent from both regions: never looked up
        const double lb = p.levels[b];
        double si = 0.0, so = 0.0;
        for (int c = lane; c < p.L; c += 32) {
            long long hi = g[c], ho = g[p.L + c];
            if ((hi | ho) == 0) continue;
            double diff = p.levels[c] - lb;
            double kv = 0.3989422804014327 * exp(p.mhH * (diff * diff));  // A * exp(-0.5*H*d^2), VRG:7,154
            si += (double)hi * kv;
            so += (double)ho * kv;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            si += __shfl_xor_sync(0xFFFFFFFFu, si, o);
            so += __shfl_xor_sync(0xFFFFFFFFu, so, o);
        }
        double pi = si / (double)n_in, po = so / (double)n_out;  // VRG:81-82
        if (lane == 0) { p.pin[b] = pi; p.pout[b] = po; }
        if (pi >= po) mybits |= 1u << (warp * (32 / WARPS) + k);  // ties go inside, VRG:87
    }
    if (lane == 0) s_bits[warp] = mybits;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t w = 0;
        for (int i = 0; i < WARPS; ++i) w |= s_bits[i];
        p.dbits[blockIdx.x] = w;
    }
}

// ---------------------------------------------------------------------------------------------
// k_decide: bands from the segmented plane, decision bit per voxel, flip flag planes.
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_decide(Params p) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    extern __shared__ uint32_t s_dbits[];
    for (int i = threadIdx.x; i < p.LW; i += BLOCK) s_dbits[i] = p.dbits[i];
    __syncthreads();
    const int par = (int)(p.ctrl[C_APPLIED] & 1);
    const uint32_t *__restrict__ S = p.seg[par];
    const uint32_t *__restrict__ E = p.excl;
    const int lane = threadIdx.x & 31;
    const int zlo = max(p.valid_lo, p.own_lo - 1), zhi = min(p.valid_hi, p.own_hi + 1);
    const long long nrows = (long long)(zhi - zlo) * p.Y * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    long long flips = 0;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const RowSeg rs = decode_row(p, r, zlo, lane);
        const uint32_t vm = rs.inrange ? valid_mask(p, rs.c) : 0u;
        uint32_t vs = 0, vn = 0, s = 0;
        if (rs.inrange) {
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = rs.zl + dz;
                if (zz < p.valid_lo || zz >= p.valid_hi) continue;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = rs.y + dy;
                    if (yy < 0 || yy >= p.Y) continue;
                    const uint32_t w = S[(long long)zz * p.plane_words + (long long)yy * p.WP + rs.c];
                    vs |= w;
                    vn |= ~w & vm;
                    if (dz == 0 && dy == 0) s = w;
                }
            }
        }
        const long long widx = (long long)rs.zl * p.plane_words + (long long)rs.y * p.WP + rs.c;
        const uint32_t e = (E != nullptr && rs.inrange) ? E[widx] : 0u;
        const uint32_t dil_s = dilate_x1(vs), dil_n = dilate_x1(vn);
        const bool active = rs.inrange && lane >= 1 && lane <= WORDS_PER_WARP;
        const uint32_t innerB = active ? (s & dil_n) : 0u;             // segmented with an unsegmented in-bounds neighbour
        const uint32_t outerB = active ? (~s & vm & ~e & dil_s) : 0u;  // unsegmented, not excluded, segmented neighbour
        const uint32_t band = innerB | outerB;
        uint32_t D = 0;
        const long long rowvox = (long long)rs.zl * p.plane_vox + (long long)rs.y * p.X;
        const int c0 = rs.c - lane;  // column of lane 0
        constexpr int U = 6;
#pragma unroll 1
        for (int j0 = 1; j0 <= WORDS_PER_WARP; j0 += U) {
            if (c0 + j0 >= p.XW) break;
            int lev[U];
            bool on[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + u, cj = c0 + j;
                const uint32_t bw = __shfl_sync(0xFFFFFFFFu, band, j);
                const int x = cj * 32 + lane;
                const bool wordon = cj < p.XW && (MODE == MODE_F64_DENSE || bw != 0u);
                on[u] = wordon;
                lev[u] = -1;
                if (wordon && x < p.X) lev[u] = level_at<MODE>(p, rowvox + x);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!on[u]) continue;  // warp-uniform
                const int l = lev[u];
                const uint32_t bit = l >= 0 ? (s_dbits[l >> 5] >> (l & 31)) & 1u : 0u;
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
                if (lane == j0 + u) D = word;
            }
        }
        const uint32_t Rw = innerB & ~D;  // inner band leaves iff in < out
        const uint32_t Aw = outerB & D;   // outer band enters iff in >= out
        if (active) { p.R[widx] = Rw; p.A0[widx] = Aw; }
        if (rs.zl >= p.own_lo && rs.zl < p.own_hi) flips += __popc(Rw) + __popc(Aw);
    }
    flips = warp_sum(flips);
    if (lane == 0 && flips) atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_FLIPS], (unsigned long long)flips);
}

// ---------------------------------------------------------------------------------------------
// k_apply: cancel rule + new segmented plane + integer statistics.
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_apply(Params p) {
    if (p.ctrl[C_STATUS] != RUNNING || !p.ctrl[C_APPLY]) return;
    const int par = (int)(p.ctrl[C_APPLIED] & 1);
    const uint32_t *__restrict__ S = p.seg[par];
    uint32_t *__restrict__ S2 = p.seg[par ^ 1];
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(p.own_hi - p.own_lo) * p.Y * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    long long d_in = 0;
    unsigned long long *hin = (unsigned long long *)p.lstats, *hout = hin + p.L;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const RowSeg rs = decode_row(p, r, p.own_lo, lane);
        uint32_t keepv = 0, s = 0, rw = 0;
        if (rs.inrange) {
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = rs.zl + dz;
                if (zz < p.valid_lo || zz >= p.valid_hi) continue;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = rs.y + dy;
                    if (yy < 0 || yy >= p.Y) continue;
                    const long long i = (long long)zz * p.plane_words + (long long)yy * p.WP + rs.c;
                    const uint32_t w = S[i], rr = p.R[i];
                    keepv |= w & ~rr;
                    if (dz == 0 && dy == 0) { s = w; rw = rr; }
                }
            }
        }
        const long long widx = (long long)rs.zl * p.plane_words + (long long)rs.y * p.WP + rs.c;
        const bool active = rs.inrange && lane >= 1 && lane <= WORDS_PER_WARP;
        const uint32_t a0 = active ? p.A0[widx] : 0u;
        const uint32_t dilk = dilate_x1(keepv);
        const uint32_t a = a0 & dilk;  // an addition needs a segmented neighbour that stays (VRG:183-190 then 198)
        if (!active) continue;
        S2[widx] = (s & ~rw) | a;
        d_in += __popc(a) - __popc(rw);
        const long long rowvox = (long long)rs.zl * p.plane_vox + (long long)rs.y * p.X + (long long)rs.c * 32;
        uint32_t m = rw;
        while (m) {  // leaves the inside region: histogram deltas replace VRG:232-247
            const int b = __ffs(m) - 1; m &= m - 1;
            const int l = level_at<MODE>(p, rowvox + b);
            atomicAdd(&hin[l], ~0ull);
            atomicAdd(&hout[l], 1ull);
        }
        m = a;
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            const int l = level_at<MODE>(p, rowvox + b);
            atomicAdd(&hin[l], 1ull);
            atomicAdd(&hout[l], ~0ull);
        }
    }
    d_in = warp_sum(d_in);
    if (lane == 0 && d_in) {
        atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_IN], (unsigned long long)d_in);
        atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_OUT], (unsigned long long)(-d_in));
    }
}

// ---------------------------------------------------------------------------------------------
// k_absorb: excluded voxels within 1 of any listed flip, or within 2 of an executed flip, become outside.
// Runs after k_apply (and after the S' halo exchange on multi-GPU); `applied` parity still names the old plane.
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_absorb(Params p) {
    if (p.ctrl[C_STATUS] != RUNNING || !p.ctrl[C_APPLY]) return;
    const int par = (int)(p.ctrl[C_APPLIED] & 1);
    const uint32_t *__restrict__ S = p.seg[par], *__restrict__ S2 = p.seg[par ^ 1];
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(p.own_hi - p.own_lo) * p.Y * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    long long n_abs = 0;
    unsigned long long *hout = (unsigned long long *)p.lstats + p.L;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const RowSeg rs = decode_row(p, r, p.own_lo, lane);
        const long long widx = (long long)rs.zl * p.plane_words + (long long)rs.y * p.WP + rs.c;
        const bool active = rs.inrange && lane >= 1 && lane <= WORDS_PER_WARP;
        const uint32_t e = active ? p.excl[widx] : 0u;
        if (__ballot_sync(0xFFFFFFFFu, e != 0u) == 0u) continue;
        uint32_t ve = 0, vf = 0;
        if (rs.inrange) {
            for (int dz = -2; dz <= 2; ++dz) {
                const int zz = rs.zl + dz;
                if (zz < p.valid_lo || zz >= p.valid_hi) continue;
                for (int dy = -2; dy <= 2; ++dy) {
                    const int yy = rs.y + dy;
                    if (yy < 0 || yy >= p.Y) continue;
                    const long long i = (long long)zz * p.plane_words + (long long)yy * p.WP + rs.c;
                    ve |= S[i] ^ S2[i];  // executed flips
                    if (dz >= -1 && dz <= 1 && dy >= -1 && dy <= 1) vf |= p.R[i] | p.A0[i];  // every listed flip
                }
            }
        }
        const uint32_t hit = dilate_x2(ve) | dilate_x1(vf);
        uint32_t ab = e & hit;
        if (!active || ab == 0u) continue;
        p.excl[widx] = e & ~ab;
        n_abs += __popc(ab);
        const long long rowvox = (long long)rs.zl * p.plane_vox + (long long)rs.y * p.X + (long long)rs.c * 32;
        while (ab) {
            const int b = __ffs(ab) - 1; ab &= ab - 1;
            atomicAdd(&hout[level_at<MODE>(p, rowvox + b)], 1ull);  // addedPoints, VRG:235,247
        }
    }
    n_abs = warp_sum(n_abs);
    if (lane == 0 && n_abs) {
        atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_OUT], (unsigned long long)n_abs);
        atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_EXCL], (unsigned long long)(-n_abs));
    }
}

// ---------------------------------------------------------------------------------------------
// k_advance: the exit tests of VRG:91-104 and the loop bookkeeping of VRG:113-117.
__global__ void k_advance(Params p) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long *c = p.ctrl;
    if (c[C_STATUS] != RUNNING) return;
    const long long *g = p.gstats + 2 * p.L;
    c[C_SWEEPS] += 1;
    if (g[ST_N_FLIPS] == 0) { c[C_STATUS] = 0; return; }   // converged, VRG:91
    if (!c[C_APPLY]) { c[C_STATUS] = 2; return; }          // max segment size, VRG:101
    const long long t = c[C_TRACE_N];
    p.trace[3 * t] = g[ST_N_FLIPS];
    p.trace[3 * t + 1] = g[ST_N_IN];
    p.trace[3 * t + 2] = g[ST_N_OUT];
    c[C_TRACE_N] = t + 1;
    c[C_APPLIED] += 1;
    c[C_ITER] += 1;
    if (c[C_ITER] > c[C_ITER_MAX]) c[C_STATUS] = 3;        // VRG:58,118
}

// ---------------------------------------------------------------------------------------------
// init branch of update(), VRG:129-145
__global__ void __launch_bounds__(BLOCK) k_init_planes(Params p, const uint8_t *__restrict__ vm, uint32_t *eraw) {
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(p.valid_hi - p.valid_lo) * p.Y * p.XW;
    const long long nwarps = (long long)gridDim.x * WARPS;
    bool bad = false;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const int c = (int)(r % p.XW);
        const long long t = r / p.XW;
        const int y = (int)(t % p.Y), zl = p.valid_lo + (int)(t / p.Y);
        const int x = c * 32 + lane;
        uint8_t v = 3;
        if (x < p.X) v = vm[(long long)zl * p.plane_vox + (long long)y * p.X + x];
        bad |= !(v == 0 || v == 3 || v == 4);
        const uint32_t s = __ballot_sync(0xFFFFFFFFu, v == 0), e = __ballot_sync(0xFFFFFFFFu, v == 4);
        if (lane == 0) {
            const long long widx = (long long)zl * p.plane_words + (long long)y * p.WP + c;
            p.seg[0][widx] = s;
            if (eraw) eraw[widx] = e;
        }
    }
    if (bad) p.lstats[2 * p.L + ST_BAD_LABEL] = 1;
}

// E = Eraw & ~dil3(S) (VRG:137), and the initial band count.  In place on p.excl (only the centre word is read).
__global__ void __launch_bounds__(BLOCK) k_init_bands(Params p) {
    const uint32_t *__restrict__ S = p.seg[0];
    const int lane = threadIdx.x & 31;
    const int zlo = max(p.valid_lo, p.own_lo - 1), zhi = min(p.valid_hi, p.own_hi + 1);
    const long long nrows = (long long)(zhi - zlo) * p.Y * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    long long nband = 0;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const RowSeg rs = decode_row(p, r, zlo, lane);
        const uint32_t vm = rs.inrange ? valid_mask(p, rs.c) : 0u;
        uint32_t vs = 0, vn = 0, s = 0;
        if (rs.inrange) {
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = rs.zl + dz;
                if (zz < p.valid_lo || zz >= p.valid_hi) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = rs.y + dy;
                    if (yy < 0 || yy >= p.Y) continue;
                    const uint32_t w = S[(long long)zz * p.plane_words + (long long)yy * p.WP + rs.c];
                    vs |= w; vn |= ~w & vm;
                    if (dz == 0 && dy == 0) s = w;
                }
            }
        }
        const uint32_t dil_s = dilate_x1(vs), dil_n = dilate_x1(vn);
        const bool active = rs.inrange && lane >= 1 && lane <= WORDS_PER_WARP;
        if (!active) continue;
        const long long widx = (long long)rs.zl * p.plane_words + (long long)rs.y * p.WP + rs.c;
        uint32_t e = 0;
        if (p.excl) { e = p.excl[widx] & ~dil_s; p.excl[widx] = e; }
        if (rs.zl >= p.own_lo && rs.zl < p.own_hi) nband += __popc(s & dil_n) + __popc(~s & vm & ~e & dil_s);
    }
    nband = warp_sum(nband);
    if (lane == 0 && nband) atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_BAND], (unsigned long long)nband);
}

// region histograms and sizes over own planes (VRG:49-52, 149-150 as integer counts)
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_init_hist(Params p, int use_smem) {
    extern __shared__ unsigned int s_h[];  // [2L] when use_smem
    if (use_smem) {
        for (int i = threadIdx.x; i < 2 * p.L; i += BLOCK) s_h[i] = 0;
        __syncthreads();
    }
    const uint32_t *__restrict__ S = p.seg[0];
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(p.own_hi - p.own_lo) * p.Y * p.XW;
    const long long nwarps = (long long)gridDim.x * WARPS;
    unsigned long long *hin = (unsigned long long *)p.lstats, *hout = hin + p.L;
    long long n_in = 0, n_out = 0, n_ex = 0;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const int c = (int)(r % p.XW);
        const long long t = r / p.XW;
        const int y = (int)(t % p.Y), zl = p.own_lo + (int)(t / p.Y);
        const long long widx = (long long)zl * p.plane_words + (long long)y * p.WP + c;
        const uint32_t s = S[widx], e = p.excl ? p.excl[widx] : 0u;
        const int x = c * 32 + lane;
        if (x >= p.X) continue;
        const int l = level_at<MODE>(p, (long long)zl * p.plane_vox + (long long)y * p.X + x);
        const uint32_t bit = 1u << lane;
        if (s & bit) { n_in++; if (use_smem) atomicAdd(&s_h[l], 1u); else atomicAdd(&hin[l], 1ull); }
        else if (!(e & bit)) { n_out++; if (use_smem) atomicAdd(&s_h[p.L + l], 1u); else atomicAdd(&hout[l], 1ull); }
        else n_ex++;
    }
    n_in = warp_sum(n_in); n_out = warp_sum(n_out); n_ex = warp_sum(n_ex);
    if (lane == 0) {
        if (n_in) atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_IN], (unsigned long long)n_in);
        if (n_out) atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_OUT], (unsigned long long)n_out);
        if (n_ex) atomicAdd((unsigned long long *)&p.lstats[2 * p.L + ST_N_EXCL], (unsigned long long)n_ex);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * p.L; i += BLOCK)
            if (s_h[i]) atomicAdd(&hin[i], (unsigned long long)s_h[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// distinct intensity levels: open-addressing hash set over the fp64 bit patterns
constexpr unsigned long long HEMPTY = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(BLOCK) k_scan_levels(const double *__restrict__ data, long long n, unsigned long long *table,
                                                       int cap_mask, int *count, int max_count, int *flags) {
    unsigned long long last = HEMPTY;
    for (long long i = (long long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * BLOCK) {
        double v = data[i] + 0.0;
        if (!isfinite(v)) { flags[0] = 1; continue; }
        unsigned long long key = (unsigned long long)__double_as_longlong(v);
        if (key == last) continue;
        last = key;
        unsigned int h = (unsigned int)mix64(key) & cap_mask;
        while (true) {
            unsigned long long cur = table[h];
            if (cur == key) break;
            if (cur == HEMPTY) {
                if (*(volatile int *)count >= max_count) { flags[1] = 1; break; }
                unsigned long long old = atomicCAS(&table[h], HEMPTY, key);
                if (old == HEMPTY) { atomicAdd(count, 1); break; }
                if (old == key) break;
            }
            h = (h + 1) & cap_mask;
        }
    }
}

__global__ void __launch_bounds__(BLOCK) k_build_index(Params p, const double *__restrict__ data, uint16_t *index, long long n) {
    for (long long i = (long long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long long)gridDim.x * BLOCK)
        index[i] = (uint16_t)level_of(p, data[i]);
}

// ---------------------------------------------------------------------------------------------
// outputs: canonical labels (VRG:21) or the 0/1 segmented map, one byte per voxel, own planes only
__global__ void __launch_bounds__(BLOCK) k_labels(Params p, uint8_t *__restrict__ out, int seg_only) {
    const int par = (int)(p.ctrl[C_APPLIED] & 1);
    const uint32_t *__restrict__ S = p.seg[par];
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)(p.own_hi - p.own_lo) * p.Y * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const RowSeg rs = decode_row(p, r, p.own_lo, lane);
        const uint32_t vm = rs.inrange ? valid_mask(p, rs.c) : 0u;
        uint32_t vs = 0, vn = 0, s = 0;
        if (rs.inrange) {
            for (int dz = -1; dz <= 1; ++dz) {
                const int zz = rs.zl + dz;
                if (zz < p.valid_lo || zz >= p.valid_hi) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = rs.y + dy;
                    if (yy < 0 || yy >= p.Y) continue;
                    const uint32_t w = S[(long long)zz * p.plane_words + (long long)yy * p.WP + rs.c];
                    vs |= w; vn |= ~w & vm;
                    if (dz == 0 && dy == 0) s = w;
                }
            }
        }
        const long long widx = (long long)rs.zl * p.plane_words + (long long)rs.y * p.WP + rs.c;
        const uint32_t e = (p.excl && rs.inrange) ? p.excl[widx] : 0u;
        const uint32_t dil_s = dilate_x1(vs), dil_n = dilate_x1(vn);
        const uint32_t inner = s & dil_n, outer = ~s & ~e & dil_s;
        const long long rowout = ((long long)(rs.zl - p.own_lo) * p.Y + rs.y) * p.X;
        const int c0 = rs.c - lane;
        for (int j = 1; j <= WORDS_PER_WARP; ++j) {
            const int cj = c0 + j;
            if (cj >= p.XW) break;
            const uint32_t sj = __shfl_sync(0xFFFFFFFFu, s, j), ij = __shfl_sync(0xFFFFFFFFu, inner, j);
            const uint32_t oj = __shfl_sync(0xFFFFFFFFu, outer, j), ej = __shfl_sync(0xFFFFFFFFu, e, j);
            const int x = cj * 32 + lane;
            if (x >= p.X) continue;
            const uint32_t bit = 1u << lane;
            uint8_t lab;
            if (seg_only) lab = (sj & bit) ? 1 : 0;
            else lab = (sj & bit) ? ((ij & bit) ? 1 : 0) : ((ej & bit) ? 4 : ((oj & bit) ? 2 : 3));
            out[rowout + x] = lab;
        }
    }
}

}  // namespace vrg
