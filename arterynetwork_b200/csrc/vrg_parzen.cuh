// Continuous-intensity mode: the brute-force Parzen path for volumes whose intensities do
// not collapse to <= 65536 levels, i.e. the reference's own arithmetic (VRG:151-155, 232-255) without a table.
//
// Every band voxel carries its two unnormalised Parzen sums (pin, pout), exactly as innerProb / outerProb do in the
// reference.  Per iteration:
//   k_cont_decide  flip flags from the stored sums: band & ((pin/n_in >= pout/n_out) ^ S)                  VRG:79-88
//   k_cancel<CONT> cancel rule + in-place flips + region sizes (shared with the table modes)               VRG:165-230
//   k_cont_rows    the rows that flipped, in canonical order (single-block compaction)
//   k_cont_band    the new band B' from S'; voxels that stay in the band / enter it are listed
//   k_cont_incr    staying voxels: pin += sum_A K - sum_R K, pout -= the same  (A entered, R left)          VRG:232-247
//   k_cont_full    entering voxels: both sums over the whole volume, O(n_new * N) fp64 exp                   VRG:249-255
// Fixed traversal orders everywhere (canonical row order for the corrections, a fixed volume partition and a fixed
// reduction tree for the full sums), so results do not depend on scheduling.  Single slab.
// Label 4: excluded voxels belong to neither region; those absorbed in an iteration (k_absorb<CONT> records them in the
// plane AB) join the outside region, i.e. every staying band voxel's pout gains their kernel terms (VRG:235,247).
#pragma once
#include "vrg_kernels.cuh"

namespace vrg {

enum { CC_NEW = 0, CC_OLD = 1, CC_ROWS = 2, CC_OVERFLOW = 3, CC_ABROWS = 4, CC_WORDS = 8 };
constexpr int CONT_TILE = 16;    // new voxels whose sums one block accumulates while streaming the volume
constexpr int CONT_SPLIT = 32;   // volume partitions per tile (second-stage sum in partition order): a few hundred entering voxels
                                 // are 10-20 tiles, and 20 x 8 blocks left most of the 148 SMs with one block of 8 warps (33 % of the
                                 // fp64 exp rate); 32 partitions fill them
constexpr long long EXIT_CONT_OVERFLOW = 98;

struct Cont {
    double *pin, *pout;      // [local voxels] unnormalised Parzen sums, meaningful where B is set
    uint32_t *B;             // band plane the sums are valid for
    int *newlist, *oldlist;  // local voxel ids of band voxels that entered / stayed
    int *rowlist;            // row-segment ids whose flip word is non-zero, ascending
    uint32_t *AB;            // voxels absorbed (label 4 -> 3) in this iteration (only with label 4 in the input)
    int *abrowlist;          // row-segment ids that hold an absorbed voxel, ascending
    int *count;              // [CC_WORDS]
    int cap;                 // capacity of the voxel lists
    double *partial;         // [tiles][CONT_SPLIT][CONT_TILE][2]
};

// resets of the per-iteration counters (a separate launch: the next kernels add to them from every block)
__global__ void k_cont_begin(Params p, Cont q) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || p.ctrl[C_STATUS] != RUNNING) return;
    if (q.count[CC_OVERFLOW]) { p.ctrl[C_STATUS] = EXIT_CONT_OVERFLOW; return; }  // a band list did not fit
    q.count[CC_NEW] = 0;  // consumed by the previous iteration's k_cont_incr / k_cont_full
    q.count[CC_OLD] = 0;
    p.ctrl[C_APPLY] = p.gstats[ST_N_IN] < p.ctrl[C_MAX_SEG];  // L == 0 in this mode: the counters start the vector
    p.lstats[ST_N_FLIPS] = 0;
    front_list(p, (int)(p.ctrl[C_SWEEPS] & 1))[0] = 0;
    dirty_list(p, (int)((p.ctrl[C_SWEEPS] + 1) & 1))[0] = 0;
}

// region sizes at init (VRG:49-52); excluded voxels (label 4) belong to neither region
__global__ void __launch_bounds__(BLOCK) k_cont_count(Params p) {
    const int lane = threadIdx.x & 31;
    const int nrows = (p.own_hi - p.own_lo) * p.Y, nwarps = gridDim.x * WARPS;
    long long n_in = 0, n_all = 0, n_ex = 0;
    for (int r = blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const long long wbase = (long long)(p.own_lo + r / p.Y) * p.plane_words + (long long)(r % p.Y) * p.WP;
        for (int c = lane; c < p.XW; c += 32) {
            const uint32_t s = p.S[wbase + c];
            n_in += __popc(s);
            n_all += __popc(valid_mask(p, c));
            if (p.E) n_ex += __popc(p.E[wbase + c] & ~s & valid_mask(p, c));
        }
    }
    n_in = warp_sum(n_in); n_all = warp_sum(n_all); n_ex = warp_sum(n_ex);
    if (lane == 0) {
        atomicAdd((unsigned long long *)&p.lstats[ST_N_IN], (unsigned long long)n_in);
        atomicAdd((unsigned long long *)&p.lstats[ST_N_OUT], (unsigned long long)(n_all - n_in - n_ex));
        if (n_ex) atomicAdd((unsigned long long *)&p.lstats[ST_N_EXCL], (unsigned long long)n_ex);
    }
}

// ascending list of the row segments that satisfy a predicate (one block; the row space of this mode is small):
// which = 0: the flip word is non-zero (row flag)  -> rowlist / CC_ROWS
// which = 1: the row holds an absorbed voxel (AB)   -> abrowlist / CC_ABROWS
__global__ void __launch_bounds__(1024) k_cont_rows(Params p, Cont q, int which) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = (p.own_hi - p.own_lo) * p.Y * p.nseg, base = p.own_lo * p.Y * p.nseg;
    int *list = which ? q.abrowlist : q.rowlist;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < nrows; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        bool flag = false;
        if (i < nrows) {
            if (!which) flag = p.rowflag[base + i] != 0;
            else {
                const int rr = base + i, sg = rr % p.nseg, t = rr / p.nseg;
                const uint32_t *row = q.AB + (long long)(t / p.Y) * p.plane_words + (long long)(t % p.Y) * p.WP;
                for (int c = sg * p.segw; c < min((sg + 1) * p.segw, p.XW); ++c) flag |= row[c] != 0u;
            }
        }
        const unsigned m = __ballot_sync(FULL, flag);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (flag) list[s_base + before + __popc(m & ((1u << lane) - 1u))] = base + i;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; ++w) t += s_warp[w]; s_base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) q.count[which ? CC_ABROWS : CC_ROWS] = s_base;
}

// appends the voxel ids of the set bits of `bits` (one word per lane) to a list, warp-aggregated
__device__ __forceinline__ void cont_append(int *list, int *counter, int cap, int *overflow, uint32_t bits, long long vox0, int lane) {
    const int n = __popc(bits);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, total);
    base = __shfl_sync(FULL, base, 0);
    if (base + total > cap) { if (lane == 0) *overflow = 1; return; }
    int at = base + incl - n;
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        list[at++] = (int)(vox0 + b);
    }
}

// the band after the flips: B is rewritten in place, entering voxels go to newlist, staying ones to oldlist
__global__ void __launch_bounds__(BLOCK) k_cont_band(Params p, Cont q) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const int lane = threadIdx.x & 31;
    const int nyb = (p.Y + ROWS_PER_UNIT - 1) / ROWS_PER_UNIT;
    const long long nunits = (long long)(p.own_hi - p.own_lo) * nyb * p.nseg;
    const long long nwarps = (long long)gridDim.x * WARPS;
    Strip st;
    for (long long u = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); u < nunits; u += nwarps) {
        const Unit un = decode_unit(p, u, p.own_lo, nyb);
        if (!unit_near_segmented(p, un.zl, un.y0, un.sg, lane)) continue;
        const int c = un.sg * p.segw - 1 + lane;
        st.begin(p, un.zl, un.y0, c, lane);
        for (int y = un.y0; y < un.y1; ++y) {
            uint32_t s, inner, outer;
            st.step(p, y, s, inner, outer);
            const long long widx = (long long)un.zl * p.plane_words + (long long)y * p.WP + c;
            if (p.E != nullptr && outer) outer &= ~p.E[widx];  // outer is non-zero on active lanes only
            const uint32_t band = inner | outer;
            const uint32_t bold = st.active ? q.B[widx] : 0u;
            if (__ballot_sync(FULL, (band | bold) != 0u) == 0u) continue;
            if (st.active && band != bold) q.B[widx] = band;
            const long long vox0 = (long long)un.zl * p.plane_vox + (long long)y * p.X + (long long)c * 32;
            cont_append(q.newlist, q.count + CC_NEW, q.cap, q.count + CC_OVERFLOW, band & ~bold, vox0, lane);
            cont_append(q.oldlist, q.count + CC_OLD, q.cap, q.count + CC_OVERFLOW, band & bold, vox0, lane);
        }
    }
}

__device__ __forceinline__ double parzen(const Params &p, double a, double b) {
    const double d = a - b;
    return 0.3989422804014327 * exp(p.mhH * (d * d));  // A * exp(-0.5 * H * d^2), VRG:7,154
}

// corrections of the voxels that stay in the band (VRG:232-247): the flipped voxels are walked in canonical order
// (ascending rows, ascending words, ascending bits), so every thread adds the same terms in the same order
__global__ void __launch_bounds__(BLOCK) k_cont_incr(Params p, Cont q) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const int n_old = min(q.count[CC_OLD], q.cap), nrows = q.count[CC_ROWS];
    for (int i = blockIdx.x * BLOCK + threadIdx.x; i < n_old; i += gridDim.x * BLOCK) {
        const int vox = q.oldlist[i];
        const double v = p.data[vox];
        double sum_a = 0.0, sum_r = 0.0;
        unsigned long long nev = 0;
        for (int k = 0; k < nrows; ++k) {
            const int rr = q.rowlist[k];
            const int sg = rr % p.nseg, t = rr / p.nseg, y = t % p.Y, zl = t / p.Y;
            const long long wrow = (long long)zl * p.plane_words + (long long)y * p.WP;
            const long long vrow = (long long)zl * p.plane_vox + (long long)y * p.X;
            for (int c = sg * p.segw; c < min((sg + 1) * p.segw, p.XW); ++c) {
                uint32_t f = p.F[wrow + c];  // executed flips (k_cancel rewrote the word)
                if (!f) continue;
                const uint32_t snew = p.S[wrow + c];
                nev += __popc(f);
                while (f) {
                    const int b = __ffs(f) - 1;
                    f &= f - 1;
                    const double kv = parzen(p, p.data[vrow + (long long)c * 32 + b], v);
                    if ((snew >> b) & 1u) sum_a += kv; else sum_r += kv;
                }
            }
        }
        double sum_ab = 0.0;  // voxels absorbed into the outside region (label 4 -> 3), same canonical order
        const int nab = p.E != nullptr ? q.count[CC_ABROWS] : 0;
        for (int k = 0; k < nab; ++k) {
            const int rr = q.abrowlist[k];
            const int sg = rr % p.nseg, t = rr / p.nseg, y = t % p.Y, zl = t / p.Y;
            const long long wrow = (long long)zl * p.plane_words + (long long)y * p.WP;
            const long long vrow = (long long)zl * p.plane_vox + (long long)y * p.X;
            for (int c = sg * p.segw; c < min((sg + 1) * p.segw, p.XW); ++c) {
                uint32_t ab = q.AB[wrow + c];
                nev += __popc(ab);
                while (ab) {
                    const int b = __ffs(ab) - 1;
                    ab &= ab - 1;
                    sum_ab += parzen(p, p.data[vrow + (long long)c * 32 + b], v);
                }
            }
        }
        q.pin[vox] = q.pin[vox] + sum_a - sum_r;            // innerProb += innerCorrection; innerProb -= outerCorrection
        q.pout[vox] = q.pout[vox] - sum_a + sum_r + sum_ab;  // outerProb -= inner...; += outer...; += addedCorrection
        if (nev) atomicAdd((unsigned long long *)&p.lstats[ST_EXP_EVALS], nev);
    }
}

// full sums of the voxels that entered the band (VRG:151-155, 249-255): block (tile, part) streams its part of the
// volume once and accumulates the sums of CONT_TILE voxels; stage 2 adds the parts in order
__global__ void __launch_bounds__(BLOCK) k_cont_full1(Params p, Cont q) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const int n_new = min(q.count[CC_NEW], q.cap);
    const int ntiles = (n_new + CONT_TILE - 1) / CONT_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part = blockIdx.y;
    const long long nwords = (long long)(p.own_hi - p.own_lo) * p.Y * p.XW;
    const long long w0 = nwords * part / CONT_SPLIT, w1 = nwords * (part + 1) / CONT_SPLIT;
    __shared__ double s_red[WARPS][CONT_TILE][2];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double vt[CONT_TILE], ain[CONT_TILE], aout[CONT_TILE];
#pragma unroll
        for (int k = 0; k < CONT_TILE; ++k) {
            const int j = tile * CONT_TILE + k;
            vt[k] = p.data[q.newlist[min(j, n_new - 1)]];
            ain[k] = 0.0; aout[k] = 0.0;
        }
        for (long long w = w0 + warp; w < w1; w += WARPS) {
            const int c = (int)(w % p.XW);
            const long long t = w / p.XW;
            const int y = (int)(t % p.Y), zl = p.own_lo + (int)(t / p.Y);
            const int x = c * 32 + lane;
            if (x >= p.X) continue;
            const long long wi = (long long)zl * p.plane_words + (long long)y * p.WP + c;
            const uint32_t s = p.S[wi], e = p.E ? p.E[wi] : 0u;
            const double val = p.data[(long long)zl * p.plane_vox + (long long)y * p.X + x];
            const bool in = (s >> lane) & 1u;
            if (!in && ((e >> lane) & 1u)) continue;  // excluded: in neither region
#pragma unroll
            for (int k = 0; k < CONT_TILE; ++k) {
                const double kv = parzen(p, val, vt[k]);
                if (in) ain[k] += kv; else aout[k] += kv;
            }
        }
#pragma unroll
        for (int k = 0; k < CONT_TILE; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ain[k] += __shfl_xor_sync(FULL, ain[k], o);
                aout[k] += __shfl_xor_sync(FULL, aout[k], o);
            }
            if (lane == 0) { s_red[warp][k][0] = ain[k]; s_red[warp][k][1] = aout[k]; }
        }
        __syncthreads();
        if (threadIdx.x < CONT_TILE * 2) {
            const int k = threadIdx.x >> 1, io = threadIdx.x & 1;
            double t = 0.0;
            for (int w = 0; w < WARPS; ++w) t += s_red[w][k][io];
            q.partial[(((size_t)tile * CONT_SPLIT + part) * CONT_TILE + k) * 2 + io] = t;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BLOCK) k_cont_full2(Params p, Cont q) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const int n_new = min(q.count[CC_NEW], q.cap);
    if (blockIdx.x == 0 && threadIdx.x == 0)  // every entering voxel met every voxel of the two regions once
        atomicAdd((unsigned long long *)&p.lstats[ST_EXP_EVALS],
                  (unsigned long long)n_new * (unsigned long long)(p.gstats[ST_N_IN] + p.gstats[ST_N_OUT]));
    for (int j = blockIdx.x * BLOCK + threadIdx.x; j < n_new; j += gridDim.x * BLOCK) {
        const int tile = j / CONT_TILE, k = j % CONT_TILE;
        double si = 0.0, so = 0.0;
        for (int part = 0; part < CONT_SPLIT; ++part) {
            const size_t at = (((size_t)tile * CONT_SPLIT + part) * CONT_TILE + k) * 2;
            si += q.partial[at];
            so += q.partial[at + 1];
        }
        const int vox = q.newlist[j];
        q.pin[vox] = si;
        q.pout[vox] = so;
    }
}

// flip flags from the stored sums (VRG:79-88): a band voxel is inside iff pin/n_in >= pout/n_out (ties inside)
__global__ void __launch_bounds__(BLOCK) k_cont_decide(Params p, Cont q) {
    if (p.ctrl[C_STATUS] != RUNNING) return;
    const int lane = threadIdx.x & 31;
    const double n_in = (double)p.gstats[ST_N_IN], n_out = (double)p.gstats[ST_N_OUT];
    const int nrows = (p.own_hi - p.own_lo) * p.Y * p.nseg, nwarps = gridDim.x * WARPS;
    long long flips = 0;
    for (int r = blockIdx.x * WARPS + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const int sg = r % p.nseg, t = r / p.nseg, y = t % p.Y, zl = p.own_lo + t / p.Y;
        const int c0 = sg * p.segw - 1, c = c0 + lane;
        const bool active = c >= 0 && c < p.XW && lane >= 1 && lane <= p.segw;
        const long long widx = (long long)zl * p.plane_words + (long long)y * p.WP + c;
        const long long ridx = ((long long)zl * p.Y + y) * p.nseg + sg;
        const uint8_t was = p.rowflag[ridx];
        const uint32_t band = active ? q.B[widx] : 0u;
        unsigned m = __ballot_sync(FULL, band != 0u);
        if (m == 0u && !was) continue;
        const uint32_t s = active ? p.S[widx] : 0u;
        const long long rowvox = (long long)zl * p.plane_vox + (long long)y * p.X + lane;
        uint32_t D = 0;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const int x = (c0 + j) * 32 + lane;
            bool bit = false;
            if (x < p.X) {
                const long long vox = rowvox + (long long)(c0 + j) * 32;
                bit = q.pin[vox] / n_in >= q.pout[vox] / n_out;
            }
            const unsigned word = __ballot_sync(FULL, bit);
            if (lane == j) D = word;
        }
        const uint32_t f = band & (D ^ s);
        store_flips(p, widx, ridx, was, f, active, true, lane);
        flips += __popc(f);
    }
    flips = warp_sum(flips);
    if (lane == 0 && flips) atomicAdd((unsigned long long *)&p.lstats[ST_N_FLIPS], (unsigned long long)flips);
}

// fp64 exp rate of the device with nothing else in the way: the roofline of the continuous mode (independent chains, the same
// expression as parzen()).  out[thread] keeps the compiler honest.
__global__ void __launch_bounds__(BLOCK) k_exp_peak(double mhH, int iters, double *out) {
    const double a = (double)threadIdx.x * 1e-3, b = (double)blockIdx.x * 1e-6;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double d = a - (b + (double)(i * 8 + k) * 1e-7);
            acc[k] += 0.3989422804014327 * exp(mhH * (d * d));
        }
    }
    double t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += acc[k];
    out[(size_t)blockIdx.x * BLOCK + threadIdx.x] = t;
}

}  // namespace vrg
