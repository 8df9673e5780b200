// Vessel-mask construction on the device: the step before the VRG path.
// Reference: Code/generateVesselVolume.py (GVV:line)
//   GVV:188-192  two cut-offs relative to the range of the vesselness volume, the first one applied only to voxels within
//                `edge_distance` (10) of the brain-mask boundary (EDT of the brain mask, GVV:183)
//   GVV:195-200  binarise, label the 26-connected components (skimage.measure.label, connectivity=3, GVV:126),
//                drop components of at most `min_size` (150) voxels
//   GVV:216      the result is saved as a uint8 mask
// and labelVolume (GVV:108-136): labelled volume + (label, size) list, components numbered in raster order.
//
// Connected components: lock-free union-find over voxel indices with randomised linking (the root with the smaller hashed
// index wins: expected logarithmic depth whatever the order of the concurrent unions); the smallest voxel index of every
// component is recorded beside its size, and the raster-order numbering of skimage / SciPy is a prefix sum over those first
// voxels.  Voxels of one x-run start out pointing at the run's first voxel (a warp scan per row), so only links between
// rows and planes need unions.
#include "../../include/vrg_b200.h"

#include "vrg_scratch.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

int vrg_edt_squared_device_internal(const uint8_t *d_mask, const int64_t *shape, int *sq_out, cudaStream_t stream);
int vrg_edt_rows_device_internal(const uint8_t *d_mask, const int64_t *shape, uint16_t *d1, uint32_t *bits, cudaStream_t stream);
void vrg_set_error_internal(const char *msg);  // vrg_b200.cu: text behind vrg_last_error()

namespace {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;
constexpr int GRID = 148 * 8;

// ---- range of the vesselness volume (GVV:187): per-block partial minima / maxima, folded on the host ---------------
__global__ void __launch_bounds__(256) k_minmax(const double *__restrict__ v, long long n, double *partial) {
    double lo = INFINITY, hi = -INFINITY;
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n; p += (long long)gridDim.x * 256) {
        const double x = v[p];
        lo = fmin(lo, x); hi = fmax(hi, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(FULLMASK, lo, o));
        hi = fmax(hi, __shfl_xor_sync(FULLMASK, hi, o));
    }
    __shared__ double s_lo[8], s_hi[8];
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { lo = fmin(lo, s_lo[i]); hi = fmax(hi, s_hi[i]); }
        partial[2 * blockIdx.x] = lo; partial[2 * blockIdx.x + 1] = hi;
    }
}

// ---- GVV:189-195: the two cut-offs and the binarisation, one pass ----------------------------------------------------
// A voxel survives iff it is not (near the brain boundary and <= t_edge), is > t_all, and is non-zero.
__global__ void __launch_bounds__(256) k_rule(const double *__restrict__ v, const int *__restrict__ edt_sq, long long n,
                                              double edge_distance, double t_edge, double t_all, uint8_t *__restrict__ out) {
    for (long long p = (long long)blockIdx.x * 256 + threadIdx.x; p < n; p += (long long)gridDim.x * 256) {
        const double x = v[p];
        const bool near_edge = sqrt((double)edt_sq[p]) <= edge_distance;
        const bool zeroed = (near_edge && x <= t_edge) || x <= t_all;
        out[p] = (!zeroed && x != 0.0) ? 1 : 0;
    }
}

// The same rule without a distance transform of the whole brain.  The distance to the brain boundary enters the rule only at
// voxels whose vesselness lies in (t_all, t_edge] (below, the voxel is zeroed anyway; above, it is kept anyway): bright vessel
// voxels, a fraction of a per cent of the volume.  k_rule_classify settles every other voxel and lists those; k_rule_near
// answers "is a non-brain voxel within edge_distance?" for each listed voxel exactly, one warp per voxel, from the packed
// bits of the brain mask: min over the rows (dz, dy) within reach of dz^2 + dy^2 + (distance to the nearest zero of that row
// within reach)^2 -- the separable transform restricted to the window that can matter.
__global__ void __launch_bounds__(256) k_rule_classify(const double *__restrict__ v, long long n, double t_edge, double t_all,
                                                       uint8_t *__restrict__ out, int *__restrict__ cand, unsigned int *__restrict__ cand_n) {
    const int lane = threadIdx.x & 31;
    const long long ngroups = (n + 3) / 4, nround = (ngroups + 31) / 32 * 32;
    const bool aligned = (((uintptr_t)v) & 15) == 0 && (((uintptr_t)out) & 3) == 0;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < nround; g += (long long)gridDim.x * 256) {
        const long long p0 = g * 4;
        double x[4] = {0.0, 0.0, 0.0, 0.0};
        if (g < ngroups) {
            if (aligned && p0 + 3 < n) {
                const double2 a = *(const double2 *)(v + p0), b = *(const double2 *)(v + p0 + 2);
                x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
            } else {
                for (int k = 0; k < 4; ++k)
                    if (p0 + k < n) x[k] = v[p0 + k];
            }
        }
        uint32_t o = 0u;
        unsigned c = 0u;  // bit k: voxel p0 + k is a candidate
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool settled = x[k] <= t_all || x[k] > t_edge || x[k] == 0.0;  // padding (0.0) is settled
            o |= (uint32_t)(x[k] > t_edge && x[k] != 0.0) << (8 * k);            // candidates start at 0
            c |= (unsigned)(!settled) << k;
        }
        if (g < ngroups) {
            if (aligned && p0 + 3 < n) *(uint32_t *)(out + p0) = o;
            else
                for (int k = 0; k < 4; ++k)
                    if (p0 + k < n) out[p0 + k] = (uint8_t)((o >> (8 * k)) & 1u);
        }
        if (__any_sync(FULLMASK, c != 0u)) {  // rare: append the candidates of this warp
            const int mine = __popc(c);
            int incl = mine;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(FULLMASK, incl, s);
                if (lane >= s) incl += t;
            }
            unsigned int slot = 0;
            if (lane == 31) slot = atomicAdd(cand_n, (unsigned int)incl);
            slot = __shfl_sync(FULLMASK, slot, 31) + (unsigned int)(incl - mine);
            for (int k = 0; k < 4; ++k)
                if (c & (1u << k)) cand[slot++] = (int)(p0 + k);
        }
    }
}
// packed foreground bits of a uint8 volume, [rows][XW] words (bits beyond the row end are 0)
__global__ void __launch_bounds__(256) k_pack_bits(const uint8_t *__restrict__ mask, uint32_t *__restrict__ bits, long long nrows, int X, int XW) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const uint8_t *row = mask + r * X;
        if ((X & 3) == 0 && (((uintptr_t)row) & 3) == 0) {  // 128 voxels per step: four bytes per lane, eight lanes per word
            for (int w0 = 0; w0 < XW; w0 += 4) {
                const int x = 32 * w0 + 4 * lane;
                uint32_t nib = 0u;
                if (x < X) {
                    const uint32_t v = *(const uint32_t *)(row + x);
                    const uint32_t nz = (((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
                    nib = ((nz >> 7) & 1u) | ((nz >> 14) & 2u) | ((nz >> 21) & 4u) | ((nz >> 28) & 8u);
                }
                uint32_t word = nib << (4 * (lane & 7));
                word |= __shfl_xor_sync(FULLMASK, word, 1);
                word |= __shfl_xor_sync(FULLMASK, word, 2);
                word |= __shfl_xor_sync(FULLMASK, word, 4);
                const int w = w0 + (lane >> 3);
                if ((lane & 7) == 0 && w < XW) bits[r * XW + w] = word;
            }
        } else {
            for (int w = 0; w < XW; ++w) {
                const int x = 32 * w + lane;
                const uint32_t word = __ballot_sync(FULLMASK, x < X && row[x] != 0);
                if (lane == 0) bits[r * XW + w] = word;
            }
        }
    }
}
// "a non-brain voxel lies in this block" for blocks of 8 planes x 8 rows x 32 voxels, one bit per x: a listed voxel whose
// surrounding blocks hold none is settled after one word per lane (the summary of a whole volume is a megabyte, L2-resident)
__global__ void __launch_bounds__(256) k_zero_summary(const uint32_t *__restrict__ bits, uint32_t *__restrict__ sum, int Z, int Y, int X, int XW) {
    const int nyb = (Y + 7) / 8, nzb = (Z + 7) / 8;
    const long long total = (long long)nzb * nyb * XW;
    const uint32_t tail = (X & 31) ? ((1u << (X & 31)) - 1u) : 0xFFFFFFFFu;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int w = (int)(i % XW), yb = (int)((i / XW) % nyb), zb = (int)(i / ((long long)XW * nyb));
        uint32_t acc = 0u;
        for (int z = 8 * zb; z < min(Z, 8 * zb + 8); ++z)
            for (int y = 8 * yb; y < min(Y, 8 * yb + 8); ++y) acc |= ~bits[((long long)z * Y + y) * XW + w];
        sum[i] = w == XW - 1 ? acc & tail : acc;
    }
}
// one warp per listed voxel, one lane per row (dz, dy) of the (2W+1)^2 rows within reach: the nearest non-brain voxel of the row
// inside [x - W, x + W] comes from one to three words of packed bits
__global__ void __launch_bounds__(256) k_rule_near(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ sum,
                                                   const int *__restrict__ cand, const unsigned int *__restrict__ cand_n, int Z, int Y, int X,
                                                   int XW, double edge_distance, int W, uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long m = *cand_n, nwarps = (long long)gridDim.x * 8;
    const int side = 2 * W + 1, nrows = side * side;
    for (long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); i < m; i += nwarps) {
        const int p = cand[i];
        const int x = p % X, y = (p / X) % Y, z = p / (X * Y);
        const int xlo = max(0, x - W), xhi = min(X - 1, x + W);
        {   // coarse: any non-brain voxel in the blocks that cover the box?  (deep inside the brain: none, and that is the answer)
            const int nyb = (Y + 7) / 8;
            const int zb0 = max(0, z - W) >> 3, zb1 = min(Z - 1, z + W) >> 3, yb0 = max(0, y - W) >> 3, yb1 = min(Y - 1, y + W) >> 3;
            const int w0 = xlo >> 5, nw = (xhi >> 5) - w0 + 1, ny = yb1 - yb0 + 1, combos = (zb1 - zb0 + 1) * ny * nw;
            bool some = false;
            for (int c0 = 0; c0 < combos && !some; c0 += 32) {
                const int c = c0 + lane;
                bool hit = false;
                if (c < combos) {
                    const int w = w0 + c % nw, yb = yb0 + (c / nw) % ny, zb = zb0 + c / (nw * ny);
                    const int b0 = max(xlo - 32 * w, 0), b1 = min(xhi - 32 * w, 31);
                    const uint32_t win = (b1 == 31 ? 0xFFFFFFFFu : ((2u << b1) - 1u)) & ~((1u << b0) - 1u);
                    hit = (sum[((long long)zb * nyb + yb) * XW + w] & win) != 0u;
                }
                some = __any_sync(FULLMASK, hit);
            }
            if (!some) {
                if (lane == 0) out[p] = 1;
                continue;
            }
        }
        bool near = false;
        for (int r0 = 0; r0 < nrows && !near; r0 += 32) {
            const int r = r0 + lane;
            bool hit = false;
            if (r < nrows) {
                const int dz = r / side - W, dy = r % side - W;
                const int zz = z + dz, yy = y + dy;
                if (zz >= 0 && zz < Z && yy >= 0 && yy < Y) {
                    const uint32_t *rw = bits + ((long long)zz * Y + yy) * XW;
                    int best = 1 << 20;  // |dx| to the nearest zero voxel of this row within the window
                    for (int w = xlo >> 5; w <= (xhi >> 5); ++w) {
                        const int b0 = max(xlo - 32 * w, 0), b1 = min(xhi - 32 * w, 31);  // window bits of this word
                        const uint32_t win = (b1 == 31 ? 0xFFFFFFFFu : ((2u << b1) - 1u)) & ~((1u << b0) - 1u);
                        const uint32_t zw = ~rw[w] & win;
                        if (!zw) continue;
                        const int xb = x - 32 * w;  // position of x relative to this word (may lie outside 0..31)
                        const uint32_t below = xb >= 31 ? zw : (xb < 0 ? 0u : zw & ((2u << xb) - 1u));
                        const uint32_t above = xb <= 0 ? zw : (xb > 31 ? 0u : zw & ~((1u << xb) - 1u));
                        if (below) best = min(best, xb - (31 - __clz(below)));
                        if (above) best = min(best, (__ffs(above) - 1) - xb);
                    }
                    if (best < (1 << 20)) {
                        const long long d2 = (long long)dz * dz + (long long)dy * dy + (long long)best * best;
                        hit = sqrt((double)d2) <= edge_distance;  // the reference compares the float64 distance
                    }
                }
            }
            near = __any_sync(FULLMASK, hit);
        }
        if (lane == 0 && !near) out[p] = 1;
    }
}

// ---- connected components ------------------------------------------------------------------------------------------------
// parent[p] = first voxel of p's x-run, foreground voxels only (nobody reads the parent of a background voxel: every consumer
// looks at the one-byte mask first, so a thin mask costs each pass 1 byte per voxel instead of 4): one warp per row, running
// maximum of the positions of background voxels
// The foreground voxels are also appended to `list` (any order): the union, flatten and count steps then run one thread per
// FOREGROUND voxel with full warps, instead of one thread per voxel with a lane or two of a warp at work.
__global__ void __launch_bounds__(256) k_cc_init(const uint8_t *__restrict__ fg, int *__restrict__ parent, int *__restrict__ size,
                                                 int *__restrict__ first, int *__restrict__ list, unsigned int *__restrict__ list_n,
                                                 long long nrows, int X) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const uint8_t *row = fg + r * X;
        int carry = -1;
        for (int x0 = 0; x0 < X; x0 += 32) {
            const int x = x0 + lane;
            const bool f = x < X && row[x] != 0;
            const unsigned fmask = __ballot_sync(FULLMASK, f);
            if (fmask == 0u) { carry = min(x0 + 31, X - 1); continue; }  // 32 voxels of background: the last one is the nearest so far
            unsigned int slot = 0;
            if (lane == 0) slot = atomicAdd(list_n, (unsigned int)__popc(fmask));
            slot = __shfl_sync(FULLMASK, slot, 0);
            if (f) list[slot + __popc(fmask & ((1u << lane) - 1u))] = (int)(r * X + x);
            int v = (x < X && !f) ? x : -1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULLMASK, v, o);
                if (lane >= o) v = max(v, t);
            }
            v = max(v, carry);
            carry = __shfl_sync(FULLMASK, v, 31);
            if (f) {  // size / first are read at roots only, and roots are foreground voxels
                parent[r * X + x] = (int)(r * X + v + 1);
                size[r * X + x] = 0;
                if (first != nullptr) first[r * X + x] = 0x7FFFFFFF;
            }
        }
    }
}

__device__ __forceinline__ int cc_find(int *parent, int a) {
    while (true) {
        const int pa = parent[a];
        if (pa == a) return a;
        const int ga = parent[pa];
        if (ga != pa) parent[a] = ga;  // path halving (a benign race: any ancestor is a valid parent)
        a = ga;
    }
}
// Which of two roots wins a union is decided by a hash of their indices, not by the indices: all links of a thin structure
// are made at the same time between singletons, and "the smaller index wins" then builds one path as long as the vessel
// (every later find walked it: 11 GB of pointer chasing at C3), while random priorities cut it into pieces of expected
// length two or three and keep the forest logarithmically shallow.  Links go from a root to a node of strictly smaller
// (priority, index), so there are no cycles; a root is hooked by compare-and-swap, so a node is hooked once.
__device__ __forceinline__ uint32_t cc_prio(int a) {
    uint32_t x = (uint32_t)a * 0x9E3779B1u;
    x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
    return x;
}
__device__ __forceinline__ void cc_union(int *parent, int a, int b) {
    while (true) {
        a = cc_find(parent, a); b = cc_find(parent, b);
        if (a == b) return;
        const uint32_t pa = cc_prio(a), pb = cc_prio(b);
        if (pa < pb || (pa == pb && a < b)) { const int t = a; a = b; b = t; }  // a loses: it is hung under b
        const int old = atomicCAS(&parent[a], a, b);
        if (old == a) return;
        a = old;  // somebody hooked a first: go on from its new parent
    }
}

// links to the previous row / plane: for each of the 4 earlier neighbour rows (dz,dy) in {(0,-1),(-1,-1),(-1,0),(-1,+1)},
// x-1, x, x+1; when the voxel straight across is foreground it already joins its own neighbours, so the diagonals are
// skipped.  Two x-runs that touch need one union, not one per touching voxel pair: a pair is linked only if p is the
// first voxel of its run or the neighbour is the first voxel of its run (every touching pair of runs has such a pair).
// Four voxels per thread and load: background words (almost all of a vessel mask) cost one 32-bit load.
__device__ __forceinline__ void cc_merge_voxel(const uint8_t *__restrict__ fg, int *parent, long long p, int Y, int X) {
    const int x = (int)(p % X), y = (int)((p / X) % Y), z = (int)(p / ((long long)X * Y));
    // the 17 mask bytes the rule below can look at are requested together (one round trip instead of a chain of them)
    const bool pl = x > 0 && fg[p - 1];
    bool ok[4], c0[4], cm1[4], cm2[4], cp1[4];
    long long qs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int dz = k == 0 ? 0 : -1, dy = k == 0 ? -1 : k - 2;
        const int zz = z + dz, yy = y + dy;
        ok[k] = zz >= 0 && yy >= 0 && yy < Y;
        const long long q = ok[k] ? ((long long)zz * Y + yy) * X + x : p;
        qs[k] = q;
        c0[k] = ok[k] && fg[q];
        cm1[k] = ok[k] && x > 0 && fg[q - 1];
        cm2[k] = ok[k] && x > 1 && fg[q - 2];
        cp1[k] = ok[k] && x + 1 < X && fg[q + 1];
    }
    const bool pstart = !pl;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const long long q = qs[k];
        if (c0[k]) {
            if (pstart || !cm1[k]) cc_union(parent, (int)p, (int)q);
        } else {
            if (cm1[k] && (pstart || x < 2 || !cm2[k])) cc_union(parent, (int)p, (int)(q - 1));
            if (cp1[k]) cc_union(parent, (int)p, (int)(q + 1));  // q + 1 starts its run (q is background)
        }
    }
}
__global__ void __launch_bounds__(256) k_cc_merge(const uint8_t *__restrict__ fg, int *parent, const int *__restrict__ list,
                                                  const unsigned int *__restrict__ list_n, int Y, int X) {
    // Every lane of a warp makes the same number of trips and the warp is brought back together at the end of each one: the
    // unions are compare-and-swap loops, and without the barrier the compiler lets the lanes of a warp drift apart for the rest
    // of the kernel (seen with ncu on the first version: one active thread per instruction, 20 x the instructions).
    const long long m = *list_n, nround = (m + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nround; i += (long long)gridDim.x * 256) {
        if (i < m) cc_merge_voxel(fg, parent, list[i], Y, X);
        __syncwarp();
    }
}

// parent[p] = root; component sizes at the roots (one atomic per group of equal roots in a warp).
// The walk to the root is read-only here: with path halving, another thread's late write of a stale grandparent
// could land on parent[p] after p's own thread stored the root, and the consumers below index by parent[p].
// `first` (optional): first[root] = smallest voxel index of the component (the numbering is in raster order of those).
__device__ __forceinline__ uint32_t cc_load4(const uint8_t *__restrict__ fg, long long p0, long long n, bool aligned) {
    if (aligned && p0 + 3 < n) return *(const uint32_t *)(fg + p0);
    uint32_t w = 0u;
    for (int b = 0; b < 4; ++b)
        if (p0 + b < n) w |= (uint32_t)(fg[p0 + b] != 0) << (8 * b);
    return w;
}
__global__ void __launch_bounds__(256) k_cc_flatten_count(int *parent, int *__restrict__ size, int *__restrict__ first,
                                                          const int *__restrict__ list, const unsigned int *__restrict__ list_n) {
    const int lane = threadIdx.x & 31;
    const long long m = *list_n, nround = (m + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nround; i += (long long)gridDim.x * 256) {
        int root = -1, p = 0x7FFFFFFF;
        if (i < m) {
            p = list[i];
            root = p;
            for (int up = parent[root]; up != root; up = parent[root]) root = up;
            parent[p] = root;
        }
        const unsigned peers = __match_any_sync(FULLMASK, root);
        const int pmin = __reduce_min_sync(peers, p);  // smallest voxel index among the lanes of the same root (the list has no order)
        if (root >= 0 && lane == __ffs(peers) - 1) {
            atomicAdd(&size[root], __popc(peers));
            if (first != nullptr) atomicMin(&first[root], pmin);
        }
    }
}
// chunk_count[c] = first voxels of components in chunk c of the volume (chunk_count zeroed by the caller)
__global__ void __launch_bounds__(256) k_cc_chunk_roots(const int *__restrict__ parent, const int *__restrict__ first,
                                                        const int *__restrict__ list, const unsigned int *__restrict__ list_n,
                                                        int *__restrict__ chunk_count, int chunk) {
    const long long m = *list_n;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < m; i += (long long)gridDim.x * 256) {
        const int p = list[i];
        if (first[parent[p]] == p) atomicAdd(&chunk_count[p / chunk], 1);
    }
}

// keep[p] = foreground and component size > min_size (GVV:198-200); counts the kept voxels and components
__global__ void __launch_bounds__(256) k_cc_filter(const uint8_t *__restrict__ fg, const int *__restrict__ parent,
                                                   const int *__restrict__ size, long long n, long long min_size,
                                                   uint8_t *__restrict__ out, unsigned long long *counts) {
    unsigned long long kept = 0, comps = 0;
    const bool aligned = (((uintptr_t)fg) & 3) == 0 && (((uintptr_t)out) & 3) == 0;
    const long long ngroups = (n + 3) / 4;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (long long)gridDim.x * 256) {
        const long long p0 = g * 4;
        const uint32_t w = cc_load4(fg, p0, n, aligned);
        uint32_t o = 0u;
        if (w != 0u) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (!((w >> (8 * b)) & 0xFFu)) continue;
                const long long p = p0 + b;
                const int root = parent[p];
                const bool keep = size[root] > min_size;
                o |= (uint32_t)keep << (8 * b);
                kept += keep;
                comps += keep && root == p;
            }
        }
        if (aligned && p0 + 3 < n) *(uint32_t *)(out + p0) = o;
        else
            for (int b = 0; b < 4; ++b)
                if (p0 + b < n) out[p0 + b] = (uint8_t)((o >> (8 * b)) & 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kept += __shfl_xor_sync(FULLMASK, kept, o);
        comps += __shfl_xor_sync(FULLMASK, comps, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(&counts[0], kept);
        if (comps) atomicAdd(&counts[1], comps);
    }
}

// raster-order numbering: roots per chunk of 1024 voxels -> exclusive scan over chunks -> labels
constexpr int CHUNK = 4096;
__device__ __forceinline__ bool cc_is_first(const uint8_t *__restrict__ fg, const int *__restrict__ parent, const int *__restrict__ first,
                                            long long p) {
    return fg[p] && first[parent[p]] == (int)p;
}
// one block: chunk_count -> exclusive prefix sums in place; total -> *total
__global__ void __launch_bounds__(1024) k_cc_scan_chunks(int *chunk_count, long long nchunks, int *total) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < nchunks; base += 1024) {
        const long long i = base + threadIdx.x;
        const int v = i < nchunks ? chunk_count[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULLMASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULLMASK, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int before = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
        if (i < nchunks) chunk_count[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
// number[root] = 1 + rank of the root in raster order (written into `size`, which is no longer needed at the roots'
// slots once `sizes_by_label` has been filled), then labels[p] = number[root(p)]
__global__ void __launch_bounds__(256) k_cc_number_roots(const uint8_t *__restrict__ fg, const int *__restrict__ parent,
                                                         const int *__restrict__ first, long long n, const int *__restrict__ chunk_base,
                                                         const int *__restrict__ total, int *__restrict__ size,
                                                         int *__restrict__ sizes_by_label) {
    const long long c0 = (long long)blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_w[8];
    for (long long c = c0; c * CHUNK < n; c += gridDim.x) {
        int running = chunk_base[c];
        if (((c + 1) * CHUNK < n ? chunk_base[c + 1] : *total) == running) continue;  // no first voxel in this chunk (block-uniform)
        for (int i0 = 0; i0 < CHUNK; i0 += 256) {  // 256 voxels at a time, in order
            const long long p = c * CHUNK + i0 + threadIdx.x;
            const bool is_root = p < n && cc_is_first(fg, parent, first, p);  // the component's first voxel stands for its root
            const unsigned b = __ballot_sync(FULLMASK, is_root);
            if (lane == 0) s_w[warp] = __popc(b);
            __syncthreads();
            int before = running;
            for (int w = 0; w < warp; ++w) before += s_w[w];
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += s_w[w];
            if (is_root) {
                const int number = before + __popc(b & ((1u << lane) - 1u)) + 1;
                const int root = parent[p];
                sizes_by_label[number - 1] = size[root];
                size[root] = number;
            }
            running += tot;
            __syncthreads();
        }
    }
}
__global__ void __launch_bounds__(256) k_cc_labels(const uint8_t *__restrict__ fg, const int *__restrict__ parent,
                                                   const int *__restrict__ number, long long n, int *__restrict__ labels) {
    const bool aligned = (((uintptr_t)fg) & 3) == 0 && (((uintptr_t)labels) & 15) == 0;
    const long long ngroups = (n + 3) / 4;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < ngroups; g += (long long)gridDim.x * 256) {
        const long long p0 = g * 4;
        const uint32_t w = cc_load4(fg, p0, n, aligned);
        int4 o = make_int4(0, 0, 0, 0);
        if (w != 0u) {
            if (w & 0xFFu) o.x = number[parent[p0]];
            if (w & 0xFF00u) o.y = number[parent[p0 + 1]];
            if (w & 0xFF0000u) o.z = number[parent[p0 + 2]];
            if (w & 0xFF000000u) o.w = number[parent[p0 + 3]];
        }
        if (aligned && p0 + 3 < n) *(int4 *)(labels + p0) = o;
        else {
            const int v[4] = {o.x, o.y, o.z, o.w};
            for (int b = 0; b < 4; ++b)
                if (p0 + b < n) labels[p0 + b] = v[b];
        }
    }
}

using DevBuf = vrg_scratch::Buf;

int status_of(cudaError_t e) {
    if (e == cudaSuccess) return VRG_OK;
    vrg_set_error_internal(cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
}
int bad_args(const char *what) {
    vrg_set_error_internal(what);
    return VRG_ERR_ARG;
}
const char *const SHAPE_MSG = "null buffer or bad shape (3 axes of 1..16384, fewer than 2^31 voxels)";

bool shape_ok(const int64_t *shape) {
    if (!shape || shape[0] <= 0 || shape[1] <= 0 || shape[2] <= 0) return false;
    if (shape[0] > 16384 || shape[1] > 16384 || shape[2] > 16384) return false;
    return (double)shape[0] * (double)shape[1] * (double)shape[2] < 2147483647.0;  // voxel indices are int32
}

// parent / size of the 26-connected components of a device uint8 volume (non-zero = foreground)
// d_list: n ints; d_list_n: one counter (the number of foreground voxels ends up there)
int components_device(const uint8_t *d_fg, const int64_t *shape, int *d_parent, int *d_size, int *d_first, int *d_list,
                      unsigned int *d_list_n, cudaStream_t st) {
    const long long Z = shape[0], Y = shape[1], X = shape[2];
    cudaMemsetAsync(d_list_n, 0, sizeof(unsigned int), st);
    k_cc_init<<<GRID, 256, 0, st>>>(d_fg, d_parent, d_size, d_first, d_list, d_list_n, Z * Y, (int)X);
    k_cc_merge<<<GRID, 256, 0, st>>>(d_fg, d_parent, d_list, d_list_n, (int)Y, (int)X);
    k_cc_flatten_count<<<GRID, 256, 0, st>>>(d_parent, d_size, d_first, d_list, d_list_n);
    return status_of(cudaGetLastError());
}

}  // namespace

// labelVolume, GVV:108-136.  binary: device uint8 (non-zero = foreground); labels_out: device int32 (0 = background,
// components 1..K in raster order of their first voxel); sizes_out (host, optional): voxel count of component k at [k-1].
extern "C" int vrg_label_components_device(int device, const uint8_t *binary_dev, const int64_t *shape, int32_t *labels_dev,
                                           int64_t *n_components, int64_t *sizes_out, int64_t sizes_cap, void *cuda_stream) {
    if (!binary_dev || !labels_dev || !shape_ok(shape)) return bad_args(SHAPE_MSG);
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n = (long long)shape[0] * shape[1] * shape[2], nchunks = (n + CHUNK - 1) / CHUNK;
    DevBuf parent, size, first, list, list_n, chunks, total, bylabel;
    cudaError_t e = parent.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = size.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = first.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = list.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = list_n.alloc(sizeof(unsigned int), st);
    if (e == cudaSuccess) e = chunks.alloc(nchunks * sizeof(int), st);
    if (e == cudaSuccess) e = total.alloc(sizeof(int), st);
    if (e != cudaSuccess) return status_of(e);
    int rc = components_device(binary_dev, shape, parent.as<int>(), size.as<int>(), first.as<int>(), list.as<int>(), list_n.as<unsigned int>(), st);
    if (rc != VRG_OK) return rc;
    cudaMemsetAsync(chunks.p, 0, nchunks * sizeof(int), st);
    k_cc_chunk_roots<<<GRID, 256, 0, st>>>(parent.as<int>(), first.as<int>(), list.as<int>(), list_n.as<unsigned int>(), chunks.as<int>(), CHUNK);
    k_cc_scan_chunks<<<1, 1024, 0, st>>>(chunks.as<int>(), nchunks, total.as<int>());
    int K = 0;
    e = cudaMemcpyAsync(&K, total.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = bylabel.alloc((size_t)K * sizeof(int), st);
    if (e != cudaSuccess) return status_of(e);
    k_cc_number_roots<<<GRID, 256, 0, st>>>(binary_dev, parent.as<int>(), first.as<int>(), n, chunks.as<int>(), total.as<int>(), size.as<int>(),
                                            bylabel.as<int>());
    k_cc_labels<<<GRID, 256, 0, st>>>(binary_dev, parent.as<int>(), size.as<int>(), n, labels_dev);
    e = cudaGetLastError();
    if (e == cudaSuccess && sizes_out && sizes_cap > 0 && K > 0) {
        std::vector<int> tmp((size_t)std::min<int64_t>(K, sizes_cap));
        e = cudaMemcpyAsync(tmp.data(), bylabel.p, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        for (size_t i = 0; i < tmp.size(); ++i) sizes_out[i] = tmp[i];
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (n_components) *n_components = K;
    return status_of(e);
}

extern "C" int vrg_label_components(int device, const uint8_t *binary_host, const int64_t *shape, int32_t *labels_host,
                                    int64_t *n_components, int64_t *sizes_out, int64_t sizes_cap) {
    if (!binary_host || !labels_host || !shape_ok(shape)) return bad_args(SHAPE_MSG);
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    const size_t n = (size_t)shape[0] * shape[1] * shape[2];
    DevBuf b, l;
    cudaError_t e = b.alloc(n, nullptr);
    if (e == cudaSuccess) e = l.alloc(n * sizeof(int32_t), nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(b.p, binary_host, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return status_of(e);
    const int rc = vrg_label_components_device(device, b.as<uint8_t>(), shape, l.as<int32_t>(), n_components, sizes_out, sizes_cap, nullptr);
    if (rc != VRG_OK) return rc;
    return status_of(cudaMemcpy(labels_host, l.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
}

// GVV:187-200 on device buffers.  vesselness: float64; brain_mask: uint8 (non-zero = brain); mask_out: uint8 0/1.
// info_out (optional, host): [0] kept voxels, [1] kept components, then the two cut-offs as doubles via thresholds_out.
extern "C" int vrg_vessel_mask_device(int device, const double *vesselness_dev, const uint8_t *brain_mask_dev, const int64_t *shape,
                                      double edge_distance, double edge_fraction, double fraction, int64_t min_size,
                                      uint8_t *mask_out_dev, int64_t *info_out, double *thresholds_out, void *cuda_stream) {
    if (!vesselness_dev || !brain_mask_dev || !mask_out_dev || !shape_ok(shape)) return bad_args(SHAPE_MSG);
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n = (long long)shape[0] * shape[1] * shape[2];
    DevBuf sq, partial, parent, size, counts, binary, list, list_n;
    cudaError_t e = partial.alloc(2 * GRID * sizeof(double), st);
    if (e == cudaSuccess) e = counts.alloc(2 * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = binary.alloc(n, st);
    if (e != cudaSuccess) return status_of(e);
    // range of the vesselness volume, GVV:187
    k_minmax<<<GRID, 256, 0, st>>>(vesselness_dev, n, partial.as<double>());
    std::vector<double> hp(2 * GRID);
    e = cudaMemcpyAsync(hp.data(), partial.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return status_of(e);
    double lo = hp[0], hi = hp[1];
    for (int i = 1; i < GRID; ++i) { lo = std::fmin(lo, hp[2 * i]); hi = std::fmax(hi, hp[2 * i + 1]); }
    if (!std::isfinite(lo) || !std::isfinite(hi)) { vrg_set_error_internal("vesselness volume holds NaN or Inf"); return VRG_ERR_NONFINITE; }
    const double t_edge = lo + edge_fraction * (hi - lo), t_all = lo + fraction * (hi - lo);  // GVV:189,191, same expression
    if (thresholds_out) { thresholds_out[0] = t_edge; thresholds_out[1] = t_all; }
    int rc = VRG_OK;
    if (edge_distance >= 0.0 && edge_distance <= 40.0) {
        // the distance to the brain-mask boundary (GVV:183) only where the rule can depend on it: see k_rule_near
        const long long Z = shape[0], Y = shape[1], X = shape[2];
        const int XW = (int)((X + 31) / 32);
        DevBuf bits, zsum, cand, cand_n;
        e = bits.alloc((size_t)Z * Y * XW * sizeof(uint32_t), st);
        if (e == cudaSuccess) e = zsum.alloc((size_t)((Z + 7) / 8) * ((Y + 7) / 8) * XW * sizeof(uint32_t), st);
        if (e == cudaSuccess) e = cand.alloc(n * sizeof(int), st);
        if (e == cudaSuccess) e = cand_n.alloc(sizeof(unsigned int), st);
        if (e != cudaSuccess) return status_of(e);
        k_pack_bits<<<GRID, 256, 0, st>>>(brain_mask_dev, bits.as<uint32_t>(), Z * Y, (int)X, XW);
        k_zero_summary<<<GRID, 256, 0, st>>>(bits.as<uint32_t>(), zsum.as<uint32_t>(), (int)Z, (int)Y, (int)X, XW);
        cudaMemsetAsync(cand_n.p, 0, sizeof(unsigned int), st);
        k_rule_classify<<<GRID, 256, 0, st>>>(vesselness_dev, n, t_edge, t_all, binary.as<uint8_t>(), cand.as<int>(), cand_n.as<unsigned int>());
        k_rule_near<<<GRID, 256, 0, st>>>(bits.as<uint32_t>(), zsum.as<uint32_t>(), cand.as<int>(), cand_n.as<unsigned int>(), (int)Z, (int)Y, (int)X, XW,
                                          edge_distance, (int)edge_distance, binary.as<uint8_t>());
    } else {
        // distance to the brain-mask boundary, GVV:183 (squared, exact), over the whole volume
        e = sq.alloc(n * sizeof(int), st);
        if (e != cudaSuccess) return status_of(e);
        rc = vrg_edt_squared_device_internal(brain_mask_dev, shape, sq.as<int>(), st);
        if (rc != VRG_OK) return rc;
        k_rule<<<GRID, 256, 0, st>>>(vesselness_dev, sq.as<int>(), n, edge_distance, t_edge, t_all, binary.as<uint8_t>());
        sq.release();  // stream-ordered: the block is reusable once k_rule has run
    }
    e = parent.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = size.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = list.alloc(n * sizeof(int), st);
    if (e == cudaSuccess) e = list_n.alloc(sizeof(unsigned int), st);
    if (e != cudaSuccess) return status_of(e);
    rc = components_device(binary.as<uint8_t>(), shape, parent.as<int>(), size.as<int>(), nullptr, list.as<int>(), list_n.as<unsigned int>(), st);
    if (rc != VRG_OK) return rc;
    cudaMemsetAsync(counts.p, 0, 2 * sizeof(unsigned long long), st);
    k_cc_filter<<<GRID, 256, 0, st>>>(binary.as<uint8_t>(), parent.as<int>(), size.as<int>(), n, min_size, mask_out_dev,
                                      counts.as<unsigned long long>());
    unsigned long long hc[2] = {0, 0};
    e = cudaMemcpyAsync(hc, counts.p, sizeof hc, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (info_out) { info_out[0] = (int64_t)hc[0]; info_out[1] = (int64_t)hc[1]; }
    return status_of(e);
}

extern "C" int vrg_vessel_mask(int device, const double *vesselness_host, const uint8_t *brain_mask_host, const int64_t *shape,
                               double edge_distance, double edge_fraction, double fraction, int64_t min_size,
                               uint8_t *mask_out_host, int64_t *info_out, double *thresholds_out) {
    if (!vesselness_host || !brain_mask_host || !mask_out_host || !shape_ok(shape)) return bad_args(SHAPE_MSG);
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    const size_t n = (size_t)shape[0] * shape[1] * shape[2];
    DevBuf v, b, o;
    cudaError_t e = v.alloc(n * sizeof(double), nullptr);
    if (e == cudaSuccess) e = b.alloc(n, nullptr);
    if (e == cudaSuccess) e = o.alloc(n, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(v.p, vesselness_host, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b.p, brain_mask_host, n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return status_of(e);
    const int rc = vrg_vessel_mask_device(device, v.as<double>(), b.as<uint8_t>(), shape, edge_distance, edge_fraction, fraction,
                                          min_size, o.as<uint8_t>(), info_out, thresholds_out, nullptr);
    if (rc != VRG_OK) return rc;
    return status_of(cudaMemcpy(mask_out_host, o.p, n, cudaMemcpyDeviceToHost));
}
