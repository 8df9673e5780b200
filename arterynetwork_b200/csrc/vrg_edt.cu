// Exact Euclidean distance transform of a binary mask: the step after the VRG path.
// Reference call sites: Code/manualCorrectionGUI.py:248 (vessel radii from the VRG output mask) and
// Code/generateVesselVolume.py:183 (distance to the brain-mask boundary), both
// scipy.ndimage.distance_transform_edt(mask) with default arguments: every non-zero voxel gets its Euclidean distance
// to the nearest zero voxel, float64, unit spacing.
//
// Separable and exact in integers (Meijster, Roerdink, Hesselink 2000):
//   pass x   g(x)  = (distance to the nearest zero voxel of the same row)^2: a prefix-max / suffix-min scan, one warp per row
//   pass y,z out(u) = min_i (u - i)^2 + in(i) along the axis: lower envelope of parabolas, one thread per line, O(n) per
//            line whatever the distances are; neighbouring threads own neighbouring x, so every access of the scan
//            (input, envelope stacks, output) is a coalesced row of the volume
//   sqrt     fp64 square root of the integer squared distance (correctly rounded, as NumPy's)
// Squared distances are int32 (axes <= 32768 would overflow: axes are limited to 16384, 3 * 16384^2 < 2^30).
#include "../../include/vrg_b200.h"

#include "vrg_scratch.cuh"

#include <cuda_runtime.h>

#include <algorithm>

void vrg_set_error_internal(const char *msg);  // vrg_b200.cu: text behind vrg_last_error()

namespace {

using vrg_scratch::Buf;

constexpr int EDT_INF = 1 << 30;  // "no zero voxel seen yet"; sums with k^2 are formed in 64 bits
constexpr unsigned FULLMASK = 0xFFFFFFFFu;

// pass x: one warp per row, 32 voxels per step; the nearest zero to the left is a running maximum of positions,
// the nearest to the right a running minimum, each a 5-step warp scan plus a carry
__global__ void __launch_bounds__(256) k_edt_rows(const uint8_t *__restrict__ mask, int *__restrict__ out, long long nrows, int X) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const int NONE_L = -(1 << 20), NONE_R = 1 << 20;  // farther than any axis: the squared distance saturates to EDT_INF
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += nwarps) {
        const uint8_t *row = mask + r * X;
        int *orow = out + r * X;
        int carry = NONE_L;
        for (int x0 = 0; x0 < X; x0 += 32) {  // left to right: store the distance to the left zero
            const int x = x0 + lane;
            int v = (x < X && row[x] == 0) ? x : NONE_L;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULLMASK, v, o);
                if (lane >= o) v = max(v, t);
            }
            v = max(v, carry);
            carry = __shfl_sync(FULLMASK, v, 31);
            if (x < X) orow[x] = x - v;  // >= 2^20 - X when there is no zero to the left
        }
        carry = NONE_R;
        for (int x0 = ((X - 1) / 32) * 32; x0 >= 0; x0 -= 32) {  // right to left
            const int x = x0 + lane;
            int v = (x < X && row[x] == 0) ? x : NONE_R;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_down_sync(FULLMASK, v, o);
                if (lane + o < 32) v = min(v, t);
            }
            v = min(v, carry);
            carry = __shfl_sync(FULLMASK, v, 0);
            if (x < X) {
                const int d = min(orow[x], v - x);
                orow[x] = d >= 32768 ? EDT_INF : d * d;
            }
        }
    }
}

// pass y / z: out(u) = min_i (u - i)^2 + f(i) along one axis: lower envelope of parabolas (Meijster).  Thread = one
// line; `line` enumerates (outer, x) so that threads of a warp sit on neighbouring x, walk u in lockstep, and every
// access is a coalesced row of the volume.
// Where f(u) = 0 the answer is 0, and a zero shields everything behind it ((u - i)^2 + f(i) > (u - z)^2 for i beyond a
// zero at z): only the zeros next to a positive voxel can matter to the envelope, so the zeros inside a stretch of
// background are neither pushed nor popped.  For a vessel mask almost every voxel is background: the pass is two reads
// and one write per voxel, and the envelope stack (s / t / g: parabola apex, start of its reign, height) holds a few
// entries per vessel crossing, at the front of the line's slots, which stay in cache.  For a solid mask (the brain) the
// runs are long and the scan is still O(n) per line whatever the distances.  A line is one dependent chain: inputs are
// fetched eight steps ahead in both directions, and the pops of the backward scan (entries below the top no longer
// change) fetch four stack entries at once.
__global__ void __launch_bounds__(128) k_edt_lines(const int *__restrict__ in, int *__restrict__ out, short *__restrict__ s,
                                                   int *__restrict__ t, int *__restrict__ g, long long nlines, int len,
                                                   long long stride, int X, long long outer_stride) {
    for (long long line = (long long)blockIdx.x * blockDim.x + threadIdx.x; line < nlines; line += (long long)gridDim.x * blockDim.x) {
        const long long base = (line / X) * outer_stride + (line % X);
        const int *f = in + base;
        short *ss = s + base;
        int *tt = t + base, *gg = g + base;
        int *o = out + base;
        // forward scan.  Registers hold the top of the stack (sq, tq, fq = f(sq)); q = depth - 1, -1 = empty.
        int q = -1, sq = 0, tq = 0, prev = 0;
        long long fq = 0;
        constexpr int PF = 8;
        for (int u0 = 0; u0 < len; u0 += PF) {
            int fpre[PF + 1];
#pragma unroll
            for (int k = 0; k <= PF; ++k) fpre[k] = f[(long long)min(u0 + k, len - 1) * stride];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int u = u0 + k;
                if (u >= len) break;
                const long long fu = fpre[k];
                const bool wanted = fu == 0 ? (prev > 0 || (u + 1 < len && fpre[k + 1] > 0)) : fu < EDT_INF;
                prev = (int)fu;
                if (!wanted) continue;  // background away from any foreground, or an infinite parabola
                while (q >= 0) {
                    // F(tq, sq) > F(tq, u): the new parabola is already lower where the top one starts to reign: pop
                    const long long a = (long long)(tq - sq) * (tq - sq) + fq, b = (long long)(tq - u) * (tq - u) + fu;
                    if (a <= b) break;
                    --q;
                    if (q >= 0) { sq = ss[(long long)q * stride]; tq = tt[(long long)q * stride]; fq = gg[(long long)q * stride]; }
                }
                long long w = 0;  // first position where parabola u is the lowest
                if (q >= 0) w = 1 + ((long long)u * u - (long long)sq * sq + fu - fq) / (2LL * (u - sq));
                if (w < len) {
                    ++q; sq = u; tq = (int)w; fq = fu;
                    ss[(long long)q * stride] = (short)u; tt[(long long)q * stride] = (int)w; gg[(long long)q * stride] = (int)fu;
                }
            }
        }
        // backward scan
        constexpr int PB = 4;
        int ps[PB], pt[PB], pg[PB], have = 0;  // entries q-1 .. q-have, fetched ahead (ps[0] is the next one to pop)
        for (int u0 = len - 1; u0 >= 0; u0 -= PF) {
            int fpre[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) fpre[k] = f[(long long)max(u0 - k, 0) * stride];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int u = u0 - k;
                if (u < 0) break;
                while (q > 0 && u < tq) {  // the top parabola reigns from tq on: below it the previous one does
                    if (have == 0) {
#pragma unroll
                        for (int j = 0; j < PB; ++j) {
                            const long long e = (long long)max(q - 1 - j, 0) * stride;
                            ps[j] = ss[e]; pt[j] = tt[e]; pg[j] = gg[e];
                        }
                        have = min(PB, q);
                    }
                    --q; --have;
                    sq = ps[0]; tq = pt[0]; fq = pg[0];
#pragma unroll
                    for (int j = 0; j + 1 < PB; ++j) { ps[j] = ps[j + 1]; pt[j] = pt[j + 1]; pg[j] = pg[j + 1]; }
                }
                long long d = EDT_INF;
                if (fpre[k] == 0) d = 0;
                else if (q >= 0) d = (long long)(u - sq) * (u - sq) + fq;
                o[(long long)u * stride] = d >= EDT_INF ? EDT_INF : (int)d;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_edt_sqrt(const int *__restrict__ sq, double *__restrict__ out, long long n, int *no_background) {
    bool inf = false;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int v = sq[p];
        inf |= v >= EDT_INF;
        out[p] = sqrt((double)v);
    }
    if (inf) *no_background = 1;
}

int edt_check(const int64_t *shape) {
    if (!shape || shape[0] <= 0 || shape[1] <= 0 || shape[2] <= 0) return VRG_ERR_ARG;
    if (shape[0] > 16384 || shape[1] > 16384 || shape[2] > 16384) return VRG_ERR_ARG;
    return VRG_OK;
}

// squared distances (int32) of a device mask; scratch = 3 int32 + 1 int16 volume.  sq_out must hold n ints.
// Asynchronous on `stream`.
int edt_squared_device(const uint8_t *d_mask, const int64_t *shape, int *sq_out, cudaStream_t stream) {
    const long long Z = shape[0], Y = shape[1], X = shape[2], n = Z * Y * X;
    Buf a, t, s, g;
    cudaError_t e = a.alloc(n * sizeof(int), stream);
    if (e == cudaSuccess) e = t.alloc(n * sizeof(int), stream);
    if (e == cudaSuccess) e = g.alloc(n * sizeof(int), stream);
    if (e == cudaSuccess) e = s.alloc(n * sizeof(short), stream);
    if (e == cudaSuccess) {
        const int grid = 148 * 8;
        const int gy = (int)std::min<long long>((Z * X + 127) / 128, 148 * 16), gz = (int)std::min<long long>((Y * X + 127) / 128, 148 * 16);
        k_edt_rows<<<grid, 256, 0, stream>>>(d_mask, sq_out, Z * Y, (int)X);                                             // mask -> sq_out
        k_edt_lines<<<gy, 128, 0, stream>>>(sq_out, a.as<int>(), s.as<short>(), t.as<int>(), g.as<int>(), Z * X, (int)Y, X, (int)X, X * Y);   // y: sq_out -> a
        k_edt_lines<<<gz, 128, 0, stream>>>(a.as<int>(), sq_out, s.as<short>(), t.as<int>(), g.as<int>(), Y * X, (int)Z, X * Y, (int)(X * Y), 0);  // z: a -> sq_out
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
    return VRG_OK;
}

int edt_device(const uint8_t *d_mask, const int64_t *shape, double *d_out, cudaStream_t stream) {
    const long long n = (long long)shape[0] * shape[1] * shape[2];
    Buf sq, flag;
    cudaError_t e = sq.alloc(n * sizeof(int), stream);
    if (e == cudaSuccess) e = flag.alloc(sizeof(int), stream);
    int rc = VRG_OK;
    if (e == cudaSuccess) {
        rc = edt_squared_device(d_mask, shape, sq.as<int>(), stream);
        if (rc == VRG_OK) {
            cudaMemsetAsync(flag.p, 0, sizeof(int), stream);
            k_edt_sqrt<<<148 * 8, 256, 0, stream>>>(sq.as<int>(), d_out, n, flag.as<int>());
            int h_flag = 0;
            e = cudaMemcpyAsync(&h_flag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            if (e == cudaSuccess && h_flag) {  // no zero voxel anywhere: the transform is undefined
                vrg_set_error_internal("distance transform: the mask has no zero voxel");
                rc = VRG_ERR_ARG;
            }
        }
    }
    if (e != cudaSuccess) {
        vrg_set_error_internal(cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
    }
    return rc;
}

int bad_args(const char *what) {
    vrg_set_error_internal(what);
    return VRG_ERR_ARG;
}

}  // namespace

// shared with vrg_mask.cu: squared EDT of a device mask into a device int32 volume
int vrg_edt_squared_device_internal(const uint8_t *d_mask, const int64_t *shape, int *sq_out, cudaStream_t stream) {
    return edt_squared_device(d_mask, shape, sq_out, stream);
}

// mask / dist live on the device
extern "C" int vrg_edt_device(int device, const uint8_t *mask_dev, const int64_t *shape, double *dist_dev, void *cuda_stream) {
    if (!mask_dev || !dist_dev || edt_check(shape) != VRG_OK) return bad_args("distance transform: null buffer or bad shape (axes 1..16384)");
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    return edt_device(mask_dev, shape, dist_dev, (cudaStream_t)cuda_stream);
}

// host buffers: mask uint8 (non-zero = foreground), dist float64, both (Z, Y, X) C order
extern "C" int vrg_edt(int device, const uint8_t *mask_host, const int64_t *shape, double *dist_host) {
    if (!mask_host || !dist_host || edt_check(shape) != VRG_OK) return bad_args("distance transform: null buffer or bad shape (axes 1..16384)");
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    vrg_scratch::pool_setup(device);
    const size_t n = (size_t)shape[0] * shape[1] * shape[2];
    Buf d_mask, d_out;
    cudaError_t e = d_mask.alloc(n, nullptr);
    if (e == cudaSuccess) e = d_out.alloc(n * sizeof(double), nullptr);
    int rc = VRG_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_mask.p, mask_host, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = edt_device(d_mask.as<uint8_t>(), shape, d_out.as<double>(), nullptr);
        if (rc == VRG_OK) e = cudaMemcpy(dist_host, d_out.p, n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) {
        vrg_set_error_internal(cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
    }
    return rc;
}

// hands the cached scratch blocks of the handle-less entry points (EDT, labelling, vessel mask) back to the driver
extern "C" int vrg_release_scratch(int device) {
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    cudaDeviceSynchronize();
    vrg_scratch::pool_trim(device);
    return VRG_OK;
}
