// Exact Euclidean distance transform of a binary mask: the step after the VRG path.
// Reference call sites: Code/manualCorrectionGUI.py:248 (vessel radii from the VRG output mask) and
// Code/generateVesselVolume.py:183 (distance to the brain-mask boundary), both
// scipy.ndimage.distance_transform_edt(mask) with default arguments: every non-zero voxel gets its Euclidean distance
// to the nearest zero voxel, float64.
//
// Separable and exact in integers: three passes  out(p) = min_k in(p + k*stride) + k^2  along x, y, z over int32
// squared distances, then one sqrt.  Each pass is a windowed search: the candidate at offset k costs at least k^2,
// so a voxel stops as soon as k^2 >= its current best -- the work per voxel is its own distance, which is small for
// vessel masks.  Threads map to x in every pass, so all loads and stores are coalesced whatever the axis.
#include "../../include/vrg_b200.h"

#include <cuda_runtime.h>

namespace {

constexpr int EDT_INF = 1 << 29;  // + k^2 (k <= 2^15) stays below 2^31

template <bool FIRST>
__global__ void __launch_bounds__(256) k_edt_pass(const uint8_t *__restrict__ mask, const int *__restrict__ in, int *__restrict__ out,
                                                  long long n, int len, long long stride) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int pos = (int)((p / stride) % len);
        int best = FIRST ? (mask[p] ? EDT_INF : 0) : in[p];
        for (int k = 1; (long long)k * k < best; ++k) {
            const bool lo = pos - k >= 0, hi = pos + k < len;
            if (!lo && !hi) break;
            const int kk = k * k;
            if (lo) {
                const long long q = p - (long long)k * stride;
                const int v = FIRST ? (mask[q] ? EDT_INF : 0) : in[q];
                best = min(best, v + kk);
            }
            if (hi) {
                const long long q = p + (long long)k * stride;
                const int v = FIRST ? (mask[q] ? EDT_INF : 0) : in[q];
                best = min(best, v + kk);
            }
        }
        out[p] = best;
    }
}

__global__ void __launch_bounds__(256) k_edt_sqrt(const int *__restrict__ sq, double *__restrict__ out, long long n, int *no_background) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int v = sq[p];
        if (v >= EDT_INF) *no_background = 1;
        out[p] = sqrt((double)v);
    }
}

int edt_device(const uint8_t *d_mask, const int64_t *shape, double *d_out, cudaStream_t stream) {
    const long long Z = shape[0], Y = shape[1], X = shape[2], n = Z * Y * X;
    int *a = nullptr, *b = nullptr, *flag = nullptr;
    cudaError_t e = cudaMalloc((void **)&a, n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&b, n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&flag, sizeof(int));
    int rc = VRG_OK;
    if (e == cudaSuccess) {
        cudaMemsetAsync(flag, 0, sizeof(int), stream);
        const int grid = 148 * 16;
        k_edt_pass<true><<<grid, 256, 0, stream>>>(d_mask, nullptr, a, n, (int)X, 1);
        k_edt_pass<false><<<grid, 256, 0, stream>>>(nullptr, a, b, n, (int)Y, X);
        k_edt_pass<false><<<grid, 256, 0, stream>>>(nullptr, b, a, n, (int)Z, X * Y);
        k_edt_sqrt<<<grid, 256, 0, stream>>>(a, d_out, n, flag);
        int h_flag = 0;
        e = cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess && h_flag) rc = VRG_ERR_ARG;  // no zero voxel anywhere: the transform is undefined
    }
    if (e != cudaSuccess) rc = e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
    cudaFree(a); cudaFree(b); cudaFree(flag);
    return rc;
}

}  // namespace

// mask / dist live on the device
extern "C" int vrg_edt_device(int device, const uint8_t *mask_dev, const int64_t *shape, double *dist_dev, void *cuda_stream) {
    if (!mask_dev || !shape || !dist_dev || shape[0] <= 0 || shape[1] <= 0 || shape[2] <= 0) return VRG_ERR_ARG;
    if (shape[0] > 32768 || shape[1] > 32768 || shape[2] > 32768) return VRG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    return edt_device(mask_dev, shape, dist_dev, (cudaStream_t)cuda_stream);
}

// host buffers: mask uint8 (non-zero = foreground), dist float64, both (Z, Y, X) C order
extern "C" int vrg_edt(int device, const uint8_t *mask_host, const int64_t *shape, double *dist_host) {
    if (!mask_host || !shape || !dist_host || shape[0] <= 0 || shape[1] <= 0 || shape[2] <= 0) return VRG_ERR_ARG;
    if (shape[0] > 32768 || shape[1] > 32768 || shape[2] > 32768) return VRG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    const size_t n = (size_t)shape[0] * shape[1] * shape[2];
    uint8_t *d_mask = nullptr;
    double *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_mask, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, n * sizeof(double));
    int rc = VRG_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_mask, mask_host, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = edt_device(d_mask, shape, d_out, 0);
        if (rc == VRG_OK) e = cudaMemcpy(dist_host, d_out, n * sizeof(double), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) rc = e == cudaErrorMemoryAllocation ? VRG_ERR_NOMEM : VRG_ERR_CUDA;
    cudaFree(d_mask); cudaFree(d_out);
    return rc;
}
