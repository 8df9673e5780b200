// Device generator of the synthetic vessel-tree phantom: the integer-exact twin of
// arterynetwork_b200/phantom.py, so a slab generated here equals the NumPy volume bit for bit.
// Bench/test input only -- nothing on the VRG path depends on it.
#include "../../include/vrg_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

constexpr int BR = 16;  // candidate-segment lists are kept per 16^3 brick

__device__ __forceinline__ unsigned long long splitmix(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void k_phantom(long long Y, long long X, long long z0, long long nz, const long long *__restrict__ seg,
                          const int *__restrict__ off, const int *__restrict__ items, int nby, int nbx, long long seed,
                          long long quantum, long long sigma_k, long long excl_k, int use_excl, double *__restrict__ data,
                          uint8_t *__restrict__ vm) {
    const long long n = nz * Y * X;
    const long long noise0 = (131070ll * sigma_k) / 37837ll;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long x = i % X, y = (i / X) % Y, zl = i / (X * Y), z = z0 + zl;
        const int b = ((int)(zl / BR) * nby + (int)(y / BR)) * nbx + (int)(x / BR);
        bool inside = false;
        for (int k = off[b]; k < off[b + 1] && !inside; ++k) {
            const long long *s = seg + 8ll * items[k];
            const long long dz = s[3] - s[0], dy = s[4] - s[1], dx = s[5] - s[2];
            const long long wz = z - s[0], wy = y - s[1], wx = x - s[2];
            const long long c1 = wz * dz + wy * dy + wx * dx, c2 = dz * dz + dy * dy + dx * dx;
            const long long w2 = wz * wz + wy * wy + wx * wx, r2 = s[6];
            if (c1 <= 0) inside = w2 <= r2;
            else if (c1 >= c2) {
                const long long ez = z - s[3], ey = y - s[4], ex = x - s[5];
                inside = ez * ez + ey * ey + ex * ex <= r2;
            } else inside = (w2 * c2 - c1 * c1) <= r2 * c2;
        }
        const unsigned long long lin = (unsigned long long)((z * Y + y) * X + x);
        const unsigned long long h = splitmix(lin + (unsigned long long)(seed + 1) * 0x9E3779B97F4A7C15ull);
        const long long s4 = (long long)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48));
        const long long kq = (inside ? quantum : 0) + (s4 * sigma_k) / 37837ll - noise0;
        data[i] = (double)kq / (double)quantum;
        vm[i] = (use_excl && kq <= excl_k) ? 4 : 3;
    }
}

__global__ void k_seeds(long long Y, long long X, long long z0, long long nz, const long long *__restrict__ roots,
                        long long n_roots, uint8_t *__restrict__ vm) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_roots * 8) return;
    const long long r = t / 8, c = t % 8;
    const long long z = roots[3 * r] + (c >> 2), y = roots[3 * r + 1] + ((c >> 1) & 1), x = roots[3 * r + 2] + (c & 1);
    if (z < z0 || z >= z0 + nz || y >= Y || x >= X) return;
    vm[((z - z0) * Y + y) * X + x] = 0;  // 2x2x2 seed cube, cf. VRG:288-289
}

}  // namespace

extern "C" int vrg_phantom_device(int device, const int64_t *shape, int64_t z0, int64_t nz, const int64_t *segments,
                                  int64_t n_segments, const int64_t *roots, int64_t n_roots, int64_t seed,
                                  int64_t quantum, int64_t sigma_k, int64_t exclude_below_k, int use_exclude,
                                  double *data_dev, uint8_t *value_map_dev) {
    if (!shape || !data_dev || !value_map_dev || nz <= 0) return VRG_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return VRG_ERR_CUDA;
    const int64_t Y = shape[1], X = shape[2];
    const int nbz = (int)((nz + BR - 1) / BR), nby = (int)((Y + BR - 1) / BR), nbx = (int)((X + BR - 1) / BR);
    std::vector<std::vector<int>> lists((size_t)nbz * nby * nbx);
    for (int64_t i = 0; i < n_segments; ++i) {
        const int64_t *s = segments + 8 * i;
        const int64_t r = (int64_t)std::ceil(std::sqrt((double)s[6])) + 1;
        const int64_t zlo = std::max(std::min(s[0], s[3]) - r, z0), zhi = std::min(std::max(s[0], s[3]) + r, z0 + nz - 1);
        const int64_t ylo = std::max<int64_t>(std::min(s[1], s[4]) - r, 0), yhi = std::min(std::max(s[1], s[4]) + r, Y - 1);
        const int64_t xlo = std::max<int64_t>(std::min(s[2], s[5]) - r, 0), xhi = std::min(std::max(s[2], s[5]) + r, X - 1);
        if (zlo > zhi || ylo > yhi || xlo > xhi) continue;
        for (int64_t bz = (zlo - z0) / BR; bz <= (zhi - z0) / BR; ++bz)
            for (int64_t by = ylo / BR; by <= yhi / BR; ++by)
                for (int64_t bx = xlo / BR; bx <= xhi / BR; ++bx) lists[((size_t)bz * nby + by) * nbx + bx].push_back((int)i);
    }
    std::vector<int> off(lists.size() + 1, 0), items;
    for (size_t b = 0; b < lists.size(); ++b) {
        off[b + 1] = off[b] + (int)lists[b].size();
        items.insert(items.end(), lists[b].begin(), lists[b].end());
    }
    long long *d_seg = nullptr, *d_roots = nullptr;
    int *d_off = nullptr, *d_items = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_seg, std::max<size_t>(1, (size_t)n_segments * 8) * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_roots, std::max<size_t>(1, (size_t)n_roots * 3) * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_off, off.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_items, std::max<size_t>(1, items.size()) * sizeof(int));
    if (e == cudaSuccess && n_segments) e = cudaMemcpy(d_seg, segments, (size_t)n_segments * 8 * sizeof(long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_roots) e = cudaMemcpy(d_roots, roots, (size_t)n_roots * 3 * sizeof(long long), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !items.empty()) e = cudaMemcpy(d_items, items.data(), items.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_phantom<<<148 * 16, 256>>>(Y, X, z0, nz, d_seg, d_off, d_items, nby, nbx, seed, quantum, sigma_k, exclude_below_k,
                                     use_exclude, data_dev, value_map_dev);
        if (n_roots) k_seeds<<<(unsigned)((n_roots * 8 + 127) / 128), 128>>>(Y, X, z0, nz, d_roots, n_roots, value_map_dev);
        e = cudaDeviceSynchronize();
    }
    cudaFree(d_seg); cudaFree(d_roots); cudaFree(d_off); cudaFree(d_items);
    return e == cudaSuccess ? VRG_OK : VRG_ERR_CUDA;
}
