// Scratch device memory of the handle-less entry points (EDT, labelling, vessel mask): stream-ordered allocations from
// the device's default memory pool, which is told to keep freed blocks (release threshold = max), so that after the
// first call an allocation costs microseconds instead of the milliseconds of cudaMalloc / cudaFree.
// vrg_release_scratch() hands the cached blocks back to the driver.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrg_scratch {

inline cudaError_t pool_setup(int device) {
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return cudaSuccess;
    cudaMemPool_t pool;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
    if (e != cudaSuccess) return e;
    uint64_t keep = UINT64_MAX;
    e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    done[device] = e == cudaSuccess;
    return e;
}

inline void pool_trim(int device) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
}

struct Buf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    Buf() = default;
    Buf(const Buf &) = delete;
    Buf &operator=(const Buf &) = delete;
    ~Buf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
    }
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        release();
        st = stream;
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, stream);
        if (e == cudaErrorMemoryAllocation) {  // cached blocks of other sizes may be in the way: give them back, try once more
            cudaGetLastError();
            cudaStreamSynchronize(stream);
            int dev = 0;
            cudaGetDevice(&dev);
            pool_trim(dev);
            e = cudaMallocAsync(&p, bytes ? bytes : 1, stream);
        }
        if (e != cudaSuccess) p = nullptr;
        return e;
    }
    template <typename T> T *as() const { return (T *)p; }
};

}  // namespace vrg_scratch
