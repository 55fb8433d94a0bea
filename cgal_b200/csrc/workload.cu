// workload.cu — device-side generators of the benchmark workloads + result checksums (include/b200work.h).
// Bit-for-bit the numpy generators of cgal_b200/workloads.py (counter-based SplitMix64, u = (x >> 11) * 2^-53).
// Build with -fmad=false like the rest of the library: the binary64 operation order is the contract.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200aabb.h"
#include "../../include/b200work.h"

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// workloads.splitmix64(seed, stream, index)
__host__ __device__ inline uint64_t splitmix64(uint32_t seed, uint32_t stream, uint64_t index) {
    uint64_t x = index + 0x9E3779B97F4A7C15ull * ((uint64_t)seed * 1315423911ull + (uint64_t)stream * 2654435761ull + 1ull);
    return mix64(x + 0x9E3779B97F4A7C15ull);
}
__host__ __device__ inline double uniform01(uint32_t seed, uint32_t stream, uint64_t index) {
    return (double)(splitmix64(seed, stream, index) >> 11) * (1.0 / 9007199254740992.0);
}

struct Box6 { double v[6]; };
struct Cam13 { double v[13]; };

__global__ void __launch_bounds__(256) bbox_points_kernel(Box6 b, uint32_t seed, uint64_t offset, uint64_t n, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < 3 * n; j += stride) {   // one coordinate per thread: coalesced 8-byte stores
        const int k = (int)(j % 3);
        const double lo = b.v[k], ext = b.v[3 + k] - b.v[k];
        out[j] = lo + uniform01(seed, 1u, 3 * offset + j) * ext;
    }
}

__global__ void __launch_bounds__(256) pinhole_rays_kernel(Cam13 c, uint32_t width, uint32_t height, uint32_t row0, uint64_t n, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double th = c.v[12];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t col = (uint32_t)(i % width), row = row0 + (uint32_t)(i / width);
        const double px = ((double)col + 0.5) / (double)width * 2.0 - 1.0;
        const double py = 1.0 - ((double)row + 0.5) / (double)height * 2.0;
        const double sx = px * th * (double)width / (double)height, sy = py * th;
        double* o = out + 6 * i;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double d = (c.v[3 + k] + sx * c.v[6 + k]) + sy * c.v[9 + k];
            o[k] = c.v[k];
            o[3 + k] = c.v[k] + d;
        }
    }
}

constexpr uint64_t kK1 = 0x9E3779B97F4A7C15ull, kK2 = 0xD1B54A32D192ED03ull;

__global__ void __launch_bounds__(256) checksum_kernel(const uint64_t* __restrict__ words, uint64_t nwords, uint64_t tail, int tail_bytes, uint64_t index0,
                                                       unsigned long long* __restrict__ sums) {
    unsigned long long s0 = 0, s1 = 0;
    const uint64_t total = nwords + (tail_bytes ? 1 : 0);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t w = i < nwords ? words[i] : tail, pos = index0 + i + 1;
        s0 += mix64(w ^ (kK1 * pos));
        s1 += mix64(w + kK2 * pos);
    }
    for (int off = 16; off > 0; off >>= 1) { s0 += __shfl_xor_sync(0xFFFFFFFFu, s0, off); s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, off); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], s0); atomicAdd(&sums[1], s1); }
}

// FP64 pipe microbenchmark: 8 independent DFMA chains per thread (SURVEY.md §8d asks for the measured FP64 peak next to the traversal's FP64 share)
__global__ void __launch_bounds__(256) dfma_kernel(double* __restrict__ out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

unsigned grid_for(uint64_t n) { const uint64_t b = (n + 255) / 256; return (unsigned)(b < 148ull * 16 ? (b ? b : 1) : 148ull * 16); }

}  // namespace

extern "C" {

int b200work_bbox_points(int device, void* cuda_stream, const double bbox[6], uint32_t seed, uint64_t offset, uint64_t n, double* out) {
    if (!bbox || (n && !out)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    if (n == 0) return B200AABB_OK;
    Box6 b;
    for (int k = 0; k < 6; ++k) b.v[k] = bbox[k];
    bbox_points_kernel<<<grid_for(3 * n), 256, 0, (cudaStream_t)cuda_stream>>>(b, seed, offset, n, out);
    return cudaGetLastError() == cudaSuccess ? B200AABB_OK : B200AABB_ERR_CUDA;
}

int b200work_pinhole_rays(int device, void* cuda_stream, const double camera[13], uint32_t width, uint32_t height, uint32_t row0, uint32_t nrows, double* out) {
    if (!camera || !width || !height || (nrows && !out)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    if (nrows == 0) return B200AABB_OK;
    Cam13 c;
    for (int k = 0; k < 13; ++k) c.v[k] = camera[k];
    const uint64_t n = (uint64_t)nrows * width;
    pinhole_rays_kernel<<<grid_for(n), 256, 0, (cudaStream_t)cuda_stream>>>(c, width, height, row0, n, out);
    return cudaGetLastError() == cudaSuccess ? B200AABB_OK : B200AABB_ERR_CUDA;
}

int b200work_dfma_tflops(int device, double* tflops) {
    if (!tflops) return B200AABB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    const int blocks = 148 * 8, iters = 1 << 14;
    double* d = nullptr;
    if (cudaMalloc(&d, (size_t)blocks * 256 * 8) != cudaSuccess) return B200AABB_ERR_OUT_OF_MEMORY;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return B200AABB_ERR_CUDA; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && (best == 0 || ms < best)) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * 256 / (best * 1e-3) / 1e12;
    return B200AABB_OK;
}

int b200work_checksum128(int device, void* cuda_stream, const void* data, uint64_t nbytes, uint64_t index0, uint64_t sum[2]) {
    if (!sum || (nbytes && !data)) return B200AABB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return B200AABB_ERR_NO_DEVICE;
    sum[0] = sum[1] = 0;
    if (nbytes == 0) return B200AABB_OK;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    unsigned long long* d_sums = nullptr;
    if (cudaMalloc(&d_sums, 16) != cudaSuccess) return B200AABB_ERR_OUT_OF_MEMORY;
    const uint64_t nwords = nbytes / 8;
    const int tail_bytes = (int)(nbytes % 8);
    uint64_t tail = 0;
    bool ok = cudaMemsetAsync(d_sums, 0, 16, st) == cudaSuccess;
    if (ok && tail_bytes) ok = cudaMemcpyAsync(&tail, (const char*)data + 8 * nwords, tail_bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    if (ok) {
        checksum_kernel<<<grid_for(nwords + 1), 256, 0, st>>>((const uint64_t*)data, nwords, tail, tail_bytes, index0, d_sums);
        ok = cudaMemcpyAsync(sum, d_sums, 16, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    }
    cudaFree(d_sums);
    return ok ? B200AABB_OK : B200AABB_ERR_CUDA;
}

}  // extern "C"
