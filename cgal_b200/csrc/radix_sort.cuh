// radix_sort.cuh — in-house onesweep LSD radix sort of (uint64 key, uint32 value) pairs for sm_100a.
//
// One pass over the data per 8-bit digit: every tile of 4096 keys ranks its keys in shared memory, publishes its
// per-digit counts, resolves its global offsets by decoupled look-back over the previous tiles' published
// counts/prefixes (Merrill & Garland 2016; Adinets & Merrill "Onesweep" 2022 — published algorithms, own code), and
// scatters keys and values straight to their final position of the pass.  Per pass and element: 12 B read + 12 B
// written, coalesced (warp-striped 8-byte key loads, run-contiguous stores).  The digit histograms of all passes
// come from one upfront kernel (8 B/key read).
//
// Used for the 63-bit Morton keys of the LBVH build (replaces the median split by std::nth_element,
// AABB_tree/include/CGAL/AABB_traits_3.h:596-626) and for query coherence sorting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

constexpr int kRsBits = 8;
constexpr int kRsBins = 1 << kRsBits;
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;           // 4096 keys per tile
constexpr uint32_t kRsFlagAgg = 1u << 30, kRsFlagIncl = 2u << 30, kRsFlagMask = 3u << 30, kRsValMask = ~kRsFlagMask;
constexpr size_t kRsSmemBytes = kRsTile * 8 + kRsWarps * kRsBins * 4 + kRsBins * 4 * 2 + 64;   // 42 KB: four blocks per SM (values reuse the key area)

// Digit histograms of all passes, accumulated by a block in shared memory (sh[8 * kRsBins], zeroed by rs_hist_begin) and flushed once.
// Kernels that PRODUCE the keys (the Morton kernel of the build) call these themselves, which saves the separate 8 B/key read.
__device__ __forceinline__ void rs_hist_begin(uint32_t* sh) {
    for (int i = threadIdx.x; i < 8 * kRsBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
}
__device__ __forceinline__ void rs_hist_add(uint32_t* sh, uint64_t k, int first_bit, int npasses) {
#pragma unroll
    for (int p = 0; p < 8; ++p)
        if (p < npasses) atomicAdd(&sh[p * kRsBins + ((k >> (first_bit + p * kRsBits)) & (kRsBins - 1))], 1u);
}
__device__ __forceinline__ void rs_hist_end(const uint32_t* sh, uint32_t* __restrict__ hist, int npasses) {
    __syncthreads();
    for (int i = threadIdx.x; i < npasses * kRsBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// hist[pass][bin] += number of keys whose digit `pass` (bits first_bit + 8 pass ..) is `bin`
__global__ void __launch_bounds__(256) rs_histogram_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ hist, int first_bit, int npasses) {
    __shared__ uint32_t sh[8 * kRsBins];
    rs_hist_begin(sh);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) rs_hist_add(sh, keys[i], first_bit, npasses);
    rs_hist_end(sh, hist, npasses);
}

// in place: hist[pass][*] -> exclusive prefix sums
__global__ void __launch_bounds__(kRsBins) rs_scan_kernel(uint32_t* __restrict__ hist) {
    __shared__ uint32_t s[kRsBins];
    uint32_t* h = hist + blockIdx.x * kRsBins;
    const uint32_t v = h[threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < kRsBins; off <<= 1) {
        const uint32_t add = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
        __syncthreads();
        s[threadIdx.x] += add;
        __syncthreads();
    }
    h[threadIdx.x] = s[threadIdx.x] - v;
}

template <bool HAS_VALUES>
__global__ void __launch_bounds__(kRsThreads, 4) rs_onesweep_kernel(const uint64_t* __restrict__ kin, uint64_t* __restrict__ kout,
                                                                 const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                                                 uint64_t n, int shift, const uint32_t* __restrict__ global_base,
                                                                 volatile uint32_t* lookback, uint32_t* tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(smem_raw);   // the values pass through the same area once the keys have left it
    uint32_t* s_whist = reinterpret_cast<uint32_t*>(s_keys + kRsTile);   // [warps][bins]: per-warp digit counts, then per-warp exclusive offsets
    uint32_t* s_dstart = s_whist + kRsWarps * kRsBins;    // exclusive prefix of the tile's digit counts
    uint32_t* s_goff = s_dstart + kRsBins;                // global index of tile-sorted position 0 of each digit run, minus s_dstart
    uint32_t* s_misc = s_goff + kRsBins;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_misc[0] = atomicAdd(tile_counter, 1u);          // tiles are numbered in start order: look-back cannot deadlock
    for (int i = tid; i < kRsWarps * kRsBins; i += kRsThreads) s_whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_misc[0];
    const uint64_t tile_base = (uint64_t)tile * kRsTile;
    const uint32_t valid = (uint32_t)min((uint64_t)kRsTile, n - tile_base);

    // warp-striped load: item j of lane l of warp w is element w*512 + j*32 + l of the tile
    uint64_t key[kRsItems];
    const uint64_t wbase = tile_base + (uint64_t)warp * (32 * kRsItems) + lane;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint64_t idx = wbase + (uint64_t)j * 32;
        key[j] = idx < n ? kin[idx] : ~0ull;                         // sentinels sort to the end of the tile
    }
    // stable rank of every key among the keys of its warp with the same digit
    uint32_t rank[kRsItems];
    uint32_t* wh = s_whist + warp * kRsBins;
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & (kRsBins - 1);
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) { base = wh[d]; wh[d] = base + __popc(peers); }
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        rank[j] = base + __popc(peers & ((1u << lane) - 1));
        __syncwarp();
    }
    __syncthreads();
    // digit `tid`: exclusive scan over warps, tile count
    uint32_t count = 0;
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) { const uint32_t c = s_whist[w * kRsBins + tid]; s_whist[w * kRsBins + tid] = count; count += c; }
    if (tile == 0) lookback[tid] = count | kRsFlagIncl;
    else lookback[(uint64_t)tile * kRsBins + tid] = count | kRsFlagAgg;
    // exclusive scan of the counts over digits
    {
        uint32_t incl = count;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) s_misc[1 + warp] = incl;
        __syncthreads();
        uint32_t wsum = 0;
        for (int w = 0; w < warp; ++w) wsum += s_misc[1 + w];
        s_dstart[tid] = wsum + incl - count;
    }
    // decoupled look-back: exclusive prefix of digit `tid` over all previous tiles
    uint32_t excl = 0;
    if (tile > 0) {
        int64_t p = (int64_t)tile - 1;
        while (true) {
            uint32_t v = lookback[(uint64_t)p * kRsBins + tid];
            while ((v & kRsFlagMask) == 0) v = lookback[(uint64_t)p * kRsBins + tid];
            excl += v & kRsValMask;
            if (v & kRsFlagIncl) break;
            --p;
        }
        lookback[(uint64_t)tile * kRsBins + tid] = ((excl + count) & kRsValMask) | kRsFlagIncl;
    }
    s_goff[tid] = global_base[tid] + excl - s_dstart[tid];
    __syncthreads();
    // scatter into tile-sorted order in shared memory
    uint32_t pos[kRsItems];
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t d = (uint32_t)(key[j] >> shift) & (kRsBins - 1);
        pos[j] = s_dstart[d] + s_whist[warp * kRsBins + d] + rank[j];
        s_keys[pos[j]] = key[j];
    }
    __syncthreads();
    // run-contiguous stores to the final position of this pass
    uint32_t g[kRsItems];
#pragma unroll
    for (int j = 0; j < kRsItems; ++j) {
        const uint32_t i = tid + j * kRsThreads;
        g[j] = 0;
        if (i < valid) {
            const uint64_t k = s_keys[i];
            const uint32_t d = (uint32_t)(k >> shift) & (kRsBins - 1);
            g[j] = s_goff[d] + i;
            kout[g[j]] = k;
        }
    }
    if (HAS_VALUES) {
        __syncthreads();                                              // every key has left shared memory
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            const uint64_t idx = wbase + (uint64_t)j * 32;
            s_vals[pos[j]] = idx < n ? vin[idx] : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            const uint32_t i = tid + j * kRsThreads;
            if (i < valid) vout[g[j]] = s_vals[i];
        }
    }
}

struct RadixSortWorkspace {
    uint32_t* hist = nullptr;        // [8][256]
    uint32_t* lookback = nullptr;    // [8][ntiles][256]
    uint32_t* counters = nullptr;    // [8]
    uint64_t ntiles = 0;
    size_t bytes = 0;
};

inline uint64_t rs_num_tiles(uint64_t n) { return (n + kRsTile - 1) / kRsTile; }
inline size_t rs_workspace_bytes(uint64_t n) { return 8 * kRsBins * 4 + 8 * rs_num_tiles(n) * kRsBins * 4 + 8 * 4; }

inline void rs_set_attributes() {
    // (per device: a process may drive several GPUs)
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(rs_onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRsSmemBytes);
        cudaFuncSetAttribute(rs_onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRsSmemBytes);
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
}

// Zeroes the workspace (histograms, look-back words, tile counters).  Call before the digit histograms are accumulated.
inline void rs_begin(void* ws, uint64_t n, cudaStream_t stream) { cudaMemsetAsync(ws, 0, rs_workspace_bytes(n), stream); }
inline uint32_t* rs_hist_of(void* ws) { return reinterpret_cast<uint32_t*>(ws); }

// The sorting passes over key bits [first_bit, first_bit + 8 npasses), given the digit histograms in the workspace (rs_histogram_kernel, or
// the kernel that produced the keys).  vals[0] == nullptr: keys only.  Returns the index (0/1) of the buffers holding the result.
inline int rs_sort_passes(uint64_t* keys[2], uint32_t* vals[2], uint64_t n, int first_bit, int npasses, void* ws, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return 0;
    const uint64_t ntiles = rs_num_tiles(n);
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws);
    uint32_t* lookback = hist + 8 * kRsBins;
    uint32_t* counters = lookback + 8 * ntiles * kRsBins;
    rs_scan_kernel<<<npasses, kRsBins, 0, stream>>>(hist);
    rs_set_attributes();
    int cur = 0;
    for (int p = 0; p < npasses; ++p) {
        if (vals && vals[0])
            rs_onesweep_kernel<true><<<(unsigned)ntiles, kRsThreads, kRsSmemBytes, stream>>>(keys[cur], keys[cur ^ 1], vals[cur], vals[cur ^ 1], n, first_bit + p * kRsBits,
                                                                                         hist + p * kRsBins, lookback + (uint64_t)p * ntiles * kRsBins, counters + p);
        else
            rs_onesweep_kernel<false><<<(unsigned)ntiles, kRsThreads, kRsSmemBytes, stream>>>(keys[cur], keys[cur ^ 1], nullptr, nullptr, n, first_bit + p * kRsBits,
                                                                                          hist + p * kRsBins, lookback + (uint64_t)p * ntiles * kRsBins, counters + p);
        cur ^= 1;
    }
    if (launches) *launches += 1 + npasses;
    return cur;
}

// Sorts n (key,value) pairs by the low `npasses`*8 key bits.  keys/vals ping-pong between the two buffers; returns
// the index (0/1) of the buffer pair holding the sorted result.  ws = rs_workspace_bytes(n) bytes of device memory.
// launches: number of kernels launched is added to *launches.
inline int radix_sort_pairs(uint64_t* keys[2], uint32_t* vals[2], uint64_t n, int npasses, void* ws, cudaStream_t stream, uint64_t* launches) {
    if (n == 0) return 0;
    rs_begin(ws, n, stream);
    const uint64_t hblocks = (n + 255) / 256;
    const int hgrid = (int)(hblocks < 148 * 8 ? hblocks : 148 * 8);
    rs_histogram_kernel<<<hgrid, 256, 0, stream>>>(keys[0], n, rs_hist_of(ws), 0, npasses);
    if (launches) *launches += 1;
    return rs_sort_passes(keys, vals, n, 0, npasses, ws, stream, launches);
}

}  // namespace b2
