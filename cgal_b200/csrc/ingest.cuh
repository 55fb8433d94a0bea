// ingest.cuh — OFF text -> vertex / face arrays on the device ("mesh ingest on device").
//
// Replaces, for plain triangle OFF files, the reference's reader Surface_mesh/include/CGAL/Surface_mesh/IO/OFF.h:323-345 ->
// Stream_support/include/CGAL/IO/OFF/File_scanner_OFF.h (vertex coordinates read with `std::istream >> double`, i.e. correctly rounded; faces as vertex counts
// + indices): the face order of the file is the face order of the Surface_mesh and of the tree's primitives (tests/golden/make_fixtures.py relies on the same).
//   1. token_flags / token_offsets kernels: a byte starts a token when it is not white space and the byte before it is; block counts -> scan -> offsets
//   2. parse kernels: one thread per token; decimal literal -> binary64 with the Eisel-Lemire algorithm (D. Lemire, "Number parsing at a gigabyte per second",
//      Software: Practice and Experience 51(8), 2021; 128-bit truncated powers of five from scripts/make_pow5_table.py), which rounds correctly for every
//      literal of at most 19 significant digits; longer literals are decided when truncation cannot change the result and are otherwise FLAGGED and parsed by
//      the caller with strtod (text parsing is not a geometric predicate: this is not the "CPU fallback" the predicates must not have)
// The literal parser is host-callable: tests/hostcheck runs it against Python's float() on adversarial strings.
#pragma once
#include <stdint.h>
#include "exact.cuh"      // B2_HD
#include "pow5_table.h"

namespace b2 {

#if defined(__CUDACC__)
__device__ static const uint64_t kPow5[2 * (B2_POW5_MAX - B2_POW5_MIN + 1)] = {B2_POW5_TABLE};
#endif
#if !defined(__CUDA_ARCH__)
namespace host_tables { static const uint64_t Pow5[2 * (B2_POW5_MAX - B2_POW5_MIN + 1)] = {B2_POW5_TABLE}; }
#endif
B2_HD uint64_t pow5_word(int i) {
#if defined(__CUDA_ARCH__)
    return kPow5[i];
#else
    return host_tables::Pow5[i];
#endif
}
B2_HD void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umul64hi(a, b);
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p; hi = (uint64_t)(p >> 64);
#endif
}
B2_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// w * 10^q (w != 0, any q) rounded to the nearest binary64, ties to even: bits of the result.  Exact for every (w, q): the 128-bit product decides except in
// cases that cannot occur for a 64-bit w (Lemire 2021, §9; Mushtak & Lemire 2023 remove the last fallback).
B2_HD uint64_t decimal_to_binary64(uint64_t w, int64_t q) {
    if (w == 0 || q < B2_POW5_MIN) return 0;
    if (q > B2_POW5_MAX) return 0x7FF0000000000000ull;
    int lz = clz64(w);
    w <<= lz;
    const int idx = 2 * (int)(q - B2_POW5_MIN);
    uint64_t hi, lo;
    mul64(w, pow5_word(idx), hi, lo);
    if ((hi & 0x1FFull) == 0x1FFull) {                       // the 55 bits kept could still change: take the low word of the power into account
        uint64_t hi2, lo2;
        mul64(w, pow5_word(idx + 1), hi2, lo2);
        lo += hi2;
        if (hi2 > lo) ++hi;
    }
    const int upperbit = (int)(hi >> 63);
    const int shift = upperbit + 64 - 52 - 3;
    uint64_t mant = hi >> shift;
    int power2 = (int)((((int64_t)(152170 + 65536) * q) >> 16) + 63) + upperbit - lz + 1023;
    if (power2 <= 0) {                                       // subnormal result
        if (-power2 + 1 >= 64) return 0;
        mant >>= -power2 + 1;
        mant += (mant & 1);
        mant >>= 1;
        return mant;                                         // (mant == 2^52 is the smallest normal number: its bit pattern is right as it stands)
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1 && (mant << shift) == hi) mant &= ~1ull;   // exactly half way: to even
    mant += (mant & 1);
    mant >>= 1;
    if (mant >= (2ull << 52)) { mant = 1ull << 52; ++power2; }
    mant &= ~(1ull << 52);
    if (power2 >= 0x7FF) return 0x7FF0000000000000ull;
    return mant | ((uint64_t)power2 << 52);
}

// One decimal floating-point literal [s, e): [+-] digits [. digits] [(e|E) [+-] digits] (at least one digit in the mantissa), as `istream >> double` accepts.
// status: 0 = ok (bits = the correctly rounded value), 1 = not a number, 2 = more than 19 significant digits and the dropped ones matter: the caller decides.
B2_HD uint64_t parse_decimal(const char* s, const char* e, int& status) {
    status = 1;
    bool neg = false;
    if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
    uint64_t w = 0;
    int ndig = 0, nsig = 0;          // digits seen, significant digits kept (<= 19)
    int64_t q = 0;
    bool dropped_nonzero = false, seen_point = false;
    for (; s < e; ++s) {
        const char c = *s;
        if (c == '.') { if (seen_point) return 0; seen_point = true; continue; }
        if (c < '0' || c > '9') break;
        ++ndig;
        if (nsig < 19) {
            if (w != 0 || c != '0') { w = w * 10 + (uint64_t)(c - '0'); ++nsig; }
            if (seen_point) --q;
        } else {
            if (c != '0') dropped_nonzero = true;
            if (!seen_point) ++q;
        }
    }
    if (ndig == 0) return 0;
    if (s < e) {
        if (*s != 'e' && *s != 'E') return 0;
        ++s;
        bool eneg = false;
        if (s < e && (*s == '-' || *s == '+')) { eneg = *s == '-'; ++s; }
        if (s == e) return 0;
        int64_t ex = 0;
        for (; s < e; ++s) {
            if (*s < '0' || *s > '9') return 0;
            if (ex < 100000) ex = ex * 10 + (*s - '0');
        }
        q += eneg ? -ex : ex;
    }
    const uint64_t sign = neg ? 0x8000000000000000ull : 0ull;
    uint64_t bits = decimal_to_binary64(w, q);
    if (dropped_nonzero) {                                   // the true value lies strictly between w and w + 1 (times 10^q)
        if (w == 0xFFFFFFFFFFFFFFFFull || decimal_to_binary64(w + 1, q) != bits) { status = 2; return 0; }
    }
    status = 0;
    return sign | bits;
}

// a non-negative decimal integer token; status as above (1 = malformed or > 2^32 - 1)
B2_HD uint32_t parse_index(const char* s, const char* e, int& status) {
    status = 1;
    if (s == e) return 0;
    uint64_t v = 0;
    for (; s < e; ++s) {
        if (*s < '0' || *s > '9') return 0;
        v = v * 10 + (uint64_t)(*s - '0');
        if (v > 0xFFFFFFFFull) return 0;
    }
    status = 0;
    return (uint32_t)v;
}

B2_HD bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

#if defined(__CUDACC__)
constexpr int kTokBlock = 256, kTokPerThread = 16, kTokTile = kTokBlock * kTokPerThread;   // bytes per block

// pass 1: tokens starting in each tile of the body [begin, len)
__global__ void __launch_bounds__(kTokBlock) token_count_kernel(const char* __restrict__ text, uint64_t begin, uint64_t len, uint32_t* __restrict__ tile_counts) {
    const uint64_t base = begin + (uint64_t)blockIdx.x * kTokTile;
    unsigned cnt = 0;
    for (int k = 0; k < kTokPerThread; ++k) {
        const uint64_t i = base + (uint64_t)k * kTokBlock + threadIdx.x;
        if (i < len) cnt += (!is_space(text[i]) && (i == begin || is_space(text[i - 1]))) ? 1u : 0u;
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    __shared__ unsigned s[kTokBlock / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = 0; for (int j = 0; j < kTokBlock / 32; ++j) t += s[j]; tile_counts[blockIdx.x] = t; }
}
// pass 2: exclusive scan of the tile counts (one block; a 1 GB file has 2.6e5 tiles)
__global__ void __launch_bounds__(1024) token_scan_kernel(uint32_t* __restrict__ tile_counts, uint32_t ntiles, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s[1024];
    unsigned long long carry = 0;
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < ntiles ? tile_counts[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const unsigned long long a = threadIdx.x >= (unsigned)d ? s[threadIdx.x - d] : 0;
            __syncthreads();
            s[threadIdx.x] += a;
            __syncthreads();
        }
        const unsigned long long incl = s[threadIdx.x];
        if (i < ntiles) tile_counts[i] = (uint32_t)(carry + incl - v);      // (a body of < 2^32 bytes has < 2^31 tokens)
        carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// pass 3: offsets of the tokens, in file order (thread t of a tile owns kTokPerThread CONSECUTIVE bytes here, so its tokens are consecutive)
__global__ void __launch_bounds__(kTokBlock) token_offsets_kernel(const char* __restrict__ text, uint64_t begin, uint64_t len, const uint32_t* __restrict__ tile_offsets,
                                                                  uint32_t* __restrict__ tok, uint64_t tok_cap) {
    const uint64_t base = begin + (uint64_t)blockIdx.x * kTokTile + (uint64_t)threadIdx.x * kTokPerThread;
    unsigned mask = 0;
    for (int k = 0; k < kTokPerThread; ++k) {
        const uint64_t i = base + k;
        if (i < len && !is_space(text[i]) && (i == begin || is_space(text[i - 1]))) mask |= 1u << k;
    }
    const unsigned mine = __popc(mask);
    // exclusive prefix over the block
    unsigned incl = mine;
    for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((threadIdx.x & 31) >= (unsigned)d) incl += v; }
    __shared__ unsigned wsum[kTokBlock / 32];
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned before = 0;
    for (unsigned j = 0; j < (threadIdx.x >> 5); ++j) before += wsum[j];
    uint64_t at = (uint64_t)tile_offsets[blockIdx.x] + before + incl - mine;
    for (int k = 0; k < kTokPerThread; ++k)
        if (mask & (1u << k)) { if (at < tok_cap) tok[at] = (uint32_t)(base + k); ++at; }
}

__device__ __forceinline__ const char* token_end(const char* text, uint64_t start, uint64_t len) {
    uint64_t i = start;
    while (i < len && !is_space(text[i])) ++i;
    return text + i;
}
// tokens [0, 3 nv): coordinates.  flags[0] |= 1 malformed token, flags[1] counts literals left to the caller (their token numbers go to `undecided`)
__global__ void __launch_bounds__(256) parse_vertices_kernel(const char* __restrict__ text, uint64_t len, const uint32_t* __restrict__ tok, uint64_t ncoord, double* __restrict__ verts,
                                                             unsigned* __restrict__ flags, uint32_t* __restrict__ undecided, uint32_t undecided_cap) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncoord) return;
    int status;
    const uint64_t bits = parse_decimal(text + tok[k], token_end(text, tok[k], len), status);
    verts[k] = __longlong_as_double((long long)bits);
    if (status == 1) atomicOr(&flags[0], 1u);
    else if (status == 2) { const unsigned at = atomicAdd(&flags[1], 1u); if (at < undecided_cap) undecided[at] = (uint32_t)k; }
}
// tokens [3 nv, 3 nv + 4 nf): "3 i j k" per face.  flags[0] |= 2 a face is not a triangle, |= 4 malformed or out-of-range index
__global__ void __launch_bounds__(256) parse_faces_kernel(const char* __restrict__ text, uint64_t len, const uint32_t* __restrict__ tok, uint64_t nf, uint64_t nv,
                                                          uint32_t* __restrict__ faces, unsigned* __restrict__ flags) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nf) return;
    const uint32_t* t = tok + 4 * j;
    int st;
    const uint32_t cnt = parse_index(text + t[0], token_end(text, t[0], len), st);
    if (st || cnt != 3) { atomicOr(&flags[0], st ? 4u : 2u); return; }
    for (int c = 0; c < 3; ++c) {
        const uint32_t v = parse_index(text + t[1 + c], token_end(text, t[1 + c], len), st);
        if (st || v >= nv) { atomicOr(&flags[0], 4u); return; }
        faces[3 * j + c] = v;
    }
}
#endif  // __CUDACC__

}  // namespace b2
