// common.cuh -- shared device/host helpers for libradar_retrieval (sm_100a only).
//
// Ranking keys.  Internally every score is a float "key" for which LARGER IS BETTER:
//     DPR    key = ip                          API score =  key
//     KL     key = X - H  ( = -KL exactly )    API score =  0 - key   (so that KL == 0 prints +0)
//     hybrid key = fma(alpha, ip, -(oma*KL))   API score =  key
// A candidate is one 64-bit "composite":  (ord(key) << 32) | (0xFFFFFFFF - local_row)  so that a plain
// unsigned DESCENDING sort realises the oracle's order (better score first, then smaller id;
// oracle/radar_oracle.c: better()).  composite 0 is never produced by a finite key and means "empty".
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/radar_retrieval.h"

namespace radar {

constexpr int kNumObs = RADAR_NUM_OBS;
constexpr int kObsPad = RADAR_OBS_PAD;

// ---- error plumbing (host) ----------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define RADAR_CUDA_CHECK(expr)                                                                   \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::radar::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                        \
            return RADAR_E_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

#define RADAR_ARG_CHECK(cond, ...)          \
    do {                                    \
        if (!(cond)) {                      \
            ::radar::set_error(__VA_ARGS__); \
            return RADAR_E_ARG;             \
        }                                   \
    } while (0)

// ---- key <-> orderable bits ---------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f + 0.0f);  // -0 -> +0 so that -0 and +0 tie like the oracle's float compare
#else
    float g = f + 0.0f;
    uint32_t u;
    memcpy(&u, &g, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

__device__ __forceinline__ uint64_t make_composite(float key, uint32_t local_row) {
    return (static_cast<uint64_t>(f2ord(key)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - local_row);
}
__device__ __forceinline__ float composite_key(uint64_t c) { return ord2f(static_cast<uint32_t>(c >> 32)); }
__device__ __forceinline__ uint32_t composite_row(uint64_t c) {
    return 0xFFFFFFFFu - static_cast<uint32_t>(c & 0xFFFFFFFFull);
}

// composite with a shard-local row -> the exchange form with the GLOBAL id (radar_search's out_packed)
__device__ __forceinline__ uint64_t packed_global(uint64_t c, int64_t idx_offset) {
    const uint32_t gid = static_cast<uint32_t>(static_cast<int64_t>(composite_row(c)) + idx_offset);
    return (c & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(0xFFFFFFFFu - gid);
}

// ---- bitonic sort, DESCENDING, of a power-of-two array in shared memory --------------------------
// `nthreads` threads (ids tid in [0,nthreads)) cooperate; `sync()` must synchronise exactly them.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* s, int n, int tid, int nthreads, SyncFn sync) {
    // the two stage loops stay rolled: callers are rare-path code whose size must not evict hot loops from the
    // instruction cache (fully unrolled, one 256-element sort was 27 KB of SASS)
#pragma unroll 1
    for (int size = 2; size <= n; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            sync();
            for (int t = tid; t < (n >> 1); t += nthreads) {
                const int lo = 2 * t - (t & (stride - 1));  // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);  // direction of this sub-sequence
                const uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    sync();
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// canonical scoring pieces (see oracle/radar_oracle.c header for the definition they implement)
__device__ __forceinline__ float canonical_key(int mode, float ip, float x, float h, float alpha, float oma) {
    if (mode == RADAR_MODE_DPR) return ip;
    if (mode == RADAR_MODE_KL) return __fsub_rn(x, h);
    const float kl = __fsub_rn(h, x);
    const float t = __fmul_rn(oma, kl);
    return __fmaf_rn(alpha, ip, -t);
}

__device__ __forceinline__ float api_score_from_key(int mode, float key) {
    return mode == RADAR_MODE_KL ? __fsub_rn(0.0f, key) : key;
}

}  // namespace radar
