// proj_gemm.cuh -- the query-side prologue on the tensor pipe:
//     y = normalize(x W^T + b)      x [B, K] fp32 (BiomedCLIP features, K = 768), W [512, K] (nn.Linear layout)
// replaces  F.normalize(nn.Linear(768, 512)(features), dim=-1)  (dpr.py:202-203, :246, :263).
//
// fp32 accuracy from bf16 tensor cores: both operands are split into bf16 hi + lo (x = x_hi + x_lo + O(2^-17 |x|)), and
// three products are accumulated in fp32 -- hi.hi + hi.lo + lo.hi; the dropped lo.lo term and the split residuals are
// < 2^-16 relative per product, far inside the 1e-5 tolerance of a unit vector's components (tests/test_gpu_parity.py).
//
//   proj_split_kernel   fp32 [rows, K] -> bf16 [rows_pad, 2 K] = [hi | lo]   (x per call; W per call, 0.4 M elements)
//   proj_gemm_kernel    one CTA per 128 rows of x: TMA (64-byte swizzle) streams K in blocks of 32 through a 2-stage
//                       ring -- per stage the [128 x 32] hi / lo tiles of x and the [512 x 32] hi / lo tiles of W --,
//                       one thread issues tcgen05.mma (cta_group::1, M = 128, N = 256, two column halves) into the
//                       512 fp32 TMEM columns = the whole [128 x 512] output tile; four epilogue warps (thread = row) add
//                       the bias, reduce the row norm, and write the normalised row as fp32 and / or as bf16 -- the
//                       bf16 rows ARE the A operand of the DPR filter (tc_filter.cuh packs exactly bf16(e_q)).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "kl_filter.cuh"
#include "tc_filter.cuh"

namespace radar {
namespace proj {

using namespace tc;

constexpr int kOut = 512;            // output features == TMEM columns
constexpr int kRows = 128;           // rows of x per CTA == TMEM lanes
constexpr int kKBlock = 32;          // K elements per stage (64-byte rows, SW64)
constexpr int kStagesP = 2;
constexpr int kABytes = kRows * 64;  // one [128 x 32] bf16 tile
constexpr int kBBytes = kOut * 64;   // one [512 x 32] bf16 tile
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // x_hi, x_lo, w_hi, w_lo = 80 KB
constexpr int kThreadsP = 192;       // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 epilogue
constexpr size_t kSmemP = 1024 + static_cast<size_t>(kStagesP) * kStageBytes + 4 * 32 * 33 * sizeof(float) + 256;
static_assert(kSmemP <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(256) proj_split_kernel(const float* __restrict__ src, int64_t rows, int64_t rows_pad, int k,
                                                         __nv_bfloat16* __restrict__ dst) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one element
    if (i >= rows_pad * k) return;
    const int64_t r = i / k;
    const int c = static_cast<int>(i - r * k);
    const float v = r < rows ? src[i] : 0.0f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
    dst[r * 2 * k + c] = hi;
    dst[r * 2 * k + k + c] = lo;
}

struct ProjArgs {
    const float* bias;        // [512] or nullptr
    int64_t b;                // real rows
    int kblocks;              // K / 32
    int k;                    // K
    float* y;                 // [b, 512] fp32 or nullptr
    __nv_bfloat16* yb;        // [b, 512] bf16 or nullptr
};

__global__ void __launch_bounds__(kThreadsP, 1)
proj_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ProjArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* stage = reinterpret_cast<float*>(smem + kStagesP * kStageBytes);  // [4 warps][32 rows][33]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * 32 * 33);
    uint64_t* full_bar = bars;               // [kStagesP]
    uint64_t* empty_bar = full_bar + kStagesP;
    uint64_t* tfull_bar = empty_bar + kStagesP;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    const uint32_t ring = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row_tile = static_cast<int64_t>(blockIdx.x) * kRows;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_w);
        for (int i = 0; i < kStagesP; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) klf::tmem_alloc1(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- TMA producer: per K block the hi / lo tiles of x (128 rows) and of W (2 x 256 rows) ----
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < a.kblocks; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (elect_one()) {
                const uint32_t base = ring + s * kStageBytes;
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(&full_bar[s], kStageBytes);
                const int c_hi = kb * kKBlock, c_lo = a.k + kb * kKBlock;
                klf::tma_load_2d_1(&map_x, fb, base, c_hi, static_cast<int>(row_tile));
                klf::tma_load_2d_1(&map_x, fb, base + kABytes, c_lo, static_cast<int>(row_tile));
                klf::tma_load_2d_1(&map_w, fb, base + 2 * kABytes, c_hi, 0);
                klf::tma_load_2d_1(&map_w, fb, base + 2 * kABytes + kBBytes / 2, c_hi, 256);
                klf::tma_load_2d_1(&map_w, fb, base + 2 * kABytes + kBBytes, c_lo, 0);
                klf::tma_load_2d_1(&map_w, fb, base + 2 * kABytes + kBBytes + kBBytes / 2, c_lo, 256);
            }
            __syncwarp();
            if (++s == kStagesP) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: 12 MMAs per K block (2 K steps x 2 column halves x {hi.hi, hi.lo, lo.hi}) ----
        constexpr uint32_t IDESC = make_idesc_mn(kRows, 256);
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < a.kblocks; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t base = ring + s * kStageBytes;
            if (elect_one()) {
                const uint64_t x_hi = make_smem_desc(base, 512, 4), x_lo = make_smem_desc(base + kABytes, 512, 4);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint64_t w_hi = make_smem_desc(base + 2 * kABytes + h * (kBBytes / 2), 512, 4);
                    const uint64_t w_lo = make_smem_desc(base + 2 * kABytes + kBBytes + h * (kBBytes / 2), 512, 4);
                    const uint32_t d = tmem_base + h * 256;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {  // K = 16 per instruction: +32 bytes inside the swizzled 64-byte row
                        klf::umma_ss1(d, x_hi + 2 * ks, w_hi + 2 * ks, IDESC, (kb | ks) ? 1u : 0u);
                        klf::umma_ss1(d, x_hi + 2 * ks, w_lo + 2 * ks, IDESC, 1u);
                        klf::umma_ss1(d, x_lo + 2 * ks, w_hi + 2 * ks, IDESC, 1u);
                    }
                }
                klf::umma_commit1(&empty_bar[s]);
                if (kb == a.kblocks - 1) klf::umma_commit1(tfull_bar);
            }
            __syncwarp();
            if (++s == kStagesP) {
                s = 0;
                ph ^= 1;
            }
        }
    } else {
        // ---- epilogue: thread = row; bias, row norm, normalised row out (staged through smem for coalesced stores) ----
        const int quad = warp & 3;
        const int r_local = quad * 32 + lane;
        const int64_t row = row_tile + r_local;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        float* st = stage + (warp - 2) * 32 * 33;
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        float ss = 0.0f;
        for (int c = 0; c < kOut / 32; ++c) {
            float v[32];
            tmem_ld_x32(t_row + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float t = v[j] + (a.bias ? __ldg(a.bias + c * 32 + j) : 0.0f);
                ss = fmaf(t, t, ss);
            }
        }
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
        const int64_t row_w0 = row_tile + quad * 32;        // first row of this warp
        for (int c = 0; c < kOut / 32; ++c) {
            float v[32];
            tmem_ld_x32(t_row + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) st[lane * 33 + j] = (v[j] + (a.bias ? __ldg(a.bias + c * 32 + j) : 0.0f)) * inv;
            __syncwarp();
            for (int rr = 0; rr < 32; ++rr) {
                const int64_t r = row_w0 + rr;
                if (r >= a.b) break;  // warp-uniform
                const float val = st[rr * 33 + lane];
                if (a.y) a.y[r * kOut + c * 32 + lane] = val;
                if (a.yb) a.yb[r * kOut + c * 32 + lane] = __float2bfloat16_rn(val);
            }
            __syncwarp();
        }
        (void)row;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) klf::tmem_dealloc1(tmem_base);
}

static inline size_t proj_workspace_bytes(int64_t b, int in_dim) {
    const int64_t b_pad = (b + kRows - 1) / kRows * kRows;
    return static_cast<size_t>(b_pad + kOut) * 2 * in_dim * sizeof(uint16_t) + 512;
}

static inline bool proj_tc_supported(int in_dim, int out_dim) { return out_dim == kOut && in_dim % kKBlock == 0 && in_dim >= kKBlock; }

static int launch_proj_tc(const float* x, const float* w, const float* bias, int64_t b, int in_dim, float* y, uint16_t* yb,
                          void* workspace, cudaStream_t st) {
    const int64_t b_pad = (b + kRows - 1) / kRows * kRows;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    __nv_bfloat16* ws = xs + b_pad * 2 * in_dim;
    proj_split_kernel<<<static_cast<unsigned>((b_pad * in_dim + 255) / 256), 256, 0, st>>>(x, b, b_pad, in_dim, xs);
    RADAR_CUDA_CHECK(cudaGetLastError());
    proj_split_kernel<<<static_cast<unsigned>((static_cast<int64_t>(kOut) * in_dim + 255) / 256), 256, 0, st>>>(w, kOut, kOut, in_dim, ws);
    RADAR_CUDA_CHECK(cudaGetLastError());
    CUtensorMap map_x, map_w;
    memset(&map_x, 0, sizeof map_x);
    memset(&map_w, 0, sizeof map_w);
    int rc = encode_2d_bf16(&map_x, xs, 2 * in_dim, b_pad, kKBlock, kRows, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = encode_2d_bf16(&map_w, ws, 2 * in_dim, kOut, kKBlock, 256, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    RADAR_CUDA_CHECK(cudaFuncSetAttribute(proj_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemP)));
    ProjArgs a{};
    a.bias = bias; a.b = b; a.kblocks = in_dim / kKBlock; a.k = in_dim; a.y = y; a.yb = reinterpret_cast<__nv_bfloat16*>(yb);
    proj_gemm_kernel<<<static_cast<unsigned>(b_pad / kRows), kThreadsP, kSmemP, st>>>(map_x, map_w, a);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

}  // namespace proj
}  // namespace radar
