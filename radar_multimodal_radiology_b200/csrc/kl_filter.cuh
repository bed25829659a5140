// kl_filter.cuh -- KL observation retrieval for MANY queries (thousands) over a corpus that fits L2: the regime of
// BASELINE config 2 (377 k cases x 65 536 queries).  There the contraction is tiny (K = 16) and the bound is how fast the
// 2.5e10 keys can be pulled out of tensor memory and looked at, so this kernel is built around a LEAN epilogue:
//
//   * CTA pair (cluster of 2, tcgen05 cta_group::2), M = 256 query rows per pair, persistent over work items
//     (query tile x corpus slab).  Both operands come from shared memory (SS form): the query tile [256 x 32 halves]
//     is loaded once per item (double-buffered), corpus tiles [BLOCK_N x K] stream through an 8-slot TMA ring.
//   * BLOCK_N = 64 E corpus rows per tile, 512 / BLOCK_N accumulator stages fill the whole tensor memory.
//   * E epilogue warps per TMEM lane quadrant SPLIT THE COLUMNS of every tile: a warp pulls its 64 columns into
//     registers (two tcgen05.ld per wait -- one load per wait caps TMEM reads at a quarter of the rate,
//     tools/micro/ldtm_bw.cu) and hands the stage back at once, so a stage is occupied for two TMEM reads, not for the
//     filtering of a whole tile; then 32 FMNMX3 reduce the 64 keys of the thread's query row to two maxima, one vote
//     decides whether any of the warp's 32 rows has a survivor, and only then does the (register-resident) chunk go
//     through the rare path: staged in shared memory with a survivor mask, appended to the row's candidate buffer,
//     compacted by the register radix select when the buffer passes 192 entries.
//   * PREPASS instantiation: same pipeline over every tile_stride-th tile, the epilogue only keeps per query row the
//     maximum key of every group of tiles; the k'-th largest group maximum (klf_group_threshold_kernel) is a valid and
//     nearly exact initial threshold for the real pass, which then sees tens of survivors per query, not thousands.
//
// Filter arithmetic (FMT):
//   0  bf16 hi/lo split of both operands, three products (v_hi L_hi + v_hi L_lo + v_lo L_hi) on the [hi|lo] table
//      `klpack` (64 B per case) -- the arithmetic of the hybrid filter;
//   1  fp16, ONE product  fp16(2^13 v) . fp16(2^11 L)  on the fp16 table `kl16` (32 B per case);
//   2  fp16, two products ([v_hi + v_lo] . L16) on the same table.
// The power-of-two scales keep every fp16 operand normal (v >= 1e-8, |L| >= 6e-8), so the result does not depend on how
// the tensor core treats subnormals; accumulators are compared in scaled units and scaled back exactly (x 2^-24).
// |filter key - canonical key| <= qerr is computed per query by klf_pack_kernel (derivation next to it).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"
#include "scan_kernels.cuh"
#include "tc_filter.cuh"

namespace radar {
namespace klf {

using namespace tc;  // PTX wrappers, descriptors, cluster helpers

#ifndef RADAR_KLF_PAIR
#define RADAR_KLF_PAIR 1
#endif
#ifndef RADAR_KLF_CS
#define RADAR_KLF_CS 4
#endif
#ifndef RADAR_KLF_TS
#define RADAR_KLF_TS 1
#endif
#ifndef RADAR_KLF_ONE_THREAD
#define RADAR_KLF_ONE_THREAD 0
#endif
// producer / issuer loops run by ONE thread (no per-tile elect + reconvergence) or by the whole warp with an elected lane
constexpr bool kOneThread = RADAR_KLF_ONE_THREAD != 0;
constexpr bool kPair = RADAR_KLF_PAIR != 0;  // CTA pairs (tcgen05 cta_group::2, M = 256) or single CTAs (M = 128, all hand-offs CTA-local)
constexpr int kCtas = kPair ? 2 : 1;
constexpr int kTileQK = kBlockM * kCtas;   // query rows per work tile
#ifndef RADAR_KLF_CPW
#define RADAR_KLF_CPW 1
#endif
constexpr int kCPW = RADAR_KLF_CPW;        // 64-column chunks a warp reads one after the other from each of its tiles
constexpr int kCS = RADAR_KLF_CS;          // warps per lane quadrant that split the columns of one tile (64 kCPW columns each)
constexpr int kTS = RADAR_KLF_TS;          // tile streams: tile g of a pair is filtered by the warps of stream g mod kTS
constexpr int kE = kCS * kTS;              // epilogue warps per TMEM lane quadrant == candidate buffers per (query, slab)
constexpr int kBlockN = 64 * kCS * kCPW;   // corpus rows per tile (whole pair / CTA)
constexpr int kStages = 512 / kBlockN > 8 ? 8 : 512 / kBlockN;  // accumulator stages
constexpr int kThreadsK = 64 + 128 * kE;   // warp 0 TMA, warp 1 MMA + TMEM alloc, 4 E epilogue warps
// a stream waits for "its" tile on an mbarrier PARITY, which is only sound when the previous phase of that barrier is
// known to be complete: with kStages % kTS == 0 the previous user of a stage is the same stream
static_assert(kStages % kTS == 0, "tile streams must divide the accumulator stages");
constexpr int kSlotsK = 8;                 // corpus tile ring
constexpr int kASlotBytes = kBlockM * 64;  // one CTA's half of a query tile: 128 rows x [hi(16) | lo(16)] halves
constexpr int kMaxKpPrepass = 48;          // the threshold kernel keeps k' values per query in shared memory
constexpr int kMaxGroupsK = 1024;
static_assert(kE >= 1 && kE <= 4 && kBlockN <= 256 && kBlockN % 16 == 0 && kStages >= 2, "KL filter geometry");

constexpr int kFmtBf16x3 = 0, kFmtF16x1 = 1, kFmtF16x2 = 2;
constexpr float kScaleV = 8192.0f, kScaleL = 2048.0f;        // 2^13, 2^11
constexpr float kAccScale = 16777216.0f;                     // 2^24 = kScaleV * kScaleL
constexpr float kAccInv = 1.0f / 16777216.0f;

__host__ __device__ constexpr int row_bytes(int fmt) { return fmt == kFmtBf16x3 ? 64 : 32; }
__host__ __device__ constexpr int slot_bytes(int fmt) { return (kBlockN / kCtas) * row_bytes(fmt); }  // per CTA
__host__ __device__ constexpr size_t smem_bytes(int fmt) {
    return 1024 + static_cast<size_t>(kSlotsK) * slot_bytes(fmt) + 2 * kASlotBytes + 4 * kE * 32 * 32 * sizeof(float) + 1024;
}
static_assert(smem_bytes(kFmtBf16x3) <= 227 * 1024, "shared memory budget");

// kind::f16 instruction descriptor with fp16 operands (format code 0) and fp32 accumulation
__host__ __device__ constexpr uint32_t make_idesc_f16_mn(int m, int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, M = 256 across the CTA pair (each CTA supplies its 128 A rows and half of the B rows)
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// single-CTA flavours of the pair helpers in tc_filter.cuh
__device__ __forceinline__ void umma_ss1(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit1(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_1(const CUtensorMap* map, uint32_t bar_addr, uint32_t dst_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_addr),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc1(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc1(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemCols) : "memory");
}
// the flavour this build uses
__device__ __forceinline__ void klf_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (kPair) umma_ss2(d, a, b, idesc, acc);
    else umma_ss1(d, a, b, idesc, acc);
}
__device__ __forceinline__ void klf_commit(uint64_t* bar) {
    if constexpr (kPair) umma_commit_pair(bar);
    else umma_commit1(bar);
}
__device__ __forceinline__ void klf_tma(const CUtensorMap* map, uint32_t bar_addr, uint32_t dst, int c0, int c1) {
    if constexpr (kPair) tma_load_2d_pair(map, bar_addr, dst, c0, c1);
    else tma_load_2d_1(map, bar_addr, dst, c0, c1);
}

// barrier helpers on precomputed 32-bit shared addresses (the epilogue loop must not re-derive them every tile)
__device__ __forceinline__ bool mbar_try_a(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_a(addr, parity)) {
        if (++spins > kSpinLimit) {
            printf("radar kl_filter: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_cluster_a(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// keeps a loop-invariant value in its register (the compiler otherwise re-materialises address arithmetic inside the
// hot loop to save registers)
__device__ __forceinline__ uint32_t pin(uint32_t x) {
    asm volatile("" : "+r"(x));
    return x;
}

// ---- query pack: [v_hi | v_lo] rows + shift + error bound ------------------------------------------------------
// FMT 1 / 2: v_hi = fp16(2^13 v), v_lo = fp16(2^13 v - v_hi); the corpus table holds fp16(2^11 L).  With v in {0} u
// [1e-8, 1] and |L| in {0} u [5.9e-8, 18.5] every non-zero v_hi and table entry is a NORMAL fp16 number, so both carry a
// relative error <= 2^-11 and their products are exact in fp32.  Per (query, case):
//   FMT 1:  |sum v_hi L16 - sum v L| <= (2^-11 + 2^-11 + 2^-22) sum |v||L|                        = 9.77e-4 kl_mag
//   FMT 2:  v_hi + v_lo = v (1 + 2^-22) or, when the residual underflows fp16 (|2^13 v - v_hi| < 6.1e-5), off by
//           < 6.1e-5 * 2^-13 = 7.5e-9 per observation:  (2^-11 + 2^-21) sum |v||L| + 7.5e-9 sum_j max|L_j|
// plus the fp32 accumulation of the tensor core and of the canonical chain (budget 2^-13 of the magnitudes, as for the
// bf16 filter) and the roundings of the final combine / shift (1e-6 of the magnitudes).  kl_mag = sum_j |v_j| max_n |L_nj|.
struct KlPackArgs {
    const float* p16;
    const float* entropy;
    int64_t q, q_pad;
    int fmt;
    float logq_col_max[kObsPad];
    uint16_t* apack;   // [q_pad][32] halves
    float* qshift;     // [q_pad]
    float* qerr;       // [q_pad]
};

__global__ void __launch_bounds__(256) klf_pack_kernel(const KlPackArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= a.q_pad) return;
    const bool valid = row < a.q;
    float sv = 0.0f, sl = 0.0f;
    if (lane < kObsPad) {
        const float v = valid ? a.p16[row * kObsPad + lane] : 0.0f;
        const float vs = v * kScaleV;  // exact
        const __half hi = __float2half_rn(vs);
        const __half lo = __float2half_rn(__fsub_rn(vs, __half2float(hi)));
        reinterpret_cast<__half*>(a.apack)[row * 32 + lane] = hi;
        reinterpret_cast<__half*>(a.apack)[row * 32 + kObsPad + lane] = lo;
        sv = fabsf(v) * a.logq_col_max[lane];
        sl = v != 0.0f ? a.logq_col_max[lane] : 0.0f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sl += __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if (lane == 0) {
        const float shift = valid ? a.entropy[row] : 0.0f;
        const float rel = a.fmt == kFmtF16x1 ? 1.13e-3f : 6.3e-4f;  // (9.77e-4 | 4.89e-4) + 2^-13, a few % of slack
        a.qshift[row] = shift;
        a.qerr[row] = rel * sv + 8e-9f * sl + 1e-6f * (fabsf(shift) + sv) + 1e-30f;
    }
}

// ---- the filter kernel -----------------------------------------------------------------------------------------
struct KlfArgs {
    const float* qshift;
    int64_t q, q_tiles, n;
    int parts;
    int64_t rows_per_part;  // multiple of kBlockN
    int kp;
    uint64_t* cand;         // [q_pad][parts][kE][kCandCap]
    uint32_t* cnt;          // [q_pad][parts][kE]
    float* thr;             // [q_pad][parts][kE]
    uint32_t* gthr;         // [q_pad] best published threshold per query (ord-encoded key units, 0 = none)
    int gthr_init;          // gthr holds prepass thresholds: read it even with a single slab
    // prepass
    uint32_t* groupmax;     // [q_tiles][groups][256] ord-encoded group maxima (0 = none)
    int groups, groups_per_slab, group_tiles, tile_stride;
    unsigned long long* clk;  // [2] SM cycles / nanoseconds CTA 0 spent in the kernel
};

template <int FMT, bool PREPASS>
__global__ void __launch_bounds__(kThreadsK, 1)
klf_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c, const KlfArgs a) {
    constexpr int LOAD_N = kBlockN / kCtas;
    constexpr int SLOT = slot_bytes(FMT);
    constexpr uint32_t IDESC = FMT == kFmtBf16x3 ? make_idesc_mn(kTileQK, kBlockN) : make_idesc_f16_mn(kTileQK, kBlockN);
    constexpr float SCALE = FMT == kFmtBf16x3 ? 1.0f : kAccScale;
    constexpr float INV = FMT == kFmtBf16x3 ? 1.0f : kAccInv;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* abuf = smem + kSlotsK * SLOT;                                         // [2][128 rows x 64 B], SW64
    float* stage = reinterpret_cast<float*>(abuf + 2 * kASlotBytes);               // [4 E warps][32 columns][32 lanes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * kE * 32 * 32);
    uint64_t* full_bar = bars;                    // [kSlotsK]  (leader's are waited on)
    uint64_t* empty_bar = full_bar + kSlotsK;     // [kSlotsK]
    uint64_t* tfull_bar = empty_bar + kSlotsK;    // [kStages]
    uint64_t* tempty_bar = tfull_bar + 8;         // [kStages]  (leader's, 8 E arrivals)
    uint64_t* afull_bar = tempty_bar + 8;         // [2]        (leader's)
    uint64_t* afree_bar = afull_bar + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afree_bar + 2);
    const uint32_t ring_addr = smem_u32(smem), abuf_addr = smem_u32(abuf);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int64_t unit = kPair ? blockIdx.x >> 1 : blockIdx.x, units = kPair ? gridDim.x >> 1 : gridDim.x;
    const int64_t items = a.q_tiles * a.parts;
    const int64_t tile_step = static_cast<int64_t>(kBlockN) * a.tile_stride;

    unsigned long long clk0 = 0, ns0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_q);
        prefetch_tmap(&map_c);
        for (int i = 0; i < kSlotsK; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4 * kCS * kCtas);  // warps reading one stage: 4 quadrants x kCS (x 2 CTAs)
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&afull_bar[i], 1);
            mbar_init(&afree_bar[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (kPair) tmem_alloc_pair(tmem_slot);
        else tmem_alloc1(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) {
        if (threadIdx.x == 0) printf("radar kl_filter: unexpected TMEM base %u\n", tmem_base);
        __trap();
    }

    if (warp == 0) {
        // ================================ TMA producer (every CTA: its halves) ================================
        // ONE thread runs the whole loop (no per-tile elect / reconvergence): a tile costs a barrier wait, an expect_tx
        // and a bulk copy
        if (!kOneThread || lane == 0) {
            uint32_t slot = 0, sph = 0, item_no = 0;
            for (int64_t item = unit; item < items; item += units, ++item_no) {
                const int64_t qtile = item % a.q_tiles;
                const int part = static_cast<int>(item / a.q_tiles);
                const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
                const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
                const uint32_t ntiles = static_cast<uint32_t>((row_end - row_begin + tile_step - 1) / tile_step);
                const uint32_t ab = item_no & 1u;
                if (item_no >= 2) mbar_wait(&afree_bar[ab], ((item_no >> 1) - 1u) & 1u);  // MMAs of item_no - 2 are done
                if (kOneThread || elect_one()) {
                    if (leader) mbar_expect_tx(&afull_bar[ab], kCtas * kASlotBytes);
                    klf_tma(&map_q, smem_u32(&afull_bar[ab]), abuf_addr + ab * kASlotBytes, 0,
                            static_cast<int>(qtile * kTileQK) + static_cast<int>(cta_rank) * kBlockM);
                }
                if (!kOneThread) __syncwarp();
                int row = static_cast<int>(row_begin) + static_cast<int>(cta_rank) * LOAD_N;
                for (uint32_t j = 0; j < ntiles; ++j, row += static_cast<int>(tile_step)) {
                    mbar_wait(&empty_bar[slot], sph ^ 1);
                    if (kOneThread || elect_one()) {
                        if (leader) mbar_expect_tx(&full_bar[slot], kCtas * SLOT);
                        klf_tma(&map_c, smem_u32(&full_bar[slot]), ring_addr + slot * SLOT, 0, row);
                    }
                    if (!kOneThread) __syncwarp();
                    if (++slot == kSlotsK) {
                        slot = 0;
                        sph ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA): one thread ================================
        if (leader && (!kOneThread || lane == 0)) {
            uint32_t slot = 0, sph = 0, as = 0, aph = 0, item_no = 0;
            for (int64_t item = unit; item < items; item += units, ++item_no) {
                const int part = static_cast<int>(item / a.q_tiles);
                const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
                const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
                const uint32_t ntiles = static_cast<uint32_t>((row_end - row_begin + tile_step - 1) / tile_step);
                const uint32_t ab = item_no & 1u;
                mbar_wait(&afull_bar[ab], (item_no >> 1) & 1u);
                tc_fence_after();
                const uint64_t a_desc = make_smem_desc(abuf_addr + ab * kASlotBytes, 512, 4);  // SW64: hi at +0, lo at +32 B
                const uint64_t b_desc0 = FMT == kFmtBf16x3 ? make_smem_desc(ring_addr, 512, 4)    // SW64 rows of 64 B
                                                           : make_smem_desc(ring_addr, 256, 6);   // SW32 rows of 32 B
                for (uint32_t j = 0; j < ntiles; ++j) {
                    mbar_wait(&tempty_bar[as], aph ^ 1);
                    mbar_wait(&full_bar[slot], sph);
                    tc_fence_after();
                    const uint32_t d = as * kBlockN;
                    const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((slot * SLOT) >> 4);
                    if (kOneThread || elect_one()) {
                        if (FMT == kFmtBf16x3) {
                            klf_mma(d, a_desc, b_desc, IDESC, 0u);           // v_hi . L_hi
                            klf_mma(d, a_desc, b_desc + 2, IDESC, 1u);       // v_hi . L_lo
                            klf_mma(d, a_desc + 2, b_desc, IDESC, 1u);       // v_lo . L_hi
                        } else {
                            klf_mma(d, a_desc, b_desc, IDESC, 0u);           // v_hi . L16
                            if (FMT == kFmtF16x2) klf_mma(d, a_desc + 2, b_desc, IDESC, 1u);  // v_lo . L16
                        }
                        klf_commit(&empty_bar[slot]);
                        klf_commit(&tfull_bar[as]);
                    }
                    if (!kOneThread) __syncwarp();
                    if (++slot == kSlotsK) {
                        slot = 0;
                        sph ^= 1;
                    }
                    if (++as == kStages) {
                        as = 0;
                        aph ^= 1;
                    }
                }
                if (kOneThread || elect_one()) klf_commit(&afree_bar[ab]);  // the query buffer may be reloaded (both CTAs)
                if (!kOneThread) __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue: E warps per lane quadrant split the columns ================================
        const int w = warp - 2;
        const int quad = warp & 3;         // TMEM lane quadrant this warp may access
        const int e = w >> 2;              // warp set: candidate buffer / group slot of this warp
        const int ts = e / kCS, cs = e % kCS;  // tile stream, 64-column slice of the stream's tiles
        const int r_in_tile = static_cast<int>(cta_rank) * kBlockM + quad * 32 + lane;
        const uint32_t tacc0 = pin(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + cs * 64 * kCPW);  // this warp's lanes and columns
        const uint32_t tfull_a = pin(smem_u32(tfull_bar));
        const uint32_t tempty_a = pin(kPair ? smem_u32(tempty_bar) & kPeerBitMask : smem_u32(tempty_bar));  // the leader CTA's barriers
        float* my_stage = stage + w * 32 * 32 + lane;  // [column * 32]: bank == lane
        const uint32_t lane0 = pin(lane == 0 ? 1u : 0u);
        uint32_t g = 0;  // running tile number of this pair over all its items (accumulator stage g mod kStages)
        for (int64_t item = unit; item < items; item += units) {
            const int64_t qtile = item % a.q_tiles;
            const int part = static_cast<int>(item / a.q_tiles);
            const int64_t qrow = qtile * kTileQK + r_in_tile;
            const bool valid = qrow < a.q;
            const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
            const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
            const float shift = a.qshift[qrow];
            // ---- real pass state ----
            float thr = -CUDART_INF_F;       // key units
            float thr_acc = CUDART_INF_F;    // accumulator units (scaled); +inf: padding rows never survive
            int cnt = 0;
            const int64_t slot_idx = (qrow * a.parts + part) * kE + e;
            uint64_t* buf = a.cand + slot_idx * kCandCap;
            if (!PREPASS) {
                const uint32_t g0 = (a.parts > 1 || a.gthr_init) ? *reinterpret_cast<volatile const uint32_t*>(a.gthr + qrow) : 0u;
                if (g0) thr = ord2f(g0);
                if (valid) {
                    if (g0) {
                        // an inherited threshold is fl(acc * 2^-24 - shift) of some case; fl(thr + shift) may round above that
                        // case's accumulator, and cases tied with it would be lost: step a few ulps down (always safe)
                        const float c = __fadd_rn(thr, shift);
                        thr_acc = (c - 4e-7f * (fabsf(c) + fabsf(shift))) * SCALE;
                    } else {
                        thr_acc = -CUDART_INF_F;
                    }
                }
            }
            // ---- prepass state ----
            float gmax = -CUDART_INF_F;
            int tg = 0;
            uint32_t gbound = static_cast<uint32_t>(a.group_tiles);  // first tile (slab-local number) of the next tile group
            uint32_t* gdst = PREPASS ? a.groupmax + (qtile * a.groups + static_cast<int64_t>(part) * a.groups_per_slab + e) * kTileQK + r_in_tile
                                     : nullptr;

            const uint32_t ntiles = static_cast<uint32_t>((row_end - row_begin + tile_step - 1) / tile_step);
            // first tile of this item that belongs to this warp's stream
            for (uint32_t j = (static_cast<uint32_t>(ts) + kTS - (g % kTS)) % kTS; j < ntiles; j += kTS) {
                const uint32_t gt = g + j;
                const uint32_t as = gt % kStages, aph = (gt / kStages) & 1u;
                mbar_wait_a(tfull_a + as * 8, aph);
                tc_fence_after();
                const uint32_t t_acc = tacc0 + as * kBlockN;
                float tile_m = -CUDART_INF_F;  // prepass: maximum over this warp's columns of the tile
#pragma unroll
                for (int cw = 0; cw < kCPW; ++cw) {
                    const int col0 = (cs * kCPW + cw) * 64;  // first tile column of this chunk
                    float v[64];
                    tmem_ld_x32(t_acc + cw * 64, v);
                    tmem_ld_x32(t_acc + cw * 64 + 32, v + 32);
                    tmem_wait_ld();
                    if (cw == kCPW - 1) {  // the stage is free once the last chunk is in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane0) mbar_arrive_cluster_a(tempty_a + as * 8);
                    }
                    // four independent chains of 16 keys each (FMNMX3: two keys per instruction)
                    float ch[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float m = v[16 * c];
#pragma unroll
                        for (int jj = 1; jj < 15; jj += 2) m = fmaxf(fmaxf(m, v[16 * c + jj]), v[16 * c + jj + 1]);
                        ch[c] = fmaxf(m, v[16 * c + 15]);
                    }
                    if (PREPASS) {
                        float m = fmaxf(fmaxf(ch[0], ch[1]), fmaxf(ch[2], ch[3]));
                        if (j + 1 == ntiles) {  // last tile of the slab: rows past the end were zero-filled by TMA (key 0 beats them all)
                            const int64_t rowc = row_begin + static_cast<int64_t>(j) * tile_step + col0;
                            if (rowc + 64 > row_end) {
                                m = -CUDART_INF_F;
#pragma unroll
                                for (int jj = 0; jj < 64; ++jj)
                                    if (rowc + jj < row_end) m = fmaxf(m, v[jj]);
                            }
                        }
                        tile_m = fmaxf(tile_m, m);
                        continue;
                    }
                    if (__any_sync(0xffffffffu, fmaxf(fmaxf(ch[0], ch[1]), fmaxf(ch[2], ch[3])) >= thr_acc)) {
                        // rare path: per 16-column quarter that holds a survivor of some lane, stage the quarter in this warp's
                        // shared-memory scratch with a survivor bit mask, then walk the set bits
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (!__any_sync(0xffffffffu, ch[c] >= thr_acc)) continue;
                            uint32_t mask = 0;
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                my_stage[jj * 32] = v[16 * c + jj];
                                mask |= (v[16 * c + jj] >= thr_acc ? 1u : 0u) << jj;
                            }
                            const int64_t row_base = row_begin + static_cast<int64_t>(j) * tile_step + col0 + 16 * c;
                            while (mask) {
                                const int jj = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const int64_t row = row_base + jj;
                                if (row < row_end)
                                    buf[cnt++] = make_composite(__fsub_rn(my_stage[jj * 32] * INV, shift), static_cast<uint32_t>(row));
                            }
                            __syncwarp();
                            unsigned need = __ballot_sync(0xffffffffu, cnt > kCandSoft);
                            while (need) {
                                const int src_lane = __ffs(need) - 1;
                                need &= need - 1;
                                uint64_t* b = reinterpret_cast<uint64_t*>(
                                    __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(buf), src_lane));
                                const int n_src = __shfl_sync(0xffffffffu, cnt, src_lane);
                                const float t = warp_compact(b, n_src, a.kp, lane);
                                if (lane == src_lane) {
                                    thr = t;
                                    thr_acc = __fadd_rn(t, shift) * SCALE;  // own threshold: every kept key was fl(acc - shift)
                                    cnt = a.kp;
                                }
                            }
                        }
                    }
                }
                if (PREPASS) {
                    while (j >= gbound) {  // tile j opens a later tile group: the finished ones get their (possibly empty) maximum
                        gdst[static_cast<int64_t>(tg) * kE * kTileQK] =
                            (valid && gmax > -CUDART_INF_F) ? f2ord(__fsub_rn(gmax * INV, shift)) : 0u;
                        gmax = -CUDART_INF_F;
                        gbound += static_cast<uint32_t>(a.group_tiles);
                        ++tg;
                    }
                    gmax = fmaxf(gmax, tile_m);
                }
            }
            g += ntiles;
            if (PREPASS) {
                // the group in progress, then "no value" for the group slots this warp's tiles never reached (short last slab)
                for (; tg * kE < a.groups_per_slab; ++tg) {
                    gdst[static_cast<int64_t>(tg) * kE * kTileQK] =
                        (valid && gmax > -CUDART_INF_F) ? f2ord(__fsub_rn(gmax * INV, shift)) : 0u;
                    gmax = -CUDART_INF_F;
                }
                continue;
            }
            a.cnt[slot_idx] = valid ? static_cast<uint32_t>(cnt) : 0u;
            a.thr[slot_idx] = thr;
            if (a.parts > 1 && valid && thr > -CUDART_INF_F) atomicMax(a.gthr + qrow, f2ord(thr));
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still touch it
    if (warp == 1) {
        if constexpr (kPair) tmem_dealloc_pair(tmem_base);
        else tmem_dealloc1(tmem_base);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.clk) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        a.clk[0] = clock64() - clk0;
        a.clk[1] = ns1 - ns0;
    }
}

// k'-th largest group maximum per query -> gthr (0 = no threshold).  One CTA per query tile, one thread per query:
// the group maxima are read coalesced ([group][256 rows]); the running best k' live in shared memory (column = thread).
// (Measured alternatives on 65 536 queries x 496 groups: this kernel 0.136 ms; a coalesced shared-memory tile + bit-wise radix
// descent per warp 0.297 ms; warp per query with strided loads + k' maximum extractions 0.167 ms; tile + extractions 0.197 ms.)
__global__ void __launch_bounds__(kTileQK) klf_group_threshold_kernel(const uint32_t* __restrict__ groupmax, int64_t q,
                                                                     int groups, int kp, uint32_t* __restrict__ gthr) {
    extern __shared__ uint32_t top[];  // [kp][256]
    const int t = threadIdx.x;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * kTileQK + t;
    const uint32_t* src = groupmax + static_cast<int64_t>(blockIdx.x) * groups * kTileQK + t;
    uint32_t minv = 0xFFFFFFFFu;
    int minpos = 0, have = 0;
    constexpr int kBatch = 8;  // independent loads in flight per thread (the insertion logic below is a dependent chain)
    for (int g0 = 0; g0 < groups; g0 += kBatch) {
        uint32_t vb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) vb[u] = g0 + u < groups ? __ldcs(src + static_cast<int64_t>(g0 + u) * kTileQK) : 0u;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (g0 + u >= groups) break;
            const uint32_t v = vb[u];
            if (have < kp) {
                top[have * kTileQK + t] = v;
                if (v < minv) {
                    minv = v;
                    minpos = have;
                }
                ++have;
            } else if (v > minv) {
                top[minpos * kTileQK + t] = v;
                minv = 0xFFFFFFFFu;
                for (int i = 0; i < kp; ++i) {
                    const uint32_t u2 = top[i * kTileQK + t];
                    if (u2 < minv) {
                        minv = u2;
                        minpos = i;
                    }
                }
            }
        }
    }
    if (qi < q) gthr[qi] = have == kp ? minv : 0u;
}

// ---- host side ---------------------------------------------------------------------------------------------
struct KlfLaunch {
    const radar_corpus_t* corpus;
    const radar_queries_t* queries;
    int fmt;
    int64_t q, q_tiles;
    int parts;
    int64_t rows_per_part;
    int kp;
    uint64_t* cand;
    uint32_t* cnt;
    float* thr;
    uint32_t* gthr;   // region of gthr_region_bytes(q_pad): [q_pad] thresholds, then the {cycles, ns} pair (zeroed here)
    float* qerr;
    uint16_t* apack;  // [q_pad][32] halves followed (256-byte aligned) by qshift [q_pad]
    int units;
    uint32_t* groupmax;
    int groups, group_tiles, tile_stride, groups_per_slab;
    cudaEvent_t ev_start, ev_stop;
    unsigned long long* clk_dev;
};

template <int FMT, bool PREPASS>
static int launch_klf_mode(const KlfLaunch& fl, const KlfArgs& fa, cudaStream_t st) {
    CUtensorMap map_q, map_c;
    memset(&map_q, 0, sizeof map_q);
    memset(&map_c, 0, sizeof map_c);
    const int64_t q_pad = fl.q_tiles * kTileQK;
    int rc = encode_2d_bf16(&map_q, fl.apack, 32, q_pad, 32, kBlockM, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    if (FMT == kFmtBf16x3)
        rc = encode_2d_bf16(&map_c, fl.corpus->klpack, RADAR_KLPACK, fl.corpus->n, RADAR_KLPACK, kBlockN / kCtas, CU_TENSOR_MAP_SWIZZLE_64B);
    else
        rc = encode_2d_bf16(&map_c, fl.corpus->kl16, kObsPad, fl.corpus->n, kObsPad, kBlockN / kCtas, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc) return rc;
    RADAR_CUDA_CHECK(cudaFuncSetAttribute(klf_kernel<FMT, PREPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem_bytes(FMT))));
    const int64_t items = fl.q_tiles * fl.parts;
    int64_t units = fl.units < 1 ? 1 : fl.units;
    if (units > items) units = items;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(units * kCtas));
    cfg.blockDim = dim3(kThreadsK);
    cfg.dynamicSmemBytes = smem_bytes(FMT);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RADAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, klf_kernel<FMT, PREPASS>, map_q, map_c, fa));
    return RADAR_OK;
}

template <bool PREPASS>
static int launch_klf_fmt(const KlfLaunch& fl, const KlfArgs& fa, cudaStream_t st) {
    if (fl.fmt == kFmtBf16x3) return launch_klf_mode<kFmtBf16x3, PREPASS>(fl, fa, st);
    if (fl.fmt == kFmtF16x1) return launch_klf_mode<kFmtF16x1, PREPASS>(fl, fa, st);
    return launch_klf_mode<kFmtF16x2, PREPASS>(fl, fa, st);
}

// pack -> [prepass -> thresholds] -> filter.  The profiled span (ev_start .. ev_stop) covers prepass, threshold selection
// and the real pass.
static int launch_kl_filter(KlfLaunch& fl, cudaStream_t st, int* launches) {
    const int64_t q_pad = fl.q_tiles * kTileQK;
    const size_t shift_off = (sizeof(uint16_t) * static_cast<size_t>(q_pad) * 32 + 255) / 256 * 256;
    float* qshift = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fl.apack) + shift_off);
    if (fl.fmt == kFmtBf16x3) {
        PackArgs pa{};
        pa.q_emb = nullptr; pa.p16 = fl.queries->p16; pa.entropy = fl.queries->entropy; pa.q = fl.q; pa.q_pad = q_pad;
        pa.d = fl.corpus->d; pa.mode = RADAR_MODE_KL; pa.alpha = 0.0f; pa.oma = 1.0f;
        pa.emb_max_norm = fl.corpus->emb_max_norm; pa.logq_max_abs = fl.corpus->logq_max_abs;
        fill_col_max(fl.corpus, pa.logq_col_max);
        pa.apack = fl.apack; pa.qshift = qshift; pa.qerr = fl.qerr;
        query_pack_kernel<<<static_cast<unsigned>((q_pad * 32 + 255) / 256), 256, 0, st>>>(pa);
    } else {
        KlPackArgs pa{};
        pa.p16 = fl.queries->p16; pa.entropy = fl.queries->entropy; pa.q = fl.q; pa.q_pad = q_pad; pa.fmt = fl.fmt;
        fill_col_max(fl.corpus, pa.logq_col_max);
        pa.apack = fl.apack; pa.qshift = qshift; pa.qerr = fl.qerr;
        klf_pack_kernel<<<static_cast<unsigned>((q_pad * 32 + 255) / 256), 256, 0, st>>>(pa);
    }
    RADAR_CUDA_CHECK(cudaGetLastError());
    RADAR_CUDA_CHECK(cudaMemsetAsync(fl.gthr, 0, gthr_region_bytes(q_pad), st));
    KlfArgs fa{};
    fa.qshift = qshift; fa.q = fl.q; fa.q_tiles = fl.q_tiles; fa.n = fl.corpus->n; fa.parts = fl.parts;
    fa.rows_per_part = fl.rows_per_part; fa.kp = fl.kp; fa.cand = fl.cand; fa.cnt = fl.cnt; fa.thr = fl.thr;
    fa.gthr = fl.gthr; fa.tile_stride = 1; fa.group_tiles = 1;
    fa.clk = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(fl.gthr) +
                                                   (sizeof(uint32_t) * static_cast<size_t>(q_pad) + 7) / 8 * 8) + kMaxUnits;
    fl.clk_dev = fa.clk;
    *launches = 2;
    int rc;
    if (fl.ev_start) RADAR_CUDA_CHECK(cudaEventRecord(fl.ev_start, st));
    if (fl.groups > 0) {
        KlfArgs fp = fa;
        fp.groupmax = fl.groupmax; fp.groups = fl.groups; fp.group_tiles = fl.group_tiles; fp.tile_stride = fl.tile_stride;
        fp.groups_per_slab = fl.groups_per_slab; fp.clk = nullptr;
        rc = launch_klf_fmt<true>(fl, fp, st);
        if (rc) return rc;
        RADAR_CUDA_CHECK(cudaFuncSetAttribute(klf_group_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(sizeof(uint32_t) * kMaxKpPrepass * kTileQK)));
        klf_group_threshold_kernel<<<static_cast<unsigned>(fl.q_tiles), kTileQK, sizeof(uint32_t) * fl.kp * kTileQK, st>>>(
            fl.groupmax, fl.q, fl.groups, fl.kp, fl.gthr);
        RADAR_CUDA_CHECK(cudaGetLastError());
        fa.gthr_init = 1;
        *launches += 2;
    }
    rc = launch_klf_fmt<false>(fl, fa, st);
    if (rc) return rc;
    if (fl.ev_stop) RADAR_CUDA_CHECK(cudaEventRecord(fl.ev_stop, st));
    return RADAR_OK;
}

}  // namespace klf
}  // namespace radar
