// tc_filter.cuh -- tcgen05 (5th-gen tensor core) filter kernel for sm_100a.
//
// One CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256) per two SMs, persistent over work items
// (query tile of 256 rows, corpus slab).
//   * the QUERY tile is the A operand and lives in TENSOR MEMORY for the whole slab sweep
//     (bf16 pairs packed in 32-bit TMEM columns, written once per item with tcgen05.st) -- shared
//     memory holds nothing but the streamed corpus tiles;
//   * CORPUS tiles are the B operand: TMA (cp.async.bulk.tensor.2d, 128B/64B swizzle) streams each CTA's
//     half (BLOCK_N/2 rows) of every tile -- d/64 K-blocks of the embedding matrix plus one block of the
//     [log q hi | log q lo] pack -- into a ring of TILE-ALIGNED shared-memory slots, so that every
//     shared-memory matrix descriptor is "slot base + compile-time constant" and the issue loop is
//     straight-line code (one mbarrier wait per two K-blocks);
//   * one elected thread of the leader CTA issues tcgen05.mma.cta_group::2.kind::f16 (A from TMEM, B from
//     smem descriptors) into one of ACC_STAGES fp32 accumulators in TMEM;
//   * four epilogue warps per CTA read the accumulator back with tcgen05.ld (one query row per thread),
//     compare against the row's running threshold and append survivors to the row's candidate buffer
//     in global memory -- the Q x N score matrix never exists outside TMEM;
//   * pairs that sweep the same slab keep within `window` tiles of each other (a progress word per pair,
//     polled by the TMA producer), so a corpus tile is fetched from HBM once and then served from L2 to
//     every query tile instead of once per query tile.
//
// Filter value for (query i, case n):
//     DPR    f = sum_t bf16(e_q[t]) bf16(e_c[t])
//     KL     f = sum_j (p_hi L_hi + p_hi L_lo + p_lo L_hi)[j]                     (lo*lo dropped: < 2^-18 |p||L|)
//     hybrid f = sum_t bf16(alpha e_q[t]) bf16(e_c[t]) + sum_j (v_hi L_hi + v_hi L_lo + v_lo L_hi)[j],  v = (1-alpha) p
// and the ranking key handed to the candidate buffer is f - shift_i, shift = H (KL) / (1-alpha) H (hybrid),
// i.e. it approximates the canonical key of common.cuh within qerr_i (see query_pack_kernel).
#pragma once
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "scan_kernels.cuh"

namespace radar {
namespace tc {

// Bring-up / measurement hooks (dense key dump, epilogue and TMA knock-outs, window switch) exist only in the
// RADAR_DEBUG flavour of the library (libradar_retrieval_dbg.so, built for the tests); in the release build the
// expressions below are the constant 0 and the code behind them is removed by the compiler.
#ifdef RADAR_DEBUG
#define RADAR_DBG(expr) (expr)
#else
#define RADAR_DBG(expr) (0)
#endif

constexpr int kBlockM = 128;          // query rows per CTA == TMEM lanes
constexpr int kTileQ = 2 * kBlockM;   // query rows per work tile (CTA pair)
constexpr int kThreads = 192;         // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue (+ warps 6..9 in KL mode)
constexpr int kTmemCols = 512;
constexpr int kAIpCols = 256;         // TMEM columns reserved for the embedding part of A (d <= 512)
constexpr int kRingBytes = 192 * 1024;  // shared-memory ring of tile slots (per CTA); KL tiles are small: 96 KB there
constexpr int kMaxSlots = 16;
constexpr int kMaxGroups = 4;         // full barriers per slot: one per two K-blocks
constexpr int kMaxUnits = 1024;       // CTA pairs a launch may use (progress words in the workspace)
constexpr int kProgressEvery = 8;     // tiles between progress publications / window checks
constexpr uint32_t kSpinLimit = 1u << 27;
constexpr uint32_t kWindowSpinLimit = 1u << 16;  // polls (x ~1 us) before a pair stops honouring the window

// KL mode geometry.  A KL tile is three K = 16 MMAs (~80 cycles each at N = 160), so there the bound is not the tensor
// pipe but (a) how long an accumulator stage stays occupied -- commit -> mbarrier -> tcgen05.ld -> remote arrive is ~1000
// cycles, of which reading the stage back is the part that can be shortened -- and (b) the threshold filter itself.
// Measured (tools/micro/ldtm_bw.cu): a tcgen05.wait::ld after every single load caps TMEM reads at ~128 B/cycle/SM;
// two loads per wait and 3-4 warps per lane quadrant reach ~470 B/cycle/SM.  Hence:
//   * kKlStages accumulator stages of kKlBlockN columns (as much of the 512 TMEM columns as possible in flight);
//   * the epilogue warp sets (four warps each, one per lane quadrant) form kKlTileStreams x kKlColSplit: tile g of a
//     pair is filtered by the kKlColSplit sets of stream g mod kKlTileStreams, each reading its own column range with
//     two tcgen05.ld per wait -- a stage is read back by several warps at once (short occupancy) while other streams
//     are still filtering the previous tiles (no warp sits in the hand-off chain of every tile).
#ifndef RADAR_KL_BLOCK_N
#define RADAR_KL_BLOCK_N 160
#endif
#ifndef RADAR_KL_STAGES
#define RADAR_KL_STAGES 3
#endif
#ifndef RADAR_KL_TILE_STREAMS
#define RADAR_KL_TILE_STREAMS 3
#endif
#ifndef RADAR_KL_COL_SPLIT
#define RADAR_KL_COL_SPLIT 1
#endif
constexpr int kKlBlockN = RADAR_KL_BLOCK_N;
constexpr int kKlStages = RADAR_KL_STAGES;
constexpr int kKlTileStreams = RADAR_KL_TILE_STREAMS;
constexpr int kKlColSplit = RADAR_KL_COL_SPLIT;
constexpr int kKlSets = kKlTileStreams * kKlColSplit;
constexpr int kKlAccCol0 = 32;        // the packed [v_hi | v_lo] query rows occupy TMEM columns 0..15
static_assert(kKlBlockN % 16 == 0 && kKlBlockN >= 32 && kKlBlockN <= 256, "KL tile width");
static_assert((kKlBlockN / kKlColSplit) % 16 == 0 && kKlBlockN % kKlColSplit == 0, "KL column split");
static_assert(kKlAccCol0 + kKlStages * kKlBlockN <= 512, "KL accumulator stages exceed TMEM");
static_assert(kKlSets >= 1 && kKlSets <= 5 && kKlStages <= 8, "KL warp sets");
// a stream waits for "its" tile on an mbarrier PARITY: it must not get two phases ahead of the barrier, which holds
// because the tiles complete in order and a stream that is done with tile g - TS knows tile g - STAGES is complete
static_assert(kKlTileStreams <= kKlStages, "KL tile streams may not outnumber the accumulator stages");

__host__ __device__ constexpr int block_n_for_mode(int mode) {
    return mode == RADAR_MODE_HYBRID ? 112 : (mode == RADAR_MODE_KL ? kKlBlockN : 128);
}
__host__ __device__ constexpr int acc_stages_for_mode(int mode) { return mode == RADAR_MODE_KL ? kKlStages : 2; }
// Epilogue warp sets (four warps each, one per TMEM lane quadrant).  DPR / hybrid: one set (the tensor pipe is the bound).
// KL: tile streams x column splits (above); every set keeps its own candidate buffers and thresholds (select_kernel
// merges them; the sets of a CTA share their best thresholds through shared memory).
__host__ __device__ constexpr int epi_sets_for_mode(int mode) { return mode == RADAR_MODE_KL ? kKlSets : 1; }
__host__ __device__ constexpr int threads_for_mode(int mode) { return 64 + 128 * epi_sets_for_mode(mode); }
__host__ __device__ constexpr int acc_col0_for_mode(int mode) {
    return mode == RADAR_MODE_DPR ? kAIpCols : (mode == RADAR_MODE_HYBRID ? kAIpCols + 32 : kKlAccCol0);
}
__host__ __device__ constexpr int a_kl_col_for_mode(int mode) { return mode == RADAR_MODE_HYBRID ? kAIpCols : 0; }
// bf16 elements per packed query row
__host__ __device__ inline int a_cols_for(int mode, int d) {
    return (mode != RADAR_MODE_KL ? d : 0) + (mode != RADAR_MODE_DPR ? 2 * kObsPad : 0);
}
// shared-memory geometry of one tile slot (per CTA: BLOCK_N/2 corpus rows)
__host__ __device__ constexpr int sub_ip_bytes(int mode) { return block_n_for_mode(mode) / 2 * 128; }  // one K-block
__host__ __device__ constexpr int sub_kl_bytes(int mode) { return block_n_for_mode(mode) / 2 * 64; }
__host__ __device__ constexpr int slot_stride_for(int mode, int kblocks) {
    return (kblocks * sub_ip_bytes(mode) + (mode != RADAR_MODE_DPR ? sub_kl_bytes(mode) : 0) + 1023) / 1024 * 1024;
}
__host__ __device__ constexpr int ring_bytes_for(int mode) { return mode == RADAR_MODE_KL ? 96 * 1024 : kRingBytes; }
__host__ __device__ constexpr int slots_for(int mode, int kblocks) {
    return ring_bytes_for(mode) / slot_stride_for(mode, kblocks) > kMaxSlots
               ? kMaxSlots
               : ring_bytes_for(mode) / slot_stride_for(mode, kblocks);
}
__host__ __device__ constexpr int groups_for(int kblocks) { return kblocks <= 1 ? 1 : (kblocks + 1) / 2; }

// dynamic shared memory of one CTA: [align slack | ring | chunk staging: 4 KB per epilogue warp | barriers, TMEM slot,
// shared thresholds]
__host__ __device__ constexpr size_t smem_bytes_for(int mode) {
    return 1024 + static_cast<size_t>(ring_bytes_for(mode)) + 4 * epi_sets_for_mode(mode) * 32 * 32 * sizeof(float) + 2048;
}
static_assert(smem_bytes_for(RADAR_MODE_DPR) <= 227 * 1024 && smem_bytes_for(RADAR_MODE_KL) <= 227 * 1024 &&
                  smem_bytes_for(RADAR_MODE_HYBRID) <= 227 * 1024,
              "shared memory budget");

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) {
            printf("radar tc_filter: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
// one lane of a converged warp (the operands of the predicated tcgen05 / TMA instruction then stay warp-uniform,
// so the compiler keeps them in uniform registers instead of emitting an R2UR waterfall loop per instruction)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// shared-memory matrix descriptor, K-major operand, swizzled canonical layout
//   SW128: rows of 128 B, 8-row atoms 1024 B apart (layout type 2);  SW64: rows of 64 B, atoms 512 B apart (type 4)
// The start-address field is (addr >> 4) in bits [0,14); shared-memory addresses stay below 256 KB, so adding
// (byte offset >> 4) to a descriptor never carries out of the field.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // start address,   bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                            // LBO (ignored for swizzled K-major), bits [16,30)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;   // SBO,             bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                            // descriptor version 1 (sm_100)
    d |= static_cast<uint64_t>(layout_type & 7u) << 61;             // swizzle mode,    bits [61,64)
    return d;
}

__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
    // kind::f16: D=F32 (bits 4-5 = 1), A=BF16 (bits 7-9 = 1), B=BF16 (bits 10-12 = 1), K-major A and B,
    // N>>3 at bits 17-22, M>>4 at bits 24-28
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- query pack: fp32 queries -> bf16 A rows + per-query shift / error bound -----------------------------
// apack row layout (bf16): [ d embedding values (scaled by alpha in hybrid) | v_hi (16) | v_lo (16) ]
// qmeta[row] = {shift, qerr}.  One warp per packed row; rows >= q are zero-filled.
struct PackArgs {
    const float* q_emb;
    const float* p16;
    const float* entropy;
    int64_t q, q_pad;
    int d, mode;
    float alpha, oma;
    float emb_max_norm, logq_max_abs;
    float logq_col_max[kObsPad];  // per-observation max |log q| over the corpus (all zero: use logq_max_abs)
    uint16_t* apack;
    float* qshift;  // [q_pad]
    float* qerr;    // [q_pad]
};

__global__ void __launch_bounds__(256) query_pack_kernel(const PackArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= a.q_pad) return;
    const bool has_ip = a.mode != RADAR_MODE_KL, has_kl = a.mode != RADAR_MODE_DPR;
    const int cols = a_cols_for(a.mode, a.d);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.apack) + row * cols;
    const bool valid = row < a.q;
    float ss = 0.0f, sv = 0.0f;
    if (has_ip) {
        const float scale = a.mode == RADAR_MODE_HYBRID ? a.alpha : 1.0f;
        for (int c = lane; c < a.d; c += 32) {
            const float v = valid ? __fmul_rn(scale, a.q_emb[row * a.d + c]) : 0.0f;
            ss = fmaf(v, v, ss);
            dst[c] = __float2bfloat16_rn(v);
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (has_kl) {
        const float scale = a.mode == RADAR_MODE_HYBRID ? a.oma : 1.0f;
        if (lane < kObsPad) {
            const float v = valid ? __fmul_rn(scale, a.p16[row * kObsPad + lane]) : 0.0f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
            dst[(has_ip ? a.d : 0) + lane] = hi;
            dst[(has_ip ? a.d : 0) + kObsPad + lane] = lo;
            sv = fabsf(v) * a.logq_col_max[lane];  // sum_j |v_j| max_n |L_nj| >= sum_j |v_j L_nj| for every case n
        }
        for (int o = 16; o > 0; o >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (lane == 0) {
        float shift = 0.0f;
        if (valid && has_kl) shift = a.mode == RADAR_MODE_HYBRID ? __fmul_rn(a.oma, a.entropy[row]) : a.entropy[row];
        // |canonical key - filter key| <= qerr  (derivation in DESIGN.md, "certificate"):
        //   bf16 RN of both embedding operands: (2^-8 + 2^-18) |a||c| ; tensor-core fp32 accumulation of <= 560
        //   exact products and the canonical fp32 chain: <= 2^-13 (|a||c| + sum|v||L|) ; hi/lo split of v and L
        //   with the lo*lo product dropped: (2^-17 + 2^-18) sum|v||L| ; canonical combine / shift roundings:
        //   2^-20 of the magnitudes involved.
        const float ip_mag = sqrtf(ss) * a.emb_max_norm;
        const float kl_mag = sv;
        const float e = 0.00403f * ip_mag + 1.65e-4f * kl_mag + 1e-6f * (fabsf(shift) + ip_mag + kl_mag) + 1e-30f;
        a.qshift[row] = shift;
        a.qerr[row] = e;
    }
}

// ---- the filter kernel ------------------------------------------------------------------------------
struct FilterArgs {
    const uint16_t* apack;
    const float* qshift;
    int64_t q;          // real queries
    int64_t q_tiles;    // work tiles of 256 query rows
    int64_t n;          // corpus rows
    int d;
    int parts;
    int64_t rows_per_part;  // multiple of BLOCK_N
    int kp;
    uint64_t* cand;     // [q_pad][parts][epilogue sets][kCandCap]
    uint32_t* cnt;      // [q_pad][parts][epilogue sets]
    float* thr;         // [q_pad][parts][epilogue sets]  final thresholds
    uint32_t* gthr;     // [q_pad] best published threshold per query (ord-encoded, 0 = none): slabs of the same
                        // query tile that run later start from it instead of -inf
    unsigned long long* progress;  // [units] (round << 32 | tiles loaded) of every pair; zeroed before the launch
    int window;         // tiles a pair may run ahead of the slowest pair sweeping the same slab (0 = unbounded)
    float* dbg_scores;  // RADAR_DEBUG builds only: optional [q_pad][n] dense dump of the filter keys
    // KL threshold prepass (see launch_filter): prepass = 1 -> the epilogue only records, per query, the maximum key of every
    // group of `group_tiles` consecutive sampled tiles of a slab (groupmax[qrow][slab][group], ord-encoded); the k'-th largest
    // group maximum is then a near-exact initial threshold for the real pass (gthr_init = 1), which so sees ~k' survivors
    // per query instead of the ~k' ln(n) of a cold start
    uint32_t* groupmax;
    int prepass, gthr_init, group_tiles, groups, groups_per_slab;
    int tile_stride;    // 1, or > 1 in the prepass: only every tile_stride-th tile of a slab is visited (a sample still
                        // yields a valid, slightly looser threshold at 1/tile_stride of the cost)
    int dbg_flags;      // RADAR_DEBUG builds only (env RADAR_TC_DBG, results are garbage): 1 = epilogue only recycles the
                        // accumulators, 2 = no TMA loads (MMAs run on whatever is in shared memory), 4 = epilogue
                        // loads the accumulators but does not look at them
    unsigned long long* clk;  // [2] SM cycles / nanoseconds CTA 0 spent in the kernel (average SM clock of the launch)
};

// cluster helpers (CTA pairs)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> even CTA
// arrive on the barrier at the same offset in the pair's leader (even) CTA; for the leader itself this is local
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's barrier (both CTAs of a pair call it)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_addr, uint32_t dst_addr, int c0,
                                                 int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(dst_addr),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kTmemCols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T ; kind::f16 with bf16 inputs, fp32 accumulate.  M = 256 across the CTA pair:
// each CTA supplies its own 128 A rows (TMEM) and half of the B rows (smem).
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit -> arrive (once every MMA issued so far by this thread has completed) on the barrier at this offset in
// BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

__device__ __forceinline__ unsigned long long ld_progress(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_progress(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// KB_T > 0: the embedding K-block count (d / 64) is a compile-time constant and the issue / load loops unroll
// completely (KB_T = 8 is BiomedCLIP's d = 512); KB_T = 0: d is read from the arguments.
#ifdef RADAR_TC_TIMING
__device__ unsigned long long g_tc_timing[296 * 2];  // experiment aid (tools/tc_timing.py): per CTA [start, end] of the last real pass
#endif
template <int MODE, int KB_T>
__global__ void __launch_bounds__(threads_for_mode(MODE), 1)
tc_filter_kernel(const __grid_constant__ CUtensorMap map_emb, const __grid_constant__ CUtensorMap map_kl,
                 const FilterArgs a) {
    constexpr bool HAS_IP = MODE != RADAR_MODE_KL;
    constexpr bool HAS_KL = MODE != RADAR_MODE_DPR;
    constexpr int BLOCK_N = block_n_for_mode(MODE);     // corpus rows per MMA tile (whole pair)
    constexpr int LOAD_N = BLOCK_N / 2;                 // corpus rows this CTA stages per tile
    constexpr int ACC_STAGES = acc_stages_for_mode(MODE);
    constexpr int EPI_SETS = epi_sets_for_mode(MODE);
    constexpr int COLS = BLOCK_N / EPI_SETS;            // accumulator columns one epilogue warp looks at
    constexpr int ACC_COL0 = acc_col0_for_mode(MODE);
    constexpr int A_KL_COL = a_kl_col_for_mode(MODE);
    constexpr uint32_t IDESC = make_idesc_mn(kTileQ, BLOCK_N);
    constexpr int SUB_IP = sub_ip_bytes(MODE), SUB_KL = sub_kl_bytes(MODE);
    static_assert(ACC_COL0 + ACC_STAGES * BLOCK_N <= kTmemCols, "TMEM budget");
    static_assert(LOAD_N % 8 == 0 && BLOCK_N % 16 == 0, "tile shape");
    static_assert(HAS_IP || KB_T == 0, "KL mode has no embedding K-blocks");

    const int kblocks = HAS_IP ? (KB_T > 0 ? KB_T : a.d / 64) : 0;
    const int groups = groups_for(kblocks);
    const int slot_stride = slot_stride_for(MODE, kblocks);
    const int slots = slots_for(MODE, kblocks);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* stage = reinterpret_cast<float*>(smem + ring_bytes_for(MODE));     // [epilogue warps][32 columns][32 lanes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 4 * EPI_SETS * 32 * 32);
    uint64_t* full_bar = bars;                                   // [kMaxSlots][kMaxGroups] (only the leader's are waited on)
    uint64_t* empty_bar = full_bar + kMaxSlots * kMaxGroups;      // [kMaxSlots]
    uint64_t* tfull_bar = empty_bar + kMaxSlots;                  // [ACC_STAGES]
    uint64_t* tempty_bar = tfull_bar + 8;                         // [ACC_STAGES]  (leader's, 8 arrivals per consuming set)
    uint64_t* aready_bar = tempty_bar + 8;                        // [1]           (leader's, 8 arrivals)
    uint64_t* adone_bar = aready_bar + 1;                         // [1] KL: every MMA of an item has completed (both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(adone_bar + 1);
    // KL: best threshold of each query row of this CTA over the warp sets, tagged with the item it belongs to:
    // (item number + 1) << 32 | ord-encoded threshold.  Sets may be one item apart; a tagged word of another item is
    // ignored by readers and can never overwrite a newer one (atomicMax on the whole word).
    unsigned long long* thr_sh = reinterpret_cast<unsigned long long*>(tmem_slot + 2);  // [kBlockM]
    const uint32_t ring_addr = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int64_t unit = blockIdx.x >> 1;        // work-tile processor id (CTA pair)
    const int64_t units = gridDim.x >> 1;
    const int64_t items = a.q_tiles * a.parts;
    const int64_t tile_step = static_cast<int64_t>(BLOCK_N) * a.tile_stride;  // the prepass samples every tile_stride-th tile

#ifdef RADAR_TC_TIMING
    if (!a.prepass && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_tc_timing[blockIdx.x * 2]));
#endif
    unsigned long long clk0 = 0, ns0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        clk0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    }
    if (warp == 0 && lane == 0) {
        if (HAS_IP) prefetch_tmap(&map_emb);
        if (HAS_KL) prefetch_tmap(&map_kl);
        for (int i = 0; i < kMaxSlots * kMaxGroups; ++i) mbar_init(&full_bar[i], 1);
        for (int i = 0; i < kMaxSlots; ++i) mbar_init(&empty_bar[i], 1);
        for (int i = 0; i < ACC_STAGES; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], MODE == RADAR_MODE_KL ? 8 * kKlColSplit : 8 * EPI_SETS);  // warps reading one stage
        }
        mbar_init(aready_bar, 8);
        mbar_init(adone_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot);
    if (MODE == RADAR_MODE_KL && threadIdx.x >= 64 && threadIdx.x < 64 + kBlockM) thr_sh[threadIdx.x - 64] = 0ull;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything is signalled remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) {  // all 512 columns are allocated, so the base is (lane 0, column 0); the MMA issuer relies on it
        if (threadIdx.x == 0) printf("radar tc_filter: unexpected TMEM base %u\n", tmem_base);
        __trap();
    }

    if (warp == 0) {
        // ================================ TMA producer (every CTA: its half of each corpus tile) =================
        // the whole warp runs the loop converged; one elected lane arms the barriers and issues the copies
        uint32_t slot = 0, sph = 0, round = 0;
        const bool publish = a.window > 0 && leader;  // progress is published even after this pair stopped waiting
        bool honour_window = publish;
        for (int64_t item = unit; item < items; item += units, ++round) {
            const int part = static_cast<int>(item / a.q_tiles);
            const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
            const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
            // pairs sweeping the same slab in this round: items [part*q_tiles, (part+1)*q_tiles) of round `round`
            const int64_t round_base = static_cast<int64_t>(round) * units;
            const int peer_lo = static_cast<int>(max(static_cast<int64_t>(0), part * a.q_tiles - round_base));
            const int peer_hi = static_cast<int>(min(min(units, items - round_base), (part + 1) * a.q_tiles - round_base));
            uint32_t j = 0;
            for (int64_t row0 = row_begin; row0 < row_end; row0 += tile_step, ++j) {
                if (publish && (j % kProgressEvery) == 0) {
                    const unsigned long long mine = (static_cast<unsigned long long>(round) << 32) | j;
                    if (lane == 0) st_progress(a.progress + unit, mine);
                    if (honour_window && j > static_cast<uint32_t>(a.window)) {
                        const unsigned long long need = mine - static_cast<unsigned long long>(a.window);
                        uint32_t polls = 0;
                        while (true) {
                            unsigned long long m = ~0ull;
                            for (int u = peer_lo + lane; u < peer_hi; u += 32) m = min(m, ld_progress(a.progress + u));
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                            if (m >= need) break;
                            if (++polls > kWindowSpinLimit) {  // never a correctness matter: stop waiting for good
                                honour_window = false;
                                break;
                            }
                            __nanosleep(500);
                        }
                    }
                }
                mbar_wait(&empty_bar[slot], sph ^ 1);
                if (RADAR_DBG(a.dbg_flags & 2)) {
                    if (leader && lane == 0)
                        for (int g = 0; g < groups; ++g) mbar_arrive(&full_bar[slot * kMaxGroups + g]);
                } else if (elect_one()) {
                    const uint32_t dst = ring_addr + slot * slot_stride;
                    const uint32_t fb = smem_u32(&full_bar[slot * kMaxGroups]);
                    const int my_row = static_cast<int>(row0) + static_cast<int>(cta_rank) * LOAD_N;
#pragma unroll
                    for (int g = 0; g < (KB_T > 0 ? groups_for(KB_T) : kMaxGroups); ++g) {
                        if (g < groups) {
                            const int kb0 = 2 * g, kb1 = min(kblocks, 2 * g + 2);
                            const bool last = g == groups - 1;
                            if (leader) {  // both halves land on the leader's barrier
                                const uint32_t bytes = (kb1 - kb0) * (LOAD_N * 128) + ((last && HAS_KL) ? LOAD_N * 64 : 0);
                                mbar_expect_tx(&full_bar[slot * kMaxGroups + g], bytes * 2);
                            }
                            if (HAS_IP) {
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk) {
                                    const int kb = kb0 + kk;
                                    if (kb < kb1) tma_load_2d_pair(&map_emb, fb + g * 8, dst + kb * SUB_IP, kb * 64, my_row);
                                }
                            }
                            if (last && HAS_KL) tma_load_2d_pair(&map_kl, fb + g * 8, dst + kblocks * SUB_IP, 0, my_row);
                        }
                    }
                }
                __syncwarp();
                if (++slot == static_cast<uint32_t>(slots)) {
                    slot = 0;
                    sph ^= 1;
                }
            }
        }
        if (a.window > 0 && leader && lane == 0) st_progress(a.progress + unit, ~0ull);  // nobody waits for a finished pair
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA only) ================================
        // converged warp; the tcgen05.mma / tcgen05.commit instructions are issued by one elected lane
        if (leader) {
            uint32_t slot = 0, sph = 0, as = 0, aph = 0, item_no = 0;
            for (int64_t item = unit; item < items; item += units, ++item_no) {
                const int part = static_cast<int>(item / a.q_tiles);
                const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
                const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
                mbar_wait(aready_bar, item_no & 1);  // this item's query tile is in TMEM (both CTAs)
                tc_fence_after();
                for (int64_t row0 = row_begin; row0 < row_end; row0 += tile_step) {
                    mbar_wait(&tempty_bar[as], aph ^ 1);  // epilogues drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = ACC_COL0 + as * BLOCK_N;  // TMEM base is 0: the pair owns all 512 columns
                    const uint32_t sbase = ring_addr + slot * slot_stride;
                    const uint64_t desc_ip = make_smem_desc(sbase, 1024, 2);
                    const uint64_t desc_kl = make_smem_desc(sbase + kblocks * SUB_IP, 512, 4);
                    uint64_t* fb = &full_bar[slot * kMaxGroups];
#pragma unroll
                    for (int g = 0; g < (KB_T > 0 ? groups_for(KB_T) : kMaxGroups); ++g) {
                        if (g < groups) {
                            const bool last = g == groups - 1;
                            mbar_wait(&fb[g], sph);
                            tc_fence_after();
                            if (elect_one()) {
                                if (HAS_IP) {
#pragma unroll
                                    for (int kk = 0; kk < 2; ++kk) {
                                        const int kb = 2 * g + kk;
                                        if (kb < kblocks) {
#pragma unroll
                                            for (int ks = 0; ks < 4; ++ks)  // 4 x (K = 16) per 64-wide block; +32 B per step
                                                umma_ts_pair(d_tmem, kb * 32 + ks * 8,
                                                             desc_ip + static_cast<uint64_t>((kb * SUB_IP + ks * 32) >> 4), IDESC,
                                                             (kb | ks) ? 1u : 0u);
                                        }
                                    }
                                }
                                if (last) {
                                    if (HAS_KL) {
                                        const uint32_t a_hi = A_KL_COL, a_lo = a_hi + 8;
                                        umma_ts_pair(d_tmem, a_hi, desc_kl, IDESC, HAS_IP ? 1u : 0u);  // v_hi . L_hi
                                        umma_ts_pair(d_tmem, a_hi, desc_kl + 2, IDESC, 1u);             // v_hi . L_lo
                                        umma_ts_pair(d_tmem, a_lo, desc_kl, IDESC, 1u);                 // v_lo . L_hi
                                    }
                                    umma_commit_pair(&empty_bar[slot]);  // slot reusable in both CTAs
                                    umma_commit_pair(&tfull_bar[as]);    // accumulator complete (both CTAs)
                                }
                            }
                            __syncwarp();
                        }
                    }
                    if (++slot == static_cast<uint32_t>(slots)) {
                        slot = 0;
                        sph ^= 1;
                    }
                    if (++as == ACC_STAGES) {
                        as = 0;
                        aph ^= 1;
                    }
                }
                if (MODE == RADAR_MODE_KL) {  // sets other than set 0 may still be consuming this item's last tiles
                    if (elect_one()) umma_commit_pair(adone_bar);
                    __syncwarp();
                }
            }
        }
    } else if constexpr (MODE == RADAR_MODE_KL) {
        // ================================ KL epilogue: tile streams x column splits ================================
        // Tile g (running number over all items of this pair; accumulator stage g mod ACC_STAGES) is filtered by the sets
        // of stream g mod TS, set (ts, cs) reading columns [cs * COLS, (cs + 1) * COLS) -- 64 columns per
        // tcgen05.wait::ld.  A stage is handed back when the last of its readers has its columns in registers.
        constexpr int TS = kKlTileStreams, CS = kKlColSplit;
        constexpr int COLS = BLOCK_N / CS;
        static_assert(COLS % 32 == 0 || COLS % 32 == 16, "KL columns per warp must be 32k or 32k + 16");
        const int quad = warp & 3;
        const int set = (warp - 2) >> 2;
        const int ts = set / CS, cs = set % CS;
        const int col0 = cs * COLS;
        const int r_in_tile = static_cast<int>(cta_rank) * kBlockM + quad * 32 + lane;
        const int r_in_cta = quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        float* my_stage = stage + (warp - 2) * 32 * 32 + lane;
        const int a_cols = a_cols_for(MODE, a.d);
        uint32_t g = 0;         // running tile number of this pair (the issuer's accumulator stage is g mod ACC_STAGES)
        uint32_t item_no = 0;
        for (int64_t item = unit; item < items; item += units, ++item_no) {
            const int64_t qtile = item % a.q_tiles;
            const int part = static_cast<int>(item / a.q_tiles);
            const int64_t qrow = qtile * kTileQ + r_in_tile;
            const bool valid = qrow < a.q;
            const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
            const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
            if (set == 0) {
                // ---- query tile -> TMEM, once every MMA of the previous item has completed ----
                if (item_no > 0) {
                    mbar_wait(adone_bar, (item_no - 1) & 1);
                    tc_fence_after();
                }
                const uint4* ksrc = reinterpret_cast<const uint4*>(a.apack + qrow * a_cols);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint4 u0 = __ldg(ksrc + 2 * c), u1 = __ldg(ksrc + 2 * c + 1);
                    const uint32_t v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    tmem_st_x8(tmem_base + lane_addr + A_KL_COL + c * 8, v);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(aready_bar);
            }
            const float shift = a.qshift[qrow];
            const uint32_t g0 = (a.parts > 1 || a.gthr_init) ? *reinterpret_cast<volatile const uint32_t*>(a.gthr + qrow) : 0u;
            uint32_t thr_ord = g0;                                          // ord-encoded canonical-key units, 0 = none
            float thr = g0 ? ord2f(g0) : -CUDART_INF_F;
            // inherited thresholds are stepped a few ulps down (see the DPR / hybrid epilogue below for why)
            auto cmp_of = [&](float t) {
                const float c = __fadd_rn(t, shift);
                return c - 4e-7f * (fabsf(c) + fabsf(shift));
            };
            float thr_cmp = valid ? (g0 ? cmp_of(thr) : -CUDART_INF_F) : CUDART_INF_F;
            int cnt = 0;
            const int64_t slot_idx = (qrow * a.parts + part) * EPI_SETS + set;
            uint64_t* buf = a.cand + slot_idx * kCandCap;
            const uint32_t ntiles = static_cast<uint32_t>((row_end - row_begin + tile_step - 1) / tile_step);
            // first tile of this item that belongs to this set's stream
            uint32_t j = (static_cast<uint32_t>(ts) + TS - (g % TS)) % TS;
            for (; j < ntiles; j += TS) {
                const uint32_t gt = g + j;
                const uint32_t stg = gt % ACC_STAGES, aph = (gt / ACC_STAGES) & 1u;
                const int64_t row0 = row_begin + static_cast<int64_t>(j) * tile_step + col0;  // first corpus row of this set's columns
                if (!a.prepass && EPI_SETS > 1) {
                    // a compaction of another set may have published a better threshold for this row since
                    const unsigned long long tw = *reinterpret_cast<volatile unsigned long long*>(thr_sh + r_in_cta);
                    const uint32_t ts = static_cast<uint32_t>(tw);
                    if (valid && static_cast<uint32_t>(tw >> 32) == item_no + 1u && ts > thr_ord) {
                        thr_ord = ts;
                        thr = ord2f(ts);
                        thr_cmp = cmp_of(thr);
                    }
                }
                mbar_wait(&tfull_bar[stg], aph);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + lane_addr + ACC_COL0 + stg * BLOCK_N + col0;
                float tmax = -CUDART_INF_F;  // prepass: maximum over the whole tile
                // one 32-column group of accumulators of this thread's query row, in registers
                auto filter32 = [&](const float (&cur)[32], const int width, const int64_t rowc) {
                    if (RADAR_DBG(a.dbg_flags & 5)) {
                        if (RADAR_DBG(a.dbg_flags & 4)) asm volatile("" ::"f"(cur[0]), "f"(cur[width - 1]));
                        return;
                    }
                    float mm[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) mm[u] = cur[u];
#pragma unroll
                    for (int jj = 8; jj < 32; ++jj)
                        if (jj < width) mm[jj & 7] = fmaxf(mm[jj & 7], cur[jj]);
                    float m = fmaxf(fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])), fmaxf(fmaxf(mm[4], mm[5]), fmaxf(mm[6], mm[7])));
                    if (a.prepass) {
                        if (rowc + width > row_end) {  // ragged last tile: rows past the end were zero-filled by TMA
                            m = -CUDART_INF_F;
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj)
                                if (jj < width && rowc + jj < row_end) m = fmaxf(m, cur[jj]);
                        }
                        tmax = fmaxf(tmax, m);
                        return;
                    }
                    if (__any_sync(0xffffffffu, m >= thr_cmp) || RADAR_DBG(a.dbg_scores != nullptr)) {
                        // rare path.  The eight partial maxima (chain u = columns u, u+8, u+16, u+24) say where the
                        // survivors are: only chains in which SOME lane has one are staged and compared (a warp-uniform
                        // branch per chain), typically one or two of the eight.
                        uint32_t mask = 0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (__any_sync(0xffffffffu, mm[u] >= thr_cmp) || RADAR_DBG(a.dbg_scores != nullptr)) {
#pragma unroll
                                for (int jj = u; jj < 32; jj += 8) {
                                    if (jj < width) {
                                        my_stage[jj * 32] = cur[jj];
                                        mask |= (cur[jj] >= thr_cmp ? 1u : 0u) << jj;
                                    }
                                }
                            }
                        }
                        if (RADAR_DBG(a.dbg_scores != nullptr) && valid) {
                            for (int jj = 0; jj < width; ++jj)
                                if (rowc + jj < a.n) a.dbg_scores[qrow * a.n + rowc + jj] = my_stage[jj * 32] - shift;
                        }
                        while (mask) {
                            const int jj = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int64_t row = rowc + jj;
                            if (row < row_end)
                                buf[cnt++] = make_composite(__fsub_rn(my_stage[jj * 32], shift), static_cast<uint32_t>(row));
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
                                thr_ord = f2ord(t);
                                thr_cmp = __fadd_rn(t, shift);  // own threshold: exact (every kept key was fl(acc - shift))
                                cnt = a.kp;
                                if (EPI_SETS > 1)
                                    atomicMax(thr_sh + r_in_cta,
                                              (static_cast<unsigned long long>(item_no + 1u) << 32) | thr_ord);
                            }
                        }
                    }
                };
                // 64 columns per tcgen05.wait::ld: a wait after every single load caps TMEM reads at ~128 B/cycle/SM,
                // two loads per wait and three or four warps per lane quadrant reach ~470 (tools/micro/ldtm_bw.cu)
#pragma unroll
                for (int c0 = 0; c0 < COLS; c0 += 64) {
                    const int w0 = COLS - c0 >= 32 ? 32 : COLS - c0;
                    const int w1 = COLS - c0 - 32 >= 32 ? 32 : (COLS - c0 - 32 > 0 ? COLS - c0 - 32 : 0);
                    float v0[32], v1[32];
                    if (w0 == 32) tmem_ld_x32(t_acc + c0, v0);
                    else tmem_ld_x16(t_acc + c0, v0);
                    if (w1 == 32) tmem_ld_x32(t_acc + c0 + 32, v1);
                    else if (w1 == 16) tmem_ld_x16(t_acc + c0 + 32, v1);
                    tmem_wait_ld();
                    if (c0 + 64 >= COLS) {  // this warp's columns are in registers (or already filtered): done with the stage
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&tempty_bar[stg]);
                    }
                    filter32(v0, w0, row0 + c0);
                    if (w1 > 0) filter32(v1, w1, row0 + c0 + 32);
                }
                if (a.prepass && !RADAR_DBG(a.dbg_flags & 5)) {
                    // every tile is its own contribution to its group's maximum (the sets see different tiles of a group)
                    if (valid && tmax > -CUDART_INF_F)
                        atomicMax(a.groupmax + qrow * a.groups + part * a.groups_per_slab + static_cast<int>(j) / a.group_tiles,
                                  f2ord(__fsub_rn(tmax, shift)));
                }
            }
            g += ntiles;
            if (a.prepass) continue;
            a.cnt[slot_idx] = valid ? static_cast<uint32_t>(cnt) : 0u;
            a.thr[slot_idx] = thr;
            if (a.parts > 1 && valid && thr > -CUDART_INF_F) atomicMax(a.gthr + qrow, f2ord(thr));
        }
    } else {
        // ================================ epilogue warps (2..5, and 6..9 with two sets) ================================
        const int quad = warp & 3;  // TMEM lane quadrant this warp may access
        const int set = (warp - 2) >> 2;   // which COLS-wide column range of every tile this warp filters
        const int col0 = set * COLS;
        const int r_in_tile = static_cast<int>(cta_rank) * kBlockM + quad * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        float* my_stage = stage + (warp - 2) * 32 * 32 + lane;  // [column * 32]: bank == lane, no conflicts
        const int a_cols = a_cols_for(MODE, a.d);  // bf16 per packed row
        uint32_t as = 0, aph = 0;
        for (int64_t item = unit; item < items; item += units) {
            const int64_t qtile = item % a.q_tiles;
            const int part = static_cast<int>(item / a.q_tiles);
            const int64_t qrow = qtile * kTileQ + r_in_tile;
            const bool valid = qrow < a.q;
            const int64_t row_begin = static_cast<int64_t>(part) * a.rows_per_part;
            const int64_t row_end = min(a.n, row_begin + a.rows_per_part);
            // ---- A tile -> TMEM (every MMA of the previous item has completed: its last accumulator was consumed) ----
            if (set == 0) {
                const uint4* src = reinterpret_cast<const uint4*>(a.apack + qrow * a_cols);
                const int ip_chunks = HAS_IP ? a.d / 16 : 0;  // 16 bf16 = 8 TMEM columns per chunk
                for (int c = 0; c < ip_chunks; ++c) {
                    const uint4 u0 = __ldg(src + 2 * c), u1 = __ldg(src + 2 * c + 1);
                    const uint32_t v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    tmem_st_x8(tmem_base + lane_addr + c * 8, v);
                }
                if (HAS_KL) {
                    const uint4* ksrc = src + 2 * ip_chunks;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint4 u0 = __ldg(ksrc + 2 * c), u1 = __ldg(ksrc + 2 * c + 1);
                        const uint32_t v[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                        tmem_st_x8(tmem_base + lane_addr + A_KL_COL + c * 8, v);
                    }
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(aready_bar);
            }
            const float shift = a.qshift[qrow];
            // start from the best threshold an earlier slab of this query published (a lower bound on the k'-th best
            // key over the whole corpus, so dropping below it is always safe)
            const uint32_t g0 = (a.parts > 1 || a.gthr_init) ? *reinterpret_cast<volatile const uint32_t*>(a.gthr + qrow) : 0u;
            float thr = g0 ? ord2f(g0) : -CUDART_INF_F;                    // in canonical-key units
            float thr_cmp = valid ? (g0 ? __fadd_rn(thr, shift) : -CUDART_INF_F) : CUDART_INF_F;  // accumulator units
            // an inherited threshold is the key of some case computed as fl(acc - shift); fl(thr + shift) may round ABOVE that
            // case's accumulator, and cases tied with it (duplicates, possibly with smaller ids) would then be dropped:
            // step a few ulps down -- a lower threshold is always safe
            if (valid && g0) thr_cmp -= 4e-7f * (fabsf(thr_cmp) + fabsf(shift));
            int cnt = 0;
            const int64_t slot_idx = (qrow * a.parts + part) * EPI_SETS + set;  // this warp set's private buffer
            uint64_t* buf = a.cand + slot_idx * kCandCap;
            float gmax = -CUDART_INF_F;  // prepass: running maximum of the current tile group

            for (int64_t row0 = row_begin; row0 < row_end; row0 += tile_step) {
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + lane_addr + ACC_COL0 + as * BLOCK_N + col0;
                const int64_t rowc = row0 + col0;  // corpus row of this warp's first column
                // The whole accumulator row goes to registers with back-to-back tcgen05.ld, then the TMEM stage is
                // handed back to the MMA issuer BEFORE the values are looked at: the threshold filter below (and its
                // occasional candidate insertions / compactions) overlaps the next tile's MMAs instead of sitting
                // between two of them.
                float v[COLS];
                if (!RADAR_DBG(a.dbg_flags & 1)) {
#pragma unroll
                    for (int c = 0; c < COLS; c += 32) {
                        if (COLS - c >= 32) tmem_ld_x32(t_acc + c, v + c);
                        else tmem_ld_x16(t_acc + c, v + c);
                    }
                    tmem_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
                if (++as == ACC_STAGES) {
                    as = 0;
                    aph ^= 1;
                }
                if (RADAR_DBG(a.dbg_flags & 5)) {
                    if (RADAR_DBG(a.dbg_flags & 4)) asm volatile("" ::"f"(v[0]), "f"(v[COLS - 1]));
                    continue;
                }
                if (a.prepass) {
                    float m;
                    if (row0 + BLOCK_N <= row_end) {
                        float mm[8];  // eight independent chains: the epilogue warp is alone on its scheduler, ILP is all it has
#pragma unroll
                        for (int u = 0; u < 8; ++u) mm[u] = v[u];
#pragma unroll
                        for (int jj = 8; jj < COLS; ++jj) mm[jj & 7] = fmaxf(mm[jj & 7], v[jj]);
                        m = fmaxf(fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3])), fmaxf(fmaxf(mm[4], mm[5]), fmaxf(mm[6], mm[7])));
                    } else {  // ragged last tile: rows past the end were zero-filled by TMA and must not count
                        m = -CUDART_INF_F;
#pragma unroll
                        for (int jj = 0; jj < COLS; ++jj)
                            if (rowc + jj < row_end) m = fmaxf(m, v[jj]);
                    }
                    gmax = fmaxf(gmax, m);
                    const int64_t tile_no = (row0 - row_begin) / tile_step;  // index among this slab's sampled tiles
                    const bool group_ends = ((tile_no + 1) % a.group_tiles) == 0 || row0 + tile_step >= row_end;
                    if (group_ends) {
                        if (valid)  // both warp sets feed the same group slot
                            atomicMax(a.groupmax + qrow * a.groups + part * a.groups_per_slab + tile_no / a.group_tiles,
                                      f2ord(__fsub_rn(gmax, shift)));
                        gmax = -CUDART_INF_F;
                    }
                    continue;
                }
#pragma unroll
                for (int c = 0; c < COLS; c += 32) {
                    const int width = (COLS - c) < 32 ? (COLS - c) : 32;  // 32 or 16 (compile time after unroll)
                    float mm[4] = {v[c], v[c + 1], v[c + 2], v[c + 3]};  // four independent chains (ILP)
#pragma unroll
                    for (int jj = 4; jj < 32; ++jj)
                        if (jj < width) mm[jj & 3] = fmaxf(mm[jj & 3], v[c + jj]);
                    const float m = fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3]));
                    if (__any_sync(0xffffffffu, m >= thr_cmp) || RADAR_DBG(a.dbg_scores != nullptr)) {
                        // rare path, written for a small instruction footprint (it used to be unrolled per column and
                        // pushed the kernel far beyond the instruction cache): the chunk is staged in this warp's
                        // shared-memory scratch together with a survivor bit mask, then a short loop walks the set bits
                        uint32_t mask = 0;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            if (jj < width) {
                                my_stage[jj * 32] = v[c + jj];
                                mask |= (v[c + jj] >= thr_cmp ? 1u : 0u) << jj;
                            }
                        }
                        if (RADAR_DBG(a.dbg_scores != nullptr) && valid) {
                            for (int jj = 0; jj < width; ++jj)
                                if (rowc + c + jj < a.n) a.dbg_scores[qrow * a.n + rowc + c + jj] = my_stage[jj * 32] - shift;
                        }
                        const int64_t row_base = rowc + c;
                        while (mask) {
                            const int jj = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int64_t row = row_base + jj;
                            if (row < row_end)
                                buf[cnt++] = make_composite(__fsub_rn(my_stage[jj * 32], shift), static_cast<uint32_t>(row));
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
                                thr_cmp = __fadd_rn(t, shift);
                                cnt = a.kp;
                            }
                        }
                    }
                }
            }
            if (a.prepass) continue;
            a.cnt[slot_idx] = valid ? static_cast<uint32_t>(cnt) : 0u;
            a.thr[slot_idx] = thr;
            if (a.parts > 1 && valid && thr > -CUDART_INF_F) atomicMax(a.gthr + qrow, f2ord(thr));
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef RADAR_TC_TIMING
    if (!a.prepass && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_tc_timing[blockIdx.x * 2 + 1]));
#endif
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still touch it
    if (warp == 1) tmem_dealloc_pair(tmem_base);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
        a.clk[0] = clock64() - clk0;
        a.clk[1] = ns1 - ns0;
    }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        RADAR_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return RADAR_E_CUDA;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    *out = fn;
    return RADAR_OK;
}

static int encode_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint32_t box_inner,
                          uint32_t box_rows, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn;
    int rc = get_encode_fn(&fn);
    if (rc) return rc;
    cuuint64_t dims[2] = {inner, rows};
    cuuint64_t strides[1] = {inner * sizeof(uint16_t)};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu rows=%llu box=%ux%u)", (int)r,
                  (unsigned long long)inner, (unsigned long long)rows, box_inner, box_rows);
        return RADAR_E_CUDA;
    }
    return RADAR_OK;
}

struct FilterLaunch {
    const radar_corpus_t* corpus;
    const radar_queries_t* queries;
    int mode;
    float alpha, oma;
    int64_t q, q_tiles;   // q_tiles = work tiles of 256 rows
    int parts;
    int64_t rows_per_part;
    int kp;
    uint64_t* cand;
    uint32_t* cnt;
    float* thr;
    uint32_t* gthr;   // [q_pad] followed (8-byte aligned) by kMaxUnits progress words; the region is zeroed here
    float* qerr;      // [q_pad]
    uint16_t* apack;  // [q_pad][a_cols] followed by qshift [q_pad] floats
    int units;        // CTA pairs to launch (<= kMaxUnits)
    int device_sms;   // SMs of the device (the window is only honoured when every CTA is resident)
    float* dbg_scores;
    uint32_t* groupmax;  // [q_pad][groups] when the KL threshold prepass is planned (groups > 0), else unused
    int groups, group_tiles, tile_stride, groups_per_slab;
    cudaEvent_t ev_start, ev_stop;  // optional: recorded around the filter kernel only
    unsigned long long* clk_dev;    // out: device address of the {cycles, ns} pair the kernel writes
};

static inline size_t gthr_region_bytes(int64_t q_pad) {
    return (sizeof(uint32_t) * static_cast<size_t>(q_pad) + 7) / 8 * 8 + sizeof(unsigned long long) * (kMaxUnits + 2);
}

template <int MODE, int KB_T>
static int launch_filter_mode(const FilterLaunch& fl, const FilterArgs& fa, cudaStream_t st) {
    constexpr int BLOCK_N = block_n_for_mode(MODE);
    constexpr int LOAD_N = BLOCK_N / 2;
    CUtensorMap map_emb, map_kl;
    memset(&map_emb, 0, sizeof map_emb);
    memset(&map_kl, 0, sizeof map_kl);
    int rc;
    if (MODE != RADAR_MODE_KL) {
        rc = encode_2d_bf16(&map_emb, fl.corpus->emb_bf16, fl.corpus->d, fl.corpus->n, 64, LOAD_N,
                            CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (MODE != RADAR_MODE_DPR) {
        rc = encode_2d_bf16(&map_kl, fl.corpus->klpack, RADAR_KLPACK, fl.corpus->n, RADAR_KLPACK, LOAD_N,
                            CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    RADAR_CUDA_CHECK(cudaFuncSetAttribute(tc_filter_kernel<MODE, KB_T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem_bytes_for(MODE))));
    const int64_t items = fl.q_tiles * fl.parts;
    int64_t units = fl.units;
    if (units < 1) units = 1;
    if (units > items) units = items;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(units * 2));
    cfg.blockDim = dim3(threads_for_mode(MODE));
    cfg.dynamicSmemBytes = smem_bytes_for(MODE);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (fl.ev_start) RADAR_CUDA_CHECK(cudaEventRecord(fl.ev_start, st));
    RADAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_filter_kernel<MODE, KB_T>, map_emb, map_kl, fa));
    if (fl.ev_stop) RADAR_CUDA_CHECK(cudaEventRecord(fl.ev_stop, st));
    return RADAR_OK;
}

// per-observation bound on |log q|: the caller's column maxima when given (any non-zero entry), else logq_max_abs
static inline void fill_col_max(const radar_corpus_t* c, float* out) {
    bool given = false;
    for (int j = 0; j < kObsPad; ++j) given |= c->logq_col_max[j] > 0.0f;
    for (int j = 0; j < kObsPad; ++j) out[j] = given ? c->logq_col_max[j] * 1.0001f : c->logq_max_abs;
}

// k'-th largest group maximum per query (one warp per query, up to 512 groups in registers, bit-wise radix descent) ->
// initial threshold of the real pass
constexpr int kMaxGroups32 = 16;
__global__ void __launch_bounds__(256) group_threshold_kernel(const uint32_t* __restrict__ groupmax, int64_t q, int groups,
                                                              int kp, uint32_t* __restrict__ gthr) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (qi >= q) return;
    const uint32_t* src = groupmax + qi * groups;
    uint32_t val[kMaxGroups32];
#pragma unroll
    for (int e = 0; e < kMaxGroups32; ++e) {
        const int i = lane + 32 * e;
        val[e] = i < groups ? src[i] : 0u;
    }
    uint32_t key = 0;
    if (groups >= kp) {
#pragma unroll 1
        for (int b = 31; b >= 0; --b) {
            const uint32_t trial = key | (1u << b);
            int c = 0;
#pragma unroll
            for (int e = 0; e < kMaxGroups32; ++e) c += val[e] >= trial ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, c) >= kp) key = trial;
        }
    }
    if (lane == 0) gthr[qi] = key;  // 0 = no threshold
}

// apack region layout: [q_pad * a_cols] uint16, then (256-byte aligned) qshift [q_pad] floats
static inline size_t apack_bytes(int64_t q_pad, int mode, int d) {
    size_t b = sizeof(uint16_t) * static_cast<size_t>(q_pad) * a_cols_for(mode, d);
    b = (b + 255) / 256 * 256;
    return b + sizeof(float) * static_cast<size_t>(q_pad);
}

static int launch_filter(FilterLaunch& fl, cudaStream_t st, int* launches) {
    const int64_t q_pad = fl.q_tiles * kTileQ;
    const int cols = a_cols_for(fl.mode, fl.corpus->d);
    size_t shift_off = (sizeof(uint16_t) * static_cast<size_t>(q_pad) * cols + 255) / 256 * 256;
    float* qshift = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(fl.apack) + shift_off);
    PackArgs pa{};
    pa.q_emb = fl.queries->emb_f32; pa.p16 = fl.queries->p16; pa.entropy = fl.queries->entropy;
    pa.q = fl.q; pa.q_pad = q_pad; pa.d = fl.corpus->d; pa.mode = fl.mode; pa.alpha = fl.alpha; pa.oma = fl.oma;
    pa.emb_max_norm = fl.corpus->emb_max_norm; pa.logq_max_abs = fl.corpus->logq_max_abs;
    fill_col_max(fl.corpus, pa.logq_col_max);
    pa.apack = fl.apack; pa.qshift = qshift; pa.qerr = fl.qerr;
    query_pack_kernel<<<static_cast<unsigned>((q_pad * 32 + 255) / 256), 256, 0, st>>>(pa);
    RADAR_CUDA_CHECK(cudaGetLastError());
    RADAR_CUDA_CHECK(cudaMemsetAsync(fl.gthr, 0, gthr_region_bytes(q_pad), st));
    FilterArgs fa{};
    fa.apack = fl.apack; fa.qshift = qshift; fa.q = fl.q; fa.q_tiles = fl.q_tiles; fa.n = fl.corpus->n;
    fa.d = fl.corpus->d; fa.parts = fl.parts; fa.rows_per_part = fl.rows_per_part; fa.kp = fl.kp;
    fa.cand = fl.cand; fa.cnt = fl.cnt; fa.thr = fl.thr; fa.gthr = fl.gthr; fa.dbg_scores = fl.dbg_scores;
#ifdef RADAR_DEBUG
    fa.dbg_flags = getenv("RADAR_TC_DBG") ? atoi(getenv("RADAR_TC_DBG")) : 0;
#endif
    fa.tile_stride = 1;
    fa.progress = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(fl.gthr) +
                                                        (sizeof(uint32_t) * static_cast<size_t>(q_pad) + 7) / 8 * 8);
    fa.clk = fa.progress + kMaxUnits;
    fl.clk_dev = fa.clk;
    // window: ~16 MB of corpus tiles per slab in flight (a few slabs are swept concurrently; L2 is 126 MB).  Only
    // when every CTA of the launch is resident at once and at least two query tiles share a slab.
    const int kblocks = fl.mode != RADAR_MODE_KL ? fl.corpus->d / 64 : 0;
    const int64_t tile_bytes = 2ll * slot_stride_for(fl.mode, kblocks);
    int64_t window = (16ll << 20) / tile_bytes;
    if (window < 4 * kProgressEvery) window = 4 * kProgressEvery;
    const bool resident = 2 * fl.units <= fl.device_sms;
    fa.window = (resident && fl.q_tiles > 1) ? static_cast<int>(window) : 0;
#ifdef RADAR_DEBUG
    if (getenv("RADAR_TC_NO_WINDOW") != nullptr) fa.window = 0;
#endif
    int rc;
    const bool d512 = fl.corpus->d == 512;
    *launches = 2;
#ifdef RADAR_DEBUG
    // KL mode of the general filter: kept for the dense key dump of the debug flavour only -- KL searches run the dedicated
    // kernel of kl_filter.cuh, and the release library does not even instantiate this mode
    if (fl.mode == RADAR_MODE_KL) return launch_filter_mode<RADAR_MODE_KL, 0>(fl, fa, st);
#else
    if (fl.mode == RADAR_MODE_KL) {
        set_error("KL searches are served by the dedicated KL filter (kl_filter.cuh)");
        return RADAR_E_ARG;
    }
#endif
    auto launch_mode = [&](FilterLaunch& l, const FilterArgs& args) {
        if (l.mode == RADAR_MODE_DPR)
            return d512 ? launch_filter_mode<RADAR_MODE_DPR, 8>(l, args, st) : launch_filter_mode<RADAR_MODE_DPR, 0>(l, args, st);
        return d512 ? launch_filter_mode<RADAR_MODE_HYBRID, 8>(l, args, st) : launch_filter_mode<RADAR_MODE_HYBRID, 0>(l, args, st);
    };
    if (fl.groups > 0 && fl.dbg_scores == nullptr) {
        // SHORT SWEEPS (a few hundred thousand rows per query tile: config 3, shards of an 8-GPU run): the cold start of the
        // running thresholds -- ~k' ln(n / k') survivors per query through the rare path -- is no longer hidden behind the
        // MMAs.  A prepass of the same kernel over every tile_stride-th tile only records per query the maximum key of every
        // group of sampled tiles (no rare path at all); the k'-th largest group maximum is a valid initial threshold, and the
        // real pass sees ~k' ln(tile_stride) survivors instead.
        RADAR_CUDA_CHECK(cudaMemsetAsync(fl.groupmax, 0, sizeof(uint32_t) * static_cast<size_t>(q_pad) * fl.groups, st));
        FilterArgs fp = fa;
        fp.prepass = 1; fp.groupmax = fl.groupmax; fp.groups = fl.groups; fp.group_tiles = fl.group_tiles;
        fp.tile_stride = fl.tile_stride; fp.groups_per_slab = fl.groups_per_slab;
        // the profiled span covers the prepass, the threshold selection and the real pass
        cudaEvent_t e1 = fl.ev_stop;
        fl.ev_stop = nullptr;
        rc = launch_mode(fl, fp);
        fl.ev_start = nullptr;
        fl.ev_stop = e1;
        if (rc) return rc;
        group_threshold_kernel<<<static_cast<unsigned>((fl.q * 32 + 255) / 256), 256, 0, st>>>(fl.groupmax, fl.q, fl.groups,
                                                                                                 fl.kp, fl.gthr);
        RADAR_CUDA_CHECK(cudaGetLastError());
        // the progress words of the window protocol must start from zero again
        RADAR_CUDA_CHECK(cudaMemsetAsync(fa.progress, 0, sizeof(unsigned long long) * kMaxUnits, st));
        fa.gthr_init = 1;
        *launches += 2;
    }
    if (fl.mode == RADAR_MODE_DPR)
        rc = d512 ? launch_filter_mode<RADAR_MODE_DPR, 8>(fl, fa, st) : launch_filter_mode<RADAR_MODE_DPR, 0>(fl, fa, st);
    else
        rc = d512 ? launch_filter_mode<RADAR_MODE_HYBRID, 8>(fl, fa, st)
                  : launch_filter_mode<RADAR_MODE_HYBRID, 0>(fl, fa, st);
    if (rc) return rc;
    return RADAR_OK;
}

}  // namespace tc
}  // namespace radar
