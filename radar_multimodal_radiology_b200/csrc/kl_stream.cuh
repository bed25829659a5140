// kl_stream.cuh -- KL observation retrieval for FEW queries (q <= 256) over a LARGE corpus: the HBM-bound regime
// ("latency regime": 64 B of log-probabilities per case and pass, AI ~ q/2 flop/B).
//
// The general filter (tc_filter.cuh) puts queries on the M side of the MMA and gives every query row a private
// candidate buffer per corpus slab; with a handful of queries that leaves 7 of 8 epilogue warps idle and lets every
// slab warm its thresholds up on its own.  Here the roles are swapped and the candidates are pooled:
//
//   boot    (tcgen05)      the stream kernel's BOOT instantiation over a strided SAMPLE of super-tiles: per query the maximum
//                          filter key of every sampled super-tile.  The k'-th largest tile maximum is a valid lower bound of
//                          the k'-th best filter key of the whole corpus (tile maxima are distinct cases) -> initial thresholds.
//   stream  (tcgen05)      CTA pairs sweep the corpus once.  A operand = 256 corpus rows of the bf16 [hi|lo] log table
//                          (TMA -> smem ring), B operand = the packed queries (resident in smem), D = [256 cases x N
//                          queries] fp32 in TMEM, 512/N accumulator stages.  Epilogue thread = one case: it compares
//                          its N keys with the per-query thresholds (broadcast LDS) -- one predicate chain and one vote
//                          per 32 queries -- and only survivors are appended (atomicAdd slot) to ONE pooled buffer per
//                          query in global memory.  Every 64 appends the appending warp tries a per-query lock
//                          and folds the new entries into the query's running best-k' list with the register radix
//                          select; the k'-th key is published with atomicMax and picked up by all warps.
//   final                  per query: canonical fp32 re-score of every pooled entry, block radix select of the k best,
//                          sort, certificate (fp32 mode) / overflow check -> uncertified queries are re-run exactly.
//
// Algorithmic bytes: 64 B per case (the [hi|lo] row) + outputs.  Roofline: HBM.
#pragma once
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "scan_kernels.cuh"
#include "tc_filter.cuh"
#include "kl_filter.cuh"

namespace radar {
namespace kls {

using namespace tc;  // PTX wrappers, descriptors, cluster helpers
using klf::kFmtBf16x3;
using klf::kFmtF16x1;
using klf::kFmtF16x2;

constexpr int kTileRows = 512;          // corpus rows per super-tile: 256 per CTA = two M = 256 MMA sub-tiles
constexpr int kSubRows = 256;           // corpus rows per MMA (128 per CTA)
constexpr int kSlots = 8;               // smem ring slots per CTA: 256 rows x 64 B ([hi|lo] bf16 table) or x 32 B (fp16 table)
__host__ __device__ constexpr int slot_bytes(int fmt) { return 256 * klf::row_bytes(fmt); }
constexpr int kMaxN = 256;              // queries per launch (MMA N)
#ifndef RADAR_KLS_ISSUERS
#define RADAR_KLS_ISSUERS 2
#endif
#ifndef RADAR_KLS_DEFER
#define RADAR_KLS_DEFER 0
#endif
// park a pooled append until the thread's next visit of the rare path instead of waiting ~700 cycles for its slot number.
// Measured: SLOWER (0.125 -> 0.151 ms on kl_latency) -- entries reach the pool late, refreshes see fewer of them, thresholds
// tighten later and more cases survive; flushing at the next chunk instead (one tile later) measured no different from not
// deferring at all.  Kept as a build option.
constexpr bool kDeferAppends = RADAR_KLS_DEFER != 0;
constexpr int kIssuers = RADAR_KLS_ISSUERS;  // MMA issuer warps (warp 1 and warp 11) taking alternate super-tiles
constexpr int kStreamThreads = 384;     // warp 0 TMA, warps 1 / 11 MMA, warps 2-5 / 6-9 epilogue sets (alternating super-tiles),
                                        // warp 10 threshold refresher
#ifndef RADAR_KLS_REFRESH_EVERY
#define RADAR_KLS_REFRESH_EVERY 64
#endif
#ifndef RADAR_KLS_RELOAD
#define RADAR_KLS_RELOAD 4
#endif
#ifndef RADAR_KLS_MAX_SAMPLE
#define RADAR_KLS_MAX_SAMPLE 2048
#endif
constexpr int kRefreshEvery = RADAR_KLS_REFRESH_EVERY;  // appends of a query between threshold refresh attempts
constexpr int kThrReload = RADAR_KLS_RELOAD;            // super-tiles between reloads of the published thresholds
constexpr int kMinRows = 1 << 16;       // smaller corpora use the general path
constexpr int kBootThreads = 256;
constexpr int kMaxSample = RADAR_KLS_MAX_SAMPLE;  // sample tiles of the boot pass (k'-th largest tile maximum = first threshold)

__host__ __device__ inline int pool_cap_for(int n_pad) {  // entries of one query's pooled buffer
    int c = (1 << 20) / n_pad;
    return c > 8192 ? 8192 : (c < 2048 ? 2048 : c);
}
inline int sample_tiles_for(int n_pad, int64_t tiles) {  // sampled super-tiles of the boot pass: ~10 % of the shard
    (void)n_pad;
    int64_t s = tiles / 10;
    if (s < 256) s = 256;
    if (s > kMaxSample) s = kMaxSample;
    return static_cast<int>(s < tiles ? s : tiles);
}

// Programmatic dependent launch (the KL stream chain is latency-bound: six short kernels around a 0.12 ms sweep).  A kernel
// launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor still runs; it must call
// pdl_wait() before it touches anything the predecessor writes (a no-op when launched without the attribute).  pdl_trigger()
// lets the NEXT kernel in the stream be scheduled from now on.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- boot threshold: k'-th largest sampled super-tile maximum per query -------------------------------------------
// The maxima come from the BOOT instantiation of the stream kernel (filter keys of a strided sample of super-tiles): the
// k'-th largest is a valid lower bound of the k'-th best filter key of the whole corpus (tile maxima are distinct cases).
struct BootArgs {
    const uint32_t* tilemax;  // [q][sample_tiles] ord-encoded filter keys (0 = no valid row)
    int sample_tiles, kp;
    uint32_t* gthr;           // [q] out
};

// one CTA per query, values in registers, radix descent two bits per step (three independent counts, one block-wide sum)
__global__ void __launch_bounds__(kBootThreads) kl_boot_threshold_kernel(const BootArgs a) {
    static_assert(kMaxSample % kBootThreads == 0, "sample tiles per thread");
    __shared__ int cnt[2][3][kBootThreads / 32];
    const int qi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();  // the stream kernel may set up its barriers / tensor memory while the thresholds are selected
    pdl_wait();     // the boot sweep's maxima
    const uint32_t* src = a.tilemax + static_cast<int64_t>(qi) * a.sample_tiles;
    uint32_t val[kMaxSample / kBootThreads];
#pragma unroll
    for (int e = 0; e < kMaxSample / kBootThreads; ++e) {
        const int i = tid + kBootThreads * e;
        val[e] = i < a.sample_tiles ? src[i] : 0u;
    }
    uint32_t key = 0;
    if (a.sample_tiles >= a.kp) {
        int buf = 0;
#pragma unroll 1
        for (int b = 30; b >= 0; b -= 2, buf ^= 1) {
            const uint32_t t1 = key | (1u << b), t2 = key | (2u << b), t3 = key | (3u << b);
            int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
            for (int e = 0; e < kMaxSample / kBootThreads; ++e) {
                c1 += val[e] >= t1 ? 1 : 0;
                c2 += val[e] >= t2 ? 1 : 0;
                c3 += val[e] >= t3 ? 1 : 0;
            }
            c1 = __reduce_add_sync(0xffffffffu, c1);
            c2 = __reduce_add_sync(0xffffffffu, c2);
            c3 = __reduce_add_sync(0xffffffffu, c3);
            if (lane == 0) {
                cnt[buf][0][warp] = c1;
                cnt[buf][1][warp] = c2;
                cnt[buf][2][warp] = c3;
            }
            __syncthreads();  // the other buffer is rewritten only after the next barrier: one barrier per step is enough
            c1 = c2 = c3 = 0;
#pragma unroll
            for (int w = 0; w < kBootThreads / 32; ++w) {
                c1 += cnt[buf][0][w];
                c2 += cnt[buf][1][w];
                c3 += cnt[buf][2][w];
            }
            key = c3 >= a.kp ? t3 : (c2 >= a.kp ? t2 : (c1 >= a.kp ? t1 : key));
        }
    }
    if (tid == 0) a.gthr[qi] = key;  // 0 = no threshold
}

// ---- stream -------------------------------------------------------------------------------------------------
struct StreamArgs {
    int64_t n, tiles;            // corpus rows, 512-row super-tiles
    int q, n_pad;                // real queries, MMA N (32 / 64 / 128 / 256)
    int kp, pool_cap;
    const float* qshift;         // [n_pad] entropy (0 for padding rows)
    uint32_t* gthr;              // [n_pad] published thresholds, ord-encoded canonical-key units (0 = none)
    uint32_t* gcnt;              // [n_pad] appended entries per query (may exceed pool_cap: overflow)
    uint32_t* lock;              // [n_pad]
    uint32_t* processed;         // [n_pad] pooled entries already folded into the best list
    uint32_t* best_n;            // [n_pad]
    uint64_t* best;              // [n_pad][kCandCap] running best-k' list of every query (touched under the lock only)
    uint64_t* pool;              // [n_pad][pool_cap]
    int reload;                  // super-tiles (of one epilogue set) between reloads of the published thresholds
    // BOOT instantiation only: `tiles` counts the SAMPLED super-tiles (sample s = super-tile s * total_tiles / tiles)
    int64_t total_tiles;
    uint32_t* tilemax;           // [q][tiles] ord-encoded maximum filter key of every sampled super-tile (zeroed by the caller)
};

constexpr size_t kStreamSmemBytes = 1024 + static_cast<size_t>(kSlots) * slot_bytes(kFmtBf16x3) + 128 * 64 /*queries*/ +
                                    8 * 32 * 32 * sizeof(float) /*chunk staging*/ + 8 * kMaxN * sizeof(float) /*thresholds*/ +
                                    kMaxN * sizeof(float) /*entropy*/ + kCandCap * sizeof(uint64_t) /*refresh scratch*/ +
                                    1024 /*barriers + refresh requests*/;
static_assert(kStreamSmemBytes <= 227 * 1024, "shared memory budget");

// D[tmem] (+)= A[smem] * B[smem]^T, M = 256 across the CTA pair
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float redux_fmax(float v) {  // sm_100a: warp-wide float maximum in one instruction
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// fold the not yet processed pooled entries of query qi into its best-k' list and publish the k'-th key
// (warp-cooperative; a no-op when another warp holds the query's lock)
__device__ __noinline__ void refresh_threshold(const StreamArgs& a, int qi, uint64_t* sc, int lane) {
    uint32_t got = 0;
    if (lane == 0) got = atomicCAS(a.lock + qi, 0u, 1u) == 0u ? 1u : 0u;
    got = __shfl_sync(0xffffffffu, got, 0);
    if (!got) return;
    __threadfence();
    uint32_t start = ld_relaxed_u32(a.processed + qi);
    int nb = static_cast<int>(ld_relaxed_u32(a.best_n + qi));
    const uint32_t end = min(ld_relaxed_u32(a.gcnt + qi), static_cast<uint32_t>(a.pool_cap));
    uint64_t* best = a.best + static_cast<int64_t>(qi) * kCandCap;
    const uint64_t* pool = a.pool + static_cast<int64_t>(qi) * a.pool_cap;
    for (int i = lane; i < nb; i += 32) sc[i] = __ldcg(best + i);
    float t = 0.0f;
    bool have = false;
    int rounds = 0;
    while (start < end && rounds < 8) {  // bounded: the lock is never held for long
        const int take = min(static_cast<int>(end - start), kCandCap - nb);
        for (int i = lane; i < take; i += 32) sc[nb + i] = __ldcg(pool + start + i);  // unwritten slots read as 0
        nb += take;
        start += take;
        ++rounds;
        __syncwarp();
        if (nb > a.kp) {
            t = warp_compact(sc, nb, a.kp, lane);
            nb = a.kp;
            have = true;
        }
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) __stcg(best + i, sc[i]);
    __syncwarp();
    if (lane == 0) {
        a.processed[qi] = start;
        a.best_n[qi] = static_cast<uint32_t>(nb);
        if (have) atomicMax(a.gthr + qi, f2ord(t));
        __threadfence();
        atomicExch(a.lock + qi, 0u);
    }
    __syncwarp();
}

// FMT (klf::kFmt*): bf16 hi/lo x 3 products on klpack (64 B rows), or fp16 x 1 / x 2 products on kl16 (32 B rows: half
// the HBM bytes per case); the fp16 accumulators are scaled by 2^24 (kl_filter.cuh), thresholds are compared in scaled units
#ifdef RADAR_KLS_TIMING
// experiment aid (tools/kls_timing.py): per CTA [start, after prologue, end, rare-path visits] of the last stream launch
__device__ unsigned long long g_kls_timing[296 * 4];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif
// BOOT: the same pipeline over a strided SAMPLE of super-tiles; the epilogue only reduces, per query, the maximum filter key of
// every sampled super-tile (redux.sync.max.f32 over the 32 cases of a warp, one atomicMax per query and warp) -> tilemax.
template <int FMT, bool BOOT>
__global__ void __launch_bounds__(kStreamThreads, 1)
kl_stream_kernel(const __grid_constant__ CUtensorMap map_kl, const __grid_constant__ CUtensorMap map_q,
                 const StreamArgs a) {
    constexpr int kSlotBytes = slot_bytes(FMT);
    constexpr int ROW_B = klf::row_bytes(FMT);
    constexpr float SCALE = FMT == kFmtBf16x3 ? 1.0f : klf::kAccScale;
    constexpr float INV = FMT == kFmtBf16x3 ? 1.0f : klf::kAccInv;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                                        // [kSlots][16 KB]
    uint8_t* qtile = ring + kSlots * slot_bytes(kFmtBf16x3);                     // [n_pad/2 rows][64 B], SW64
    float* stage = reinterpret_cast<float*>(qtile + 128 * 64);                   // [8 warps][32 cols][32 lanes]
    float* thr_s = stage + 8 * 32 * 32;                                          // [8 warps][kMaxN] accumulator units
    float* h_s = thr_s + 8 * kMaxN;                                              // [kMaxN]
    uint64_t* scratch = reinterpret_cast<uint64_t*>(h_s + kMaxN);                // [kCandCap] (refresher warp)
    uint64_t* bars = scratch + kCandCap;
    uint64_t* full_bar = bars;                // [kSlots]
    uint64_t* empty_bar = full_bar + kSlots;  // [kSlots]
    uint64_t* tfull_bar = empty_bar + kSlots; // [8] one per accumulator stage PAIR
    uint64_t* tempty_bar = tfull_bar + 8;     // [8] 8 arrivals: the 4 warps of one epilogue set x 2 CTAs
    uint64_t* qfull_bar = tempty_bar + 8;     // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(qfull_bar + 1);
    uint32_t* refresh_req = tmem_slot + 2;    // [kMaxN / 32] bit per query: "fold my new entries, publish a threshold"
    uint32_t* epi_done = refresh_req + kMaxN / 32;  // epilogue warps that have finished

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int64_t unit = blockIdx.x >> 1, units = gridDim.x >> 1;
    const int N = a.n_pad;
    const int spairs = kTmemCols / (2 * N);  // accumulator stage pairs (1 .. 8)
    const uint32_t idesc = FMT == kFmtBf16x3 ? make_idesc_mn(kSubRows, N) : klf::make_idesc_f16_mn(kSubRows, N);

#ifdef RADAR_KLS_TIMING
    if (!BOOT && threadIdx.x == 0) { g_kls_timing[blockIdx.x * 4 + 0] = gtimer(); g_kls_timing[blockIdx.x * 4 + 3] = 0ull; }
#endif
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_kl);
        prefetch_tmap(&map_q);
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 8; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);
        }
        mbar_init(qfull_bar, 1);
        fence_barrier_init();
        for (int i = 0; i < kMaxN / 32; ++i) refresh_req[i] = 0u;
        *epi_done = 0u;
    }
    pdl_trigger();  // BOOT: the threshold kernel's CTAs may be scheduled; stream: the final kernel's (they wait for this grid)
    if (warp == 1) tmem_alloc_pair(tmem_slot);
    for (int i = threadIdx.x; i < N; i += kStreamThreads) h_s[i] = a.qshift[i];  // written by the pack kernel: long complete
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (!BOOT) pdl_wait();  // everything above overlapped the threshold kernel; the thresholds are read below
#ifdef RADAR_KLS_TIMING
    if (!BOOT && threadIdx.x == 0) g_kls_timing[blockIdx.x * 4 + 1] = gtimer();
#endif
    if (tmem_base != 0) {
        if (threadIdx.x == 0) printf("radar kl_stream: unexpected TMEM base %u\n", tmem_base);
        __trap();
    }

    if (warp == 0) {
        // ================================ TMA producer: one 16 KB box (256 rows) per CTA and super-tile ==============
        if (elect_one()) {
            if (leader) mbar_expect_tx(qfull_bar, static_cast<uint32_t>(N / 2 * 64) * 2);
            tma_load_2d_pair(&map_q, smem_u32(qfull_bar), smem_u32(qtile), 0, static_cast<int>(cta_rank) * (N / 2));
        }
        __syncwarp();
        uint32_t slot = 0, sph = 0;
        for (int64_t t = unit; t < a.tiles; t += units) {
            mbar_wait(&empty_bar[slot], sph ^ 1);
            if (elect_one()) {
                if (leader) mbar_expect_tx(&full_bar[slot], kSlotBytes * 2);
                const int64_t tt = BOOT ? t * a.total_tiles / a.tiles : t;
                tma_load_2d_pair(&map_kl, smem_u32(&full_bar[slot]), smem_u32(ring + slot * kSlotBytes), 0,
                                 static_cast<int>(tt * kTileRows) + static_cast<int>(cta_rank) * 256);
            }
            __syncwarp();
            if (++slot == kSlots) {
                slot = 0;
                sph ^= 1;
            }
        }
    } else if (warp == 1 || warp == 11) {
        // ================================ MMA issuers (leader CTA) ================================
        // A super-tile costs its issuer two barrier waits, the MMAs and two commits -- several hundred cycles of serial
        // latency that has nothing to do with the tensor pipe.  Two issuer warps take alternate super-tiles (tcgen05.commit
        // tracks the MMAs of the issuing thread, so each warp signals exactly its own tiles); slots and stage pairs are
        // revisited by the same issuer (their counts are even), which keeps the parity waits sound.  A single stage pair
        // (256 queries) is shared by consecutive tiles: one issuer only.
        const uint32_t ii = warp == 1 ? 0u : 1u;
        const uint32_t nissue = (kIssuers > 1 && spairs > 1) ? 2u : 1u;
        if (leader && ii < nissue) {
            mbar_wait(qfull_bar, 0);
            tc_fence_after();
            const uint64_t bdesc = make_smem_desc(smem_u32(qtile), 512, 4);
            const uint32_t sp_mask = static_cast<uint32_t>(spairs) - 1u, sp_shift = 31u - __clz(static_cast<uint32_t>(spairs));
            for (uint32_t j = ii;; j += nissue) {
                const int64_t t = unit + static_cast<int64_t>(j) * units;
                if (t >= a.tiles) break;
                const uint32_t slot = j % kSlots, sph = (j / kSlots) & 1u;
                const uint32_t sp = j & sp_mask, aph = (j >> sp_shift) & 1u;
                mbar_wait(&tempty_bar[sp], aph ^ 1);
                mbar_wait(&full_bar[slot], sph);
                tc_fence_after();
                const uint64_t adesc = FMT == kFmtBf16x3 ? make_smem_desc(smem_u32(ring + slot * kSlotBytes), 512, 4)   // SW64
                                                         : make_smem_desc(smem_u32(ring + slot * kSlotBytes), 256, 6);  // SW32
                const uint32_t d0 = static_cast<uint32_t>(sp * 2 * N);
                if (elect_one()) {
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {  // rows [sub*128, sub*128+128) of each CTA's box
                        const uint64_t ad = adesc + static_cast<uint64_t>((sub * 128 * ROW_B) >> 4);
                        const uint32_t d = d0 + static_cast<uint32_t>(sub * N);
                        if (FMT == kFmtBf16x3) {
                            umma_ss_pair(d, ad, bdesc, idesc, 0u);      // L_hi . v_hi
                            umma_ss_pair(d, ad + 2, bdesc, idesc, 1u);  // L_lo . v_hi
                            umma_ss_pair(d, ad, bdesc + 2, idesc, 1u);  // L_hi . v_lo
                        } else {
                            umma_ss_pair(d, ad, bdesc, idesc, 0u);      // L16 . v_hi
                            if (FMT == kFmtF16x2) umma_ss_pair(d, ad, bdesc + 2, idesc, 1u);  // L16 . v_lo
                        }
                    }
                    umma_commit_pair(&empty_bar[slot]);
                    umma_commit_pair(&tfull_bar[sp]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 10) {
        // ================================ threshold refresher ================================
        // Folding pooled entries into a query's best-k' list takes microseconds; done by an epilogue warp it would hold
        // up the accumulator hand-off of the whole CTA pair, so the epilogue warps only raise a request bit.
        while (!BOOT) {
            bool any = false;
            for (int w = 0; w < N / 32; ++w) {
                uint32_t bits = 0;
                if (lane == 0) bits = atomicExch(&refresh_req[w], 0u);
                bits = __shfl_sync(0xffffffffu, bits, 0);
                any |= bits != 0u;
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    refresh_threshold(a, w * 32 + b, scratch, lane);
                }
            }
            if (!any) {
                uint32_t d = 0;
                if (lane == 0) d = *reinterpret_cast<volatile uint32_t*>(epi_done);
                if (__shfl_sync(0xffffffffu, d, 0) == 8u) break;
                __nanosleep(200);
            }
        }
    } else {
        // ===================== epilogue warps: thread = corpus row ====================================================
        // Two sets of four warps (2-5 and 6-9, one warp per TMEM lane quadrant) alternate over the super-tiles of the
        // pair; a warp reads BOTH M = 256 sub-tiles of its super-tile (two back-to-back tcgen05.ld per 32 queries), so
        // one barrier round trip and one threshold fetch serve 2 x 32 cases per lane quadrant.
        const int quad = warp & 3;
        const int ew = warp - 2;           // 0..7
        const int set = ew >> 2;           // which super-tiles (local index parity) this warp takes
        const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
        float* my_stage = stage + ew * 32 * 32 + lane;
        float* my_thr = thr_s + ew * kMaxN;
        const uint32_t my_thr_addr = smem_u32(my_thr);
        const int nq32 = N / 32;  // threshold words per lane
        // thresholds in accumulator units: key + H
        uint32_t pending[kMaxN / 32];
#pragma unroll
        for (int i = 0; i < kMaxN / 32; ++i) pending[i] = (!BOOT && i < nq32) ? ld_relaxed_u32(a.gthr + lane + 32 * i) : 0u;
        auto commit_thresholds = [&]() {
#pragma unroll
            for (int i = 0; i < kMaxN / 32; ++i) {
                if (i < nq32) {
                    const int qi = lane + 32 * i;
                    const float tk = pending[i] ? ord2f(pending[i]) : -CUDART_INF_F;
                    // a few ulps below fl(key + H): fl(fl(acc - H) + H) may round above acc, and cases tied with the
                    // threshold must not be lost (a lower threshold is always safe)
                    const float ta = __fadd_rn(tk, h_s[qi]);
                    my_thr[qi] = qi < a.q ? (pending[i] ? (ta - 4e-7f * (fabsf(ta) + fabsf(h_s[qi]))) * SCALE : -CUDART_INF_F)
                                          : CUDART_INF_F;
                }
            }
            __syncwarp();
        };
        // A pooled append needs a slot number from a global atomicAdd (~700 cycles).  The thread does not wait for it: the
        // atomic is issued, the entry is parked in registers, and the store into the pool happens when the thread next
        // comes through here (or at the end) -- by then the slot number has long arrived.  A slot that is still unwritten
        // when a refresh looks at it reads as 0 ("empty"); the final pass rescans the whole pool, so nothing is lost.
        uint32_t pend_slot = 0, pend_qi = 0;
        uint64_t pend_comp = 0ull;  // 0 = nothing parked
        auto flush_pending = [&]() {
            if (pend_comp != 0ull) {
                if (pend_slot < static_cast<uint32_t>(a.pool_cap))
                    a.pool[static_cast<int64_t>(pend_qi) * a.pool_cap + pend_slot] = pend_comp;
                if (((pend_slot + 1) % kRefreshEvery) == 0 && pend_slot + 1 >= static_cast<uint32_t>(a.kp))
                    atomicOr(&refresh_req[pend_qi >> 5], 1u << (pend_qi & 31));
                pend_comp = 0ull;
            }
        };
        // rare path of one 32-case x 32-query chunk held in registers.  hits = which of the four predicate chains (query
        // column mod 4) saw a survivor in this lane: only the columns of chains that fired in SOME lane are staged and
        // compared again (typically 8 of 32), then the survivor bits are walked.
        auto append_survivors = [&](const float (&v)[32], int cb, int64_t row, bool row_ok, uint32_t hits) {
            uint32_t mask = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (!__any_sync(0xffffffffu, (hits >> r) & 1u)) continue;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int c = 4 * c4 + r;
                    my_stage[c * 32] = v[c];
                    mask |= (v[c] >= my_thr[cb * 32 + c] ? 1u : 0u) << c;
                }
            }
            if (!row_ok) mask = 0;
            while (mask) {
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                const int qi = cb * 32 + c;
                const float key = __fsub_rn(my_stage[c * 32] * INV, h_s[qi]);
                flush_pending();
                pend_slot = atomicAdd(a.gcnt + qi, 1u);
                pend_qi = static_cast<uint32_t>(qi);
                pend_comp = make_composite(key, static_cast<uint32_t>(row));
                if (!kDeferAppends) flush_pending();
            }
            __syncwarp();
        };
        if (!BOOT) commit_thresholds();
        // A waiter must see every phase of its barriers: with an even number of stage pairs (a power of two) a set meets
        // each of "its" stage pairs on every use; with a single stage pair (N = 256) set 0 takes every super-tile.
        const uint32_t jstep = spairs > 1 ? 2u : 1u;
        const uint32_t sp_mask = static_cast<uint32_t>(spairs) - 1u, sp_shift = 31u - __clz(static_cast<uint32_t>(spairs));
        const bool idle = spairs == 1 && set != 0;
        uint32_t since = 0;
        for (uint32_t j = spairs > 1 ? static_cast<uint32_t>(set) : 0u; !idle; j += jstep) {
            const int64_t t = unit + static_cast<int64_t>(j) * units;
            if (t >= a.tiles) break;
            const uint32_t sp = j & sp_mask, aph = (j >> sp_shift) & 1u;
            if (!BOOT && ++since == static_cast<uint32_t>(a.reload)) {  // use the values requested a few super-tiles ago, request fresh ones
                since = 0;
                commit_thresholds();
#pragma unroll
                for (int i = 0; i < kMaxN / 32; ++i) pending[i] = i < nq32 ? ld_relaxed_u32(a.gthr + lane + 32 * i) : 0u;
            }
            const int64_t row0 = (BOOT ? t * a.total_tiles / a.tiles : t) * kTileRows + static_cast<int64_t>(cta_rank) * 256 + quad * 32 + lane;
            const int64_t row1 = row0 + 128;
            const bool ok0 = row0 < a.n, ok1 = row1 < a.n;
            mbar_wait(&tfull_bar[sp], aph);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + lane_addr + static_cast<uint32_t>(sp * 2 * N);
            for (int cb = 0; cb < nq32; ++cb) {
                float v0[32], v1[32];
                tmem_ld_x32(t_acc + cb * 32, v0);
                tmem_ld_x32(t_acc + N + cb * 32, v1);
                tmem_wait_ld();
                if (cb == nq32 - 1) {  // both accumulator stages are free again once their last columns are in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tempty_bar[sp]);
                }
                if (BOOT) {
                    if (!__all_sync(0xffffffffu, ok0 && ok1)) {  // rows past the end of the corpus arrive as zeros from TMA
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            if (!ok0) v0[c] = -CUDART_INF_F;
                            if (!ok1) v1[c] = -CUDART_INF_F;
                        }
                    }
                    float mine = -CUDART_INF_F;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float m = redux_fmax(fmaxf(v0[c], v1[c]));
                        if (lane == c) mine = m;
                    }
                    const int qi = cb * 32 + lane;
                    if (qi < a.q && mine > -CUDART_INF_F)
                        atomicMax(a.tilemax + static_cast<int64_t>(qi) * a.tiles + t, f2ord(__fsub_rn(mine * INV, h_s[qi])));
                    continue;
                }
                bool g0 = false, g1 = false, g2 = false, g3 = false;  // independent predicate chains, sub-tile 0
                bool h0 = false, h1 = false, h2 = false, h3 = false;  // sub-tile 1
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 th = lds_f4(my_thr_addr + (cb * 32 + 4 * c4) * 4);  // broadcast
                    g0 |= v0[4 * c4 + 0] >= th.x;
                    g1 |= v0[4 * c4 + 1] >= th.y;
                    g2 |= v0[4 * c4 + 2] >= th.z;
                    g3 |= v0[4 * c4 + 3] >= th.w;
                    h0 |= v1[4 * c4 + 0] >= th.x;
                    h1 |= v1[4 * c4 + 1] >= th.y;
                    h2 |= v1[4 * c4 + 2] >= th.z;
                    h3 |= v1[4 * c4 + 3] >= th.w;
                }
                const bool hit0 = (g0 | g1 | g2 | g3) && ok0, hit1 = (h0 | h1 | h2 | h3) && ok1;
#ifdef RADAR_KLS_NORARE
                if (false) {
#else
                if (__any_sync(0xffffffffu, hit0 | hit1)) {
#endif
#ifdef RADAR_KLS_TIMING
                    if (lane == 0) atomicAdd(&g_kls_timing[blockIdx.x * 4 + 3], 1ull);
#endif
                    if (__any_sync(0xffffffffu, hit0))
                        append_survivors(v0, cb, row0, ok0, ok0 ? (g0 ? 1u : 0u) | (g1 ? 2u : 0u) | (g2 ? 4u : 0u) | (g3 ? 8u : 0u) : 0u);
                    if (__any_sync(0xffffffffu, hit1))
                        append_survivors(v1, cb, row1, ok1, ok1 ? (h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u) : 0u);
                }
            }
        }
        if (!BOOT) {
            flush_pending();
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(epi_done, 1u);
        }
    }
    tc_fence_before();
    __syncthreads();
#ifdef RADAR_KLS_TIMING
    if (!BOOT && threadIdx.x == 0) g_kls_timing[blockIdx.x * 4 + 2] = gtimer();
#endif
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base);
}

// ---- final: survivors of the last published threshold -> canonical re-score -> best k, certificate ---------------
struct StreamFinalArgs {
    const float* p16;
    const float* entropy;
    const float* logq16;
    const float* qerr;        // [n_pad]
    const uint32_t* gthr;
    const uint32_t* gcnt;
    const uint64_t* pool;
    int q, k, pool_cap, certify;
    int64_t idx_offset;
    float* out_scores;
    int64_t* out_idx;
    uint64_t* out_packed;  // nullable
    uint32_t* uncert_count;
    uint32_t* uncert_list;
};

constexpr int kFinThreads = 256;
constexpr int kFinCap = 2048;  // survivors the sort can take; more (no refresh ever ran and the pool is large) -> exact re-run

__global__ void __launch_bounds__(kFinThreads) kl_stream_final_kernel(const StreamFinalArgs a) {
    __shared__ float ps[kObsPad];
    __shared__ uint64_t surv[kFinCap];
    __shared__ int surv_n;
    const int qi = blockIdx.x, tid = threadIdx.x;
    pdl_wait();  // the stream kernel's pools, counts and final thresholds
    if (tid < kObsPad) ps[tid] = a.p16[qi * kObsPad + tid];
    if (tid == 0) surv_n = 0;
    __syncthreads();
    const float h = a.entropy[qi];
    const uint32_t cnt = a.gcnt[qi];
    const uint32_t g = a.gthr[qi];
    const int n_e = static_cast<int>(min(cnt, static_cast<uint32_t>(a.pool_cap)));
    const uint64_t* pool = a.pool + static_cast<int64_t>(qi) * a.pool_cap;
    // The final threshold is a lower bound of the k'-th best filter key, so only pooled entries at or above it can be
    // among the best k' by filter key: everything else was admitted under an older, looser threshold.
    for (int i = tid; i < n_e; i += kFinThreads) {
        const uint64_t c = pool[i];
        if (c != 0ull && static_cast<uint32_t>(c >> 32) >= g) {
            const int pos = atomicAdd(&surv_n, 1);
            if (pos < kFinCap) {
                const uint32_t row = composite_row(c);
                const float4* ll = reinterpret_cast<const float4*>(a.logq16 + static_cast<int64_t>(row) * kObsPad);
                const float4 l0 = __ldg(ll), l1 = __ldg(ll + 1), l2 = __ldg(ll + 2), l3 = __ldg(ll + 3);
                const float lv[16] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w,
                                      l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
                float x = 0.0f;
#pragma unroll
                for (int j = 0; j < kNumObs; ++j) x = __fmaf_rn(ps[j], lv[j], x);
                surv[pos] = make_composite(__fsub_rn(x, h), row);  // canonical key
            }
        }
    }
    __syncthreads();
    const int ns_all = surv_n;
    const int ns = min(ns_all, kFinCap);
    const int P = max(next_pow2(max(ns, 1)), 2);
    for (int i = ns + tid; i < P; i += kFinThreads) surv[i] = 0ull;
    bitonic_sort_desc(surv, P, tid, kFinThreads, [] { __syncthreads(); });
    for (int j = tid; j < a.k; j += kFinThreads) {
        const uint64_t c = j < ns ? surv[j] : 0ull;
        a.out_scores[static_cast<int64_t>(qi) * a.k + j] = c ? api_score_from_key(RADAR_MODE_KL, composite_key(c)) : CUDART_INF_F;
        a.out_idx[static_cast<int64_t>(qi) * a.k + j] = c ? static_cast<int64_t>(composite_row(c)) + a.idx_offset : -1;
        if (a.out_packed) a.out_packed[static_cast<int64_t>(qi) * a.k + j] = c ? packed_global(c, a.idx_offset) : 0ull;
    }
    if (tid == 0) {
        // an overflowed pool may have lost candidates; too many survivors cannot be sorted here
        bool ok = cnt <= static_cast<uint32_t>(a.pool_cap) && ns_all <= kFinCap && ns >= a.k;
        if (ok && a.certify && g != 0u) {
            // every case that is NOT among the survivors has filter key < the final threshold, |canonical - filter| <= qerr
            ok = ord2f(g) + a.qerr[qi] < composite_key(surv[a.k - 1]);
        }
        if (!ok) {
            const uint32_t slot = atomicAdd(a.uncert_count, 1u);
            a.uncert_list[slot] = static_cast<uint32_t>(qi);
        }
    }
}

}  // namespace kls
}  // namespace radar
