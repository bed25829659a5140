// scan_kernels.cuh -- the exact CUDA-core scan (canonical fp32 arithmetic) and the candidate
// post-processing kernels shared with the tensor-core filter.
//
// Pipeline of one search call (Q x N scores are never written to HBM):
//   scan / filter kernel : per (query tile, corpus slab) keeps a thresholded candidate buffer
//                          cand[query][slab][C] of 64-bit composites, compacted to the best k' whenever
//                          it passes C_SOFT entries (threshold <- k'-th best key)
//   select_kernel        : per query, best R composites over all slabs (+ bound on everything dropped)
//   rescore_kernel       : (filter path only) canonical fp32 key of each selected candidate
//   final_kernel         : per query sort by canonical key, write top-k (+ certificate on the filter path)
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace radar {

constexpr int kCandCap = 256;   // capacity C of one (query, slab) candidate buffer
constexpr int kCandSoft = 192;  // compaction trigger; C - C_SOFT >= rows a tile can append per query

constexpr int kScanTQ = 64;  // queries per CTA tile
constexpr int kScanTC = 64;  // corpus rows per tile
constexpr int kScanKC = 32;  // embedding columns staged per step
constexpr int kScanThreads = 256;
constexpr int kScanLd = kScanTQ + 4;

struct ScanArgs {
    const float* q_emb;
    const float* p16;
    const float* entropy;
    const float* c_emb;
    const float* logq16;
    const uint32_t* qmap;  // nullable: tile row -> query id (exact re-run of uncertified queries)
    const uint32_t* nq_dev;  // nullable: DEVICE-side number of tile rows in use (<= nq); lets the re-run of uncertified
                             // queries be enqueued without a host round trip -- CTAs beyond the count exit at once
    uint32_t nq_skip;      // with nq_dev: the first nq_skip listed queries belong to another launch (count = *nq_dev - nq_skip)
    int64_t nq;            // tile rows in use (upper bound when nq_dev is set)
    int64_t n;             // corpus rows
    int d;
    int mode;
    float alpha, oma;
    int parts;
    int64_t rows_per_part;  // multiple of kScanTC
    int kp;                 // entries kept by a compaction (= k on this exact path)
    uint64_t* cand;         // [nq][parts][kCandCap]
    uint32_t* cnt;          // [nq][parts]
    // "search after" (radar_queries_t::after_*, nullable): only cases ranking strictly after (score, id) are kept
    const float* after_scores;
    const int64_t* after_idx;
    int64_t idx_offset;
};

constexpr size_t kScanSmemBytes =
    sizeof(float) * (2 * kScanKC * kScanLd + 2 * kObsPad * kScanTQ + 2 * kScanTQ) + sizeof(uint32_t) * 3 * kScanTQ +
    sizeof(long long) * kScanTQ;

// warp-cooperative compaction of one candidate buffer (n <= kCandCap composites in global memory, written by lanes
// of this warp before a __syncwarp): keep exactly the best kp, return the kp-th key.  Selection, not sorting: the
// composites sit in registers (8 per lane) and the kp-th largest is found by a bit-wise radix descent with one
// warp-wide population count per bit -- first over the 32 key bits, then (only when equal keys straddle the cut)
// over the 32 row bits.  About 4x cheaper than the shared-memory bitonic sort it replaces, and it needs no scratch.
// (not inlined: it is rare-path code and its callers' hot loops must stay inside the instruction cache)
__device__ __noinline__ float warp_compact(uint64_t* buf, int n, int kp, int lane) {
    constexpr int kPer = kCandCap / 32;
    uint32_t hi[kPer], lo[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        const int i = lane + 32 * e;
        const uint64_t c = i < n ? buf[i] : 0ull;
        hi[e] = static_cast<uint32_t>(c >> 32);
        lo[e] = static_cast<uint32_t>(c);
    }
    uint32_t key = 0;  // largest value with at least kp keys >= it  ==  the kp-th largest key
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
        const uint32_t trial = key | (1u << b);
        int c = 0;
#pragma unroll
        for (int e = 0; e < kPer; ++e) c += hi[e] >= trial ? 1 : 0;
        if (__reduce_add_sync(0xffffffffu, c) >= kp) key = trial;
    }
    int above = 0, ties = 0;
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        above += hi[e] > key ? 1 : 0;
        ties += hi[e] == key ? 1 : 0;
    }
    above = __reduce_add_sync(0xffffffffu, above);
    ties = __reduce_add_sync(0xffffffffu, ties);
    uint32_t low = 0;  // among equal keys the smaller row wins, i.e. the larger low word
    if (above + ties > kp) {
        const int need = kp - above;  // >= 1
#pragma unroll 1
        for (int b = 31; b >= 0; --b) {
            const uint32_t trial = low | (1u << b);
            int c = 0;
#pragma unroll
            for (int e = 0; e < kPer; ++e) c += (hi[e] == key && lo[e] >= trial) ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, c) >= need) low = trial;
        }
    }
    int keep = 0;
#pragma unroll
    for (int e = 0; e < kPer; ++e) keep += (hi[e] > key || (hi[e] == key && lo[e] >= low)) ? 1 : 0;
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    int w = incl - keep;
    __syncwarp();  // every lane has its registers loaded before the buffer is overwritten
#pragma unroll
    for (int e = 0; e < kPer; ++e)
        if (hi[e] > key || (hi[e] == key && lo[e] >= low)) buf[w++] = (static_cast<uint64_t>(hi[e]) << 32) | lo[e];
    __syncwarp();
    return ord2f(key);
}

template <bool HAS_IP, bool HAS_KL>
__global__ void __launch_bounds__(kScanThreads) simt_scan_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);  // [KC][LD]  query chunk, k-major
    float* Bs = As + kScanKC * kScanLd;              // [KC][LD]  corpus chunk, k-major
    float* Ps = Bs + kScanKC * kScanLd;              // [16][TQ]
    float* Ls = Ps + kObsPad * kScanTQ;              // [16][TC]
    float* Hs = Ls + kObsPad * kScanTC;              // [TQ]
    float* thr = Hs + kScanTQ;                       // [TQ]
    uint32_t* cnt_s = reinterpret_cast<uint32_t*>(thr + kScanTQ);  // [TQ]
    uint32_t* qid_s = cnt_s + kScanTQ;                             // [TQ]
    uint32_t* ubk_s = qid_s + kScanTQ;                             // [TQ] search-after bound: orderable key bits
    long long* ubr_s = reinterpret_cast<long long*>(ubk_s + kScanTQ);  // [TQ] ... and LOCAL row (may lie outside the shard)

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int part = blockIdx.y;
    const int64_t q0 = static_cast<int64_t>(blockIdx.x) * kScanTQ;
    const int64_t nq = a.nq_dev ? min(static_cast<int64_t>(*a.nq_dev > a.nq_skip ? *a.nq_dev - a.nq_skip : 0u), a.nq) : a.nq;
    if (q0 >= nq) return;
    const int64_t part_begin = static_cast<int64_t>(part) * a.rows_per_part;
    const int64_t part_end = min(a.n, part_begin + a.rows_per_part);

    if (tid < kScanTQ) {
        const int64_t qi = q0 + tid;
        const bool valid = qi < nq;
        const uint32_t qid = valid ? (a.qmap ? a.qmap[qi] : static_cast<uint32_t>(qi)) : 0xFFFFFFFFu;
        qid_s[tid] = qid;
        thr[tid] = -CUDART_INF_F;
        cnt_s[tid] = 0;
        Hs[tid] = (valid && HAS_KL) ? a.entropy[qid] : 0.0f;
        ubk_s[tid] = 0xFFFFFFFFu;   // no bound: every finite key orders below it
        ubr_s[tid] = -1;
        if (valid && a.after_idx && a.after_idx[qid] >= 0) {
            const float sc = a.after_scores[qid];
            ubk_s[tid] = f2ord(a.mode == RADAR_MODE_KL ? __fsub_rn(0.0f, sc) : sc);
            ubr_s[tid] = a.after_idx[qid] - a.idx_offset;
        }
#pragma unroll
        for (int j = 0; j < kObsPad; ++j)
            Ps[j * kScanTQ + tid] = (valid && HAS_KL) ? a.p16[static_cast<int64_t>(qid) * kObsPad + j] : 0.0f;
    }
    __syncthreads();

    const int ld_r = tid >> 2;         // tile row this thread stages (0..63)
    const int ld_k = (tid & 3) * 8;    // first of the 8 columns it stages
    const uint32_t ld_qid = qid_s[ld_r];

    for (int64_t row0 = part_begin; row0 < part_end; row0 += kScanTC) {
        float ip[4][4], x[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ip[i][j] = 0.0f;
                x[i][j] = 0.0f;
            }
        if (HAS_KL) {
            const int64_t row = row0 + ld_r;
            const int jq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < a.n) v = __ldg(reinterpret_cast<const float4*>(a.logq16 + row * kObsPad + jq));
            Ls[(jq + 0) * kScanTC + ld_r] = v.x;
            Ls[(jq + 1) * kScanTC + ld_r] = v.y;
            Ls[(jq + 2) * kScanTC + ld_r] = v.z;
            Ls[(jq + 3) * kScanTC + ld_r] = v.w;
        }
        if (HAS_IP) {
            for (int k0 = 0; k0 < a.d; k0 += kScanKC) {
                const int64_t row = row0 + ld_r;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kk = ld_k + 4 * h;
                    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                    if (k0 + kk < a.d) {
                        if (ld_qid != 0xFFFFFFFFu)
                            va = __ldg(reinterpret_cast<const float4*>(a.q_emb + static_cast<int64_t>(ld_qid) * a.d + k0 + kk));
                        if (row < a.n) vb = __ldg(reinterpret_cast<const float4*>(a.c_emb + row * a.d + k0 + kk));
                    }
                    As[(kk + 0) * kScanLd + ld_r] = va.x;
                    As[(kk + 1) * kScanLd + ld_r] = va.y;
                    As[(kk + 2) * kScanLd + ld_r] = va.z;
                    As[(kk + 3) * kScanLd + ld_r] = va.w;
                    Bs[(kk + 0) * kScanLd + ld_r] = vb.x;
                    Bs[(kk + 1) * kScanLd + ld_r] = vb.y;
                    Bs[(kk + 2) * kScanLd + ld_r] = vb.z;
                    Bs[(kk + 3) * kScanLd + ld_r] = vb.w;
                }
                __syncthreads();
#pragma unroll 8
                for (int kk = 0; kk < kScanKC; ++kk) {
                    const float4 av = *reinterpret_cast<const float4*>(As + kk * kScanLd + ty * 4);
                    const float4 bv = *reinterpret_cast<const float4*>(Bs + kk * kScanLd + tx * 4);
                    const float aa[4] = {av.x, av.y, av.z, av.w};
                    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) ip[i][j] = __fmaf_rn(aa[i], bb[j], ip[i][j]);
                }
                __syncthreads();
            }
        } else {
            __syncthreads();
        }
        if (HAS_KL) {
#pragma unroll
            for (int j = 0; j < kNumObs; ++j) {
                const float4 pv = *reinterpret_cast<const float4*>(Ps + j * kScanTQ + ty * 4);
                const float4 lv = *reinterpret_cast<const float4*>(Ls + j * kScanTC + tx * 4);
                const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
                const float ll[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) x[i][jj] = __fmaf_rn(pp[i], ll[jj], x[i][jj]);
            }
        }
        // threshold filter + append
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ql = ty * 4 + i;
            if (qid_s[ql] == 0xFFFFFFFFu) continue;
            const float t = thr[ql];
            const float h = Hs[ql];
            uint64_t* buf = a.cand + ((q0 + ql) * a.parts + part) * kCandCap;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t row = row0 + tx * 4 + j;
                const float key = canonical_key(a.mode, ip[i][j], x[i][j], h, a.alpha, a.oma);
                if (row < part_end && key >= t) {
                    // search-after: drop what ranks at or before the bound (better key, or equal key and id <= bound id)
                    const uint32_t o = f2ord(key), ub = ubk_s[ql];
                    if (o < ub || (o == ub && row > ubr_s[ql])) {
                        const uint32_t slot = atomicAdd(&cnt_s[ql], 1u);
                        buf[slot] = make_composite(key, static_cast<uint32_t>(row));
                    }
                }
            }
        }
        __syncthreads();
        for (int ql = warp; ql < kScanTQ; ql += kScanThreads / 32) {
            const int c = static_cast<int>(cnt_s[ql]);
            if (c > kCandSoft) {
                uint64_t* buf = a.cand + ((q0 + ql) * a.parts + part) * kCandCap;
                const float t = warp_compact(buf, c, a.kp, lane);
                if (lane == 0) {
                    thr[ql] = t;
                    cnt_s[ql] = a.kp;
                }
            }
        }
        __syncthreads();
    }
    if (tid < kScanTQ && q0 + tid < nq) a.cnt[(q0 + tid) * a.parts + part] = cnt_s[tid];
}

// ---------------------------------------------------------------------------------------------------
// select: per query, the best R composites over all slabs, sorted descending, into sel[qi][R] (0 = empty).
// bound[qi] = upper bound on the key of every candidate that was ever dropped for this query
// (-inf when nothing was dropped): max over slabs of the slab's final threshold, and of anything
// dropped here.  One warp per query.
// ---------------------------------------------------------------------------------------------------
constexpr int kSelWarps = 8;  // queries per CTA (one warp each)

__global__ void __launch_bounds__(kSelWarps * 32) select_kernel(const uint64_t* __restrict__ cand,
                                                                const uint32_t* __restrict__ cnt,
                                                                const float* __restrict__ thr_final, int64_t nq,
                                                                const uint32_t* __restrict__ nq_dev, int parts, int cap,
                                                                int R, uint64_t* __restrict__ sel,
                                                                float* __restrict__ bound, uint32_t nq_skip = 0,
                                                                const uint8_t* __restrict__ done = nullptr) {
    __shared__ uint64_t work[kSelWarps][kCandCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * kSelWarps + warp;
    if (nq_dev) nq = min(nq, static_cast<int64_t>(*nq_dev > nq_skip ? *nq_dev - nq_skip : 0u));
    if (qi >= nq) return;  // warp-uniform
    if (done && done[qi]) return;  // finished by the fused KL kernel (kl_finish_kernel)
    uint64_t* w = work[warp];
    float b = -CUDART_INF_F;
    if (thr_final) {
        for (int p = lane; p < parts; p += 32) b = fmaxf(b, thr_final[qi * parts + p]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    // The largest final threshold b is a lower bound of the k'-th best key of the whole corpus, and every case with a key >= b
    // still sits in some slab's buffer (a case is only ever dropped below a threshold <= b).  So entries below b -- most of what
    // the buffers hold: each keeps up to 192 admitted under older, looser thresholds -- are skipped on the way in, and the
    // register radix select below (issue-bound: ~600 instructions per call) runs once instead of once per 200 entries read.
    // Every compaction returns the R-th best key so far: a running threshold, so that the stream of entries thins out as it is
    // read (a short sweep never tightens the prepass threshold inside the filter: ~2 900 entries per query reach this kernel and
    // cost 14 compactions without the running threshold, 2 - 3 with it).
    uint32_t keep_from = (thr_final && b > -CUDART_INF_F) ? f2ord(b) : 0u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    int fill = 0;
    auto append = [&](uint64_t c, bool valid) {  // warp-wide: compacting append of the lanes' entries that can still matter
        const bool keep = valid && c != 0ull && static_cast<uint32_t>(c >> 32) >= keep_from;
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        const int n = __popc(m);
        if (n == 0) return;
        if (fill + n > kCandCap) {
            __syncwarp();
            const float t = warp_compact(w, fill, R, lane);  // everything it dropped is <= t; ties with t stay admissible
            b = fmaxf(b, t);
            keep_from = max(keep_from, f2ord(t));
            fill = R;
        }
        if (keep) w[fill + __popc(m & lt_mask)] = c;  // (entries of this chunk that fell below the new threshold are sorted out later)
        fill += n;
    };
    uint32_t cnt_l = 0;  // counts of 32 slabs at a time, one per lane: the per-slab loop below never waits for a count
    // Slabs are taken four at a time and a slab's whole buffer (up to kCandCap entries = 8 loads per lane) is requested before the
    // first entry is appended: 32 loads in flight per lane, so a query with many slabs (a single-query search sweeps 74 x 2 of
    // them) pays one load round trip per four slabs instead of several per slab.
    constexpr int kBatch = 4, kChunks = kCandCap / 32;
    for (int p0 = 0; p0 < parts; p0 += kBatch) {
        int c[kBatch];
        uint64_t v[kBatch][kChunks];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int p = p0 + u;
            if ((p & 31) == 0) cnt_l = p + lane < parts ? cnt[qi * parts + p + lane] : 0u;  // kBatch divides 32
            c[u] = p < parts ? static_cast<int>(__shfl_sync(0xffffffffu, cnt_l, p & 31)) : 0;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const uint64_t* src = cand + (qi * parts + p0 + u) * cap;
#pragma unroll
            for (int j = 0; j < kChunks; ++j) v[u][j] = 32 * j + lane < c[u] ? src[32 * j + lane] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
#pragma unroll
            for (int j = 0; j < kChunks; ++j)
                if (32 * j < c[u]) append(v[u][j], 32 * j + lane < c[u]);
            const uint64_t* src = cand + (qi * parts + p0 + u) * cap;
            for (int i0 = kCandCap; i0 < c[u]; i0 += 32)  // (a caller with larger buffers than kCandCap)
                append(i0 + lane < c[u] ? src[i0 + lane] : 0ull, i0 + lane < c[u]);
        }
    }
    __syncwarp();
    if (fill > R) {
        b = fmaxf(b, warp_compact(w, fill, R, lane));
        fill = R;
    }
    const int P = max(next_pow2(fill), 2);
    for (int i = fill + lane; i < P; i += 32) w[i] = 0ull;
    bitonic_sort_desc(w, P, lane, 32, [] { __syncwarp(); });
    for (int i = lane; i < R; i += 32) sel[qi * R + i] = i < fill ? w[i] : 0ull;
    if (bound && lane == 0) bound[qi] = b;
}

// ---------------------------------------------------------------------------------------------------
// rescore: canonical fp32 key of every selected candidate (one thread per candidate, fma chains in
// index order -- bit-identical to oracle/radar_oracle.c).
// ---------------------------------------------------------------------------------------------------
struct RescoreArgs {
    const float* q_emb;
    const float* p16;
    const float* entropy;
    const float* c_emb;
    const float* logq16;
    const uint32_t* qmap;
    int64_t nq;
    int d;
    int mode;
    float alpha, oma;
    int R;
    uint64_t* sel;
    const uint8_t* done;  // nullable: queries already finished by kl_finish_kernel
    // pruning (all nullable / 0): sel rows arrive sorted by FILTER key.  With f_k the k-th best filter key, a candidate with
    // f < f_k - 2 qerr cannot reach the canonical top-k (k candidates have canonical >= f_k - qerr, its own is < f_k - qerr),
    // so its 2 KB embedding row is never fetched: the entry is emptied and its filter key joins the bound of dropped
    // candidates, where the certificate of final_kernel sees it like any other dropped case.
    const float* qerr;
    float* bound;
    int k;
};

// One warp per query, lanes = candidates.  The embedding rows of the (up to) 32 candidates of a group are fetched
// COOPERATIVELY -- for every candidate row the warp issues one coalesced 512-byte load (LDG.128 per lane) into a padded
// shared-memory tile -- and then each lane walks its own row in index order, so the fma chain is the canonical one while
// HBM/L2 see full-line requests instead of 16-byte strided ones.
constexpr int kRsWarps = 4;                 // queries per CTA
constexpr int kRsChunk = 128;               // floats of a row staged per step
constexpr int kRsLd = kRsChunk + 4;         // padded row pitch: LDS.128 of 32 different rows is conflict free
constexpr size_t kRescoreSmemBytes = sizeof(float) * kRsWarps * (32 * kRsLd + kRsChunk);

__global__ void __launch_bounds__(kRsWarps * 32) rescore_kernel(const RescoreArgs a) {
    extern __shared__ __align__(16) float rs_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * kRsWarps + warp;
    if (qi >= a.nq) return;  // warp-uniform
    if (a.done && a.done[qi]) return;
    float* tile = rs_smem + warp * (32 * kRsLd + kRsChunk);
    float* qrow = tile + 32 * kRsLd;
    const int64_t qid = a.qmap ? a.qmap[qi] : qi;
    const bool has_ip = a.mode != RADAR_MODE_KL, has_kl = a.mode != RADAR_MODE_DPR;
    const float h = has_kl ? a.entropy[qid] : 0.0f;
    float prune_below = -CUDART_INF_F, pruned_max = -CUDART_INF_F;
    if (a.qerr && a.k >= 1 && a.k <= a.R) {
        const uint64_t ck = a.sel[qi * a.R + a.k - 1];
        if (ck != 0ull) prune_below = composite_key(ck) - 2.0f * a.qerr[qid];
    }
    for (int g0 = 0; g0 < a.R; g0 += 32) {
        const int ci = g0 + lane;
        const uint64_t c = ci < a.R ? a.sel[qi * a.R + ci] : 0ull;
        bool live = c != 0ull;
        if (live && composite_key(c) < prune_below) {
            pruned_max = fmaxf(pruned_max, composite_key(c));
            a.sel[qi * a.R + ci] = 0ull;
            live = false;
        }
        const uint32_t row = live ? composite_row(c) : 0u;
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        if (live_mask == 0u) continue;
        float ip = 0.0f, x = 0.0f;
        if (has_ip) {
            for (int t0 = 0; t0 < a.d; t0 += kRsChunk) {
                const int w = min(kRsChunk, a.d - t0);  // multiple of 4
                __syncwarp();
                if (4 * lane < w)
                    *reinterpret_cast<float4*>(qrow + 4 * lane) =
                        __ldg(reinterpret_cast<const float4*>(a.q_emb + qid * a.d + t0) + lane);
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const uint32_t rr = __shfl_sync(0xffffffffu, row, r);
                    if (((live_mask >> r) & 1u) && 4 * lane < w)
                        *reinterpret_cast<float4*>(tile + r * kRsLd + 4 * lane) =
                            __ldg(reinterpret_cast<const float4*>(a.c_emb + static_cast<int64_t>(rr) * a.d + t0) + lane);
                }
                __syncwarp();
                if (live) {
                    const float* mine = tile + lane * kRsLd;
                    for (int t = 0; t < w; t += 4) {
                        const float4 u = *reinterpret_cast<const float4*>(qrow + t);
                        const float4 v = *reinterpret_cast<const float4*>(mine + t);
                        ip = __fmaf_rn(u.x, v.x, ip);
                        ip = __fmaf_rn(u.y, v.y, ip);
                        ip = __fmaf_rn(u.z, v.z, ip);
                        ip = __fmaf_rn(u.w, v.w, ip);
                    }
                }
            }
        }
        if (live) {
            if (has_kl) {
                const float* pp = a.p16 + qid * kObsPad;
                const float4* ll = reinterpret_cast<const float4*>(a.logq16 + static_cast<int64_t>(row) * kObsPad);
                const float4 l0 = __ldg(ll), l1 = __ldg(ll + 1), l2 = __ldg(ll + 2), l3 = __ldg(ll + 3);
                const float lv[16] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w,
                                      l2.x, l2.y, l2.z, l2.w, l3.x, l3.y, l3.z, l3.w};
#pragma unroll
                for (int j = 0; j < kNumObs; ++j) x = __fmaf_rn(pp[j], lv[j], x);
            }
            a.sel[qi * a.R + ci] = make_composite(canonical_key(a.mode, ip, x, h, a.alpha, a.oma), row);
        }
    }
    if (a.bound) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pruned_max = fmaxf(pruned_max, __shfl_xor_sync(0xffffffffu, pruned_max, o));
        if (lane == 0 && pruned_max > -CUDART_INF_F) a.bound[qi] = fmaxf(a.bound[qi], pruned_max);
    }
}

// ---------------------------------------------------------------------------------------------------
// final: sort the R canonical composites of a query, write the top k in API form.
// Certificate (filter path, FP32 precision): every dropped candidate has filter key <= bound and
// |canonical - filter| <= qerr, so the result is provably the canonical top-k when
// bound + qerr < (k-th best canonical key).  Otherwise the query is appended to uncert_list.
// ---------------------------------------------------------------------------------------------------
struct FinalArgs {
    const uint64_t* sel;
    int R;
    int k;
    int mode;
    int sort;  // 0: sel rows are already in final order
    const uint32_t* qmap;
    const uint32_t* nq_dev;  // nullable: device-side number of queries (CTAs beyond it exit)
    uint32_t nq_skip;        // with nq_dev: count = *nq_dev - nq_skip (the first nq_skip listed queries belong to another launch)
    int64_t idx_offset;
    float* out_scores;
    int64_t* out_idx;
    uint64_t* out_packed;  // nullable: (key bits << 32) | (0xFFFFFFFF - global id), 0 = padding
    // certificate (all nullable)
    const float* bound;
    const float* qerr;
    uint32_t* uncert_count;
    uint32_t* uncert_list;
    const uint8_t* done;  // nullable: queries already finished by kl_finish_kernel
};

constexpr int kFinalThreads = 128;
constexpr int kFinalCap = 512;

__global__ void __launch_bounds__(kFinalThreads) final_kernel(const FinalArgs a) {
    __shared__ uint64_t s[kFinalCap];
    const int64_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    if (a.nq_dev && qi >= static_cast<int64_t>(*a.nq_dev > a.nq_skip ? *a.nq_dev - a.nq_skip : 0u)) return;
    if (a.done && a.done[qi]) return;
    const int P = max(next_pow2(a.R), 2);
    for (int i = tid; i < P; i += kFinalThreads) s[i] = i < a.R ? a.sel[qi * a.R + i] : 0ull;
    if (a.sort) bitonic_sort_desc(s, P, tid, kFinalThreads, [] { __syncthreads(); });
    else __syncthreads();
    const int64_t qid = a.qmap ? a.qmap[qi] : qi;
    for (int j = tid; j < a.k; j += kFinalThreads) {
        const uint64_t c = s[j];
        float sc;
        int64_t id;
        if (c != 0ull) {
            sc = api_score_from_key(a.mode, composite_key(c));
            id = static_cast<int64_t>(composite_row(c)) + a.idx_offset;
        } else {
            sc = a.mode == RADAR_MODE_KL ? CUDART_INF_F : -CUDART_INF_F;
            id = -1;
        }
        a.out_scores[qid * a.k + j] = sc;
        a.out_idx[qid * a.k + j] = id;
        if (a.out_packed) a.out_packed[qid * a.k + j] = c != 0ull ? packed_global(c, a.idx_offset) : 0ull;
    }
    if (a.bound && tid == 0) {
        const float b = a.bound[qi];
        bool ok = true;
        if (b > -CUDART_INF_F) {
            const uint64_t ck = s[a.k - 1];
            ok = (ck != 0ull) && (b + a.qerr[qid] < composite_key(ck));
        }
        if (!ok) {
            const uint32_t slot = atomicAdd(a.uncert_count, 1u);
            a.uncert_list[slot] = static_cast<uint32_t>(qid);
        }
    }
}

// R <= 32: one WARP per query (a CTA per query is mostly launch overhead when R is a few dozen entries): every lane
// holds one composite, its rank is the number of better ones (composites are distinct; empty slots rank by lane).
constexpr int kFinalWarps = 8;

__global__ void __launch_bounds__(kFinalWarps * 32) final_warp_kernel(const FinalArgs a, int64_t nq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * kFinalWarps + warp;
    if (a.nq_dev) nq = min(nq, static_cast<int64_t>(*a.nq_dev > a.nq_skip ? *a.nq_dev - a.nq_skip : 0u));
    if (qi >= nq) return;  // warp-uniform
    if (a.done && a.done[qi]) return;
    const uint64_t c = lane < a.R ? a.sel[qi * a.R + lane] : 0ull;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint64_t o = __shfl_sync(0xffffffffu, c, j);
        rank += (o > c || (o == c && j < lane)) ? 1 : 0;
    }
    const int64_t qid = a.qmap ? a.qmap[qi] : qi;
    if (rank < a.k) {
        float sc;
        int64_t id;
        if (c != 0ull) {
            sc = api_score_from_key(a.mode, composite_key(c));
            id = static_cast<int64_t>(composite_row(c)) + a.idx_offset;
        } else {
            sc = a.mode == RADAR_MODE_KL ? CUDART_INF_F : -CUDART_INF_F;
            id = -1;
        }
        a.out_scores[qid * a.k + rank] = sc;
        a.out_idx[qid * a.k + rank] = id;
        if (a.out_packed) a.out_packed[qid * a.k + rank] = c != 0ull ? packed_global(c, a.idx_offset) : 0ull;
    }
    if (a.bound && rank == a.k - 1) {  // exactly one lane holds the k-th best entry
        const float b = a.bound[qi];
        bool ok = true;
        if (b > -CUDART_INF_F) ok = (c != 0ull) && (b + a.qerr[qid] < composite_key(c));
        if (!ok) {
            const uint32_t slot = atomicAdd(a.uncert_count, 1u);
            a.uncert_list[slot] = static_cast<uint32_t>(qid);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// KL filter path, short candidate lists: select + rescore + final in ONE kernel, one warp per query.  After a prepass the
// buffers of a query hold a few dozen candidates in total; when they are at most 64 (two per lane) the warp gathers them
// all, computes their canonical keys (every candidate, not only the best R by filter key: nothing is dropped here, so the
// bound is just the largest final threshold), ranks them by counting, writes the top k and checks the certificate.
// Queries with more candidates are left to the three generic kernels (done[qi] = 0).
// ---------------------------------------------------------------------------------------------------
struct KlFinishArgs {
    const uint64_t* cand;     // [q][parts][kCandCap]
    const uint32_t* cnt;      // [q][parts]
    const float* thr_final;   // [q][parts]
    const float* p16;
    const float* entropy;
    const float* logq16;
    const float* qerr;        // nullable: no certificate (filter-only precision)
    int64_t nq;
    int parts, k;
    int64_t idx_offset;
    float* out_scores;
    int64_t* out_idx;
    uint64_t* out_packed;     // nullable
    uint32_t* uncert_count;
    uint32_t* uncert_list;
    uint8_t* done;            // [q] out
};

constexpr int kFinishWarps = 8;

__global__ void __launch_bounds__(kFinishWarps * 32) kl_finish_kernel(const KlFinishArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * kFinishWarps + warp;
    if (qi >= a.nq) return;  // warp-uniform
    // counts and final thresholds of up to 32 buffers per round (parts is a few units)
    int total = 0;
    float b = -CUDART_INF_F;
    for (int p0 = 0; p0 < a.parts; p0 += 32) {
        const int p = p0 + lane;
        const int c = p < a.parts ? static_cast<int>(a.cnt[qi * a.parts + p]) : 0;
        if (p < a.parts) b = fmaxf(b, a.thr_final[qi * a.parts + p]);
        total += __reduce_add_sync(0xffffffffu, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (total > 64 || a.parts > 32) {  // warp-uniform: the generic kernels take this query
        if (lane == 0) a.done[qi] = 0;
        return;
    }
    const int my_cnt = lane < a.parts ? static_cast<int>(a.cnt[qi * a.parts + lane]) : 0;
    // gather: entry j of the concatenated buffers -> (buffer, offset)
    uint64_t c[2] = {0ull, 0ull};
    {
        int base = 0;
        for (int p = 0; p < a.parts; ++p) {
            const int cp = __shfl_sync(0xffffffffu, my_cnt, p);
            const uint64_t* src = a.cand + (qi * a.parts + p) * kCandCap;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = lane + 32 * e;
                if (j >= base && j < base + cp) c[e] = src[j - base];
            }
            base += cp;
        }
    }
    // canonical keys (the fma chain of oracle/radar_oracle.c: index order, fp32)
    const float h = a.entropy[qi];
    const float pv = lane < kObsPad ? a.p16[qi * kObsPad + lane] : 0.0f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const uint32_t row = composite_row(c[e]);
        const bool live = c[e] != 0ull;
        float lv[16];
        if (live) {
            const float4* ll = reinterpret_cast<const float4*>(a.logq16 + static_cast<int64_t>(row) * kObsPad);
            const float4 l0 = __ldg(ll), l1 = __ldg(ll + 1), l2 = __ldg(ll + 2), l3 = __ldg(ll + 3);
            lv[0] = l0.x; lv[1] = l0.y; lv[2] = l0.z; lv[3] = l0.w; lv[4] = l1.x; lv[5] = l1.y; lv[6] = l1.z; lv[7] = l1.w;
            lv[8] = l2.x; lv[9] = l2.y; lv[10] = l2.z; lv[11] = l2.w; lv[12] = l3.x; lv[13] = l3.y; lv[14] = l3.z; lv[15] = l3.w;
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) lv[j] = 0.0f;
        }
        float x = 0.0f;
#pragma unroll
        for (int j = 0; j < kNumObs; ++j) x = __fmaf_rn(__shfl_sync(0xffffffffu, pv, j), lv[j], x);
        if (live) c[e] = make_composite(__fsub_rn(x, h), row);
    }
    // rank by counting (composites are distinct; empty slots rank behind everything, by position)
    int rank[2] = {0, 0};
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        for (int j = 0; j < 32; ++j) {
            const uint64_t o = __shfl_sync(0xffffffffu, c[f], j);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int pos_o = j + 32 * f, pos_me = lane + 32 * e;
                rank[e] += (o > c[e] || (o == c[e] && pos_o < pos_me)) ? 1 : 0;
            }
        }
    }
    bool cert_ok = true;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        if (rank[e] < a.k) {
            float sc;
            int64_t id;
            if (c[e] != 0ull) {
                sc = api_score_from_key(RADAR_MODE_KL, composite_key(c[e]));
                id = static_cast<int64_t>(composite_row(c[e])) + a.idx_offset;
            } else {
                sc = CUDART_INF_F;
                id = -1;
            }
            a.out_scores[qi * a.k + rank[e]] = sc;
            a.out_idx[qi * a.k + rank[e]] = id;
            if (a.out_packed) a.out_packed[qi * a.k + rank[e]] = c[e] != 0ull ? packed_global(c[e], a.idx_offset) : 0ull;
        }
        if (a.qerr && rank[e] == a.k - 1 && b > -CUDART_INF_F)  // exactly one slot holds the k-th best entry
            cert_ok = (c[e] != 0ull) && (b + a.qerr[qi] < composite_key(c[e]));
    }
    if (!cert_ok) {
        const uint32_t slot = atomicAdd(a.uncert_count, 1u);
        a.uncert_list[slot] = static_cast<uint32_t>(qi);
    }
    if (lane == 0) a.done[qi] = 1;
}

static inline cudaError_t launch_final(const FinalArgs& a, int64_t nq, cudaStream_t st) {
    if (a.R <= 32 && a.k <= 32)
        final_warp_kernel<<<static_cast<unsigned>((nq + kFinalWarps - 1) / kFinalWarps), kFinalWarps * 32, 0, st>>>(a, nq);
    else
        final_kernel<<<static_cast<unsigned>(nq), kFinalThreads, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// merge of per-shard top-k lists ([parts][q][k_in], the ncclAllGather layout).  One CTA per query.
// ---------------------------------------------------------------------------------------------------
constexpr int kMergeCap = 2048;

__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx,
                                                    int64_t nq, int parts, int k_in, int k_out, int ascending,
                                                    float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    __shared__ uint64_t s[kMergeCap];
    const int64_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = parts * k_in;
    const int P = max(next_pow2(total), 2);
    for (int i = tid; i < P; i += 256) {
        uint64_t c = 0ull;
        if (i < total) {
            const int p = i / k_in, j = i - p * k_in;
            const int64_t o = (static_cast<int64_t>(p) * nq + qi) * k_in + j;
            const int64_t id = idx[o];
            if (id >= 0) {
                const float sc = scores[o];
                c = make_composite(ascending ? __fsub_rn(0.0f, sc) : sc, static_cast<uint32_t>(id));
            }
        }
        s[i] = c;
    }
    bitonic_sort_desc(s, P, tid, 256, [] { __syncthreads(); });
    for (int j = tid; j < k_out; j += 256) {
        const uint64_t c = s[j];
        if (c != 0ull) {
            const float key = composite_key(c);
            out_scores[qi * k_out + j] = ascending ? __fsub_rn(0.0f, key) : key;
            out_idx[qi * k_out + j] = static_cast<int64_t>(composite_row(c));
        } else {
            out_scores[qi * k_out + j] = ascending ? CUDART_INF_F : -CUDART_INF_F;
            out_idx[qi * k_out + j] = -1;
        }
    }
}

// the same merge on packed words (radar_search's out_packed; one 8-byte all-gather instead of scores + ids)
__global__ void __launch_bounds__(256) merge_packed_kernel(const uint64_t* __restrict__ cand, int64_t nq, int parts,
                                                           int k_in, int k_out, int mode, float* __restrict__ out_scores,
                                                           int64_t* __restrict__ out_idx) {
    __shared__ uint64_t s[kMergeCap];
    const int64_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = parts * k_in;
    const int P = max(next_pow2(total), 2);
    for (int i = tid; i < P; i += 256) {
        uint64_t c = 0ull;
        if (i < total) {
            const int p = i / k_in, j = i - p * k_in;
            c = cand[(static_cast<int64_t>(p) * nq + qi) * k_in + j];
        }
        s[i] = c;
    }
    bitonic_sort_desc(s, P, tid, 256, [] { __syncthreads(); });
    for (int j = tid; j < k_out; j += 256) {
        const uint64_t c = s[j];
        if (c != 0ull) {
            out_scores[qi * k_out + j] = api_score_from_key(mode, composite_key(c));
            out_idx[qi * k_out + j] = static_cast<int64_t>(composite_row(c));
        } else {
            out_scores[qi * k_out + j] = mode == RADAR_MODE_KL ? CUDART_INF_F : -CUDART_INF_F;
            out_idx[qi * k_out + j] = -1;
        }
    }
}

}  // namespace radar
