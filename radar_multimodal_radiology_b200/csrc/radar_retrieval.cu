// radar_retrieval.cu -- C-ABI entry points of libradar_retrieval.so (see include/radar_retrieval.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "prep_kernels.cuh"
#include "scan_kernels.cuh"
#include "tc_filter.cuh"
#include "kl_filter.cuh"
#include "kl_stream.cuh"
#include "proj_gemm.cuh"

namespace radar {

static thread_local char g_err[512] = "";
// optional CUDA events recorded around the main scan / filter kernel of the next radar_search calls
static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
static thread_local int g_prof_dev = -1;  // device the event pair was created on (events belong to one device)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct DeviceInfo {
    int sms = 0, major = 0, minor = 0;
    bool ok = false;
};

static int get_device_info(DeviceInfo* out) {
    static thread_local DeviceInfo cache[64];
    int dev = 0;
    RADAR_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) {
        set_error("device ordinal %d out of range", dev);
        return RADAR_E_CUDA;
    }
    if (!cache[dev].ok) {
        RADAR_CUDA_CHECK(cudaDeviceGetAttribute(&cache[dev].sms, cudaDevAttrMultiProcessorCount, dev));
        RADAR_CUDA_CHECK(cudaDeviceGetAttribute(&cache[dev].major, cudaDevAttrComputeCapabilityMajor, dev));
        RADAR_CUDA_CHECK(cudaDeviceGetAttribute(&cache[dev].minor, cudaDevAttrComputeCapabilityMinor, dev));
        cache[dev].ok = true;
    }
    *out = cache[dev];
    return RADAR_OK;
}

// ---- search planning -------------------------------------------------------------------------------
struct Plan {
    int algo;
    int kp;            // candidates a compaction keeps (k on the exact path, k' on the filter path)
    int R;             // candidates selected per query for re-scoring (= kp)
    int64_t q_tiles;   // query tiles of the main pass
    int tile_q;        // rows per query tile (64 scan / 256 filter)
    int units;         // CTA pairs the filter launches
    int parts;
    int buf_parts;     // candidate buffers per query row: slabs x epilogue warp sets
    int64_t rows_per_part;
    // workspace offsets (bytes)
    size_t off_cand, off_cnt, off_thr, off_gthr, off_sel, off_bound, off_qerr, off_apack, off_ucount, off_ulist;
    size_t off_groupmax;  // KL threshold prepass
    int groups, group_tiles, tile_stride, groups_per_slab;
    size_t off_done;   // KL filter path: per-query "finished by kl_finish_kernel" flags
    int klf;           // KL mode: the dedicated many-queries kernel (kl_filter.cuh) runs the main pass
    int kl_fmt;        // its filter arithmetic (klf::kFmt*); also the KL stream path's table choice
    size_t off_fb_cand, off_fb_cnt, off_fb_sel;  // exact re-run of uncertified queries
    // KL stream path
    int n_pad, pool_cap, sample_tiles;
    int64_t tiles;
    size_t off_ks_zero, ks_zero_bytes, off_ks_pool, off_ks_best, off_ks_tilemax;
    int fb_parts;
    int64_t fb_rows_per_part;
    // exact re-run, first wave: the first kRerunHead uncertified queries get their own launch with MANY corpus slabs (a
    // handful of failed certificates must not be scanned by one CTA each: that is 4.5 ms for ONE query over 188 k rows and
    // 0.2 s over 10 M); the rest -- if there ever are more -- go through the all-queries plan above
    int fa_parts;
    int64_t fa_rows_per_part;
    size_t off_fa_cand, off_fa_cnt, off_fa_sel;
    size_t total;
};

static int auto_overfetch(int k) {
    int kp = ((2 * k + 22) + 15) / 16 * 16;
    if (kp < 32) kp = 32;
    if (kp > kCandSoft) kp = kCandSoft;
    return kp;
}

// KL keys come out of the filter almost exactly (bf16 hi/lo split of both operands: ~1e-4 relative), so a handful of
// spare candidates is enough; fewer candidates = tighter thresholds = fewer survivors through the filter's rare path
static int auto_overfetch_kl(int k) {
    int kp = (k + 6 + 7) / 8 * 8;
    if (kp < 16) kp = 16;
    if (kp > kCandSoft) kp = kCandSoft;
    return kp;
}

// KL keys of the fp16 filters carry a 4-7x larger error bound than the bf16 hi/lo split (kl_filter.cuh): the certificate
// "k'-th filter key + qerr < k-th canonical key" needs a wider gap between rank k and rank k', i.e. more candidates
static int auto_overfetch_kl_fmt(int k, int fmt) {
    if (fmt == klf::kFmtBf16x3) return auto_overfetch_kl(k);
    int kp = fmt == klf::kFmtF16x2 ? (3 * k / 2 + 8 + 7) / 8 * 8 : (2 * k + 12 + 7) / 8 * 8;
    if (kp < 24) kp = 24;
    if (kp > kCandSoft) kp = kCandSoft;
    return kp;
}

static bool tc_supported(const radar_corpus_t* c, int mode, const DeviceInfo& di) {
    if (di.major != 10) return false;
    if (mode != RADAR_MODE_KL) {
        if (!c->emb_bf16 || c->d % 64 != 0 || c->d > 512 || c->d <= 0) return false;
    }
    if (mode == RADAR_MODE_HYBRID && !c->klpack) return false;
    if (mode == RADAR_MODE_KL && !c->klpack && !c->kl16) return false;
    return c->n >= 1;
}

// filter arithmetic of the KL-only tensor-core paths: what the caller asked for, else what the corpus carries
// AUTO: the many-queries filter is select-bound, not tensor- or bandwidth-bound, so the arithmetic with the tightest error
// bound (bf16 hi/lo x 3: fewest candidates per query) is the fastest there; the HBM-bound stream path prefers the 32-byte
// fp16 rows (prefer_f16)
static int resolve_kl_fmt(const radar_corpus_t* c, const radar_search_params_t* p, bool prefer_f16, int* fmt) {
    switch (p->kl_variant) {
        case RADAR_KL_AUTO:
            if (prefer_f16) *fmt = c->kl16 ? klf::kFmtF16x2 : klf::kFmtBf16x3;
            else *fmt = c->klpack ? klf::kFmtBf16x3 : klf::kFmtF16x2;
            break;
        case RADAR_KL_BF16X3: *fmt = klf::kFmtBf16x3; break;
        case RADAR_KL_F16X1: *fmt = klf::kFmtF16x1; break;
        case RADAR_KL_F16X2: *fmt = klf::kFmtF16x2; break;
        default: set_error("bad kl_variant %d", p->kl_variant); return RADAR_E_ARG;
    }
    if (*fmt == klf::kFmtBf16x3 ? !c->klpack : !c->kl16) {
        set_error("kl_variant %d needs corpus.%s", p->kl_variant, *fmt == klf::kFmtBf16x3 ? "klpack" : "kl16");
        return RADAR_E_ARG;
    }
    return RADAR_OK;
}

static void plan_parts(int64_t q_tiles, int64_t n, int tile_rows, int sms, int waves, int* parts,
                       int64_t* rows_per_part) {
    const int64_t c_tiles = ceil_div64(n, tile_rows);
    int64_t p = 1;
    if (q_tiles < static_cast<int64_t>(sms) * waves) p = (static_cast<int64_t>(sms) * waves) / q_tiles;
    if (p > c_tiles) p = c_tiles;
    if (p < 1) p = 1;
    if (p > 1024) p = 1024;
    const int64_t tiles_per_part = ceil_div64(c_tiles, p);
    p = ceil_div64(c_tiles, tiles_per_part);
    *parts = static_cast<int>(p);
    *rows_per_part = tiles_per_part * tile_rows;
}

// Filter path: slabs are chosen so that (query tiles x slabs) fills the work units (CTAs or CTA pairs) evenly --
// thresholds are shared across the slabs of a query (FilterArgs::gthr), so extra slabs cost little.
static void plan_parts_filter(int64_t q_tiles, int64_t n, int tile_rows, int units, int* parts, int64_t* rows_per_part) {
    const int64_t c_tiles = ceil_div64(n, tile_rows);
    int64_t max_p = c_tiles / 64;  // keep >= 64 corpus tiles per slab
    if (max_p < 1) max_p = 1;
    if (max_p > 64) max_p = 64;
    int64_t best_p = 1;
    double best_waste = 1e30;
    for (int64_t p = 1; p <= max_p; ++p) {
        const int64_t items = q_tiles * p;
        const int64_t rounds = ceil_div64(items, units);
        const double waste = static_cast<double>(rounds * units) / static_cast<double>(items) - 1.0;
        // prefer fewer slabs (fewer candidate buffers per query, longer work items) unless the gain in occupancy is > 1.5 %;
        // while a slab keeps >= 512 corpus tiles a work item is long enough for 0.5 % to pay: 16 384 queries over 10 M rows run
        // as 64 x 15 = 960 items on 74 CTA pairs (13 rounds, 0.2 % idle) instead of 64 x 8 = 512 (7 rounds, 1.2 % idle) --
        // measured 118.6 -> 117.6 ms, a 1.25 M-row shard 15.4 -> 15.1 ms
        const double min_gain = c_tiles / p >= 512 ? 0.005 : 0.015;
        if (waste < best_waste - min_gain) {
            best_waste = waste;
            best_p = p;
        }
    }
    // tiny corpora / many idle units: fall back to spreading over the units
    if (q_tiles * best_p < units && c_tiles > best_p) {
        int64_t p = units / q_tiles;
        if (p > c_tiles) p = c_tiles;
        if (p > 1024) p = 1024;
        if (p > best_p) best_p = p;
    }
    const int64_t tiles_per_part = ceil_div64(c_tiles, best_p);
    best_p = ceil_div64(c_tiles, tiles_per_part);
    *parts = static_cast<int>(best_p);
    *rows_per_part = tiles_per_part * tile_rows;
}

constexpr int64_t kRerunHead = 1024;

// workspace of the exact re-run chain (both waves); `carve` is make_plan's allocator
template <typename Carve>
static void plan_rerun(Plan* pl, const radar_corpus_t* c, int64_t q, int k, int sms, Carve carve) {
    const int64_t fa_q = q < kRerunHead ? q : kRerunHead;
    const int64_t fa_tiles = ceil_div64(fa_q, kScanTQ);
    plan_parts(fa_tiles, c->n, kScanTC, sms, 8, &pl->fa_parts, &pl->fa_rows_per_part);
    pl->off_fa_cand = carve(sizeof(uint64_t) * fa_tiles * kScanTQ * pl->fa_parts * kCandCap);
    pl->off_fa_cnt = carve(sizeof(uint32_t) * fa_tiles * kScanTQ * pl->fa_parts);
    pl->off_fa_sel = carve(sizeof(uint64_t) * fa_q * k);
    const int64_t fb_q = q > kRerunHead ? q - kRerunHead : 0;
    if (fb_q > 0) {
        const int64_t f_tiles = ceil_div64(fb_q, kScanTQ);
        plan_parts(f_tiles, c->n, kScanTC, sms, 2, &pl->fb_parts, &pl->fb_rows_per_part);
        pl->off_fb_cand = carve(sizeof(uint64_t) * f_tiles * kScanTQ * pl->fb_parts * kCandCap);
        pl->off_fb_cnt = carve(sizeof(uint32_t) * f_tiles * kScanTQ * pl->fb_parts);
        pl->off_fb_sel = carve(sizeof(uint64_t) * fb_q * k);
    }
}

static bool kl_stream_supported(const radar_corpus_t* c, int64_t q, int mode, const DeviceInfo& di) {
    return di.major == 10 && mode == RADAR_MODE_KL && (c->klpack || c->kl16) && c->logq16 && q >= 1 && q <= kls::kMaxN &&
           c->n >= kls::kMinRows;
}

static int make_plan(const radar_corpus_t* c, int64_t q, const radar_search_params_t* p, const DeviceInfo& di,
                     Plan* pl, bool search_after = false, bool legacy_kl = false) {
    memset(pl, 0, sizeof *pl);
    const int sms = p->num_sms > 0 ? p->num_sms : di.sms;
    int algo = p->algo;
    if (search_after) {  // paging through a ranking is served by the exact scan only
        if (algo != RADAR_ALGO_AUTO && algo != RADAR_ALGO_SIMT_EXACT) {
            set_error("queries.after_scores/after_idx need RADAR_ALGO_AUTO or RADAR_ALGO_SIMT_EXACT");
            return RADAR_E_ARG;
        }
        algo = RADAR_ALGO_SIMT_EXACT;
    }
    if (algo == RADAR_ALGO_AUTO) {
        if (kl_stream_supported(c, q, p->mode, di) && p->num_sms == 0) algo = RADAR_ALGO_KL_STREAM;
        else algo = tc_supported(c, p->mode, di) ? RADAR_ALGO_TC_FILTER : RADAR_ALGO_SIMT_EXACT;
    }
    if (algo == RADAR_ALGO_KL_STREAM && !kl_stream_supported(c, q, p->mode, di)) {
        set_error("RADAR_ALGO_KL_STREAM needs an sm_100 device, KL mode, klpack, 1..%d queries and >= %d cases", kls::kMaxN,
                  kls::kMinRows);
        return di.major != 10 ? RADAR_E_ARCH : RADAR_E_ARG;
    }
    if (algo == RADAR_ALGO_TC_FILTER && !tc_supported(c, p->mode, di)) {
        set_error("RADAR_ALGO_TC_FILTER needs an sm_100 device, emb_bf16 (d %% 64 == 0, d <= 512) and/or klpack");
        return di.major != 10 ? RADAR_E_ARCH : RADAR_E_ARG;
    }
    pl->algo = algo;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    if (algo == RADAR_ALGO_KL_STREAM) {
        // AUTO: certified results want the tightest error bound (fewest exact re-runs, each a 10M-row scan): bf16 hi/lo x3;
        // filter-only precision streams the 32-byte fp16 rows (half the HBM bytes, same speed: the path is not HBM-bound
        // any more once the rows are this small)
        int rc = resolve_kl_fmt(c, p, p->precision == RADAR_PREC_BF16, &pl->kl_fmt);
        if (rc) return rc;
        pl->kp = p->overfetch > 0 ? p->overfetch : auto_overfetch_kl_fmt(p->k, pl->kl_fmt);
        if (pl->kp < p->k) pl->kp = p->k;
        if (pl->kp > kCandSoft) pl->kp = kCandSoft;
        pl->R = pl->kp;
        int n_pad = 32;
        while (n_pad < q) n_pad <<= 1;
        pl->n_pad = n_pad;
        pl->pool_cap = kls::pool_cap_for(n_pad);
        pl->tiles = ceil_div64(c->n, kls::kTileRows);
        pl->sample_tiles = kls::sample_tiles_for(n_pad, pl->tiles);
        pl->units = sms / 2 < 1 ? 1 : sms / 2;
        pl->tile_q = n_pad;
        pl->q_tiles = 1;
        pl->parts = 1;
        const int64_t qp = n_pad;
        pl->ks_zero_bytes = align_up(sizeof(uint32_t) * (5 * qp + 8), 256) + sizeof(uint64_t) * qp * pl->pool_cap +
                            sizeof(uint32_t) * qp * pl->sample_tiles;
        pl->off_ks_zero = carve(pl->ks_zero_bytes);  // [gcnt | lock | processed | best_n | gthr | - | ucount], the pools, the sample maxima
        pl->off_ks_pool = pl->off_ks_zero + align_up(sizeof(uint32_t) * (5 * qp + 8), 256);
        pl->off_ks_tilemax = pl->off_ks_pool + sizeof(uint64_t) * qp * pl->pool_cap;
        pl->off_ks_best = carve(sizeof(uint64_t) * qp * kCandCap);
        pl->off_qerr = carve(sizeof(float) * qp);
        pl->off_ucount = pl->off_ks_zero + sizeof(uint32_t) * (5 * qp + 4);
        pl->off_ulist = carve(sizeof(uint32_t) * qp);
        pl->off_apack = carve(tc::apack_bytes(qp, RADAR_MODE_KL, c->d));
        plan_rerun(pl, c, q, p->k, sms, carve);
        pl->total = off;
        return RADAR_OK;
    }
    if (algo == RADAR_ALGO_SIMT_EXACT) {
        pl->kp = p->k;
        pl->R = p->k;
        pl->tile_q = kScanTQ;
        pl->q_tiles = ceil_div64(q, kScanTQ);
        plan_parts(pl->q_tiles, c->n, kScanTC, sms, 2, &pl->parts, &pl->rows_per_part);
    } else {
        if (p->mode == RADAR_MODE_KL && !legacy_kl) {
            int rc = resolve_kl_fmt(c, p, false, &pl->kl_fmt);
            if (rc) return rc;
            pl->klf = 1;
            pl->kp = p->overfetch > 0 ? p->overfetch : auto_overfetch_kl_fmt(p->k, pl->kl_fmt);
        } else if (p->mode == RADAR_MODE_KL) {  // debug key dump: the KL mode of the general filter (bf16 hi/lo, klpack)
            if (!c->klpack) {
                set_error("the KL key dump needs corpus.klpack");
                return RADAR_E_ARG;
            }
            pl->kp = p->overfetch > 0 ? p->overfetch : auto_overfetch_kl(p->k);
        } else {
            pl->kp = p->overfetch > 0 ? p->overfetch : auto_overfetch(p->k);
        }
        if (pl->kp < p->k) pl->kp = p->k;
        if (pl->kp > kCandSoft) pl->kp = kCandSoft;
        pl->R = pl->kp;
        pl->tile_q = pl->klf ? klf::kTileQK : tc::kTileQ;
        pl->q_tiles = ceil_div64(q, pl->tile_q);
        int units = (pl->klf && !klf::kPair) ? sms : sms / 2;  // one CTA pair (tcgen05 cta_group::2) per two SMs, or single CTAs
        if (units < 1) units = 1;
        if (units > tc::kMaxUnits) units = tc::kMaxUnits;
        pl->units = units;
        plan_parts_filter(pl->q_tiles, c->n, pl->klf ? klf::kBlockN : tc::block_n_for_mode(p->mode), units, &pl->parts,
                          &pl->rows_per_part);
    }
    const int64_t q_pad = pl->q_tiles * pl->tile_q;
    pl->buf_parts = pl->parts * (algo == RADAR_ALGO_TC_FILTER ? (pl->klf ? klf::kE : tc::epi_sets_for_mode(p->mode)) : 1);
    pl->off_cand = carve(sizeof(uint64_t) * q_pad * pl->buf_parts * kCandCap);
    pl->off_cnt = carve(sizeof(uint32_t) * q_pad * pl->buf_parts);
    pl->off_thr = carve(sizeof(float) * q_pad * pl->buf_parts);
    pl->off_gthr = carve(tc::gthr_region_bytes(q_pad));
    pl->off_sel = carve(sizeof(uint64_t) * q * pl->R);
    pl->off_bound = carve(sizeof(float) * q);
    pl->off_qerr = carve(sizeof(float) * q_pad);
    pl->off_ucount = carve(sizeof(uint32_t) * 4);
    pl->off_ulist = carve(sizeof(uint32_t) * q);
    if (algo == RADAR_ALGO_TC_FILTER) {
        pl->off_apack = carve(tc::apack_bytes(q_pad, p->mode, c->d));
        if (pl->klf) pl->off_done = carve(static_cast<size_t>(q));
        // KL with many queries over a small corpus: the cold-start survivors (~k' ln n per query) cost more than a second
        // (partial) sweep of the cheap KL contraction, so a prepass over every tile_stride-th tile collects group maxima
        // and the real pass starts from near-exact thresholds (kl_filter.cuh).  About 256 groups per query (measured: 512 groups
        // tighten the thresholds a little and cost more in the threshold kernel than they save: 3.66 vs 3.60 ms per step).
        if (pl->klf && c->n <= (2ll << 20) && pl->q_tiles * pl->tile_q >= 2048 && pl->kp <= klf::kMaxKpPrepass) {
#ifndef RADAR_KLF_PREPASS_STRIDE
#define RADAR_KLF_PREPASS_STRIDE 3
#endif
            const int64_t stride = RADAR_KLF_PREPASS_STRIDE;
            const int64_t slab_tiles = ceil_div64(ceil_div64(pl->rows_per_part, klf::kBlockN), stride);
#ifndef RADAR_KLF_GROUPS
#define RADAR_KLF_GROUPS 256
#endif
            int64_t tgs = RADAR_KLF_GROUPS / (static_cast<int64_t>(pl->parts) * klf::kE);
            if (tgs > slab_tiles) tgs = slab_tiles;
            if (tgs < 1) tgs = 1;
            const int64_t gt = ceil_div64(slab_tiles, tgs);
            tgs = ceil_div64(slab_tiles, gt);
            const int64_t groups_per_slab = tgs * klf::kE;
            const int64_t groups = groups_per_slab * pl->parts;
            if (groups >= 4 * pl->kp && groups <= klf::kMaxGroupsK) {
                pl->groups = static_cast<int>(groups);
                pl->group_tiles = static_cast<int>(gt);
                pl->tile_stride = static_cast<int>(stride);
                pl->groups_per_slab = static_cast<int>(groups_per_slab);
                pl->off_groupmax = carve(sizeof(uint32_t) * q_pad * groups);
            }
        }
        // DPR / hybrid: sampled threshold prepass of the general filter (tc_filter.cuh: launch_filter) -- about 80 k sampled
        // rows per query (one tile in 16 on a 1.25 M-row shard, one in 128 on 10 M rows), ~256 groups per query.
        // On short sweeps the cold start of the thresholds is NOT hidden behind the MMAs (1.25 M-row shards of an 8-GPU
        // run: 16.2 -> 15.3 ms per step; config 3: 21.0 -> 19.4 ms); on a 10 M-row sweep it is, and the 0.8 % of extra MMAs
        // pays for itself (118.7 -> 117.9 ms).  RADAR_TC_PREPASS_MAX_ROWS can switch it off above a shard size.
#ifndef RADAR_TC_PREPASS_MAX_ROWS
#define RADAR_TC_PREPASS_MAX_ROWS (1ll << 40)
#endif
        if (!pl->klf && c->n <= RADAR_TC_PREPASS_MAX_ROWS && c->n >= (1ll << 17) && pl->kp <= 128) {
#ifndef RADAR_TC_PREPASS_STRIDE
#define RADAR_TC_PREPASS_STRIDE 16
#endif
#ifndef RADAR_TC_PREPASS_ROWS
#define RADAR_TC_PREPASS_ROWS 78125
#endif
            int64_t stride = c->n / RADAR_TC_PREPASS_ROWS;
            if (stride < RADAR_TC_PREPASS_STRIDE) stride = RADAR_TC_PREPASS_STRIDE;
            if (stride > 256) stride = 256;
            const int bn = tc::block_n_for_mode(p->mode);
            const int64_t slab_tiles = ceil_div64(ceil_div64(pl->rows_per_part, bn), stride);
            // ~256 groups per query, 4 k' of them when many candidates are kept (k' = 96 for top-32), at most what the threshold kernel holds
            int64_t want = 4ll * pl->kp > 256 ? 4ll * pl->kp : 256;
            if (want > 32 * tc::kMaxGroups32) want = 32 * tc::kMaxGroups32;
            int64_t tgs = want / pl->parts;
            if (tgs > slab_tiles) tgs = slab_tiles;
            if (tgs < 1) tgs = 1;
            const int64_t gt = ceil_div64(slab_tiles, tgs);
            tgs = ceil_div64(slab_tiles, gt);
            const int64_t groups = tgs * pl->parts;
            if (stride > 1 && groups >= 2 * pl->kp && groups <= 32 * tc::kMaxGroups32) {
                pl->groups = static_cast<int>(groups);
                pl->group_tiles = static_cast<int>(gt);
                pl->tile_stride = static_cast<int>(stride);
                pl->groups_per_slab = static_cast<int>(tgs);
                pl->off_groupmax = carve(sizeof(uint32_t) * q_pad * groups);
            }
        }
        if (p->precision == RADAR_PREC_FP32) {
            // exact re-run of uncertified queries: enqueued unconditionally with a device-side count, sized for all q
            plan_rerun(pl, c, q, p->k, sms, carve);
        }
    }
    pl->total = off;
    return RADAR_OK;
}

static int validate_search(const radar_corpus_t* c, int64_t q, const radar_search_params_t* p) {
    RADAR_ARG_CHECK(c && p, "null corpus/params");
    RADAR_ARG_CHECK(p->mode >= RADAR_MODE_DPR && p->mode <= RADAR_MODE_HYBRID, "bad mode %d", p->mode);
    RADAR_ARG_CHECK(p->precision == RADAR_PREC_BF16 || p->precision == RADAR_PREC_FP32, "bad precision %d",
                    p->precision);
    RADAR_ARG_CHECK(p->algo >= RADAR_ALGO_AUTO && p->algo <= RADAR_ALGO_KL_STREAM, "bad algo %d", p->algo);
    RADAR_ARG_CHECK(q >= 0, "negative query count");
    RADAR_ARG_CHECK(c->n >= 1 && c->n < 0xFFFFFFFFll, "corpus rows %lld out of range [1, 2^32-1)", (long long)c->n);
    RADAR_ARG_CHECK(p->k >= 1 && p->k <= RADAR_MAX_K, "k=%d out of range [1,%d]", p->k, RADAR_MAX_K);
    RADAR_ARG_CHECK(p->k <= c->n, "k=%d exceeds corpus rows %lld (clamp k in the caller as dpr.py:308 does)", p->k,
                    (long long)c->n);
    if (p->mode != RADAR_MODE_KL) {
        RADAR_ARG_CHECK(c->emb_f32, "corpus.emb_f32 is required for DPR/hybrid");
        RADAR_ARG_CHECK(c->d >= 4 && c->d % 4 == 0, "embedding dim %d must be a positive multiple of 4", c->d);
    }
    if (p->mode != RADAR_MODE_DPR) RADAR_ARG_CHECK(c->logq16, "corpus.logq16 is required for KL/hybrid");
    RADAR_ARG_CHECK(c->idx_offset >= 0 && c->idx_offset + c->n < 0xFFFFFFFFll,
                    "idx_offset + n must stay below 2^32-1");
    return RADAR_OK;
}

// launch with programmatic stream serialization: the kernel may be scheduled before its predecessor in the stream has finished
// and synchronises on the device (griddepcontrol.wait) before reading the predecessor's output
template <typename Args>
static cudaError_t launch_pdl(void (*kernel)(Args), dim3 grid, dim3 block, cudaStream_t st, const Args& args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

static int launch_scan(const ScanArgs& a, int64_t q_tiles, cudaStream_t st) {
    dim3 grid(static_cast<unsigned>(q_tiles), static_cast<unsigned>(a.parts));
    if (a.mode == RADAR_MODE_DPR) simt_scan_kernel<true, false><<<grid, kScanThreads, kScanSmemBytes, st>>>(a);
    else if (a.mode == RADAR_MODE_KL) simt_scan_kernel<false, true><<<grid, kScanThreads, kScanSmemBytes, st>>>(a);
    else simt_scan_kernel<true, true><<<grid, kScanThreads, kScanSmemBytes, st>>>(a);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

// Exact re-run (canonical CUDA-core scan) of the queries listed in ulist[0 .. *ucount): enqueued unconditionally, the
// kernels read the count on the device and exit at once when it is zero -- no host round trip inside a search call.
// One wave: queries ulist[skip .. skip + q_max) scanned with `parts` corpus slabs.
static int launch_exact_rerun_wave(const radar_corpus_t* corpus, const radar_queries_t* queries, int mode, int k, float alpha,
                                   float oma, int64_t q_max, uint32_t skip, const uint32_t* ucount, const uint32_t* ulist,
                                   int parts, int64_t rows_per_part, uint64_t* cand, uint32_t* cnt, uint64_t* sel,
                                   float* out_scores, int64_t* out_idx, uint64_t* out_packed, cudaStream_t st) {
    ScanArgs a{};
    a.q_emb = queries->emb_f32; a.p16 = queries->p16; a.entropy = queries->entropy;
    a.c_emb = corpus->emb_f32; a.logq16 = corpus->logq16; a.qmap = ulist + skip; a.nq_dev = ucount; a.nq_skip = skip;
    a.nq = q_max; a.n = corpus->n; a.d = corpus->d; a.mode = mode; a.alpha = alpha; a.oma = oma;
    a.parts = parts; a.rows_per_part = rows_per_part; a.kp = k; a.cand = cand; a.cnt = cnt;
    int rc = launch_scan(a, ceil_div64(q_max, kScanTQ), st);
    if (rc) return rc;
    select_kernel<<<static_cast<unsigned>(ceil_div64(q_max, kSelWarps)), kSelWarps * 32, 0, st>>>(
        cand, cnt, nullptr, q_max, ucount, parts, kCandCap, k, sel, nullptr, skip);
    RADAR_CUDA_CHECK(cudaGetLastError());
    FinalArgs g{};
    g.sel = sel; g.R = k; g.k = k; g.mode = mode; g.sort = 0; g.qmap = ulist + skip; g.nq_dev = ucount; g.nq_skip = skip;
    g.idx_offset = corpus->idx_offset; g.out_scores = out_scores; g.out_idx = out_idx; g.out_packed = out_packed;
    RADAR_CUDA_CHECK(launch_final(g, q_max, st));
    return RADAR_OK;
}

// Two waves (see Plan::fa_*): the first kRerunHead listed queries with many corpus slabs, the rest with the all-queries plan.
static int launch_exact_rerun(const radar_corpus_t* corpus, const radar_queries_t* queries, int mode, int k, float alpha,
                              float oma, int64_t q, const uint32_t* ucount, const uint32_t* ulist, const Plan& pl, uint8_t* ws,
                              float* out_scores, int64_t* out_idx, uint64_t* out_packed, cudaStream_t st, int* launches) {
    const int64_t fa_q = q < kRerunHead ? q : kRerunHead;
    int rc = launch_exact_rerun_wave(corpus, queries, mode, k, alpha, oma, fa_q, 0u, ucount, ulist, pl.fa_parts,
                                     pl.fa_rows_per_part, reinterpret_cast<uint64_t*>(ws + pl.off_fa_cand),
                                     reinterpret_cast<uint32_t*>(ws + pl.off_fa_cnt),
                                     reinterpret_cast<uint64_t*>(ws + pl.off_fa_sel), out_scores, out_idx, out_packed, st);
    if (rc) return rc;
    *launches += 3;
    if (q > kRerunHead) {
        rc = launch_exact_rerun_wave(corpus, queries, mode, k, alpha, oma, q - kRerunHead, static_cast<uint32_t>(kRerunHead),
                                     ucount, ulist, pl.fb_parts, pl.fb_rows_per_part,
                                     reinterpret_cast<uint64_t*>(ws + pl.off_fb_cand),
                                     reinterpret_cast<uint32_t*>(ws + pl.off_fb_cnt),
                                     reinterpret_cast<uint64_t*>(ws + pl.off_fb_sel), out_scores, out_idx, out_packed, st);
        if (rc) return rc;
        *launches += 3;
    }
    return RADAR_OK;
}

}  // namespace radar

using namespace radar;

extern "C" {

const char* radar_last_error(void) { return g_err; }
int radar_abi_version(void) { return RADAR_ABI_VERSION; }

int radar_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    if (sm_count) *sm_count = di.sms;
    if (cc_major) *cc_major = di.major;
    if (cc_minor) *cc_minor = di.minor;
    return RADAR_OK;
}

int radar_set_device(int device) {
    RADAR_CUDA_CHECK(cudaSetDevice(device));
    return RADAR_OK;
}

int radar_get_device(int* device) {
    RADAR_ARG_CHECK(device, "get_device: null pointer");
    RADAR_CUDA_CHECK(cudaGetDevice(device));
    return RADAR_OK;
}

static void prof_destroy() {
    if (g_prof_start) {
        cudaEventDestroy(g_prof_start);
        cudaEventDestroy(g_prof_stop);
        g_prof_start = g_prof_stop = nullptr;
        g_prof_dev = -1;
    }
}

// the event pair lives on the device that is current when it is (re)created; radar_search re-creates it when it
// runs on another device
static int prof_ensure_on_current_device() {
    int dev = 0;
    RADAR_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_prof_start && g_prof_dev == dev) return RADAR_OK;
    prof_destroy();
    RADAR_CUDA_CHECK(cudaEventCreate(&g_prof_start));
    RADAR_CUDA_CHECK(cudaEventCreate(&g_prof_stop));
    g_prof_dev = dev;
    return RADAR_OK;
}

int radar_profile_enable(int enable) {
    if (enable) return prof_ensure_on_current_device();
    prof_destroy();
    return RADAR_OK;
}

int radar_profile_kernel_ms(float* ms_out) {
    RADAR_ARG_CHECK(ms_out && g_prof_start, "profile_kernel_ms: profiling is not enabled");
    RADAR_CUDA_CHECK(cudaEventSynchronize(g_prof_stop));
    RADAR_CUDA_CHECK(cudaEventElapsedTime(ms_out, g_prof_start, g_prof_stop));
    return RADAR_OK;
}

int radar_pack_embeddings(const float* emb_f32, int64_t n, int d, uint16_t* emb_bf16, float* max_norm,
                          void* stream) {
    RADAR_ARG_CHECK(emb_f32 && n >= 0 && d >= 4 && d % 4 == 0, "pack_embeddings: need emb, n >= 0, d %% 4 == 0");
    if (n == 0) return RADAR_OK;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    const int64_t warps_needed = n;
    int64_t blocks = ceil_div64(warps_needed, 8);
    const int64_t cap = static_cast<int64_t>(di.sms) * 16;
    if (blocks > cap) blocks = cap;
    pack_embeddings_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        emb_f32, n, d, reinterpret_cast<__nv_bfloat16*>(emb_bf16), max_norm);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_kl_prepare_corpus(const float* probs, int64_t n, int n_obs, float eps, int normalize, float* logq16,
                            uint16_t* klpack, uint16_t* kl16, void* stream) {
    RADAR_ARG_CHECK(probs && logq16 && n >= 0, "kl_prepare_corpus: null pointer");
    RADAR_ARG_CHECK(n_obs >= 1 && n_obs <= RADAR_NUM_OBS, "n_obs=%d out of range [1,%d]", n_obs, RADAR_NUM_OBS);
    RADAR_ARG_CHECK(eps > 0.0f && eps < 1.0f, "eps must be in (0,1)");
    if (n == 0) return RADAR_OK;
    kl_prepare_corpus_kernel<<<static_cast<unsigned>(ceil_div64(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        probs, n, n_obs, eps, normalize, logq16, reinterpret_cast<__nv_bfloat16*>(klpack), reinterpret_cast<__half*>(kl16));
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_kl_prepare_queries(const float* probs, const uint8_t* mask, int64_t q, int n_obs, float eps,
                             int normalize, float* p16, float* entropy, void* stream) {
    RADAR_ARG_CHECK(probs && p16 && entropy && q >= 0, "kl_prepare_queries: null pointer");
    RADAR_ARG_CHECK(n_obs >= 1 && n_obs <= RADAR_NUM_OBS, "n_obs=%d out of range [1,%d]", n_obs, RADAR_NUM_OBS);
    RADAR_ARG_CHECK(eps > 0.0f && eps < 1.0f, "eps must be in (0,1)");
    if (q == 0) return RADAR_OK;
    kl_prepare_queries_kernel<<<static_cast<unsigned>(ceil_div64(q, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        probs, mask, q, n_obs, eps, normalize, p16, entropy);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

size_t radar_search_workspace_bytes(const radar_corpus_t* corpus, int64_t q, const radar_search_params_t* params) {
    if (validate_search(corpus, q, params)) return 0;
    DeviceInfo di;
    if (get_device_info(&di)) return 0;
    Plan pl;
    if (make_plan(corpus, q, params, di, &pl)) return 0;
    size_t total = pl.total;
#ifdef RADAR_DEBUG
    // the key dump of KL mode runs the general filter, whose plan differs from the dedicated KL kernel's
    if (params->mode == RADAR_MODE_KL && corpus->klpack && pl.algo == RADAR_ALGO_TC_FILTER) {
        Plan legacy;
        if (make_plan(corpus, q, params, di, &legacy, false, true) == RADAR_OK && legacy.total > total) total = legacy.total;
    }
#endif
    return total + 256;
}

int radar_search(const radar_corpus_t* corpus, const radar_queries_t* queries, const radar_search_params_t* params,
                 float* out_scores, int64_t* out_idx, uint64_t* out_packed, void* workspace, size_t workspace_bytes,
                 radar_search_stats_t* stats, void* stream) {
    RADAR_ARG_CHECK(queries, "null queries");
    const int64_t q = queries->q;
    int rc = validate_search(corpus, q, params);
    if (rc) return rc;
    RADAR_ARG_CHECK(out_scores && out_idx, "null output pointer");
    if (params->mode != RADAR_MODE_KL) RADAR_ARG_CHECK(queries->emb_f32, "queries.emb_f32 is required for DPR/hybrid");
    if (params->mode != RADAR_MODE_DPR)
        RADAR_ARG_CHECK(queries->p16 && queries->entropy, "queries.p16/entropy are required for KL/hybrid");
    if (stats) memset(stats, 0, sizeof *stats);
    if (q == 0) return RADAR_OK;
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc) return rc;
    if (g_prof_start) {
        rc = prof_ensure_on_current_device();
        if (rc) return rc;
    }
    const bool search_after = queries->after_idx != nullptr;
    RADAR_ARG_CHECK(!search_after || queries->after_scores, "queries.after_idx needs queries.after_scores");
    Plan pl;
    rc = make_plan(corpus, q, params, di, &pl, search_after);
    if (rc) return rc;
    if (!workspace || workspace_bytes < pl.total) {
        set_error("workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
        return RADAR_E_WORKSPACE;
    }
    RADAR_ARG_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint64_t* cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(ws + pl.off_cnt);
    float* thr = reinterpret_cast<float*>(ws + pl.off_thr);
    uint64_t* sel = reinterpret_cast<uint64_t*>(ws + pl.off_sel);
    float* bound = reinterpret_cast<float*>(ws + pl.off_bound);
    float* qerr = reinterpret_cast<float*>(ws + pl.off_qerr);
    uint32_t* ucount = reinterpret_cast<uint32_t*>(ws + pl.off_ucount);
    uint32_t* ulist = reinterpret_cast<uint32_t*>(ws + pl.off_ulist);
    const float alpha = params->alpha;
    const float oma = 1.0f - alpha;
    int launches = 0;
    int64_t uncertified = 0;
    bool have_ucount = false;  // ucount holds the number of exactly re-run queries (read back only for statistics)
    unsigned long long* clk_dev = nullptr;

    if (pl.algo == RADAR_ALGO_KL_STREAM) {
        // ---- few queries, large corpus: pooled-candidate tcgen05 stream (kl_stream.cuh) ------------------------
        const int64_t qp = pl.n_pad;
        uint32_t* zero = reinterpret_cast<uint32_t*>(ws + pl.off_ks_zero);
        uint32_t *gcnt = zero, *lock = zero + qp, *processed = zero + 2 * qp, *best_n = zero + 3 * qp, *gthr = zero + 4 * qp;
        uint64_t* pool = reinterpret_cast<uint64_t*>(ws + pl.off_ks_pool);
        uint64_t* best = reinterpret_cast<uint64_t*>(ws + pl.off_ks_best);
        uint32_t* tilemax = reinterpret_cast<uint32_t*>(ws + pl.off_ks_tilemax);
        uint16_t* apack = reinterpret_cast<uint16_t*>(ws + pl.off_apack);
        const size_t shift_off = align_up(sizeof(uint16_t) * static_cast<size_t>(qp) * RADAR_KLPACK, 256);
        float* qshift = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(apack) + shift_off);
        RADAR_CUDA_CHECK(cudaMemsetAsync(zero, 0, pl.ks_zero_bytes, st));
        if (pl.kl_fmt == klf::kFmtBf16x3) {
            tc::PackArgs pa{};
            pa.q_emb = nullptr; pa.p16 = queries->p16; pa.entropy = queries->entropy; pa.q = q; pa.q_pad = qp;
            pa.d = corpus->d; pa.mode = RADAR_MODE_KL; pa.alpha = alpha; pa.oma = oma;
            pa.emb_max_norm = corpus->emb_max_norm; pa.logq_max_abs = corpus->logq_max_abs;
            tc::fill_col_max(corpus, pa.logq_col_max);
            pa.apack = apack; pa.qshift = qshift; pa.qerr = qerr;
            tc::query_pack_kernel<<<static_cast<unsigned>((qp * 32 + 255) / 256), 256, 0, st>>>(pa);
        } else {
            klf::KlPackArgs pa{};
            pa.p16 = queries->p16; pa.entropy = queries->entropy; pa.q = q; pa.q_pad = qp; pa.fmt = pl.kl_fmt;
            tc::fill_col_max(corpus, pa.logq_col_max);
            pa.apack = apack; pa.qshift = qshift; pa.qerr = qerr;
            klf::klf_pack_kernel<<<static_cast<unsigned>((qp * 32 + 255) / 256), 256, 0, st>>>(pa);
        }
        RADAR_CUDA_CHECK(cudaGetLastError());
        CUtensorMap map_kl, map_q;
        memset(&map_kl, 0, sizeof map_kl);
        memset(&map_q, 0, sizeof map_q);
        if (pl.kl_fmt == klf::kFmtBf16x3)
            rc = tc::encode_2d_bf16(&map_kl, corpus->klpack, RADAR_KLPACK, corpus->n, RADAR_KLPACK, 256, CU_TENSOR_MAP_SWIZZLE_64B);
        else
            rc = tc::encode_2d_bf16(&map_kl, corpus->kl16, kObsPad, corpus->n, kObsPad, 256, CU_TENSOR_MAP_SWIZZLE_32B);
        if (rc) return rc;
        rc = tc::encode_2d_bf16(&map_q, apack, RADAR_KLPACK, qp, RADAR_KLPACK, static_cast<uint32_t>(qp / 2),
                                CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        kls::StreamArgs sa{};
        sa.n = corpus->n; sa.tiles = pl.tiles; sa.q = static_cast<int>(q); sa.n_pad = pl.n_pad; sa.kp = pl.kp;
        sa.pool_cap = pl.pool_cap; sa.qshift = qshift; sa.gthr = gthr; sa.gcnt = gcnt; sa.lock = lock;
        sa.processed = processed; sa.best_n = best_n; sa.best = best; sa.pool = pool;
        sa.reload = kls::kThrReload;  // measured: 4 beats 1, 2, 8, 16 on kl_latency (fresher thresholds vs reload cost)
        sa.total_tiles = pl.tiles; sa.tilemax = tilemax;
        auto stream_kernel = pl.kl_fmt == klf::kFmtBf16x3 ? kls::kl_stream_kernel<klf::kFmtBf16x3, false>
                             : pl.kl_fmt == klf::kFmtF16x1 ? kls::kl_stream_kernel<klf::kFmtF16x1, false>
                                                           : kls::kl_stream_kernel<klf::kFmtF16x2, false>;
        auto boot_kernel = pl.kl_fmt == klf::kFmtBf16x3 ? kls::kl_stream_kernel<klf::kFmtBf16x3, true>
                           : pl.kl_fmt == klf::kFmtF16x1 ? kls::kl_stream_kernel<klf::kFmtF16x1, true>
                                                         : kls::kl_stream_kernel<klf::kFmtF16x2, true>;
        RADAR_CUDA_CHECK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(kls::kStreamSmemBytes)));
        RADAR_CUDA_CHECK(cudaFuncSetAttribute(boot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(kls::kStreamSmemBytes)));
        cudaLaunchConfig_t cfg{};
        cfg.blockDim = dim3(kls::kStreamThreads);
        cfg.dynamicSmemBytes = kls::kStreamSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        {
            // boot: the same tcgen05 pipeline over a strided sample of super-tiles -> per-query maxima -> initial thresholds
            kls::StreamArgs sb = sa;
            sb.tiles = pl.sample_tiles;
            int64_t bunits = pl.units;
            if (bunits > sb.tiles) bunits = sb.tiles;
            cfg.gridDim = dim3(static_cast<unsigned>(bunits * 2));
            RADAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, boot_kernel, map_kl, map_q, sb));
            kls::BootArgs ba{};
            ba.tilemax = tilemax; ba.sample_tiles = pl.sample_tiles; ba.kp = pl.kp; ba.gthr = gthr;
            RADAR_CUDA_CHECK(launch_pdl(kls::kl_boot_threshold_kernel, dim3(static_cast<unsigned>(q)), dim3(kls::kBootThreads), st, ba));
        }
        int64_t units = pl.units;
        if (units > pl.tiles) units = pl.tiles;
        cfg.gridDim = dim3(static_cast<unsigned>(units * 2));
        // programmatic dependent launch along boot -> thresholds -> stream -> final (kl_stream.cuh: pdl_wait / pdl_trigger); with
        // the profiling events in place the stream kernel is launched the plain way so that the events bracket exactly its run
        cudaLaunchAttribute attr_pdl[2];
        attr_pdl[0] = attr[0];
        attr_pdl[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr_pdl[1].val.programmaticStreamSerializationAllowed = 1;
        if (!g_prof_start) {
            cfg.attrs = attr_pdl;
            cfg.numAttrs = 2;
        }
        if (g_prof_start) RADAR_CUDA_CHECK(cudaEventRecord(g_prof_start, st));
        RADAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, stream_kernel, map_kl, map_q, sa));
        if (g_prof_stop) RADAR_CUDA_CHECK(cudaEventRecord(g_prof_stop, st));
        kls::StreamFinalArgs fa{};
        fa.p16 = queries->p16; fa.entropy = queries->entropy; fa.logq16 = corpus->logq16; fa.qerr = qerr; fa.gthr = gthr;
        fa.gcnt = gcnt; fa.pool = pool; fa.q = static_cast<int>(q); fa.k = params->k; fa.pool_cap = pl.pool_cap;
        fa.certify = params->precision == RADAR_PREC_FP32 ? 1 : 0; fa.idx_offset = corpus->idx_offset;
        fa.out_scores = out_scores; fa.out_idx = out_idx; fa.out_packed = out_packed;
        fa.uncert_count = ucount; fa.uncert_list = ulist;
        if (g_prof_stop) {
            kls::kl_stream_final_kernel<<<static_cast<unsigned>(q), kls::kFinThreads, 0, st>>>(fa);
            RADAR_CUDA_CHECK(cudaGetLastError());
        } else {
            RADAR_CUDA_CHECK(launch_pdl(kls::kl_stream_final_kernel, dim3(static_cast<unsigned>(q)), dim3(kls::kFinThreads), st, fa));
        }
        launches += 5;
        // queries whose pool overflowed or (fp32 mode) whose certificate failed are re-run by the exact scan
        rc = launch_exact_rerun(corpus, queries, RADAR_MODE_KL, params->k, alpha, oma, q, ucount, ulist, pl, ws, out_scores,
                                out_idx, out_packed, st, &launches);
        if (rc) return rc;
        have_ucount = true;
    } else if (pl.algo == RADAR_ALGO_SIMT_EXACT) {
        ScanArgs a{};
        a.q_emb = queries->emb_f32; a.p16 = queries->p16; a.entropy = queries->entropy;
        a.c_emb = corpus->emb_f32; a.logq16 = corpus->logq16; a.qmap = nullptr;
        a.nq = q; a.n = corpus->n; a.d = corpus->d; a.mode = params->mode; a.alpha = alpha; a.oma = oma;
        a.parts = pl.parts; a.rows_per_part = pl.rows_per_part; a.kp = pl.kp; a.cand = cand; a.cnt = cnt;
        a.after_scores = queries->after_scores; a.after_idx = queries->after_idx; a.idx_offset = corpus->idx_offset;
        if (g_prof_start) RADAR_CUDA_CHECK(cudaEventRecord(g_prof_start, st));
        rc = launch_scan(a, pl.q_tiles, st);
        if (rc) return rc;
        if (g_prof_stop) RADAR_CUDA_CHECK(cudaEventRecord(g_prof_stop, st));
        ++launches;
        select_kernel<<<static_cast<unsigned>(ceil_div64(q, kSelWarps)), kSelWarps * 32, 0, st>>>(
            cand, cnt, nullptr, q, nullptr, pl.parts, kCandCap, pl.R, sel, nullptr);
        RADAR_CUDA_CHECK(cudaGetLastError());
        ++launches;
        FinalArgs f{};
        f.sel = sel; f.R = pl.R; f.k = params->k; f.mode = params->mode; f.sort = 0; f.qmap = nullptr;
        f.idx_offset = corpus->idx_offset; f.out_scores = out_scores; f.out_idx = out_idx; f.out_packed = out_packed;
        RADAR_CUDA_CHECK(launch_final(f, q, st));
        ++launches;
    } else {
        const bool certify = params->precision == RADAR_PREC_FP32;
        RADAR_CUDA_CHECK(cudaMemsetAsync(ucount, 0, sizeof(uint32_t) * 4, st));
        tc::FilterLaunch fl{};
        fl.corpus = corpus; fl.queries = queries; fl.mode = params->mode; fl.alpha = alpha; fl.oma = oma;
        fl.q = q; fl.q_tiles = pl.q_tiles; fl.parts = pl.parts; fl.rows_per_part = pl.rows_per_part;
        fl.kp = pl.kp; fl.cand = cand; fl.cnt = cnt; fl.thr = thr; fl.qerr = qerr;
        fl.gthr = reinterpret_cast<uint32_t*>(ws + pl.off_gthr);
        fl.apack = reinterpret_cast<uint16_t*>(ws + pl.off_apack);
        fl.units = pl.units; fl.device_sms = di.sms;
        fl.dbg_scores = nullptr;
        fl.groups = pl.groups; fl.group_tiles = pl.group_tiles; fl.tile_stride = pl.tile_stride;
        fl.groups_per_slab = pl.groups_per_slab;
        fl.groupmax = pl.groups ? reinterpret_cast<uint32_t*>(ws + pl.off_groupmax) : nullptr;
        fl.ev_start = g_prof_start; fl.ev_stop = g_prof_stop;
        int nl = 0;
        if (pl.klf) {
            klf::KlfLaunch kl{};
            kl.corpus = corpus; kl.queries = queries; kl.fmt = pl.kl_fmt; kl.q = q; kl.q_tiles = pl.q_tiles;
            kl.parts = pl.parts; kl.rows_per_part = pl.rows_per_part; kl.kp = pl.kp; kl.cand = cand; kl.cnt = cnt;
            kl.thr = thr; kl.gthr = fl.gthr; kl.qerr = qerr; kl.apack = fl.apack; kl.units = pl.units;
            kl.groupmax = fl.groupmax; kl.groups = pl.groups; kl.group_tiles = pl.group_tiles;
            kl.tile_stride = pl.tile_stride; kl.groups_per_slab = pl.groups_per_slab;
            kl.ev_start = g_prof_start; kl.ev_stop = g_prof_stop;
            rc = klf::launch_kl_filter(kl, st, &nl);
            if (rc) return rc;
            clk_dev = kl.clk_dev;
        } else {
            rc = tc::launch_filter(fl, st, &nl);
            if (rc) return rc;
            clk_dev = fl.clk_dev;
        }
        launches += nl;
        const uint8_t* done = nullptr;
        if (pl.klf && pl.buf_parts <= 32 && params->k <= 64) {
            // short candidate lists (the usual case after a prepass): gather + canonical re-score + rank + certificate in
            // one warp-per-query kernel; queries with more than 64 candidates fall through to the generic kernels below
            KlFinishArgs ka{};
            ka.cand = cand; ka.cnt = cnt; ka.thr_final = thr; ka.p16 = queries->p16; ka.entropy = queries->entropy;
            ka.logq16 = corpus->logq16; ka.qerr = certify ? qerr : nullptr; ka.nq = q; ka.parts = pl.buf_parts; ka.k = params->k;
            ka.idx_offset = corpus->idx_offset; ka.out_scores = out_scores; ka.out_idx = out_idx; ka.out_packed = out_packed;
            ka.uncert_count = ucount; ka.uncert_list = ulist; ka.done = ws + pl.off_done;
            kl_finish_kernel<<<static_cast<unsigned>(ceil_div64(q, kFinishWarps)), kFinishWarps * 32, 0, st>>>(ka);
            RADAR_CUDA_CHECK(cudaGetLastError());
            ++launches;
            done = ws + pl.off_done;
        }
        select_kernel<<<static_cast<unsigned>(ceil_div64(q, kSelWarps)), kSelWarps * 32, 0, st>>>(
            cand, cnt, thr, q, nullptr, pl.buf_parts, kCandCap, pl.R, sel, bound, 0u, done);
        RADAR_CUDA_CHECK(cudaGetLastError());
        ++launches;
        RescoreArgs r{};
        r.q_emb = queries->emb_f32; r.p16 = queries->p16; r.entropy = queries->entropy;
        r.c_emb = corpus->emb_f32; r.logq16 = corpus->logq16; r.qmap = nullptr; r.nq = q; r.d = corpus->d;
        r.mode = params->mode; r.alpha = alpha; r.oma = oma; r.R = pl.R; r.sel = sel; r.done = done;
        r.qerr = qerr; r.bound = bound; r.k = params->k;
        RADAR_CUDA_CHECK(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              static_cast<int>(kRescoreSmemBytes)));
        rescore_kernel<<<static_cast<unsigned>(ceil_div64(q, kRsWarps)), kRsWarps * 32, kRescoreSmemBytes, st>>>(r);
        RADAR_CUDA_CHECK(cudaGetLastError());
        ++launches;
        FinalArgs f{};
        f.sel = sel; f.R = pl.R; f.k = params->k; f.mode = params->mode; f.sort = 1; f.qmap = nullptr; f.done = done;
        f.idx_offset = corpus->idx_offset; f.out_scores = out_scores; f.out_idx = out_idx; f.out_packed = out_packed;
        if (certify) {
            f.bound = bound; f.qerr = qerr; f.uncert_count = ucount; f.uncert_list = ulist;
        }
        RADAR_CUDA_CHECK(launch_final(f, q, st));
        ++launches;
        if (certify) {
            rc = launch_exact_rerun(corpus, queries, params->mode, params->k, alpha, oma, q, ucount, ulist, pl, ws,
                                    out_scores, out_idx, out_packed, st, &launches);
            if (rc) return rc;
            have_ucount = true;
        }
    }
    if (stats) {
        if (have_ucount) {
            uint32_t h_count = 0;
            RADAR_CUDA_CHECK(cudaMemcpyAsync(&h_count, ucount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            RADAR_CUDA_CHECK(cudaStreamSynchronize(st));
            uncertified = h_count;
        }
        RADAR_CUDA_CHECK(cudaStreamSynchronize(st));
        stats->algo_used = pl.algo;
        stats->kernel_launches = launches;
        stats->uncertified = uncertified;
        stats->parts = pl.parts;
        stats->kprime = pl.kp;
        if (clk_dev) {
            unsigned long long h[2] = {0, 0};
            RADAR_CUDA_CHECK(cudaMemcpyAsync(h, clk_dev, sizeof h, cudaMemcpyDeviceToHost, st));
            RADAR_CUDA_CHECK(cudaStreamSynchronize(st));
            if (h[1] > 0) stats->filter_sm_mhz = static_cast<float>(1e3 * static_cast<double>(h[0]) / static_cast<double>(h[1]));
        }
    }
    return RADAR_OK;
}

#ifdef RADAR_DEBUG
int radar_debug_filter_keys(const radar_corpus_t* corpus, const radar_queries_t* queries,
                            const radar_search_params_t* params, float* out_keys, void* workspace,
                            size_t workspace_bytes, void* stream) {
    RADAR_ARG_CHECK(queries && out_keys, "debug_filter_keys: null pointer");
    const int64_t q = queries->q;
    int rc = validate_search(corpus, q, params);
    if (rc) return rc;
    DeviceInfo di;
    rc = get_device_info(&di);
    if (rc) return rc;
    radar_search_params_t p2 = *params;
    p2.algo = RADAR_ALGO_TC_FILTER;
    Plan pl;
    rc = make_plan(corpus, q, &p2, di, &pl, false, true);
    if (rc) return rc;
    if (!workspace || workspace_bytes < pl.total) {
        set_error("workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
        return RADAR_E_WORKSPACE;
    }
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    tc::FilterLaunch fl{};
    fl.corpus = corpus; fl.queries = queries; fl.mode = p2.mode; fl.alpha = p2.alpha; fl.oma = 1.0f - p2.alpha;
    fl.q = q; fl.q_tiles = pl.q_tiles; fl.parts = pl.parts; fl.rows_per_part = pl.rows_per_part;
    fl.kp = pl.kp; fl.gthr = reinterpret_cast<uint32_t*>(ws + pl.off_gthr);
    fl.cand = reinterpret_cast<uint64_t*>(ws + pl.off_cand);
    fl.cnt = reinterpret_cast<uint32_t*>(ws + pl.off_cnt);
    fl.thr = reinterpret_cast<float*>(ws + pl.off_thr);
    fl.qerr = reinterpret_cast<float*>(ws + pl.off_qerr);
    fl.apack = reinterpret_cast<uint16_t*>(ws + pl.off_apack);
    fl.units = pl.units; fl.device_sms = di.sms;
    fl.dbg_scores = out_keys;
    int nl = 0;
    return tc::launch_filter(fl, static_cast<cudaStream_t>(stream), &nl);
}
#endif  // RADAR_DEBUG

int radar_merge_topk(const float* cand_scores, const int64_t* cand_idx, int64_t q, int parts, int k_in, int k_out,
                     int ascending, float* out_scores, int64_t* out_idx, void* stream) {
    RADAR_ARG_CHECK(cand_scores && cand_idx && out_scores && out_idx, "merge_topk: null pointer");
    RADAR_ARG_CHECK(q >= 0 && parts >= 1 && k_in >= 1 && k_out >= 1, "merge_topk: bad sizes");
    RADAR_ARG_CHECK(static_cast<int64_t>(parts) * k_in <= kMergeCap, "merge_topk: parts*k_in=%lld exceeds %d",
                    (long long)parts * k_in, kMergeCap);
    RADAR_ARG_CHECK(k_out <= parts * k_in, "merge_topk: k_out exceeds the candidate count");
    if (q == 0) return RADAR_OK;
    merge_kernel<<<static_cast<unsigned>(q), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        cand_scores, cand_idx, q, parts, k_in, k_out, ascending, out_scores, out_idx);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_merge_packed(const uint64_t* cand, int64_t q, int parts, int k_in, int k_out, int mode, float* out_scores,
                       int64_t* out_idx, void* stream) {
    RADAR_ARG_CHECK(cand && out_scores && out_idx, "merge_packed: null pointer");
    RADAR_ARG_CHECK(q >= 0 && parts >= 1 && k_in >= 1 && k_out >= 1, "merge_packed: bad sizes");
    RADAR_ARG_CHECK(mode >= RADAR_MODE_DPR && mode <= RADAR_MODE_HYBRID, "merge_packed: bad mode %d", mode);
    RADAR_ARG_CHECK(static_cast<int64_t>(parts) * k_in <= kMergeCap, "merge_packed: parts*k_in=%lld exceeds %d",
                    (long long)parts * k_in, kMergeCap);
    RADAR_ARG_CHECK(k_out <= parts * k_in, "merge_packed: k_out exceeds the candidate count");
    if (q == 0) return RADAR_OK;
    merge_packed_kernel<<<static_cast<unsigned>(q), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        cand, q, parts, k_in, k_out, mode, out_scores, out_idx);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_rerank_overlap(const uint16_t* case_bits, const uint16_t* missing_bits, int64_t q, int k,
                         double* out_scores, int32_t* out_order, void* stream) {
    RADAR_ARG_CHECK(case_bits && missing_bits && out_scores && out_order, "rerank_overlap: null pointer");
    RADAR_ARG_CHECK(q >= 0 && k >= 1 && k <= 1024, "rerank_overlap: bad sizes");
    if (q == 0) return RADAR_OK;
    rerank_overlap_kernel<<<static_cast<unsigned>(ceil_div64(q, 128)), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        case_bits, missing_bits, q, k, out_scores, out_order);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_gather_bits(const uint16_t* table, int64_t n, const int64_t* idx, int64_t q, int k, int64_t idx_offset,
                      uint16_t* out, void* stream) {
    RADAR_ARG_CHECK(table && idx && out && n >= 0 && q >= 0 && k >= 1, "gather_bits: bad arguments");
    const int64_t total = q * k;
    if (total == 0) return RADAR_OK;
    gather_bits_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        table, n, idx, total, idx_offset, out);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

int radar_project_normalize(const float* x, const float* w, const float* bias, int64_t b, int in_dim, int out_dim,
                            float* y, void* stream) {
    RADAR_ARG_CHECK(x && w && y && b >= 0 && in_dim >= 1 && out_dim >= 1, "project_normalize: bad arguments");
    RADAR_ARG_CHECK(in_dim <= 8192, "project_normalize: in_dim too large");
    if (b == 0) return RADAR_OK;
    constexpr int kT = 256;
    const size_t smem = sizeof(float) * (in_dim + kT / 32);
    project_normalize_kernel<kT><<<static_cast<unsigned>(b), kT, smem, static_cast<cudaStream_t>(stream)>>>(
        x, w, bias, in_dim, out_dim, y);
    RADAR_CUDA_CHECK(cudaGetLastError());
    return RADAR_OK;
}

size_t radar_project_workspace_bytes(int64_t b, int in_dim, int out_dim) {
    if (b < 0 || !proj::proj_tc_supported(in_dim, out_dim)) return 0;
    return proj::proj_workspace_bytes(b, in_dim);
}

int radar_project_normalize_tc(const float* x, const float* w, const float* bias, int64_t b, int in_dim, int out_dim,
                               float* y, uint16_t* y_bf16, void* workspace, size_t workspace_bytes, void* stream) {
    RADAR_ARG_CHECK(x && w && (y || y_bf16) && b >= 0, "project_normalize_tc: bad arguments");
    RADAR_ARG_CHECK(proj::proj_tc_supported(in_dim, out_dim), "project_normalize_tc: needs out_dim == 512 and in_dim %% 32 == 0");
    if (b == 0) return RADAR_OK;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    if (di.major != 10) {
        set_error("project_normalize_tc needs an sm_100 device");
        return RADAR_E_ARCH;
    }
    if (!workspace || workspace_bytes < proj::proj_workspace_bytes(b, in_dim)) {
        set_error("project_normalize_tc: workspace too small: need %zu bytes, got %zu", proj::proj_workspace_bytes(b, in_dim),
                  workspace_bytes);
        return RADAR_E_WORKSPACE;
    }
    return proj::launch_proj_tc(x, w, bias, b, in_dim, y, y_bf16, workspace, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

#ifdef RADAR_KLS_TIMING
extern "C" int radar_debug_kls_timing(unsigned long long* out) {  // experiment builds only (tools/kls_timing.py)
    return cudaMemcpyFromSymbol(out, radar::kls::g_kls_timing, sizeof(unsigned long long) * 296 * 4) == cudaSuccess ? 0 : 1;
}
#endif

#ifdef RADAR_TC_TIMING
extern "C" int radar_debug_tc_timing(unsigned long long* out) {  // experiment builds only (tools/tc_timing.py)
    return cudaMemcpyFromSymbol(out, radar::tc::g_tc_timing, sizeof(unsigned long long) * 296 * 2) == cudaSuccess ? 0 : 1;
}
#endif
