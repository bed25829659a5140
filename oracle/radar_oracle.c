/*
 * radar_oracle.c -- CPU restatement of RADAR's case-retrieval arithmetic in the CANONICAL fp32 form
 * that the CUDA path promises to reproduce bit for bit.  TEST INFRASTRUCTURE ONLY: nothing under
 * radar_multimodal_radiology_b200/ may link or load this file; tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg use it as the checker.
 *
 * What is restated (paths relative to /root/reference):
 *   - faiss.IndexFlatIP.search as called at annotate_retrieve/modeling_dense_passage_retrieval.py:313
 *     (exact inner product + k largest, descending, int64 ids in insertion order).  faiss is a
 *     third-party dependency absent from /root/reference and unpinned there; its published
 *     brute-force semantics are restated.
 *   - KL / hybrid / mask: NO reference code exists (src/knowledge/__init__.py is 0 bytes;
 *     hybrid_alpha at modeling_dense_passage_retrieval.py:187 is never read).  The definitions are
 *     SURVEY.md section 8c's frozen specification -> parity unpinned for these pieces.
 *
 * Canonical arithmetic (every operation IEEE-754 binary32, round to nearest even, no contraction
 * other than the explicit fmaf calls; compile with -ffp-contract=off):
 *   corpus log table   L[n][j] = (float)log((double)clamp(q[n][j], eps, 1)),  j < 14;  L[n][14..15] = 0
 *   query table        pc[i][j] = mask ? clamp(p[i][j], eps, 1) : 0 ; lp = mask ? (float)log((double)pc) : 0
 *   entropy            H[i] = fma-chain_{j=0..13}  H = fmaf(pc[j], lp[j], H),       H0 = 0
 *   cross term         X[i][n] = fma-chain_{j=0..13} X = fmaf(pc[i][j], L[n][j], X), X0 = 0
 *   KL                 KL = H - X
 *   inner product      ip = fma-chain_{t=0..D-1} ip = fmaf(eq[i][t], ec[n][t], ip), ip0 = 0
 *   hybrid             s = fmaf(alpha, ip, -((1.0f - alpha) * KL))
 *   ranking            DPR / hybrid: (score descending, id ascending); KL: (KL ascending, id ascending)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NUM_OBS 14
#define OBS_PAD 16
#define MODE_DPR 0
#define MODE_KL 1
#define MODE_HYBRID 2
#define ROWBLK 16

int radar_oracle_version(void) { return 1; }

/* ---- tiny pthread parallel-for (libgomp is not in this image) -------------------------------- */
static int g_threads = 0; /* 0 = all online cores */

void radar_oracle_set_threads(int t) { g_threads = t; }

int radar_oracle_max_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)(n > 256 ? 256 : n) : 1;
}

typedef void (*range_fn)(int64_t begin, int64_t end, void* ctx);
typedef struct { range_fn fn; void* ctx; int64_t n, chunk; volatile int64_t* next; } pf_job;

static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(j->next, j->chunk, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        int64_t e = b + j->chunk < j->n ? b + j->chunk : j->n;
        j->fn(b, e, j->ctx);
    }
    return NULL;
}

static void parallel_for(int64_t n, int64_t chunk, range_fn fn, void* ctx) {
    int nt = radar_oracle_max_threads();
    if (chunk < 1) chunk = 1;
    if (nt > (n + chunk - 1) / chunk) nt = (int)((n + chunk - 1) / chunk);
    if (nt <= 1) { if (n > 0) fn(0, n, ctx); return; }
    volatile int64_t next = 0;
    pf_job job = {fn, ctx, n, chunk, &next};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nt - 1; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

static inline float clampf(float x, float lo, float hi) {
    x = x < lo ? lo : x; /* NaN stays NaN: not part of the domain */
    return x > hi ? hi : x;
}

static void normalize_row(const float* src, float* dst, int n_obs) {
    float s = 0.0f;
    for (int j = 0; j < n_obs; ++j) s = s + src[j];
    for (int j = 0; j < n_obs; ++j) dst[j] = src[j] / s;
}

/* corpus side: float32[N,n_obs] probabilities -> float32[N,16] log table */
typedef struct { const float* probs; const uint8_t* mask; int n_obs; float eps; int normalize;
                 float* out16; float* entropy; } prep_ctx;

static void prep_corpus_range(int64_t b, int64_t e, void* vc) {
    prep_ctx* c = (prep_ctx*)vc;
    for (int64_t r = b; r < e; ++r) {
        float row[OBS_PAD];
        if (c->normalize) normalize_row(c->probs + r * c->n_obs, row, c->n_obs);
        else memcpy(row, c->probs + r * c->n_obs, sizeof(float) * c->n_obs);
        for (int j = 0; j < OBS_PAD; ++j) {
            float v = 0.0f;
            if (j < c->n_obs) v = (float)log((double)clampf(row[j], c->eps, 1.0f));
            c->out16[r * OBS_PAD + j] = v;
        }
    }
}

void radar_oracle_prepare_corpus(const float* probs, int64_t n, int n_obs, float eps, int normalize,
                                 float* logq16) {
    prep_ctx c = {probs, NULL, n_obs, eps, normalize, logq16, NULL};
    parallel_for(n, 4096, prep_corpus_range, &c);
}

/* query side: probabilities (+ optional uint8 mask [Q,n_obs]) -> p16 float32[Q,16], entropy float32[Q] */
static void prep_query_range(int64_t b, int64_t e, void* vc) {
    prep_ctx* c = (prep_ctx*)vc;
    for (int64_t r = b; r < e; ++r) {
        float row[OBS_PAD];
        if (c->normalize) normalize_row(c->probs + r * c->n_obs, row, c->n_obs);
        else memcpy(row, c->probs + r * c->n_obs, sizeof(float) * c->n_obs);
        float h = 0.0f;
        for (int j = 0; j < OBS_PAD; ++j) {
            float pc = 0.0f, lp = 0.0f;
            if (j < c->n_obs && (!c->mask || c->mask[r * c->n_obs + j])) {
                pc = clampf(row[j], c->eps, 1.0f);
                lp = (float)log((double)pc);
            }
            c->out16[r * OBS_PAD + j] = pc;
            if (j < c->n_obs) h = fmaf(pc, lp, h);
        }
        c->entropy[r] = h;
    }
}

void radar_oracle_prepare_queries(const float* probs, const uint8_t* mask, int64_t q, int n_obs,
                                  float eps, int normalize, float* p16, float* entropy) {
    prep_ctx c = {probs, mask, n_obs, eps, normalize, p16, entropy};
    parallel_for(q, 4096, prep_query_range, &c);
}

/* ranking key, larger is better, for one query against a block of ROWBLK corpus rows.
 * ct  : transposed embedding block  [d][ROWBLK]   (NULL when mode == KL)
 * lt  : transposed log-table block  [16][ROWBLK]  (NULL when mode == DPR)            */
static inline void score_block(int mode, const float* eq, const float* p16, float h, const float* ct,
                               const float* lt, int d, float alpha, float oma, float* api_score) {
    float ip[ROWBLK], x[ROWBLK];
    for (int r = 0; r < ROWBLK; ++r) { ip[r] = 0.0f; x[r] = 0.0f; }
    if (mode != MODE_KL) {
        for (int t = 0; t < d; ++t) {
            const float a = eq[t];
            const float* c = ct + (size_t)t * ROWBLK;
            for (int r = 0; r < ROWBLK; ++r) ip[r] = fmaf(a, c[r], ip[r]);
        }
    }
    if (mode != MODE_DPR) {
        for (int j = 0; j < NUM_OBS; ++j) {
            const float a = p16[j];
            const float* c = lt + (size_t)j * ROWBLK;
            for (int r = 0; r < ROWBLK; ++r) x[r] = fmaf(a, c[r], x[r]);
        }
    }
    for (int r = 0; r < ROWBLK; ++r) {
        if (mode == MODE_DPR) api_score[r] = ip[r];
        else {
            const float kl = h - x[r];
            if (mode == MODE_KL) api_score[r] = kl;
            else {
                const float t = oma * kl;
                api_score[r] = fmaf(alpha, ip[r], -t);
            }
        }
    }
}

/* (score, id) "a is better than b" under the tie rule */
static inline int better(int descending, float sa, int64_t ia, float sb, int64_t ib) {
    if (sa != sb) return descending ? (sa > sb) : (sa < sb);
    return ia < ib;
}

/* insert into a sorted (best first) list of length *cnt <= k */
static inline void topk_insert(int descending, int k, float* ls, int64_t* li, int* cnt, float s,
                               int64_t id) {
    int c = *cnt;
    if (c == k && !better(descending, s, id, ls[k - 1], li[k - 1])) return;
    int pos = c < k ? c : k - 1;
    while (pos > 0 && better(descending, s, id, ls[pos - 1], li[pos - 1])) {
        ls[pos] = ls[pos - 1];
        li[pos] = li[pos - 1];
        --pos;
    }
    ls[pos] = s;
    li[pos] = id;
    if (c < k) *cnt = c + 1;
}

typedef struct {
    int mode, d, k;
    const float *q_emb, *p16, *entropy, *c_emb, *logq16;
    float *ct, *lt;
    int64_t nq, n, nblk, idx_offset;
    float alpha, oma;
    float* out_scores;
    int64_t* out_idx;
    const int64_t* ids;
} search_ctx;

static void transpose_emb_range(int64_t b0, int64_t b1, void* vc) {
    search_ctx* c = (search_ctx*)vc;
    for (int64_t b = b0; b < b1; ++b)
        for (int r = 0; r < ROWBLK; ++r) {
            int64_t row = b * ROWBLK + r;
            if (row >= c->n) break;
            for (int t = 0; t < c->d; ++t)
                c->ct[((size_t)b * c->d + t) * ROWBLK + r] = c->c_emb[(size_t)row * c->d + t];
        }
}

static void transpose_logq_range(int64_t b0, int64_t b1, void* vc) {
    search_ctx* c = (search_ctx*)vc;
    for (int64_t b = b0; b < b1; ++b)
        for (int r = 0; r < ROWBLK; ++r) {
            int64_t row = b * ROWBLK + r;
            if (row >= c->n) break;
            for (int j = 0; j < OBS_PAD; ++j)
                c->lt[((size_t)b * OBS_PAD + j) * ROWBLK + r] = c->logq16[(size_t)row * OBS_PAD + j];
        }
}

static void search_range(int64_t i0, int64_t i1, void* vc) {
    search_ctx* c = (search_ctx*)vc;
    const int descending = c->mode != MODE_KL;
    for (int64_t i = i0; i < i1; ++i) {
        float* ls = c->out_scores + (size_t)i * c->k;
        int64_t* li = c->out_idx + (size_t)i * c->k;
        int cnt = 0;
        float sc[ROWBLK];
        for (int64_t b = 0; b < c->nblk; ++b) {
            score_block(c->mode, c->q_emb ? c->q_emb + (size_t)i * c->d : NULL,
                        c->p16 ? c->p16 + (size_t)i * OBS_PAD : NULL, c->entropy ? c->entropy[i] : 0.0f,
                        c->ct ? c->ct + (size_t)b * c->d * ROWBLK : NULL,
                        c->lt ? c->lt + (size_t)b * OBS_PAD * ROWBLK : NULL, c->d, c->alpha, c->oma, sc);
            for (int r = 0; r < ROWBLK; ++r) {
                int64_t row = b * ROWBLK + r;
                if (row >= c->n) break;
                topk_insert(descending, c->k, ls, li, &cnt, sc[r], row);
            }
        }
        for (int j = 0; j < c->k; ++j) li[j] += c->idx_offset;
    }
}

/*
 * Brute-force canonical search.  q_emb [Q,d] / c_emb [N,d] may be NULL in KL mode; p16/entropy/logq16
 * may be NULL in DPR mode.  out_scores [Q,k] float32 (API sign), out_idx [Q,k] int64 (+ idx_offset).
 * k must be <= N.  Returns 0 on success.
 */
int radar_oracle_search(int mode, const float* q_emb, const float* p16, const float* entropy,
                        const float* c_emb, const float* logq16, int64_t nq, int64_t n, int d, int k,
                        float alpha, int64_t idx_offset, float* out_scores, int64_t* out_idx) {
    if (k <= 0 || k > n || nq < 0) return 1;
    search_ctx c;
    memset(&c, 0, sizeof c);
    c.mode = mode; c.d = d; c.k = k; c.q_emb = q_emb; c.p16 = p16; c.entropy = entropy;
    c.c_emb = c_emb; c.logq16 = logq16; c.nq = nq; c.n = n; c.nblk = (n + ROWBLK - 1) / ROWBLK;
    c.idx_offset = idx_offset; c.alpha = alpha; c.oma = 1.0f - alpha;
    c.out_scores = out_scores; c.out_idx = out_idx;
    /* transposed, zero-padded copies of the corpus so the inner loops vectorise across rows while each
     * row's fma chain keeps its canonical order */
    if (mode != MODE_KL) {
        c.ct = (float*)calloc((size_t)c.nblk * d * ROWBLK, sizeof(float));
        if (!c.ct) return 2;
        parallel_for(c.nblk, 64, transpose_emb_range, &c);
    }
    if (mode != MODE_DPR) {
        c.lt = (float*)calloc((size_t)c.nblk * OBS_PAD * ROWBLK, sizeof(float));
        if (!c.lt) { free(c.ct); return 2; }
        parallel_for(c.nblk, 256, transpose_logq_range, &c);
    }
    parallel_for(nq, 1, search_range, &c);
    free(c.ct);
    free(c.lt);
    return 0;
}

static void pairs_range(int64_t i0, int64_t i1, void* vc) {
    search_ctx* c = (search_ctx*)vc;
    for (int64_t i = i0; i < i1; ++i)
        for (int j = 0; j < c->k; ++j) {
            int64_t row = c->ids[(size_t)i * c->k + j];
            float ip = 0.0f, x = 0.0f, s;
            if (c->mode != MODE_KL)
                for (int t = 0; t < c->d; ++t)
                    ip = fmaf(c->q_emb[(size_t)i * c->d + t], c->c_emb[(size_t)row * c->d + t], ip);
            if (c->mode != MODE_DPR)
                for (int t = 0; t < NUM_OBS; ++t)
                    x = fmaf(c->p16[(size_t)i * OBS_PAD + t], c->logq16[(size_t)row * OBS_PAD + t], x);
            if (c->mode == MODE_DPR) s = ip;
            else {
                float kl = c->entropy[i] - x;
                if (c->mode == MODE_KL) s = kl;
                else { float t2 = c->oma * kl; s = fmaf(c->alpha, ip, -t2); }
            }
            c->out_scores[(size_t)i * c->k + j] = s;
        }
}

/* canonical scores of explicit (query, case) pairs: out[i][j] for ids[i][j] (ids are LOCAL row numbers).
 * Used by tests to check returned scores without a full scan. */
int radar_oracle_score_pairs(int mode, const float* q_emb, const float* p16, const float* entropy,
                             const float* c_emb, const float* logq16, int64_t nq, int d, int k,
                             float alpha, const int64_t* ids, float* out) {
    search_ctx c;
    memset(&c, 0, sizeof c);
    c.mode = mode; c.d = d; c.k = k; c.q_emb = q_emb; c.p16 = p16; c.entropy = entropy;
    c.c_emb = c_emb; c.logq16 = logq16; c.alpha = alpha; c.oma = 1.0f - alpha;
    c.out_scores = out; c.ids = ids;
    parallel_for(nq, 16, pairs_range, &c);
    return 0;
}

/* top-k of the union of `parts` per-shard lists ([parts][Q][k_in] scores + int64 ids; id < 0 = padding) */
int radar_oracle_merge_topk(const float* scores, const int64_t* idx, int64_t nq, int parts, int k_in,
                            int k_out, int ascending, float* out_scores, int64_t* out_idx) {
    if (k_out > parts * k_in) return 1;
    for (int64_t i = 0; i < nq; ++i) {
        int cnt = 0;
        for (int p = 0; p < parts; ++p)
            for (int j = 0; j < k_in; ++j) {
                size_t o = ((size_t)p * nq + i) * k_in + j;
                if (idx[o] < 0) continue;
                topk_insert(!ascending, k_out, out_scores + (size_t)i * k_out, out_idx + (size_t)i * k_out,
                            &cnt, scores[o], idx[o]);
            }
        for (int j = cnt; j < k_out; ++j) {
            out_scores[(size_t)i * k_out + j] = ascending ? INFINITY : -INFINITY;
            out_idx[(size_t)i * k_out + j] = -1;
        }
    }
    return 0;
}
