/*
 * radar_retrieval.h -- C ABI of libradar_retrieval.so: the B200 (sm_100a) replacement for the
 * scoring + top-k operator underneath RADAR's case-retrieval API.
 *
 * The reference (MOsama10/radar-multimodal-radiology, 100 % Python) has no native boundary; the
 * operator it calls is third-party faiss:
 *     faiss.IndexFlatIP(d)                 annotate_retrieve/modeling_dense_passage_retrieval.py:297
 *     index.add(float32[N,d])              annotate_retrieve/modeling_dense_passage_retrieval.py:298
 *     index.ntotal                         annotate_retrieve/modeling_dense_passage_retrieval.py:300
 *     index.search(float32[nq,d], k)       annotate_retrieve/modeling_dense_passage_retrieval.py:313
 * plus the pieces north_star names that the reference only declares:
 *     RetrievalConfig.hybrid_alpha         annotate_retrieve/modeling_dense_passage_retrieval.py:187  (never read)
 *     KL observation retrieval             src/knowledge/__init__.py (0 bytes), README.md:64
 *     re-retrieval rounds                  annotate_retrieve/modeling_iterative_rag.py:236-237
 *     overlap re-rank                      annotate_retrieve/modeling_iterative_rag.py:127-152
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no exceptions cross the boundary.
 *   - every function returns 0 on success or a RADAR_E_* code; radar_last_error() returns a
 *     thread-local message for the last failure.
 *   - ALL data pointers are DEVICE pointers on the current CUDA device unless a parameter is
 *     documented "host".  The caller (PyTorch) owns every buffer; the library allocates nothing
 *     that outlives a call and keeps no pointer after returning.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it.
 *   - row-major, contiguous, 16-byte aligned tensors.
 *   - there is NO CPU fallback: on a machine without an sm_100 device the compute entry points
 *     return RADAR_E_CUDA / RADAR_E_ARCH.
 */
#ifndef RADAR_RETRIEVAL_H_
#define RADAR_RETRIEVAL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RADAR_ABI_VERSION 4

#define RADAR_NUM_OBS 14 /* CheXpert-14, train_expert_models.py:50-65 */
#define RADAR_OBS_PAD 16 /* K=16 padded contraction */
#define RADAR_KLPACK 32  /* bf16 [log q hi (16) | log q lo (16)] per corpus row = 64 B */
#define RADAR_MAX_K 128  /* largest top-k per radar_search call; larger k: page with radar_queries.after_* */

enum radar_status {
    RADAR_OK = 0,
    RADAR_E_ARG = 1,       /* invalid argument (message says which) */
    RADAR_E_CUDA = 2,      /* CUDA runtime / driver error */
    RADAR_E_ARCH = 3,      /* device is not sm_100 (tcgen05 / TMA path unavailable) */
    RADAR_E_WORKSPACE = 4, /* workspace too small */
    RADAR_E_UNCERTIFIED = 5 /* internal: certificate failed and fallback disabled */
};

enum radar_mode {
    RADAR_MODE_DPR = 0,   /* inner product, descending   (faiss.IndexFlatIP.search) */
    RADAR_MODE_KL = 1,    /* KL(p_query || q_case), ascending */
    RADAR_MODE_HYBRID = 2 /* alpha*ip - (1-alpha)*KL, descending */
};

enum radar_precision {
    RADAR_PREC_BF16 = 0, /* bf16 tensor-core filter, k' over-fetched candidates re-scored in canonical fp32 */
    RADAR_PREC_FP32 = 1  /* result identical to the canonical fp32 definition (certified filter or exact scan) */
};

enum radar_algo {
    RADAR_ALGO_AUTO = 0,
    RADAR_ALGO_SIMT_EXACT = 1, /* CUDA-core exact scan with canonical arithmetic (always exact) */
    RADAR_ALGO_TC_FILTER = 2,  /* tcgen05 bf16 filter + canonical re-score (+ certificate in FP32 mode) */
    RADAR_ALGO_KL_STREAM = 3   /* KL, <= 256 queries, >= 65 536 cases: tcgen05 stream with pooled candidates (HBM-bound) */
};

enum radar_kl_variant {
    RADAR_KL_AUTO = 0,
    RADAR_KL_BF16X3 = 1, /* bf16 hi/lo split of both operands, three products, table klpack (64 B per case) */
    RADAR_KL_F16X1 = 2,  /* fp16, one product, table kl16 (32 B per case) */
    RADAR_KL_F16X2 = 3   /* fp16, query split hi/lo (two products), table kl16 */
};

/* Corpus shard resident in HBM.  Pointers a mode does not need may be NULL.
 * Built once per index by radar_pack_embeddings / radar_kl_prepare_corpus
 * (replaces faiss.IndexFlatIP.add, modeling_dense_passage_retrieval.py:298). */
typedef struct radar_corpus {
    int64_t n;                /* rows in this shard */
    int32_t d;                /* embedding dim (multiple of 64, <= 512 for the tensor-core path) */
    int32_t reserved0;
    const float* emb_f32;     /* [n,d]   canonical embeddings (re-score / exact scan) */
    const uint16_t* emb_bf16; /* [n,d]   bf16 RN copy (tensor-core filter); NULL => SIMT only */
    const float* logq16;      /* [n,16]  log of clamped probabilities, cols 14,15 = 0 */
    const uint16_t* klpack;   /* [n,32]  bf16 hi|lo split of logq16 (tensor-core filter); NULL => SIMT only */
    float emb_max_norm;       /* max_n ||emb_f32[n]||_2 (host value; from radar_pack_embeddings) */
    float logq_max_abs;       /* max |logq16| (host value; <= |log eps|) */
    int64_t idx_offset;       /* added to every returned id (global id of row 0 of this shard) */
    float logq_col_max[RADAR_OBS_PAD]; /* per observation j: max_n |logq16[n][j]| (host values); all zero => logq_max_abs
                                          is used for every column.  Only tightens the filter's error bound (fewer exact
                                          re-runs in FP32 mode, tighter initial thresholds on the KL stream path). */
    const uint16_t* kl16;     /* [n,16]  fp16 RN of 2048 * logq16 (32 B per case): the table the KL-only tensor-core paths
                                 stream (radar_kl_prepare_corpus); NULL => they use klpack */
} radar_corpus_t;

/* Query batch.  Pointers a mode does not need may be NULL. */
typedef struct radar_queries {
    int64_t q;            /* number of queries */
    const float* emb_f32; /* [q,d]   */
    const float* p16;     /* [q,16]  masked, clamped probabilities (radar_kl_prepare_queries) */
    const float* entropy; /* [q]     sum_j p16_j log p16_j (radar_kl_prepare_queries) */
    /* "search after" (both NULL: off).  When given, query i only returns cases that rank STRICTLY AFTER the pair
     * (after_scores[i], after_idx[i]) -- a score/id pair as a previous radar_search returned it (API sign, global
     * id) -- under the (score, id) order; after_idx[i] < 0 switches the bound off for that query.  This pages
     * through a ranking: IndexFlatIP.search accepts any k (dpr.py:313, retrieve_with_hard_negatives dpr.py:326
     * asks for k + num_negatives), so k > RADAR_MAX_K is served as ceil(k / RADAR_MAX_K) calls.  Exact scan only
     * (RADAR_ALGO_AUTO selects it; other algorithms return RADAR_E_ARG). */
    const float* after_scores; /* [q] */
    const int64_t* after_idx;  /* [q] */
} radar_queries_t;

typedef struct radar_search_params {
    int32_t mode;      /* enum radar_mode */
    int32_t precision; /* enum radar_precision */
    int32_t algo;      /* enum radar_algo */
    int32_t k;         /* 1..RADAR_MAX_K, k <= corpus.n */
    float alpha;       /* hybrid weight (RetrievalConfig.hybrid_alpha, dpr.py:187); ignored unless HYBRID */
    int32_t overfetch; /* candidates re-scored per query on the filter path; 0 = automatic */
    int32_t num_sms;   /* 0 = all SMs of the device (tests use small values to force multi-part merges) */
    int32_t kl_variant; /* enum radar_kl_variant: filter arithmetic of the KL-only tensor-core paths; 0 = automatic */
} radar_search_params_t;

/* Per-call statistics written to HOST memory when the pointer is non-NULL (forces a stream sync). */
typedef struct radar_search_stats {
    int32_t algo_used;       /* enum radar_algo actually run for the main pass */
    int32_t kernel_launches; /* kernels this call enqueued */
    int64_t uncertified;     /* FP32/TC_FILTER: queries whose certificate failed and were re-run exactly */
    int32_t parts;           /* corpus slabs per query tile */
    int32_t kprime;          /* candidates kept per query by the filter */
    float filter_sm_mhz;     /* average SM clock during the tensor-core filter kernel (clock64 / globaltimer), 0 if not run */
    int32_t reserved;
} radar_search_stats_t;

const char* radar_last_error(void);
int radar_abi_version(void);

/* number of SMs / compute capability of the current device (host outputs) */
int radar_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* make `device` current for this thread inside the library's CUDA runtime (call before any other entry
 * point when the process drives more than one GPU; one process per GPU under torchrun needs it once). */
int radar_set_device(int device);
/* the device that is current for this thread in the library's CUDA runtime (host output); callers that switch
 * devices around a call use it to restore the previous one */
int radar_get_device(int* device);

/* measurement aid: while enabled, every radar_search call of this thread records a pair of CUDA events
 * (owned by the library) immediately before / after its dominant kernel (exact scan or tensor-core
 * filter) on the call's stream; radar_profile_kernel_ms waits for the last pair and returns the elapsed
 * milliseconds (host output), so a benchmark can time that kernel alone without a profiler. */
int radar_profile_enable(int enable);
int radar_profile_kernel_ms(float* ms_out);

/* ---- index build (replaces IndexFlatIP.add, dpr.py:298; K1 corpus side) ------------------------- */

/* fp32 [n,d] -> bf16 RN [n,d]; also atomically max-reduces the row L2 norms into *max_norm (device
 * float, caller zero-initialises).  emb_bf16 may be NULL (norm only). */
int radar_pack_embeddings(const float* emb_f32, int64_t n, int d, uint16_t* emb_bf16, float* max_norm,
                          void* stream);

/* probabilities [n,n_obs] (n_obs <= 16, normally 14) -> logq16 [n,16]; klpack [n,32] and kl16 [n,16] may be NULL.
 * q <- clamp(q, eps, 1); if normalize, rows are divided by their fp32 left-to-right sum first. */
int radar_kl_prepare_corpus(const float* probs, int64_t n, int n_obs, float eps, int normalize,
                            float* logq16, uint16_t* klpack, uint16_t* kl16, void* stream);

/* ---- query preparation (K1/K3 query side) ------------------------------------------------------- */

/* probabilities [q,n_obs] + optional mask uint8 [q,n_obs] (0 = observation masked out) ->
 * p16 [q,16], entropy [q]. */
int radar_kl_prepare_queries(const float* probs, const uint8_t* mask, int64_t q, int n_obs, float eps,
                             int normalize, float* p16, float* entropy, void* stream);

/* ---- search (replaces IndexFlatIP.search, dpr.py:313; K1, K2, K3) -------------------------------- */

size_t radar_search_workspace_bytes(const radar_corpus_t* corpus, int64_t q,
                                    const radar_search_params_t* params);

/* out_scores [q,k] float32 in the API's sign (DPR: ip desc; KL: KL asc; hybrid: fused desc),
 * out_idx [q,k] int64 (+ corpus->idx_offset).  workspace: device memory of at least
 * radar_search_workspace_bytes(...).  stats: optional HOST pointer (NULL: the call only enqueues work and never
 * synchronises -- it can be captured in a CUDA graph; queries whose FP32 certificate fails, or whose pooled buffer
 * overflowed on the KL stream path, are re-run by the exact scan with a device-side count, without a host round trip).
 * out_packed (nullable) [q,k] uint64: the same result as one sortable word per entry,
 *   (orderable bits of the ranking key << 32) | (0xFFFFFFFF - global id), 0 = padding; larger word = better rank.
 *   It is what row-sharded ranks exchange (ONE all-gather of 8 B per entry) before radar_merge_packed.
 * RADAR_ALGO_AUTO picks RADAR_ALGO_KL_STREAM for KL with <= 256 queries over >= 65 536 cases, otherwise the tcgen05
 * filter when the corpus carries the bf16 tables, otherwise the exact scan. */
int radar_search(const radar_corpus_t* corpus, const radar_queries_t* queries,
                 const radar_search_params_t* params, float* out_scores, int64_t* out_idx,
                 uint64_t* out_packed, void* workspace, size_t workspace_bytes, radar_search_stats_t* stats,
                 void* stream);

#ifdef RADAR_DEBUG
/* bring-up / test aid, exported ONLY by the RADAR_DEBUG flavour of the library (libradar_retrieval_dbg.so; the
 * release library neither declares nor exports it and reads no environment variables): dense dump of the
 * tensor-core filter keys (canonical-key units) for every (query, case) pair into out_keys [q,n] -- only sensible
 * for small q*n.  Same workspace as radar_search with algo = RADAR_ALGO_TC_FILTER. */
int radar_debug_filter_keys(const radar_corpus_t* corpus, const radar_queries_t* queries,
                            const radar_search_params_t* params, float* out_keys, void* workspace,
                            size_t workspace_bytes, void* stream);
#endif

/* ---- shard merge (no reference counterpart) --------------------------------- */

/* cand_scores / cand_idx: [parts, q, k_in] (the layout ncclAllGather produces); idx < 0 = padding.
 * Writes the best k_out (<= min(parts*k_in, 1024)) per query under (score, id) order. */
int radar_merge_topk(const float* cand_scores, const int64_t* cand_idx, int64_t q, int parts, int k_in,
                     int k_out, int ascending, float* out_scores, int64_t* out_idx, void* stream);

/* cand: [parts, q, k_in] packed words as radar_search's out_packed writes them (0 = padding).  Writes the best
 * k_out (<= parts*k_in <= 2048) per query in API form: scores in the mode's sign (mode = enum radar_mode), ids. */
int radar_merge_packed(const uint64_t* cand, int64_t q, int parts, int k_in, int k_out, int mode,
                       float* out_scores, int64_t* out_idx, void* stream);

/* ---- iterative-RAG re-rank (replaces TargetedRetriever.rank_retrieved_passages, rag.py:127-152) - */

/* case_bits uint16 [q,k]: 14-bit observation set of every retrieved case; missing_bits uint16 [q].
 * score = overlap/(m+1e-8) + 0.2*min(overlap/max(m,1),1) in float64 (Python float), 0.5 when m == 0;
 * out_scores [q,k] float64 in the ORIGINAL passage order; order [q,k] int32 = stable descending argsort
 * of the scores (Python's list.sort(reverse=True) semantics). */
int radar_rerank_overlap(const uint16_t* case_bits, const uint16_t* missing_bits, int64_t q, int k,
                         double* out_scores, int32_t* out_order, void* stream);

/* gather uint16 bits of retrieved ids: out[q,k] = table[idx[q,k] - idx_offset] (0 for idx < 0) */
int radar_gather_bits(const uint16_t* table, int64_t n, const int64_t* idx, int64_t q, int k,
                      int64_t idx_offset, uint16_t* out, void* stream);

/* ---- query-side prologue (next row f1: Linear(768->512) + L2 normalise, dpr.py:202-203, :246) ---- */

/* y[b,:] = normalize(x[b,:] @ W^T + bias), x [b,in], W [out,in] (nn.Linear layout), bias [out] or NULL;
 * fp32 in / fp32 out, eps = 1e-12 as torch.nn.functional.normalize. */
int radar_project_normalize(const float* x, const float* w, const float* bias, int64_t b, int in_dim,
                            int out_dim, float* y, void* stream);

/* The same operator on the tensor pipe (tcgen05, bf16 hi/lo split of both operands, three products, fp32 accumulation:
 * components within 1e-5 of the fp32 result): needs out_dim == 512 and in_dim % 32 == 0 (BiomedCLIP: 768 -> 512) and
 * a scratch buffer of radar_project_workspace_bytes(b, in_dim, out_dim) bytes (0 = shape not supported: use
 * radar_project_normalize).  Outputs (each nullable, not both): y fp32 [b,512]; y_bf16 [b,512] = bf16 RN of the
 * normalised rows, i.e. the A-operand rows the DPR filter would otherwise pack from y. */
size_t radar_project_workspace_bytes(int64_t b, int in_dim, int out_dim);
int radar_project_normalize_tc(const float* x, const float* w, const float* bias, int64_t b, int in_dim,
                               int out_dim, float* y, uint16_t* y_bf16, void* workspace, size_t workspace_bytes,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADAR_RETRIEVAL_H_ */
