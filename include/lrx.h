/*
 * lrx.h -- C ABI of the B200-native hybrid-retrieval hot path.
 *
 * Drop-in boundary for MET4L-DS/Legal-RAG-engine's `RetrievalEngine`
 * (reference: src/retrieval/retrieval_engine.py:23-96, create_vector_store.py:14-83,
 * src/retrieval/orchestrator.py:38-62).  The reference has no FFI of its own: every
 * FLOP of this path runs inside three PyPI wheels.  Each entry point below names the
 * reference call site it replaces; INTEGRATION.md shows the ctypes stub a maintainer
 * would add to the reference.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative LRX_E* code;
 *    lrx_last_error() returns the message of the last failure (handle-owned storage,
 *    or a process-wide slot when the handle is NULL).
 *  - one handle per GPU rank.  The handle owns its workspaces and the CUDA stream
 *    binding; it never frees caller memory.  Every entry point calls cudaSetDevice,
 *    so a handle may be created on one host thread and used (sequentially) from
 *    another (reference: src/server/app.py:65-70 vs :110-120).
 *  - pointers named dev_* are device pointers on the handle's GPU; host_* are host
 *    pointers (pinned or pageable).  Nothing here falls back to the CPU: a missing
 *    GPU or a non-sm_100 device is LRX_E_DEVICE at lrx_open.
 *  - ids are GLOBAL chunk ids (int64) = id_base + local row; -1 pads short lists.
 *  - all kernels run on the stream set with lrx_set_stream (default: the legacy
 *    default stream), so PyTorch code can order them with its own work and with
 *    the NCCL all-gather between lrx_search_local and lrx_search_finish.
 */
#ifndef LRX_H
#define LRX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRX_DIM 384                 /* all-MiniLM-L6-v2 embedding width */
#define LRX_MAX_BATCH 64            /* sub-queries per lrx_search_* call */
#define LRX_MAX_DEPTH 256           /* max candidate depth K (= 2k) per list */
#define LRX_MAX_QUERY_TERMS 64      /* default token capacity per sub-query of the device-pointer entry
                                     * points (raise it with lrx_set_query_capacity); the host-buffer
                                     * entry points size the capacity from the batch: no limit */
#define LRX_MAX_WORLD 16            /* shards (GPUs of one box) in a peer exchange */
#define LRX_IPC_HANDLE_BYTES 64     /* sizeof(cudaIpcMemHandle_t) */

enum {
    LRX_OK = 0,
    LRX_E_ARG = -1,        /* bad argument */
    LRX_E_DEVICE = -2,     /* no CUDA device / not sm_100 */
    LRX_E_CUDA = -3,       /* CUDA runtime error (message has the detail) */
    LRX_E_STATE = -4,      /* corpus / postings / weights not set */
    LRX_E_AMBIGUOUS = -5,  /* exactness guard could not be met */
    LRX_E_NOMEM = -6,
    LRX_E_PEER = -7        /* a peer shard did not publish its block in time */
};

enum { LRX_FUSE_LINEAR = 0, LRX_FUSE_RRF = 1 };

typedef struct lrx_handle lrx_handle;

typedef struct lrx_config {
    int32_t device;        /* CUDA ordinal */
    int32_t dim;           /* must be LRX_DIM */
    int32_t rank;          /* shard index of this handle (0 when unsharded) */
    int32_t world;         /* number of shards (1 when unsharded) */
} lrx_config;

/* One candidate of one sub-query on one shard; the unit the all-gather moves
 * (SURVEY.md section 8e).  24 bytes. */
typedef struct lrx_record {
    int64_t id;            /* global chunk id, -1 = empty slot */
    double dense;          /* exact inner product (see oracle/flat_ip.py)   */
    double bm25;           /* exact BM25Okapi score of that chunk            */
} lrx_record;

/* ---- lifetime ---------------------------------------------------------- */
int lrx_open(const lrx_config* cfg, lrx_handle** out);
int lrx_close(lrx_handle* h);
const char* lrx_last_error(const lrx_handle* h);
int lrx_set_stream(lrx_handle* h, void* cuda_stream);
/* ABI / build identification: "lrx <version> sm_100a". */
const char* lrx_version(void);

/* ---- index residency (replaces faiss.read_index + pickle.load(bm25.pkl),
 *      retrieval_engine.py:36-47; the data stay owned by the caller) ------- */
/* Row-major fp16 [n_local, 384] chunk matrix (768 B/row, 16-byte aligned). */
int lrx_set_corpus(lrx_handle* h, const void* dev_x_fp16, int64_t n_local, int64_t id_base,
                   int32_t dim);
/* Optional int8 pre-filter of K2 (up to 4 queries, K <= 64): the scan streams an int8 SHADOW of
 * the matrix -- per-row scaled int8 rows, 388 B/row instead of 768 -- and the exact float64
 * re-score of the merged candidate list reads the fp16 rows; the exactness guard uses the
 * rigorous error bound of the shadow, so results stay bit-identical to the plain scan (a query
 * whose candidates the bound cannot separate falls back to the fp16 scan through the widening
 * retry).  The shadow lives in caller-owned device memory of lrx_dense_prefilter_bytes(n_local)
 * bytes, 16-byte aligned.  lrx_build_dense_prefilter quantises the matrix set by lrx_set_corpus
 * into it (synchronous), returns host_bounds_out[2] = {max row error norm, max row norm} and
 * enables the pre-filter; lrx_set_dense_prefilter attaches a shadow built through another handle
 * over the same matrix (dev_buf NULL: back to the plain fp16 scan).  lrx_set_corpus detaches it.
 * (No counterpart in the reference: faiss.IndexFlatIP scans fp32 rows, retrieval_engine.py:64.) */
int64_t lrx_dense_prefilter_bytes(int64_t n_local);
int lrx_build_dense_prefilter(lrx_handle* h, void* dev_buf, int64_t bytes, double* host_bounds_out);
int lrx_set_dense_prefilter(lrx_handle* h, const void* dev_buf, int64_t bytes,
                            const double* host_bounds);
/* Term-major CSR postings restricted to this shard's documents, doc ids LOCAL and
 * ascending within a term.  dev_postings: nnz 8-byte entries {u32 doc_local, u16 tf,
 * u16 doc_len} (lrx_bm25_build_postings fills them from (doc, tf) pairs and the document
 * lengths), 16-byte aligned and readable up to nnz rounded up to an even count (the scan
 * moves 16-byte units).  The query-independent BM25Okapi factor
 *   impact(tf, len) = tf*(k1+1) / (tf + k1*(1 - b + b*len/avgdl))          (float64)
 * is recomputed by the scan from (tf, len) and the GLOBAL avgdl given here, with rank_bm25's
 * float64 operation order.  dev_idf: GLOBAL idf per term (rank_bm25 semantics: epsilon
 * floor applied).  max_doc_len <= 65535.  Replaces pickle.load(bm25.pkl). */
int lrx_set_postings(lrx_handle* h, const uint64_t* dev_term_ptr, const void* dev_postings,
                     const double* dev_idf, int64_t n_terms, int64_t nnz, double avgdl, double k1,
                     double b, int32_t max_doc_len);
/* Index build (create_vector_store.py:60-61, the BM25Okapi constructor's per-document
 * frequency tables): dev_doc_tf = nnz pairs {u32 doc_local, u32 tf} in CSR order ->
 * dev_postings_out (nnz x 8 bytes).  LRX_E_ARG if a tf or a length exceeds 65535. */
int lrx_bm25_build_postings(lrx_handle* h, const uint32_t* dev_doc_tf, int64_t nnz,
                            const uint32_t* dev_doc_len, void* dev_postings_out);

/* ---- K1: encoder (replaces SentenceTransformer("all-MiniLM-L6-v2").encode followed by
 *      faiss.normalize_L2; retrieval_engine.py:28,61-62, create_vector_store.py:33-34,45,51) --
 * Weights: DEVICE pointers to float32 tensors in the HuggingFace BertModel state_dict layout
 * (nn.Linear weight = [out, in]); the library packs its own fp16 copies, so the caller may
 * free them when lrx_set_encoder_weights returns.  Fixed architecture: hidden 384, 12 heads,
 * FFN 1536, 6 layers, LayerNorm eps 1e-12, exact-erf GELU, token type 0. */
#define LRX_BERT_LAYERS 6
typedef struct lrx_bert_layer {
    const float *wq, *bq, *wk, *bk, *wv, *bv;   /* attention.self.{query,key,value}  [384,384],[384] */
    const float *wo, *bo;                       /* attention.output.dense           [384,384],[384] */
    const float *ln1_g, *ln1_b;                 /* attention.output.LayerNorm       [384]           */
    const float *w1, *b1;                       /* intermediate.dense               [1536,384],[1536] */
    const float *w2, *b2;                       /* output.dense                     [384,1536],[384] */
    const float *ln2_g, *ln2_b;                 /* output.LayerNorm                 [384]           */
} lrx_bert_layer;
typedef struct lrx_bert_weights {
    int32_t vocab_size;        /* 30522 for all-MiniLM-L6-v2 */
    int32_t max_positions;     /* 512 */
    const float* word_emb;     /* embeddings.word_embeddings        [vocab, 384] */
    const float* pos_emb;      /* embeddings.position_embeddings    [max_positions, 384] */
    const float* type_emb;     /* embeddings.token_type_embeddings  [2, 384] (row 0 used) */
    const float *emb_ln_g, *emb_ln_b;
    lrx_bert_layer layers[LRX_BERT_LAYERS];
} lrx_bert_weights;
int lrx_set_encoder_weights(lrx_handle* h, const lrx_bert_weights* w);
/* dev_ids int32 [B,S] WordPiece ids ([CLS] ... [SEP], 0-padded), dev_lens int32 [B] = number of
 * attended tokens per sequence (the attention mask is a prefix mask, as the tokenizer emits).
 * Outputs (either may be NULL): float32 [B,384] and fp16 [B,384] unit vectors (the fp16 copy is
 * the query operand of lrx_dense_topk / lrx_search_*). */
int lrx_encode(lrx_handle* h, const int32_t* dev_ids, const int32_t* dev_lens, int32_t B, int32_t S,
               float* dev_out_f32, void* dev_out_f16);
/* Host-buffer form: H2D of ids/lens, the forward pass, D2H of the embeddings, synchronised. */
int lrx_encode_host(lrx_handle* h, const int32_t* host_ids, const int32_t* host_lens, int32_t B,
                    int32_t S, float* host_out_f32);
/* The tensor-core GEMM stage on its own: out[M,N] = epi(A[M,K] * W[N,K]^T), fp16 row-major
 * operands, fp32 accumulation.  epi: 0 +bias -> fp16, 1 +bias, GELU(erf) -> fp16,
 * 2 +bias +residual, LayerNorm (N == 384) -> fp16, 3 raw accumulators -> float32.
 * K % 64 == 0, N % 128 == 0. */
int lrx_gemm_f16(lrx_handle* h, const void* dev_a, const void* dev_w, int32_t M, int32_t N, int32_t K,
                 int32_t epi, const float* dev_bias, const void* dev_residual,
                 const float* dev_gamma, const float* dev_beta, float eps, void* dev_out);

/* ---- stage kernels (device pointers) ----------------------------------- */
/* K2: replaces IndexFlatIP.search(x, K)  (retrieval_engine.py:64).
 * dev_q_fp16 [B,384]; outputs [B,K]: exact float64 score, float32 D, int64 I,
 * best first, (score desc, id asc); pads (-inf, -FLT_MAX, -1).
 * dev_flags [B]: 0 = exact, 1 = exactness guard failed (caller should retry with
 * lrx_dense_topk_ex and a larger width). */
int lrx_dense_topk(lrx_handle* h, const void* dev_q_fp16, int32_t B, int32_t K,
                   double* dev_exact, float* dev_D, int64_t* dev_I, int32_t* dev_flags);
int lrx_dense_topk_ex(lrx_handle* h, const void* dev_q_fp16, int32_t B, int32_t K,
                      int32_t width, double* dev_exact, float* dev_D, int64_t* dev_I,
                      int32_t* dev_flags);
/* K2b: the same search for LARGE batches (B up to 4096; FAISS's BLAS regime, nq >= 20) on the
 * tensor cores: two tcgen05 GEMM passes (tile maxima -> per-query threshold; candidates above
 * it) + exact float64 re-score.  Same outputs and ordering as lrx_dense_topk.  stride: pass 1
 * samples every stride-th 256-row tile (0 = automatic).  dev_flags[b] = 1: more than 1024
 * candidates survived for query b -- rerun with stride 1. */
int lrx_dense_topk_batched(lrx_handle* h, const void* dev_q_fp16, int32_t B, int32_t K, int32_t stride,
                           double* dev_exact, float* dev_D, int64_t* dev_I, int32_t* dev_flags);
/* Exact inner products at given global ids (ids outside the shard or -1 -> -inf). */
int lrx_dense_at(lrx_handle* h, const void* dev_q_fp16, int32_t B, const int64_t* dev_ids,
                 int32_t n, double* dev_out);

/* Token capacity of a query batch for the DEVICE-pointer entry points (lrx_bm25, lrx_search_local*,
 * lrx_search_sharded): the library cannot see dev_q_ptr[B] from the host, so the rows of its
 * per-token tables are sized from this figure; 0 (default) = B * LRX_MAX_QUERY_TERMS.  Every token
 * is scored, in order, as the reference does (retrieval_engine.py:67-68 has no limit).  A batch with
 * MORE tokens than the capacity is never scored short silently: lrx_bm25 returns NaN maxima and the
 * search entry points set bit 2 of the status words.  The host-buffer entry points size the
 * capacity from host_q_ptr themselves. */
int lrx_set_query_capacity(lrx_handle* h, int32_t max_total_terms);

/* K3: replaces BM25Okapi.get_scores + max()  (retrieval_engine.py:68,74).
 * dev_q_terms: concatenated term ids (-1 = out of vocabulary), dev_q_ptr [B+1].
 * Emits (i) exact scores at dev_cand_ids [B,n_cand] (may be NULL / n_cand 0; ids outside
 * the shard or -1 give 0), (ii) dev_max [B]: max over the shard's POSITIVE scores, 0 when
 * none, (iii) the shard's top-K positive-score list [B,K] (score desc, id asc; pads
 * (0,-1)); K may be 0. */
int lrx_bm25(lrx_handle* h, const int32_t* dev_q_terms, const int32_t* dev_q_ptr, int32_t B,
             const int64_t* dev_cand_ids, int32_t n_cand, double* dev_cand_scores,
             double* dev_max, int32_t K, double* dev_top_scores, int64_t* dev_top_ids);

/* ---- K5: the whole of RetrievalEngine.search for a batch of sub-queries ---
 * lrx_search_local: K2 + K3 on this shard.  Writes, per sub-query b, the shard's
 *   records: dev_records [B][2][K]  ([b][0] = dense list, [b][1] = BM25 list, the
 *   latter all-empty in linear mode) and dev_maxbm25 [B].
 * (all-gather dev_records / dev_maxbm25 across shards -- NCCL, done by the caller)
 * lrx_search_finish: K4.  dev_records_all [world][B][2][K], dev_max_all [world][B];
 *   outputs [B,k]: ids (-1 pad), fused score, semantic, keyword; dev_status [B] is the
 *   OR of the exactness flags (non-zero: rerun with a larger width). */
int lrx_search_local(lrx_handle* h, const void* dev_q_fp16, const int32_t* dev_q_terms,
                     const int32_t* dev_q_ptr, int32_t B, int32_t k, int32_t mode,
                     int32_t width, lrx_record* dev_records, double* dev_maxbm25,
                     int32_t* dev_flags);
int lrx_search_finish(lrx_handle* h, const lrx_record* dev_records_all,
                      const double* dev_max_all, const int32_t* dev_flags_all, int32_t world,
                      int32_t B, int32_t k, int32_t mode, const double* dev_weights,
                      int64_t* dev_ids, double* dev_score, double* dev_sem, double* dev_kw,
                      int32_t* dev_status);

/* Packed form: the shard's records, maxima and flags in ONE contiguous block of
 * lrx_packed_bytes(B,k) bytes ([B][2][2k] records | [B] double | [B] int32, 16-byte
 * padded), so the exchange is a single all-gather per query batch; the gathered buffer
 * [world][packed] goes straight into lrx_search_finish_packed. */
int64_t lrx_packed_bytes(int32_t B, int32_t k);
int lrx_search_local_packed(lrx_handle* h, const void* dev_q_fp16, const int32_t* dev_q_terms,
                            const int32_t* dev_q_ptr, int32_t B, int32_t k, int32_t mode,
                            int32_t width, void* dev_packed);
int lrx_search_finish_packed(lrx_handle* h, const void* dev_packed_all, int32_t world, int32_t B,
                             int32_t k, int32_t mode, const double* dev_weights, int64_t* dev_ids,
                             double* dev_score, double* dev_sem, double* dev_kw,
                             int32_t* dev_status);

/* ---- sharded search with the exchange fused into the kernels (world > 1, one box).
 * Replaces "lrx_search_local_packed -> NCCL all-gather -> lrx_search_finish_packed": every rank
 * owns an exchange region (two parities x world slots of one packed block, sequence flags and the
 * call counter); the packing kernel stores the rank's records straight into the slot `rank` of EVERY
 * peer's region over NVLink (peer memory opened through CUDA IPC), its last CTA publishes the
 * call's sequence number with a release store, and the fusion kernel acquires the world flags
 * before it merges.  The sequence number lives on the device, so the whole chain is one captured
 * CUDA graph replayed per batch: no collective library call, no host synchronisation, one launch.
 * Calls must be made in the same order by all ranks (SPMD), each on its own stream.  dev_status[b]:
 * 0 ok, bit 0 = exactness guard (rerun with a larger width), bit 1 = token capacity exceeded,
 * -1 = a peer's block did not arrive within the exchange timeout.  world == 1 is accepted (no
 * exchange): the same call then serves every shard count.
 *   lrx_exchange_export  allocates this rank's region for batches up to (B_max, k_max) and
 *                        returns its IPC handle (LRX_IPC_HANDLE_BYTES bytes);
 *   lrx_exchange_import  takes all ranks' handles in rank order (exchanged by the host, e.g. one
 *                        torch.distributed all-gather at start-up) and opens the peers' regions;
 *   lrx_search_sharded   K2 + K3 on this shard, exchange, K4: the replicated fused result. */
int lrx_exchange_export(lrx_handle* h, int32_t B_max, int32_t k_max, void* host_handle_out);
int lrx_exchange_import(lrx_handle* h, const void* host_handles_all);
int lrx_search_sharded(lrx_handle* h, const void* dev_q_fp16, const int32_t* dev_q_terms,
                       const int32_t* dev_q_ptr, const double* dev_weights, int32_t B, int32_t k,
                       int32_t mode, int32_t width, int64_t* dev_ids, double* dev_score,
                       double* dev_sem, double* dev_kw, int32_t* dev_status);

/* Bounded spin of the fusion kernel on a peer's sequence flag (default 2000 ms); expiry is reported
 * as status -1 / LRX_E_PEER, never as a hung GPU. */
int lrx_set_exchange_timeout(lrx_handle* h, int32_t milliseconds);

/* Host-buffer form of the whole search: copies the inputs in, runs K2..K4 (exchange included when
 * the handle is a shard, world > 1: every rank makes the same call), copies the results out and
 * synchronises; widens the dense candidate lists and reruns when the exactness guard trips (the
 * status is replicated, so all ranks rerun together).  This is the call `RetrievalEngine.search` /
 * `search_batch` makes (retrieval_engine.py:59).  Any number of tokens per query. */
int lrx_search_batch_host(lrx_handle* h, const void* host_q_fp16, const int32_t* host_q_terms,
                          const int32_t* host_q_ptr, const double* host_weights, int32_t B,
                          int32_t k, int32_t mode, int64_t* host_ids, double* host_score,
                          double* host_sem, double* host_kw);

/* The same under the name the sharded callers use (world >= 1). */
int lrx_search_sharded_host(lrx_handle* h, const void* host_q_fp16, const int32_t* host_q_terms,
                            const int32_t* host_q_ptr, const double* host_weights, int32_t B,
                            int32_t k, int32_t mode, int64_t* host_ids, double* host_score,
                            double* host_sem, double* host_kw);

/* Split form for pipelining: _begin stages the inputs in pinned memory and enqueues H2D + the chain
 * + D2H on the handle's stream and returns at once; _end synchronises, checks the status words
 * (rerunning wider if needed) and copies the results out.  One search in flight per handle; several
 * handles over the same index (same corpus / postings pointers, each with its own stream and
 * exchange region) keep several batches in flight on one GPU. */
int lrx_search_host_begin(lrx_handle* h, const void* host_q_fp16, const int32_t* host_q_terms,
                          const int32_t* host_q_ptr, const double* host_weights, int32_t B, int32_t k,
                          int32_t mode);
int lrx_search_text_host_begin(lrx_handle* h, const int32_t* host_tok_ids, const int32_t* host_tok_lens,
                               int32_t S, const int32_t* host_q_terms, const int32_t* host_q_ptr,
                               const double* host_weights, int32_t B, int32_t k, int32_t mode);
int lrx_search_host_end(lrx_handle* h, int64_t* host_ids, double* host_score, double* host_sem,
                        double* host_kw);

/* The whole of RetrievalEngine.search for a batch of query STRINGS already tokenised on the host
 * (world >= 1; a shard encodes the strings redundantly): H2D of WordPiece ids [B,S] + lens [B] and of the BM25 term ids, K1 (encoder) ->
 * K2 -> K3 -> K4, D2H of the fused results, synchronised.  Replaces retrieval_engine.py:59-96
 * end to end; `search_batch` with B = 4 is the orchestrator fan-out (orchestrator.py:38-62). */
int lrx_search_text_host(lrx_handle* h, const int32_t* host_tok_ids, const int32_t* host_tok_lens,
                         int32_t S, const int32_t* host_q_terms, const int32_t* host_q_ptr,
                         const double* host_weights, int32_t B, int32_t k, int32_t mode,
                         int64_t* host_ids, double* host_score, double* host_sem, double* host_kw);

/* Kernel-tuning aid: when set (device int64[128], caller-owned; NULL to clear), CTA 0 of every
 * tensor-core GEMM launch writes clock64() stamps of its pipeline events there. */
int lrx_debug_set_trace(lrx_handle* h, void* dev_int64_128);

/* Test hook for K3: the scan divides with the branch-free fast path of the float64 division
 * (same FMA sequence as the compiler's, without the exponent-range test); this counts the
 * (tf, len) pairs in [0,n_tf) x [0,n_len) whose BM25 factor differs from the IEEE division. */
int lrx_debug_bm25_divcheck(lrx_handle* h, double avgdl, double k1, double b, int32_t n_tf,
                            int32_t n_len, uint64_t* host_mismatches);

/* Number of kernels launched by this handle since lrx_open (bench.py's gpu_launches). */
int64_t lrx_launch_count(const lrx_handle* h);


/* Per-kernel timing for bench.py's roofline line: when enabled, CUDA events are recorded
 * on the launch stream around every launch of the two streaming kernels
 * (which = 0: dense_scan_kernel, 1: bm25_scan_kernel).  lrx_profile_read synchronises,
 * returns the summed device time and launch count since the last read, and resets. */
int lrx_profile_enable(lrx_handle* h, int32_t on);
int lrx_profile_read(lrx_handle* h, int32_t which, double* total_ms, int64_t* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* LRX_H */
