/*
 * Oracle / CPU baseline (TEST INFRASTRUCTURE, see oracle/__init__.py; parity unpinned).
 *
 * C restatement of what faiss-cpu's IndexFlatIP.search does for a small query batch
 * (reference call site: src/retrieval/retrieval_engine.py:64; faiss-cpu>=1.7.4 is a
 * PyPI dependency, source not under /root/reference).  Published algorithm for
 * nq < 20 ("exhaustive_inner_product_seq"): every query scans all rows sequentially
 * with a SIMD fp32 dot product and keeps the best k in a binary min-heap; OpenMP
 * parallelises over QUERIES only, so one query uses one core.
 *
 * Used by bench.py as the timed CPU baseline for the dense stage, and by the tests
 * as a second opinion on ranking (fp32 accumulation: scores agree with the exact
 * oracle to ~1e-6, ids agree wherever scores are separated by more than that).
 */
#include <stdint.h>
#include <stdlib.h>
#include <float.h>
#include <immintrin.h>
#include <pthread.h>

static inline float dot_f32(const float* a, const float* b, int d) {
#if defined(__AVX2__) && defined(__FMA__)
    __m256 acc0 = _mm256_setzero_ps(), acc1 = _mm256_setzero_ps();
    int i = 0;
    for (; i + 16 <= d; i += 16) {
        acc0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), acc0);
        acc1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), acc1);
    }
    acc0 = _mm256_add_ps(acc0, acc1);
    __m128 s = _mm_add_ps(_mm256_castps256_ps128(acc0), _mm256_extractf128_ps(acc0, 1));
    s = _mm_hadd_ps(s, s);
    s = _mm_hadd_ps(s, s);
    float r = _mm_cvtss_f32(s);
    for (; i < d; ++i) r += a[i] * b[i];
    return r;
#else
    float r = 0.f;
    for (int i = 0; i < d; ++i) r += a[i] * b[i];
    return r;
#endif
}

/* min-heap on (score, then LARGER id is "smaller"): root = current worst kept */
static inline int worse(float s1, int64_t i1, float s2, int64_t i2) {
    return (s1 < s2) || (s1 == s2 && i1 > i2);
}
static void heap_replace_root(float* hs, int64_t* hi, int k, float s, int64_t id) {
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        float ms = s; int64_t mi = id;
        if (l < k && worse(hs[l], hi[l], ms, mi)) { m = l; ms = hs[l]; mi = hi[l]; }
        if (r < k && worse(hs[r], hi[r], ms, mi)) { m = r; }
        if (m == i) break;
        hs[i] = hs[m]; hi[i] = hi[m];
        i = m;
    }
    hs[i] = s; hi[i] = id;
}
static int cmp_desc(const void* pa, const void* pb, void* ctx) {
    (void)ctx;
    const float* a = (const float*)pa; const float* b = (const float*)pb;
    return (a[0] < b[0]) - (a[0] > b[0]);
}

typedef struct {
    const float* x; int64_t n; int d; const float* q; int k; float* D; int64_t* I;
    int b0, b1;
} scan_job;

static void scan_one(const scan_job* J, int b) {
    const int k = J->k, d = J->d;
    float* hs = J->D + (int64_t)b * k;
    int64_t* hi = J->I + (int64_t)b * k;
    for (int j = 0; j < k; ++j) { hs[j] = -FLT_MAX; hi[j] = INT64_MAX; }
    const float* qb = J->q + (int64_t)b * d;
    for (int64_t r = 0; r < J->n; ++r) {
        const float s = dot_f32(J->x + r * d, qb, d);
        if (worse(hs[0], hi[0], s, r)) heap_replace_root(hs, hi, k, s, r);
    }
    /* sort best first (selection sort on k <= a few hundred) */
    for (int a = 0; a < k; ++a) {
        int best = a;
        for (int c = a + 1; c < k; ++c)
            if (worse(hs[best], hi[best], hs[c], hi[c])) best = c;
        float ts = hs[a]; hs[a] = hs[best]; hs[best] = ts;
        int64_t ti = hi[a]; hi[a] = hi[best]; hi[best] = ti;
    }
    for (int a = 0; a < k; ++a) if (hi[a] == INT64_MAX) hi[a] = -1;
}
static void* scan_thread(void* arg) {
    const scan_job* J = (const scan_job*)arg;
    for (int b = J->b0; b < J->b1; ++b) scan_one(J, b);
    return 0;
}

/* x: fp32 [n,d] row-major; q: fp32 [nq,d]; outputs D [nq,k], I [nq,k] best first,
 * ties by ascending id; pads (-FLT_MAX, -1).  Parallel over QUERIES only (as FAISS):
 * min(nq, nthreads) threads, one query per core at a time. */
void oracle_flat_ip_search(const float* x, int64_t n, int d, const float* q, int nq, int k,
                           float* D, int64_t* I, int nthreads) {
    if (nthreads <= 0 || nthreads > nq) nthreads = nq;
    if (nthreads > 64) nthreads = 64;
    pthread_t th[64];
    scan_job jobs[64];
    for (int t = 0; t < nthreads; ++t) {
        scan_job J = {x, n, d, q, k, D, I, (int)((int64_t)nq * t / nthreads),
                      (int)((int64_t)nq * (t + 1) / nthreads)};
        jobs[t] = J;
        if (t > 0) pthread_create(&th[t], 0, scan_thread, &jobs[t]);
    }
    scan_thread(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
    (void)cmp_desc;
}

/* CSR BM25 ("a fair CPU"): same float64 operations as oracle/bm25.py, term by term. */
void oracle_bm25_scores(const int64_t* term_ptr, const int64_t* post_doc, const int64_t* post_tf,
                        const double* idf, const double* doc_norm, double k1,
                        const int32_t* q_terms, int n_q_terms, double* score /* [n_docs], zeroed */) {
    const double k1p1 = k1 + 1.0;
    for (int s = 0; s < n_q_terms; ++s) {
        const int t = q_terms[s];
        if (t < 0) continue;
        const double w = idf[t];
        if (w == 0.0) continue;
        for (int64_t p = term_ptr[t]; p < term_ptr[t + 1]; ++p) {
            const double tf = (double)post_tf[p];
            const int64_t dd = post_doc[p];
            const double num = tf * k1p1;
            const double den = tf + doc_norm[dd];
            score[dd] = score[dd] + w * (num / den);
        }
    }
}
