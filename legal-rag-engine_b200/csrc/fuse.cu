// K4: cross-shard merge of candidate records + score fusion.
//
// linear : reference src/retrieval/retrieval_engine.py:71-96, operation for
//          operation in float64 (no FMA contraction) -- candidates are the dense
//          top-2k only, semantic = float32-rounded inner product, keyword =
//          bm25/max_bm25, stable descending sort (ties keep flat-IP order).
// rrf    : README.md:39,82-83 (no upstream code): sum of 1/(60+rank) over the dense
//          and BM25 lists, order (rrf desc, id asc).  See oracle/fusion.py.
//
// One CTA per sub-query; every shard runs the same deterministic merge on the
// all-gathered records, so the result is replicated bit for bit.  The shards' lists arrive
// sorted, so merging is rank arithmetic (binary searches), not sorting.
#include "common.cuh"
#include "handle.h"

namespace lrx {

constexpr int kFuseThreads = 128;
constexpr int kFuseMaxIn = 2048;   // world * K records per list
constexpr int kFuseMaxK = LRX_MAX_DEPTH;

// One WARP per (query, list position): packs the dense record and the BM25 record of that
// position; for rrf the BM25 hit's exact dense score (shown as "semantic") is computed here, one
// float64 dot product by the warp (x == NULL: taken from bm_dense instead).
__global__ void pack_records_kernel(int B, int K, int mode, const double* __restrict__ dense_exact,
                                    const int64_t* __restrict__ dense_ids,
                                    const double* __restrict__ dense_bm25,
                                    const double* __restrict__ bm_scores,
                                    const int64_t* __restrict__ bm_ids,
                                    const double* __restrict__ bm_dense,
                                    const unsigned char* __restrict__ x, int64_t n_rows,
                                    int64_t id_base, const __half* __restrict__ q,
                                    lrx_record* __restrict__ records) {
    pdl_trigger();                       // the exchange / fusion kernel may be scheduled
    pdl_wait();                          // bm25_at_kernel's scores are visible
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= B * K) return;
    const int b = i / K, j = i - b * K;
    lrx_record r1;
    r1.id = -1;
    r1.dense = -INFINITY;
    r1.bm25 = 0.0;
    if (mode == LRX_FUSE_RRF && bm_ids[i] >= 0) {
        r1.id = bm_ids[i];
        r1.bm25 = bm_scores[i];
        if (x != nullptr) {
            const int64_t row = r1.id - id_base;
            if (row >= 0 && row < n_rows) {
                const uint2* rowp = reinterpret_cast<const uint2*>(x + row * (int64_t)(LRX_DIM * 2));
                const uint2* qp = reinterpret_cast<const uint2*>(q + (size_t)b * LRX_DIM);
                double acc = 0.0;
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const uint2 v = rowp[s * 32 + lane], w = qp[s * 32 + lane];
                    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
                    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
                    const float2 e = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
                    acc = fma((double)a.x, (double)e.x, acc);
                    acc = fma((double)a.y, (double)e.y, acc);
                    acc = fma((double)c.x, (double)f.x, acc);
                    acc = fma((double)c.y, (double)f.y, acc);
                }
#pragma unroll
                for (int lb = 16; lb > 0; lb >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, lb);
                r1.dense = acc;
            }
        } else {
            r1.dense = bm_dense[i];
        }
    }
    if (lane == 0) {
        lrx_record r0;
        r0.id = dense_ids[i];
        r0.dense = dense_exact[i];
        r0.bm25 = (r0.id >= 0) ? dense_bm25[i] : 0.0;
        records[((size_t)b * 2 + 0) * K + j] = r0;
        records[((size_t)b * 2 + 1) * K + j] = r1;
    }
}

// key = order image (64) | ~id (32) | source slot (32): sorts by (value desc, id asc)
__device__ __forceinline__ u128 rec_key(double v, int64_t id, uint32_t src) {
    return ((u128)f64_ord(v) << 64) | ((u128)(uint32_t)(~(uint32_t)id) << 32) | (u128)src;
}

// Number of keys greater than `key` in a DESCENDING run of n keys (empty keys = 0 at its end).
__device__ __forceinline__ int count_greater_sorted(const u128* run, int n, u128 key) {
    int lo = 0, hi = n;                  // first position whose key is <= `key`
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (run[mid] > key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// No sorting networks here: every list that arrives is already sorted per shard (K2's re-score and
// K3's merge emit (score desc, id asc)), so the rank of a record in the merged order is a sum of
// binary searches, and the final order of the <= 2K fused scores is a rank by counting.  Keys are
// unique (they end in the source slot), so ranks are a permutation.
__global__ void __launch_bounds__(kFuseThreads)
fuse_kernel(const lrx_record* __restrict__ rec_all, const double* __restrict__ max_all,
            const int32_t* __restrict__ flags_all, int64_t shard_stride /* bytes; 0 = dense */,
            int world, int B, int K, int k, int mode,
            const double* __restrict__ weights, int64_t* __restrict__ out_ids,
            double* __restrict__ out_score, double* __restrict__ out_sem,
            double* __restrict__ out_kw, int32_t* __restrict__ out_status,
            const unsigned long long* wait_flags /* [world] or NULL */, unsigned long long wait_seq,
            int self) {
    pdl_wait();                                  // the packing / exchange kernel has finished
    extern __shared__ __align__(16) unsigned char fuse_dyn[];
    u128* keys = reinterpret_cast<u128*>(fuse_dyn);   // [kFuseMaxIn]
    __shared__ lrx_record dsel[kFuseMaxK];      // global dense top-K
    __shared__ lrx_record ssel[kFuseMaxK];      // global BM25 top-K (rrf)
    __shared__ double fscore[2 * kFuseMaxK];
    __shared__ int out_slot[2 * kFuseMaxK];     // fused rank -> union slot
    __shared__ int n_dense, n_sparse, n_fused;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int n_in = world * K;
    if (tid == 0) { n_dense = 0; n_sparse = 0; n_fused = 0; }
    // peer exchange: the blocks of the other shards were stored into this GPU's memory by THEIR
    // kernels; acquire every shard's sequence flag before touching its block (bounded spin: a
    // lost peer surfaces as status -1, not as a hung GPU)
    __shared__ int wait_failed;
    if (tid == 0) wait_failed = 0;
    __syncthreads();
    if (wait_flags != nullptr && tid < world && tid != self) {
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_flags + tid) : "memory");
            if (v >= wait_seq) break;
            if (clock64() - t0 > 4000000000ll) { wait_failed = 1; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();

    // max_bm25 = max(scores) if max(scores) > 0 else 1.0   (retrieval_engine.py:74)
    double maxbm = 0.0;
    int status = 0;
    for (int w = 0; w < world; ++w) {
        const double* mw = shard_stride
            ? reinterpret_cast<const double*>(reinterpret_cast<const char*>(max_all) + w * shard_stride)
            : max_all + (size_t)w * B;
        maxbm = fmax(maxbm, __ldcg(mw + b));
        if (flags_all != nullptr) {
            const int32_t* fw = shard_stride
                ? reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(flags_all) + w * shard_stride)
                : flags_all + (size_t)w * B;
            status |= __ldcg(fw + b);
        }
    }
    if (!(maxbm > 0.0)) maxbm = 1.0;

    // by value, through L2 (ld.cg): the records may have been written by another GPU
    auto rec_at = [&](int list, int src) -> lrx_record {
        const int w = src / K, j = src - w * K;
        const lrx_record* rw = shard_stride
            ? reinterpret_cast<const lrx_record*>(reinterpret_cast<const char*>(rec_all) + w * shard_stride)
            : rec_all + (size_t)w * B * 2 * K;
        const lrx_record* r = rw + ((size_t)b * 2 + list) * K + j;
        lrx_record out;
        out.id = __ldcg(&r->id);
        out.dense = __ldcg(&r->dense);
        out.bm25 = __ldcg(&r->bm25);
        return out;
    };
    // merged top-K of the shards' sorted lists `list` (0 dense, 1 BM25) into sel[]; count in *n_sel
    auto merge_lists = [&](int list, lrx_record* sel, int* n_sel) {
        __syncthreads();                                      // keys[] free
        for (int i = tid; i < n_in; i += kFuseThreads) {
            const lrx_record r = rec_at(list, i);
            keys[i] = (r.id >= 0) ? rec_key(list ? r.bm25 : r.dense, r.id, (uint32_t)i) : (u128)0;
        }
        __syncthreads();
        for (int i = tid; i < n_in; i += kFuseThreads) {
            const u128 key = keys[i];
            if (key == 0) continue;
            const int w_own = i / K;
            int rank = i - w_own * K;                         // position in its own sorted run
            for (int w = 0; w < world && rank < K; ++w)
                if (w != w_own) rank += count_greater_sorted(keys + w * K, K, key);
            if (rank < K) {
                sel[rank] = rec_at(list, i);
                atomicAdd(n_sel, 1);
            }
        }
        __syncthreads();
    };
    // rank by counting of the n keys in keys[]: out_slot[rank] = low 32 bits of the key
    auto order_keys = [&](int n) {
        __syncthreads();
        for (int i = tid; i < n; i += kFuseThreads) {
            const u128 key = keys[i];
            if (key == 0) continue;
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += (keys[j] > key) ? 1 : 0;
            out_slot[rank] = (int)(uint32_t)key;
            atomicAdd(&n_fused, 1);
        }
        __syncthreads();
    };

    merge_lists(0, dsel, &n_dense);
    const int nd = n_dense;

    if (mode == LRX_FUSE_LINEAR) {
        const double w = weights[b];
        const double one_m_w = __dsub_rn(1.0, w);
        for (int j = tid; j < nd; j += kFuseThreads) {
            const double sem = (double)__double2float_rn(dsel[j].dense);   // float(dist)
            const double kw = __ddiv_rn(dsel[j].bm25, maxbm);
            const double s = __dadd_rn(__dmul_rn(sem, one_m_w), __dmul_rn(kw, w));
            fscore[j] = s;
            // stable descending sort: ties keep flat-IP order j
            keys[j] = ((u128)f64_ord(s) << 64) | ((u128)(uint32_t)(~(uint32_t)j) << 32) | (u128)(uint32_t)j;
        }
        order_keys(nd);
        const int n_out = min(k, n_fused);
        for (int i = tid; i < k; i += kFuseThreads) {
            const size_t o = (size_t)b * k + i;
            if (i < n_out) {
                const int j = out_slot[i];
                out_ids[o] = dsel[j].id;
                out_score[o] = fscore[j];
                out_sem[o] = (double)__double2float_rn(dsel[j].dense);
                out_kw[o] = __ddiv_rn(dsel[j].bm25, maxbm);
            } else {
                out_ids[o] = -1;
                out_score[o] = 0.0;
                out_sem[o] = 0.0;
                out_kw[o] = 0.0;
            }
        }
    } else {
        merge_lists(1, ssel, &n_sparse);
        const int ns = n_sparse;
        // union slots: [0, nd) dense entries, [nd, nd + ns) BM25 entries (unused when the
        // document is already in the dense list).  Dense term first: 0.0 + 1/(60 + rank).
        for (int j = tid; j < nd + ns; j += kFuseThreads)
            fscore[j] = (j < nd) ? __dadd_rn(0.0, __ddiv_rn(1.0, 60.0 + (double)(j + 1))) : -1.0;
        __syncthreads();
        for (int r = tid; r < ns; r += kFuseThreads) {
            const double term = __ddiv_rn(1.0, 60.0 + (double)(r + 1));
            int hit = -1;
            for (int j = 0; j < nd; ++j)
                if (dsel[j].id == ssel[r].id) hit = j;
            if (hit >= 0) fscore[hit] = __dadd_rn(fscore[hit], term);      // unique hit per r
            else fscore[nd + r] = __dadd_rn(0.0, term);
        }
        __syncthreads();
        for (int i = tid; i < nd + ns; i += kFuseThreads) {
            u128 key = 0;
            if (fscore[i] > 0.0) {
                const int64_t id = (i < nd) ? dsel[i].id : ssel[i - nd].id;
                key = rec_key(fscore[i], id, (uint32_t)i);
            }
            keys[i] = key;
        }
        order_keys(nd + ns);
        const int n_out = min(k, n_fused);
        for (int i = tid; i < k; i += kFuseThreads) {
            const size_t o = (size_t)b * k + i;
            if (i < n_out) {
                const int src = out_slot[i];
                const lrx_record& r = (src < nd) ? dsel[src] : ssel[src - nd];
                out_ids[o] = r.id;
                out_score[o] = fscore[src];
                out_sem[o] = (double)__double2float_rn(r.dense);
                out_kw[o] = __ddiv_rn(r.bm25, maxbm);
            } else {
                out_ids[o] = -1;
                out_score[o] = 0.0;
                out_sem[o] = 0.0;
                out_kw[o] = 0.0;
            }
        }
    }
    if (tid == 0) out_status[b] = wait_failed ? -1 : status;
}

// Peer exchange, sending side: CTA j stores this rank's packed block into slot `rank` of peer
// (rank + 1 + j) % world over NVLink, then publishes the call's sequence number there with a
// release store at system scope (fence cumulativity carries the whole CTA's stores).
__global__ void exchange_kernel(const uint4* __restrict__ mine, int n16, void* const* __restrict__ peers,
                                int rank, int world, size_t slot_off, size_t flag_off,
                                unsigned long long seq) {
    pdl_trigger();                       // the fusion kernel may be scheduled (it waits for us)
    pdl_wait();                          // pack_records_kernel's block is visible
    const int p = (rank + 1 + (int)blockIdx.x) % world;
    char* base = reinterpret_cast<char*>(peers[p]);
    uint4* dst = reinterpret_cast<uint4*>(base + slot_off);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = mine[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned long long* flag = reinterpret_cast<unsigned long long*>(base + flag_off);
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(seq) : "memory");
    }
}

cudaError_t launch_exchange(lrx_handle* h, const void* mine, size_t bytes, size_t slot_off,
                            size_t flag_off, unsigned long long seq) {
    if (h->world < 2) return cudaSuccess;
    cudaError_t e = launch_pdl(exchange_kernel, dim3(h->world - 1), dim3(256), 0, h->stream,
                               (const uint4*)mine, (int)(bytes / 16), (void* const*)h->xchg_peer_dev, h->rank,
                               h->world, slot_off, flag_off, seq);
    h->launches++;
    return e;
}

cudaError_t launch_pack_records(lrx_handle* h, int B, int K, int mode, const double* dense_exact,
                                const int64_t* dense_ids, const double* dense_bm25,
                                const double* bm_scores, const int64_t* bm_ids,
                                const double* bm_dense, const void* q, lrx_record* records) {
    // q != NULL: the exact dense scores of the BM25 hits are computed in the kernel (bm_dense unused)
    const int n = B * K;
    cudaError_t e = launch_pdl(pack_records_kernel, dim3((n + 7) / 8), dim3(256), 0, h->stream,
                               B, K, mode, dense_exact, dense_ids, dense_bm25, bm_scores, bm_ids, bm_dense,
                               q ? (const unsigned char*)h->x : (const unsigned char*)nullptr, h->n_local,
                               h->id_base, (const __half*)q, records);
    h->launches++;
    return e;
}

cudaError_t launch_fuse(lrx_handle* h, const lrx_record* records_all, const double* max_all,
                        const int32_t* flags_all, int64_t shard_stride, int world, int B, int K,
                        int k, int mode, const double* weights, int64_t* ids, double* score,
                        double* sem, double* kw, int32_t* status, const unsigned long long* wait_flags,
                        unsigned long long wait_seq, int self) {
    static bool attr_dev[64] = {false};   // function attributes are per device
    bool& attr = attr_dev[h->device & 63];
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(kFuseMaxIn * sizeof(u128)));
        if (e != cudaSuccess) return e;
        attr = true;
    }
    if (world * K > kFuseMaxIn || K > kFuseMaxK) return cudaErrorInvalidValue;
    cudaError_t e = launch_pdl(fuse_kernel, dim3(B), dim3(kFuseThreads), kFuseMaxIn * sizeof(u128), h->stream,
                               records_all, max_all, flags_all, shard_stride, world, B, K, k, mode, weights,
                               ids, score, sem, kw, status, wait_flags, wait_seq, self);
    h->launches++;
    return e;
}

}  // namespace lrx
