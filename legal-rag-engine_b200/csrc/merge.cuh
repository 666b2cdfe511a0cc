// Block-wide merge of sorted candidate lists into the best `width` keys of one query (second stage
// of the fused top-k of K2 and K3), as a device function so that the dense re-score (dense.cu) and
// the BM25 finalisation (bm25.cu) run in the same kernel as their merge.
//
// The lists are sorted descending, so no sorting network is needed anywhere:
//   1. a lower bound L of the width-th best key = the width-th largest of the heads of up to 256
//      lists (any `width` distinct keys bound it from below) -- found by rank counting;
//   2. the keys >= L are a prefix of every list: a few hundred keys, collected into shared memory;
//   3. their order is again a rank by counting (keys are unique); rank r < width lands in out[r].
// Rare overflows (more than kMergeCap survivors, e.g. thousands of exactly tied scores) fall back
// to a block-wide bitonic sort that tightens L.
#pragma once
#include "common.cuh"

namespace lrx {

constexpr int kMergeThreads = 512;
constexpr int kMergeCap = 4096;
constexpr int kMergeCountMax = 1024;   // rank counting up to this many keys, sorting above

template <typename KeyT>
__device__ __forceinline__ void merge_sort_desc_any(KeyT* buf, int n_valid, int tid, int warp, int lane) {
    // pads to a power of two; warp sort up to 128 entries, block sort above.  Block-uniform.
    const int p2 = max(32, next_pow2(n_valid));
    for (int i = n_valid + tid; i < p2; i += kMergeThreads) buf[i] = 0;
    __syncthreads();
    if (p2 <= 128) {                       // tiny: one warp, no block barriers
        if (warp == 0) warp_bitonic_sort_desc<KeyT>(buf, p2, lane);
        __syncthreads();
    } else {
        block_bitonic_sort_desc<KeyT>(buf, p2, 1, p2, tid, kMergeThreads);
    }
}

// Rank of buf[i] among the n unique keys of buf (0 = largest); keys equal to 0 are skipped by the
// callers.  Every thread reads the same buf[j] at the same time: a shared-memory broadcast.
template <typename KeyT>
__device__ __forceinline__ int merge_rank_of(const KeyT* buf, int n, KeyT key) {
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < n; ++j) rank += (buf[j] > key) ? 1 : 0;
    return rank;
}

// part: list l of this query is part[(list0 + l * list_stride) * list_width .. + list_width),
// sorted descending, empty keys (0) at the end.  buf: kMergeCap keys of shared memory.  out_s:
// `width` keys of shared memory (may not alias buf) -- the merged best keys, descending, 0-padded;
// `width` may exceed `list_width` (the int8 pre-filter of K2a keeps short per-CTA lists and a long
// merged one).  All kMergeThreads threads of the block must call; returns synced.
template <typename KeyT>
__device__ __forceinline__ void merge_lists_block(const KeyT* __restrict__ part, int n_lists,
                                                  int list_stride, int list0, int list_width, int width,
                                                  KeyT* buf, KeyT* out_s, int* s_count,
                                                  int* s_overflow, KeyT* s_bound) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto key_at = [&](int list, int pos) -> KeyT {
        return part[((size_t)list0 + (size_t)list * list_stride) * list_width + pos];
    };
    for (int i = tid; i < width; i += kMergeThreads) out_s[i] = 0;

    // ---- 1. lower bound from the heads of (up to 256 of) the lists
    const int nl = min(n_lists, 256);
    int depth = (nl > 0) ? (width + nl - 1) / nl : 0;
    if (depth > list_width) depth = list_width;
    const int ns = depth * nl;
    if (tid == 0) *s_bound = 0;
    for (int i = tid; i < ns; i += kMergeThreads) buf[i] = key_at(i / depth, i % depth);
    __syncthreads();
    if (ns >= width) {
        if (ns <= kMergeCountMax) {
            for (int i = tid; i < ns; i += kMergeThreads) {
                const KeyT key = buf[i];
                if (key != 0 && merge_rank_of<KeyT>(buf, ns, key) == width - 1) *s_bound = key;
            }
            __syncthreads();
        } else {
            merge_sort_desc_any<KeyT>(buf, ns, tid, warp, lane);
            if (tid == 0) *s_bound = buf[width - 1];
            __syncthreads();
        }
    }
    KeyT L = *s_bound;                       // 0 when fewer than `width` sampled keys exist
    __syncthreads();

    // ---- 2. collect every key >= L (a prefix of each sorted list); tighten L if the
    //         buffer overflows (each retry drops >= kMergeCap - width keys)
    for (;;) {
        if (tid == 0) {
            *s_count = 0;
            *s_overflow = 0;
        }
        __syncthreads();
        for (int list = tid; list < n_lists; list += kMergeThreads) {
            KeyT k = key_at(list, 0);
            KeyT k_next = (list_width > 1) ? key_at(list, 1) : (KeyT)0;      // second load in flight
            for (int pos = 0; pos < list_width; ++pos) {
                if (k == 0 || k < L) break;
                const int p = atomicAdd(s_count, 1);
                if (p < kMergeCap) buf[p] = k; else *s_overflow = 1;
                k = k_next;
                k_next = (pos + 2 < list_width) ? key_at(list, pos + 2) : (KeyT)0;
            }
        }
        __syncthreads();
        if (!*s_overflow) break;
        merge_sort_desc_any<KeyT>(buf, kMergeCap, tid, warp, lane);
        if (tid == 0) *s_bound = buf[width - 1];
        __syncthreads();
        L = *s_bound;
        __syncthreads();
    }
    // ---- 3. order the survivors, keep the best `width`
    const int n = *s_count;
    if (n <= kMergeCountMax) {
        for (int i = tid; i < n; i += kMergeThreads) {
            const KeyT key = buf[i];
            const int r = merge_rank_of<KeyT>(buf, n, key);
            if (r < width) out_s[r] = key;
        }
    } else {
        merge_sort_desc_any<KeyT>(buf, n, tid, warp, lane);
        for (int i = tid; i < width && i < n; i += kMergeThreads) out_s[i] = buf[i];
    }
    __syncthreads();
}

}  // namespace lrx
