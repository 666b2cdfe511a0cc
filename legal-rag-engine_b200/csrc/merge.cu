// Merge of per-CTA / per-chunk sorted candidate lists into the best `width` keys of a
// query (second stage of the fused top-k of K2 and K3).  One CTA per query.
//
// The lists are sorted, so a valid lower bound L of the width-th best key is cheap: the
// width-th largest of the lists' first ceil(width/n_lists) entries.  Only the list
// prefixes >= L (a few hundred keys, not n_lists*width) are collected and sorted -- by
// one warp, without block barriers, when they fit 1024 entries.
#include "common.cuh"
#include "handle.h"

namespace lrx {

constexpr int kMergeThreads = 512;
constexpr int kMergeCap = 4096;

template <typename KeyT>
__device__ __forceinline__ void sort_desc_any(KeyT* buf, int n_valid, int tid, int warp, int lane) {
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

template <typename KeyT>
__global__ void __launch_bounds__(kMergeThreads, 1)
merge_keys_kernel(const KeyT* __restrict__ part, int n_lists, int list_stride /* in lists */,
                  const int* __restrict__ q_start /* or NULL */, int width,
                  KeyT* __restrict__ out /* [nq][width] */) {
    // two layouts: list l of query qi at part[(l * list_stride + qi) * width] (q_start == NULL:
    // the same number of lists for every query), or the lists q_start[qi] .. q_start[qi + 1] of a
    // flat array of lists (K3: warps are split between the queries by work)
    extern __shared__ __align__(128) unsigned char merge_raw[];
    KeyT* buf = reinterpret_cast<KeyT*>(merge_raw);
    __shared__ int count;
    __shared__ int overflow;
    __shared__ KeyT bound;
    const int qi = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int list0 = qi;
    if (q_start != nullptr) {
        list0 = q_start[qi];
        n_lists = q_start[qi + 1] - list0;
        list_stride = 1;
    }
    auto key_at = [&](int list, int pos) -> KeyT {
        return part[((size_t)list * list_stride + list0) * width + pos];
    };

    // ---- 1. lower bound from the heads of the lists
    int depth = (width + n_lists - 1) / n_lists;
    if (depth > width) depth = width;
    int ns = depth * n_lists;
    KeyT L = 0;
    if (ns <= kMergeCap) {
        for (int i = tid; i < ns; i += kMergeThreads) buf[i] = key_at(i / depth, i % depth);
        __syncthreads();
        sort_desc_any<KeyT>(buf, ns, tid, warp, lane);
        if (ns >= width) L = buf[width - 1];     // 0 when fewer than `width` keys exist
    }
    __syncthreads();

    // ---- 2. collect every key >= L (a prefix of each sorted list); tighten L if the
    //         buffer overflows (each retry drops >= kMergeCap - width keys)
    for (;;) {
        if (tid == 0) {
            count = 0;
            overflow = 0;
        }
        __syncthreads();
        for (int list = tid; list < n_lists; list += kMergeThreads) {
            for (int pos = 0; pos < width; ++pos) {
                const KeyT k = key_at(list, pos);
                if (k == 0 || k < L) break;
                const int p = atomicAdd(&count, 1);
                if (p < kMergeCap) buf[p] = k; else overflow = 1;
            }
        }
        __syncthreads();
        if (!overflow) break;
        sort_desc_any<KeyT>(buf, kMergeCap, tid, warp, lane);
        if (tid == 0) bound = buf[width - 1];
        __syncthreads();
        L = bound;
        __syncthreads();
    }
    // ---- 3. sort the survivors, emit the best `width`
    const int n = count;
    sort_desc_any<KeyT>(buf, n, tid, warp, lane);
    for (int i = tid; i < width; i += kMergeThreads)
        out[(size_t)qi * width + i] = (i < n) ? buf[i] : (KeyT)0;
}

template __global__ void merge_keys_kernel<uint64_t>(const uint64_t*, int, int, const int*, int, uint64_t*);
template __global__ void merge_keys_kernel<u128>(const u128*, int, int, const int*, int, u128*);

cudaError_t launch_merge_u64(cudaStream_t st, const uint64_t* part, int n_lists, int list_stride,
                             int width, int nq, uint64_t* out) {
    merge_keys_kernel<uint64_t><<<nq, kMergeThreads, kMergeCap * sizeof(uint64_t), st>>>(
        part, n_lists, list_stride, nullptr, width, out);
    return cudaGetLastError();
}
cudaError_t launch_merge_u128(cudaStream_t st, const void* part, const int* q_start, int width,
                              int nq, void* out) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(merge_keys_kernel<u128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kMergeCap * sizeof(u128)));
        attr = true;
    }
    merge_keys_kernel<u128><<<nq, kMergeThreads, kMergeCap * sizeof(u128), st>>>(
        (const u128*)part, 0, 1, q_start, width, (u128*)out);
    return cudaGetLastError();
}

}  // namespace lrx
