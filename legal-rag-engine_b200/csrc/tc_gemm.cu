// K1 GEMMs: out[M,N] = epilogue(A[M,K] * W[N,K]^T) on the 5th-gen tensor cores.
//
// Replaces the torch-CPU sgemm calls inside SentenceTransformer.encode
// (reference: src/retrieval/retrieval_engine.py:61, create_vector_store.py:45): the six
// BertLayer projections of all-MiniLM-L6-v2.  fp16 operands, fp32 accumulation in TMEM.
//
// One CTA = one 128 x BN output tile, 192 threads, warp-specialised:
//   warp 4   TMA producer: 2-D tiled loads (SWIZZLE_128B) of the A tile (128 x 64 halves)
//            and the W tile (BN x 64 halves) into a 3-stage shared-memory ring, completion
//            counted on full[] mbarriers, slots recycled through empty[] mbarriers.
//   warp 5   allocates TMEM and issues tcgen05.mma (UMMA 128 x N x 16, cta_group::1) from ONE
//            lane: 4 k-steps per stage; tcgen05.commit releases the stage / publishes the
//            accumulator.
//   warps 0-3 epilogue: thread r owns output row r (= TMEM lane r): tcgen05.ld 32 columns at
//            a time, then
//              EPI_BIAS        + bias                              -> fp16   (QKV projection)
//              EPI_BIAS_GELU   + bias, exact-erf GELU              -> fp16   (FFN up)
//              EPI_BIAS_RES_LN + bias + residual, LayerNorm(384)   -> fp16   (attention output /
//                              FFN down; BN = 384 = the whole row, so mean/variance are
//                              per-thread sums; the pre-norm row is parked in TMEM between
//                              the three passes with tcgen05.st)
//              EPI_F32         raw fp32 accumulators                         (tests / K2b)
//
// Roofline: tensor pipe.  flops per launch = 2*M*N*K.
#include <cstdio>

#include "handle.h"
#include "tc.cuh"

namespace lrx {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kABytes = kBM * kBK * 2;   // 16 KB

struct GemmEpi {
    const float* bias;        // [N]
    const __half* residual;   // [M, ld_res]      (LN epilogue)
    const float* gamma;       // [N]
    const float* beta;        // [N]
    void* out;                // fp16 [M, ld_out] (fp32 for EPI_F32)
    int ld_out, ld_res;
    int M;                    // valid rows
    float eps;
};

template <int BN>
struct GemmCfg {
    static constexpr int kStages = (BN >= 256) ? 3 : 3;
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = (BN <= 128) ? 128 : (BN <= 256 ? 256 : 512);
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void store_row32_f16(__half* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h0 = __floats2half2_rn(v[8 * i + 0], v[8 * i + 1]);
        __half2 h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
        __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]);
        __half2 h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        d4[i] = u;
    }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               int num_k_blocks, GemmEpi ep) {
    using Cfg = GemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char gemm_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;
    unsigned char* sB = base + kStages * kABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + kStages * Cfg::kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* tmem_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x;
    const int n_blk = blockIdx.y;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], (uint32_t)Cfg::kStageBytes);
                tma_load_2d(sA + s * kABytes, &tma_a, kb * kBK, m_blk * kBM, &full[s]);
#pragma unroll
                for (int nb = 0; nb < BN / 128; ++nb)
                    tma_load_2d(sB + s * Cfg::kBBytes + nb * (128 * kBK * 2), &tma_b, kb * kBK,
                                n_blk * BN + nb * 128, &full[s]);
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one lane)
        if (lane == 0) {
            constexpr int N0 = (BN > 256) ? 256 : BN;
            constexpr int N1 = BN - N0;
            constexpr uint32_t idesc0 = umma_idesc_f16(kBM, N0);
            constexpr uint32_t idesc1 = umma_idesc_f16(kBM, N1 > 0 ? N1 : 16);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * kABytes);
                const uint32_t b_addr = smem_u32(sB + s * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    const uint64_t ad = umma_desc_sw128(a_addr + k * 32);
                    const uint64_t bd = umma_desc_sw128(b_addr + k * 32);
                    const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                    umma_f16(tmem_base, ad, bd, idesc0, acc);
                    if (N1 > 0) {
                        const uint64_t bd1 = umma_desc_sw128(b_addr + N0 * 128 + k * 32);
                        umma_f16(tmem_base + N0, ad, bd1, idesc1, acc);
                    }
                }
                umma_commit(&empty[s]);                       // smem slot free when the MMAs retire
                if (kb == num_k_blocks - 1) umma_commit(tmem_full);   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: thread = output row = TMEM lane
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int64_t grow = (int64_t)m_blk * kBM + row;
        const bool ok = grow < ep.M;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int n0 = n_blk * BN;
        uint32_t r[32];
        float v[32];
        if (EPI == 3) {
            float* out = reinterpret_cast<float*>(ep.out) + grow * ep.ld_out + n0;
            for (int c = 0; c < BN; c += 32) {
                tmem_ld32(trow + c, r);
                tmem_wait_ld();
                if (ok) {
                    float4* o4 = reinterpret_cast<float4*>(out + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o4[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                }
            }
        } else if (EPI == 0 || EPI == 1) {
            __half* out = reinterpret_cast<__half*>(ep.out) + grow * ep.ld_out + n0;
            for (int c = 0; c < BN; c += 32) {
                tmem_ld32(trow + c, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(r[j]) + __ldg(ep.bias + n0 + c + j);
                    v[j] = (EPI == 1) ? gelu_erf(x) : x;
                }
                if (ok) store_row32_f16(out + c, v);
            }
        } else {
            // bias + residual, LayerNorm over the whole row (BN == N)
            const __half* res = ep.residual + grow * ep.ld_res;
            float sum = 0.f;
            for (int c = 0; c < BN; c += 32) {
                tmem_ld32(trow + c, r);
                uint4 rr[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    rr[i] = ok ? __ldg(reinterpret_cast<const uint4*>(res + c) + i) : make_uint4(0, 0, 0, 0);
                tmem_wait_ld();
                const __half2* rh = reinterpret_cast<const __half2*>(rr);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 f = __half22float2(rh[j]);
                    const float x0 = __uint_as_float(r[2 * j]) + __ldg(ep.bias + c + 2 * j) + f.x;
                    const float x1 = __uint_as_float(r[2 * j + 1]) + __ldg(ep.bias + c + 2 * j + 1) + f.y;
                    sum += x0 + x1;
                    r[2 * j] = __float_as_uint(x0);
                    r[2 * j + 1] = __float_as_uint(x1);
                }
                tmem_st32(trow + c, r);
            }
            tmem_wait_st();
            const float mean = sum * (1.0f / BN);
            float var = 0.f;
            for (int c = 0; c < BN; c += 32) {
                tmem_ld32(trow + c, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float d = __uint_as_float(r[j]) - mean;
                    var = fmaf(d, d, var);
                }
            }
            const float rstd = 1.0f / sqrtf(var * (1.0f / BN) + ep.eps);
            __half* out = reinterpret_cast<__half*>(ep.out) + grow * ep.ld_out;
            for (int c = 0; c < BN; c += 32) {
                tmem_ld32(trow + c, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = (__uint_as_float(r[j]) - mean) * rstd * __ldg(ep.gamma + c + j) +
                           __ldg(ep.beta + c + j);
                if (ok) store_row32_f16(out + c, v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

// fp16 row-major [rows, cols] with leading dimension ld (elements); box = 64 cols x box_rows,
// SWIZZLE_128B; out-of-range rows/cols read as zero.
cudaError_t make_tmap_f16(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int64_t ld,
                          int box_rows) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (fn == nullptr) return cudaErrorNotSupported;
    if (rows <= 0) rows = 1;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BN, int EPI>
static cudaError_t launch_one(cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M,
                              int N, int K, const GemmEpi& ep) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, EPI>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)GemmCfg<BN>::kSmem);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    dim3 grid((M + kBM - 1) / kBM, N / BN);
    tc_gemm_kernel<BN, EPI><<<grid, kGemmThreads, GemmCfg<BN>::kSmem, st>>>(ta, tb, K / kBK, ep);
    return cudaGetLastError();
}

// epi: 0 bias, 1 bias+GELU, 2 bias+residual+LayerNorm (N must be 384), 3 raw fp32.
cudaError_t launch_tc_gemm(lrx_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                           int K, int epi, const float* bias, const __half* residual, int ld_res,
                           const float* gamma, const float* beta, float eps, void* out, int ld_out) {
    if (M <= 0) return cudaSuccess;
    if (K % kBK != 0 || N % 128 != 0) return cudaErrorInvalidValue;
    GemmEpi ep;
    ep.bias = bias; ep.residual = residual; ep.gamma = gamma; ep.beta = beta;
    ep.out = out; ep.ld_out = ld_out; ep.ld_res = ld_res; ep.M = M; ep.eps = eps;
    cudaError_t e;
    switch (epi) {
        case 0: e = launch_one<128, 0>(h->stream, ta, tb, M, N, K, ep); break;
        case 1: e = launch_one<128, 1>(h->stream, ta, tb, M, N, K, ep); break;
        case 2:
            if (N != 384) return cudaErrorInvalidValue;
            e = launch_one<384, 2>(h->stream, ta, tb, M, N, K, ep);
            break;
        case 3: e = launch_one<128, 3>(h->stream, ta, tb, M, N, K, ep); break;
        default: return cudaErrorInvalidValue;
    }
    h->launches++;
    return e;
}

// One-off form (tests, stage benchmarks): builds the two tensor maps per call.
cudaError_t gemm_f16_adhoc(lrx_handle* h, const void* a, const void* w, int M, int N, int K, int epi,
                           const float* bias, const void* residual, const float* gamma,
                           const float* beta, float eps, void* out) {
    CUtensorMap ta, tb;
    cudaError_t e = make_tmap_f16(&ta, a, M, K, K, 128);
    if (e != cudaSuccess) return e;
    e = make_tmap_f16(&tb, w, N, K, K, 128);
    if (e != cudaSuccess) return e;
    return launch_tc_gemm(h, ta, tb, M, N, K, epi, bias, (const __half*)residual, N, gamma, beta, eps,
                          out, N);
}

}  // namespace lrx
