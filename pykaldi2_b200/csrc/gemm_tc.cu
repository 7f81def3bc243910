// bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = A[M,K] * B[N,K]^T (+ bias[N])     A, B bf16 row-major (K contiguous), fp32 accumulate
//
// This is the dense contraction of the BLSTM acoustic model (reference models/lstm.py:46-61,
// which reaches cuDNN/cuBLAS through nn.LSTM / nn.Linear): input projections X*W_ih^T,
// the output layer, and the weight / input gradients of the backward pass.
//
// Persistent, warp-specialised kernel (one CTA per SM, 256 threads):
//   warp 0   TMA producer  : cp.async.bulk.tensor (SWIZZLE_128B) A/B tiles -> 4-stage smem ring
//   warp 1   MMA issuer    : one elected thread issues tcgen05.mma (M=128, N=256, K=16) into TMEM
//   warp 2   TMEM allocator (2 accumulator stages x 256 fp32 columns = all 512 columns)
//   warps 4-7 epilogue     : tcgen05.ld TMEM -> registers -> (+bias) -> global (fp32 or bf16)
// smem full/empty mbarriers couple TMA and MMA; tmem full/empty mbarriers couple MMA and the
// epilogue so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <mutex>

namespace {

constexpr int BM = 128, BN = 256, BK = 64;      // BK * 2 B = 128 B = one swizzle row
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kThreads = 256;
constexpr uint32_t kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
constexpr uint32_t kTmemCols = kAccStages * BN;  // 512: the whole TMEM

using namespace tc;

struct __align__(8) PipeBars {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[kAccStages];
    uint64_t tmem_empty[kAccStages];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    void* __restrict__ Cout, const float* __restrict__ bias, int M, int N, int K, int ldc,
                    int c_bf16, int lstm_T, int lstm_B, int lstm_H) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    PipeBars* bars = reinterpret_cast<PipeBars*>(smem + kStages * kStageBytes);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);   // uniform register for the tcgen05 ops

    if (warp == 0) {
        // ===== TMA producer =====  (whole warp runs the loop, one elected lane issues: in a divergent
        // `if (lane == 0)` region ptxas wraps every UTMALDG / UTCHMMA in an ELECT/R2UR waterfall loop)
        uint32_t stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile % num_m) * BM, n0 = (tile / num_m) * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&bars->empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&bars->full[stage], kStageBytes);
                    uint8_t* sa = smem + stage * kStageBytes;
                    tma_load_2d(&map_a, &bars->full[stage], sa, kb * BK, m0);
                    tma_load_2d(&map_b, &bars->full[stage], sa + kStageBytesA, kb * BK, n0);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(BM, BN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&bars->full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                const uint64_t adesc = make_sw128_desc(sa);
                const uint64_t bdesc = make_sw128_desc(sa + kStageBytesA);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in 16 B units
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                    (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(&bars->empty[stage]);           // frees the smem slot when the MMAs retire
                    if (kb == num_kb - 1) tc_commit(&bars->tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int ew = warp - 4;                    // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m0 = (tile % num_m) * BM, n0 = (tile / num_m) * BN;
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m0 + ew * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc * BN + c * 32 + ((uint32_t)(ew * 32) << 16);
                tc_ld_32x32b_x32(taddr, v);
                const int col0 = n0 + c * 32;
                if (row < M && col0 < N) {
                    if (col0 + 32 <= N) {
                        if (c_bf16) {
                            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cout) + (int64_t)row * ldc + col0;
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                float f[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[j + q]) + (bias ? __ldg(&bias[col0 + j + q]) : 0.f);
                                uint4 o;
                                __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]);
                                __nv_bfloat162 p1 = __floats2bfloat162_rn(f[2], f[3]);
                                __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]);
                                __nv_bfloat162 p3 = __floats2bfloat162_rn(f[6], f[7]);
                                o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                                o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                                if ((reinterpret_cast<uintptr_t>(dst + j) & 15) == 0) {
                                    *reinterpret_cast<uint4*>(dst + j) = o;
                                } else {
#pragma unroll
                                    for (int q = 0; q < 8; ++q) dst[j + q] = __float2bfloat16(f[q]);
                                }
                            }
                        } else {
                            float* dst = reinterpret_cast<float*>(Cout) + (int64_t)row * ldc + col0;
                            if (lstm_T > 0) {
                                // LSTM gate layout [T][2][H/32][B][128]: row = b*T + t, col = dir*4H + gate*H + unit
                                const int b = row / lstm_T, t = row - b * lstm_T;
                                const int dir = col0 / (4 * lstm_H), r = col0 - dir * 4 * lstm_H;
                                const int gate = r / lstm_H, unit = r - gate * lstm_H;
                                dst = reinterpret_cast<float*>(Cout) +
                                      ((((int64_t)t * 2 + dir) * (lstm_H >> 5) + (unit >> 5)) * lstm_B + b) * 128 + gate * 32;
                            }
                            const bool al = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 o;
                                o.x = __uint_as_float(v[j + 0]) + (bias ? __ldg(&bias[col0 + j + 0]) : 0.f);
                                o.y = __uint_as_float(v[j + 1]) + (bias ? __ldg(&bias[col0 + j + 1]) : 0.f);
                                o.z = __uint_as_float(v[j + 2]) + (bias ? __ldg(&bias[col0 + j + 2]) : 0.f);
                                o.w = __uint_as_float(v[j + 3]) + (bias ? __ldg(&bias[col0 + j + 3]) : 0.f);
                                if (al) *reinterpret_cast<float4*>(dst + j) = o;
                                else { dst[j] = o.x; dst[j + 1] = o.y; dst[j + 2] = o.z; dst[j + 3] = o.w; }
                            }
                        }
                    } else {
                        for (int j = 0; j < 32 && col0 + j < N; ++j) {
                            const float f = __uint_as_float(v[j]) + (bias ? __ldg(&bias[col0 + j]) : 0.f);
                            if (c_bf16) reinterpret_cast<__nv_bfloat16*>(Cout)[(int64_t)row * ldc + col0 + j] = __float2bfloat16(f);
                            else reinterpret_cast<float*>(Cout)[(int64_t)row * ldc + col0 + j] = f;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    PK2_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PK2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
    return 0;
}

int g_num_sms = 0;
thread_local int g_lstm_T = 0, g_lstm_B = 0, g_lstm_H = 0;
thread_local int g_max_ctas = 0;      // 0 = all SMs; set to leave SMs to a concurrently running persistent kernel

}  // namespace

extern "C" int pk2_gemm_bf16_nt(const void* A, const void* B, void* C, const float* bias, int M, int N, int K,
                                int lda, int ldb, int ldc, int flags, void* stream) {
    PK2_REQUIRE(A && B && C, "pk2_gemm_bf16_nt: null argument");
    PK2_REQUIRE(M > 0 && N > 0 && K > 0, "pk2_gemm_bf16_nt: empty problem %dx%dx%d", M, N, K);
    PK2_REQUIRE((flags & 1) == 0, "pk2_gemm_bf16_nt: accumulate flag not supported");
    PK2_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "pk2_gemm_bf16_nt: lda/ldb must be multiples of 8 elements (16 B)");
    PK2_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "pk2_gemm_bf16_nt: A/B must be 16-byte aligned");
    CUtensorMap ma, mb;
    if (make_map(&ma, A, M, K, lda, BM)) return 2;
    if (make_map(&mb, B, N, K, ldb, BN)) return 2;
    if (g_num_sms == 0) {
        int dev = 0;
        PK2_CHECK(cudaGetDevice(&dev));
        PK2_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = kStages * kStageBytes + sizeof(PipeBars) + 1024;
    static bool attr = false;
    if (!attr) {
        PK2_CHECK(cudaFuncSetAttribute(gemm_bf16_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int grid = tiles < g_num_sms ? tiles : g_num_sms;
    if (g_max_ctas > 0 && grid > g_max_ctas) grid = g_max_ctas;
    gemm_bf16_nt_kernel<<<grid, kThreads, smem, pk2::as_stream(stream)>>>(ma, mb, C, bias, M, N, K, ldc, (flags >> 1) & 1,
                                                                          g_lstm_T, g_lstm_B, g_lstm_H);
    PK2_POST_LAUNCH();
    return 0;
}

// Cap the number of CTAs of subsequent GEMM launches from this thread (0 = no cap): used when weight-gradient
// GEMMs run on a side stream next to the persistent recurrence kernels, which own part of the SMs.
extern "C" int pk2_gemm_set_max_ctas(int n) { g_max_ctas = n > 0 ? n : 0; return 0; }

extern "C" int pk2_lstm_input_proj(const void* x, const void* wih, const float* bias, float* gx, int B, int T,
                                   int I, int H, int ldx, void* stream) {
    PK2_REQUIRE(H % 32 == 0 && H > 0, "pk2_lstm_input_proj: hidden size must be a multiple of 32");
    PK2_REQUIRE((8 * H) % 128 == 0, "pk2_lstm_input_proj: 8*H must be a multiple of 128");
    g_lstm_T = T; g_lstm_B = B; g_lstm_H = H;
    const int rc = pk2_gemm_bf16_nt(x, wih, gx, bias, B * T, 8 * H, I, ldx, I, 8 * H, 0, stream);
    g_lstm_T = g_lstm_B = g_lstm_H = 0;
    return rc;
}
