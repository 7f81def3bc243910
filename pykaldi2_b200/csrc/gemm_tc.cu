// bf16 GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   C[M,N] = A[M,K] * B[N,K]^T (+ bias[N])     A, B bf16 row-major (K contiguous), fp32 accumulate
//   C[M,N] = At[K,M]^T * Bt[K,N]               "TN" form (PK2_GEMM_TN): both operands stored with the contraction
//                                              index as the ROW (M / N contiguous) and consumed in place through
//                                              MN-major UMMA descriptors -- the weight gradients dW = dY^T X, whose
//                                              operands are activations [frames, features]; no transposed copies
//   optional c_row_map: row r of the product is stored to row c_row_map[r] of C (scatter of compacted rows)
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
#include <map>
#include <mutex>

namespace {

constexpr int BM = 128, BN = 256, BK = 64;      // per-CTA tile rows, tile columns; BK * 2 B = 128 B = one swizzle row
constexpr int kAccStages = 2;
constexpr int kThreads1 = 256, kThreads2 = 384;   // 4 control warps + 4 (single CTA) / 8 (pair) epilogue warps
constexpr int kMaxStages = 6;
constexpr size_t kEpiTile = 32 * 32 * sizeof(float);   // per-warp epilogue transpose tile, 16-byte chunks XOR-swizzled by row
constexpr uint32_t kStageBytesA = BM * BK * 2;
constexpr uint32_t kTmemCols = kAccStages * BN;  // 512: the whole TMEM

using namespace tc;

struct __align__(8) PipeBars {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t tmem_full[kAccStages];
    uint64_t tmem_empty[kAccStages];
    uint32_t tmem_base;
};

// CTAS = 1: one CTA per 128 x 256 tile, 4-stage ring of (A 16 KB + B 32 KB).
// CTAS = 2: a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) per 256 x 256 tile: each CTA loads its own 128 rows
//           of A and HALF of the B tile (128 of the 256 columns' rows), the leader CTA issues M=256 MMAs that read both
//           CTAs' shared memory and write both CTAs' tensor memory.  Per CTA and k-block 32 KB arrive instead of
//           48 KB (the L2->SM operand traffic was the limiter of the 1-CTA kernel on the M-long shapes:
//           profiles/kernel_bench_r1_v2.jsonl) and the ring is 6 stages deep.
// Tile rasterisation: bands of kBandRows rows of A are swept over ALL column tiles before the next band starts
// (m fastest inside a band).  With the plain "all m for one n, then the next n" order every column pass streams the
// whole of A from HBM again (A of the model's shapes is about the size of L2): 16 x 118 MB for the input projections.
constexpr int kBandRows = 4096;
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int band, int* mt, int* nt) {
    const int per_band = band * num_n;
    const int b = tile / per_band, r = tile - b * per_band;
    const int bs = min(band, num_m - b * band);          // the last band may be short
    *nt = r / bs;
    *mt = b * band + (r - *nt * bs);
}

template <int CTAS, bool TN>
__global__ void __launch_bounds__(CTAS == 2 ? kThreads2 : kThreads1, 1)
gemm_bf16_nt_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    void* __restrict__ Cout, const float* __restrict__ bias, int M, int N, int K, int ldc,
                    int c_bf16, int lstm_T, int lstm_B, int lstm_H, int splits, const int32_t* __restrict__ c_row_map) {
    // splits > 1 (split-K): work item = (tile, K range); item s of a tile writes its fp32 partial product, plain
    // layout [M][N], to Cout + s*M*N (a workspace); splitk_reduce_kernel adds the partials in a fixed order.
    constexpr int kStages = CTAS == 2 ? 6 : 4;
    constexpr uint32_t kStageBytesB = (BN / CTAS) * BK * 2;
    constexpr uint32_t kStageBytes = kStageBytesA + kStageBytesB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for SWIZZLE_128B tiles
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    PipeBars* bars = reinterpret_cast<PipeBars*>(smem + kStages * kStageBytes);
    constexpr int EW = CTAS == 2 ? 8 : 4;              // epilogue warps: warps w and w+4 share a TMEM lane quadrant and split the columns
    float* epi_smem = reinterpret_cast<float*>(smem + kStages * kStageBytes + ((sizeof(PipeBars) + 15) & ~15));   // EW x [32][32]

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
    const int num_m = (M + CTAS * BM - 1) / (CTAS * BM), num_n = (N + BN - 1) / BN;   // tiles of CTAS*128 x 256
    const int num_mn = num_m * num_n;
    const int num_tiles = num_mn * splits;
    const int num_kb_all = (K + BK - 1) / BK;
    const int kb_per = (num_kb_all + splits - 1) / splits;
    const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;
    constexpr int kBand = kBandRows / (CTAS * BM);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
        // tmem_empty: one arrival per epilogue warp of every CTA of the pair (the peer arrives remotely on the leader's)
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&bars->tmem_full[i], 1); mbar_init(&bars->tmem_empty[i], EW * CTAS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CTAS == 2) tmem_alloc_2cta(&bars->tmem_base, kTmemCols);
        else tmem_alloc(&bars->tmem_base, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();          // the peer's barriers are initialised before anything lands on them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);   // uniform register for the tcgen05 ops

    if (warp == 0) {
        // ===== TMA producer =====  (whole warp runs the loop, one elected lane issues: in a divergent
        // `if (lane == 0)` region ptxas wraps every UTMALDG / UTCHMMA in an ELECT/R2UR waterfall loop)
        uint32_t stage = 0, phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int mt, nt;
            tile_coords(tile % num_mn, num_m, num_n, kBand, &mt, &nt);
            const int m0 = (mt * CTAS + rank) * BM, n0 = nt * BN;
            const int kb0 = (tile / num_mn) * kb_per, kb1 = min(num_kb_all, kb0 + kb_per);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&bars->empty[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem + stage * kStageBytes;
                    if constexpr (CTAS == 2) {
                        // bytes of both CTAs' loads are counted on the leader's barrier; the leader arms it
                        const uint32_t bar = mapa_u32(smem_u32(&bars->full[stage]), 0);
                        if (rank == 0) mbar_expect_tx(&bars->full[stage], 2 * kStageBytes);
                        if constexpr (TN) {
                            // MN-major tiles: one box = 64 contiguous m (128 B swizzle row) x BK k-rows = 8 KB
#pragma unroll
                            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(&map_a, bar, sa + c * 8192, m0 + c * 64, kb * BK);
#pragma unroll
                            for (int c = 0; c < BN / 2 / 64; ++c)
                                tma_load_2d_2sm(&map_b, bar, sa + kStageBytesA + c * 8192, n0 + rank * (BN / 2) + c * 64, kb * BK);
                        } else {
                            tma_load_2d_2sm(&map_a, bar, sa, kb * BK, m0);
                            tma_load_2d_2sm(&map_b, bar, sa + kStageBytesA, kb * BK, n0 + rank * (BN / 2));
                        }
                    } else {
                        mbar_expect_tx(&bars->full[stage], kStageBytes);
                        if constexpr (TN) {
#pragma unroll
                            for (int c = 0; c < BM / 64; ++c) tma_load_2d(&map_a, &bars->full[stage], sa + c * 8192, m0 + c * 64, kb * BK);
#pragma unroll
                            for (int c = 0; c < BN / 64; ++c)
                                tma_load_2d(&map_b, &bars->full[stage], sa + kStageBytesA + c * 8192, n0 + c * 64, kb * BK);
                        } else {
                            tma_load_2d(&map_a, &bars->full[stage], sa, kb * BK, m0);
                            tma_load_2d(&map_b, &bars->full[stage], sa + kStageBytesA, kb * BK, n0);
                        }
                    }
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (pair: the leader CTA only) =====
        if (CTAS == 1 || rank == 0) {
            constexpr uint32_t idesc = make_idesc(CTAS * BM, BN) | (TN ? ((1u << 15) | (1u << 16)) : 0u);   // a_major / b_major = MN
            // descriptor advance per K = 16 step, in 16-byte units: 32 B inside the swizzle row (K-major) or
            // two 8-row groups of 1024 B (MN-major)
            constexpr uint64_t kstep = TN ? 128 : 2;
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                if constexpr (CTAS == 2) mbar_wait_cluster(&bars->tmem_empty[acc], acc_phase ^ 1);
                else mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int kb0 = (tile / num_mn) * kb_per, kb1 = min(num_kb_all, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t adesc = TN ? make_sw128_mn_desc(sa, 8192, 1024) : make_sw128_desc(sa);
                    const uint64_t bdesc = TN ? make_sw128_mn_desc(sa + kStageBytesA, 8192, 1024) : make_sw128_desc(sa + kStageBytesA);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if constexpr (CTAS == 2)
                                tc_mma_bf16_2cta(d_tmem, adesc + (uint64_t)k * kstep, bdesc + (uint64_t)k * kstep, idesc,
                                                 (uint32_t)(((kb - kb0) | k) != 0));
                            else
                                tc_mma_bf16(d_tmem, adesc + (uint64_t)k * kstep, bdesc + (uint64_t)k * kstep, idesc,
                                            (uint32_t)(((kb - kb0) | k) != 0));
                        }
                        // frees the smem slot (in both CTAs of a pair) when the MMAs retire
                        if constexpr (CTAS == 2) {
                            tc_commit_2cta(&bars->empty[stage], 3);
                            if (kb == kb1 - 1) tc_commit_2cta(&bars->tmem_full[acc], 3);
                        } else {
                            tc_commit(&bars->empty[stage]);
                            if (kb == kb1 - 1) tc_commit(&bars->tmem_full[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int ew = (warp - 4) & 3;              // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        const int c_begin = ((warp - 4) >> 2) * (BN / 32 / (EW / 4)), c_end = c_begin + BN / 32 / (EW / 4);
        uint32_t acc = 0, acc_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int mt, nt;
            tile_coords(tile % num_mn, num_m, num_n, kBand, &mt, &nt);
            const int m0 = (mt * CTAS + rank) * BM, n0 = nt * BN;
            mbar_wait(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            // Coalesced epilogue: the accumulator chunk arrives with lane = row (32 columns per lane); it is
            // transposed through a per-warp shared-memory tile so that one store instruction writes 4 rows x 128 B
            // (8 lanes x 16 B per row) instead of 32 rows x 16 B: the row-per-lane stores cost one sector
            // transaction per lane and made the K = 1024 shapes epilogue-bound (profiles/kernel_bench_gemm_r1_v22).
            const int row = m0 + ew * 32 + lane;
            float* tsm = epi_smem + (warp - 4) * (32 * 32);
            void* Cout_t = splits > 1 ? static_cast<void*>(reinterpret_cast<float*>(Cout) + (int64_t)(tile / num_mn) * M * N) : Cout;
            // transposed phase: this lane stores rows sub_r + 4i (i < 8), columns sub_c..sub_c+3 of every 32-column
            // chunk.  Destination = rowbase[i] + chunk term (both layouts are separable in row and column).
            const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
            int64_t rowbase[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = m0 + ew * 32 + sub_r + 4 * i;
                if (rr >= M) rowbase[i] = -1;
                else if (lstm_T > 0) {
                    // LSTM gate layout [T][2][H/32][B][128]: row = b*T + t
                    const int lb = rr / lstm_T, lt = rr - lb * lstm_T;
                    rowbase[i] = ((int64_t)lt * 2 * (lstm_H >> 5) * lstm_B + lb) * 128;
                } else rowbase[i] = (int64_t)((c_row_map && splits == 1) ? __ldg(&c_row_map[rr]) : rr) * ldc;
            }
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + acc * BN + c * 32 + ((uint32_t)(ew * 32) << 16);
                tc_ld_32x32b_x32(taddr, v);
                const int col0 = n0 + c * 32;
                if (col0 >= N) continue;                               // warp-uniform
                if (col0 + 32 <= N) {
                    int64_t cterm = col0;
                    if (lstm_T > 0) {
                        // col = dir*4H + gate*H + unit
                        const int dir = col0 / (4 * lstm_H), r = col0 - dir * 4 * lstm_H;
                        const int gate = r / lstm_H, unit = r - gate * lstm_H;
                        cterm = ((int64_t)(dir * (lstm_H >> 5) + (unit >> 5)) * lstm_B) * 128 + gate * 32;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(tsm + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    __syncwarp();
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias) bv = make_float4(__ldg(&bias[col0 + sub_c]), __ldg(&bias[col0 + sub_c + 1]),
                                               __ldg(&bias[col0 + sub_c + 2]), __ldg(&bias[col0 + sub_c + 3]));
                    float4 x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(tsm + (sub_r + 4 * i) * 32 + ((((lane & 7)) ^ ((sub_r + 4 * i) & 7)) << 2));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        x[i].x += bv.x; x[i].y += bv.y; x[i].z += bv.z; x[i].w += bv.w;
                        if (rowbase[i] >= 0) {
                            const int64_t o = rowbase[i] + cterm + sub_c;
                            if (c_bf16) {
                                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Cout_t) + o;
                                __nv_bfloat162 p0 = __floats2bfloat162_rn(x[i].x, x[i].y), p1 = __floats2bfloat162_rn(x[i].z, x[i].w);
                                if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                                    uint2 w;
                                    w.x = *reinterpret_cast<uint32_t*>(&p0); w.y = *reinterpret_cast<uint32_t*>(&p1);
                                    *reinterpret_cast<uint2*>(dst) = w;
                                } else {
                                    dst[0] = __float2bfloat16(x[i].x); dst[1] = __float2bfloat16(x[i].y);
                                    dst[2] = __float2bfloat16(x[i].z); dst[3] = __float2bfloat16(x[i].w);
                                }
                            } else {
                                float* dst = reinterpret_cast<float*>(Cout_t) + o;
                                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) *reinterpret_cast<float4*>(dst) = x[i];
                                else { dst[0] = x[i].x; dst[1] = x[i].y; dst[2] = x[i].z; dst[3] = x[i].w; }
                            }
                        }
                    }
                } else if (row < M) {
                    const int64_t orow = (c_row_map && splits == 1) ? __ldg(&c_row_map[row]) : row;
                    for (int j = 0; j < 32 && col0 + j < N; ++j) {
                        const float f = __uint_as_float(v[j]) + (bias ? __ldg(&bias[col0 + j]) : 0.f);
                        if (c_bf16) reinterpret_cast<__nv_bfloat16*>(Cout_t)[orow * ldc + col0 + j] = __float2bfloat16(f);
                        else reinterpret_cast<float*>(Cout_t)[orow * ldc + col0 + j] = f;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // relaxed: only the tcgen05.ld reads (complete after wait::ld) must precede the leader's next MMAs; a
                // release here would wait for all of this warp's global stores
                if (CTAS == 2 && rank != 0) mbar_arrive_remote_relaxed(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
                else mbar_arrive(&bars->tmem_empty[acc]);
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();          // the leader's last MMAs / commits touch the peer
    if (warp == 2) {
        if constexpr (CTAS == 2) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// C[m, n] = sum_s partial[s][m][n] (+ bias[n]), fixed order; fp32 or bf16 output with leading dimension ldc
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t MN, int N, const float* __restrict__ bias,
                                     void* __restrict__ C, int ldc, int c_bf16, const int32_t* __restrict__ c_row_map) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < MN; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = part[i];
        for (int s2 = 1; s2 < splits; ++s2) acc += part[(int64_t)s2 * MN + i];
        int64_t m = i / N;
        const int n = (int)(i - m * N);
        if (c_row_map) m = c_row_map[m];
        if (bias) acc += bias[n];
        if (c_bf16) reinterpret_cast<__nv_bfloat16*>(C)[m * ldc + n] = __float2bfloat16(acc);
        else reinterpret_cast<float*>(C)[m * ldc + n] = acc;
    }
}

struct SplitWs { float* ptr = nullptr; size_t bytes = 0; };
std::map<cudaStream_t, SplitWs> g_split_ws;     // one workspace per stream (GEMMs on different streams overlap)
std::mutex g_split_mu;

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

template <bool TN>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mb2, bool pair, void* C, const float* bias,
                int M, int N, int K, int ldc, int c_bf16, const int32_t* c_row_map, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        PK2_CHECK(cudaFuncSetAttribute(gemm_bf16_nt_kernel<1, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        PK2_CHECK(cudaFuncSetAttribute(gemm_bf16_nt_kernel<2, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr = true;
    }
    if (pair) {
        constexpr int EWH = 8;
        const size_t smem = 6 * (kStageBytesA + (BN / 2) * BK * 2) + sizeof(PipeBars) + 16 + EWH * kEpiTile + 1024;
        const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
        int max_pairs = g_num_sms / 2;
        if (g_max_ctas > 0 && max_pairs > g_max_ctas / 2) max_pairs = g_max_ctas / 2 > 0 ? g_max_ctas / 2 : 1;
        // split-K: few tiles with a long contraction (the recurrent weight gradients: 16 tiles, K = B*T) would leave
        // most SMs idle; partial products go to a per-stream workspace and are added in a fixed order
        const int num_kb = (K + BK - 1) / BK;
        int splits = 1;
        static const bool no_split = getenv("PK2_GEMM_NO_SPLITK") != nullptr;
        if (!no_split && g_lstm_T == 0 && tiles * 2 <= max_pairs && num_kb >= 64) {
            splits = max_pairs / tiles;
            if (splits > 8) splits = 8;
            if (splits > num_kb / 16) splits = num_kb / 16;
            if (splits < 1) splits = 1;
        }
        float* ws = nullptr;
        if (splits > 1) {
            const size_t need = (size_t)splits * M * N * sizeof(float);
            std::lock_guard<std::mutex> lk(g_split_mu);
            SplitWs& w = g_split_ws[st];
            if (need > w.bytes) {
                if (w.ptr) { PK2_CHECK(cudaStreamSynchronize(st)); cudaFree(w.ptr); }
                PK2_CHECK(cudaMalloc(&w.ptr, need));
                w.bytes = need;
            }
            ws = w.ptr;
        }
        const int items = tiles * splits;
        const int pairs = items < max_pairs ? items : max_pairs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(kThreads2);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (splits > 1) {
            PK2_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_nt_kernel<2, TN>, ma, mb2, static_cast<void*>(ws),
                                         static_cast<const float*>(nullptr), M, N, K, N, 0, 0, 0, 0, splits,
                                         static_cast<const int32_t*>(nullptr)));
            PK2_LAUNCHED();
            const int64_t MN = (int64_t)M * N;
            const int rb = (int)((MN + 255) / 256 < 2048 ? (MN + 255) / 256 : 2048);
            splitk_reduce_kernel<<<rb, 256, 0, st>>>(ws, splits, MN, N, bias, C, ldc, c_bf16, c_row_map);
        } else {
            PK2_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_nt_kernel<2, TN>, ma, mb2, C, bias, M, N, K, ldc, c_bf16,
                                         (int)g_lstm_T, (int)g_lstm_B, (int)g_lstm_H, 1, c_row_map));
        }
    } else {
        constexpr int EWH = 4;
        const size_t smem = 4 * (kStageBytesA + BN * BK * 2) + sizeof(PipeBars) + 16 + EWH * kEpiTile + 1024;
        const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
        int grid = tiles < g_num_sms ? tiles : g_num_sms;
        if (g_max_ctas > 0 && grid > g_max_ctas) grid = g_max_ctas;
        gemm_bf16_nt_kernel<1, TN><<<grid, kThreads1, smem, st>>>(ma, mb, C, bias, M, N, K, ldc, c_bf16,
                                                                  g_lstm_T, g_lstm_B, g_lstm_H, 1, c_row_map);
    }
    PK2_POST_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int pk2_gemm_bf16_ex(const void* A, const void* B, void* C, const float* bias, int M, int N, int K,
                                int lda, int ldb, int ldc, int flags, const int32_t* c_row_map, void* stream) {
    PK2_REQUIRE(A && B && C, "pk2_gemm_bf16: null argument");
    PK2_REQUIRE(M > 0 && N > 0 && K > 0, "pk2_gemm_bf16: empty problem %dx%dx%d", M, N, K);
    PK2_REQUIRE((flags & 1) == 0, "pk2_gemm_bf16: accumulate flag not supported");
    PK2_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "pk2_gemm_bf16: lda/ldb must be multiples of 8 elements (16 B)");
    PK2_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
                "pk2_gemm_bf16: A/B must be 16-byte aligned");
    const bool tn = (flags & PK2_GEMM_TN) != 0;
    PK2_REQUIRE(!tn || g_lstm_T == 0, "pk2_gemm_bf16: TN form has no LSTM gate-layout epilogue");
    PK2_REQUIRE(!tn || (lda >= M && ldb >= N), "pk2_gemm_bf16: TN form needs lda >= M and ldb >= N (row strides of At[K,M], Bt[K,N])");
    if (g_num_sms == 0) {
        int dev = 0;
        PK2_CHECK(cudaGetDevice(&dev));
        PK2_CHECK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // CTA pairs (256 x 256 tiles) unless the problem is a single row of tiles or PK2_GEMM_1CTA is set
    static const bool force1 = getenv("PK2_GEMM_1CTA") != nullptr;
    const bool pair = !force1 && M > BM && g_num_sms >= 2;
    const int c_bf16 = (flags >> 1) & 1;
    CUtensorMap ma, mb, mb2;
    cudaStream_t st = pk2::as_stream(stream);
    if (tn) {
        // operands [K rows][M or N contiguous]: boxes of 64 contiguous elements x BK rows
        if (make_map(&ma, A, K, M, lda, BK)) return 2;
        if (make_map(&mb, B, K, N, ldb, BK)) return 2;
        mb2 = mb;
        return launch_gemm<true>(ma, mb, mb2, pair, C, bias, M, N, K, ldc, c_bf16, c_row_map, st);
    }
    if (make_map(&ma, A, M, K, lda, BM)) return 2;
    if (make_map(&mb, B, N, K, ldb, BN)) return 2;
    if (pair && make_map(&mb2, B, N, K, ldb, BN / 2)) return 2;
    return launch_gemm<false>(ma, mb, mb2, pair, C, bias, M, N, K, ldc, c_bf16, c_row_map, st);
}

extern "C" int pk2_gemm_bf16_nt(const void* A, const void* B, void* C, const float* bias, int M, int N, int K,
                                int lda, int ldb, int ldc, int flags, void* stream) {
    return pk2_gemm_bf16_ex(A, B, C, bias, M, N, K, lda, ldb, ldc, flags & ~PK2_GEMM_TN, nullptr, stream);
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
