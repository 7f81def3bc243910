// Persistent bidirectional-LSTM recurrence on tcgen05 tensor cores (sm_100a).
//
// Replaces the cuDNN RNN the reference reaches through nn.LSTM(batch_first, bidirectional)
// (models/lstm.py:46-58): for every layer, direction d and time step t
//     gates = Gx[t] + h_{t-1} W_hh^T ;  i,f,o = sigmoid, g = tanh ;  c_t = f c_{t-1} + i g ;  h_t = o tanh(c_t)
// Gx (= x W_ih^T + b_ih + b_hh for all t) is one large GEMM (gemm_tc.cu); this file is the
// serial part: T dependent step-GEMMs [B,H]x[H,4H] per direction.
//
// One launch per layer, all T steps, both directions:
//   grid = (H/32 CTAs, 2 directions, G batch groups of NB sequences), all CTAs co-resident.
//   CTA (c, d, g) owns hidden units [32c, 32c+32) of direction d for batch rows [g*NB, g*NB+NB):
//   * its 128 gate rows of W_hh (4 gates x 32 units, bf16, K-major SWIZZLE_128B) stay in shared
//     memory for the whole sequence;
//   * per step it computes gates^T[128, NB] = W_slice[128, H] * h_{t-1}^T on the tensor cores:
//     tcgen05.mma M=128, N=NB, K=16 issued by one thread, accumulator in TMEM; h_{t-1} (all H
//     columns, written by the H/32 CTAs of the direction) is fetched with TMA straight out of the
//     layer output tensor y[B,T,2H] (3-D tensor map -> swizzled shared tile = the B operand);
//   * the epilogue (tcgen05.ld, + Gx tile prefetched by a bulk copy, activations, cell update)
//     keeps c_t in registers, writes h_t (bf16) into y and the gates / cell state for backward;
//   * step t+1 of the other CTAs is released through a global arrive counter
//     (red.release.gpu / ld.acquire.gpu) -- no grid-wide barrier, no per-step launch.
// The backward kernel has the same structure with the roles swapped: dgates_t (all 4H columns,
// bf16) is the streamed A operand, W_hh^T[32 units, 4H] the resident B operand.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <mutex>
#include <stdlib.h>

using namespace tc;

namespace {

constexpr int kEpiThreads = 128;
constexpr int kThreads = 192;          // warps 0-3 epilogue, warp 4 producer, warp 5 MMA issuer

struct FwdDev {
    int B, T, H;
    const float* gx;            // [T][2][H/32][B][128]
    __nv_bfloat16* y;           // [B][T][2H]
    __nv_bfloat16* gates;       // [2][T][B][4][H]   post-activation
    float* cstate;              // [2][T][B][H]
    unsigned* counters;         // [G][2]
    long long* prof;            // optional per-step clock64 trace of CTA (0,0,0) (profiling only)
    int a_tmem;                 // 1: keep the W_hh slice (A operand) in tensor memory instead of shared memory
};

struct BwdDev {
    int B, T, H;
    const float* dy;            // [B][T][2H] fp32
    const __nv_bfloat16* gates;
    const float* cstate;
    __nv_bfloat16* dgates;      // [B][T][2][4H]
    unsigned* counters;
    long long* prof;            // optional clock64 trace (profiling only)
    int dbg;                    // PK2_LSTM_RS_DBG experiment bits (profiling only)
};

// tanh.approx.f32: one MUFU op, max relative error 2^-11 -- below the bf16 rounding the gates and h
// receive anyway.  sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5.
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_fast(float x) {
    // 2*sigmoid(2x) - 1, exact enough for the 1e-3 parity budget and safe for large |x|
    return __fdividef(2.0f, 1.0f + __expf(-2.0f * x)) - 1.0f;
}

__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
    while (ld_acquire_gpu(ctr) < target) { }
}

// Batch rows are processed in groups of NB = 32 (the MMA's N).  Each CTA interleaves NG groups: while
// the h_t of group A travels (publish -> peers' acquire -> TMA), the CTA runs the MMA + epilogue of
// group B, so the inter-CTA hand-off latency is hidden behind useful work.
constexpr int NB = 32;

// --------------------------------------------------------------------------- forward ----
template <int NG>
__global__ void __launch_bounds__(kThreads, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y, FwdDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const int H = p.H, T = p.T, B = p.B;
    const int KB = H / 64;                                   // k-blocks of 64
    uint8_t* Ws = smem;                                      // KB x [128 x 64] bf16
    uint8_t* Hs0 = Ws + KB * 16384;                          // NG x KB x [NB x 64] bf16
    const int hs_bytes = KB * NB * 128;
    float* gxs0 = reinterpret_cast<float*>(Hs0 + NG * hs_bytes);   // NG x [NB][128] fp32
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(gxs0) + NG * NB * 512);
    uint64_t* wbar = bars + 0;          // W slice landed
    uint64_t* hbar = bars + 1;          // [NG] h_{t-1} tiles landed
    uint64_t* gbar = hbar + NG;         // [NG] Gx tile landed
    uint64_t* mbar = gbar + NG;         // [NG] step MMAs retired
    uint64_t* gfree = mbar + NG;        // [NG] Gx buffer may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gfree + NG);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int cta = blockIdx.x, dir = blockIdx.y;
    const int nctas = gridDim.x;
    const int u0 = cta * 32;
    const int grp0 = blockIdx.z * NG;                        // first batch group of this CTA
    int nga = 0;                                             // groups that hold at least one row
    for (int g = 0; g < NG; ++g) nga += ((grp0 + g) * NB < B) ? 1 : 0;

    if (threadIdx.x == 0) {
        mbar_init(wbar, 1);
        for (int g = 0; g < NG; ++g) { mbar_init(&hbar[g], 1); mbar_init(&gbar[g], 1); mbar_init(&mbar[g], 1); mbar_init(&gfree[g], 1); }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, NG * NB < 32 ? 32 : NG * NB);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // uniform register: no R2UR waterfall per tcgen05 op

    if (warp == 4) {
        // ===== producer: TMA / bulk loads + inter-CTA step flags =====
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)KB * 16384u);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(&map_w, wbar, Ws + kb * 16384, kb * 64, (dir * nctas + cta) * 128);
            uint32_t ph_free = 0;
            for (int s = 0; s < T; ++s) {
                const int tt = dir ? (T - 1 - s) : s;
                for (int g = 0; g < nga; ++g) {
                    const int b0 = (grp0 + g) * NB, nbv = min(NB, B - b0);
                    if (s > 0) mbar_wait(&gfree[g], ph_free);
                    const float* src = p.gx + ((((int64_t)tt * 2 + dir) * nctas + cta) * B + b0) * 128;
                    mbar_expect_tx(&gbar[g], (uint32_t)nbv * 512u);
                    bulk_load(gxs0 + g * NB * 128, src, (uint32_t)nbv * 512u, &gbar[g]);
                    if (s > 0) {
                        const int tp = dir ? tt + 1 : tt - 1;
                        spin_until(p.counters + ((grp0 + g) * 2 + dir), (unsigned)(nctas * s));
                        fence_proxy_async();
                        mbar_expect_tx(&hbar[g], (uint32_t)hs_bytes);
                        // one 4-D box {64 cols, NB rows, 1 step, KB k-blocks} -> KB swizzled [NB x 64] tiles
                        tma_load_4d(&map_y, &hbar[g], Hs0 + g * hs_bytes, 0, b0, tp, dir * KB);
                    }
                }
                if (s > 0) ph_free ^= 1;
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        {   // whole warp in the loop, one elected lane issues (see lstm_fwd_cluster_kernel)
            const uint32_t idesc = make_idesc(128, NB);
            mbar_wait(wbar, 0);
            const uint64_t ad0 = make_sw128_desc(smem_u32(Ws));
            uint32_t ph_h = 0;
            for (int s = 1; s < T; ++s) {
                for (int g = 0; g < nga; ++g) {
                    mbar_wait(&hbar[g], ph_h);
                    tc_fence_after();
                    const uint64_t bd0 = make_sw128_desc(smem_u32(Hs0 + g * hs_bytes));
                    if (elect_one()) {
                        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                tc_mma_bf16(tmem_base + g * NB, ad0 + (uint64_t)(kb * (16384 / 16) + k * 2),
                                            bd0 + (uint64_t)(kb * (NB * 128 / 16) + k * 2), idesc, (uint32_t)((kb | k) != 0));
                        }
                        tc_commit(&mbar[g]);
                    }
                    __syncwarp();
                }
                ph_h ^= 1;
            }
        }
    } else {
        // ===== epilogue warps 0-3: thread r = gate row (gate = r/32, unit = r%32) = TMEM lane r =====
        const int r = threadIdx.x;
        const int gate = warp;
        float cst[NG][NB / 4];
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int k = 0; k < NB / 4; ++k) cst[g][k] = 0.f;
        uint32_t ph_g = 0, ph_m = 0;
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? (T - 1 - s) : s;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g >= nga) break;
                const int b0 = (grp0 + g) * NB, nbv = min(NB, B - b0);
                float* gxs = gxs0 + g * NB * 128;
                float acc[NB];
                if (s > 0) {
                    mbar_wait(&mbar[g], ph_m);
                    tc_fence_after();
                    uint32_t v[32];
                    tc_ld_32x32b_x32(tmem_base + g * NB + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(v[j]);
                    tc_fence_before();
                } else {
#pragma unroll
                    for (int j = 0; j < NB; ++j) acc[j] = 0.f;
                }
                mbar_wait(&gbar[g], ph_g);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float v = acc[b] + gxs[b * 128 + r];
                    gxs[b * 128 + r] = (gate == 2) ? tanhf_fast(v) : sigmoidf_fast(v);
                }
                named_bar_sync(1, kEpiThreads);
                // cell update: lane = unit, warp w handles batch rows w, w+4, ...  (critical path: h_t -> y)
                {
                    __nv_bfloat16* yo = p.y + ((int64_t)b0 * T + tt) * 2 * H + dir * H + u0 + lane;
#pragma unroll
                    for (int k = 0; k < NB / 4; ++k) {
                        const int b = warp + 4 * k;
                        const float ig = gxs[b * 128 + lane], fg = gxs[b * 128 + 32 + lane];
                        const float gg = gxs[b * 128 + 64 + lane], og = gxs[b * 128 + 96 + lane];
                        const float c = fmaf(fg, cst[g][k], ig * gg);
                        cst[g][k] = c;
                        const float h = og * tanhf_fast(c);
                        if (b < nbv) yo[(int64_t)b * T * 2 * H] = __float2bfloat16(h);
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, kEpiThreads);
                if (threadIdx.x == 0) red_release_gpu_add(p.counters + ((grp0 + g) * 2 + dir), 1u);   // publish h_t
                // off the critical path: save gates and cell state for the backward pass
                {
                    float* co = p.cstate + (((int64_t)dir * T + tt) * B + b0) * H + u0 + lane;
                    __nv_bfloat16* gout = p.gates + (((int64_t)dir * T + tt) * B + b0) * 4 * H + u0 + lane;
#pragma unroll
                    for (int k = 0; k < NB / 4; ++k) {
                        const int b = warp + 4 * k;
                        if (b < nbv) {
                            co[(int64_t)b * H] = cst[g][k];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                gout[((int64_t)b * 4 + q) * H] = __float2bfloat16(gxs[b * 128 + q * 32 + lane]);
                        }
                    }
                }
                named_bar_sync(1, kEpiThreads);
                if (threadIdx.x == 0) mbar_arrive(&gfree[g]);
            }
            ph_g ^= 1;
            if (s > 0) ph_m ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, NG * NB < 32 ? 32 : NG * NB);
}

// ------------------------------------------------------------ forward, cluster + DSMEM ----
// Same recurrence, but the H/32 CTAs of one (direction, batch group) form ONE thread-block cluster and
// exchange h_t through distributed shared memory instead of global memory: after the cell update each
// CTA stages its [NB x 32] bf16 slice and sends it to every CTA of the cluster with one bulk copy
// (cp.async.bulk.shared::cluster) whose completion bytes land on the RECEIVER's mbarrier.  The receiver's
// MMA warp simply waits on that mbarrier: no global flag, no acquire poll, no TMA round trip through L2
// on the critical path.  The h operand uses the SWIZZLE_NONE K-major layout
// [K/8 chunks][NB/8 row groups][8 rows][16 B] so that each sender's slice is one contiguous 2 KB block.
// h tiles are double buffered (a sender may run one step ahead of a receiver's MMA).
// EW epilogue warps (4 or 8): warps w and w+4 share TMEM lane quadrant w&3 (= gate) and split the batch
// columns; ACC independent accumulators (1 or 4) summed in the epilogue.
template <int EW, int ACC>
__global__ void __launch_bounds__((EW + 2) * 32, 1)
lstm_fwd_cluster_kernel(const __grid_constant__ CUtensorMap map_w, FwdDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const int H = p.H, T = p.T, B = p.B;
    const int KB = H / 64;
    constexpr int kChunk = NB / 8 * 128;                     // bytes of one 16-byte K-chunk over NB rows
    constexpr int kSlice = NB * 64;                          // bytes one CTA contributes per step
    constexpr int CPW = NB / (EW / 4);                       // batch columns per epilogue warp in phase 1
    constexpr int RPW = NB / EW;                             // batch rows per epilogue warp in phase 2
    constexpr int kEpi = EW * 32;
    const int hs_bytes = H * NB * 2;
    uint8_t* Ws = smem;                                      // KB x [128 x 64] bf16, SWIZZLE_128B
    uint8_t* Hb = Ws + KB * 16384;                           // 2 x [H/8 chunks][NB/8][8][16 B]
    float* gxs = reinterpret_cast<float*>(Hb + 2 * hs_bytes);    // [NB][128] fp32
    uint8_t* stg = reinterpret_cast<uint8_t*>(gxs) + NB * 512;   // 2 x kSlice
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + 2 * kSlice);
    uint64_t* wbar = bars + 0;
    uint64_t* hfull = bars + 1;          // [2]
    uint64_t* gbar = bars + 3;
    uint64_t* mbar = bars + 4;
    uint64_t* gfree = bars + 5;
    uint64_t* abar = bars + 6;           // W slice copied into tensor memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int CS = gridDim.x;                                // cluster size = H/32
    const int cta = (int)cluster_ctarank();                  // == blockIdx.x
    const int dir = blockIdx.y, grp = blockIdx.z;
    const int u0 = cta * 32, b0 = grp * NB;
    const int nbv = min(NB, B - b0);
    const uint32_t step_bytes = (uint32_t)CS * kSlice;
    long long* prof = (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.prof : nullptr;
#define PK2_PROF(e) do { if (prof && s >= 64 && s < 72) prof[(s - 64) * 16 + (e)] = clock64(); } while (0)

    if (threadIdx.x == 0) {
        mbar_init(wbar, 1); mbar_init(&hfull[0], 1); mbar_init(&hfull[1], 1);
        mbar_init(gbar, 1); mbar_init(mbar, 1); mbar_init(gfree, 1); mbar_init(abar, EW);
        fence_barrier_init();
        if (T >= 2) mbar_expect_tx(&hfull[0], step_bytes);    // h_0
        if (T >= 3) mbar_expect_tx(&hfull[1], step_bytes);    // h_1
    }
    for (int i = threadIdx.x; i < NB * 128; i += (EW + 2) * 32) gxs[i] = 0.f;   // padded batch rows stay finite
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // before the bulk copies write gxs
    // D: ACC accumulators of NB columns; A (W_hh slice, bf16 pairs): H/2 columns at column 256
    const uint32_t tmem_cols = p.a_tmem ? 512u : 128u;
    if (warp == EW + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                      // every CTA's barriers are initialised
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // uniform register: no R2UR waterfall per tcgen05 op

    if (warp == EW) {
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)KB * 16384u);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(&map_w, wbar, Ws + kb * 16384, kb * 64, (dir * CS + cta) * 128);
            uint32_t ph_free = 0;
            for (int s = 0; s < T; ++s) {
                const int tt = dir ? (T - 1 - s) : s;
                if (s > 0) { mbar_wait(gfree, ph_free); ph_free ^= 1; }
                const float* src = p.gx + ((((int64_t)tt * 2 + dir) * CS + cta) * B + b0) * 128;
                mbar_expect_tx(gbar, (uint32_t)nbv * 512u);
                bulk_load(gxs, src, (uint32_t)nbv * 512u, gbar);
            }
        }
    } else if (warp == EW + 1) {
        // The whole warp runs this loop (warp-uniform control flow, operands in uniform registers); one elected
        // lane issues.  Under `if (lane == 0)` every tcgen05.mma cost ~92 cycles of ELECT/R2UR waterfall.
        const uint32_t idesc = make_idesc(128, NB);
        mbar_wait(wbar, 0);
        if (p.a_tmem) { mbar_wait(abar, 0); }
        tc_fence_after();
        const uint64_t ad0 = make_sw128_desc(smem_u32(Ws));
        const bool a_tmem = p.a_tmem != 0;
        for (int s = 1; s < T; ++s) {
            const int buf = (s - 1) & 1;
            mbar_wait(&hfull[buf], (uint32_t)(((s - 1) >> 1) & 1));      // h_{s-1} from all CTAs
            if (lane == 0) PK2_PROF(0);
            tc_fence_after();
            // descriptors are affine in kk: bases built once, one 64-bit add per MMA
            const uint64_t bd0 = make_nosw_desc(smem_u32(Hb + buf * hs_bytes), kChunk, 128);
            if (elect_one()) {
                if (s + 1 <= T - 2) mbar_expect_tx(&hfull[buf], step_bytes);  // re-arm for h_{s+1}
                if (a_tmem) {
#pragma unroll 8
                    for (int kk = 0; kk < H / 16; ++kk)
                        tc_mma_bf16_ts(tmem_base + (uint32_t)(kk & (ACC - 1)) * NB, tmem_base + 256 + kk * 8,
                                       bd0 + (uint64_t)(kk * (2 * kChunk / 16)), idesc, (uint32_t)(kk >= ACC));
                } else {
#pragma unroll 8
                    for (int kk = 0; kk < H / 16; ++kk)
                        tc_mma_bf16(tmem_base + (uint32_t)(kk & (ACC - 1)) * NB,
                                    ad0 + (uint64_t)((kk >> 2) * (16384 / 16) + (kk & 3) * 2),
                                    bd0 + (uint64_t)(kk * (2 * kChunk / 16)), idesc, (uint32_t)(kk >= ACC));
                }
                tc_commit(mbar);
            }
            __syncwarp();
            if (lane == 0) PK2_PROF(1);
        }
    } else {
        const int gate = warp & 3;                               // TMEM lane quadrant = gate
        const int r = gate * 32 + lane;                          // gate row of this thread (phase 1)
        const int c0 = (warp >> 2) * CPW;                        // first batch column of this warp (phase 1)
        float cst[RPW];
#pragma unroll
        for (int k = 0; k < RPW; ++k) cst[k] = 0.f;
        if (p.a_tmem) {
            // one-time: my gate row of the W_hh slice -> tensor memory lane r, two bf16 per 32-bit column,
            // read back out of the 128B-swizzled shared-memory tiles the TMA wrote (warps sharing a quadrant
            // split the k-blocks)
            mbar_wait(wbar, 0);
            const int kpw = KB / (EW / 4);
            for (int kb = (warp >> 2) * kpw; kb < (warp >> 2) * kpw + kpw; ++kb) {
                uint32_t w[32];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint4 q = *reinterpret_cast<const uint4*>(Ws + kb * 16384 + r * 128 + ((ch ^ (r & 7)) << 4));
                    w[ch * 4 + 0] = q.x; w[ch * 4 + 1] = q.y; w[ch * 4 + 2] = q.z; w[ch * 4 + 3] = q.w;
                }
                tc_st_32x32b_x32(tmem_base + 256 + kb * 32 + ((uint32_t)(gate * 32) << 16), w);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(abar);
        }
        uint32_t ph_g = 0, ph_m = 0;
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? (T - 1 - s) : s;
            float acc[CPW];
            if (s > 0) {
                mbar_wait(mbar, ph_m); ph_m ^= 1;
                if (threadIdx.x == 0) PK2_PROF(2);
                tc_fence_after();
                uint32_t v[ACC][CPW];
#pragma unroll
                for (int q = 0; q < ACC; ++q) {
                    const uint32_t ta = tmem_base + q * NB + c0 + ((uint32_t)(gate * 32) << 16);
                    if constexpr (CPW == 32) tc_ld_32x32b_x32_nowait(ta, v[q]);
                    else tc_ld_32x32b_x16_nowait(ta, v[q]);
                }
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < CPW; ++j) {
                    float t = __uint_as_float(v[0][j]);
#pragma unroll
                    for (int q = 1; q < ACC; ++q) t += __uint_as_float(v[q][j]);
                    acc[j] = t;
                }
                tc_fence_before();
            } else {
#pragma unroll
                for (int j = 0; j < CPW; ++j) acc[j] = 0.f;
            }
            if (threadIdx.x == 0) PK2_PROF(3);
            mbar_wait(gbar, ph_g); ph_g ^= 1;
            if (threadIdx.x == 0) PK2_PROF(4);
            // all loads first, then independent activations, then all stores (the in-place
            // read-modify-write loop serialised on the load->MUFU->store latency chain: 97 cycles/element)
            {
                float gxr[CPW];
#pragma unroll
                for (int b = 0; b < CPW; ++b) gxr[b] = gxs[(c0 + b) * 128 + r];
                if (gate == 2) {
#pragma unroll
                    for (int b = 0; b < CPW; ++b) gxr[b] = tanh_approx(acc[b] + gxr[b]);
                } else {
#pragma unroll
                    for (int b = 0; b < CPW; ++b) gxr[b] = sigmoid_approx(acc[b] + gxr[b]);
                }
#pragma unroll
                for (int b = 0; b < CPW; ++b) gxs[(c0 + b) * 128 + r] = gxr[b];
            }
            named_bar_sync(1, kEpi);
            if (threadIdx.x == 0) PK2_PROF(5);
            // cell update (lane = unit, warp w handles rows w, w+EW, ...); h_t goes to the staging slice.
            // The activations are copied to registers so the Gx buffer can be refilled right away.
            __nv_bfloat16 hv[RPW];
            __nv_bfloat16 gv[RPW][4];
            uint8_t* st = stg + (s & 1) * kSlice;
            {
                float gi[RPW], gf[RPW], gg_[RPW], go[RPW], tc_[RPW];
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int b = warp + EW * k;
                    gi[k] = gxs[b * 128 + lane]; gf[k] = gxs[b * 128 + 32 + lane];
                    gg_[k] = gxs[b * 128 + 64 + lane]; go[k] = gxs[b * 128 + 96 + lane];
                }
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    cst[k] = fmaf(gf[k], cst[k], gi[k] * gg_[k]);
                    tc_[k] = tanh_approx(cst[k]);
                }
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int b = warp + EW * k;
                    gv[k][0] = __float2bfloat16(gi[k]); gv[k][1] = __float2bfloat16(gf[k]);
                    gv[k][2] = __float2bfloat16(gg_[k]); gv[k][3] = __float2bfloat16(go[k]);
                    hv[k] = __float2bfloat16(go[k] * tc_[k]);
                    *reinterpret_cast<__nv_bfloat16*>(st + (lane >> 3) * kChunk + (b >> 3) * 128 + (b & 7) * 16 + (lane & 7) * 2) = hv[k];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_bar_sync(1, kEpi);
            if (threadIdx.x == 0) { PK2_PROF(6); mbar_arrive(gfree); }   // Gx buffer consumed: prefetch the next step now
            constexpr int kStride = kEpi / 16;
            if (s < T - 1 && (threadIdx.x % kStride) == 0 && (int)(threadIdx.x / kStride) < CS) {
                // my slice -> CTA `rank` of the cluster (Hb[s&1] + cta*kSlice there); bytes are counted on its
                // hfull[s&1].  The copies are issued by 16 threads spread over the epilogue warps.
                const uint32_t rank = threadIdx.x / kStride;
                const uint32_t dst = mapa_u32(smem_u32(Hb + (s & 1) * hs_bytes + cta * kSlice), rank);
                const uint32_t bar = mapa_u32(smem_u32(&hfull[s & 1]), rank);
                dsmem_bulk_copy(dst, smem_u32(st), (uint32_t)kSlice, bar);
            }
            if (threadIdx.x == 0) PK2_PROF(7);
            // off the critical path: layer output, gates and cell state to global memory
            {
                __nv_bfloat16* yo = p.y + ((int64_t)b0 * T + tt) * 2 * H + dir * H + u0 + lane;
                float* co = p.cstate + (((int64_t)dir * T + tt) * B + b0) * H + u0 + lane;
                __nv_bfloat16* gout = p.gates + (((int64_t)dir * T + tt) * B + b0) * 4 * H + u0 + lane;
#pragma unroll
                for (int k = 0; k < RPW; ++k) {
                    const int b = warp + EW * k;
                    if (b < nbv) {
                        yo[(int64_t)b * T * 2 * H] = hv[k];
                        co[(int64_t)b * H] = cst[k];
#pragma unroll
                        for (int q = 0; q < 4; ++q) gout[((int64_t)b * 4 + q) * H] = gv[k][q];
                    }
                }
            }
            if (threadIdx.x == 0) PK2_PROF(8);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // peers have consumed everything this CTA sent
    if (warp == EW + 1) tmem_dealloc(tmem_base, tmem_cols);
#undef PK2_PROF
}

// -------------------------------------------------------------------------- backward ----
// dgates_t of the previous step is streamed as the A operand in chunks of CH k-blocks of [NB x 64] bf16
// through a 2-deep ring (shared by the NG interleaved groups); W_hh^T[32 units, 4H] is the resident B operand.
constexpr int kChunkBytes = 32768;
constexpr int kRing = 2;
constexpr int kASlack = 16384;           // the M=128 MMA reads 16 KB from each tile base (rows >= NB ignored)

template <int NG>
__global__ void __launch_bounds__(kThreads, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap map_wt, const __grid_constant__ CUtensorMap map_dg, BwdDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const int H = p.H, T = p.T, B = p.B;
    const int KB = 4 * H / 64;                               // k-blocks over the 4H gate columns
    const int CH = min(kChunkBytes / (NB * 128), KB);        // k-blocks per chunk
    const int NCH = KB / CH;                                 // chunks per step
    const uint32_t chunk_bytes = (uint32_t)CH * NB * 128u;
    uint8_t* Wt = smem;                                      // KB x [32 x 64] bf16   (W_hh^T slice)
    uint8_t* As = Wt + KB * 4096;                            // kRing chunks + slack
    float* dhs0 = reinterpret_cast<float*>(As + kRing * kChunkBytes + kASlack);   // NG x [NB][33]
    constexpr int kDhs = NB * 33;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(dhs0) + ((NG * kDhs * 4 + 7) & ~7));
    uint64_t* wbar = bars + 0;
    uint64_t* mbar = bars + 1;                               // [NG]
    uint64_t* afull = mbar + NG;                             // [kRing]
    uint64_t* aempty = afull + kRing;                        // [kRing]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty + kRing);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int cta = blockIdx.x, dir = blockIdx.y;
    const int nctas = gridDim.x;
    const int u0 = cta * 32;
    const int grp0 = blockIdx.z * NG;
    int nga = 0;
    for (int g = 0; g < NG; ++g) nga += ((grp0 + g) * NB < B) ? 1 : 0;

    if (threadIdx.x == 0) {
        mbar_init(wbar, 1);
        for (int g = 0; g < NG; ++g) mbar_init(&mbar[g], 1);
        for (int i = 0; i < kRing; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, NG * 32 < 32 ? 32 : NG * 32);
    // the slack behind the ring is read (never written) by the MMAs: keep it finite
    for (int i = threadIdx.x; i < (kRing * kChunkBytes + kASlack) / 16; i += kThreads)
        reinterpret_cast<uint4*>(As)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // uniform register: no R2UR waterfall per tcgen05 op

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(wbar, (uint32_t)KB * 4096u);
            for (int kb = 0; kb < KB; ++kb)
                tma_load_2d(&map_wt, wbar, Wt + kb * 4096, kb * 64, dir * H + u0);
            uint32_t stage = 0, phase = 0;
            for (int s = 1; s < T; ++s) {
                // step s consumes dgates of the step processed before it (time tprev)
                const int tt = dir ? s : (T - 1 - s);
                const int tprev = dir ? tt - 1 : tt + 1;
                for (int g = 0; g < nga; ++g) {
                    spin_until(p.counters + ((grp0 + g) * 2 + dir), (unsigned)(nctas * s));
                    fence_proxy_async();
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait(&aempty[stage], phase ^ 1);
                        mbar_expect_tx(&afull[stage], chunk_bytes);
                        tma_load_4d(&map_dg, &afull[stage], As + stage * kChunkBytes, 0, (grp0 + g) * NB, tprev, dir * KB + c * CH);
                        if (++stage == kRing) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 5) {
        {   // whole warp in the loop, one elected lane issues (see lstm_fwd_cluster_kernel)
            const uint32_t idesc = make_idesc(128, 32);
            mbar_wait(wbar, 0);
            const uint64_t bdw = make_sw128_desc(smem_u32(Wt));
            uint32_t stage = 0, phase = 0;
            for (int s = 1; s < T; ++s) {
                for (int g = 0; g < nga; ++g) {
                    for (int c = 0; c < NCH; ++c) {
                        mbar_wait(&afull[stage], phase);
                        tc_fence_after();
                        const uint64_t ad0 = make_sw128_desc(smem_u32(As + stage * kChunkBytes));
                        const uint64_t bd0 = bdw + (uint64_t)(c * CH * (4096 / 16));
                        if (elect_one()) {
                            for (int q = 0; q < CH; ++q) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc_mma_bf16(tmem_base + g * 32, ad0 + (uint64_t)(q * (NB * 128 / 16) + k * 2),
                                                bd0 + (uint64_t)(q * (4096 / 16) + k * 2), idesc, (uint32_t)((c | q | k) != 0));
                            }
                            tc_commit(&aempty[stage]);
                            if (c == NCH - 1) tc_commit(&mbar[g]);
                        }
                        __syncwarp();
                        if (++stage == kRing) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===== epilogue: TMEM lane = batch row, column = unit; then lane = unit, warp strides batch rows =====
        float dcn[NG][NB / 4];
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int k = 0; k < NB / 4; ++k) dcn[g][k] = 0.f;
        uint32_t ph_m = 0;
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? s : (T - 1 - s);
            const int tfp = dir ? tt + 1 : tt - 1;          // time of c_{prev} in forward order
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                if (g >= nga) break;
                const int b0 = (grp0 + g) * NB, nbv = min(NB, B - b0);
                float* dhs = dhs0 + g * kDhs;
                // everything that does not depend on dh_rec is fetched and folded before the MMAs retire
                float cO[NB / 4], a1[NB / 4], cI[NB / 4], cF[NB / 4], cG[NB / 4], fgv[NB / 4], dyv[NB / 4];
#pragma unroll
                for (int k = 0; k < NB / 4; ++k) {
                    const int b = warp + 4 * k;
                    cO[k] = a1[k] = cI[k] = cF[k] = cG[k] = fgv[k] = dyv[k] = 0.f;
                    if (b < nbv) {
                        const int64_t bb = b0 + b;
                        const __nv_bfloat16* gp = p.gates + ((((int64_t)dir * T + tt) * B + bb) * 4) * H + u0 + lane;
                        const float ig = __bfloat162float(gp[0]), fg = __bfloat162float(gp[H]);
                        const float gg = __bfloat162float(gp[2 * H]), og = __bfloat162float(gp[3 * H]);
                        const float c = p.cstate[(((int64_t)dir * T + tt) * B + bb) * H + u0 + lane];
                        const float cp = (tfp >= 0 && tfp < T)
                                             ? p.cstate[(((int64_t)dir * T + tfp) * B + bb) * H + u0 + lane] : 0.f;
                        dyv[k] = p.dy[(bb * T + tt) * 2 * H + dir * H + u0 + lane];
                        const float tc_ = tanhf_fast(c);
                        cO[k] = tc_ * og * (1.0f - og);
                        a1[k] = og * (1.0f - tc_ * tc_);
                        cI[k] = gg * ig * (1.0f - ig);
                        cF[k] = cp * fg * (1.0f - fg);
                        cG[k] = ig * (1.0f - gg * gg);
                        fgv[k] = fg;
                    }
                }
                if (s > 0) {
                    mbar_wait(&mbar[g], ph_m);
                    tc_fence_after();
                    if (warp == 0) {                        // NB = 32 rows live in TMEM lanes 0..31
                        uint32_t v[32];
                        tc_ld_32x32b_x32(tmem_base + g * 32, v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) dhs[lane * 33 + j] = __uint_as_float(v[j]);
                    }
                    tc_fence_before();
                } else {
                    for (int i = threadIdx.x; i < kDhs; i += kEpiThreads) dhs[i] = 0.f;
                }
                named_bar_sync(1, kEpiThreads);
#pragma unroll
                for (int k = 0; k < NB / 4; ++k) {
                    const int b = warp + 4 * k;
                    if (b < nbv) {
                        const int64_t bb = b0 + b;
                        const float dh = dhs[b * 33 + lane] + dyv[k];
                        const float dc = fmaf(dh, a1[k], dcn[g][k]);
                        dcn[g][k] = dc * fgv[k];
                        __nv_bfloat16* dp = p.dgates + ((bb * T + tt) * 2 + dir) * 4 * H + u0 + lane;
                        dp[0] = __float2bfloat16(dc * cI[k]);
                        dp[H] = __float2bfloat16(dc * cF[k]);
                        dp[2 * H] = __float2bfloat16(dc * cG[k]);
                        dp[3 * H] = __float2bfloat16(dh * cO[k]);
                    }
                }
                fence_proxy_async();
                named_bar_sync(1, kEpiThreads);
                if (threadIdx.x == 0) red_release_gpu_add(p.counters + ((grp0 + g) * 2 + dir), 1u);
            }
            if (s > 0) ph_m ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, NG * 32 < 32 ? 32 : NG * 32);
}

// ------------------------------------------- backward, cluster + DSMEM, K-split / reduce-scatter ----
// dh_{t-1}[b, j] = sum_n dgates_t[b, n] W_hh[n, j].  Instead of all-gathering dgates (4H columns), each CTA
// contracts ONLY over the 128 gate columns it has just produced itself, for ALL H output units:
//     P_c[j, b] = sum_{n in slice c} W_hh[n, j] dgates_t[b, n]
// computed TRANSPOSED: A = W_hh[slice c, :]^T (resident, 128 KB, M = 128 output units per tile, H/128 tiles),
// B = this step's dgates tile [32 batch rows x 128] (8 KB, local), N = 32: 8 x H/128 small MMAs (16 cycles each)
// instead of 16 M128xN256 ones, and the accumulator has the output UNIT on the TMEM lane, so every epilogue
// warp drains its own lane quadrant (with the batch on the lanes only warp 0 could, 2.1 k cycles per step).
// The partial sums are reduce-scattered through distributed shared memory: tile (units of CTA d) x (32 rows)
// goes to CTA d as one 2 KB bulk copy, issued by the warp that drained it; CTA d adds the CS tiles it receives.
// Receivers release their receive buffer to the senders with remote mbarrier arrives (one per warp).
constexpr int NBR = 32;                  // batch rows per cluster
constexpr int kRsChunk = NBR / 8 * 128 + 16;   // B-tile K-chunk stride, +16 B: the 4 lane groups of a store hit different banks

template <int EW>                        // epilogue warps: 4 or 8 (warps w and w+4 share a TMEM lane quadrant)
__global__ void __launch_bounds__((EW + 2) * 32, 1)
lstm_bwd_rs_kernel(const __grid_constant__ CUtensorMap map_wt, BwdDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    constexpr int RPT = NBR / EW;                            // batch rows per thread (thread = unit lane, rows RPT*w ..)
    constexpr int kChunk = kRsChunk;
    constexpr int kBTile = 16 * kChunk;                      // [128 k' / 8][4 row groups][8][16 B] (+ padding)
    constexpr int kTile = NBR * 32 * 2;                      // 2 KB: one [32 units x 32 rows] bf16 partial tile
    const int H = p.H, T = p.T, B = p.B;
    const int NM = H / 128;                                  // M tiles of 128 output units
    const int TPW = NM / (EW / 4);                           // tiles drained per warp
    const int CS = gridDim.x;
    uint8_t* Wb = smem;                                      // 2 k-blocks x [H rows x 64] bf16, SWIZZLE_128B
    uint8_t* Bt = Wb + 2 * H * 128;                          // dgates tile of the step (B operand)
    uint8_t* stg = Bt + ((kBTile + 1023) & ~1023);           // [CS] outgoing partial tiles
    uint8_t* rcv = stg + CS * kTile;                         // [CS] incoming partial tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(rcv + CS * kTile);
    uint64_t* wbar = bars + 0;
    uint64_t* aready = bars + 1;         // B tile of the step written
    uint64_t* dready = bars + 2;         // MMAs of the step retired
    uint64_t* dfree = bars + 3;          // accumulator drained (one arrival per epilogue warp)
    uint64_t* rfull = bars + 4;          // all partial tiles of the step landed
    uint64_t* rfree = bars + 5;          // every receiver has consumed my tiles (one arrival per CTA)
    uint64_t* abar = bars + 6;           // W_hh^T slice copied into tensor memory (one arrival per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // shfl: warp-uniform for ptxas
    const int cta = (int)cluster_ctarank();
    const int dir = blockIdx.y, grp = blockIdx.z;
    const int u0 = cta * 32, b0 = grp * NBR;
    const int nbv = min(NBR, B - b0);
    const uint32_t step_bytes = (uint32_t)CS * kTile;
    // accumulators: NM x 32 columns; A operand (W_hh^T slice, bf16 pairs): NM x 64 columns at column 128
    const uint32_t tmem_cols = NM > 2 ? 512u : 256u;

    if (threadIdx.x == 0) {
        mbar_init(wbar, 1); mbar_init(aready, 1); mbar_init(dready, 1); mbar_init(dfree, EW);
        mbar_init(rfull, 1); mbar_init(rfree, (uint32_t)CS); mbar_init(abar, EW);
        fence_barrier_init();
        if (T >= 2) mbar_expect_tx(rfull, step_bytes);
    }
    if (warp == EW + 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // uniform register: no R2UR waterfall per tcgen05 op
    const uint32_t tmem_a = tmem_base + 128;

    if (warp == EW) {
        if (lane == 0) {
            // W_hh[slice c, :]^T as A operand: rows j (all H), K = my 128 permuted gate columns
            mbar_expect_tx(wbar, (uint32_t)(2 * H * 128));
            for (int kb = 0; kb < 2; ++kb)
                for (int h = 0; h < H / 256; ++h)
                    tma_load_2d(&map_wt, wbar, Wb + kb * H * 128 + h * 256 * 128, cta * 128 + kb * 64, dir * H + h * 256);
        }
    } else if (warp == EW + 1) {
        // whole warp in the loop, one elected lane issues (see lstm_fwd_cluster_kernel)
        const uint32_t idesc = make_idesc(128, NBR);
        mbar_wait(abar, 0);                                      // W_hh^T slice is in tensor memory
        tc_fence_after();
        const uint64_t bd0 = make_nosw_desc(smem_u32(Bt), kChunk, 128);
        for (int s = 0; s + 1 < T; ++s) {
            mbar_wait(aready, (uint32_t)(s & 1));
            if (s > 0) mbar_wait(dfree, (uint32_t)((s - 1) & 1));
            tc_fence_after();
            if (elect_one()) {
                for (int m = 0; m < NM; ++m) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        tc_mma_bf16_ts(tmem_base + m * 32, tmem_a + m * 64 + kk * 8,
                                       bd0 + (uint64_t)(kk * (2 * kChunk / 16)), idesc, (uint32_t)(kk != 0));
                }
                tc_commit(dready);
            }
            __syncwarp();
        }
    } else {
        // Epilogue mapping: lane = unit (u0 + lane), warp w owns batch rows RPT*w .. RPT*w + RPT - 1.
        const int q = warp & 3;                                   // TMEM lane quadrant of this warp
        const int t0 = (warp >> 2) * TPW;                         // first M tile this warp drains
        {
            // one-time: rows 128m + 32q + lane of the W_hh^T slice -> tensor memory lane, two bf16 per 32-bit
            // column, read back out of the 128B-swizzled tiles the TMA wrote (A from TMEM: no per-MMA smem read of A)
            mbar_wait(wbar, 0);
            const int r = q * 32 + lane;
            for (int m = t0; m < t0 + TPW; ++m)
                for (int kb = 0; kb < 2; ++kb) {
                    uint32_t w[32];
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint4 x = *reinterpret_cast<const uint4*>(Wb + kb * H * 128 + (m * 128 + r) * 128 + ((ch ^ (r & 7)) << 4));
                        w[ch * 4 + 0] = x.x; w[ch * 4 + 1] = x.y; w[ch * 4 + 2] = x.z; w[ch * 4 + 3] = x.w;
                    }
                    tc_st_32x32b_x32(tmem_a + m * 64 + kb * 32 + ((uint32_t)(q * 32) << 16), w);
                }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(abar);
        }
        float dcn[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) dcn[i] = 0.f;
        long long* prof = (p.prof && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) ? p.prof : nullptr;
#define PK2_PROF(e) do { if (prof && s >= 64 && s < 72) prof[(s - 64) * 16 + (e)] = clock64(); } while (0)
        // RAW operands of one step, fetched a whole step ahead.  Nothing touches these registers until the next
        // step (no conversion, no select): a dependent instruction right after the load would stall the warp for
        // the full HBM latency (measured: 7 k cycles per step, profiles/lstm_bwd_rs_trace_r1_v10_dbg.txt).
        unsigned short q_g[RPT][4];
        float q_c[RPT], q_cp[RPT], q_dy[RPT];
        auto fetch = [&](int s) {
            const int tt = dir ? s : (T - 1 - s);
            const int tfp = dir ? tt + 1 : tt - 1;
            const int tcp = (tfp >= 0 && tfp < T) ? tfp : tt;          // always a valid address; masked at use
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int b = min(RPT * warp + i, nbv - 1);
                const int64_t bb = b0 + b;
                const unsigned short* gp = reinterpret_cast<const unsigned short*>(p.gates) +
                                           ((((int64_t)dir * T + tt) * B + bb) * 4) * H + u0 + lane;
#pragma unroll
                for (int g = 0; g < 4; ++g) q_g[i][g] = gp[g * H];
                q_c[i] = p.cstate[(((int64_t)dir * T + tt) * B + bb) * H + u0 + lane];
                q_cp[i] = p.cstate[(((int64_t)dir * T + tcp) * B + bb) * H + u0 + lane];
                q_dy[i] = p.dy[(bb * T + tt) * 2 * H + dir * H + u0 + lane];
            }
        };
        fetch(0);
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);        // 16-byte chunk swizzle of the partial tiles
        const uint32_t hsw = (uint32_t)((lane >> 3) & 1);        // 8-byte half swap inside a chunk
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? s : (T - 1 - s);
            const int tfp = dir ? tt + 1 : tt - 1;
            const float cpm = (tfp >= 0 && tfp < T) ? 1.f : 0.f;
            float cO[RPT], a1[RPT], cI[RPT], cF[RPT], cG[RPT], fgv[RPT], dh[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float live = (RPT * warp + i < nbv) ? 1.f : 0.f;
                const float ig = __uint_as_float((uint32_t)q_g[i][0] << 16);
                const float fg = __uint_as_float((uint32_t)q_g[i][1] << 16);
                const float gg = __uint_as_float((uint32_t)q_g[i][2] << 16);
                const float og = __uint_as_float((uint32_t)q_g[i][3] << 16);
                const float cp = q_cp[i] * cpm;
                const float tc_ = tanh_approx(q_c[i]);
                cO[i] = live * tc_ * og * (1.0f - og);
                a1[i] = live * og * (1.0f - tc_ * tc_);
                cI[i] = gg * ig * (1.0f - ig);
                cF[i] = cp * fg * (1.0f - fg);
                cG[i] = ig * (1.0f - gg * gg);
                fgv[i] = fg;
                dh[i] = live * q_dy[i];
            }
            if (s + 1 < T) fetch(s + 1);                 // in flight during this whole step
            if (s > 0) {
                // reduce: dh_rec[b, u] = sum over the source CTAs of their partial tile
                mbar_wait(rfull, (uint32_t)((s - 1) & 1));
                PK2_PROF(0);
                if (threadIdx.x == 0 && s + 1 < T) mbar_expect_tx(rfull, step_bytes);   // re-arm for this step's tiles
                float acc[RPT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) acc[i] = 0.f;
                if (EW == 4) {
                    // rows 8w..8w+7 = chunk w of my unit's 64-byte row
                    const uint8_t* rp = rcv + lane * 64 + (((uint32_t)warp ^ swz) << 4);
#pragma unroll 4
                    for (int src = 0; src < CS; ++src) {
                        uint4 v = *reinterpret_cast<const uint4*>(rp + src * kTile);
                        if (hsw) { uint32_t t; t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
                        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j2 = 0; j2 < 4; ++j2) {
                            acc[(2 * j2) % RPT] += __uint_as_float(wv[j2] << 16);
                            acc[(2 * j2 + 1) % RPT] += __uint_as_float(wv[j2] & 0xffff0000u);
                        }
                    }
                } else {
                    // rows 4w..4w+3 = half (w & 1) of chunk w >> 1
                    const uint8_t* rp = rcv + lane * 64 + ((((uint32_t)warp >> 1) ^ swz) << 4) + ((((uint32_t)warp & 1) ^ hsw) << 3);
#pragma unroll 4
                    for (int src = 0; src < CS; ++src) {
                        const uint2 v = *reinterpret_cast<const uint2*>(rp + src * kTile);
                        acc[0] += __uint_as_float(v.x << 16);
                        acc[1 % RPT] += __uint_as_float(v.x & 0xffff0000u);
                        acc[2 % RPT] += __uint_as_float(v.y << 16);
                        acc[3 % RPT] += __uint_as_float(v.y & 0xffff0000u);
                    }
                }
#pragma unroll
                for (int i = 0; i < RPT; ++i) dh[i] += acc[i];
                PK2_PROF(6);
            }
            unsigned short dg[RPT][4];
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int b = RPT * warp + i;
                const float dc = fmaf(dh[i], a1[i], dcn[i]);
                dcn[i] = dc * fgv[i];
                const float o[4] = {dc * cI[i], dc * cF[i], dc * cG[i], dh[i] * cO[i]};
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    dg[i][g] = __bfloat16_as_ushort(__float2bfloat16_rn(o[g]));
                    if (s + 1 < T) {
                        const int col = g * 32 + lane;                // k' within my slice = gate*32 + unit
                        *reinterpret_cast<unsigned short*>(Bt + (col >> 3) * kChunk + (b >> 3) * 128 + (b & 7) * 16 + (col & 7) * 2) = dg[i][g];
                    }
                }
            }
            if (s + 1 < T) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_bar_sync(1, EW * 32);
                if (threadIdx.x == 0) { PK2_PROF(2); mbar_arrive(aready); }
                // every epilogue warp is past its reads of rcv: release it to the senders, one arrival per
                // destination, spread over the warps (16 remote arrives from ONE warp serialise: 1.2 k cycles)
                constexpr int kStride = EW * 32 / 16;
                if (s > 0 && (threadIdx.x % kStride) == 0 && (int)(threadIdx.x / kStride) < CS)
                    mbar_arrive_remote_relaxed(mapa_u32(smem_u32(rfree), threadIdx.x / kStride));   // reads consumed before the barrier above
                PK2_PROF(1);
            }
            // off the critical path: dgates in the natural layout for the weight-gradient GEMMs
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int b = RPT * warp + i;
                if (b < nbv) {
                    unsigned short* dp = reinterpret_cast<unsigned short*>(p.dgates) +
                                         (((int64_t)(b0 + b) * T + tt) * 2 + dir) * 4 * H + u0 + lane;
#pragma unroll
                    for (int g = 0; g < 4; ++g) dp[g * H] = dg[i][g];
                }
            }
            if (s + 1 < T) {
                // partial sums of this step: TMEM lane = output unit, column = batch row
                mbar_wait(dready, (uint32_t)(s & 1));
                PK2_PROF(3);
                tc_fence_after();
                if (s > 0) mbar_wait_cluster(rfree, (uint32_t)((s - 1) & 1));   // receivers consumed my previous tiles
                for (int t = t0; t < t0 + TPW; ++t) {
                    uint32_t v[32];
                    tc_ld_32x32b_x32(tmem_base + t * 32 + ((uint32_t)(q * 32) << 16), v);
                    uint8_t* dstp = stg + (4 * t + q) * kTile + lane * 64;   // tile for CTA 4t+q: [unit = lane][32 rows]
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]));
                        __nv_bfloat162 p1 = __floats2bfloat162_rn(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]));
                        __nv_bfloat162 p2 = __floats2bfloat162_rn(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]));
                        __nv_bfloat162 p3 = __floats2bfloat162_rn(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]));
                        uint4 o;
                        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                        if (hsw) { uint32_t x; x = o.x; o.x = o.z; o.z = x; x = o.y; o.y = o.w; o.w = x; }
                        *reinterpret_cast<uint4*>(dstp + (((uint32_t)c ^ swz) << 4)) = o;
                    }
                }
                tc_fence_before();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                PK2_PROF(4);
                if (lane == 0) mbar_arrive(dfree);
                if (lane < TPW) {
                    // tile for CTA `rank` -> slot `cta` of its receive buffer; bytes counted on its rfull
                    const uint32_t rank = (uint32_t)(4 * (t0 + lane) + q);
                    dsmem_bulk_copy(mapa_u32(smem_u32(rcv + cta * kTile), rank), smem_u32(stg + rank * kTile),
                                    (uint32_t)kTile, mapa_u32(smem_u32(rfull), rank));
                }
                PK2_PROF(5);
            }
        }
#undef PK2_PROF
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == EW + 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------ host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// bf16 tensor map, rank 2 or 3, innermost box = 64 elements (128 B), SWIZZLE_128B
int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
    EncodeTiledFn enc = get_encode();
    PK2_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                     strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PK2_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

int check_dims(const char* who, int B, int T, int H, int ng, int num_sms) {
    PK2_REQUIRE(H % 64 == 0 && H >= 64 && H <= 512, "%s: hidden size %d unsupported (multiple of 64, <= 512)", who, H);
    PK2_REQUIRE(B > 0 && T > 0, "%s: empty batch", who);
    const int G = (B + NB * ng - 1) / (NB * ng);
    PK2_REQUIRE((H / 32) * 2 * G <= num_sms, "%s: B=%d needs %d co-resident CTAs (> %d SMs); split the batch", who, B,
                (H / 32) * 2 * G, num_sms);
    return 0;
}

long long* g_prof = nullptr;    // set through pk2_lstm_set_profile_buffer (profiling only)
long long* g_prof_bwd = nullptr;

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int NG>
int launch_fwd(const pk2_lstm_fwd_args* a, cudaStream_t st) {
    const int H = a->H, T = a->T, B = a->B, KB = H / 64;
    const int groups = (B + NB - 1) / NB, G = (groups + NG - 1) / NG;
    CUtensorMap mw, my;
    {   // packed W_hh rows: [(dir*H/32 + cta)*128 + gate*32 + ul][H]
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)(2 * 4 * H)};
        cuuint64_t str[1] = {(cuuint64_t)H * 2};
        cuuint32_t box[2] = {64, 128};
        if (make_map(&mw, a->whh, 2, dims, str, box)) return 2;
    }
    {   // y[B][T][2H] viewed as (64 cols, b, t, 64-col block): one box = KB swizzled [NB x 64] tiles
        cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)T, (cuuint64_t)(2 * H / 64)};
        cuuint64_t str[3] = {(cuuint64_t)T * 2 * H * 2, (cuuint64_t)(2 * H) * 2, 128};
        cuuint32_t box[4] = {64, (cuuint32_t)NB, 1, (cuuint32_t)KB};
        if (make_map(&my, a->y, 4, dims, str, box)) return 2;
    }
    const size_t smem = (size_t)KB * 16384 + (size_t)NG * KB * NB * 128 + (size_t)NG * NB * 512 + 128 + 1024;
    PK2_CHECK(cudaFuncSetAttribute(lstm_fwd_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PK2_CHECK(cudaMemsetAsync(a->sync, 0, sizeof(unsigned) * 2 * groups, st));
    FwdDev d;
    d.B = B; d.T = T; d.H = H; d.gx = a->gx;
    d.y = static_cast<__nv_bfloat16*>(a->y);
    d.gates = static_cast<__nv_bfloat16*>(a->gates);
    d.cstate = a->cstate; d.counters = a->sync; d.prof = nullptr; d.a_tmem = 0;
    lstm_fwd_kernel<NG><<<dim3(H / 32, 2, G), kThreads, smem, st>>>(mw, my, d);
    PK2_POST_LAUNCH();
    return 0;
}

// Cluster/DSMEM forward: returns 0 on success, -1 if this device cannot co-schedule the clusters.
template <int EW, int ACC>
int launch_fwd_cluster_t(const pk2_lstm_fwd_args* a, cudaStream_t st) {
    const int H = a->H, T = a->T, B = a->B, KB = H / 64, CS = H / 32;
    const int G = (B + NB - 1) / NB;
    CUtensorMap mw;
    {
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)(2 * 4 * H)};
        cuuint64_t str[1] = {(cuuint64_t)H * 2};
        cuuint32_t box[2] = {64, 128};
        if (make_map(&mw, a->whh, 2, dims, str, box)) return 2;
    }
    const size_t smem = (size_t)KB * 16384 + 2 * (size_t)H * NB * 2 + (size_t)NB * 512 + 2 * NB * 64 + 128 + 1024;
    static bool attr_done = false, usable = true;
    if (!attr_done) {
        attr_done = true;
        if (cudaFuncSetAttribute(lstm_fwd_cluster_kernel<EW, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess ||
            cudaFuncSetAttribute(lstm_fwd_cluster_kernel<EW, ACC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            usable = false;
        }
    }
    if (!usable) return -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, 2, G);
    cfg.blockDim = dim3((EW + 2) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, lstm_fwd_cluster_kernel<EW, ACC>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (max_clusters < 1) return -1;             // clusters are independent: fewer resident ones just run in waves
    FwdDev d;
    d.B = B; d.T = T; d.H = H; d.gx = a->gx;
    d.y = static_cast<__nv_bfloat16*>(a->y);
    d.gates = static_cast<__nv_bfloat16*>(a->gates);
    d.cstate = a->cstate; d.counters = a->sync; d.prof = g_prof;
    static const bool a_smem = getenv("PK2_LSTM_A_SMEM") != nullptr;       // debug: A operand from shared memory
    d.a_tmem = (!a_smem && H <= 512 && (H / 64) % (EW / 4) == 0) ? 1 : 0;
    PK2_CHECK(cudaLaunchKernelEx(&cfg, lstm_fwd_cluster_kernel<EW, ACC>, mw, d));
    PK2_LAUNCHED();
    return 0;
}

int launch_fwd_cluster(const pk2_lstm_fwd_args* a, cudaStream_t st) {
    const int CS = a->H / 32;
    if (CS > 16 || (CS & (CS - 1)) != 0) return -1;
    // experiment knobs: PK2_LSTM_FWD_EW = 4 | 8 epilogue warps, PK2_LSTM_FWD_ACC = 1 | 4 accumulators
    static int ew = -1, acc = -1;
    if (ew < 0) {
        const char* e = getenv("PK2_LSTM_FWD_EW"); ew = (e && atoi(e) == 4) ? 4 : 8;
        const char* c = getenv("PK2_LSTM_FWD_ACC"); acc = (c && atoi(c) == 4) ? 4 : 1;
    }
    if (ew == 4) return acc == 4 ? launch_fwd_cluster_t<4, 4>(a, st) : launch_fwd_cluster_t<4, 1>(a, st);
    return acc == 4 ? launch_fwd_cluster_t<8, 4>(a, st) : launch_fwd_cluster_t<8, 1>(a, st);
}

template <int NG>
int launch_bwd(const pk2_lstm_bwd_args* a, cudaStream_t st) {
    const int H = a->H, T = a->T, B = a->B, KB = 4 * H / 64;
    const int groups = (B + NB - 1) / NB, G = (groups + NG - 1) / NG;
    CUtensorMap mwt, mdg;
    {   // W_hh^T: [dir*H + j][4H]
        cuuint64_t dims[2] = {(cuuint64_t)(4 * H), (cuuint64_t)(2 * H)};
        cuuint64_t str[1] = {(cuuint64_t)(4 * H) * 2};
        cuuint32_t box[2] = {64, 32};
        if (make_map(&mwt, a->whh_t, 2, dims, str, box)) return 2;
    }
    {   // dgates[B][T][2*4H] viewed as (64 cols, b, t, 64-col block): one box = CH swizzled [NB x 64] tiles
        cuuint64_t dims[4] = {64, (cuuint64_t)B, (cuuint64_t)T, (cuuint64_t)(8 * H / 64)};
        cuuint64_t str[3] = {(cuuint64_t)T * 8 * H * 2, (cuuint64_t)(8 * H) * 2, 128};
        const int ch = (kChunkBytes / (NB * 128)) < KB ? (kChunkBytes / (NB * 128)) : KB;
        cuuint32_t box[4] = {64, (cuuint32_t)NB, 1, (cuuint32_t)ch};
        PK2_REQUIRE(KB % ch == 0, "pk2_lstm_layer_bwd: hidden size %d not supported", H);
        if (make_map(&mdg, a->dgates, 4, dims, str, box)) return 2;
    }
    const size_t smem = (size_t)KB * 4096 + (size_t)kRing * kChunkBytes + kASlack + (size_t)((NG * NB * 33 * 4 + 7) & ~7) + 128 + 1024;
    PK2_CHECK(cudaFuncSetAttribute(lstm_bwd_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PK2_CHECK(cudaMemsetAsync(a->sync, 0, sizeof(unsigned) * 2 * groups, st));
    BwdDev d;
    d.B = B; d.T = T; d.H = H; d.dy = a->dy;
    d.gates = static_cast<const __nv_bfloat16*>(a->gates);
    d.cstate = a->cstate;
    d.dgates = static_cast<__nv_bfloat16*>(a->dgates);
    d.counters = a->sync; d.prof = nullptr; d.dbg = 0;
    lstm_bwd_kernel<NG><<<dim3(H / 32, 2, G), kThreads, smem, st>>>(mwt, mdg, d);
    PK2_POST_LAUNCH();
    return 0;
}

// K-split / reduce-scatter backward on clusters; needs the permuted W_hh^T (a->whh_t_perm).
template <int EW>
int launch_bwd_rs_t(const pk2_lstm_bwd_args* a, cudaStream_t st) {
    const int H = a->H, T = a->T, B = a->B, CS = H / 32;
    const int G = (B + NBR - 1) / NBR;
    CUtensorMap mwt;
    {
        cuuint64_t dims[2] = {(cuuint64_t)(4 * H), (cuuint64_t)(2 * H)};
        cuuint64_t str[1] = {(cuuint64_t)(4 * H) * 2};
        cuuint32_t box[2] = {64, 256};
        if (make_map(&mwt, a->whh_t_perm, 2, dims, str, box)) return 2;
    }
    const size_t smem = (size_t)2 * H * 128 + ((16 * kRsChunk + 1023) & ~1023) + 2 * (size_t)CS * 2048 + 128 + 1024;
    static bool attr_done = false, usable = true;
    if (!attr_done) {
        attr_done = true;
        if (cudaFuncSetAttribute(lstm_bwd_rs_kernel<EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448) != cudaSuccess ||
            cudaFuncSetAttribute(lstm_bwd_rs_kernel<EW>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            usable = false;
        }
    }
    if (!usable || smem > 232448) return -1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS, 2, G);
    cfg.blockDim = dim3((EW + 2) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, lstm_bwd_rs_kernel<EW>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    if (max_clusters < 1) return -1;
    BwdDev d;
    d.B = B; d.T = T; d.H = H; d.dy = a->dy;
    d.gates = static_cast<const __nv_bfloat16*>(a->gates);
    d.cstate = a->cstate;
    d.dgates = static_cast<__nv_bfloat16*>(a->dgates);
    d.counters = a->sync; d.prof = g_prof_bwd;
    d.dbg = 0;
    PK2_CHECK(cudaLaunchKernelEx(&cfg, lstm_bwd_rs_kernel<EW>, mwt, d));
    PK2_LAUNCHED();
    return 0;
}

int launch_bwd_rs(const pk2_lstm_bwd_args* a, cudaStream_t st) {
    const int H = a->H, CS = H / 32;
    // H = 256 (8 CTAs, 2 M tiles) or 512 (16 CTAs, 4 M tiles); the resident W_hh^T slice is 256*H bytes
    if ((H != 256 && H != 512) || CS > 16 || a->whh_t_perm == nullptr) return -1;
    static int ew = -1;
    if (ew < 0) { const char* e = getenv("PK2_LSTM_RS_EW"); ew = (e && atoi(e) == 4) ? 4 : 8; }
    return ew == 4 ? launch_bwd_rs_t<4>(a, st) : launch_bwd_rs_t<8>(a, st);
}

}  // namespace

// Profiling aid: device buffer of 8*16 int64 that receives clock64() stamps of steps 64..71 of CTA (0,0,0)
// of the cluster forward kernel (see tools/lstm_trace.py).  NULL switches it off.
extern "C" int pk2_lstm_set_profile_buffer(void* buf) {
    g_prof = static_cast<long long*>(buf);
    g_prof_bwd = buf ? static_cast<long long*>(buf) + 128 : nullptr;      // second half: backward trace
    return 0;
}

extern "C" int pk2_lstm_layer_fwd(const pk2_lstm_fwd_args* a, void* stream) {
    PK2_REQUIRE(a && a->gx && a->whh && a->y && a->gates && a->cstate && a->sync, "pk2_lstm_layer_fwd: null argument");
    // Measured (profiles/kernel_bench_r1_v2.jsonl): interleaving two groups per CTA does not hide the
    // hand-off latency (period = latency + work either way) and serialises the groups; one group per
    // CTA, groups side by side on different SMs, is faster.  NG = 2 is kept for batches that would
    // otherwise not fit the SMs (B > 128).
    const int ng = a->B > 4 * NB ? 2 : 1;
    if (check_dims("pk2_lstm_layer_fwd", a->B, a->T, a->H, ng, num_sms())) return 2;
    static const bool no_cluster = getenv("PK2_LSTM_NO_CLUSTER") != nullptr;
    if (!no_cluster) {
        const int rc = launch_fwd_cluster(a, pk2::as_stream(stream));
        if (rc >= 0) return rc;                   // -1: clusters of H/32 CTAs not schedulable -> global-memory exchange
    }
    return ng == 1 ? launch_fwd<1>(a, pk2::as_stream(stream)) : launch_fwd<2>(a, pk2::as_stream(stream));
}

extern "C" int pk2_lstm_layer_bwd(const pk2_lstm_bwd_args* a, void* stream) {
    PK2_REQUIRE(a && a->dy && a->whh_t && a->gates && a->cstate && a->dgates && a->sync, "pk2_lstm_layer_bwd: null argument");
    const int ng = a->B > 4 * NB ? 2 : 1;
    if (check_dims("pk2_lstm_layer_bwd", a->B, a->T, a->H, ng, num_sms())) return 2;
    static const bool no_cluster = getenv("PK2_LSTM_NO_CLUSTER") != nullptr;
    // Default: K-split / reduce-scatter kernel on clusters (9.5 k cycles per step, profiles/lstm_bwd_rs_trace_r1_v11.txt);
    // PK2_LSTM_NO_RS=1 or an unschedulable cluster falls back to the global-memory kernel (8.9 us per step).
    static const bool use_rs = getenv("PK2_LSTM_NO_RS") == nullptr;
    if (!no_cluster && use_rs) {
        const int rc = launch_bwd_rs(a, pk2::as_stream(stream));
        if (rc >= 0) return rc;                   // -1: not applicable (H % 256, clusters) -> global-memory kernel
    }
    return ng == 1 ? launch_bwd<1>(a, pk2::as_stream(stream)) : launch_bwd<2>(a, pk2::as_stream(stream));
}
