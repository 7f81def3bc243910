// LF-MMI numerator forward-backward on the GPU (sm_100a).
//
// Replaces the numerator half of kaldi_chain.compute_chain_objf_and_deriv (reference
// ops/ops.py:265; Kaldi chain-numerator.cc runs it on the CPU in the log domain).
// Supervision FSTs are epsilon-free with time-stamped states sorted by time, so the
// recursion is level-synchronous: one warp per sequence walks the time levels, one lane
// per state of the level, log-add in double as Kaldi does.  The work is tiny next to the
// denominator (<= ~10 states per level); it runs concurrently with it.
#include "common.cuh"

namespace {

// 8 warps (sequences) per CTA: the kernel runs next to the denominator kernels, whose CTAs own a whole SM
// each and are placed as clusters of 8 inside a GPC.  64 one-warp CTAs would be spread over 64 SMs and keep
// most clusters from being placed until they finish; 8 CTAs take one SM in each GPC at most.
constexpr int kNumWarps = 8;

__global__ void __launch_bounds__(32 * kNumWarps)
numfb_kernel(pk2_sup_batch sup, const float* __restrict__ loglikes, int N, int64_t row_stride_b,
             float deriv_scale, double* alpha, double* beta,
             float* __restrict__ grad, double* __restrict__ logz, float* __restrict__ arc_post) {
    const int b = blockIdx.x * kNumWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= sup.n_seq) return;
    const int T = sup.num_frames[b];
    const int32_t* lvl = sup.level_off + sup.lvl_base[b];   // lvl[t]..lvl[t+1] = states of time t
    const float* ll = loglikes + (int64_t)b * row_stride_b * N;
    float* g = grad + (int64_t)b * row_stride_b * N;
    const int s_begin = sup.seq_state_off[b];

    // level 0: the start state is the first state of the sequence
    for (int s = lvl[0] + lane; s < lvl[1]; s += 32) alpha[s] = (s == s_begin) ? 0.0 : -INFINITY;
    __syncwarp();
    for (int t = 1; t <= T; ++t) {
        const float* row = ll + (int64_t)(t - 1) * N;
        for (int s = lvl[t] + lane; s < lvl[t + 1]; s += 32) {
            double acc = -INFINITY;
            for (int k = sup.in_off[s]; k < sup.in_off[s + 1]; ++k) {
                const double sc = alpha[sup.in_src[k]] + (double)row[sup.in_pdf[k]] - (double)sup.in_w[k];
                acc = pk2::log_add(acc, sc);
            }
            alpha[s] = acc;
        }
        __syncwarp();
    }
    // total over final states (time T)
    double z = -INFINITY;
    for (int s = lvl[T] + lane; s < lvl[T + 1]; s += 32) {
        const float fc = sup.final_cost[s];
        const double bt = (fc < INFINITY) ? -(double)fc : -INFINITY;
        beta[s] = bt;
        z = pk2::log_add(z, alpha[s] + bt);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z = pk2::log_add(z, __shfl_xor_sync(0xffffffffu, z, o));
    if (lane == 0) logz[b] = z;
    __syncwarp();
    for (int t = T - 1; t >= 0; --t) {
        const float* row = ll + (int64_t)t * N;
        for (int s = lvl[t] + lane; s < lvl[t + 1]; s += 32) {
            double acc = -INFINITY;
            const double a = alpha[s];
            for (int k = sup.out_off[s]; k < sup.out_off[s + 1]; ++k) {
                const int p = sup.out_pdf[k];
                const double sc = (double)row[p] - (double)sup.out_w[k] + beta[sup.out_dst[k]];
                acc = pk2::log_add(acc, sc);
                const double post = exp(a + sc - z);
                if (arc_post) arc_post[k] = (float)post;          // split mode: scatter later
                else if (post > 0.0) atomicAdd(&g[(int64_t)t * N + p], deriv_scale * (float)post);
            }
            beta[s] = acc;
        }
        __syncwarp();
    }
}

// grad[b, time(s), pdf] += scale * arc_post[k] for every out-arc k of every state s (split mode)
__global__ void num_scatter_kernel(pk2_sup_batch sup, int total_states, const float* __restrict__ arc_post, int N,
                                   int64_t row_stride_b, float scale, float* __restrict__ grad) {
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < total_states; s += gridDim.x * blockDim.x) {
        int lo = 0, hi = sup.n_seq;                 // sequence of this state
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (sup.seq_state_off[mid] <= s) lo = mid; else hi = mid;
        }
        float* g = grad + ((int64_t)lo * row_stride_b + sup.state_time[s]) * N;
        for (int k = sup.out_off[s]; k < sup.out_off[s + 1]; ++k) {
            const float v = arc_post[k];
            if (v > 0.f) atomicAdd(&g[sup.out_pdf[k]], scale * v);
        }
    }
}

// Kaldi's fallback for a sequence whose objective is not finite (chain-training.cc: derivatives <- 0): zero its
// gradient rows on the device, without a host round trip.  Blocks of healthy sequences return at once.
__global__ void chain_guard_kernel(const double* __restrict__ logz_den, const double* __restrict__ logz_num,
                                   int64_t row_elems, float* __restrict__ grad) {
    const int b = blockIdx.y;
    if (isfinite(logz_den[b]) && isfinite(logz_num[b])) return;
    float* g = grad + (int64_t)b * row_elems;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < row_elems; i += (int64_t)gridDim.x * blockDim.x) g[i] = 0.f;
}

}  // namespace

extern "C" int pk2_chain_guard(const double* logz_den, const double* logz_num, int n_seq, int64_t row_elems,
                               float* grad, void* stream) {
    PK2_REQUIRE(logz_den && logz_num && grad, "pk2_chain_guard: null argument");
    if (n_seq <= 0 || row_elems <= 0) return 0;
    chain_guard_kernel<<<dim3(32, n_seq), 256, 0, pk2::as_stream(stream)>>>(logz_den, logz_num, row_elems, grad);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_numfb_post(const pk2_sup_batch* sup, const float* loglikes, int num_pdfs, int64_t row_stride_b,
                              double* ws_alpha, double* ws_beta, float* arc_post, double* logz, void* stream) {
    PK2_REQUIRE(sup && loglikes && ws_alpha && ws_beta && arc_post && logz, "pk2_numfb_post: null argument");
    PK2_REQUIRE(sup->n_seq > 0, "pk2_numfb_post: empty batch");
    numfb_kernel<<<(sup->n_seq + kNumWarps - 1) / kNumWarps, 32 * kNumWarps, 0, pk2::as_stream(stream)>>>(*sup, loglikes, num_pdfs, row_stride_b, 0.f, ws_alpha,
                                                              ws_beta, nullptr, logz, arc_post);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_numfb_scatter(const pk2_sup_batch* sup, int total_states, const float* arc_post, int num_pdfs,
                                 int64_t row_stride_b, float deriv_scale, float* grad, void* stream) {
    PK2_REQUIRE(sup && arc_post && grad, "pk2_numfb_scatter: null argument");
    if (total_states <= 0) return 0;
    int blocks = (total_states + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    num_scatter_kernel<<<blocks, 256, 0, pk2::as_stream(stream)>>>(*sup, total_states, arc_post, num_pdfs, row_stride_b,
                                                                 deriv_scale, grad);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_numfb(const pk2_sup_batch* sup, const float* loglikes, int num_pdfs,
                         int64_t row_stride_b, float deriv_scale, double* ws_alpha, double* ws_beta,
                         float* grad, double* logz, void* stream) {
    PK2_REQUIRE(sup && loglikes && ws_alpha && ws_beta && grad && logz, "pk2_numfb: null argument");
    PK2_REQUIRE(sup->n_seq > 0, "pk2_numfb: empty batch");
    numfb_kernel<<<(sup->n_seq + kNumWarps - 1) / kNumWarps, 32 * kNumWarps, 0, pk2::as_stream(stream)>>>(*sup, loglikes, num_pdfs, row_stride_b,
                                                              deriv_scale, ws_alpha, ws_beta, grad, logz, nullptr);
    PK2_POST_LAUNCH();
    return 0;
}
