// Lattice forward-backward + MMI posterior merge on the GPU (sm_100a).
//
// Replaces lattice_forward_backward_mmi(..., drop_frames=True, convert_to_pdf_ids=False,
// cancel=True) + Posterior.to_pdf_matrix (reference ops/ops.py:57-62; Kaldi
// lat/lattice-functions.cc + hmm/posterior.cc on the CPU) and removes the per-utterance
// D2H/H2D copies of ops/ops.py:55,64.  Recursion: SURVEY.md Appendix B, alpha/beta in double.
// Lattice states are time-stamped, so the recursion is level-synchronous: one CTA per
// utterance walks the time levels, one thread per state of the level; same-level epsilon
// arcs (rare) are applied serially in topological order between levels.  The arc's acoustic
// score is gathered from the loglike row already on the device.  Output: dense gradient
// rows  grad[t,:] = den_post(t,:) - num_post(t,:)  for kept frames, 0 for dropped frames.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
latfb_kernel(pk2_lat_batch lat, const float* __restrict__ loglikes, int N, int64_t row_stride_b,
             float lm, float ac, double* alpha, double* beta, float* __restrict__ grad,
             double* __restrict__ tot_out) {
    __shared__ double s_tot;
    __shared__ double s_red[kThreads / 32];
    const int b = blockIdx.x;
    const int T = lat.num_frames[b];
    const int32_t* lvl = lat.level_off + lat.lvl_base[b];
    const int32_t* eoff = lat.eps_off + lat.lvl_base[b];   // eps arcs of level t: [eoff[t], eoff[t+1])
    const float* ll = loglikes + (int64_t)b * row_stride_b * N;
    float* g = grad + (int64_t)b * row_stride_b * N;
    const int s_begin = lat.seq_state_off[b];
    const int fb = lat.frame_base[b];
    const int tid_x = threadIdx.x;

    // ---- forward
    for (int s = lvl[0] + tid_x; s < lvl[1]; s += kThreads) alpha[s] = (s == s_begin) ? 0.0 : -INFINITY;
    __syncthreads();
    if (tid_x == 0)
        for (int k = eoff[0]; k < eoff[1]; ++k) {
            const int d = lat.eps_dst[k];
            alpha[d] = pk2::log_add(alpha[d], alpha[lat.eps_src[k]] - (double)(lm * lat.eps_gc[k]));
        }
    __syncthreads();
    for (int t = 1; t <= T; ++t) {
        const float* row = ll + (int64_t)(t - 1) * N;
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double acc = -INFINITY;
            for (int k = lat.in_off[s]; k < lat.in_off[s + 1]; ++k) {
                const double like = -(double)(lm * lat.in_gc[k]) +
                                    (double)ac * (double)row[lat.tid2pdf[lat.in_tid[k]]];
                acc = pk2::log_add(acc, alpha[lat.in_src[k]] + like);
            }
            alpha[s] = acc;
        }
        __syncthreads();
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0)
                for (int k = eoff[t]; k < eoff[t + 1]; ++k) {
                    const int d = lat.eps_dst[k];
                    alpha[d] = pk2::log_add(alpha[d], alpha[lat.eps_src[k]] - (double)(lm * lat.eps_gc[k]));
                }
            __syncthreads();
        }
    }
    // ---- total and beta at the last level
    double z = -INFINITY;
    for (int s = lvl[T] + tid_x; s < lvl[T + 1]; s += kThreads) {
        const float fc = lat.final_cost[s];
        const double bt = (fc < INFINITY) ? -(double)(lm * fc) : -INFINITY;
        beta[s] = bt;
        z = pk2::log_add(z, alpha[s] + bt);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z = pk2::log_add(z, __shfl_xor_sync(0xffffffffu, z, o));
    if ((tid_x & 31) == 0) s_red[tid_x >> 5] = z;
    __syncthreads();
    if (tid_x == 0) {
        double zz = s_red[0];
        for (int i = 1; i < kThreads / 32; ++i) zz = pk2::log_add(zz, s_red[i]);
        s_tot = zz;
        tot_out[b] = zz;
        for (int k = eoff[T + 1] - 1; k >= eoff[T]; --k) {      // reverse topological order
            const int s = lat.eps_src[k];
            beta[s] = pk2::log_add(beta[s], beta[lat.eps_dst[k]] - (double)(lm * lat.eps_gc[k]));
        }
    }
    __syncthreads();
    const double tot = s_tot;
    // ---- backward + posteriors
    for (int t = T - 1; t >= 0; --t) {
        const float* row = ll + (int64_t)t * N;
        float* grow = g + (int64_t)t * N;
        const bool keep = lat.keep[fb + t] != 0;
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double acc = -INFINITY;
            const double a = alpha[s];
            for (int k = lat.out_off[s]; k < lat.out_off[s + 1]; ++k) {
                const int p = lat.tid2pdf[lat.out_tid[k]];
                const double like = -(double)(lm * lat.out_gc[k]) + (double)ac * (double)row[p];
                const double ab = beta[lat.out_dst[k]] + like;
                acc = pk2::log_add(acc, ab);
                if (keep) {
                    const double post = exp(a + ab - tot);
                    if (post > 0.0) atomicAdd(&grow[p], (float)post);
                }
            }
            beta[s] = acc;
        }
        __syncthreads();
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0)
                for (int k = eoff[t + 1] - 1; k >= eoff[t]; --k) {
                    const int s = lat.eps_src[k];
                    beta[s] = pk2::log_add(beta[s], beta[lat.eps_dst[k]] - (double)(lm * lat.eps_gc[k]));
                }
            __syncthreads();
        }
    }
    // ---- numerator: -1 at (t, pdf(num_ali[t])) on kept frames
    for (int t = tid_x; t < T; t += kThreads)
        if (lat.keep[fb + t]) atomicAdd(&g[(int64_t)t * N + lat.tid2pdf[lat.num_ali[fb + t]]], -1.0f);
}

// sMBR / MPFE (reference ops/ops.py:119-156 -> Kaldi LatticeForwardBackwardMpeVariants, criterion "smbr" / "mpfe",
// one_silence_class = true).  Same level-synchronous walk, two more recursions: alpha_smbr / beta_smbr = expected
// frame accuracy of the partial paths into / out of a state.  The per-arc frame accuracy (0/1: pdf or phone of the
// arc equals that of the reference alignment, or both are silence phones) is index work done on the host
// (graphs.Lattice.frame_acc); acc_in / acc_out hold it in in-arc / out-arc order.
// grad[t, pdf] += deriv_scale * posterior(arc) * (alpha_smbr[src] + acc + beta_smbr[dst] - expected accuracy).
__global__ void __launch_bounds__(kThreads)
latfb_mpe_kernel(pk2_lat_batch lat, const uint8_t* __restrict__ acc_in, const uint8_t* __restrict__ acc_out,
                 const float* __restrict__ loglikes, int N, int64_t row_stride_b, float lm, float ac,
                 double* alpha, double* beta, double* alpha_s, double* beta_s, float deriv_scale,
                 float* __restrict__ grad, double* __restrict__ tot_out, double* __restrict__ score_out) {
    __shared__ double s_tot, s_score;
    __shared__ double s_red[kThreads / 32], s_red2[kThreads / 32];
    const int b = blockIdx.x;
    const int T = lat.num_frames[b];
    const int32_t* lvl = lat.level_off + lat.lvl_base[b];
    const int32_t* eoff = lat.eps_off + lat.lvl_base[b];
    const float* ll = loglikes + (int64_t)b * row_stride_b * N;
    float* g = grad + (int64_t)b * row_stride_b * N;
    const int s_begin = lat.seq_state_off[b];
    const int tid_x = threadIdx.x;

    // epsilon arcs of one level, serial in topological order: first alpha, then (with final alphas) alpha_smbr
    auto eps_forward = [&](int t) {
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0) {
                for (int k = eoff[t]; k < eoff[t + 1]; ++k) {
                    const int d = lat.eps_dst[k];
                    alpha[d] = pk2::log_add(alpha[d], alpha[lat.eps_src[k]] - (double)(lm * lat.eps_gc[k]));
                }
            }
            __syncthreads();
        }
    };
    auto eps_forward_s = [&](int t) {
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0) {
                for (int k = eoff[t]; k < eoff[t + 1]; ++k) {
                    const int s = lat.eps_src[k], d = lat.eps_dst[k];
                    alpha_s[d] += exp(alpha[s] - (double)(lm * lat.eps_gc[k]) - alpha[d]) * alpha_s[s];
                }
            }
            __syncthreads();
        }
    };

    // ---- forward: alpha, then alpha_smbr of the same level
    for (int s = lvl[0] + tid_x; s < lvl[1]; s += kThreads) { alpha[s] = (s == s_begin) ? 0.0 : -INFINITY; alpha_s[s] = 0.0; }
    __syncthreads();
    eps_forward(0);
    eps_forward_s(0);
    for (int t = 1; t <= T; ++t) {
        const float* row = ll + (int64_t)(t - 1) * N;
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double acc = -INFINITY;
            for (int k = lat.in_off[s]; k < lat.in_off[s + 1]; ++k) {
                const double like = -(double)(lm * lat.in_gc[k]) + (double)ac * (double)row[lat.tid2pdf[lat.in_tid[k]]];
                acc = pk2::log_add(acc, alpha[lat.in_src[k]] + like);
            }
            alpha[s] = acc;
        }
        __syncthreads();
        eps_forward(t);
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double sc = 0.0;
            const double a = alpha[s];
            for (int k = lat.in_off[s]; k < lat.in_off[s + 1]; ++k) {
                const int u = lat.in_src[k];
                const double like = -(double)(lm * lat.in_gc[k]) + (double)ac * (double)row[lat.tid2pdf[lat.in_tid[k]]];
                sc += exp(alpha[u] + like - a) * (alpha_s[u] + (double)acc_in[k]);
            }
            alpha_s[s] = sc;
        }
        __syncthreads();
        eps_forward_s(t);
    }
    // ---- totals; beta / beta_smbr at the last level
    double z = -INFINITY;
    for (int s = lvl[T] + tid_x; s < lvl[T + 1]; s += kThreads) {
        const float fc = lat.final_cost[s];
        const double bt = (fc < INFINITY) ? -(double)(lm * fc) : -INFINITY;
        beta[s] = bt;
        beta_s[s] = 0.0;
        z = pk2::log_add(z, alpha[s] + bt);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z = pk2::log_add(z, __shfl_xor_sync(0xffffffffu, z, o));
    if ((tid_x & 31) == 0) s_red[tid_x >> 5] = z;
    __syncthreads();
    if (tid_x == 0) {
        double zz = s_red[0];
        for (int i = 1; i < kThreads / 32; ++i) zz = pk2::log_add(zz, s_red[i]);
        s_tot = zz;
        tot_out[b] = zz;
    }
    __syncthreads();
    const double tot = s_tot;
    double sc_sum = 0.0;
    for (int s = lvl[T] + tid_x; s < lvl[T + 1]; s += kThreads) {
        const float fc = lat.final_cost[s];
        if (fc < INFINITY) sc_sum += exp(alpha[s] - (double)(lm * fc) - tot) * alpha_s[s];
    }
    sc_sum = pk2::warp_sum_d(sc_sum);
    if ((tid_x & 31) == 0) s_red2[tid_x >> 5] = sc_sum;
    __syncthreads();
    if (tid_x == 0) {
        double zz = 0.0;
        for (int i = 0; i < kThreads / 32; ++i) zz += s_red2[i];
        s_score = zz;
        score_out[b] = zz;
        // epsilon arcs of the last level: beta in reverse topological order, then beta_smbr with final betas
        for (int k = eoff[T + 1] - 1; k >= eoff[T]; --k) {
            const int s = lat.eps_src[k];
            beta[s] = pk2::log_add(beta[s], beta[lat.eps_dst[k]] - (double)(lm * lat.eps_gc[k]));
        }
        for (int k = eoff[T + 1] - 1; k >= eoff[T]; --k) {
            const int s = lat.eps_src[k], d = lat.eps_dst[k];
            beta_s[s] += exp(beta[d] - (double)(lm * lat.eps_gc[k]) - beta[s]) * beta_s[d];
        }
    }
    __syncthreads();
    const double score = s_score;
    // ---- backward: beta, then beta_smbr and the posteriors of the level
    for (int t = T - 1; t >= 0; --t) {
        const float* row = ll + (int64_t)t * N;
        float* grow = g + (int64_t)t * N;
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double acc = -INFINITY;
            for (int k = lat.out_off[s]; k < lat.out_off[s + 1]; ++k) {
                const double like = -(double)(lm * lat.out_gc[k]) + (double)ac * (double)row[lat.tid2pdf[lat.out_tid[k]]];
                acc = pk2::log_add(acc, beta[lat.out_dst[k]] + like);
            }
            beta[s] = acc;
        }
        __syncthreads();
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0)
                for (int k = eoff[t + 1] - 1; k >= eoff[t]; --k) {
                    const int s = lat.eps_src[k];
                    beta[s] = pk2::log_add(beta[s], beta[lat.eps_dst[k]] - (double)(lm * lat.eps_gc[k]));
                }
            __syncthreads();
        }
        for (int s = lvl[t] + tid_x; s < lvl[t + 1]; s += kThreads) {
            double bs = 0.0;
            const double a = alpha[s], as = alpha_s[s], bt = beta[s];
            for (int k = lat.out_off[s]; k < lat.out_off[s + 1]; ++k) {
                const int p = lat.tid2pdf[lat.out_tid[k]];
                const int d = lat.out_dst[k];
                const double like = -(double)(lm * lat.out_gc[k]) + (double)ac * (double)row[p];
                const double ab = beta[d] + like;
                const double fa = (double)acc_out[k];
                bs += exp(ab - bt) * (beta_s[d] + fa);
                const double post = exp(a + ab - tot) * (as + fa + beta_s[d] - score);
                if (post != 0.0) atomicAdd(&grow[p], deriv_scale * (float)post);
            }
            beta_s[s] = bs;
        }
        __syncthreads();
        if (eoff[t + 1] > eoff[t]) {
            if (tid_x == 0)
                for (int k = eoff[t + 1] - 1; k >= eoff[t]; --k) {
                    const int s = lat.eps_src[k], d = lat.eps_dst[k];
                    beta_s[s] += exp(beta[d] - (double)(lm * lat.eps_gc[k]) - beta[s]) * beta_s[d];
                }
            __syncthreads();
        }
    }
}

}  // namespace

extern "C" int pk2_latfb_mmi(const pk2_lat_batch* lat, const float* loglikes, int num_pdfs,
                             int max_frames, int64_t row_stride_b, float lm_scale, float ac_scale,
                             double* ws_alpha, double* ws_beta, float* grad, double* tot, void* stream) {
    PK2_REQUIRE(lat && loglikes && ws_alpha && ws_beta && grad && tot, "pk2_latfb_mmi: null argument");
    PK2_REQUIRE(lat->n_seq > 0 && max_frames > 0, "pk2_latfb_mmi: empty batch");
    PK2_REQUIRE(row_stride_b >= max_frames, "pk2_latfb_mmi: row_stride_b < max_frames");
    cudaStream_t st = pk2::as_stream(stream);
    // rows are addressed as b*row_stride_b + t: zero everything up to the last sequence's max_frames
    const size_t rows = (size_t)(lat->n_seq - 1) * (size_t)row_stride_b + (size_t)max_frames;
    PK2_CHECK(cudaMemsetAsync(grad, 0, rows * (size_t)num_pdfs * sizeof(float), st));
    latfb_kernel<<<lat->n_seq, kThreads, 0, st>>>(*lat, loglikes, num_pdfs, row_stride_b, lm_scale,
                                                 ac_scale, ws_alpha, ws_beta, grad, tot);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_latfb_mpe(const pk2_lat_batch* lat, const uint8_t* acc_in, const uint8_t* acc_out,
                             const float* loglikes, int num_pdfs, int max_frames, int64_t row_stride_b,
                             float lm_scale, float ac_scale, double* ws /* [4][total_states] */, int64_t total_states,
                             float deriv_scale, float* grad, double* tot_like, double* tot_score, void* stream) {
    PK2_REQUIRE(lat && acc_in && acc_out && loglikes && ws && grad && tot_like && tot_score, "pk2_latfb_mpe: null argument");
    PK2_REQUIRE(lat->n_seq > 0 && max_frames > 0 && total_states > 0, "pk2_latfb_mpe: empty batch");
    PK2_REQUIRE(row_stride_b >= max_frames, "pk2_latfb_mpe: row_stride_b < max_frames");
    cudaStream_t st = pk2::as_stream(stream);
    const size_t rows = (size_t)(lat->n_seq - 1) * (size_t)row_stride_b + (size_t)max_frames;
    PK2_CHECK(cudaMemsetAsync(grad, 0, rows * (size_t)num_pdfs * sizeof(float), st));
    latfb_mpe_kernel<<<lat->n_seq, kThreads, 0, st>>>(*lat, acc_in, acc_out, loglikes, num_pdfs, row_stride_b, lm_scale,
                                                     ac_scale, ws, ws + total_states, ws + 2 * total_states,
                                                     ws + 3 * total_states, deriv_scale, grad, tot_like, tot_score);
    PK2_POST_LAUNCH();
    return 0;
}
