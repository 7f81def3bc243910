/* pk2.h -- C ABI of libpk2.so, the B200 (sm_100a) hot path of pykaldi2.
 *
 * The reference (jzlianglu/pykaldi2) has no FFI: its hot path is Python calling
 * PyTorch/cuDNN, PyKaldi->Kaldi and Horovod.  Each entry point below names the
 * reference call site it replaces (path:line into the reference tree).  All
 * signatures are plain C: device pointers (unless suffixed _h = host pointer),
 * sizes, and a CUDA stream passed as void* (cudaStream_t; NULL = default stream).
 * Every function returns 0 on success, non-zero on error; pk2_last_error() gives
 * the message (thread-local).  Calls are asynchronous on the given stream unless
 * stated.  Tensors are owned by the caller (torch caching allocator on the Python side).
 *
 * Python binding: pykaldi2_b200/_lib.py (ctypes).  See INTEGRATION.md.
 */
#ifndef PK2_H_
#define PK2_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int pk2_version(void);
const char* pk2_last_error(void);
/* number of kernel launches this library has issued in this process (bench "gpu_launches") */
int64_t pk2_launch_count(void);

/* ---------------------------------------------------------------- fbank ----
 * Replaces DataGeneratorTrain._logfbank_extractor (data/sr_dataset.py:279-296)
 * + stft/_enframe (simulation/freq_analysis.py:41-150) on the CPU.
 * plan: Hamming(400), 256-pt twiddles and the compacted [257,80] mel matrix
 * (mel_h = host float32 [257*80], row-major, already multiplied by nothing:
 * the 32768^2 scale and the +1 floor are applied by the kernel). */
int pk2_fbank_plan_create(const float* mel_h, void** plan);
int pk2_fbank_plan_destroy(void* plan);
/* wav: concatenated float32 waveforms; wav_off[n_utts+1] sample offsets (device,
 * int64); frame_off[n_utts+1] cumulative frame counts (device, int32), frames of
 * utterance u = max(0, ceil((n_u-1-400)/160)+1); out [total_frames, 80]. */
int pk2_fbank(void* plan, const float* wav, const int64_t* wav_off, const int32_t* frame_off,
              int n_utts, int total_frames, float* out, void* stream);
/* per-utterance column means over time: reader.preprocess.cmn (reader/preprocess.py:34-41)
 * called with axis=0 at data/sr_dataset.py:365-366.  mean [n_utts, dim]. */
int pk2_colmean(const float* feats, const int32_t* frame_off, int n_utts, int dim,
                float* mean, void* stream);
/* out[r,:] = row_src[r] < 0 ? 0 : (feats[row_src[r],:] - mean[row_utt[r],:] - mvn_mean) * mvn_istd
 * Fuses CMN, GlobalMeanVarianceNormalization.apply_on_ndarray (reader/preprocess.py:211-229),
 * _utt2seg chunking (data/sr_dataset.py:40-52), SeqDataloader zero padding
 * (data/dataloader.py:96-103) and chain roll+subsample (bin/train_chain.py:251-255)
 * into one gather.  mean / mvn_mean / mvn_istd may be NULL. */
int pk2_gather_norm(const float* feats, const int32_t* row_src, const int32_t* row_utt,
                    const float* mean, const float* mvn_mean, const float* mvn_istd,
                    int n_rows, int dim, float* out, void* stream);

/* ------------------------------------------------------- CE softmax+NLL ----
 * Replaces nn.CrossEntropyLoss(ignore_index=-100) fwd+bwd (bin/train_ce.py:134,189;
 * reduction='sum' at bin/train_se.py:214,235).  loss_rows[r] = lse - logit[label]
 * (0 for ignored rows); grad = scale * (softmax - onehot) (0 for ignored rows);
 * grad may be NULL; it must not alias logits. */
int pk2_ce_softmax(const float* logits, const int64_t* labels, int64_t n_rows, int n_cols,
                   float scale, float* loss_rows, float* grad, void* stream);

/* ----------------------------------------------------- LF-MMI denominator --
 * Replaces kaldi_chain.DenominatorGraph(den_fst, num_pdfs) (bin/train_chain.py:167,202)
 * and the denominator half of kaldi_chain.compute_chain_objf_and_deriv
 * (ops/ops.py:265).  Host arrays are Kaldi's forward-transition CSR:
 * fwd_off[S+1], and per arc (prob, pdf, dst state); init[S] initial probs. */
int pk2_den_graph_create(int num_states, int num_pdfs, const int32_t* fwd_off_h,
                         const float* fwd_prob_h, const int32_t* fwd_pdf_h,
                         const int32_t* fwd_state_h, const float* init_h, void** graph);
int pk2_den_graph_destroy(void* graph);
/* bytes of alpha workspace pk2_denfb needs for n_seq sequences of at most max_frames */
size_t pk2_denfb_workspace_bytes(void* graph, int n_seq, int max_frames);
/* loglikes/grad: row t of sequence b at base + (b*row_stride_b + t)*num_pdfs floats.
 * num_frames[b] (device int32) <= max_frames.  Writes grad[b,t,:] = deriv_scale *
 * gamma_den(t,:) for t < num_frames[b] and 0 for num_frames[b] <= t < max_frames,
 * logz[b] (double) = log Z_den.  cluster = CTAs per sequence (1, 2, 4 or 8); 0 = auto.
 * 8 (and auto, where the graph qualifies: num_states % 256 == 0, num_states <= 8192, num_pdfs % 4 == 0,
 * per-CTA tables fit shared memory): persistent clusters of 8 CTAs with the arc records in registers, each
 * working through a list of sequences; with the host copy num_frames_h, auto also runs the shortest
 * sequences as single-CTA kernels on the SMs the clusters leave free.  Otherwise (1, 2, 4, or auto on other
 * graphs): streaming kernels, one cluster per sequence; with num_frames_h, auto schedules length-aware
 * (long sequences get 4-CTA clusters, short ones 1) as up to three concurrent launches. */
int pk2_denfb(void* graph, const float* loglikes, const int32_t* num_frames,
              const int32_t* num_frames_h, int n_seq, int max_frames, int64_t row_stride_b,
              float leaky, float deriv_scale, void* workspace, float* grad, double* logz,
              int cluster, void* stream);

/* ------------------------------------------------------- LF-MMI numerator --
 * Replaces the numerator half of compute_chain_objf_and_deriv (ops/ops.py:265):
 * log-domain forward-backward over per-sequence supervision FSTs (epsilon-free,
 * states sorted by time).  Concatenated CSR over all sequences:
 *  seq_state_off[n_seq+1], seq_arc_off[n_seq+1];
 *  per state (global index): out_off[S_tot+1], in_off[S_tot+1], final_cost[S_tot];
 *  state_time[S_tot]; level_off: for sequence b, states of time t are
 *  [level_off[lvl_base[b]+t], level_off[lvl_base[b]+t+1]) (global state index);
 *  out arcs (sorted by src): out_dst, out_pdf, out_w (cost); in arcs (sorted by dst):
 *  in_src, in_pdf, in_w.
 * grad[b,t,p] += deriv_scale * gamma_num(t,p)  (atomic add; call after pk2_denfb);
 * logz[b] (double) = log Z_num.  ws_alpha/ws_beta: double [S_tot] each. */
typedef struct {
    int n_seq;
    const int32_t* seq_state_off;
    const int32_t* lvl_base;
    const int32_t* level_off;
    const int32_t* num_frames;
    const int32_t* out_off; const int32_t* out_dst; const int32_t* out_pdf; const float* out_w;
    const int32_t* in_off;  const int32_t* in_src;  const int32_t* in_pdf;  const float* in_w;
    const float* final_cost;
    const int32_t* state_time;
} pk2_sup_batch;
int pk2_numfb(const pk2_sup_batch* sup, const float* loglikes, int num_pdfs, int64_t row_stride_b,
              float deriv_scale, double* ws_alpha, double* ws_beta, float* grad, double* logz,
              void* stream);
/* Kaldi's fallback for sequences whose objective is not finite (derivatives <- 0; the caller sets
 * objf <- -10 * weight * frames): zero grad[b, :, :] (row_elems = max_frames * num_pdfs floats per sequence)
 * where logz_den[b] or logz_num[b] is not finite.  Device-side, no host synchronisation. */
int pk2_chain_guard(const double* logz_den, const double* logz_num, int n_seq, int64_t row_elems,
                    float* grad, void* stream);

/* Split form, so that the numerator can run on a side stream while the denominator kernels run:
 * pk2_numfb_post computes log Z_num and the per-arc posteriors arc_post[A_tot] (no write to grad);
 * pk2_numfb_scatter adds deriv_scale * arc_post into grad afterwards (after pk2_denfb). */
int pk2_numfb_post(const pk2_sup_batch* sup, const float* loglikes, int num_pdfs, int64_t row_stride_b,
                   double* ws_alpha, double* ws_beta, float* arc_post, double* logz, void* stream);
int pk2_numfb_scatter(const pk2_sup_batch* sup, int total_states, const float* arc_post, int num_pdfs,
                      int64_t row_stride_b, float deriv_scale, float* grad, void* stream);

/* ------------------------------------------------------------ lattice MMI --
 * Replaces lattice_forward_backward_mmi + Posterior.to_pdf_matrix (ops/ops.py:57-62)
 * for a batch of topologically sorted lattices with states ordered by time.
 * Arc like = -lm_scale*graph_cost + ac_scale*loglikes[time(src), tid2pdf[tid]] (tid>0).
 * Same concatenated layout as pk2_sup_batch; eps_* list same-level epsilon arcs
 * (tid 0) of each level in topological order of src.
 * grad (pre-zeroed by this call) [b,t,:] = -(num - den) merged posteriors with
 * drop_frames / cancel semantics; keep[b*max_frames + t] = 0 for dropped frames
 * (index tensor computed on host, bit-exact); tot[b] = lattice total log-like. */
typedef struct {
    int n_seq;
    const int32_t* seq_state_off;
    const int32_t* lvl_base;
    const int32_t* level_off;
    const int32_t* num_frames;
    const int32_t* out_off; const int32_t* out_dst; const int32_t* out_tid; const float* out_gc;
    const int32_t* in_off;  const int32_t* in_src;  const int32_t* in_tid;  const float* in_gc;
    const int32_t* eps_off; const int32_t* eps_src; const int32_t* eps_dst; const float* eps_gc;
    const float* final_cost;
    const int32_t* state_time;
    const int32_t* tid2pdf;
    const int32_t* num_ali;     /* [sum T] alignment tids, offset by frame_base[b] */
    const int32_t* frame_base;  /* [n_seq+1] */
    const uint8_t* keep;        /* [sum T] 1 = frame kept, 0 = dropped */
} pk2_lat_batch;
/* ws: pk2_latfb_workspace_bytes(total_states, total_arcs, mpe) bytes of device memory (alpha / beta per state,
 * per-arc scores in in-arc and out-arc order, per-arc pdf); total_arcs = number of non-epsilon arcs
 * (= in_off[total_states]), total_frames = sum of num_frames. */
size_t pk2_latfb_workspace_bytes(int64_t total_states, int64_t total_arcs, int mpe);
int pk2_latfb_mmi(const pk2_lat_batch* lat, const float* loglikes, int num_pdfs, int max_frames,
                  int64_t row_stride_b, float lm_scale, float ac_scale,
                  void* ws, int64_t total_states, int64_t total_arcs, int64_t total_frames,
                  float* grad, double* tot, void* stream);

/* sMBR / MPFE.  Replaces lattice_forward_backward_mpe_variants(trans_model, silence_phones,
 * lattice, trans_ids, criterion, one_silence_class=True) + Posterior.to_pdf_matrix of sMBRFunction.forward
 * (reference ops/ops.py:130-147; Kaldi lat/lattice-functions.cc LatticeForwardBackwardMpeVariants on the CPU).
 * acc_in / acc_out: per non-epsilon arc frame accuracy (0/1) in the order of lat->in_* / lat->out_* (host index
 * work: graphs.Lattice.frame_acc).  ws: pk2_latfb_workspace_bytes(total_states, total_arcs, 1) bytes.  grad (zeroed by the call)[b,t,pdf] =
 * deriv_scale * sum over the arcs (t, pdf) of posterior * (accuracy through the arc - expected accuracy);
 * tot_like[b] = lattice log-likelihood, tot_score[b] = expected frame accuracy (the op's return value).
 * The reference applies no lattice_scale on this path: pass lm_scale = ac_scale = 1. */
int pk2_latfb_mpe(const pk2_lat_batch* lat, const uint8_t* acc_in, const uint8_t* acc_out,
                  const float* loglikes, int num_pdfs, int max_frames, int64_t row_stride_b,
                  float lm_scale, float ac_scale, void* ws, int64_t total_states, int64_t total_arcs,
                  float deriv_scale, float* grad, double* tot_like, double* tot_score, void* stream);

/* ------------------------------------------------------------------ BLSTM --
 * Replaces nn.LSTM(batch_first, bidirectional) + nn.Linear forward/backward
 * (models/lstm.py:46-61 -> cuDNN RNN / cuBLAS).  See pykaldi2_b200/csrc/blstm.cu. */
/* C[M,N] (+)= A[M,K] * B[N,K]^T (+ bias[N]), bf16 operands row-major (K contiguous),
 * fp32 accumulate on tcgen05 tensor cores; C fp32 or bf16 (c_bf16).
 * flags: bit0 accumulate into C, bit1 C is bf16.  */
/* cap on the CTAs of later GEMM launches of the calling thread (0 = all SMs) */
int pk2_gemm_set_max_ctas(int n);
int pk2_gemm_bf16_nt(const void* A, const void* B, void* C, const float* bias,
                     int M, int N, int K, int lda, int ldb, int ldc, int flags, void* stream);
/* Extended form.  flags bit2 (PK2_GEMM_TN): C[M,N] = At[K,M]^T * Bt[K,N] -- both operands are stored with the
 * contraction index as the row (lda / ldb = row strides >= M / N) and are consumed in place through MN-major tcgen05
 * descriptors: the weight gradients dW = dY^T X of nn.Linear / nn.LSTM backward (reference models/lstm.py:46-61 ->
 * cuBLAS gemm with op(A) = T), whose operands are the activation matrices [frames, features] as they lie in memory.
 * c_row_map (device int32[M], may be NULL): row r of the product goes to row c_row_map[r] of C -- the output layer is
 * evaluated on the valid (unpadded) frames only and scattered back into the padded [B, Tmax, N] layout. */
#define PK2_GEMM_BF16_OUT 2
#define PK2_GEMM_TN 4
int pk2_gemm_bf16_ex(const void* A, const void* B, void* C, const float* bias,
                     int M, int N, int K, int lda, int ldb, int ldc, int flags, const int32_t* c_row_map, void* stream);

/* x[B*T, I] (bf16, row stride ldx) * W_ih_cat[8H, I]^T + bias[8H] -> gx in the recurrent kernel's
 * layout [T][2][H/32][B][128] fp32 (128 = 4 gates x 32 units of one CTA).  W_ih_cat = [W_ih_fwd; W_ih_bwd],
 * bias = b_ih + b_hh of both directions. */
int pk2_lstm_input_proj(const void* x, const void* wih, const float* bias, float* gx, int B, int T,
                        int I, int H, int ldx, void* stream);

typedef struct {
    int B, T, H;            /* batch, time steps, hidden size per direction (multiple of 64, <= 512) */
    const float* gx;        /* [T][2][H/32][B][128] fp32 input projections incl. both biases */
    const void* whh;        /* bf16 [2*4H, H] recurrent weights, rows packed per CTA:
                               row (dir*H/32 + cta)*128 + gate*32 + ul = W_hh[dir][gate*H + cta*32 + ul] */
    void* y;                /* bf16 [B,T,2H] layer output (fwd | bwd halves); also the h exchange */
    void* gates;            /* bf16 [2,T,B,4,H] post-activation gates i,f,g,o (saved for backward) */
    float* cstate;          /* fp32 [2,T,B,H] cell states (saved for backward) */
    unsigned int* sync;     /* [2*ceil(B/32)] step counters (zeroed by the call) */
} pk2_lstm_fwd_args;
int pk2_lstm_layer_fwd(const pk2_lstm_fwd_args* a, void* stream);
/* profiling aid: device int64[128] receiving clock64() stamps of 8 steps of one CTA; NULL = off */
int pk2_lstm_set_profile_buffer(void* buf);
/* Host-side schedule of pk2_denfb's register-resident path (no device work; used by the CPU tests):
 * assign[i] = cluster that processes sequence i, or -1 = single-CTA kernel on one of `spare_sms` SMs.
 * Returns the largest cluster load in frames, -1 on bad arguments. */
long long pk2_den_plan(const int32_t* frames_h, int n_seq, int n_clusters, int spare_sms, int32_t* assign_h);

/* SM budget of later pk2_denfb calls on this graph (register-resident path): at most `max_clusters` resident
 * clusters of 8 CTAs (0 = as many as fit) and `reserve_sms` SMs kept free of single-CTA kernels, so that kernels of
 * OTHER streams (the 16-CTA BLSTM recurrence clusters of the other half-batch, pipeline.chain_step_overlapped) can
 * run next to the denominator.  No reference counterpart: Kaldi's chain computation owns the GPU
 * (reference ops/ops.py:255-269 is a blocking call). */
int pk2_den_set_sm_budget(void* graph, int max_clusters, int reserve_sms);

/* profiling aid: device int64[128]; frames 64..71 of the first cluster of every pk2_denfb launch stamp clock64()
 * at 7 points of the forward ([0,64)) and backward ([64,128)) frame loop; NULL = off */
int pk2_den_set_profile_buffer(void* buf);

typedef struct {
    int B, T, H;
    const float* dy;        /* fp32 [B,T,2H] grad wrt layer output */
    const void* whh_t;      /* bf16 [2*H, 4H]: row dir*H + j = W_hh[dir][:, j] (transposed recurrent weights) */
    const void* gates;      /* from forward */
    const float* cstate;    /* from forward */
    void* dgates;           /* bf16 [B,T,2,4H] grad wrt gate pre-activations (output; also the exchange) */
    unsigned int* sync;     /* [2*ceil(B/32)] */
    const void* whh_t_perm; /* bf16 [2*H, 4H]: whh_t with the 4H index permuted to cta*128 + gate*32 + unit
                               (cta = unit/32): operand of the cluster/DSMEM kernel; NULL = global-memory kernel */
} pk2_lstm_bwd_args;
int pk2_lstm_layer_bwd(const pk2_lstm_bwd_args* a, void* stream);

/* elementwise / layout helpers of the BLSTM path */
int pk2_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
/* src fp32 or bf16 [R, C] (row stride lds) -> dst bf16 [C, ldd] (dst[c, r]); src_bf16 selects the input type */
int pk2_transpose_bf16(const void* src, int src_bf16, void* dst, int R, int C, int lds, int ldd, void* stream);
/* hprevT[dir][j][b*T+t] = y[b][t -/+ 1][dir*H + j] (0 at the sequence boundary): the h_{t-1} operand of dW_hh */
int pk2_lstm_hprev_t(const void* y, void* hprev_t, int B, int T, int H, int ldd, void* stream);
/* bf16 operand copies of one BLSTM layer's weights, one launch.  params: host array of 8 device pointers in nn.LSTM's
 * order (weight_ih, weight_hh, bias_ih, bias_hh of the forward, then of the reverse direction; fp32, reference
 * models/lstm.py:46-52).  Outputs:
 *   wih_cat [8H, I]      rows = [W_ih fwd; W_ih rev]                          (input projection, B operand)
 *   bias_cat[8H] fp32    b_ih + b_hh of both directions
 *   whh_p   [2*4H, H]    recurrent weights in per-CTA order: row (dir*H/32 + cta)*128 + gate*32 + ul
 *   whh_t   [2H, 4H]     W_hh^T per direction                                 (backward recurrence)
 *   whh_tp  [2H, 4H]     W_hh^T with the 4H index in per-CTA order            (cluster backward kernel)
 *   wih_t   [I, ldk]     [W_ih fwd; W_ih rev]^T, may be NULL (bottom layer)   (input gradient GEMM) */
int pk2_lstm_pack_layer(const float* const* params, int H, int I, void* wih_cat, float* bias_cat, void* whh_p,
                        void* whh_t, void* whh_tp, void* wih_t, int ldk, void* stream);
/* hprev[b*T+t][dir*H + j] = y[b][t -/+ 1][dir*H + j] (0 at the sequence boundary), bf16 [B*T, 2H]: the h_{t-1} operand
 * of dW_hh in the layout the TN GEMM consumes in place */
int pk2_lstm_hprev(const void* y, void* hprev, int B, int T, int H, void* stream);
/* dst[r, :] = bf16(src[rows[r], :]), src fp32 or bf16 [*, C] contiguous rows, rows device int32[R] or NULL (identity):
 * compaction of the valid (unpadded) frames of dlogits / the top layer's output for the output-layer GEMMs */
int pk2_gather_rows_bf16(const void* src, int src_bf16, const int32_t* rows, void* dst, int64_t R, int C, void* stream);
/* zero rows t >= lengths[b] of p[B][T][row_bytes] (lengths device int32[B], row_bytes % 16 == 0) */
int pk2_zero_pad_rows(void* p, const int32_t* lengths, int B, int T, int64_t row_bytes, void* stream);
/* out[c] = sum_r src[r, c]  (bf16 src [R, C], fp32 out) */
int pk2_colsum_bf16(const void* src, float* out, int64_t R, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PK2_H_ */
