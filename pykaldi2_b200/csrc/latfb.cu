// Lattice forward-backward + MMI posterior merge on the GPU (sm_100a).
//
// Replaces lattice_forward_backward_mmi(..., drop_frames=True, convert_to_pdf_ids=False,
// cancel=True) + Posterior.to_pdf_matrix (reference ops/ops.py:57-62; Kaldi
// lat/lattice-functions.cc + hmm/posterior.cc on the CPU) and removes the per-utterance
// D2H/H2D copies of ops/ops.py:55,64.  Recursion: SURVEY.md Appendix B, alpha/beta in double.
// Lattice states are time-stamped, so the recursion is level-synchronous.  Output: dense gradient
// rows  grad[t,:] = den_post(t,:) - num_post(t,:)  for kept frames, 0 for dropped frames.
//
// Round-2 structure (three launches, the serial part reduced to what is serial):
//   lat_arc_like_kernel   one thread per state, all SMs: per-arc score  -lm*graph + ac*loglike[t, pdf(tid)]  for the
//                         in-arc and the out-arc order and the arc's pdf -- the only part that touches the T x N
//                         log-likelihood matrix (random 4-byte gathers, independent of the recursion)
//   lat_chain_kernel      TWO CTAs per utterance, running concurrently: the alpha chain (levels 0..T, in-arc CSR) and the
//                         beta chain (levels T..0, out-arc CSR) do not depend on each other.  Per level a thread owns one
//                         state; its arcs (source / destination state, score) were prefetched into registers one level
//                         ahead and the CSR offsets two levels ahead, so the per-level critical path is: alpha of the
//                         previous level from L1 -> max / expf / logf -> store -> one block barrier.  Values are kept
//                         in double, the transcendental part of each log-sum-exp runs in fp32 on the shifted terms
//                         (|error| ~1e-7 per level).  Same-level epsilon arcs (rare) are applied serially in
//                         topological order between levels.
//   lat_post_kernel       one thread per state, all SMs: arc posteriors exp(alpha + score + beta - tot) scattered into
//                         the dense gradient rows (atomicAdd: two arcs of a frame may share a pdf), numerator -1
// The round-1 kernels (one CTA per utterance doing everything, fp64 transcendental per arc, 4 dependent global loads
// per level: 24.6 ms against 2.6-2.8 ms on the C3 batch, profiles/lattice_kernels_r2_v5.jsonl) were removed after the A/B.
#include "common.cuh"

namespace {

// =================================================================== round-2 kernels ====
constexpr int kChain = 128;      // threads of a chain CTA (one state per thread and level; wider levels loop)
constexpr int kPF = 6;           // arcs per state staged in shared memory ahead of time
constexpr int kDepth = 7;        // levels the asynchronous arc copies run ahead of the recursion (kDepth + 1 = 8 slots)
constexpr int kValW = 1024;      // states per level whose value is exchanged through shared memory
constexpr int kOffSlots = 16;    // >= 2 * kDepth, power of two

// one arc as the chain and posterior kernels read it: score, state at the other end (global index, and its index
// inside its level | frame accuracy of sMBR / MPFE << 30)
struct __align__(16) ArcRec { double like; int peer; int aux; };
__device__ __forceinline__ int rec_local(const ArcRec& r) { return r.aux & 0x3fffffff; }
__device__ __forceinline__ int rec_acc(const ArcRec& r) { return r.aux >> 30; }

__device__ __forceinline__ int seq_of_state(const pk2_lat_batch& lat, int s) {
    int lo = 0, hi = lat.n_seq - 1;                   // largest b with seq_state_off[b] <= s
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&lat.seq_state_off[mid]) <= s) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
lat_arc_like_kernel(pk2_lat_batch lat, const uint8_t* __restrict__ acc_in, const uint8_t* __restrict__ acc_out,
                    const float* __restrict__ loglikes, int N, int64_t row_stride_b, float lm, float ac,
                    ArcRec* __restrict__ rec_in, ArcRec* __restrict__ rec_out, int32_t* __restrict__ pdf_out,
                    int total_states) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total_states) return;
    const int b = seq_of_state(lat, s);
    const int t = lat.state_time[s];
    const float* ll = loglikes + (int64_t)b * row_stride_b * N;
    const int32_t* lvl = lat.level_off + lat.lvl_base[b];
    if (t < lat.num_frames[b]) {
        const float* row = ll + (int64_t)t * N;
        const int next0 = lvl[t + 1];                     // out-arcs end in level t + 1
        for (int k = lat.out_off[s]; k < lat.out_off[s + 1]; ++k) {
            const int p = __ldg(&lat.tid2pdf[lat.out_tid[k]]);
            ArcRec r;
            r.like = -(double)(lm * lat.out_gc[k]) + (double)ac * (double)__ldg(&row[p]);
            r.peer = lat.out_dst[k];
            r.aux = (r.peer - next0) | ((acc_out ? (int)acc_out[k] : 0) << 30);
            rec_out[k] = r;
            pdf_out[k] = p;
        }
    }
    if (t >= 1) {
        const float* row = ll + (int64_t)(t - 1) * N;
        const int prev0 = lvl[t - 1];                     // in-arcs start in level t - 1
        for (int k = lat.in_off[s]; k < lat.in_off[s + 1]; ++k) {
            ArcRec r;
            r.like = -(double)(lm * lat.in_gc[k]) + (double)ac * (double)__ldg(&row[lat.tid2pdf[lat.in_tid[k]]]);
            r.peer = lat.in_src[k];
            r.aux = (r.peer - prev0) | ((acc_in ? (int)acc_in[k] : 0) << 30);
            rec_in[k] = r;
        }
    }
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// shared-memory carve-up of a chain CTA
struct ChainSmem {
    int* lvl;          // [T + 2] first state of every level
    int* eoff;         // [T + 2] first epsilon arc of every level
    double* sv;        // [2][kValW] value (alpha / beta) of the previous and the current level
    double* sv_s;      // [2][kValW] accuracy recursion (sMBR / MPFE)
    int2* off;         // [kOffSlots][kChain] CSR range of the thread's state, 2 * kDepth levels ahead
    ArcRec* rec;       // [kDepth + 1][kPF][kChain] the state's first kPF arcs, kDepth levels ahead
};

// One direction of the recursion for one utterance.  FWD: level t from the in-arcs (sources at level t-1);
// !FWD: level t from the out-arcs (destinations at level t+1).  `val` = alpha or beta, `val_s` = the expected-accuracy
// recursion of sMBR / MPFE (MPE only).  The arcs of level t are copied into the thread's private shared-memory slots
// kDepth levels ahead with cp.async (the CSR offsets they depend on 2 * kDepth levels ahead), so the loads never sit
// on the dependency chain; the previous level's values come from shared memory.
template <bool FWD, bool MPE>
__device__ void lat_chain(const pk2_lat_batch& lat, int b, const int32_t* __restrict__ arc_off,
                          const ArcRec* __restrict__ arc_rec, float lm, double* val, double* val_s,
                          double* __restrict__ tot_out, double* __restrict__ score_out, const ChainSmem& sm) {
    __shared__ double s_red[kChain / 32], s_red2[kChain / 32];
    __shared__ double s_tot;
    const int T = lat.num_frames[b];
    const int tid = threadIdx.x;
    int* s_lvl = sm.lvl;
    int* s_eoff = sm.eoff;
    {
        const int32_t* lvl_g = lat.level_off + lat.lvl_base[b];
        const int32_t* eoff_g = lat.eps_off + lat.lvl_base[b];
        for (int i = tid; i < T + 2; i += kChain) { s_lvl[i] = lvl_g[i]; s_eoff[i] = eoff_g[i]; }
    }
    __syncthreads();

    // value of a state of level `lt` (the level being computed, or its neighbour): shared memory for the first kValW
    // states of a level, global memory beyond
    auto sv_slot = [&](int lt) { return (lt & 1) * kValW; };
    auto get = [&](int s, int lt) { const int li = s - s_lvl[lt]; return li < kValW ? sm.sv[sv_slot(lt) + li] : val[s]; };
    auto get_s = [&](int s, int lt) { const int li = s - s_lvl[lt]; return li < kValW ? sm.sv_s[sv_slot(lt) + li] : val_s[s]; };
    auto put = [&](int s, int lt, double v) { const int li = s - s_lvl[lt]; if (li < kValW) sm.sv[sv_slot(lt) + li] = v; val[s] = v; };
    auto put_s = [&](int s, int lt, double v) { const int li = s - s_lvl[lt]; if (li < kValW) sm.sv_s[sv_slot(lt) + li] = v; val_s[s] = v; };

    // same-level epsilon arcs, serial in (reverse) topological order; first the log-domain value, then, with those
    // final, the accuracy recursion
    auto eps = [&](int t) {
        if (s_eoff[t + 1] > s_eoff[t]) {
            if (tid == 0) {
                if (FWD) {
                    for (int k = s_eoff[t]; k < s_eoff[t + 1]; ++k) {
                        const int u = lat.eps_src[k], d = lat.eps_dst[k];
                        put(d, t, pk2::log_add(get(d, t), get(u, t) - (double)(lm * lat.eps_gc[k])));
                    }
                } else {
                    for (int k = s_eoff[t + 1] - 1; k >= s_eoff[t]; --k) {
                        const int u = lat.eps_src[k], d = lat.eps_dst[k];
                        put(u, t, pk2::log_add(get(u, t), get(d, t) - (double)(lm * lat.eps_gc[k])));
                    }
                }
            }
            __syncthreads();
        }
    };
    auto eps_s = [&](int t) {
        if (MPE && s_eoff[t + 1] > s_eoff[t]) {
            if (tid == 0) {
                if (FWD) {
                    for (int k = s_eoff[t]; k < s_eoff[t + 1]; ++k) {
                        const int u = lat.eps_src[k], d = lat.eps_dst[k];
                        put_s(d, t, get_s(d, t) + exp(get(u, t) - (double)(lm * lat.eps_gc[k]) - get(d, t)) * get_s(u, t));
                    }
                } else {
                    for (int k = s_eoff[t + 1] - 1; k >= s_eoff[t]; --k) {
                        const int u = lat.eps_src[k], d = lat.eps_dst[k];
                        put_s(u, t, get_s(u, t) + exp(get(d, t) - (double)(lm * lat.eps_gc[k]) - get(u, t)) * get_s(d, t));
                    }
                }
            }
            __syncthreads();
        }
    };
    const int step = FWD ? 1 : -1;
    // a state whose arcs are not staged (levels wider than the CTA): arcs straight from memory.  pass 0: value,
    // pass 1 (MPE): accuracy recursion with the weights normalised by their own sum (see below)
    auto slow_state = [&](int s, int t) {
        const int k0 = arc_off[s], k1 = arc_off[s + 1];
        double m = -INFINITY;
        for (int k = k0; k < k1; ++k) m = fmax(m, get(arc_rec[k].peer, t - step) + arc_rec[k].like);
        float sum = 0.f;
        if (m > -INFINITY) for (int k = k0; k < k1; ++k) sum += expf((float)(get(arc_rec[k].peer, t - step) + arc_rec[k].like - m));
        put(s, t, (m > -INFINITY) ? m + (double)logf(sum) : -INFINITY);
    };
    auto slow_state_s = [&](int s, int t) {
        const int k0 = arc_off[s], k1 = arc_off[s + 1];
        double m = -INFINITY;
        for (int k = k0; k < k1; ++k) m = fmax(m, get(arc_rec[k].peer, t - step) + arc_rec[k].like);
        double num = 0.0, den = 0.0;
        if (m > -INFINITY)
            for (int k = k0; k < k1; ++k) {
                const ArcRec r = arc_rec[k];
                const double w = exp(get(r.peer, t - step) + r.like - m);
                den += w;
                num += w * (get_s(r.peer, t - step) + (double)rec_acc(r));
            }
        double sc = 0.0;
        if (den > 0.0) {
            sc = num / den;
            // the value stored by slow_state is m + logf(fp32 sum); v exceeds it when epsilon arcs enter the state
            float sumf = 0.f;
            for (int k = k0; k < k1; ++k) sumf += expf((float)(get(arc_rec[k].peer, t - step) + arc_rec[k].like - m));
            const double a_reg = m + (double)logf(sumf), v = get(s, t);
            if (v != a_reg) sc *= exp(a_reg - v);
        }
        put_s(s, t, sc);
    };

    // ---- first level
    const int t_first = FWD ? 0 : T;
    if (FWD) {
        const int s_begin = lat.seq_state_off[b];
        for (int s = s_lvl[0] + tid; s < s_lvl[1]; s += kChain) { put(s, 0, (s == s_begin) ? 0.0 : -INFINITY); if (MPE) put_s(s, 0, 0.0); }
    } else {
        for (int s = s_lvl[T] + tid; s < s_lvl[T + 1]; s += kChain) {
            const float fc = lat.final_cost[s];
            put(s, T, (fc < INFINITY) ? -(double)(lm * fc) : -INFINITY);
            if (MPE) put_s(s, T, 0.0);
        }
    }
    __syncthreads();
    eps(t_first);
    eps_s(t_first);

    // ---- asynchronous staging.  Level index n = 1, 2, ... <-> level t = t_first + n * step, n <= T.
    // Offsets ring: kOffSlots (16) slots, filled 2 * kDepth levels ahead; arc ring: kDepth + 1 (8) slots.
    static_assert(kOffSlots >= 2 * kDepth + 1 && (kOffSlots & (kOffSlots - 1)) == 0, "offset ring");
    static_assert(((kDepth + 1) & kDepth) == 0, "arc ring depth must be a power of two");
    auto level_of = [&](int n) { return t_first + n * step; };
    auto issue_off = [&](int n) {             // CSR range of this thread's state of level index n
        int2* dst = &sm.off[(n & (kOffSlots - 1)) * kChain + tid];
        bool ok = false;
        if (n <= T) {
            const int t = level_of(n);
            const int s = s_lvl[t] + tid;
            if (s < s_lvl[t + 1]) { cp_async4(&dst->x, arc_off + s); cp_async4(&dst->y, arc_off + s + 1); ok = true; }
        }
        if (!ok) *dst = make_int2(0, 0);
    };
    auto issue_arcs = [&](int n, int2 o) {    // first kPF arcs of that state
        ArcRec* dst = sm.rec + (size_t)(n & kDepth) * kPF * kChain + tid;
        const int cnt = o.y - o.x;
#pragma unroll
        for (int i = 0; i < kPF; ++i) if (i < cnt) cp_async16(dst + i * kChain, arc_rec + o.x + i);
    };
    for (int n = 1; n <= 2 * kDepth; ++n) issue_off(n);
    cp_async_commit();
    cp_async_wait<0>();
    for (int n = 1; n <= kDepth; ++n) issue_arcs(n, sm.off[(n & (kOffSlots - 1)) * kChain + tid]);
    cp_async_commit();
    cp_async_wait<0>();

    for (int n = 1; n <= T; ++n) {
        const int t = level_of(n);
        cp_async_wait<kDepth - 1>();                      // the copies issued kDepth iterations ago have landed
        const int2 oc = sm.off[(n & (kOffSlots - 1)) * kChain + tid];                    // this level
        const int2 ob = sm.off[((n + kDepth) & (kOffSlots - 1)) * kChain + tid];         // kDepth levels ahead
        issue_arcs(n + kDepth, ob);
        issue_off(n + 2 * kDepth);
        cp_async_commit();

        const int lv0 = s_lvl[t], lv1 = s_lvl[t + 1];
        const int pw = s_lvl[t - step + 1] - s_lvl[t - step];          // width of the neighbouring level
        const int s = lv0 + tid;
        const int c_k0 = oc.x, c_n = oc.y - oc.x;
        const ArcRec* mine = sm.rec + (size_t)(n & kDepth) * kPF * kChain + tid;
        // fast path (block-uniform): every state of the level has its own thread and the neighbouring level's values
        // all live in shared memory -- plain indexed reads, no range checks
        const bool fast = (lv1 - lv0 <= kChain) && (pw <= kValW);
        const double* pv = sm.sv + sv_slot(t - step);
        double x[kPF], m = -INFINITY;
        float sum = 0.f;
        if (s < lv1) {
            if (fast) {
#pragma unroll
                for (int i = 0; i < kPF; ++i) {
                    x[i] = -INFINITY;
                    if (i < c_n) {
                        const ArcRec r = mine[i * kChain];
                        x[i] = pv[rec_local(r)] + r.like;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < kPF; ++i) {
                    x[i] = -INFINITY;
                    if (i < c_n) { const ArcRec r = mine[i * kChain]; x[i] = get(r.peer, t - step) + r.like; }
                }
            }
            m = fmax(fmax(fmax(x[0], x[1]), fmax(x[2], x[3])), fmax(x[4], x[5]));
            for (int k = c_k0 + kPF; k < c_k0 + c_n; ++k) m = fmax(m, get(arc_rec[k].peer, t - step) + arc_rec[k].like);
            if (m > -INFINITY) {
                // terms at -inf give exp(-inf) = 0; the spread inside one log-sum-exp is what fp32 sees
#pragma unroll
                for (int i = 0; i < kPF; ++i) sum += __expf((float)(x[i] - m));
                for (int k = c_k0 + kPF; k < c_k0 + c_n; ++k) sum += __expf((float)(get(arc_rec[k].peer, t - step) + arc_rec[k].like - m));
            }
            const double v = (m > -INFINITY) ? m + (double)logf(sum) : -INFINITY;
            if (fast) { sm.sv[sv_slot(t) + tid] = v; val[s] = v; } else put(s, t, v);
        }
        for (int s2 = s + kChain; s2 < lv1; s2 += kChain) slow_state(s2, t);
        __syncthreads();
        eps(t);
        if (MPE) {
            // accuracy recursion: sum_k w_k (val_s[peer_k] + acc_k), w_k = exp(x_k - value).  The weights are formed in
            // DOUBLE as exp(x_k - m) / sum_k exp(x_k - m): exactly normalised and accurate to 1e-16 -- the recursion
            // runs over thousands of levels and the posterior below multiplies DIFFERENCES of these values
            // (fp32 weights left 1e-6 absolute errors in the sMBR derivatives at T = 800)
            if (s < lv1) {
                double num = 0.0, den = 0.0;
                if (m > -INFINITY) {
#pragma unroll
                    for (int i = 0; i < kPF; ++i)
                        if (i < c_n) {
                            const ArcRec r = mine[i * kChain];
                            const double w = exp(x[i] - m);
                            den += w;
                            num += w * (get_s(r.peer, t - step) + (double)rec_acc(r));
                        }
                    for (int k = c_k0 + kPF; k < c_k0 + c_n; ++k) {
                        const ArcRec r = arc_rec[k];
                        const double w = exp(get(r.peer, t - step) + r.like - m);
                        den += w;
                        num += w * (get_s(r.peer, t - step) + (double)rec_acc(r));
                    }
                }
                double sc = 0.0;
                if (den > 0.0) {
                    sc = num / den;
                    const double a_reg = m + (double)logf(sum), v = get(s, t);
                    if (v != a_reg) sc *= exp(a_reg - v);          // epsilon arcs entered this state
                }
                put_s(s, t, sc);
            }
            for (int s2 = s + kChain; s2 < lv1; s2 += kChain) slow_state_s(s2, t);
            __syncthreads();
            eps_s(t);
        }
    }
    cp_async_wait<0>();

    if (FWD) {
        // ---- total log-likelihood (and expected accuracy) from the final states
        double z = -INFINITY;
        for (int s = s_lvl[T] + tid; s < s_lvl[T + 1]; s += kChain) {
            const float fc = lat.final_cost[s];
            if (fc < INFINITY) z = pk2::log_add(z, val[s] - (double)(lm * fc));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z = pk2::log_add(z, __shfl_xor_sync(0xffffffffu, z, o));
        if ((tid & 31) == 0) s_red[tid >> 5] = z;
        __syncthreads();
        if (tid == 0) {
            double zz = s_red[0];
            for (int i = 1; i < kChain / 32; ++i) zz = pk2::log_add(zz, s_red[i]);
            s_tot = zz;
            tot_out[b] = zz;
        }
        __syncthreads();
        if (MPE) {
            const double tot = s_tot;
            double sc = 0.0;
            for (int s = s_lvl[T] + tid; s < s_lvl[T + 1]; s += kChain) {
                const float fc = lat.final_cost[s];
                if (fc < INFINITY) sc += exp(val[s] - (double)(lm * fc) - tot) * val_s[s];
            }
            sc = pk2::warp_sum_d(sc);
            if ((tid & 31) == 0) s_red2[tid >> 5] = sc;
            __syncthreads();
            if (tid == 0) {
                double zz = 0.0;
                for (int i = 0; i < kChain / 32; ++i) zz += s_red2[i];
                score_out[b] = zz;
            }
        }
    }
}

struct LatWs {
    double *alpha, *beta, *alpha_s, *beta_s;
    ArcRec *rec_in, *rec_out;
    int32_t* pdf_out;
};

size_t chain_smem_bytes(int max_frames, bool mpe) {
    size_t b = 2 * (size_t)(max_frames + 2) * sizeof(int);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(mpe ? 4 : 2) * kValW * sizeof(double);
    b += (size_t)kOffSlots * kChain * sizeof(int2);
    b += (size_t)(kDepth + 1) * kPF * kChain * sizeof(ArcRec);
    return b;
}

template <bool MPE>
__global__ void __launch_bounds__(kChain)
lat_chain_kernel(pk2_lat_batch lat, LatWs w, float lm, double* __restrict__ tot_out, double* __restrict__ score_out,
                 int max_frames) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    ChainSmem sm;
    size_t o = 0;
    sm.lvl = reinterpret_cast<int*>(s_dyn);
    sm.eoff = sm.lvl + (max_frames + 2);
    o = (2 * (size_t)(max_frames + 2) * sizeof(int) + 15) & ~(size_t)15;
    sm.sv = reinterpret_cast<double*>(s_dyn + o); o += 2 * kValW * sizeof(double);
    sm.sv_s = reinterpret_cast<double*>(s_dyn + o); if (MPE) o += 2 * kValW * sizeof(double);
    sm.off = reinterpret_cast<int2*>(s_dyn + o); o += (size_t)kOffSlots * kChain * sizeof(int2);
    sm.rec = reinterpret_cast<ArcRec*>(s_dyn + o);
    const int b = blockIdx.x >> 1;
    if ((blockIdx.x & 1) == 0)
        lat_chain<true, MPE>(lat, b, lat.in_off, w.rec_in, lm, w.alpha, w.alpha_s, tot_out, score_out, sm);
    else
        lat_chain<false, MPE>(lat, b, lat.out_off, w.rec_out, lm, w.beta, w.beta_s, tot_out, score_out, sm);
}

template <bool MPE>
__global__ void __launch_bounds__(256)
lat_post_kernel(pk2_lat_batch lat, LatWs w, int N, int64_t row_stride_b,
                float deriv_scale, const double* __restrict__ tot_in, const double* __restrict__ score_in,
                float* __restrict__ grad, int total_states, int total_frames) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < total_states) {
        const int b = seq_of_state(lat, s);
        const int t = lat.state_time[s];
        if (t < lat.num_frames[b] && (MPE || lat.keep[lat.frame_base[b] + t] != 0)) {
            float* grow = grad + ((int64_t)b * row_stride_b + t) * N;
            const double a = w.alpha[s] - tot_in[b];
            if (a > -INFINITY) {
                const double as = MPE ? (w.alpha_s[s] - score_in[b]) : 0.0;
                for (int k = lat.out_off[s]; k < lat.out_off[s + 1]; ++k) {
                    const ArcRec r = w.rec_out[k];
                    const float post = expf((float)(a + r.like + w.beta[r.peer]));
                    if (MPE) {
                        const float v = post * (float)(as + (double)rec_acc(r) + w.beta_s[r.peer]);
                        if (v != 0.f) atomicAdd(&grow[w.pdf_out[k]], deriv_scale * v);
                    } else if (post > 0.f) {
                        atomicAdd(&grow[w.pdf_out[k]], post);
                    }
                }
            }
        }
    }
    if (!MPE && s < total_frames) {
        // numerator: -1 at (t, pdf(num_ali[t])) on kept frames; s enumerates the frames of the batch
        if (lat.keep[s]) {
            int lo = 0, hi = lat.n_seq - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (__ldg(&lat.frame_base[mid]) <= s) lo = mid; else hi = mid - 1;
            }
            const int t = s - lat.frame_base[lo];
            atomicAdd(&grad[((int64_t)lo * row_stride_b + t) * N + lat.tid2pdf[lat.num_ali[s]]], -1.0f);
        }
    }
}

size_t lat_ws_layout(int64_t S, int64_t A, bool mpe, void* base, LatWs* w) {
    const size_t nd = (size_t)(mpe ? 4 : 2) * S;
    const size_t vals = (nd * sizeof(double) + 15) & ~(size_t)15;
    if (w) {
        double* d = static_cast<double*>(base);
        w->alpha = d; w->beta = d + S;
        w->alpha_s = mpe ? d + 2 * S : nullptr; w->beta_s = mpe ? d + 3 * S : nullptr;
        ArcRec* arcs = reinterpret_cast<ArcRec*>(static_cast<char*>(base) + vals);
        w->rec_in = arcs; w->rec_out = arcs + A;
        w->pdf_out = reinterpret_cast<int32_t*>(arcs + 2 * A);
    }
    return vals + 2 * (size_t)A * sizeof(ArcRec) + (size_t)A * sizeof(int32_t) + 16;
}

template <bool MPE>
int launch_lat_v1(const pk2_lat_batch* lat, const uint8_t* acc_in, const uint8_t* acc_out, const float* loglikes,
                  int num_pdfs, int max_frames, int64_t row_stride_b, float lm, float ac, void* ws, int64_t S, int64_t A,
                  float deriv_scale, float* grad, double* tot, double* score, int total_frames, cudaStream_t st) {
    PK2_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "pk2_latfb: workspace must be 16-byte aligned");
    LatWs w;
    lat_ws_layout(S, A, MPE, ws, &w);
    const int nb = (int)((S + 255) / 256);
    lat_arc_like_kernel<<<nb, 256, 0, st>>>(*lat, acc_in, acc_out, loglikes, num_pdfs, row_stride_b, lm, ac, w.rec_in, w.rec_out,
                                           w.pdf_out, (int)S);
    PK2_POST_LAUNCH();
    const size_t smem = chain_smem_bytes(max_frames, MPE);
    PK2_REQUIRE(smem <= 227 * 1024, "pk2_latfb: utterances of %d frames exceed the chain kernel's shared memory", max_frames);
    PK2_CHECK(cudaFuncSetAttribute(lat_chain_kernel<MPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lat_chain_kernel<MPE><<<2 * lat->n_seq, kChain, smem, st>>>(*lat, w, lm, tot, score, max_frames);
    PK2_POST_LAUNCH();
    const int64_t np = S > total_frames ? S : total_frames;
    lat_post_kernel<MPE><<<(int)((np + 255) / 256), 256, 0, st>>>(*lat, w, num_pdfs, row_stride_b, deriv_scale, tot, score,
                                                                grad, (int)S, total_frames);
    PK2_POST_LAUNCH();
    return 0;
}

}  // namespace

extern "C" size_t pk2_latfb_workspace_bytes(int64_t total_states, int64_t total_arcs, int mpe) {
    if (total_states <= 0 || total_arcs < 0) return 0;
    return lat_ws_layout(total_states, total_arcs, mpe != 0, nullptr, nullptr);
}

extern "C" int pk2_latfb_mmi(const pk2_lat_batch* lat, const float* loglikes, int num_pdfs,
                             int max_frames, int64_t row_stride_b, float lm_scale, float ac_scale,
                             void* ws, int64_t total_states, int64_t total_arcs, int64_t total_frames,
                             float* grad, double* tot, void* stream) {
    PK2_REQUIRE(lat && loglikes && ws && grad && tot, "pk2_latfb_mmi: null argument");
    PK2_REQUIRE(lat->n_seq > 0 && max_frames > 0 && total_states > 0, "pk2_latfb_mmi: empty batch");
    PK2_REQUIRE(row_stride_b >= max_frames, "pk2_latfb_mmi: row_stride_b < max_frames");
    cudaStream_t st = pk2::as_stream(stream);
    // rows are addressed as b*row_stride_b + t: zero everything up to the last sequence's max_frames
    const size_t rows = (size_t)(lat->n_seq - 1) * (size_t)row_stride_b + (size_t)max_frames;
    PK2_CHECK(cudaMemsetAsync(grad, 0, rows * (size_t)num_pdfs * sizeof(float), st));
    return launch_lat_v1<false>(lat, nullptr, nullptr, loglikes, num_pdfs, max_frames, row_stride_b, lm_scale, ac_scale, ws,
                                total_states, total_arcs, 1.0f, grad, tot, nullptr, (int)total_frames, st);
}

extern "C" int pk2_latfb_mpe(const pk2_lat_batch* lat, const uint8_t* acc_in, const uint8_t* acc_out,
                             const float* loglikes, int num_pdfs, int max_frames, int64_t row_stride_b,
                             float lm_scale, float ac_scale, void* ws, int64_t total_states, int64_t total_arcs,
                             float deriv_scale, float* grad, double* tot_like, double* tot_score, void* stream) {
    PK2_REQUIRE(lat && acc_in && acc_out && loglikes && ws && grad && tot_like && tot_score, "pk2_latfb_mpe: null argument");
    PK2_REQUIRE(lat->n_seq > 0 && max_frames > 0 && total_states > 0, "pk2_latfb_mpe: empty batch");
    PK2_REQUIRE(row_stride_b >= max_frames, "pk2_latfb_mpe: row_stride_b < max_frames");
    cudaStream_t st = pk2::as_stream(stream);
    const size_t rows = (size_t)(lat->n_seq - 1) * (size_t)row_stride_b + (size_t)max_frames;
    PK2_CHECK(cudaMemsetAsync(grad, 0, rows * (size_t)num_pdfs * sizeof(float), st));
    return launch_lat_v1<true>(lat, acc_in, acc_out, loglikes, num_pdfs, max_frames, row_stride_b, lm_scale, ac_scale, ws,
                               total_states, total_arcs, deriv_scale, grad, tot_like, tot_score, 0, st);
}
