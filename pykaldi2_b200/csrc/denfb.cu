// LF-MMI denominator forward-backward on the GPU (sm_100a).
//
// Replaces kaldi_chain.DenominatorGraph (reference bin/train_chain.py:167,202) and the
// denominator half of kaldi_chain.compute_chain_objf_and_deriv (ops/ops.py:265), i.e.
// Kaldi's chain-denominator.cc / chain-kernels.cu (one launch per frame, one thread per
// block when called with one sequence, as the reference does).  Recursion: SURVEY.md
// Appendix C (scaled-probability leaky-HMM, fp32, per-frame 1/sum(alpha) scaling).
//
// Design: one thread-block CLUSTER (K = 1, 2 or 4 CTAs) per sequence, persistent over all
// frames, no per-frame launches and no grid-wide sync.  alpha'(t) / beta(t) (S floats) and
// exp(loglikes[t]) (N floats) live in shared memory, replicated in every CTA of the
// cluster; each CTA owns a contiguous range of rows, computes them, and broadcasts the
// results into its peers' shared memory (DSMEM) followed by one cluster barrier per frame.
// The arc table is stored in SELL-32 (sliced ELLPACK, rows sorted by degree) so that a
// warp streams 32 rows in lock step with fully coalesced 8-byte arc records
// {prob, idxA | idxB<<16}; three tables exist: rows = destination states (alpha pass),
// rows = source states (beta pass), rows = pdfs (occupancy pass, so gamma needs no atomics).
// The arc tables (~0.5 MB each) are shared by all sequences and stay L2-resident; per
// frame HBM traffic is the loglike row (cp.async prefetched), the alpha row (written in the
// forward, re-read in the backward) and the gradient row.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <vector>
#include <mutex>
#include <thread>
#include <atomic>
#include <string.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxK = 4;          // largest cluster of the streaming kernels
constexpr int kMaxParts = 8;     // largest row partition of a table (register-resident kernels: clusters of 8)

struct SellDev {
    const uint2* arcs;        // SELL arc records
    const int* slice_off;     // [n_slices + 1] offsets into arcs (multiple of 32)
    const int* slice_row;     // [n_slices * 32] row id or -1
    int part_slice[kMaxParts + 1];  // slices of part c: [part_slice[c], part_slice[c+1])
    int part_row[kMaxParts + 1];    // rows of part c
};

struct SellHost {
    uint2* arcs = nullptr;
    int* slice_off = nullptr;
    int* slice_row = nullptr;
    int part_slice[kMaxParts + 1];
    int part_row[kMaxParts + 1];
    size_t n_arcs = 0;               // arc records (with padding)
    std::vector<int> slice_len;      // host copy: arcs per lane of every slice
    void free_dev() { cudaFree(arcs); cudaFree(slice_off); cudaFree(slice_row); }
    SellDev dev() const {
        SellDev d;
        d.arcs = arcs; d.slice_off = slice_off; d.slice_row = slice_row;
        for (int i = 0; i <= kMaxParts; ++i) { d.part_slice[i] = part_slice[i]; d.part_row[i] = part_row[i]; }
        return d;
    }
};

struct Arc3 { float w; int a, b; };
inline unsigned f2u(float f) { unsigned u; memcpy(&u, &f, sizeof(u)); return u; }   // generic arc of a row: weight, gather index A, gather index B

inline int part_bound(int R, int c, int K) {
    if (c >= K) return R;
    return (int)(((int64_t)R * c / K) & ~31LL);
}


// Shared-memory wavefronts of one gather of a warp: the largest number of DISTINCT addresses that fall into
// one of the 32 banks (same address = broadcast).  idx: the 32 lanes' word indices.
inline int gather_wavefronts(const unsigned* idx) {
    unsigned char cnt[32] = {0};
    int worst = 0;
    for (int l = 0; l < 32; ++l) {
        bool dup = false;
        for (int m = 0; m < l; ++m) if (idx[m] == idx[l]) { dup = true; break; }
        if (dup) continue;
        const int c = ++cnt[idx[l] & 31];
        if (c > worst) worst = c;
    }
    return worst;
}

// Second pass over a slice after the greedy fill: hill-climb on the order of the arcs INSIDE each lane (any
// order gives the same row sums; padding records may sit anywhere) to minimise the shared-memory wavefronts
// of the two gathers of every step.  The greedy order alone leaves ~2.4 wavefronts per gather on the
// BASELINE graph (3.4 unscheduled), this pass ~2.0 -- the arc passes are bound by the shared-memory pipe
// (profiles/ncu_den_full_r1_v24.md).  Deterministic (fixed seed).
inline void refine_slice(uint2* rec, int len) {
    if (len < 2) return;
    auto step_cost = [&](int k) {
        unsigned a[32], b[32];
        for (int l = 0; l < 32; ++l) { a[l] = rec[k * 32 + l].y & 0xffffu; b[l] = rec[k * 32 + l].y >> 16; }
        return gather_wavefronts(a) + gather_wavefronts(b);
    };
    std::vector<int> cost(len);
    for (int k = 0; k < len; ++k) cost[k] = step_cost(k);
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (unsigned)(rng >> 32); };
    const int sweeps = 48;
    for (int sw = 0; sw < sweeps; ++sw) {
        for (int l = 0; l < 32; ++l) {
            for (int rep = 0; rep < 2; ++rep) {
                const int i = (int)(next() % (unsigned)len), j = (int)(next() % (unsigned)len);
                if (i == j) continue;
                uint2& x = rec[i * 32 + l];
                uint2& y = rec[j * 32 + l];
                if (x.x == y.x && x.y == y.y) continue;
                const int old = cost[i] + cost[j];
                std::swap(x, y);
                const int ci = step_cost(i), cj = step_cost(j);
                if (ci + cj <= old) { cost[i] = ci; cost[j] = cj; }
                else std::swap(x, y);
            }
        }
    }
}

// Build a SELL-32 table from per-row arc lists, partitioned into K contiguous row ranges.
int build_sell(const std::vector<std::vector<Arc3>>& rows, int K, SellHost* out) {
    const int R = (int)rows.size();
    std::vector<uint2> arcs;
    std::vector<int> slice_off(1, 0), slice_row;
    for (int i = 0; i <= kMaxParts; ++i) { out->part_slice[i] = 0; out->part_row[i] = R; }
    for (int c = 0; c < K; ++c) {
        const int r0 = part_bound(R, c, K), r1 = part_bound(R, c + 1, K);
        out->part_row[c] = r0;
        out->part_slice[c] = (int)slice_off.size() - 1;
        std::vector<int> order(r1 - r0);
        for (int i = 0; i < r1 - r0; ++i) order[i] = r0 + i;
        std::stable_sort(order.begin(), order.end(),
                         [&](int x, int y) { return rows[x].size() > rows[y].size(); });
        for (size_t s = 0; s < order.size(); s += 32) {
            size_t len = rows[order[s]].size();      // longest row of the slice (sorted desc)
            const size_t base = arcs.size();
            arcs.resize(base + len * 32, make_uint2(0u, 0u));
            // Schedule the arcs of the slice: at step k every lane issues one arc of its row.  The two
            // shared-memory gathers of a step conflict when lanes hit the same bank (index mod 32), so
            // each lane greedily takes, among its remaining arcs, the one whose banks are least used
            // in this step (summation order inside a row is free).
            std::vector<std::vector<Arc3>> rem(32);
            for (int l = 0; l < 32; ++l) {
                if (s + l < order.size()) {
                    slice_row.push_back(order[s + l]);
                    rem[l] = rows[order[s + l]];
                } else {
                    slice_row.push_back(-1);
                }
            }
            for (size_t k = 0; k < len; ++k) {
                int ca[32] = {0}, cb[32] = {0};
                for (int l0 = 0; l0 < 32; ++l0) {
                    const int l = (int)((l0 + k) & 31);          // rotate the lane that chooses first
                    if (rem[l].empty()) continue;
                    size_t best = 0;
                    int bc = 1 << 30;
                    for (size_t q = 0; q < rem[l].size(); ++q) {
                        const int c = ca[rem[l][q].a & 31] + cb[rem[l][q].b & 31];
                        if (c < bc) { bc = c; best = q; }
                    }
                    const Arc3 a = rem[l][best];
                    rem[l][best] = rem[l].back();
                    rem[l].pop_back();
                    ca[a.a & 31]++; cb[a.b & 31]++;
                    uint2 rec;
                    rec.x = f2u(a.w);
                    rec.y = (unsigned)a.a | ((unsigned)a.b << 16);
                    arcs[base + k * 32 + l] = rec;
                }
            }
            slice_off.push_back((int)arcs.size());
        }
    }
    {   // second pass (independent per slice): host threads
        const int nsl = (int)slice_off.size() - 1;
        const int nth = std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        std::atomic<int> nextsl(0);
        uint2* data = arcs.data();
        for (int t = 0; t < nth; ++t)
            pool.emplace_back([&]() {
                for (int i = nextsl.fetch_add(1); i < nsl; i = nextsl.fetch_add(1))
                    refine_slice(data + slice_off[i], (slice_off[i + 1] - slice_off[i]) / 32);
            });
        for (auto& th : pool) th.join();
    }
    for (int c = K; c <= kMaxParts; ++c) { out->part_slice[c] = (int)slice_off.size() - 1; out->part_row[c] = R; }
    out->n_arcs = arcs.size();
    out->slice_len.clear();
    for (size_t i = 0; i + 1 < slice_off.size(); ++i) out->slice_len.push_back((slice_off[i + 1] - slice_off[i]) / 32);
    if (arcs.empty()) arcs.push_back(make_uint2(0u, 0u));
    if (slice_row.empty()) slice_row.push_back(-1);
    PK2_CHECK(cudaMalloc(&out->arcs, sizeof(uint2) * arcs.size()));
    PK2_CHECK(cudaMalloc(&out->slice_off, sizeof(int) * slice_off.size()));
    PK2_CHECK(cudaMalloc(&out->slice_row, sizeof(int) * slice_row.size()));
    PK2_CHECK(cudaMemcpy(out->arcs, arcs.data(), sizeof(uint2) * arcs.size(), cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(out->slice_off, slice_off.data(), sizeof(int) * slice_off.size(), cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(out->slice_row, slice_row.data(), sizeof(int) * slice_row.size(), cudaMemcpyHostToDevice));
    return 0;
}

struct RegSmemHost;
struct DenGraph {
    int S = 0, N = 0;
    int64_t A = 0;
    std::vector<std::vector<Arc3>> rows_fwd, rows_bwd, rows_pdf;
    float* init = nullptr;          // device [S]
    float init_sum = 0.f;
    bool built[kMaxParts + 1] = {};
    SellHost t_fwd[kMaxParts + 1], t_bwd[kMaxParts + 1], t_pdf[kMaxParts + 1];
    std::mutex mu;
    std::mutex launch_mu;            // the side streams / events / start flag of a graph serve one call at a time
    int* start_flag = nullptr;       // device int, hybrid schedule
    int epoch = 0;
    int reg_state = 0;               // register-resident path: 0 = not planned, 1 = usable, -1 = not usable
    RegSmemHost* reg = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};     // side streams for the mixed-cluster schedule
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    int budget_clusters = 0;         // pk2_den_set_sm_budget: cap on resident clusters of 8 (0 = all that fit)
    int budget_reserve = 0;          //                        SMs left to kernels of other streams
};

// ---------------------------------------------------------------- device helpers ----

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// asynchronous copy of n floats gmem -> smem (16 B chunks when aligned, else plain loads)
__device__ __forceinline__ void row_prefetch(float* dst, const float* src, int n) {
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        for (int i = threadIdx.x * 4; i < n; i += kThreads * 4) cp_async16(dst + i, src + i);
    } else {
        for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = __ldg(src + i);
    }
    cp_async_commit();
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    // deterministic: warp shuffle tree, then warp 0 adds the 32 partials in order
    v = pk2::warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < kWarps) ? red[threadIdx.x] : 0.f;
    if (warp == 0) {
        t = pk2::warp_sum(t);
        if (lane == 0) red[kWarps] = t;
    }
    __syncthreads();
    return red[kWarps];
}

template <int K>
__device__ __forceinline__ void cluster_barrier() {
    if constexpr (K == 1) __syncthreads();
    else cg::this_cluster().sync();
}

// Per-CTA view of one SELL table, staged once into shared memory: slice (offset, length) and the row
// ids as uint16 (0xFFFF = padding row).  Re-reading this metadata from L2 every frame (L1 is flushed by
// every cluster barrier) was the largest fixed per-frame cost of the first version.
struct TabView {
    const uint2* arcs;
    const int2* meta;        // [nsl] {arc offset, slice length}
    const uint16_t* rows;    // [nsl * 32]
    int nsl;
};

__device__ __forceinline__ void stage_table(const SellDev& tb, int part, int2* meta, uint16_t* rows) {
    const int s0 = tb.part_slice[part], n = tb.part_slice[part + 1] - s0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const int off = __ldg(&tb.slice_off[s0 + i]);
        meta[i] = make_int2(off, (__ldg(&tb.slice_off[s0 + i + 1]) - off) >> 5);
    }
    for (int i = threadIdx.x; i < n * 32; i += kThreads) {
        const int r = __ldg(&tb.slice_row[s0 * 32 + i]);
        rows[i] = r < 0 ? (uint16_t)0xFFFFu : (uint16_t)r;
    }
}

// Sum over the arcs of every row of this CTA's part; body(row, acc) per valid row.  A warp streams TWO
// slices at a time (8 independent 8-byte loads in flight per lane) to cover the L2 latency of the arc
// records; the two shared-memory gathers per arc are scheduled bank-aware at build time.
// MODE 0: acc  = sum ga[a] * (w * gb[b])
// MODE 1: also acc2 = sum w * gb[b]          (beta pass with the leaky term applied lazily)
// MODE 2: also acc2 = sum ga[a] * w          (occupancy pass with the leaky term applied lazily)
template <int MODE, class Body>
__device__ __forceinline__ void sell_pass(const TabView& tv, const float* __restrict__ ga,
                                          const float* __restrict__ gb, Body&& body, int debug = 0,
                                          const float* __restrict__ rowvec = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (debug == 2) return;
    for (int ia = warp; ia < tv.nsl; ia += 2 * kWarps) {
        const int ib = ia + kWarps;
        const bool hasb = ib < tv.nsl;
        const int2 ma = tv.meta[ia];
        const int2 mb = hasb ? tv.meta[ib] : make_int2(0, 0);
        const uint2* pa = tv.arcs + ma.x + lane;
        const uint2* pb = tv.arcs + mb.x + lane;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, a2 = 0.f, b2 = 0.f;
        // row ids are known up front: fetch the per-row value of `rowvec` now, its L2 latency hides
        // behind the arc loop
        const unsigned rowa = tv.rows[ia * 32 + lane];
        const unsigned rowb = hasb ? tv.rows[ib * 32 + lane] : 0xFFFFu;
        float va = 0.f, vb = 0.f;
        if (rowvec) {
            if (rowa != 0xFFFFu) va = __ldg(&rowvec[rowa]);
            if (rowb != 0xFFFFu) vb = __ldg(&rowvec[rowb]);
        }
        const int lmax = debug == 1 ? 0 : max(ma.y, mb.y);
        for (int k = 0; k < lmax; k += 4) {
            uint2 ra[4], rb[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ra[q] = (k + q < ma.y) ? __ldg(pa + (k + q) * 32) : make_uint2(0u, 0u);
                rb[q] = (k + q < mb.y) ? __ldg(pb + (k + q) * 32) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float wa = __uint_as_float(ra[q].x), wb = __uint_as_float(rb[q].x);
                const float ga_a = ga[ra[q].y & 0xffffu], ga_b = ga[rb[q].y & 0xffffu];
                const float ta = wa * gb[ra[q].y >> 16], tb = wb * gb[rb[q].y >> 16];
                if (q & 1) { a1 = fmaf(ga_a, ta, a1); b1 = fmaf(ga_b, tb, b1); }
                else       { a0 = fmaf(ga_a, ta, a0); b0 = fmaf(ga_b, tb, b0); }
                if (MODE == 1) { a2 += ta; b2 += tb; }
                if (MODE == 2) { a2 = fmaf(ga_a, wa, a2); b2 = fmaf(ga_b, wb, b2); }
            }
        }
        if (rowa != 0xFFFFu) body((int)rowa, a0 + a1, a2, va);
        if (rowb != 0xFFFFu) body((int)rowb, b0 + b1, b2, vb);
    }
}

struct DenArgs {
    SellDev fwd, bwd, pdf;
    const float* init;
    int S, N;
    const float* ll;
    const int32_t* num_frames;
    int64_t row_stride_b;
    int max_frames;
    float leaky, deriv_scale;
    float* alpha_ws;     // [n_seq][max_frames][S]  alpha'(t), t < T
    float* asum_ws;      // [n_seq][max_frames + 2] A(t), t <= T ; slot max_frames+1 = totp
    float* grad;
    double* logz;
    const int32_t* seq_map;   // sequence handled by cluster i (NULL = identity)
    const float* e;           // register-resident kernels: exp(clamp(loglikes)) [n_seq][max_frames][N]
    const int32_t* work;      // register-resident kernels: [n_clusters + 1] offsets, then sequence ids
    int work_ids;             //   offset of the ids inside `work`
    int* start_flag;          // hybrid schedule: set to `epoch` when the forward cluster kernel has started
    int epoch;
    int debug;                // profiling only (PK2_DEN_DEBUG): 1 = skip the arc loops, 2 = skip the passes
    long long* prof;          // profiling only: clock64 stamps of frames 64..71 of cluster 0 (pk2_den_set_profile_buffer)
};

constexpr int kInitRegs = 8;     // init[] values a thread keeps in registers (covers S <= 8192)

// End of a pass: every CTA of the cluster needs (a) the rows the other CTAs computed and (b) the sum
// of one scalar per CTA.  Each CTA has written its own rows [r0, r1) into its LOCAL copy of `vec`;
// one thread then pushes that contiguous row block (and a 16-byte slot with the CTA's partial sum)
// into every peer's shared memory with a DSMEM bulk copy whose completion bytes are counted on the
// PEER's mbarrier.  No scattered remote stores, no cluster barrier on the per-frame path.
// Returns the cluster-wide sum (added in a fixed order: bit-identical in all CTAs).
struct XchgState {
    uint64_t* bar;       // [2] one mbarrier per frame parity
    float* parts;        // [2][kMaxK * 4] partial sums (16-byte slots)
    float* wsum;         // [kWarps]
    uint32_t phase;      // bit par = phase parity of bar[par]
    uint32_t expect;     // bytes received from the peers per frame
};

template <int K>
__device__ __forceinline__ float cluster_exchange(XchgState& x, float* vec, const float* src_rows, int r0, int r1,
                                                  float warp_partial, int par, int c) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) x.wsum[warp] = warp_partial;
    if constexpr (K > 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my row writes -> async proxy
    __syncthreads();
    float* parts = x.parts + par * kMaxK * 4;
    if (warp == 0) {
        float t = pk2::warp_sum(x.wsum[lane]);
        if (lane == 0) parts[c * 4] = t;
        if constexpr (K > 1) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const uint32_t rows_l = tc::smem_u32(vec + r0), src_l = tc::smem_u32(src_rows);
                const uint32_t part_l = tc::smem_u32(parts + c * 4);
                const uint32_t bar_l = tc::smem_u32(&x.bar[par]);
#pragma unroll
                for (int q = 1; q < K; ++q) {
                    const uint32_t peer = (uint32_t)((c + q) % K);
                    const uint32_t bar_r = tc::mapa_u32(bar_l, peer);
                    tc::dsmem_bulk_copy(tc::mapa_u32(rows_l, peer), src_l, (uint32_t)(r1 - r0) * 4u, bar_r);
                    tc::dsmem_bulk_copy(tc::mapa_u32(part_l, peer), part_l, 16u, bar_r);
                }
            }
        }
    }
    if constexpr (K > 1) {
        tc::mbar_wait(&x.bar[par], (x.phase >> par) & 1u);
        x.phase ^= (1u << par);
        __syncthreads();                       // everyone has seen this phase before it is re-armed
        if (threadIdx.x == 0) tc::mbar_expect_tx(&x.bar[par], x.expect);   // arm for frame t+2
    } else {
        __syncthreads();
    }
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < K; ++q) tot += parts[q * 4];
    return tot;
}

// -------------------------------------------------------------------- forward ----
template <int K>
__device__ __forceinline__ void den_forward_body(const DenArgs& a) {
    extern __shared__ __align__(16) float smem[];
    const int S = a.S, N = a.N;
    const int Sp = (S + 3) & ~3, Np = (N + 3) & ~3;
    const int c = (K == 1) ? 0 : (int)cg::this_cluster().block_rank();
    const int nsl = a.fwd.part_slice[c + 1] - a.fwd.part_slice[c];
    float* buf0 = smem;
    float* buf1 = buf0 + Sp;
    float* ev = buf1 + Sp;          // exp(loglikes[t])
    float* lraw = ev + Np;          // prefetched raw loglikes[t+1]
    float* wsum = lraw + Np;        // [kWarps] + parts [2][kMaxK*4]
    float* red = wsum + 2 * kMaxK * kWarps;   // [kWarps + 1]
    uint64_t* xbar = reinterpret_cast<uint64_t*>(red + kWarps + 4);          // [2] exchange mbarriers
    int2* meta = reinterpret_cast<int2*>(xbar + 2);
    uint16_t* rows = reinterpret_cast<uint16_t*>(meta + ((S + 31) / 32 + 2));
    const int rcap = ((S + K - 1) / K + 64 + 3) & ~3;                         // rows per CTA (padded)
    // outgoing copy of this CTA's rows, double buffered: alpha(t+1) is updated in place (leaky term)
    // while the bulk copy of the raw rows may still be reading its source
    float* stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(rows + rcap + 32) + 15) & ~(uintptr_t)15);

    const int b = a.seq_map ? a.seq_map[blockIdx.x / K] : (int)(blockIdx.x / K);
    const int T = a.num_frames[b];
    const float* ll = a.ll + (int64_t)b * a.row_stride_b * N;
    float* aws = a.alpha_ws + (int64_t)b * a.max_frames * S;
    float* asum = a.asum_ws + (int64_t)b * (a.max_frames + 2);
    const int r0 = a.fwd.part_row[c], r1 = a.fwd.part_row[c + 1];

    stage_table(a.fwd, c, meta, rows);
    TabView tv;
    tv.arcs = a.fwd.arcs; tv.meta = meta; tv.rows = rows; tv.nsl = nsl;
    XchgState xs;
    xs.bar = xbar; xs.wsum = wsum; xs.parts = wsum + kWarps; xs.phase = 0;
    xs.expect = (uint32_t)(S - (r1 - r0)) * 4u + (uint32_t)(K - 1) * 16u;
    if (K > 1 && threadIdx.x == 0) {
        tc::mbar_init(&xbar[0], 1); tc::mbar_init(&xbar[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&xbar[0], xs.expect);
        tc::mbar_expect_tx(&xbar[1], xs.expect);
    }

    float* cur = buf0;
    float* nxt = buf1;

    // alpha(0) = init, A(0) = sum(init), alpha'(0) = alpha(0) + leaky*A(0)*init ; init kept in registers
    float rinit[kInitRegs];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < kInitRegs; ++q) {
        const int j = threadIdx.x + q * kThreads;
        rinit[q] = (j < S) ? __ldg(&a.init[j]) : 0.f;
        s += rinit[q];
    }
    for (int j = threadIdx.x + kInitRegs * kThreads; j < S; j += kThreads) s += __ldg(&a.init[j]);
    float A = block_sum(s, red);
#pragma unroll
    for (int q = 0; q < kInitRegs; ++q) {
        const int j = threadIdx.x + q * kThreads;
        if (j < S) cur[j] = rinit[q] + a.leaky * A * rinit[q];
    }
    for (int j = threadIdx.x + kInitRegs * kThreads; j < S; j += kThreads) {
        const float v = __ldg(&a.init[j]);
        cur[j] = v + a.leaky * A * v;
    }
    if (T > 0) row_prefetch(lraw, ll, N);
    double logsum = 0.0;
    cp_async_wait_all();
    __syncthreads();
    for (int p = threadIdx.x; p < N; p += kThreads) ev[p] = __expf(fminf(fmaxf(lraw[p], -30.f), 30.f));
    cluster_barrier<K>();           // peers' shared memory is live from here on; ev / cur complete
    if (T > 1) row_prefetch(lraw, ll + (int64_t)N, N);

    long long* prof = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? a.prof : nullptr;
#define PK2_DPROF(e) do { if (prof && t >= 64 && t < 72) prof[(t - 64) * 8 + (e)] = clock64(); } while (0)
    for (int t = 0; t < T; ++t) {
        PK2_DPROF(0);
        // cur = alpha'(t) (complete in every CTA), ev = e(t), lraw <- loglikes[t+1] in flight
        for (int j = r0 + threadIdx.x; j < r1; j += kThreads) aws[(int64_t)t * S + j] = cur[j];
        if (threadIdx.x == 0 && c == 0) asum[t] = A;

        const float invA = 1.0f / A;
        float local = 0.f;
        const int par = t & 1;
        float* stg = stage + par * rcap;
        sell_pass<0>(tv, cur, ev, [&](int row, float acc, float, float) {
            const float v = acc * invA;
            local += v;
            nxt[row] = v;
            if (K > 1) stg[row - r0] = v;
        }, a.debug);
        PK2_DPROF(1);
        const float An = cluster_exchange<K>(xs, nxt, stg, r0, r1, pk2::warp_sum(local), par, c);
        PK2_DPROF(2);
        // alpha'(t+1) in place, e(t+1) from the prefetched row
        const float lk = a.leaky * An;
#pragma unroll
        for (int q = 0; q < kInitRegs; ++q) {
            const int j = threadIdx.x + q * kThreads;
            if (j < S) nxt[j] = fmaf(lk, rinit[q], nxt[j]);
        }
        for (int j = threadIdx.x + kInitRegs * kThreads; j < S; j += kThreads)
            nxt[j] = fmaf(lk, __ldg(&a.init[j]), nxt[j]);
        PK2_DPROF(3);
        if (t + 1 < T) {
            cp_async_wait_all();
            __syncthreads();                          // lraw visible to all threads; nxt update done
            PK2_DPROF(4);
            for (int p = threadIdx.x; p < N; p += kThreads) ev[p] = __expf(fminf(fmaxf(lraw[p], -30.f), 30.f));
        }
        A = An;
        float* tmp = cur; cur = nxt; nxt = tmp;
        PK2_DPROF(5);
        __syncthreads();
        if (t + 2 < T) row_prefetch(lraw, ll + (int64_t)(t + 2) * N, N);
        PK2_DPROF(6);
    }
#undef PK2_DPROF
    // total probability: sum_j alpha'(T, j)
    float s2 = 0.f;
    for (int j = threadIdx.x; j < S; j += kThreads) s2 += cur[j];
    const float totp = block_sum(s2, red);
    if (c == 0) {
        // log Z = log totp + sum_t log A(t): the per-frame normalisers were stored by thread 0; sum their logs
        // here in double, in parallel, instead of one double log per frame on the critical path
        __threadfence_block();
        __syncthreads();
        double part = 0.0;
        for (int t = threadIdx.x; t < T; t += kThreads) part += log((double)asum[t]);
        part = pk2::warp_sum_d(part);
        double* dred = reinterpret_cast<double*>(ev);         // ev is dead now
        if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 0; w < kWarps; ++w) logsum += dred[w];
            asum[T] = A;
            asum[a.max_frames + 1] = totp;
            a.logz[b] = log((double)totp) + logsum;
        }
    }
    cluster_barrier<K>();           // no CTA exits while peers may still write into it
}

// ------------------------------------------------------------------- backward ----
template <int K>
__global__ void __launch_bounds__(kThreads, 1) den_forward_kernel(DenArgs a) { den_forward_body<K>(a); }

template <int K>
__device__ __forceinline__ void den_backward_body(const DenArgs& a) {
    extern __shared__ __align__(16) float smem[];
    const int S = a.S, N = a.N;
    const int Sp = (S + 3) & ~3, Np = (N + 3) & ~3;
    const int c = (K == 1) ? 0 : (int)cg::this_cluster().block_rank();
    float* buf0 = smem;
    float* buf1 = buf0 + Sp;
    float* al = buf1 + Sp;          // alpha'(t)
    float* ev = al + Sp;            // exp(loglikes[t])
    float* lraw = ev + Np;          // prefetched raw loglikes
    float* gbuf = lraw + Np;        // gamma staging for this CTA's pdf range
    const int gcap = ((N + K - 1) / K + 64 + 3) & ~3;
    // second alpha buffer (K > 1, where it fits): alpha'(t-1) is prefetched a whole frame ahead
    constexpr bool kAl2 = (K > 1);
    float* al2 = gbuf + gcap;
    float* wsum = al2 + (kAl2 ? Sp : 0);       // [kWarps] + parts [2][kMaxK*4]
    float* red = wsum + 2 * kMaxK * kWarps;
    uint64_t* xbar = reinterpret_cast<uint64_t*>(red + kWarps + 4);
    int2* meta_b = reinterpret_cast<int2*>(xbar + 2);
    int2* meta_g = meta_b + ((S + 31) / 32 + 2);
    uint16_t* rows_b = reinterpret_cast<uint16_t*>(meta_g + ((N + 31) / 32 + 2));
    const int nsl_b = a.bwd.part_slice[c + 1] - a.bwd.part_slice[c];
    const int nsl_g = a.pdf.part_slice[c + 1] - a.pdf.part_slice[c];
    uint16_t* rows_g = rows_b + nsl_b * 32;

    const int b = a.seq_map ? a.seq_map[blockIdx.x / K] : (int)(blockIdx.x / K);
    const int T = a.num_frames[b];
    const float* ll = a.ll + (int64_t)b * a.row_stride_b * N;
    float* grad = a.grad + (int64_t)b * a.row_stride_b * N;
    const float* aws = a.alpha_ws + (int64_t)b * a.max_frames * S;
    const float* asum = a.asum_ws + (int64_t)b * (a.max_frames + 2);
    const int p0 = a.pdf.part_row[c], p1 = a.pdf.part_row[c + 1];

    // zero-fill the padded frames of this CTA's pdf range
    for (int t = T; t < a.max_frames; ++t)
        for (int p = p0 + threadIdx.x; p < p1; p += kThreads) grad[(int64_t)t * N + p] = 0.f;
    if (T <= 0) return;   // uniform across the cluster (same sequence)

    stage_table(a.bwd, c, meta_b, rows_b);
    stage_table(a.pdf, c, meta_g, rows_g);
    TabView tb, tg;
    tb.arcs = a.bwd.arcs; tb.meta = meta_b; tb.rows = rows_b; tb.nsl = nsl_b;
    tg.arcs = a.pdf.arcs; tg.meta = meta_g; tg.rows = rows_g; tg.nsl = nsl_g;
    const int r0 = a.bwd.part_row[c], r1 = a.bwd.part_row[c + 1];
    XchgState xs;
    xs.bar = xbar; xs.wsum = wsum; xs.parts = wsum + kWarps; xs.phase = 0;
    xs.expect = (uint32_t)(S - (r1 - r0)) * 4u + (uint32_t)(K - 1) * 16u;
    if (K > 1 && threadIdx.x == 0) {
        tc::mbar_init(&xbar[0], 1); tc::mbar_init(&xbar[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&xbar[0], xs.expect);
        tc::mbar_expect_tx(&xbar[1], xs.expect);
    }

    float* cur = buf0;
    float* nxt = buf1;

    // beta'(T) = 1/totp ; beta(T) = beta'(T) + leaky * sum_k beta'(T,k) init[k]
    float s = 0.f;
    for (int j = threadIdx.x; j < S; j += kThreads) s += __ldg(&a.init[j]);
    const float isum = block_sum(s, red);
    const float totp = asum[a.max_frames + 1];
    // cur holds beta'(t+1) WITHOUT the leaky term; beta = beta' + lk with the scalar lk = leaky * sum_k beta'_k init_k.
    // (beta' rows are exchanged by asynchronous bulk copies, so they are never modified in place.)
    const float bT = 1.0f / totp;
    float lk = a.leaky * isum * bT;
    for (int j = threadIdx.x; j < S; j += kThreads) cur[j] = bT;
    row_prefetch(lraw, ll + (int64_t)(T - 1) * N, N);
    row_prefetch(al, aws + (int64_t)(T - 1) * S, S);
    cluster_barrier<K>();
    float* alc = al;                 // alpha'(t)
    float* aln = kAl2 ? al2 : al;    // alpha'(t-1) being prefetched
    float asum_t = asum[T - 1];

    long long* prof = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? a.prof + 64 : nullptr;
#define PK2_DPROF(e) do { if (prof && t >= 64 && t < 72) prof[(t - 64) * 8 + (e)] = clock64(); } while (0)
    for (int t = T - 1; t >= 0; --t) {
        PK2_DPROF(0);
        cp_async_wait_all();
        __syncthreads();
        PK2_DPROF(1);
        for (int p = threadIdx.x; p < N; p += kThreads)
            ev[p] = __expf(fminf(fmaxf(lraw[p], -30.f), 30.f));
        const float invA = 1.0f / asum_t;
        if (t > 0) asum_t = asum[t - 1];                 // next frame's normaliser: load now, use next frame
        __syncthreads();
        if (t > 0) {
            row_prefetch(lraw, ll + (int64_t)(t - 1) * N, N);
            if (kAl2) row_prefetch(aln, aws + (int64_t)(t - 1) * S, S);
        }
        PK2_DPROF(2);

        // beta'(t, i) for this CTA's source states
        float local = 0.f;
        sell_pass<1>(tb, cur, ev, [&](int row, float acc, float acc2, float init_row) {
            const float v = fmaf(lk, acc2, acc) * invA;
            local = fmaf(v, init_row, local);
            nxt[row] = v;
        }, a.debug, a.init);
        PK2_DPROF(3);
        // pdf occupancies gamma(t, p) for this CTA's pdf range
        const float gs = a.deriv_scale * invA;
        sell_pass<2>(tg, alc, cur, [&](int row, float acc, float acc2, float) {
            gbuf[row - p0] = fmaf(lk, acc2, acc) * ev[row] * gs;
        }, a.debug);
        PK2_DPROF(4);
        const int par = t & 1;
        const float dot = cluster_exchange<K>(xs, nxt, nxt + r0, r0, r1, pk2::warp_sum(local), par, c);
        PK2_DPROF(5);
        // (block barrier passed inside: gbuf / al reads of this frame are complete in this CTA)
        for (int p = p0 + threadIdx.x; p < p1; p += kThreads) grad[(int64_t)t * N + p] = gbuf[p - p0];
        if (kAl2) { float* t2 = alc; alc = aln; aln = t2; }
        else if (t > 0) row_prefetch(al, aws + (int64_t)(t - 1) * S, S);
        lk = a.leaky * dot;
        float* tmp = cur; cur = nxt; nxt = tmp;
        PK2_DPROF(6);
    }
#undef PK2_DPROF
    cluster_barrier<K>();
}

template <int K>
__global__ void __launch_bounds__(kThreads, 1) den_backward_kernel(DenArgs a) { den_backward_body<K>(a); }

// Forward and backward of single-CTA sequences in ONE launch (hybrid schedule: the shortest sequences of a
// batch run on the SMs the clusters of 8 leave free; a single launch is placed once, so its CTAs can never
// land on the SMs a cluster kernel is about to need).
__global__ void __launch_bounds__(kThreads, 1) den_fb1_kernel(DenArgs a) {
    den_forward_body<1>(a);
    __threadfence();
    __syncthreads();
    den_backward_body<1>(a);
}

// Side-stream gate: returns once the cluster kernel of this call has started (its CTAs are placed), so that
// the single-CTA kernel behind it only gets SMs the clusters do not use.  Bounded wait.
__global__ void den_gate_kernel(const int* flag, int epoch) {
    for (int i = 0; i < (1 << 20); ++i) {
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v - epoch >= 0) return;
        __nanosleep(200);
    }
}

// =====================================================================================================
// Register-resident variant (clusters of 8 CTAs x 512 threads).
//
// The streaming kernels above re-read every arc record from L2 once per frame (3 x 0.6 MB per sequence and
// frame): at 4 CTAs per sequence the three arc passes still take 21 k of the 32 k cycles of a frame pair
// (profiles/den_trace_r1_v19.txt).  With 8 CTAs per sequence a CTA owns S/8 <= 1024 rows, i.e. two rows per
// thread, and the ~8 arcs of a row fit the register file: the alpha and beta passes read no arc record from
// memory at all (kRegArcs records per row live in registers for the whole kernel, longer rows keep the
// remainder in shared memory), and the pdf-occupancy table of the CTA (~65 KB) is copied into shared memory
// once.  What remains per frame is the two shared-memory gathers per arc (shared-memory bandwidth is the
// bound of these kernels), one row exchange over DSMEM and two block barriers:
//   * forward: the leaky-HMM term is applied lazily.  Peers exchange alpha(t) WITHOUT the leaky term and
//     every row adds leaky*A(t) * z(t,j), z(t,j) = sum_arcs w*init[src]*e(t,pdf): one more register and FMA
//     per arc instead of a pass over all S states in every CTA.
//   * backward: the gamma pass is taken off the dependency chain: it runs while the beta rows travel.
//   * the copies to the 7 peers are issued by 7 different warps (a single thread issuing 14 bulk copies
//     costs ~2 k cycles).
// A cluster is persistent over a LIST of sequences (host-side LPT assignment by length), so the 64
// sequences of a batch are processed by the ~16 clusters that fit the GPU in one wave.
constexpr int kRK = 8;             // CTAs per cluster
constexpr int kRT = 512;           // threads per CTA
constexpr int kRW = kRT / 32;      // warps per CTA
constexpr int kRegArcs = 12;       // arc records per row held in registers
constexpr int kRowsPerThread = 2;  // SELL slices per warp

struct RegSlice {                  // one warp's slice of a SELL table
    float w[kRegArcs];             // arc probability
    uint32_t off[kRegArcs];        // byte offsets of the two gathers: (a * 4) | (b * 4) << 16
    int len;                       // arcs per lane in this slice (warp-uniform); 0 = no slice
    int row;                       // row of this lane, -1 = padding
    const uint4* ovf;              // shared memory: records beyond kRegArcs, [len - kRegArcs][32] of {w, wi, off, -}
};

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
// keep a loop-invariant address in a register: the compiler otherwise re-derives it (8-12 instructions) at every use
__device__ __forceinline__ uint32_t pin_u32(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t prescale(uint32_t y) { return ((y & 0xffffu) << 2) | ((y >> 16) << 18); }

// Load slices (warp, warp + kRW) of part c into registers; wi[] = w * init[a] when WI (forward table: a = source state).
template <bool WI>
__device__ __forceinline__ void load_reg_slices(const SellDev& tb, int c, const float* __restrict__ init,
                                                RegSlice (&sl)[kRowsPerThread], float (&wi)[kRowsPerThread][kRegArcs],
                                                uint4* ovf_area, int* wcnt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = tb.part_slice[c], nsl = tb.part_slice[c + 1] - s0;
    int off[kRowsPerThread], extra[kRowsPerThread];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
        const int si = r == 0 ? warp : 2 * kRW - 1 - warp;       // slices are sorted by length: pair long with short
        sl[r].len = 0; sl[r].row = -1; off[r] = 0;
        if (si < nsl) {
            off[r] = __ldg(&tb.slice_off[s0 + si]);
            sl[r].len = (__ldg(&tb.slice_off[s0 + si + 1]) - off[r]) >> 5;
            sl[r].row = __ldg(&tb.slice_row[(s0 + si) * 32 + lane]);
        }
#pragma unroll
        for (int k = 0; k < kRegArcs; ++k) {
            const uint2 rec = (k < sl[r].len) ? __ldg(tb.arcs + off[r] + k * 32 + lane) : make_uint2(0u, 0u);
            sl[r].w[k] = __uint_as_float(rec.x);
            sl[r].off[k] = prescale(rec.y);
            wi[r][k] = WI ? sl[r].w[k] * __ldg(&init[rec.y & 0xffffu]) : 0.f;
        }
        extra[r] = max(sl[r].len - kRegArcs, 0);
    }
    if (lane == 0) { wcnt[warp] = extra[0] * 32; wcnt[2 * kRW - 1 - warp] = extra[1] * 32; }
    __syncthreads();
    // overflow area order: slice 0..2*kRW-1 (wcnt index = slice index inside the part)
    const int c0 = wcnt[lane];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) {
        const int si = r == 0 ? warp : 2 * kRW - 1 - warp;
        const int pre = warp_sum_i(lane < si ? c0 : 0);
        for (int k = 0; k < extra[r]; ++k) {
            const uint2 rec = __ldg(tb.arcs + off[r] + (kRegArcs + k) * 32 + lane);
            const float w = __uint_as_float(rec.x);
            const float x = WI ? w * __ldg(&init[rec.y & 0xffffu]) : 0.f;
            ovf_area[pre + k * 32 + lane] = make_uint4(rec.x, __float_as_uint(x), prescale(rec.y), 0u);
        }
        sl[r].ovf = ovf_area + pre;
    }
    __syncwarp();
}

// One slice: acc += sum g_a * (w * g_b); accz += sum wz * g_b where wz = wi (forward: lazy leaky term) or
// w (backward: lazy leaky term of beta).  base_a / base_b: shared-memory byte addresses of the gathered vectors.
template <bool WI>
__device__ __forceinline__ void reg_slice_pass(const RegSlice& s, const float (&wi)[kRegArcs], uint32_t base_a,
                                               uint32_t base_b, float& acc, float& accz) {
    const int lane = threadIdx.x & 31;
    float a0 = 0.f, a1 = 0.f, z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int g = 0; g < kRegArcs; g += 4) {
        if (g < s.len) {
            float gb[4], ga[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                gb[u] = lds_f32(base_b + (s.off[g + u] >> 16));
                uint32_t lo = s.off[g + u] & 0xffffu;
                // forward kernels are short of registers: keep the compiler from hoisting 24 masked copies of the
                // offsets out of the frame loop (one more LOP per arc instead)
                if (WI) asm volatile("and.b32 %0, %1, 0xffff;" : "=r"(lo) : "r"(s.off[g + u]));
                ga[u] = lds_f32(base_a + lo);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float t = s.w[g + u] * gb[u];
                if (u & 1) { a1 = fmaf(ga[u], t, a1); z1 = WI ? fmaf(wi[g + u], gb[u], z1) : z1 + t; }
                else       { a0 = fmaf(ga[u], t, a0); z0 = WI ? fmaf(wi[g + u], gb[u], z0) : z0 + t; }
            }
        }
    }
    for (int k = kRegArcs; k < s.len; k += 4) {              // rows longer than the register slots (rare)
        uint4 r[4];
        float gb[4], ga[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            r[u] = (k + u < s.len) ? s.ovf[(k + u - kRegArcs) * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 4; ++u) { gb[u] = lds_f32(base_b + (r[u].z >> 16)); ga[u] = lds_f32(base_a + (r[u].z & 0xffffu)); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float t = __uint_as_float(r[u].x) * gb[u];
            if (u & 1) { a1 = fmaf(ga[u], t, a1); z1 = WI ? fmaf(__uint_as_float(r[u].y), gb[u], z1) : z1 + t; }
            else       { a0 = fmaf(ga[u], t, a0); z0 = WI ? fmaf(__uint_as_float(r[u].y), gb[u], z0) : z0 + t; }
        }
    }
    acc = a0 + a1;
    accz = z0 + z1;
}

template <int NT>
__device__ __forceinline__ void row_prefetch_nt(float* dst, const float* src, int n) {
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        for (int i = threadIdx.x * 4; i < n; i += NT * 4) cp_async16(dst + i, src + i);
    } else {
        for (int i = threadIdx.x; i < n; i += NT) dst[i] = __ldg(src + i);
    }
    cp_async_commit();
}
// exp(clamp(x)) of a row: straight from global memory (first frame of a sequence) ...
template <int NT>
__device__ __forceinline__ void exp_row_direct(float* dst, const float* src, int n) {
    for (int i = threadIdx.x; i < n; i += NT) dst[i] = __expf(fminf(fmaxf(__ldg(src + i), -30.f), 30.f));
}
// ... or in place on the chunks THIS thread fetched with row_prefetch_nt (same thread -> element mapping, so
// no barrier is needed between the copy landing and the exp)
template <int NT>
__device__ __forceinline__ void exp_inplace_own(float* dst, const float* src, int n) {
    const bool vec = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
        for (int i = threadIdx.x * 4; i < n; i += NT * 4) {
            float4 x = *reinterpret_cast<float4*>(dst + i);
            x.x = __expf(fminf(fmaxf(x.x, -30.f), 30.f)); x.y = __expf(fminf(fmaxf(x.y, -30.f), 30.f));
            x.z = __expf(fminf(fmaxf(x.z, -30.f), 30.f)); x.w = __expf(fminf(fmaxf(x.w, -30.f), 30.f));
            *reinterpret_cast<float4*>(dst + i) = x;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += NT) dst[i] = __expf(fminf(fmaxf(dst[i], -30.f), 30.f));
    }
}
template <int NW>
__device__ __forceinline__ float block_sum_nw(float v, float* red) {
    v = pk2::warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const float t = pk2::warp_sum(lane < NW ? red[lane] : 0.f);     // every warp: identical bits
    __syncthreads();
    return t;
}

// mbarrier wait with a watchdog: a protocol error traps (cudaErrorLaunchFailure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = tc::smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 27)) __trap();      // seconds: far beyond any legitimate wait (micro-seconds)
    }
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Row exchange: warp q (q = 0..6), one elected lane, pushes this CTA's row block and its 16-byte partial-sum
// slot into peer (c + 1 + q) % 8; completion is counted on the peer's mbarrier.  Every issuing thread writes
// the (identical) partial sum itself so that the write is ordered before its own copies.
__device__ __forceinline__ void reg_send(const float* src_rows, float* dst_rows, int nrow, float* part_slot,
                                         float own, uint64_t* bar, int c, int warp) {
    *part_slot = own;
    fence_async_smem();
    const uint32_t peer = (uint32_t)((c + 1 + warp) % kRK);
    const uint32_t bar_r = tc::mapa_u32(tc::smem_u32(bar), peer);
    tc::dsmem_bulk_copy(tc::mapa_u32(tc::smem_u32(dst_rows), peer), tc::smem_u32(src_rows), (uint32_t)nrow * 4u, bar_r);
    tc::dsmem_bulk_copy(tc::mapa_u32(tc::smem_u32(part_slot), peer), tc::smem_u32(part_slot), 16u, bar_r);
}

// cluster-wide sum of the 8 partial sums of a frame (16-byte slots); the own slot may not be visible yet
__device__ __forceinline__ float sum_parts(uint32_t parts_s, int c, float own) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kRK; ++k) {
        const float v = lds_f32(parts_s + k * 16);
        t += (k == c) ? own : v;
    }
    return t;
}

// e = exp(clamp(loglikes, -30, 30)) for the valid frames of every sequence, once per call: the frame loops
// of the register-resident kernels then bulk-copy e rows straight into shared memory
__global__ void __launch_bounds__(256) den_exp_kernel(const float* __restrict__ ll, float* __restrict__ e,
                                                      const int32_t* __restrict__ num_frames, int64_t row_stride_b,
                                                      int max_frames, int N) {
    const int b = blockIdx.y, t = blockIdx.x;
    if (t >= num_frames[b]) return;
    const float* src = ll + ((int64_t)b * row_stride_b + t) * N;
    float* dst = e + ((int64_t)b * max_frames + t) * N;
    for (int i = threadIdx.x * 4; i < N; i += 256 * 4) {
        float4 x = *reinterpret_cast<const float4*>(src + i);
        x.x = __expf(fminf(fmaxf(x.x, -30.f), 30.f)); x.y = __expf(fminf(fmaxf(x.y, -30.f), 30.f));
        x.z = __expf(fminf(fmaxf(x.z, -30.f), 30.f)); x.w = __expf(fminf(fmaxf(x.w, -30.f), 30.f));
        *reinterpret_cast<float4*>(dst + i) = x;
    }
}

struct RegSmem {                   // host-computed shared-memory carve-up (element counts)
    int ovf_f, ovf_b;              // overflow records (forward / backward table)
    int garcs;                     // pdf-table records per CTA
    int nslg;                      // pdf-table slices per CTA
    int gcap;                      // pdf rows per CTA
};

#define PK2_RPROF(e) do { if (prof && wi_ == w0 && t >= 64 && t < 72) prof[(t - 64) * 8 + (e)] = clock64(); } while (0)

__global__ void __launch_bounds__(kRT, 1) den_forward_reg_kernel(DenArgs a, RegSmem rs) {
    extern __shared__ __align__(16) float smem[];
    const int S = a.S, N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (int)tc::cluster_ctarank();
    const int q = blockIdx.x / kRK;
    const int r0 = a.fwd.part_row[c], r1 = a.fwd.part_row[c + 1], nrow = r1 - r0;
    float* buf = smem;                      // [2][S]  alpha (without the leaky term)
    float* E = buf + 2 * S;                 // [2][N]  e(t), e(t+1)   (N % 4 == 0)
    float* wsum = E + 2 * N;                // [2][32] per-warp partial sums, by frame parity
    float* parts = wsum + 64;               // [2][kRK * 4]
    float* red = parts + 2 * kRK * 4;       // [40]
    int* wcnt = reinterpret_cast<int*>(red + 40);                  // [32]
    double* dred = reinterpret_cast<double*>(wcnt + 32);           // [32]
    uint64_t* xbar = reinterpret_cast<uint64_t*>(dred + 32);       // [2] row exchange
    uint64_t* ebar = xbar + 2;                                     // [2] e rows
    uint4* ovf = reinterpret_cast<uint4*>(xbar + 4);

    if (a.start_flag && blockIdx.x == 0 && threadIdx.x == 0) {      // this kernel's CTAs are placed (hybrid schedule)
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.start_flag), "r"(a.epoch) : "memory");
    }
    RegSlice sl[kRowsPerThread];
    float wi[kRowsPerThread][kRegArcs];
    load_reg_slices<true>(a.fwd, c, a.init, sl, wi, ovf, wcnt);
    float init_own[2];                      // init of the rows this thread stores to the alpha workspace
    float s = 0.f;
    for (int j = threadIdx.x; j < S; j += kRT) s += __ldg(&a.init[j]);
#pragma unroll
    for (int r = 0; r < 2; ++r) init_own[r] = (threadIdx.x + r * kRT < nrow) ? __ldg(&a.init[r0 + threadIdx.x + r * kRT]) : 0.f;
    const float A0 = block_sum_nw<kRW>(s, red);
    const uint32_t expect = (uint32_t)(S - nrow) * 4u + (uint32_t)(kRK - 1) * 16u;
    const uint32_t ebytes = (uint32_t)N * 4u;
    if (threadIdx.x == 0) {
        tc::mbar_init(&xbar[0], 1); tc::mbar_init(&xbar[1], 1);
        tc::mbar_init(&ebar[0], 1); tc::mbar_init(&ebar[1], 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&xbar[0], expect);
        tc::mbar_expect_tx(&xbar[1], expect);
    }
    tc::cluster_sync_all();                 // every peer's barriers are armed before the first copy is sent

    const int w0 = a.work[q], w1 = a.work[q + 1];
    const int32_t* ids = a.work + a.work_ids;
    long long* prof = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? a.prof : nullptr;
    const uint32_t buf_s = tc::smem_u32(buf), E_s = tc::smem_u32(E), parts_s = tc::smem_u32(parts);
    uint32_t f = 0, phase = 0, ephase = 0;  // global frame counter: parity of buffers and barriers
    for (int wi_ = w0; wi_ < w1; ++wi_) {
        const int b = ids[wi_];
        const int T = a.num_frames[b];
        const float* e = a.e + (int64_t)b * a.max_frames * N;
        float* aws = a.alpha_ws + (int64_t)b * a.max_frames * S;
        float* asum = a.asum_ws + (int64_t)b * (a.max_frames + 2);
        {
            float* cur = buf + (f & 1) * S;             // alpha(0) = init
            for (int j = threadIdx.x; j < S; j += kRT) cur[j] = __ldg(&a.init[j]);
            if (threadIdx.x == 0) {
                if (T > 0) { tc::mbar_expect_tx(&ebar[f & 1], ebytes); tc::bulk_load(E + (f & 1) * N, e, ebytes, &ebar[f & 1]); }
                if (T > 1) { tc::mbar_expect_tx(&ebar[(f + 1) & 1], ebytes); tc::bulk_load(E + ((f + 1) & 1) * N, e + N, ebytes, &ebar[(f + 1) & 1]); }
            }
        }
        float A = A0;
        float lk = a.leaky * A0;                        // alpha'(t) = alpha(t) + lk * init
        __syncthreads();
        for (int t = 0; t < T; ++t, ++f) {
            PK2_RPROF(0);
            const int par = (int)(f & 1);
            float* cur = buf + par * S;
            float* nxt = buf + (par ^ 1) * S;
            // alpha'(t) of this CTA's rows -> workspace (read by the backward kernel), coalesced
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = threadIdx.x + r * kRT;
                if (i < nrow) aws[(int64_t)t * S + r0 + i] = fmaf(lk, init_own[r], cur[r0 + i]);
            }
            if (threadIdx.x == 0 && c == 0) asum[t] = A;
            const float invA = 1.0f / A;
            const uint32_t cur_s = pin_u32(buf_s + (uint32_t)(par * S) * 4u), Ec_s = pin_u32(E_s + (uint32_t)(par * N) * 4u);
            mbar_wait_guard(&ebar[par], (ephase >> par) & 1u);          // e(t) has landed
            ephase ^= (1u << par);
            float vsum = 0.f;
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                float acc, z;
                reg_slice_pass<true>(sl[r], wi[r], cur_s, Ec_s, acc, z);
                if (sl[r].row >= 0) {
                    const float v = fmaf(lk, z, acc) * invA;      // alpha(t+1, row)
                    nxt[sl[r].row] = v;
                    vsum += v;
                }
            }
            const float ws = pk2::warp_sum(vsum);
            if (lane == 0) wsum[par * 32 + warp] = ws;
            fence_async_smem();                              // rows -> async proxy
            PK2_RPROF(1);
            __syncthreads();                                 // pass done in this CTA
            const float own = pk2::warp_sum(lane < kRW ? wsum[par * 32 + lane] : 0.f);     // block total, bit-identical in every warp
            if (lane == 0) {
                if (warp < kRK - 1) {
                    reg_send(nxt + r0, nxt + r0, nrow, parts + par * kRK * 4 + c * 4, own, &xbar[par], c, warp);
                } else if (warp == kRK - 1 && t + 2 < T) {   // e(t+2) into the buffer the pass has just released
                    tc::mbar_expect_tx(&ebar[par], ebytes);
                    tc::bulk_load(E + par * N, e + (int64_t)(t + 2) * N, ebytes, &ebar[par]);
                }
            }
            PK2_RPROF(2);
            mbar_wait_guard(&xbar[par], (phase >> par) & 1u);
            phase ^= (1u << par);
            if (threadIdx.x == 0) tc::mbar_expect_tx(&xbar[par], expect);      // arm for frame f + 2
            PK2_RPROF(3);
            const float An = sum_parts(parts_s + (uint32_t)par * kRK * 16u, c, own);
            lk = a.leaky * An;
            A = An;
            PK2_RPROF(4);
        }
        // total probability sum_j alpha'(T, j) = A(T) + lk * sum(init) and log Z
        const float totp = fmaf(lk, A0, A);
        if (c == 0) {
            __threadfence_block();
            __syncthreads();
            double part = 0.0;
            for (int t = threadIdx.x; t < T; t += kRT) part += log((double)asum[t]);
            part = pk2::warp_sum_d(part);
            if (lane == 0) dred[warp] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                double logsum = 0.0;
                for (int w = 0; w < kRW; ++w) logsum += dred[w];
                asum[T] = A;
                asum[a.max_frames + 1] = totp;
                a.logz[b] = log((double)totp) + logsum;
            }
        }
        __syncthreads();
    }
    tc::cluster_sync_all();                 // no CTA exits while a peer may still copy into it
}

// ---- forward, two sequences interleaved per cluster -------------------------------------------------------
// In den_forward_reg_kernel a third of the frame is spent waiting for the DSMEM row exchange (28 KB out and
// in per CTA at ~17 B/clk) with nothing else to do.  Here a cluster works on TWO sequences ("slots", own
// alpha / e buffers and mbarriers, the same arc registers): while the rows of one slot travel, the pass of
// the other slot runs.  Each slot pulls the next sequence of the cluster's work list when its own ends.
struct FSlot {
    int b, T, t;                 // sequence, its frames, current frame; T < 0: slot idle (work list exhausted)
    int pend;                    // rows of frame t have been sent, the exchange is not yet complete
    float A, lk, own;
    uint32_t f, phase, ephase;   // per-slot frame counter: parity of buffers and barriers
};

__global__ void __launch_bounds__(kRT, 1) den_forward_reg2_kernel(DenArgs a, RegSmem rs) {
    extern __shared__ __align__(16) float smem[];
    const int S = a.S, N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (int)tc::cluster_ctarank();
    const int q = blockIdx.x / kRK;
    const int r0 = a.fwd.part_row[c], r1 = a.fwd.part_row[c + 1], nrow = r1 - r0;
    // slot block: buf [2][S] | E [2][N] | wsum [2][32] | parts [2][kRK*4] | xbar [2] | ebar [2]
    const int slot_floats = 2 * S + 2 * N + 64 + 2 * kRK * 4 + 8;
    float* red = smem + 2 * slot_floats;                           // [40]
    int* wcnt = reinterpret_cast<int*>(red + 40);                  // [32]
    double* dred = reinterpret_cast<double*>(wcnt + 32);           // [32]
    uint4* ovf = reinterpret_cast<uint4*>(dred + 32);

    if (a.start_flag && blockIdx.x == 0 && threadIdx.x == 0) {      // this kernel's CTAs are placed (hybrid schedule)
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(a.start_flag), "r"(a.epoch) : "memory");
    }
    RegSlice sl[kRowsPerThread];
    float wi[kRowsPerThread][kRegArcs];
    load_reg_slices<true>(a.fwd, c, a.init, sl, wi, ovf, wcnt);
    float init_own[2];
    float s0 = 0.f;
    for (int j = threadIdx.x; j < S; j += kRT) s0 += __ldg(&a.init[j]);
#pragma unroll
    for (int r = 0; r < 2; ++r) init_own[r] = (threadIdx.x + r * kRT < nrow) ? __ldg(&a.init[r0 + threadIdx.x + r * kRT]) : 0.f;
    const float A0 = block_sum_nw<kRW>(s0, red);
    const uint32_t expect = (uint32_t)(S - nrow) * 4u + (uint32_t)(kRK - 1) * 16u;
    const uint32_t ebytes = (uint32_t)N * 4u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + s * slot_floats + 2 * S + 2 * N + 64 + 2 * kRK * 4);
            tc::mbar_init(&xbar[0], 1); tc::mbar_init(&xbar[1], 1);
            tc::mbar_init(&xbar[2], 1); tc::mbar_init(&xbar[3], 1);
        }
        tc::fence_barrier_init();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            uint64_t* xbar = reinterpret_cast<uint64_t*>(smem + s * slot_floats + 2 * S + 2 * N + 64 + 2 * kRK * 4);
            tc::mbar_expect_tx(&xbar[0], expect);
            tc::mbar_expect_tx(&xbar[1], expect);
        }
    }
    tc::cluster_sync_all();

    const int w1 = a.work[q + 1];
    int widx = a.work[q];
    const int32_t* ids = a.work + a.work_ids;
    const uint32_t smem_s = tc::smem_u32(smem);

    FSlot st[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) { st[s].T = -1; st[s].t = 0; st[s].pend = 0; st[s].f = 0; st[s].phase = 0; st[s].ephase = 0;
                                  st[s].b = 0; st[s].A = A0; st[s].lk = 0.f; st[s].own = 0.f; }

    // log Z of a finished (or empty) sequence; all threads of the CTA call it
    auto finish = [&](FSlot& z) {
        float* asum = a.asum_ws + (int64_t)z.b * (a.max_frames + 2);
        const float totp = fmaf(z.lk, A0, z.A);          // sum_j alpha'(T, j) = A(T) + lk * sum(init)
        if (c == 0) {
            __threadfence_block();
            __syncthreads();
            double part = 0.0;
            for (int t = threadIdx.x; t < z.T; t += kRT) part += log((double)asum[t]);
            part = pk2::warp_sum_d(part);
            if (lane == 0) dred[warp] = part;
            __syncthreads();
            if (threadIdx.x == 0) {
                double logsum = 0.0;
                for (int w = 0; w < kRW; ++w) logsum += dred[w];
                asum[z.T] = z.A;
                asum[a.max_frames + 1] = totp;
                a.logz[z.b] = log((double)totp) + logsum;
            }
            __syncthreads();
        }
    };
    // pull the next sequence of the work list into slot s (slot buffers are free: the caller has synchronised)
    auto start_next = [&](FSlot& z, float* slot) {
        for (;;) {
            if (widx >= w1) { z.T = -1; return; }
            z.b = ids[widx++];
            z.T = a.num_frames[z.b];
            z.t = 0; z.pend = 0;
            z.A = A0; z.lk = a.leaky * A0;
            if (z.T > 0) break;
            finish(z);                                     // empty sequence: log Z = log sum alpha'(0)
        }
        float* cur = slot + (z.f & 1) * S;
        for (int j = threadIdx.x; j < S; j += kRT) cur[j] = __ldg(&a.init[j]);
        if (threadIdx.x == 0) {
            const float* e = a.e + (int64_t)z.b * a.max_frames * N;
            float* E = slot + 2 * S;
            uint64_t* ebar = reinterpret_cast<uint64_t*>(slot + 2 * S + 2 * N + 64 + 2 * kRK * 4) + 2;
            tc::mbar_expect_tx(&ebar[z.f & 1], ebytes);
            tc::bulk_load(E + (z.f & 1) * N, e, ebytes, &ebar[z.f & 1]);
            if (z.T > 1) {
                tc::mbar_expect_tx(&ebar[(z.f + 1) & 1], ebytes);
                tc::bulk_load(E + ((z.f + 1) & 1) * N, e + N, ebytes, &ebar[(z.f + 1) & 1]);
            }
        }
    };

#pragma unroll
    for (int s = 0; s < 2; ++s) start_next(st[s], smem + s * slot_floats);
    __syncthreads();

    while (st[0].T >= 0 || st[1].T >= 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            FSlot& z = st[s];
            if (z.T < 0) continue;
            float* slot = smem + s * slot_floats;
            const uint32_t slot_s = smem_s + (uint32_t)(s * slot_floats) * 4u;
            float* wsum = slot + 2 * S + 2 * N;
            float* parts = wsum + 64;
            uint64_t* xbar = reinterpret_cast<uint64_t*>(parts + 2 * kRK * 4);
            uint64_t* ebar = xbar + 2;
            if (z.pend) {
                // complete frame t of this slot: the peers' rows travelled while the other slot was computing
                const int par = (int)(z.f & 1);
                mbar_wait_guard(&xbar[par], (z.phase >> par) & 1u);
                z.phase ^= (1u << par);
                if (threadIdx.x == 0) tc::mbar_expect_tx(&xbar[par], expect);
                const float An = sum_parts(slot_s + (uint32_t)(2 * S + 2 * N + 64 + par * kRK * 4) * 4u, c, z.own);
                z.lk = a.leaky * An;
                z.A = An;
                ++z.t; ++z.f; z.pend = 0;
                if (z.t == z.T) {
                    finish(z);
                    __syncthreads();
                    start_next(z, slot);
                    __syncthreads();
                    if (z.T < 0) continue;
                }
            }
            const int par = (int)(z.f & 1);
            const int t = z.t;
            float* cur = slot + par * S;
            float* nxt = slot + (par ^ 1) * S;
            float* aws = a.alpha_ws + (int64_t)z.b * a.max_frames * S;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int i = threadIdx.x + r * kRT;
                if (i < nrow) aws[(int64_t)t * S + r0 + i] = fmaf(z.lk, init_own[r], cur[r0 + i]);
            }
            if (threadIdx.x == 0 && c == 0) a.asum_ws[(int64_t)z.b * (a.max_frames + 2) + t] = z.A;
            const float invA = 1.0f / z.A;
            const uint32_t cur_s = pin_u32(slot_s + (uint32_t)(par * S) * 4u);
            const uint32_t Ec_s = pin_u32(slot_s + (uint32_t)(2 * S + par * N) * 4u);
            mbar_wait_guard(&ebar[par], (z.ephase >> par) & 1u);          // e(t) has landed
            z.ephase ^= (1u << par);
            float vsum = 0.f;
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                float acc, zz;
                reg_slice_pass<true>(sl[r], wi[r], cur_s, Ec_s, acc, zz);
                if (sl[r].row >= 0) {
                    const float v = fmaf(z.lk, zz, acc) * invA;
                    nxt[sl[r].row] = v;
                    vsum += v;
                }
            }
            const float ws = pk2::warp_sum(vsum);
            if (lane == 0) wsum[par * 32 + warp] = ws;
            fence_async_smem();
            __syncthreads();
            z.own = pk2::warp_sum(lane < kRW ? wsum[par * 32 + lane] : 0.f);
            if (lane == 0) {
                if (warp < kRK - 1) {
                    reg_send(nxt + r0, nxt + r0, nrow, parts + par * kRK * 4 + c * 4, z.own, &xbar[par], c, warp);
                } else if (warp == kRK - 1 && t + 2 < z.T) {
                    const float* e = a.e + (int64_t)z.b * a.max_frames * N;
                    tc::mbar_expect_tx(&ebar[par], ebytes);
                    tc::bulk_load(slot + 2 * S + par * N, e + (int64_t)(t + 2) * N, ebytes, &ebar[par]);
                }
            }
            z.pend = 1;
        }
    }
    tc::cluster_sync_all();
}

size_t reg_fwd2_smem_bytes(int S, int N, const RegSmem& rs) {
    return sizeof(float) * (2 * (size_t)(2 * S + 2 * N + 64 + 2 * kRK * 4 + 8) + 40) + sizeof(int) * 32 +
           sizeof(double) * 32 + sizeof(uint4) * (size_t)rs.ovf_f + 16;
}

// Sum over the arcs of the rows of a SELL table held in SHARED memory (pdf-occupancy pass):
// acc = sum ga[a] * (w * gb[b]), acc2 = sum ga[a] * w.  Slices are sorted by length: odd rounds go to the
// warps in reverse order so that every warp gets about the same number of arcs.
template <class Body>
__device__ __forceinline__ void sell_pass_smem(const uint2* arcs, const int2* meta, const uint16_t* rows, int nsl,
                                               uint32_t base_a, uint32_t base_b, Body&& body) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int rd = 0; rd * kRW < nsl; ++rd) {
        const int i = rd * kRW + ((rd & 1) ? kRW - 1 - warp : warp);
        if (i >= nsl) continue;
        const int2 m = meta[i];
        const uint2* p = arcs + m.x + lane;
        const unsigned row = rows[i * 32 + lane];
        float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
        int k = 0;
        for (; k + 4 <= m.y; k += 4) {
            uint2 r[4];
            float ga[4], gb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = p[(k + u) * 32];
#pragma unroll
            for (int u = 0; u < 4; ++u) { ga[u] = lds_f32(base_a + (r[u].y & 0xffffu)); gb[u] = lds_f32(base_b + (r[u].y >> 16)); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float w = __uint_as_float(r[u].x);
                if (u & 1) { a1 = fmaf(ga[u], w * gb[u], a1); c1 = fmaf(ga[u], w, c1); }
                else       { a0 = fmaf(ga[u], w * gb[u], a0); c0 = fmaf(ga[u], w, c0); }
            }
        }
        for (; k < m.y; ++k) {
            const uint2 r = p[k * 32];
            const float w = __uint_as_float(r.x);
            const float g = lds_f32(base_a + (r.y & 0xffffu));
            a0 = fmaf(g, w * lds_f32(base_b + (r.y >> 16)), a0);
            c0 = fmaf(g, w, c0);
        }
        if (row != 0xFFFFu) body((int)row, a0 + a1, c0 + c1);
    }
}

__global__ void __launch_bounds__(kRT, 1) den_backward_reg_kernel(DenArgs a, RegSmem rs) {
    extern __shared__ __align__(16) float smem[];
    const int S = a.S, N = a.N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = (int)tc::cluster_ctarank();
    const int q = blockIdx.x / kRK;
    const int r0 = a.bwd.part_row[c], r1 = a.bwd.part_row[c + 1], nrow = r1 - r0;
    const int p0 = a.pdf.part_row[c], p1 = a.pdf.part_row[c + 1];
    float* buf = smem;                      // [2][S]  beta' (without the leaky term)
    float* al = buf + 2 * S;                // [S]     alpha'(t)
    float* E = al + S;                      // [2][N]  e(t), e(t-1)
    float* gbuf = E + 2 * N;                // [gcap]  gamma of this CTA's pdf range
    float* wsum = gbuf + rs.gcap;           // [32]
    float* parts = wsum + 32;               // [2][kRK * 4]
    float* red = parts + 2 * kRK * 4;       // [40]
    int* wcnt = reinterpret_cast<int*>(red + 40);                  // [32]
    uint64_t* xbar = reinterpret_cast<uint64_t*>(wcnt + 32);       // [2] row exchange
    uint64_t* cbar = xbar + 2;                                     // [2] credits
    uint64_t* ebar = xbar + 4;                                     // [2] e rows
    uint64_t* abar = xbar + 6;                                     // [1] alpha row (+1 pad)
    uint4* ovf = reinterpret_cast<uint4*>(xbar + 8);               // [ovf_b]
    uint2* garcs = reinterpret_cast<uint2*>(ovf + rs.ovf_b);       // [garcs]  (offsets pre-scaled)
    int2* meta_g = reinterpret_cast<int2*>(garcs + rs.garcs);      // [nslg]
    uint16_t* rows_g = reinterpret_cast<uint16_t*>(meta_g + rs.nslg);   // [nslg * 32]

    RegSlice sl[kRowsPerThread];
    float wi[kRowsPerThread][kRegArcs];     // unused (WI = false): optimised away
    load_reg_slices<false>(a.bwd, c, a.init, sl, wi, ovf, wcnt);
    float init_row[kRowsPerThread];
#pragma unroll
    for (int r = 0; r < kRowsPerThread; ++r) init_row[r] = sl[r].row >= 0 ? __ldg(&a.init[sl[r].row]) : 0.f;
    // pdf table of this CTA -> shared memory (slice offsets relative to the CTA's first record)
    const int gs0 = a.pdf.part_slice[c], nslg = a.pdf.part_slice[c + 1] - gs0;
    const int gbase = __ldg(&a.pdf.slice_off[gs0]);
    const int gcnt = __ldg(&a.pdf.slice_off[gs0 + nslg]) - gbase;
    for (int i = threadIdx.x; i < nslg; i += kRT) {
        const int off = __ldg(&a.pdf.slice_off[gs0 + i]);
        meta_g[i] = make_int2(off - gbase, (__ldg(&a.pdf.slice_off[gs0 + i + 1]) - off) >> 5);
    }
    for (int i = threadIdx.x; i < nslg * 32; i += kRT) {
        const int r = __ldg(&a.pdf.slice_row[gs0 * 32 + i]);
        rows_g[i] = r < 0 ? (uint16_t)0xFFFFu : (uint16_t)r;
    }
    for (int i = threadIdx.x; i < gcnt; i += kRT) {
        const uint2 rec = __ldg(a.pdf.arcs + gbase + i);
        garcs[i] = make_uint2(rec.x, prescale(rec.y));
    }
    float s = 0.f;
    for (int j = threadIdx.x; j < S; j += kRT) s += __ldg(&a.init[j]);
    const float isum = block_sum_nw<kRW>(s, red);
    const uint32_t expect = (uint32_t)(S - nrow) * 4u + (uint32_t)(kRK - 1) * 16u;
    const uint32_t ebytes = (uint32_t)N * 4u, abytes = (uint32_t)S * 4u;
    if (threadIdx.x == 0) {
        tc::mbar_init(&xbar[0], 1); tc::mbar_init(&xbar[1], 1);
        tc::mbar_init(&cbar[0], kRK - 1); tc::mbar_init(&cbar[1], kRK - 1);
        tc::mbar_init(&ebar[0], 1); tc::mbar_init(&ebar[1], 1); tc::mbar_init(abar, 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&xbar[0], expect);
        tc::mbar_expect_tx(&xbar[1], expect);
    }
    tc::cluster_sync_all();

    const int w0 = a.work[q], w1 = a.work[q + 1];
    const int32_t* ids = a.work + a.work_ids;
    long long* prof = (a.prof && blockIdx.x == 0 && threadIdx.x == 0) ? a.prof + 64 : nullptr;
    const uint32_t buf_s = tc::smem_u32(buf), E_s = tc::smem_u32(E), al_s = tc::smem_u32(al), parts_s = tc::smem_u32(parts);
    uint32_t f = 0, phase = 0, ephase = 0, aphase = 0;
    for (int wi_ = w0; wi_ < w1; ++wi_) {
        const int b = ids[wi_];
        const int T = a.num_frames[b];
        const float* e = a.e + (int64_t)b * a.max_frames * N;
        float* grad = a.grad + (int64_t)b * a.row_stride_b * N;
        const float* aws = a.alpha_ws + (int64_t)b * a.max_frames * S;
        const float* asum = a.asum_ws + (int64_t)b * (a.max_frames + 2);
        for (int t = T; t < a.max_frames; ++t)
            for (int p = p0 + threadIdx.x; p < p1; p += kRT) grad[(int64_t)t * N + p] = 0.f;
        if (T <= 0) continue;               // uniform across the cluster
        const float totp = asum[a.max_frames + 1];
        const float bT = 1.0f / totp;
        float lk = a.leaky * isum * bT;      // beta = beta' + lk (leaky term applied lazily)
        {
            float* cur = buf + (f & 1) * S;
            for (int j = threadIdx.x; j < S; j += kRT) cur[j] = bT;
            if (threadIdx.x == 0) {
                tc::mbar_expect_tx(&ebar[f & 1], ebytes);
                tc::bulk_load(E + (f & 1) * N, e + (int64_t)(T - 1) * N, ebytes, &ebar[f & 1]);
                if (T > 1) {
                    tc::mbar_expect_tx(&ebar[(f + 1) & 1], ebytes);
                    tc::bulk_load(E + ((f + 1) & 1) * N, e + (int64_t)(T - 2) * N, ebytes, &ebar[(f + 1) & 1]);
                }
                tc::mbar_expect_tx(abar, abytes);
                tc::bulk_load(al, aws + (int64_t)(T - 1) * S, abytes, abar);
            }
        }
        float asum_t = asum[T - 1];
        __syncthreads();
        for (int t = T - 1; t >= 0; --t, ++f) {
            PK2_RPROF(0);
            const int par = (int)(f & 1);
            float* nxt = buf + (par ^ 1) * S;   // beta'(t)
            float* Ec = E + par * N;            // e(t)
            const uint32_t cur_s = pin_u32(buf_s + (uint32_t)(par * S) * 4u), Ec_s = pin_u32(E_s + (uint32_t)(par * N) * 4u);
            const float invA = 1.0f / asum_t;
            if (t > 0) asum_t = asum[t - 1];
            mbar_wait_guard(&ebar[par], (ephase >> par) & 1u);          // e(t) has landed
            ephase ^= (1u << par);
            float d = 0.f;
#pragma unroll
            for (int r = 0; r < kRowsPerThread; ++r) {
                float acc, acc2;
                reg_slice_pass<false>(sl[r], wi[r], cur_s, Ec_s, acc, acc2);
                if (sl[r].row >= 0) {
                    const float v = fmaf(lk, acc2, acc) * invA;
                    nxt[sl[r].row] = v;
                    d = fmaf(v, init_row[r], d);
                }
            }
            const float ws = pk2::warp_sum(d);
            if (lane == 0) wsum[warp] = ws;
            fence_async_smem();
            PK2_RPROF(1);
            __syncthreads();                                 // [1] beta rows of this CTA complete
            const float own = pk2::warp_sum(lane < kRW ? wsum[lane] : 0.f);
            if (warp < kRK - 1 && lane == 0) {
                // The peer's gamma pass of frame f-1 reads the buffer this frame's rows land in: wait for its
                // credit.  Credits alternate between two barriers: a peer can give its credit for frame f
                // before this CTA has looked at the credits of frame f-1 (it does not need our frame-f rows
                // for that), and a single barrier two phases ahead would look "not complete" to a parity wait.
                if (f > 0) mbar_wait_guard(&cbar[(f - 1) & 1u], ((f - 1) >> 1) & 1u);
                reg_send(nxt + r0, nxt + r0, nrow, parts + par * kRK * 4 + c * 4, own, &xbar[par], c, warp);
            }
            PK2_RPROF(2);
            // off the dependency chain, while the rows travel: pdf occupancies gamma(t, p) of this CTA's pdf range
            mbar_wait_guard(abar, aphase & 1u);                         // alpha'(t) has landed
            aphase ^= 1u;
            const float gs = a.deriv_scale * invA;
            sell_pass_smem(garcs, meta_g, rows_g, nslg, al_s, cur_s, [&](int row, float ac, float ac2) {
                gbuf[row - p0] = fmaf(lk, ac2, ac) * Ec[row] * gs;
            });
            PK2_RPROF(3);
            __syncthreads();                                 // [2] reads of al / cur / Ec done, gbuf complete
            if (lane == 0) {
                if (warp < kRK - 1) {
                    tc::mbar_arrive_remote_relaxed(tc::mapa_u32(tc::smem_u32(&cbar[par]), (uint32_t)((c + 1 + warp) % kRK)));
                } else if (warp == kRK - 1) {
                    if (t > 0) { tc::mbar_expect_tx(abar, abytes); tc::bulk_load(al, aws + (int64_t)(t - 1) * S, abytes, abar); }
                    if (t > 1) { tc::mbar_expect_tx(&ebar[par], ebytes); tc::bulk_load(Ec, e + (int64_t)(t - 2) * N, ebytes, &ebar[par]); }
                }
            }
            for (int p = threadIdx.x; p < p1 - p0; p += kRT) grad[(int64_t)t * N + p0 + p] = gbuf[p];
            PK2_RPROF(4);
            mbar_wait_guard(&xbar[par], (phase >> par) & 1u);
            phase ^= (1u << par);
            if (threadIdx.x == 0) tc::mbar_expect_tx(&xbar[par], expect);
            lk = a.leaky * sum_parts(parts_s + (uint32_t)par * kRK * 16u, c, own);
            PK2_RPROF(5);
        }
        __syncthreads();
    }
    tc::cluster_sync_all();
}
#undef PK2_RPROF

size_t reg_fwd_smem_bytes(int S, int N, const RegSmem& rs) {
    return sizeof(float) * (2 * (size_t)S + 2 * (size_t)N + 64 + 2 * kRK * 4 + 40) +
           sizeof(int) * 32 + sizeof(double) * 32 + sizeof(uint64_t) * 4 + sizeof(uint4) * (size_t)rs.ovf_f + 16;
}
size_t reg_bwd_smem_bytes(int S, int N, const RegSmem& rs) {
    return sizeof(float) * (2 * (size_t)S + S + 2 * (size_t)N + rs.gcap + 32 + 2 * kRK * 4 + 40) +
           sizeof(int) * 32 + sizeof(uint64_t) * 8 + sizeof(uint4) * (size_t)rs.ovf_b + sizeof(int2) * (size_t)rs.nslg +
           sizeof(uint2) * (size_t)rs.garcs + sizeof(uint16_t) * (size_t)rs.nslg * 32 + 16;
}

size_t fwd_smem_bytes(int S, int N, int K) {
    const size_t Sp = (S + 3) & ~3, Np = (N + 3) & ~3;
    const size_t rows_part = (size_t)((S + K - 1) / K + 64);            // rows staged per CTA (+ slice padding)
    const size_t rcap = (rows_part + 3) & ~(size_t)3;
    return sizeof(float) * (2 * Sp + 2 * Np + 2 * kMaxK * kWarps + kWarps + 4 + (K > 1 ? 2 * rcap : 0)) +
           sizeof(int2) * ((S + 31) / 32 + 2) + sizeof(uint16_t) * (rows_part + 32 + 32) + 64 + 32;
}
size_t bwd_smem_bytes(int S, int N, int K) {
    const size_t Sp = (S + 3) & ~3, Np = (N + 3) & ~3;
    const size_t gcap = ((N + K - 1) / K + 64 + 3) & ~3;
    const size_t rows_part = (size_t)((S + K - 1) / K + 64) + (size_t)((N + K - 1) / K + 64);
    return sizeof(float) * ((K > 1 ? 4 : 3) * Sp + 2 * Np + gcap + 2 * kMaxK * kWarps + kWarps + 4) +
           sizeof(int2) * ((S + 31) / 32 + (N + 31) / 32 + 4) + sizeof(uint16_t) * (rows_part + 64) + 64 + 16;
}

long long* g_den_prof = nullptr;     // pk2_den_set_profile_buffer (profiling only)

template <int K>
int launch_den(const DenArgs& args, int n_seq, cudaStream_t st) {
    if (n_seq <= 0) return 0;
    const size_t sf = fwd_smem_bytes(args.S, args.N, K), sb = bwd_smem_bytes(args.S, args.N, K);
    PK2_REQUIRE(sb <= 227 * 1024, "pk2_denfb: graph too large for shared memory (S=%d N=%d needs %zu B)",
                args.S, args.N, sb);
    PK2_CHECK(cudaFuncSetAttribute(den_forward_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf));
    PK2_CHECK(cudaFuncSetAttribute(den_backward_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_seq * K);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = K; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.dynamicSmemBytes = sf;
    PK2_CHECK(cudaLaunchKernelEx(&cfg, den_forward_kernel<K>, args));
    PK2_LAUNCHED();
    cfg.dynamicSmemBytes = sb;
    PK2_CHECK(cudaLaunchKernelEx(&cfg, den_backward_kernel<K>, args));
    PK2_LAUNCHED();
    return 0;
}

int launch_den_k(int K, const DenArgs& a, int n_seq, cudaStream_t st) {
    if (K == 1) return launch_den<1>(a, n_seq, st);
    if (K == 2) return launch_den<2>(a, n_seq, st);
    return launch_den<4>(a, n_seq, st);
}

int ensure_tables(DenGraph* g, int K) {
    std::lock_guard<std::mutex> lk(g->mu);
    if (!g->built[K]) {
        if (build_sell(g->rows_fwd, K, &g->t_fwd[K])) return 1;
        if (build_sell(g->rows_bwd, K, &g->t_bwd[K])) return 1;
        if (build_sell(g->rows_pdf, K, &g->t_pdf[K])) return 1;
        g->built[K] = true;
    }
    return 0;
}

// Mixed-cluster schedule: the kernels are persistent per sequence, so a launch lasts as long as its
// longest sequence.  Give long sequences more CTAs (K = 4), short ones fewer (K = 1) so that all
// clusters of the batch fit the SMs in ONE wave and finish at about the same time.  Relative
// per-frame cost of a cluster of K CTAs.  Isolated measurements give 1.46 / 1.0 / 0.72 (K = 1 / 2 / 4,
// profiles/kernel_bench_den_r1_v6.jsonl) but with ~140 CTAs sharing L2 the schedule planned with
// 1.9 / 1.0 / 0.62 is faster end to end (14.7 vs 15.7 ms, kernel_bench_den_r1_v9.jsonl).
void plan_clusters(const int32_t* frames, int n, int budget, std::vector<int>* ks) {
    static const double cost[5] = {0, 1.9, 1.0, 0, 0.62};
    ks->assign(n, 1);
    int used = n;
    for (;;) {
        int worst = -1;
        double wt = -1.0;
        for (int i = 0; i < n; ++i) {
            const double t = frames[i] * cost[(*ks)[i]];
            if (t > wt) { wt = t; worst = i; }
        }
        if (worst < 0 || (*ks)[worst] == 4) break;
        const int k = (*ks)[worst], extra = k;            // 1->2 costs 1 CTA, 2->4 costs 2
        if (used + extra > budget) break;
        (*ks)[worst] = 2 * k;
        used += extra;
    }
}


struct RegSmemHost { RegSmem rs; size_t smem_f, smem_b, smem_f2; int max_clusters; bool two_slot; };

int launch_cluster8(void (*kern)(DenArgs, RegSmem), const DenArgs& args, const RegSmem& rs, int n_clusters,
                    size_t smem, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_clusters * kRK);
    cfg.blockDim = dim3(kRT);
    cfg.stream = st;
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kRK; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // per launch, not per graph: the attribute belongs to the function, and graphs of different sizes share it
    PK2_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PK2_CHECK(cudaLaunchKernelEx(&cfg, kern, args, rs));
    PK2_LAUNCHED();
    return 0;
}

// Decide once per graph whether the register-resident kernels apply: S splits into 8 parts of whole
// 32-row slices with at most one row per thread, indices fit the packed records, and the backward
// kernel's working set (beta x2, alpha, e x2, the CTA's pdf table) fits the 227 KB of shared memory.
int plan_reg(DenGraph* g) {
    if (g->reg_state != 0) return 0;
    g->reg_state = -1;
    const int S = g->S, N = g->N;
    if (S % (kRK * 32) != 0 || S / kRK > kRT * kRowsPerThread || S >= 16384 || N >= 16384 || N % 4 != 0) return 0;
    if (ensure_tables(g, kRK)) return 1;
    RegSmem rs = {0, 0, 0, 0, 0};
    const SellHost &tf = g->t_fwd[kRK], &tb = g->t_bwd[kRK], &tp = g->t_pdf[kRK];
    for (int c = 0; c < kRK; ++c) {
        int of = 0, ob = 0, ga = 0;
        if (tf.part_slice[c + 1] - tf.part_slice[c] > kRW * kRowsPerThread || tb.part_slice[c + 1] - tb.part_slice[c] > kRW * kRowsPerThread) return 0;
        for (int i = tf.part_slice[c]; i < tf.part_slice[c + 1]; ++i) of += std::max(tf.slice_len[i] - kRegArcs, 0) * 32;
        for (int i = tb.part_slice[c]; i < tb.part_slice[c + 1]; ++i) ob += std::max(tb.slice_len[i] - kRegArcs, 0) * 32;
        for (int i = tp.part_slice[c]; i < tp.part_slice[c + 1]; ++i) ga += tp.slice_len[i] * 32;
        rs.ovf_f = std::max(rs.ovf_f, of); rs.ovf_b = std::max(rs.ovf_b, ob); rs.garcs = std::max(rs.garcs, ga);
        rs.nslg = std::max(rs.nslg, tp.part_slice[c + 1] - tp.part_slice[c]);
        rs.gcap = std::max(rs.gcap, tp.part_row[c + 1] - tp.part_row[c]);
    }
    rs.ovf_f = (rs.ovf_f + 1) & ~1; rs.ovf_b = (rs.ovf_b + 1) & ~1;
    rs.nslg = (rs.nslg + 1) & ~1;
    rs.gcap = (rs.gcap + 3) & ~3;
    RegSmemHost* h = new RegSmemHost();
    h->rs = rs;
    h->smem_f = reg_fwd_smem_bytes(S, N, rs);
    h->smem_b = reg_bwd_smem_bytes(S, N, rs);
    h->smem_f2 = reg_fwd2_smem_bytes(S, N, rs);
    h->two_slot = h->smem_f2 <= 227 * 1024;
    if (h->two_slot)
        PK2_CHECK(cudaFuncSetAttribute(den_forward_reg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_f2));
    if (h->smem_f > 227 * 1024 || h->smem_b > 227 * 1024) { delete h; return 0; }
    PK2_CHECK(cudaFuncSetAttribute(den_forward_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_f));
    PK2_CHECK(cudaFuncSetAttribute(den_backward_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_b));
    // how many clusters of 8 can be resident at once (GPC granularity: fewer than SMs / 8)
    int nmax = 1 << 30;
    for (int pass = 0; pass < 2; ++pass) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kRK * 64); cfg.blockDim = dim3(kRT);
        cfg.dynamicSmemBytes = pass ? h->smem_b : h->smem_f;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kRK; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        PK2_CHECK(cudaOccupancyMaxActiveClusters(&n, pass ? den_backward_reg_kernel : den_forward_reg_kernel, &cfg));
        nmax = std::min(nmax, n);
    }
    if (nmax < 1) { delete h; return 0; }
    h->max_clusters = nmax;
    g->reg = h;
    g->reg_state = 1;
    return 0;
}

// Work lists: longest sequence first onto the least loaded cluster (LPT), then local search (move / swap
// between the most loaded cluster and the others) -- with ~3 sequences per cluster plain LPT leaves the
// largest load 5-10 % above the mean, the refinement ~2 %.  Returns the largest load (frames).
long long plan_work(const int32_t* frames, std::vector<int> ids, int ncl, std::vector<int32_t>* work) {
    if (frames) std::stable_sort(ids.begin(), ids.end(), [&](int x, int y) { return frames[x] > frames[y]; });
    auto cost = [&](int i) { return (long long)(frames ? frames[i] : 1) + 8; };      // + per-sequence set-up
    std::vector<std::vector<int>> lists(ncl);
    std::vector<long long> load(ncl, 0);
    for (int i : ids) {
        int best = 0;
        for (int k = 1; k < ncl; ++k) if (load[k] < load[best]) best = k;
        lists[best].push_back(i);
        load[best] += cost(i);
    }
    for (int it = 0; it < 1000; ++it) {
        const int m = (int)(std::max_element(load.begin(), load.end()) - load.begin());
        bool improved = false;
        for (size_t x = 0; x < lists[m].size() && !improved; ++x) {
            const int i = lists[m][x];
            for (int k = 0; k < ncl && !improved; ++k) {
                if (k == m) continue;
                if (load[k] + cost(i) < load[m]) {                      // move i to cluster k
                    lists[k].push_back(i); lists[m].erase(lists[m].begin() + x);
                    load[k] += cost(i); load[m] -= cost(i);
                    improved = true;
                    break;
                }
                for (size_t y = 0; y < lists[k].size(); ++y) {          // swap i with a shorter sequence of k
                    const int j = lists[k][y];
                    const long long d = cost(i) - cost(j);
                    if (d > 0 && load[k] + d < load[m]) {
                        lists[m][x] = j; lists[k][y] = i;
                        load[m] -= d; load[k] += d;
                        improved = true;
                        break;
                    }
                }
            }
        }
        if (!improved) break;
    }
    if (work) {
        work->assign(ncl + 1, 0);
        for (int k = 0; k < ncl; ++k) {
            if (frames) std::stable_sort(lists[k].begin(), lists[k].end(), [&](int x, int y) { return frames[x] > frames[y]; });
            (*work)[k + 1] = (*work)[k] + (int)lists[k].size();
            for (int i : lists[k]) work->push_back(i);
        }
    }
    return *std::max_element(load.begin(), load.end());
}

// Hybrid schedule: the clusters of 8 leave `spare` SMs unused (GPC granularity).  The n1 shortest sequences
// run there as single-CTA streaming kernels (kSingleCost x the per-frame time of a cluster, but on one SM
// instead of eight); n1 minimises max(time of the longest single sequence, time of the cluster pool).
constexpr double kSingleCost = 5.6;     // per-frame time of den_fb1_kernel / per-frame time of a cluster of 8
void plan_hybrid(const int32_t* frames, int n, int ncl, int spare, std::vector<int>* pool, std::vector<int>* single) {
    std::vector<int> asc(n);
    for (int i = 0; i < n; ++i) asc[i] = i;
    std::stable_sort(asc.begin(), asc.end(), [&](int x, int y) { return frames[x] < frames[y]; });
    int best = 0;
    double best_t = (double)plan_work(frames, asc, std::min(ncl, n), nullptr);
    for (int n1 = 1; n1 <= spare && n1 < n; ++n1) {
        std::vector<int> rest(asc.begin() + n1, asc.end());
        const double tp = (double)plan_work(frames, rest, std::min(ncl, (int)rest.size()), nullptr);
        const double t = std::max(tp, ((double)frames[asc[n1 - 1]] + 8.0) * kSingleCost);
        if (t < best_t) { best_t = t; best = n1; }
    }
    single->assign(asc.begin(), asc.begin() + best);
    pool->assign(asc.begin() + best, asc.end());
}

int ensure_side(DenGraph* g) {
    if (!g->side[0]) {
        for (int i = 0; i < 2; ++i) {
            PK2_CHECK(cudaStreamCreateWithFlags(&g->side[i], cudaStreamNonBlocking));
            PK2_CHECK(cudaEventCreateWithFlags(&g->ev_join[i], cudaEventDisableTiming));
        }
        PK2_CHECK(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
    }
    return 0;
}

}  // namespace

extern "C" int pk2_den_graph_create(int S, int N, const int32_t* fwd_off, const float* fwd_prob,
                                    const int32_t* fwd_pdf, const int32_t* fwd_state,
                                    const float* init_h, void** graph) {
    PK2_REQUIRE(fwd_off && fwd_prob && fwd_pdf && fwd_state && init_h && graph, "pk2_den_graph_create: null argument");
    PK2_REQUIRE(S > 0 && S <= 65535 && N > 0 && N <= 65535,
                "pk2_den_graph_create: num_states=%d / num_pdfs=%d outside the 16-bit packed index range", S, N);
    DenGraph* g = new DenGraph();
    g->S = S; g->N = N; g->A = fwd_off[S];
    g->rows_fwd.resize(S); g->rows_bwd.resize(S); g->rows_pdf.resize(N);
    for (int i = 0; i < S; ++i) {
        for (int k = fwd_off[i]; k < fwd_off[i + 1]; ++k) {
            const int j = fwd_state[k], p = fwd_pdf[k];
            const float w = fwd_prob[k];
            if (j < 0 || j >= S || p < 0 || p >= N) {
                delete g;
                pk2::set_error("pk2_den_graph_create: arc %d out of range (dst %d pdf %d)", k, j, p);
                return 2;
            }
            g->rows_fwd[j].push_back({w, i, p});   // alpha pass: gather alpha'[src], e[pdf]
            g->rows_bwd[i].push_back({w, j, p});   // beta pass:  gather beta[dst],  e[pdf]
            g->rows_pdf[p].push_back({w, i, j});   // gamma pass: gather alpha'[src], beta[dst]
        }
    }
    double isum = 0.0;
    for (int i = 0; i < S; ++i) isum += init_h[i];
    g->init_sum = (float)isum;
    PK2_CHECK(cudaMalloc(&g->init, sizeof(float) * S));
    PK2_CHECK(cudaMemcpy(g->init, init_h, sizeof(float) * S, cudaMemcpyHostToDevice));
    *graph = g;
    return 0;
}

// Host-side schedule of one pk2_denfb call, exported for the CPU tests (no device work): assign[i] = cluster
// (0 .. n_clusters-1) that processes sequence i, or -1 if it runs as a single-CTA kernel on a spare SM;
// returns the largest cluster load in frames (incl. the per-sequence set-up allowance) or -1 on bad arguments.
extern "C" long long pk2_den_plan(const int32_t* frames, int n_seq, int n_clusters, int spare_sms, int32_t* assign) {
    if (!frames || !assign || n_seq <= 0 || n_clusters <= 0) return -1;
    std::vector<int> pool, single;
    if (spare_sms > 0 && n_seq > n_clusters) plan_hybrid(frames, n_seq, n_clusters, spare_sms, &pool, &single);
    else { pool.resize(n_seq); for (int i = 0; i < n_seq; ++i) pool[i] = i; }
    const int ncl = std::min((int)pool.size(), n_clusters);
    std::vector<int32_t> work;
    const long long worst = plan_work(frames, pool, ncl, &work);
    for (int i = 0; i < n_seq; ++i) assign[i] = -2;
    for (int k = 0; k < ncl; ++k)
        for (int j = work[k]; j < work[k + 1]; ++j) assign[work[ncl + 1 + j]] = k;
    for (int i : single) assign[i] = -1;
    return worst;
}

extern "C" int pk2_den_set_sm_budget(void* graph, int max_clusters, int reserve_sms) {
    PK2_REQUIRE(graph && max_clusters >= 0 && reserve_sms >= 0, "pk2_den_set_sm_budget: bad argument");
    DenGraph* g = static_cast<DenGraph*>(graph);
    std::lock_guard<std::mutex> launch_lock(g->launch_mu);
    g->budget_clusters = max_clusters;
    g->budget_reserve = reserve_sms;
    return 0;
}

extern "C" int pk2_den_set_profile_buffer(void* buf) {
    g_den_prof = static_cast<long long*>(buf);
    return 0;
}

extern "C" int pk2_den_graph_destroy(void* graph) {
    if (!graph) return 0;
    DenGraph* g = static_cast<DenGraph*>(graph);
    for (int k = 1; k <= kMaxParts; ++k)
        if (g->built[k]) { g->t_fwd[k].free_dev(); g->t_bwd[k].free_dev(); g->t_pdf[k].free_dev(); }
    cudaFree(g->init);
    cudaFree(g->start_flag);
    delete g->reg;
    delete g;
    return 0;
}

extern "C" size_t pk2_denfb_workspace_bytes(void* graph, int n_seq, int max_frames) {
    if (!graph || n_seq <= 0 || max_frames <= 0) return 0;
    DenGraph* g = static_cast<DenGraph*>(graph);
    size_t alpha = (size_t)n_seq * (size_t)max_frames * (size_t)g->S * sizeof(float);
    size_t asum = (size_t)n_seq * (size_t)(max_frames + 2) * sizeof(float);
    size_t maps = ((size_t)n_seq + 256) * sizeof(int32_t);      // sequence map / work lists
    size_t ebuf = (size_t)n_seq * (size_t)max_frames * (size_t)g->N * sizeof(float);   // exp(loglikes), register-resident path
    return ((alpha + 255) & ~(size_t)255) + ((asum + 255) & ~(size_t)255) + ((maps + 255) & ~(size_t)255) + ebuf;
}

extern "C" int pk2_denfb(void* graph, const float* loglikes, const int32_t* num_frames,
                         const int32_t* num_frames_h, int n_seq, int max_frames, int64_t row_stride_b,
                         float leaky, float deriv_scale, void* workspace, float* grad, double* logz,
                         int cluster, void* stream) {
    PK2_REQUIRE(graph && loglikes && num_frames && workspace && grad && logz, "pk2_denfb: null argument");
    PK2_REQUIRE(n_seq > 0 && max_frames > 0, "pk2_denfb: empty batch");
    PK2_REQUIRE(row_stride_b >= max_frames, "pk2_denfb: row_stride_b < max_frames");
    PK2_REQUIRE(cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8,
                "pk2_denfb: cluster must be 0, 1, 2, 4 or 8 (got %d)", cluster);
    DenGraph* g = static_cast<DenGraph*>(graph);
    cudaStream_t st = pk2::as_stream(stream);
    DenArgs a;
    a.init = g->init; a.S = g->S; a.N = g->N;
    a.ll = loglikes; a.num_frames = num_frames; a.row_stride_b = row_stride_b; a.max_frames = max_frames;
    a.leaky = leaky; a.deriv_scale = deriv_scale;
    size_t alpha = (size_t)n_seq * (size_t)max_frames * (size_t)g->S * sizeof(float);
    alpha = (alpha + 255) & ~(size_t)255;
    size_t asum = (size_t)n_seq * (size_t)(max_frames + 2) * sizeof(float);
    asum = (asum + 255) & ~(size_t)255;
    a.alpha_ws = static_cast<float*>(workspace);
    a.asum_ws = reinterpret_cast<float*>(static_cast<char*>(workspace) + alpha);
    int32_t* maps_dev = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + alpha + asum);
    const size_t maps_bytes = ((((size_t)n_seq + 256) * sizeof(int32_t)) + 255) & ~(size_t)255;
    float* e_dev = reinterpret_cast<float*>(static_cast<char*>(workspace) + alpha + asum + maps_bytes);
    a.grad = grad; a.logz = logz; a.seq_map = nullptr; a.work = nullptr; a.work_ids = 0; a.e = nullptr; a.start_flag = nullptr; a.epoch = 0;
    { const char* e = getenv("PK2_DEN_DEBUG"); a.debug = e ? atoi(e) : 0; }
    a.prof = g_den_prof;

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);

    std::lock_guard<std::mutex> launch_lock(g->launch_mu);
    // clusters of 8 with register-resident arcs (default where the graph fits, see plan_reg)
    {
        static const bool reg_off = []() { const char* e = getenv("PK2_DEN_REG"); return e && atoi(e) == 0; }();
        if (cluster == 8 || (cluster == 0 && !reg_off)) {
            if (plan_reg(g)) return 1;
            if (g->reg_state == 1) {
                { static bool said = false; if (!said && getenv("PK2_DEN_VERBOSE")) { said = true;
                    fprintf(stderr, "pk2_denfb: register-resident path, %d resident clusters of 8 (%d SMs), smem fwd %zu bwd %zu B\n",
                            g->reg->max_clusters, sms, g->reg->smem_f, g->reg->smem_b); } }
                static const bool hyb_off = []() { const char* e = getenv("PK2_DEN_HYBRID"); return e && atoi(e) == 0; }();
                std::vector<int> pool, single;
                // 8 of the spare SMs are left to the numerator kernel that runs next to this call (ops.py)
                int max_cl = g->reg->max_clusters;
                if (g->budget_clusters > 0) max_cl = std::min(max_cl, g->budget_clusters);
                const int spare = sms - kRK * max_cl - 8 - g->budget_reserve;
                if (cluster == 0 && !hyb_off && num_frames_h && spare > 0 && n_seq > max_cl &&
                    fwd_smem_bytes(g->S, g->N, 1) <= 227 * 1024 && bwd_smem_bytes(g->S, g->N, 1) <= 227 * 1024) {
                    plan_hybrid(num_frames_h, n_seq, max_cl, spare, &pool, &single);
                } else {
                    pool.resize(n_seq);
                    for (int i = 0; i < n_seq; ++i) pool[i] = i;
                }
                const int ncl = std::min((int)pool.size(), max_cl);
                std::vector<int32_t> work;
                plan_work(num_frames_h, pool, ncl, &work);
                if (getenv("PK2_DEN_VERBOSE")) {
                    static int said2 = 0;
                    if (said2++ < 2) fprintf(stderr, "pk2_denfb: %d sequences on %d clusters, %d single-CTA sequences\n",
                                             (int)pool.size(), ncl, (int)single.size());
                }
                const int single_off = (int)work.size();
                for (int i : single) work.push_back(i);
                PK2_CHECK(cudaMemcpyAsync(maps_dev, work.data(), sizeof(int32_t) * work.size(), cudaMemcpyHostToDevice, st));
                a.fwd = g->t_fwd[kRK].dev(); a.bwd = g->t_bwd[kRK].dev(); a.pdf = g->t_pdf[kRK].dev();
                a.work = maps_dev; a.work_ids = ncl + 1;
                a.e = e_dev;
                if (!single.empty()) {
                    if (ensure_side(g) || ensure_tables(g, 1)) return 1;
                    if (!g->start_flag) {
                        PK2_CHECK(cudaMalloc(&g->start_flag, sizeof(int)));
                        PK2_CHECK(cudaMemset(g->start_flag, 0, sizeof(int)));
                    }
                    a.start_flag = g->start_flag; a.epoch = ++g->epoch;
                    PK2_CHECK(cudaEventRecord(g->ev_fork, st));          // work lists are on the device
                }
                den_exp_kernel<<<dim3(max_frames, n_seq), 256, 0, st>>>(loglikes, e_dev, num_frames, row_stride_b, max_frames, g->N);
                PK2_POST_LAUNCH();
                static const bool two_off = []() { const char* e = getenv("PK2_DEN_FWD2"); return e && atoi(e) == 0; }();
                if (g->reg->two_slot && !two_off) {
                    if (launch_cluster8(den_forward_reg2_kernel, a, g->reg->rs, ncl, g->reg->smem_f2, st)) return 1;
                } else if (launch_cluster8(den_forward_reg_kernel, a, g->reg->rs, ncl, g->reg->smem_f, st)) return 1;
                if (launch_cluster8(den_backward_reg_kernel, a, g->reg->rs, ncl, g->reg->smem_b, st)) return 1;
                if (!single.empty()) {
                    // single-CTA sequences: behind a gate that opens when the forward cluster kernel has started
                    cudaStream_t ss = g->side[0];
                    PK2_CHECK(cudaStreamWaitEvent(ss, g->ev_fork, 0));
                    den_gate_kernel<<<1, 1, 0, ss>>>(g->start_flag, a.epoch);
                    PK2_POST_LAUNCH();
                    DenArgs a1 = a;
                    a1.fwd = g->t_fwd[1].dev(); a1.bwd = g->t_bwd[1].dev(); a1.pdf = g->t_pdf[1].dev();
                    a1.seq_map = maps_dev + single_off;
                    a1.work = nullptr; a1.start_flag = nullptr;
                    const size_t sm1 = std::max(fwd_smem_bytes(g->S, g->N, 1), bwd_smem_bytes(g->S, g->N, 1));
                    PK2_CHECK(cudaFuncSetAttribute(den_fb1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
                    den_fb1_kernel<<<(int)single.size(), kThreads, sm1, ss>>>(a1);
                    PK2_POST_LAUNCH();
                    PK2_CHECK(cudaEventRecord(g->ev_join[0], ss));
                    PK2_CHECK(cudaStreamWaitEvent(st, g->ev_join[0], 0));
                }
                return 0;
            }
            PK2_REQUIRE(cluster != 8, "pk2_denfb: cluster = 8 needs num_states %% 256 == 0, num_states <= 8192, num_pdfs %% 4 == 0 and a graph "
                        "whose per-CTA tables fit shared memory (S=%d N=%d)", g->S, g->N);
        }
    }
    if (g->S % 4 != 0) {            // the DSMEM row exchange moves 16-byte multiples: single-CTA clusters only
        PK2_REQUIRE(cluster == 0 || cluster == 1, "pk2_denfb: num_states %% 4 != 0 supports cluster = 1 only");
        cluster = 1;
    }
    if (cluster != 0 || num_frames_h == nullptr || n_seq > sms) {
        // uniform cluster size
        int K = cluster;
        if (K == 0) K = (n_seq * 4 <= sms) ? 4 : ((n_seq * 2 <= sms) ? 2 : 1);
        if (ensure_tables(g, K)) return 1;
        a.fwd = g->t_fwd[K].dev(); a.bwd = g->t_bwd[K].dev(); a.pdf = g->t_pdf[K].dev();
        return launch_den_k(K, a, n_seq, st);
    }

    // length-aware mixed schedule: up to three concurrent launches (K = 4, 2, 1), longest first
    std::vector<int> ks;
    plan_clusters(num_frames_h, n_seq, sms, &ks);
    std::vector<int32_t> order;
    int count[5] = {0, 0, 0, 0, 0};
    for (int K : {4, 2, 1}) {
        std::vector<int> idx;
        for (int i = 0; i < n_seq; ++i) if (ks[i] == K) idx.push_back(i);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return num_frames_h[x] > num_frames_h[y]; });
        for (int i : idx) order.push_back(i);
        count[K] = (int)idx.size();
    }
    PK2_CHECK(cudaMemcpyAsync(maps_dev, order.data(), sizeof(int32_t) * n_seq, cudaMemcpyHostToDevice, st));
    if (ensure_side(g)) return 1;
    PK2_CHECK(cudaEventRecord(g->ev_fork, st));
    int off = 0, lane = 0;
    for (int K : {4, 2, 1}) {
        if (count[K] == 0) continue;
        if (ensure_tables(g, K)) return 1;
        DenArgs ak = a;
        ak.fwd = g->t_fwd[K].dev(); ak.bwd = g->t_bwd[K].dev(); ak.pdf = g->t_pdf[K].dev();
        ak.seq_map = maps_dev + off;
        cudaStream_t s = st;
        if (lane > 0) {
            s = g->side[lane - 1];
            PK2_CHECK(cudaStreamWaitEvent(s, g->ev_fork, 0));
        }
        if (launch_den_k(K, ak, count[K], s)) return 1;
        if (lane > 0) {
            PK2_CHECK(cudaEventRecord(g->ev_join[lane - 1], s));
            PK2_CHECK(cudaStreamWaitEvent(st, g->ev_join[lane - 1], 0));
        }
        off += count[K];
        ++lane;
    }
    return 0;
}
