// 80-dim log-mel fbank on the GPU (sm_100a).
//
// Replaces DataGeneratorTrain._logfbank_extractor (reference data/sr_dataset.py:279-296)
// and stft/_enframe (simulation/freq_analysis.py:41-150): pre-emphasis 0.96 (sample 0
// dropped), 400-sample frames / hop 160 / no centering / last partial frame zero padded,
// symmetric Hamming(400), 512-pt real FFT, |.|^2, mel[257,80]*32768^2, +1, log.
// One warp per frame: the 512-pt real FFT is a 256-pt complex radix-2 FFT in shared
// memory plus the split post-pass; the mel matrix is stored compacted per filter
// (483 non-zeros).  HBM-bound: 4 B/sample in, 320 B/frame out.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace {

constexpr int kFrameLen = 400;
constexpr int kHop = 160;
constexpr int kNfft = 512;
constexpr int kBins = 257;
constexpr int kMel = 80;
constexpr int kWarps = 4;

struct FbankPlan {
    float* hamming;      // [400]
    float2* tw;          // [257] exp(-2 pi i k / 512)
    int* fstart;         // [80] first bin of filter f
    int* flen;           // [80]
    int* foff;           // [80] offset into fw
    float* fw;           // compacted weights (already * 32768^2)
    int max_len;
};

struct cplx { float x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cplx cshfl(cplx a, int lane) {
    return {__shfl_sync(0xffffffffu, a.x, lane), __shfl_sync(0xffffffffu, a.y, lane)};
}
__device__ __forceinline__ cplx cshfl_xor(cplx a, int m) {
    return {__shfl_xor_sync(0xffffffffu, a.x, m), __shfl_xor_sync(0xffffffffu, a.y, m)};
}

// 8-point DFT in registers (decimation in frequency, three radix-2 stages); a[k] = sum_n a[n] W8^(nk)
__device__ __forceinline__ void dft8(cplx (&a)[8]) {
    const float h = 0.70710678118654752f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const cplx u = cadd(a[i], a[i + 4]), d = csub(a[i], a[i + 4]);
        a[i] = u;
        if (i == 0) a[4] = d;
        else if (i == 1) a[5] = {h * (d.x + d.y), h * (d.y - d.x)};          // * (1 - i)/sqrt2
        else if (i == 2) a[6] = {d.y, -d.x};                                 // * -i
        else a[7] = {h * (d.y - d.x), -h * (d.x + d.y)};                     // * (-1 - i)/sqrt2
    }
#pragma unroll
    for (int base = 0; base < 8; base += 4) {
        const cplx u0 = cadd(a[base], a[base + 2]), d0 = csub(a[base], a[base + 2]);
        const cplx u1 = cadd(a[base + 1], a[base + 3]), d1 = csub(a[base + 1], a[base + 3]);
        a[base] = u0; a[base + 1] = u1; a[base + 2] = d0; a[base + 3] = {d1.y, -d1.x};
    }
#pragma unroll
    for (int base = 0; base < 8; base += 2) {
        const cplx u = cadd(a[base], a[base + 1]), d = csub(a[base], a[base + 1]);
        a[base] = u; a[base + 1] = d;
    }
    // bit-reversed -> natural order (register renaming)
    cplx t = a[1]; a[1] = a[4]; a[4] = t;
    t = a[3]; a[3] = a[6]; a[6] = t;
}

// One warp per frame, the FFT in registers: the 512-pt real FFT is a 256-pt complex FFT z[n] = x[2n] + i x[2n+1],
// n = 32 n1 + lane: an 8-point DFT over n1 in the registers of each lane, the twiddle W256^(lane k1), and a 32-point
// DFT ACROSS the lanes (five radix-2 decimation-in-frequency stages on warp shuffles); lane l then holds
// Z[k1 + 8 bitrev5(l)], k1 = 0..7.  The split post-pass needs Z[256 - k], which sits in register 8 - k1 of lane
// 31 - l (one shuffle).  Shared memory is only used for the 257 power values the mel filters read (padded
// index k + k/32: conflict-free).  The first version kept the FFT in shared memory (8 stages x 4 butterflies x
// 10 accesses per lane) and reached 3.6 % of the HBM roofline (profiles/README_r1.md).
__global__ void __launch_bounds__(kWarps * 32)
fbank_kernel(const float* __restrict__ wav, const int64_t* __restrict__ wav_off,
             const int32_t* __restrict__ frame_off, int n_utts, int total_frames,
             FbankPlan plan, float* __restrict__ out) {
    __shared__ float s_p[kWarps][kBins + 12];
    __shared__ float2 s_ham[kFrameLen / 2];
    __shared__ float2 s_tw[kBins];          // W512^k
    __shared__ float2 s_tw256[256];         // W256^m

    for (int i = threadIdx.x; i < kFrameLen / 2; i += blockDim.x) s_ham[i] = make_float2(plan.hamming[2 * i], plan.hamming[2 * i + 1]);
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_tw[i] = plan.tw[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_tw256[i] = (i <= 128) ? plan.tw[2 * i] : make_float2(plan.tw[512 - 2 * i].x, -plan.tw[512 - 2 * i].y);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* pw = s_p[warp];
    // twiddles of the five cross-lane stages: W_{2 span}^(lane mod span), span = 16, 8, 4, 2, 1
    cplx stw[5];
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int span = 16 >> s;
        float sn, cs;
        sincospif(-(float)(lane & (span - 1)) / (float)span, &sn, &cs);
        stw[s] = {cs, sn};
    }
    const int k2 = (int)(__brev((unsigned)lane) >> 27);                       // lane l ends up with k = k1 + 8 k2
    const int lane_k0 = (int)(__brev((unsigned)((32 - k2) & 31)) >> 27);      // lane that holds Z[256 - 8 k2]

    for (int fr = blockIdx.x * kWarps + warp; fr < total_frames; fr += gridDim.x * kWarps) {
        // utterance of this frame: largest u with frame_off[u] <= fr
        int lo = 0, hi = n_utts;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(&frame_off[mid]) <= fr) lo = mid; else hi = mid;
        }
        const int u = lo;
        const int t = fr - __ldg(&frame_off[u]);
        const int64_t w0 = __ldg(&wav_off[u]);
        const int64_t n = __ldg(&wav_off[u + 1]) - w0;      // samples in the utterance
        const float* x = wav + w0;
        const int64_t base = (int64_t)t * kHop;

        // pre-emphasis + window: z[32 n1 + lane] = (v[2n], v[2n+1]), v[j] = (x[j+1] - .96 x[j]) * hamming[j]
        cplx a[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            const int j = 64 * n1 + 2 * lane;
            a[n1] = {0.f, 0.f};
            if (j < kFrameLen) {
                const int64_t m = base + j;                 // index into wav[1:] - .96 wav[:-1]
                const float2 hw = s_ham[j >> 1];
                if (m + 1 < n - 1) {
                    const float x0 = __ldg(&x[m]), x1 = __ldg(&x[m + 1]), x2 = __ldg(&x[m + 2]);
                    a[n1].x = __fmul_rn(__fsub_rn(x1, __fmul_rn(0.96f, x0)), hw.x);
                    a[n1].y = __fmul_rn(__fsub_rn(x2, __fmul_rn(0.96f, x1)), hw.y);
                } else if (m < n - 1) {
                    a[n1].x = __fmul_rn(__fsub_rn(__ldg(&x[m + 1]), __fmul_rn(0.96f, __ldg(&x[m]))), hw.x);
                }
            }
        }
        dft8(a);
#pragma unroll
        for (int k1 = 1; k1 < 8; ++k1) {
            const float2 w = s_tw256[(lane * k1) & 255];
            a[k1] = cmul(a[k1], cplx{w.x, w.y});
        }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int span = 16 >> s;
            const bool upper = (lane & span) == 0;
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                const cplx b = cshfl_xor(a[k1], span);
                a[k1] = upper ? cadd(a[k1], b) : cmul(csub(b, a[k1]), stw[s]);
            }
        }
        // split post-pass -> power spectrum of the 512-pt real FFT
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const int k = k1 + 8 * k2;
            const cplx zm = (k1 == 0) ? cshfl(a[0], lane_k0) : cshfl(a[8 - k1], 31 - lane);   // Z[256 - k]
            const cplx z = a[k1];
            const float er = 0.5f * (z.x + zm.x), ei = 0.5f * (z.y - zm.y);      // (Z + conj Zm) / 2
            const float dr = 0.5f * (z.x - zm.x), di = 0.5f * (z.y + zm.y);      // (Z - conj Zm) / 2
            const float orr = di, oi = -dr;                                       // / i
            const float2 w = s_tw[k];
            const float xr = er + orr * w.x - oi * w.y;
            const float xi = ei + orr * w.y + oi * w.x;
            pw[k + (k >> 5)] = xr * xr + xi * xi;
            if (k == 0) pw[256 + 8] = (z.x - z.y) * (z.x - z.y);                  // bin 256
        }
        __syncwarp();
        // mel + log
        for (int f = lane; f < kMel; f += 32) {
            const int st = plan.fstart[f], ln = plan.flen[f];
            const float* wv = plan.fw + plan.foff[f];
            float acc = 0.f;
            for (int k = 0; k < ln; ++k) { const int b = st + k; acc = fmaf(pw[b + (b >> 5)], __ldg(&wv[k]), acc); }
            out[(int64_t)fr * kMel + f] = logf(acc + 1.0f);
        }
        __syncwarp();
    }
}

// ---- per-utterance column means (fixed-order tree -> deterministic) ----
__global__ void colmean_direct_kernel(const float* __restrict__ feats,
                                      const int32_t* __restrict__ frame_off, int dim,
                                      float* __restrict__ mean) {
    // one block per utterance; threads = dim lanes x frame groups, fixed-order tree
    extern __shared__ float sm[];
    const int u = blockIdx.x;
    const int f0 = frame_off[u], f1 = frame_off[u + 1];
    const int groups = blockDim.x / dim;       // blockDim.x is a multiple of dim
    const int d = threadIdx.x % dim, g = threadIdx.x / dim;
    float acc = 0.f;
    if (g < groups)
        for (int f = f0 + g; f < f1; f += groups) acc += feats[(int64_t)f * dim + d];
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0) {
        double tot = 0.0;
        for (int k = 0; k < groups; ++k) tot += (double)sm[k * dim + d];
        mean[(int64_t)u * dim + d] = (f1 > f0) ? (float)(tot / (double)(f1 - f0)) : 0.f;
    }
}

__global__ void gather_norm_kernel(const float* __restrict__ feats,
                                   const int32_t* __restrict__ row_src,
                                   const int32_t* __restrict__ row_utt,
                                   const float* __restrict__ mean,
                                   const float* __restrict__ mvn_mean,
                                   const float* __restrict__ mvn_istd,
                                   int n_rows, int dim, float* __restrict__ out) {
    const int64_t total = (int64_t)n_rows * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / dim), d = (int)(i % dim);
        const int s = __ldg(&row_src[r]);
        float v = 0.f;
        if (s >= 0) {
            v = feats[(int64_t)s * dim + d];
            if (mean) v -= __ldg(&mean[(int64_t)__ldg(&row_utt[r]) * dim + d]);
            if (mvn_mean) v = (v - __ldg(&mvn_mean[d])) * __ldg(&mvn_istd[d]);
        }
        out[i] = v;
    }
}

}  // namespace

extern "C" int pk2_fbank_plan_create(const float* mel_h, void** plan_out) {
    PK2_REQUIRE(mel_h && plan_out, "pk2_fbank_plan_create: null argument");
    std::vector<float> ham(kFrameLen);
    for (int i = 0; i < kFrameLen; ++i)
        ham[i] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * (double)i / (double)(kFrameLen - 1)));
    std::vector<float2> tw(kBins);
    for (int k = 0; k < kBins; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)kNfft;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<int> fstart(kMel), flen(kMel), foff(kMel);
    std::vector<float> fw;
    int max_len = 0;
    const float scale = 32768.0f * 32768.0f;
    for (int f = 0; f < kMel; ++f) {
        int lo = kBins, hi = -1;
        for (int k = 0; k < kBins; ++k)
            if (mel_h[k * kMel + f] != 0.f) { lo = lo < k ? lo : k; hi = k; }
        fstart[f] = hi < 0 ? 0 : lo;
        flen[f] = hi < 0 ? 0 : hi - lo + 1;
        foff[f] = (int)fw.size();
        for (int k = 0; k < flen[f]; ++k) fw.push_back(mel_h[(fstart[f] + k) * kMel + f] * scale);
        max_len = max_len > flen[f] ? max_len : flen[f];
    }
    if (fw.empty()) fw.push_back(0.f);
    FbankPlan* p = new FbankPlan();
    p->max_len = max_len;
    PK2_CHECK(cudaMalloc(&p->hamming, sizeof(float) * kFrameLen));
    PK2_CHECK(cudaMalloc(&p->tw, sizeof(float2) * kBins));
    PK2_CHECK(cudaMalloc(&p->fstart, sizeof(int) * kMel));
    PK2_CHECK(cudaMalloc(&p->flen, sizeof(int) * kMel));
    PK2_CHECK(cudaMalloc(&p->foff, sizeof(int) * kMel));
    PK2_CHECK(cudaMalloc(&p->fw, sizeof(float) * fw.size()));
    PK2_CHECK(cudaMemcpy(p->hamming, ham.data(), sizeof(float) * kFrameLen, cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(p->tw, tw.data(), sizeof(float2) * kBins, cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(p->fstart, fstart.data(), sizeof(int) * kMel, cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(p->flen, flen.data(), sizeof(int) * kMel, cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(p->foff, foff.data(), sizeof(int) * kMel, cudaMemcpyHostToDevice));
    PK2_CHECK(cudaMemcpy(p->fw, fw.data(), sizeof(float) * fw.size(), cudaMemcpyHostToDevice));
    *plan_out = p;
    return 0;
}

extern "C" int pk2_fbank_plan_destroy(void* plan) {
    if (!plan) return 0;
    FbankPlan* p = static_cast<FbankPlan*>(plan);
    cudaFree(p->hamming); cudaFree(p->tw); cudaFree(p->fstart); cudaFree(p->flen);
    cudaFree(p->foff); cudaFree(p->fw);
    delete p;
    return 0;
}

extern "C" int pk2_fbank(void* plan, const float* wav, const int64_t* wav_off,
                         const int32_t* frame_off, int n_utts, int total_frames, float* out,
                         void* stream) {
    PK2_REQUIRE(plan && wav && wav_off && frame_off && out, "pk2_fbank: null argument");
    if (total_frames <= 0 || n_utts <= 0) return 0;
    FbankPlan* p = static_cast<FbankPlan*>(plan);
    int blocks = (total_frames + kWarps - 1) / kWarps;
    const int cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    fbank_kernel<<<blocks, kWarps * 32, 0, pk2::as_stream(stream)>>>(
        wav, wav_off, frame_off, n_utts, total_frames, *p, out);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_colmean(const float* feats, const int32_t* frame_off, int n_utts, int dim,
                           float* mean, void* stream) {
    PK2_REQUIRE(feats && frame_off && mean, "pk2_colmean: null argument");
    PK2_REQUIRE(dim > 0 && dim <= 256, "pk2_colmean: dim %d out of range", dim);
    if (n_utts <= 0) return 0;
    int threads = (1024 / dim) * dim;
    colmean_direct_kernel<<<n_utts, threads, threads * sizeof(float), pk2::as_stream(stream)>>>(
        feats, frame_off, dim, mean);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_gather_norm(const float* feats, const int32_t* row_src, const int32_t* row_utt,
                               const float* mean, const float* mvn_mean, const float* mvn_istd,
                               int n_rows, int dim, float* out, void* stream) {
    PK2_REQUIRE(feats && row_src && out, "pk2_gather_norm: null argument");
    PK2_REQUIRE(!mean || row_utt, "pk2_gather_norm: mean given without row_utt");
    PK2_REQUIRE((mvn_mean == nullptr) == (mvn_istd == nullptr), "pk2_gather_norm: mvn_mean/istd must come together");
    if (n_rows <= 0) return 0;
    int64_t total = (int64_t)n_rows * dim;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    gather_norm_kernel<<<blocks, 256, 0, pk2::as_stream(stream)>>>(
        feats, row_src, row_utt, mean, mvn_mean, mvn_istd, n_rows, dim, out);
    PK2_POST_LAUNCH();
    return 0;
}
