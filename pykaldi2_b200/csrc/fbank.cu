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

__global__ void __launch_bounds__(kWarps * 32)
fbank_kernel(const float* __restrict__ wav, const int64_t* __restrict__ wav_off,
             const int32_t* __restrict__ frame_off, int n_utts, int total_frames,
             FbankPlan plan, float* __restrict__ out) {
    __shared__ float s_re[kWarps][256];
    __shared__ float s_im[kWarps][256];
    __shared__ float s_p[kWarps][kBins + 3];
    __shared__ float s_ham[kFrameLen];
    __shared__ float2 s_tw[kBins];

    for (int i = threadIdx.x; i < kFrameLen; i += blockDim.x) s_ham[i] = plan.hamming[i];
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) s_tw[i] = plan.tw[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* re = s_re[warp];
    float* im = s_im[warp];
    float* pw = s_p[warp];

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

        // pre-emphasis + window, scattered to bit-reversed complex order
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = lane + 32 * i;
            float v = 0.f;
            if (j < kFrameLen) {
                const int64_t m = base + j;                 // index into wav[1:] - .96 wav[:-1]
                if (m < n - 1) {
                    const float a = __ldg(&x[m + 1]);
                    const float b = __ldg(&x[m]);
                    v = __fmul_rn(__fsub_rn(a, __fmul_rn(0.96f, b)), s_ham[j]);
                }
            }
            const int nn = j >> 1;
            const int br = __brev((unsigned)nn) >> 24;
            if (j & 1) im[br] = v; else re[br] = v;
        }
        __syncwarp();
        // 256-pt radix-2 DIT
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const int half = 1 << s;
            const int tstep = 256 >> s;                     // index step into the 512-pt table
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int b = lane + 32 * q;
                const int pos = b & (half - 1);
                const int i0 = ((b >> s) << (s + 1)) + pos;
                const int i1 = i0 + half;
                const float2 w = s_tw[pos * tstep];
                const float xr = re[i1], xi = im[i1];
                const float tr = xr * w.x - xi * w.y;
                const float ti = xr * w.y + xi * w.x;
                const float ur = re[i0], ui = im[i0];
                re[i0] = ur + tr; im[i0] = ui + ti;
                re[i1] = ur - tr; im[i1] = ui - ti;
            }
            __syncwarp();
        }
        // split post-pass -> power spectrum
        for (int k = lane; k < kBins; k += 32) {
            const int k0 = k & 255, k1 = (256 - k) & 255;
            const float zr = re[k0], zi = im[k0];
            const float cr = re[k1], ci = -im[k1];
            const float er = 0.5f * (zr + cr), ei = 0.5f * (zi + ci);   // Xe
            const float dr = 0.5f * (zr - cr), di = 0.5f * (zi - ci);   // (Z - conj Z')/2
            const float orr = di, oi = -dr;                              // / i  -> Xo
            const float2 w = s_tw[k];
            const float xr = er + orr * w.x - oi * w.y;
            const float xi = ei + orr * w.y + oi * w.x;
            pw[k] = xr * xr + xi * xi;
        }
        __syncwarp();
        // mel + log
        for (int f = lane; f < kMel; f += 32) {
            const int st = plan.fstart[f], ln = plan.flen[f];
            const float* wv = plan.fw + plan.foff[f];
            float acc = 0.f;
            for (int k = 0; k < ln; ++k) acc = fmaf(pw[st + k], __ldg(&wv[k]), acc);
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
