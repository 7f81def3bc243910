// Fused log-softmax + NLL forward/backward over rows of logits (sm_100a).
//
// Replaces nn.CrossEntropyLoss(ignore_index=-100) (reference bin/train_ce.py:134,189;
// reduction='sum' at bin/train_se.py:214,235).  One CTA per row; the row is staged in
// shared memory once (coalesced float4 loads when n_cols % 4 == 0), reduced (max, sum exp), and the gradient
// scale*(softmax - onehot) is written back.  HBM-bound: 4N B read + 4N B written per row.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
ce_softmax_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                  int64_t n_rows, int n_cols, float scale, float* __restrict__ loss_rows,
                  float* __restrict__ grad) {
    extern __shared__ __align__(16) float row[];
    __shared__ float red[kThreads / 32];
    __shared__ float bc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = (n_cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                     (!grad || (reinterpret_cast<uintptr_t>(grad) & 15) == 0);
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const int64_t lab = labels[r];
        const float* src = logits + r * (int64_t)n_cols;
        float* dst = grad ? grad + r * (int64_t)n_cols : nullptr;
        if (lab < 0) {   // ignore_index
            if (dst) for (int c = threadIdx.x; c < n_cols; c += kThreads) dst[c] = 0.f;
            if (threadIdx.x == 0 && loss_rows) loss_rows[r] = 0.f;
            continue;
        }
        float m = -INFINITY;
        if (vec) {
            for (int c = threadIdx.x * 4; c < n_cols; c += kThreads * 4) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src + c));     // streamed once: evict first
                *reinterpret_cast<float4*>(row + c) = v;
                m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            }
        } else {
            for (int c = threadIdx.x; c < n_cols; c += kThreads) {
                const float v = src[c];
                row[c] = v;
                m = fmaxf(m, v);
            }
        }
        m = pk2::warp_max(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mm = red[0];
            for (int i = 1; i < kThreads / 32; ++i) mm = fmaxf(mm, red[i]);
            bc = mm;
        }
        __syncthreads();
        m = bc;
        float s = 0.f;
        if (vec) {
            for (int c = threadIdx.x * 4; c < n_cols; c += kThreads * 4) {       // same thread -> element mapping as the load
                float4 v = *reinterpret_cast<float4*>(row + c);
                v.x = __expf(v.x - m); v.y = __expf(v.y - m); v.z = __expf(v.z - m); v.w = __expf(v.w - m);
                *reinterpret_cast<float4*>(row + c) = v;
                s += (v.x + v.y) + (v.z + v.w);
            }
        } else {
            for (int c = threadIdx.x; c < n_cols; c += kThreads) {
                const float e = __expf(row[c] - m);
                row[c] = e;
                s += e;
            }
        }
        s = pk2::warp_sum(s);
        __syncthreads();
        if (lane == 0) red[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float ss = 0.f;
            for (int i = 0; i < kThreads / 32; ++i) ss += red[i];
            bc = ss;
        }
        __syncthreads();
        s = bc;
        if (threadIdx.x == 0 && loss_rows) loss_rows[r] = logf(s) + m - src[lab];
        if (dst) {
            const float inv = scale / s;
            if (vec) {
                for (int c = threadIdx.x * 4; c < n_cols; c += kThreads * 4) {
                    float4 v = *reinterpret_cast<float4*>(row + c);
                    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
                    const int d = (int)lab - c;
                    if (d == 0) v.x -= scale; else if (d == 1) v.y -= scale; else if (d == 2) v.z -= scale; else if (d == 3) v.w -= scale;
                    *reinterpret_cast<float4*>(dst + c) = v;
                }
            } else {
                for (int c = threadIdx.x; c < n_cols; c += kThreads)
                    dst[c] = row[c] * inv - (c == (int)lab ? scale : 0.f);
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int pk2_ce_softmax(const float* logits, const int64_t* labels, int64_t n_rows,
                              int n_cols, float scale, float* loss_rows, float* grad,
                              void* stream) {
    PK2_REQUIRE(logits && labels, "pk2_ce_softmax: null argument");
    PK2_REQUIRE(n_cols > 0 && n_cols <= 16384, "pk2_ce_softmax: n_cols %d out of range", n_cols);
    if (n_rows <= 0) return 0;
    const size_t smem = sizeof(float) * (size_t)n_cols;
    static bool attr_set = false;
    if (!attr_set) {
        PK2_CHECK(cudaFuncSetAttribute(ce_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        attr_set = true;
    }
    int64_t blocks = n_rows < 148 * 8 ? n_rows : 148 * 8;
    ce_softmax_kernel<<<(int)blocks, kThreads, smem, pk2::as_stream(stream)>>>(
        logits, labels, n_rows, n_cols, scale, loss_rows, grad);
    PK2_POST_LAUNCH();
    return 0;
}
