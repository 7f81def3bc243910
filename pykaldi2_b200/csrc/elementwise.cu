// Layout / precision helpers of the BLSTM path (sm_100a): fp32->bf16 cast, transposing casts that
// produce the K-major operands of the weight-gradient GEMMs, the shifted h_{t-1} operand of dW_hh,
// and the bias gradient (column sums).  All are single-pass, HBM-bound, coalesced on both sides
// (32x32 shared-memory tiles for the transposes).
#include "common.cuh"
#include <map>
#include <mutex>

namespace {

__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(src)[i];
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        reinterpret_cast<uint2*>(dst)[i] = o;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16(src[i]);
}

// dst[c, r] = src[r, c]
template <bool SRC_BF16>
__global__ void transpose_bf16_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C,
                                      int lds, int ldd) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        float v = 0.f;
        if (r < R && c < C) {
            if (SRC_BF16) v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[(int64_t)r * lds + c]);
            else v = reinterpret_cast<const float*>(src)[(int64_t)r * lds + c];
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[(int64_t)c * ldd + r] = __float2bfloat16(tile[threadIdx.x][i]);
    }
}

// hprevT[(dir*H + j), b*T + t] = y[b, tprev, dir*H + j],  tprev = t-1 (dir 0) / t+1 (dir 1), 0 outside
__global__ void hprev_t_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, int B, int T, int H,
                               int ldd) {
    __shared__ __nv_bfloat16 tile[32][34];
    const int M = B * T;
    const int c0 = blockIdx.x * 32;          // column of y (0 .. 2H)
    const int r0 = blockIdx.y * 32;          // row m = b*T + t
    const int dir = c0 / H;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int m = r0 + i, c = c0 + threadIdx.x;
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (m < M) {
            const int b = m / T, t = m - b * T;
            const int tp = dir ? t + 1 : t - 1;
            if (tp >= 0 && tp < T) v = y[((int64_t)b * T + tp) * 2 * H + c];
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, m = r0 + threadIdx.x;
        if (m < M) out[(int64_t)c * ldd + m] = tile[threadIdx.x][i];
    }
}

// hprev[b*T + t, dir*H + j] = y[b, t -/+ 1, dir*H + j] (0 at the sequence boundary); 8 bf16 per thread
__global__ void hprev_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ out, int B, int T, int H) {
    const int vec_per_row = 2 * H / 8;
    const int64_t total = (int64_t)B * T * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / vec_per_row;
        const int c8 = (int)(i - m * vec_per_row);
        const int b = (int)(m / T), t = (int)(m - (int64_t)b * T);
        const int tp = (c8 * 8 >= H) ? t + 1 : t - 1;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (tp >= 0 && tp < T) v = reinterpret_cast<const uint4*>(y)[((int64_t)b * T + tp) * vec_per_row + c8];
        reinterpret_cast<uint4*>(out)[i] = v;
    }
}

// dst[r, :] = bf16(src[rows[r], :])   (rows == nullptr: identity); one warp-stride loop over 4-column groups
template <bool SRC_BF16>
__global__ void gather_rows_bf16_kernel(const void* __restrict__ src, const int32_t* __restrict__ rows,
                                        __nv_bfloat16* __restrict__ dst, int64_t R, int C) {
    const int c4n = C >> 2;
    for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
        const int64_t sr = rows ? (int64_t)__ldg(&rows[r]) : r;
        if (SRC_BF16) {
            const uint2* s2 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(src) + sr * C);
            uint2* d2 = reinterpret_cast<uint2*>(dst + r * C);
            for (int c = threadIdx.x; c < c4n; c += blockDim.x) d2[c] = s2[c];
        } else {
            const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + sr * C);
            uint2* d2 = reinterpret_cast<uint2*>(dst + r * C);
            for (int c = threadIdx.x; c < c4n; c += blockDim.x) {
                const float4 v = s4[c];
                __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
                uint2 o;
                o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
                d2[c] = o;
            }
        }
    }
}

// zero rows t >= lengths[b] of a [B, T, row_bytes] tensor (row_bytes % 16 == 0)
__global__ void zero_pad_rows_kernel(uint4* __restrict__ p, const int32_t* __restrict__ lengths, int T, int vec_per_row) {
    const int b = blockIdx.y;
    const int len = lengths[b];
    for (int t = len + blockIdx.x; t < T; t += gridDim.x) {
        uint4* row = p + ((int64_t)b * T + t) * vec_per_row;
        for (int c = threadIdx.x; c < vec_per_row; c += blockDim.x) row[c] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// All bf16 operand copies of one BLSTM layer's weights in ONE launch (they were ~15 torch ops per layer on the
// critical path of every step): see pk2_lstm_pack_layer in pk2.h for the layouts.
struct PackArgs {
    const float *wih[2], *whh[2], *bih[2], *bhh[2];
    __nv_bfloat16 *wih_cat, *whh_p, *whh_t, *whh_tp, *wih_t;
    float* bias_cat;
    int H, I, ldk;
};
__global__ void __launch_bounds__(256) lstm_pack_kernel(PackArgs a) {
    const int H = a.H, I = a.I, H4 = 4 * a.H;
    const int64_t n_ih = (int64_t)2 * H4 * I, n_hh = (int64_t)2 * H4 * H, n_b = 2 * H4;
    const int64_t total = n_ih + n_hh + n_b;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < n_ih) {
            const int r = (int)(i / I), c = (int)(i - (int64_t)r * I);          // r = dir*4H + gate row
            const int d = r / H4;
            const __nv_bfloat16 v = __float2bfloat16(a.wih[d][(int64_t)(r - d * H4) * I + c]);
            a.wih_cat[i] = v;
            if (a.wih_t) a.wih_t[(int64_t)c * a.ldk + r] = v;
        } else if (i < n_ih + n_hh) {
            const int64_t j = i - n_ih;
            const int d = (int)(j / ((int64_t)H4 * H));
            const int64_t jj = j - (int64_t)d * H4 * H;
            const int row = (int)(jj / H), k = (int)(jj - (int64_t)row * H);   // row = gate*H + unit
            const int g = row / H, u = row - g * H;
            const int cta = u >> 5, ul = u & 31;
            const int prow = cta * 128 + g * 32 + ul;                          // index inside the direction, per-CTA order
            const __nv_bfloat16 v = __float2bfloat16(a.whh[d][jj]);
            a.whh_p[((int64_t)d * H4 + prow) * H + k] = v;
            a.whh_t[((int64_t)d * H + k) * H4 + row] = v;
            a.whh_tp[((int64_t)d * H + k) * H4 + prow] = v;
        } else {
            const int r = (int)(i - n_ih - n_hh);
            const int d = r / H4, rr = r - d * H4;
            a.bias_cat[r] = a.bih[d][rr] + a.bhh[d][rr];
        }
    }
}

__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ partial, int64_t R, int C,
                                   int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
    if (c >= C) return;
    float acc = 0.f;
    for (int64_t r = r0; r < r1; ++r) acc += __bfloat162float(src[r * C + c]);
    partial[(int64_t)blockIdx.y * C + c] = acc;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int p = 0; p < nparts; ++p) acc += partial[(int64_t)p * C + c];   // fixed order: deterministic
    out[c] = acc;
}

// scratch for the two-stage column sum, one buffer per stream (calls on different streams may overlap)
struct Scratch { float* ptr = nullptr; size_t bytes = 0; };
std::map<cudaStream_t, Scratch> g_scratch;
std::mutex g_scratch_mu;

}  // namespace

extern "C" int pk2_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
    PK2_REQUIRE(src && dst, "pk2_cast_bf16: null argument");
    if (n <= 0) return 0;
    PK2_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                "pk2_cast_bf16: misaligned pointers");
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    cast_bf16_kernel<<<(int)blocks, 256, 0, pk2::as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_transpose_bf16(const void* src, int src_bf16, void* dst, int R, int C, int lds, int ldd, void* stream) {
    PK2_REQUIRE(src && dst, "pk2_transpose_bf16: null argument");
    if (R <= 0 || C <= 0) return 0;
    dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
    PK2_REQUIRE(grid.y <= 65535, "pk2_transpose_bf16: too many rows (%d)", R);
    if (src_bf16) transpose_bf16_kernel<true><<<grid, block, 0, pk2::as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), R, C, lds, ldd);
    else transpose_bf16_kernel<false><<<grid, block, 0, pk2::as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), R, C, lds, ldd);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_lstm_hprev_t(const void* y, void* hprev_t, int B, int T, int H, int ldd, void* stream) {
    PK2_REQUIRE(y && hprev_t, "pk2_lstm_hprev_t: null argument");
    PK2_REQUIRE(H % 32 == 0, "pk2_lstm_hprev_t: H must be a multiple of 32");
    const int M = B * T;
    dim3 grid(2 * H / 32, (M + 31) / 32), block(32, 8);
    PK2_REQUIRE(grid.y <= 65535, "pk2_lstm_hprev_t: too many rows (%d)", M);
    hprev_t_kernel<<<grid, block, 0, pk2::as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y),
                                                            static_cast<__nv_bfloat16*>(hprev_t), B, T, H, ldd);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_lstm_pack_layer(const float* const* params, int H, int I, void* wih_cat, float* bias_cat, void* whh_p,
                                   void* whh_t, void* whh_tp, void* wih_t, int ldk, void* stream) {
    PK2_REQUIRE(params && wih_cat && bias_cat && whh_p && whh_t && whh_tp, "pk2_lstm_pack_layer: null argument");
    PK2_REQUIRE(H > 0 && H % 32 == 0 && I > 0, "pk2_lstm_pack_layer: H must be a positive multiple of 32");
    PK2_REQUIRE(!wih_t || ldk >= 8 * H, "pk2_lstm_pack_layer: ldk < 8H");
    PackArgs a;
    for (int d = 0; d < 2; ++d) {
        a.wih[d] = params[4 * d + 0]; a.whh[d] = params[4 * d + 1]; a.bih[d] = params[4 * d + 2]; a.bhh[d] = params[4 * d + 3];
        PK2_REQUIRE(a.wih[d] && a.whh[d] && a.bih[d] && a.bhh[d], "pk2_lstm_pack_layer: null parameter");
    }
    a.wih_cat = static_cast<__nv_bfloat16*>(wih_cat); a.whh_p = static_cast<__nv_bfloat16*>(whh_p);
    a.whh_t = static_cast<__nv_bfloat16*>(whh_t); a.whh_tp = static_cast<__nv_bfloat16*>(whh_tp);
    a.wih_t = static_cast<__nv_bfloat16*>(wih_t); a.bias_cat = bias_cat;
    a.H = H; a.I = I; a.ldk = ldk;
    lstm_pack_kernel<<<148 * 8, 256, 0, pk2::as_stream(stream)>>>(a);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_lstm_hprev(const void* y, void* hprev, int B, int T, int H, void* stream) {
    PK2_REQUIRE(y && hprev, "pk2_lstm_hprev: null argument");
    PK2_REQUIRE(H % 8 == 0 && B > 0 && T > 0, "pk2_lstm_hprev: H must be a multiple of 8");
    hprev_kernel<<<148 * 8, 256, 0, pk2::as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(y),
                                                            static_cast<__nv_bfloat16*>(hprev), B, T, H);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_gather_rows_bf16(const void* src, int src_bf16, const int32_t* rows, void* dst, int64_t R, int C,
                                    void* stream) {
    PK2_REQUIRE(src && dst, "pk2_gather_rows_bf16: null argument");
    if (R <= 0 || C <= 0) return 0;
    PK2_REQUIRE(C % 4 == 0, "pk2_gather_rows_bf16: row length must be a multiple of 4 (got %d)", C);
    PK2_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                "pk2_gather_rows_bf16: misaligned pointers");
    const int grid = (int)(R < 148 * 16 ? R : 148 * 16);
    const int threads = C >= 1024 ? 256 : 128;
    if (src_bf16) gather_rows_bf16_kernel<true><<<grid, threads, 0, pk2::as_stream(stream)>>>(src, rows, static_cast<__nv_bfloat16*>(dst), R, C);
    else gather_rows_bf16_kernel<false><<<grid, threads, 0, pk2::as_stream(stream)>>>(src, rows, static_cast<__nv_bfloat16*>(dst), R, C);
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_zero_pad_rows(void* p, const int32_t* lengths, int B, int T, int64_t row_bytes, void* stream) {
    PK2_REQUIRE(p && lengths, "pk2_zero_pad_rows: null argument");
    if (B <= 0 || T <= 0 || row_bytes <= 0) return 0;
    PK2_REQUIRE(row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0, "pk2_zero_pad_rows: rows must be 16-byte multiples");
    PK2_REQUIRE(B <= 65535, "pk2_zero_pad_rows: batch too large");
    dim3 grid(T < 64 ? T : 64, B);
    zero_pad_rows_kernel<<<grid, 256, 0, pk2::as_stream(stream)>>>(static_cast<uint4*>(p), lengths, T, (int)(row_bytes / 16));
    PK2_POST_LAUNCH();
    return 0;
}

extern "C" int pk2_colsum_bf16(const void* src, float* out, int64_t R, int C, void* stream) {
    PK2_REQUIRE(src && out, "pk2_colsum_bf16: null argument");
    if (R <= 0 || C <= 0) return 0;
    const int rows_per_block = 256;
    const int nparts = (int)((R + rows_per_block - 1) / rows_per_block);
    const size_t need = (size_t)nparts * C * sizeof(float);
    float* g_partial = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_scratch_mu);
        Scratch& sc = g_scratch[pk2::as_stream(stream)];
        if (need > sc.bytes) {
            if (sc.ptr) { PK2_CHECK(cudaStreamSynchronize(pk2::as_stream(stream))); cudaFree(sc.ptr); }
            PK2_CHECK(cudaMalloc(&sc.ptr, need));
            sc.bytes = need;
        }
        g_partial = sc.ptr;
    }
    dim3 grid((C + 127) / 128, nparts);
    PK2_REQUIRE(grid.y <= 65535, "pk2_colsum_bf16: too many rows");
    colsum_bf16_kernel<<<grid, 128, 0, pk2::as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(src), g_partial, R, C, rows_per_block);
    PK2_POST_LAUNCH();
    colsum_final_kernel<<<(C + 127) / 128, 128, 0, pk2::as_stream(stream)>>>(g_partial, out, nparts, C);
    PK2_POST_LAUNCH();
    return 0;
}
