// Encoder-side kernels that are not GEMMs (whisper.cpp encoder graph):
//   k_im2col_conv1   mel window gather + f32 -> 16-bit, rows = [x(t-1,:) | x(t,:) | x(t+1,:)]
//   k_im2col_conv2   stride-2 row gather of the conv1 output
//   k_layernorm      f32 residual stream -> 16-bit normalised activations (eps 1e-5)
//   k_attn_enc       non-causal MHA, d_head 64, flash-style online softmax, tensor cores via
//                    mma.sync.m16n8k16 (first version; the projections use tcgen05)
#include "common.cuh"
#include "decoder.cuh"

namespace sb {
extern std::atomic<uint64_t> g_launches;

// ------------------------------------------------------------------------------------------
// conv1 im2col.  Window w reads clip clip_of[w] starting at mel frame seek[w].
//   mel frame f of the clip:  f < n_calc  -> stored value
//                             f < n_len   -> per-clip floor value (frames that only saw zero pad)
//                             else        -> 0  (whisper_encode_internal zero-fills past n_len)
// out row (w, t) = [ x(t-1, 0..n_mel) | x(t, :) | x(t+1, :) ], zero outside [0, 3000).
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_im2col_conv1(Im2col1Args a, T* out) {
    // CTA: 32 output rows (t0..t0+31) of one window; needs frames t0-1 .. t0+32 (34)
    extern __shared__ float s_tile[];   // [n_mel][35]
    const int w = blockIdx.y;
    const int t0 = blockIdx.x * 32;
    const int clip = a.clip_of[w];
    const int seek = a.seek[w];
    const int ncalc = a.n_calc[clip], nlen = a.n_len[clip];
    // raw mel (engine path): global max - 8 clamp and (x + 4) / 4 applied here, the same two operations k_logmel_norm performs
    const bool raw = a.clip_max != nullptr;
    const float mmax = raw ? key_float(__ldg(a.clip_max + clip)) - 8.0f : 0.f;
    const float floor_v = raw ? (fmaxf(-10.0f, mmax) + 4.0f) * 0.25f : a.floor_val[clip];
    const float* mel = a.mel + (int64_t)clip * a.mel_clip_stride;
    for (int i = threadIdx.x; i < a.n_mel * 34; i += blockDim.x) {
        const int c = i / 34, j = i - c * 34;
        const int t = t0 + j - 1;
        float v = 0.0f;
        if (t >= 0 && t < a.n_frames) {
            const int f = seek + t;
            if (f < ncalc) { v = __ldg(mel + (int64_t)c * a.mel_stride + f); if (raw) v = (fmaxf(v, mmax) + 4.0f) * 0.25f; }
            else if (f < nlen) v = floor_v;
        }
        s_tile[c * 35 + j] = v;
    }
    __syncthreads();
    const int K = 3 * a.n_mel;
    for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
        const int r = i / K, kc = i - r * K;
        const int t = t0 + r;
        if (t >= a.n_frames) continue;
        const int k = kc / a.n_mel, c = kc - k * a.n_mel;
        out[((int64_t)w * a.n_frames + t) * K + kc] = Op16<T>::from_f32(s_tile[c * 35 + r + k]);
    }
}

// conv2 im2col: in [W*n_in, d] (n_in = 3000), out [W*n_out, 3d] (n_out = 1500), row (w,t) =
// [ y(2t-1,:) | y(2t,:) | y(2t+1,:) ], y(-1) = 0.
template <typename T>
__global__ void __launch_bounds__(256) k_im2col_conv2(const T* in, T* out, int n_in, int n_out, int d, int64_t total_vec) {
    const int vec_per_row = 3 * d / 8;     // uint4 = 8 elements
    const int dv = d / 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / vec_per_row;
        const int v = (int)(i - row * vec_per_row);
        const int k = v / dv, c = v - k * dv;
        const int w = (int)(row / n_out), t = (int)(row - (int64_t)w * n_out);
        const int src = 2 * t + k - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (src >= 0 && src < n_in)
            val = __ldg(reinterpret_cast<const uint4*>(in + ((int64_t)w * n_in + src) * d) + c);
        reinterpret_cast<uint4*>(out)[i] = val;
    }
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass in registers (mean, then centred variance).
// ------------------------------------------------------------------------------------------
template <typename T, int VPL /* float4 per lane */>
__global__ void __launch_bounds__(256) k_layernorm(const float* x, const float* gamma, const float* beta, T* out16,
                                                   float* out32, int rows, int d, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_wait();
    pdl_trigger();
    if (warp >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)warp * d);
    const int n4 = d >> 2;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        v[i] = idx < n4 ? __ldcg(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);   // L2 path: safe under PDL overlap
        s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
            float4 y;
            y.x = v[i].x * rstd * g.x + b.x; y.y = v[i].y * rstd * g.y + b.y;
            y.z = v[i].z * rstd * g.z + b.z; y.w = v[i].w * rstd * g.w + b.w;
            if (out16) {
                uint2 u;
                u.x = Op16<T>::pack2(y.x, y.y); u.y = Op16<T>::pack2(y.z, y.w);
                reinterpret_cast<uint2*>(out16 + (int64_t)warp * d)[idx] = u;
            }
            if (out32) reinterpret_cast<float4*>(out32 + (int64_t)warp * d)[idx] = y;
        }
    }
}

// ---- launchers -------------------------------------------------------------------------
template <typename T>
int im2col_conv1(const Im2col1Args& a, T* out, int n_windows, cudaStream_t st) {
    dim3 grid(ceil_div(a.n_frames, 32), n_windows);
    k_im2col_conv1<T><<<grid, 256, a.n_mel * 35 * sizeof(float), st>>>(a, out);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int im2col_conv2(const T* in, T* out, int n_windows, int n_in, int n_out, int d, cudaStream_t st) {
    const int64_t total = (int64_t)n_windows * n_out * (3 * d / 8);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    k_im2col_conv2<T><<<blocks, 256, 0, st>>>(in, out, n_in, n_out, d, total);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int layernorm(const float* x, const float* g, const float* b, T* out16, float* out32, int rows, int d, cudaStream_t st) {
    SB_CHECK_ARG(d % 4 == 0 && d <= 32 * 4 * 12, "layernorm: d must be a multiple of 4 and <= 1536");
    const int blocks = ceil_div(rows, 8);
    const int n4 = d / 4;
    if (n4 <= 32 * 2) launch_pdl(k_layernorm<T, 2>, dim3(blocks), dim3(256), 0, st, x, g, b, out16, out32, rows, d, 1e-5f);
    else if (n4 <= 32 * 6) launch_pdl(k_layernorm<T, 6>, dim3(blocks), dim3(256), 0, st, x, g, b, out16, out32, rows, d, 1e-5f);
    else if (n4 <= 32 * 10) launch_pdl(k_layernorm<T, 10>, dim3(blocks), dim3(256), 0, st, x, g, b, out16, out32, rows, d, 1e-5f);
    else launch_pdl(k_layernorm<T, 12>, dim3(blocks), dim3(256), 0, st, x, g, b, out16, out32, rows, d, 1e-5f);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

#define SB_INST(T)                                                                                          \
    template int im2col_conv1<T>(const Im2col1Args&, T*, int, cudaStream_t);                                \
    template int im2col_conv2<T>(const T*, T*, int, int, int, int, cudaStream_t);                           \
    template int layernorm<T>(const float*, const float*, const float*, T*, float*, int, int, cudaStream_t);

SB_INST(__nv_bfloat16)
SB_INST(__half)

}  // namespace sb

// stage entry points for parity tests
extern "C" int sb_layernorm_dev(int dtype, const float* x, const float* gamma, const float* beta, void* out16,
                                float* out32, int rows, int d, void* stream) {
    SB_CHECK_ARG(x && gamma && beta && (out16 || out32), "null pointer");
    if (dtype == SB_DTYPE_F16)
        return sb::layernorm<__half>(x, gamma, beta, (__half*)out16, out32, rows, d, (cudaStream_t)stream);
    return sb::layernorm<__nv_bfloat16>(x, gamma, beta, (__nv_bfloat16*)out16, out32, rows, d, (cudaStream_t)stream);
}

extern "C" int sb_attn_enc_dev(int dtype, const void* qkv, void* out, int n_windows, int n_ctx, int d_model,
                               int n_head, void* stream) {
    SB_CHECK_ARG(qkv && out, "null pointer");
    if (dtype == SB_DTYPE_F16)
        return sb::attn_enc_tc<__half>((const __half*)qkv, (__half*)out, n_windows, n_ctx, d_model, n_head, (cudaStream_t)stream);
    return sb::attn_enc_tc<__nv_bfloat16>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, n_windows, n_ctx, d_model, n_head,
                                          (cudaStream_t)stream);
}
