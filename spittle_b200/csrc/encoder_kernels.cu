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
    const float floor_v = a.floor_val[clip];
    const float* mel = a.mel + (int64_t)clip * a.mel_clip_stride;
    for (int i = threadIdx.x; i < a.n_mel * 34; i += blockDim.x) {
        const int c = i / 34, j = i - c * 34;
        const int t = t0 + j - 1;
        float v = 0.0f;
        if (t >= 0 && t < a.n_frames) {
            const int f = seek + t;
            if (f < ncalc) v = __ldg(mel + (int64_t)c * a.mel_stride + f);
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

// ------------------------------------------------------------------------------------------
// Encoder attention (non-causal), d_head = 64.
// qkv: [n_win * n_ctx, 3*d_model] rows = (window, t), columns [Q | K | V] each [head][64].
// out: [n_win * n_ctx, d_model].
// grid (ceil(n_ctx/64), n_head, n_win), 128 threads: each warp owns 16 query rows.
// ------------------------------------------------------------------------------------------
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
    __device__ __forceinline__ static void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};
template <> struct MmaOp<__half> {
    __device__ __forceinline__ static void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

constexpr int kAttnBQ = 64, kAttnBK = 64, kAttnD = 64, kAttnPad = 8;
constexpr int kAttnLd = kAttnD + kAttnPad;   // 72 elements = 144 B rows

template <typename T>
__global__ void __launch_bounds__(128) k_attn_enc(const T* __restrict__ qkv, T* __restrict__ out, int n_ctx,
                                                  int d_model, float scale_log2e) {
    __shared__ __align__(16) T sQ[kAttnBQ * kAttnLd];
    __shared__ __align__(16) T sK[2][kAttnBK * kAttnLd];
    __shared__ __align__(16) T sV[2][kAttnBK * kAttnLd];
    const int qt = blockIdx.x, head = blockIdx.y, win = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ld = 3 * (int64_t)d_model;
    const T* base = qkv + (int64_t)win * n_ctx * ld + head * kAttnD;
    const int q0 = qt * kAttnBQ;

    auto load_tile = [&](T* dst, const T* src_col, int row0) {
        // 64 rows x 8 chunks of 16 B
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 128 * i;
            const int r = idx >> 3, c = idx & 7;
            const int row = row0 + r;
            const bool ok = row < n_ctx;
            const T* src = src_col + (int64_t)(ok ? row : 0) * ld + c * 8;
            cp_async16((uint32_t)__cvta_generic_to_shared(dst + r * kAttnLd + c * 8), src, ok);
        }
    };
    load_tile(sQ, base, q0);
    load_tile(sK[0], base + d_model, 0);
    load_tile(sV[0], base + 2 * d_model, 0);
    cp_async_commit();

    const int n_kt = (n_ctx + kAttnBK - 1) / kAttnBK;
    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) {
            load_tile(sK[buf ^ 1], base + d_model, (kt + 1) * kAttnBK);
            load_tile(sV[buf ^ 1], base + 2 * d_model, (kt + 1) * kAttnBK);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(
                    sQ + (warp * 16 + (lane & 15)) * kAttnLd + ks * 16 + (lane >> 4) * 8);
                ldsm_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {   // pairs of key n-tiles
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(
                    sK[buf] + (jp * 16 + (lane & 7) + (lane >> 4) * 8) * kAttnLd + ks * 16 + ((lane >> 3) & 1) * 8);
                ldsm_x4(addr, b0, b1, b2, b3);
                MmaOp<T>::mma(s[2 * jp], qf[ks], b0, b1);
                MmaOp<T>::mma(s[2 * jp + 1], qf[ks], b2, b3);
            }
        }
        // mask keys beyond n_ctx (last tile only)
        const int kbase = kt * kAttnBK;
        if (kbase + kAttnBK > n_ctx) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = kbase + j * 8 + (lane & 3) * 2;
                if (key >= n_ctx) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
                if (key + 1 >= n_ctx) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
            }
        }
        // online softmax (rows g = lane/4 and g + 8)
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f((m0 - mx0) * scale_log2e), c1 = exp2f((m1 - mx1) * scale_log2e);
        m0 = mx0; m1 = mx1;
        const float ms0 = mx0 * scale_log2e, ms1 = mx1 * scale_log2e;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t pf[4][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(fmaf(s[j][0], scale_log2e, -ms0));
            const float p1 = exp2f(fmaf(s[j][1], scale_log2e, -ms0));
            const float p2 = exp2f(fmaf(s[j][2], scale_log2e, -ms1));
            const float p3 = exp2f(fmaf(s[j][3], scale_log2e, -ms1));
            // round to the operand type first so the row sum matches what the PV mma consumes
            const uint32_t u01 = Op16<T>::pack2(p0, p1), u23 = Op16<T>::pack2(p2, p3);
            const float2 f01 = Op16<T>::unpack2(u01), f23 = Op16<T>::unpack2(u23);
            rs0 += f01.x + f01.y; rs1 += f23.x + f23.y;
            pf[j >> 1][(j & 1) * 2 + 0] = u01;
            pf[j >> 1][(j & 1) * 2 + 1] = u23;
        }
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
        // O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {       // 16 keys per step
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {   // pairs of d n-tiles
                uint32_t b0, b1, b2, b3;
                const uint32_t addr = (uint32_t)__cvta_generic_to_shared(
                    sV[buf] + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kAttnLd + jp * 16 + (lane >> 4) * 8);
                ldsm_x4_t(addr, b0, b1, b2, b3);
                MmaOp<T>::mma(o[2 * jp], pf[ks], b0, b1);
                MmaOp<T>::mma(o[2 * jp + 1], pf[ks], b2, b3);
            }
        }
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
    T* obase = out + (int64_t)win * n_ctx * d_model + head * kAttnD + (lane & 3) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (r0 < n_ctx) *reinterpret_cast<uint32_t*>(obase + (int64_t)r0 * d_model + j * 8) = Op16<T>::pack2(o[j][0] * inv0, o[j][1] * inv0);
        if (r1 < n_ctx) *reinterpret_cast<uint32_t*>(obase + (int64_t)r1 * d_model + j * 8) = Op16<T>::pack2(o[j][2] * inv1, o[j][3] * inv1);
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
template <typename T>
int attn_enc(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st) {
    SB_CHECK_ARG(d_model == n_head * kAttnD, "attention: d_head must be 64");
    dim3 grid(ceil_div(n_ctx, kAttnBQ), n_head, n_windows);
    const float scale_log2e = (1.0f / 8.0f) * 1.4426950408889634f;
    k_attn_enc<T><<<grid, 128, 0, st>>>(qkv, out, n_ctx, d_model, scale_log2e);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

#define SB_INST(T)                                                                                          \
    template int im2col_conv1<T>(const Im2col1Args&, T*, int, cudaStream_t);                                \
    template int im2col_conv2<T>(const T*, T*, int, int, int, int, cudaStream_t);                           \
    template int layernorm<T>(const float*, const float*, const float*, T*, float*, int, int, cudaStream_t); \
    template int attn_enc<T>(const T*, T*, int, int, int, int, cudaStream_t);
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
    if (sb::use_tc_attention()) {
        if (dtype == SB_DTYPE_F16)
            return sb::attn_enc_tc<__half>((const __half*)qkv, (__half*)out, n_windows, n_ctx, d_model, n_head, (cudaStream_t)stream);
        return sb::attn_enc_tc<__nv_bfloat16>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, n_windows, n_ctx, d_model, n_head,
                                              (cudaStream_t)stream);
    }
    if (dtype == SB_DTYPE_F16)
        return sb::attn_enc<__half>((const __half*)qkv, (__half*)out, n_windows, n_ctx, d_model, n_head, (cudaStream_t)stream);
    return sb::attn_enc<__nv_bfloat16>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, n_windows, n_ctx, d_model, n_head,
                                       (cudaStream_t)stream);
}
