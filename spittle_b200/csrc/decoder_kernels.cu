// Decoder-step kernels (whisper.cpp decoder graph + sampler, SURVEY.md App. C.3 / C.4, rows
// a7 / a8 of 8(a)).  One "step" advances every live sequence of the batch by one token; all
// sequences share the same position (prompt is identical), finished ones are masked.
//
//   k_dec_embed           x = token_embedding[tok] + positional_embedding[pos]
//   k_skinny_gemm         Y[B,N] = X[B,K] W[N,K]^T (+bias, GELU, +residual); B <= 64 per pass,
//                         HBM-bound weight streaming; tensor cores via mma.sync with the weight
//                         rows as the M operand, 8 warps split K, smem reduction, fused epilogue
//   k_dec_self_attn       append K/V to the cache, causal attention over <= 448 positions
//   k_dec_cross_attn      attention over the 1500 cached encoder keys (the dominant HBM stream)
//   k_logits_filter_argmax  whisper_process_logits + whisper_sample_token(best) + the per-token
//                         bookkeeping of whisper_full (seek_delta / result_len / has_ts / stop)
#include "common.cuh"
#include "decoder.cuh"

namespace sb {
extern std::atomic<uint64_t> g_launches;

// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_dec_embed(const T* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                                   const int* __restrict__ tokens, const int* __restrict__ pos_ptr,
                                                   float* __restrict__ x, int d) {
    const int b = blockIdx.x;
    pdl_wait();
    pdl_trigger();
    const int tok = __ldcg(tokens + b);
    const int pos = __ldcg(pos_ptr);
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        x[(int64_t)b * d + i] = Op16<T>::to_f32(tok_emb[(int64_t)tok * d + i]) + pos_emb[(int64_t)pos * d + i];
}

// ------------------------------------------------------------------------------------------
// skinny GEMM.  grid.x = ceil(N/16) row tiles, grid.y = batch chunks of 64.  256 threads.
// Fragment trick: both operands are read with 16-byte vector loads of 8 consecutive k; the
// k-permutation is the same for A and B so the products pair up correctly.
// ------------------------------------------------------------------------------------------
template <typename T> struct MmaOpD;
template <> struct MmaOpD<__nv_bfloat16> {
    __device__ __forceinline__ static void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};
template <> struct MmaOpD<__half> {
    __device__ __forceinline__ static void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) k_skinny_gemm(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw,
                                                     int Bn, int N, int K, SkinnyEpilogue ep) {
    __shared__ float s_red[8][16][65];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = blockIdx.x * 16;
    const int b0 = blockIdx.y * 64;
    const int nb = min(64, Bn - b0);
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    const int r_lo = min(row0 + g, N - 1), r_hi = min(row0 + g + 8, N - 1);
    const T* w_lo = W + (int64_t)r_lo * ldw + t * 8;
    const T* w_hi = W + (int64_t)r_hi * ldw + t * 8;
    const T* xb = X + (int64_t)b0 * ldx + t * 8;
    const int n_blk = K / 32;      // K % 32 == 0 enforced by the host
    // this warp's k-blocks: warp, warp + 8, ...  Weight loads are issued kWB blocks ahead of
    // their use (the HBM stream must be in flight before anything waits on it); the activation
    // rows come from L2 and are double-buffered one block ahead.
    constexpr int kWB = 4;
    const int n_it = (n_blk - warp + 7) / 8;       // iterations of this warp (may be 0)
    uint4 wlo[kWB], whi[kWB];
#pragma unroll
    for (int i = 0; i < kWB; ++i)
        if (i < n_it) { wlo[i] = ldg_nc_v4(w_lo + (warp + 8 * i) * 32); whi[i] = ldg_nc_v4(w_hi + (warp + 8 * i) * 32); }
    pdl_wait();      // weights are immutable: only the activations depend on the previous kernel
    // epilogue operands do not depend on the main loop: fetch them now so their L2 round trip is hidden
    const int e_n = tid >> 2, e_rq = (tid & 3) * 4;
    float e_res[4] = {0.f, 0.f, 0.f, 0.f}, e_bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (e_n < nb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + e_rq + i;
            if (row < N) {
                if (ep.bias) e_bias[i] = __ldg(ep.bias + row);
                if (ep.residual) e_res[i] = __ldcg(ep.residual + (int64_t)(b0 + e_n) * ep.ldr + row);
            }
        }
    }
    // activation fragments: kXB k-blocks in flight (L2 latency ~ 0.4 us per round trip)
    constexpr int kXB = 3;
    uint4 xq[kXB][8];
#pragma unroll
    for (int i = 0; i < kXB; ++i)
        if (i < n_it) {
            const int k1 = (warp + 8 * i) * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = j * 8 + g;
                xq[i][j] = n < nb ? __ldcg(reinterpret_cast<const uint4*>(xb + (int64_t)n * ldx + k1)) : make_uint4(0, 0, 0, 0);
            }
        }
    // n_it is a multiple of nothing in particular: rotate the two register rings with fully unrolled
    // bodies of lcm(kWB, kXB) = 12 iterations so every ring index is a compile-time constant
    for (int it0 = 0; it0 < n_it; it0 += 12) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int it = it0 + i;
            if (it >= n_it) break;
            const uint4 alo = wlo[i % kWB], ahi = whi[i % kWB];
            if (it + kWB < n_it) {
                wlo[i % kWB] = ldg_nc_v4(w_lo + (warp + 8 * (it + kWB)) * 32);
                whi[i % kWB] = ldg_nc_v4(w_hi + (warp + 8 * (it + kWB)) * 32);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                MmaOpD<T>::mma(acc[j], alo.x, ahi.x, alo.y, ahi.y, xq[i % kXB][j].x, xq[i % kXB][j].y);
                MmaOpD<T>::mma(acc[j], alo.z, ahi.z, alo.w, ahi.w, xq[i % kXB][j].z, xq[i % kXB][j].w);
            }
            if (it + kXB < n_it) {     // refill this slot: kXB k-blocks of activations stay in flight
                const int k1 = (warp + 8 * (it + kXB)) * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = j * 8 + g;
                    xq[i % kXB][j] = n < nb ? __ldcg(reinterpret_cast<const uint4*>(xb + (int64_t)n * ldx + k1)) : make_uint4(0, 0, 0, 0);
                }
            }
        }
    }
    pdl_trigger();   // main loop done: let the next kernel get scheduled and prefetch its weights
    // acc[j]: c0,c1 = (row g, batch j*8+2t, +1), c2,c3 = (row g+8, ...)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s_red[warp][g][j * 8 + 2 * t] = acc[j][0];
        s_red[warp][g][j * 8 + 2 * t + 1] = acc[j][1];
        s_red[warp][g + 8][j * 8 + 2 * t] = acc[j][2];
        s_red[warp][g + 8][j * 8 + 2 * t + 1] = acc[j][3];
    }
    __syncthreads();
    // epilogue: thread -> (batch n = tid / 4, 4 consecutive rows r = (tid % 4) * 4)
    const int n = e_n, rq = e_rq;
    if (n < nb) {
        const int b = b0 + n;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rq + i, row = row0 + r;
            if (row >= N) break;
            float v = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) v += s_red[w8][r][n];
            v += e_bias[i];
            if (ep.act == 1) v = gelu_tanh(v);
            v += e_res[i];
            if (ep.out32) ep.out32[(int64_t)b * ep.ldo32 + row] = v;
            if (ep.out16) reinterpret_cast<T*>(ep.out16)[(int64_t)b * ep.ldo16 + row] = Op16<T>::from_f32(v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// self attention for one new token per sequence.  grid = B * n_head / 4, 128 threads (warp per
// (b, head)).  qkv: [B, 3d] (this step); cache K/V: [B][n_text_ctx][d].
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) k_dec_self_attn(const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc,
                                                       T* __restrict__ out, const int* __restrict__ pos_ptr, int Bn,
                                                       int n_head, int d, int n_text_ctx) {
    __shared__ float s_p[4][448];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = blockIdx.x * 4 + warp;
    pdl_wait();
    pdl_trigger();
    if (idx >= Bn * n_head) return;
    const int b = idx / n_head, h = idx - b * n_head;
    const int pos = __ldcg(pos_ptr);          // index of the new token; attends to [0, pos]
    const T* q = qkv + (int64_t)b * 3 * d + h * 64;
    T* kb = kc + ((int64_t)b * n_text_ctx) * d + h * 64;
    T* vb = vc + ((int64_t)b * n_text_ctx) * d + h * 64;
    // append this step's K, V (each lane moves 2 elements)
    reinterpret_cast<uint32_t*>(kb + (int64_t)pos * d)[lane] = __ldcg(reinterpret_cast<const uint32_t*>(q + d) + lane);
    reinterpret_cast<uint32_t*>(vb + (int64_t)pos * d)[lane] = __ldcg(reinterpret_cast<const uint32_t*>(q + 2 * d) + lane);
    __syncwarp();
    // scores: lane <-> key
    float qf[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float2 f = Op16<T>::unpack2(__ldcg(reinterpret_cast<const uint32_t*>(q) + i));
        qf[2 * i] = f.x; qf[2 * i + 1] = f.y;
    }
    const int n_keys = pos + 1;
    float mx = -INFINITY;
    for (int k = lane; k < n_keys; k += 32) {
        const uint4* kr = reinterpret_cast<const uint4*>(kb + (int64_t)k * d);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldcg(kr + c);
            float2 f;
            f = Op16<T>::unpack2(u.x); s = fmaf(qf[c * 8 + 0], f.x, s); s = fmaf(qf[c * 8 + 1], f.y, s);
            f = Op16<T>::unpack2(u.y); s = fmaf(qf[c * 8 + 2], f.x, s); s = fmaf(qf[c * 8 + 3], f.y, s);
            f = Op16<T>::unpack2(u.z); s = fmaf(qf[c * 8 + 4], f.x, s); s = fmaf(qf[c * 8 + 5], f.y, s);
            f = Op16<T>::unpack2(u.w); s = fmaf(qf[c * 8 + 6], f.x, s); s = fmaf(qf[c * 8 + 7], f.y, s);
        }
        s *= 0.125f;
        s_p[warp][k] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < n_keys; k += 32) {
        const float p = __expf(s_p[warp][k] - mx);
        s_p[warp][k] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    // PV: lane <-> 2 output dims; probabilities rounded to the operand type like ggml's mul_mat.
    // Keys are taken 16 at a time with all 16 value loads issued before the FMAs (the loop is
    // otherwise a chain of exposed L2 latencies).
    float o0 = 0.f, o1 = 0.f;
    for (int k0 = 0; k0 < n_keys; k0 += 16) {
        uint32_t vv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = min(k0 + i, n_keys - 1);
            vv[i] = __ldcg(reinterpret_cast<const uint32_t*>(vb + (int64_t)k * d) + lane);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (k0 + i < n_keys) {
                const float p = Op16<T>::to_f32(Op16<T>::from_f32(s_p[warp][k0 + i] * inv));
                const float2 f = Op16<T>::unpack2(vv[i]);
                o0 = fmaf(p, f.x, o0); o1 = fmaf(p, f.y, o1);
            }
        }
    }
    reinterpret_cast<uint32_t*>(out + (int64_t)b * d + h * 64)[lane] = Op16<T>::pack2(o0, o1);
}

// ------------------------------------------------------------------------------------------
// cross attention.  grid = (n_head, B), 256 threads.  Kc/Vc rows are strided (ld_kv) inside the
// fused cross-KV buffer [W*1500, L*2*d]; K then V are each streamed exactly once through a
// cp.async double-buffered shared-memory tile.
// ------------------------------------------------------------------------------------------
constexpr int kXKeysPerTile = 128;
constexpr int kXLd = 72;   // padded row (elements)

__device__ __forceinline__ void cp_async16_d(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}

// With fq.x != nullptr the kernel also performs the cross-attention LayerNorm and the query
// projection of its own (sequence, head): q_h = Wq[h*64 .. h*64+63, :] . round16(LN(x_b)) + bq
// (two fewer launches per decoder layer; the extra prologue hides behind the K-tile stream).
template <typename T>
__global__ void __launch_bounds__(256) k_dec_cross_attn(const T* __restrict__ q, int ldq, const T* __restrict__ kbase,
                                                        const T* __restrict__ vbase, int64_t ld_kv, int64_t win_stride,
                                                        T* __restrict__ out, int d, int n_ctx, FusedQ fq) {
    __shared__ __align__(16) T s_tile[2][kXKeysPerTile * kXLd];
    __shared__ float s_sc[1504];      // scores; doubles as the normalised activation row in the fused prologue
    __shared__ float s_q[64];
    __shared__ float s_red[8];
    __shared__ float s_o[4][64];
    const int h = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* kp = kbase + (int64_t)b * win_stride + h * 64;
    const T* vp = vbase + (int64_t)b * win_stride + h * 64;
    const int n_tiles = (n_ctx + kXKeysPerTile - 1) / kXKeysPerTile;

    auto load_tile = [&](int buf, const T* src, int key0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i;      // 128 rows x 8 chunks
            const int r = idx >> 3, c = idx & 7;
            const int key = key0 + r;
            const bool ok = key < n_ctx;
            cp_async16_d((uint32_t)__cvta_generic_to_shared(&s_tile[buf][r * kXLd + c * 8]),
                         src + (int64_t)(ok ? key : 0) * ld_kv + c * 8, ok);
        }
        asm volatile("cp.async.commit_group;");
    };

    // ---- pass 1: scores ----
    load_tile(0, kp, 0);          // the encoder wrote K/V long ago: start the stream before the dependency wait
    pdl_wait();
    if (fq.x == nullptr) {
        if (tid < 32) {
            const float2 f = Op16<T>::unpack2(__ldcg(reinterpret_cast<const uint32_t*>(q + (int64_t)b * ldq + h * 64) + tid));
            s_q[2 * tid] = f.x * 0.125f; s_q[2 * tid + 1] = f.y * 0.125f;
        }
    } else {
        // LayerNorm of row b (two-pass, values held in registers; d <= 1536)
        float xv[6];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int k = tid + 256 * i;
            xv[i] = k < d ? __ldcg(fq.x + (int64_t)b * d + k) : 0.f;
            sum += xv[i];
        }
        sum = warp_sum(sum);
        if (lane == 0) s_red[warp] = sum;
        __syncthreads();
        float mean = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) mean += s_red[w];
        mean /= (float)d;
        __syncthreads();
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int k = tid + 256 * i;
            if (k < d) { xv[i] -= mean; sq += xv[i] * xv[i]; }
        }
        sq = warp_sum(sq);
        if (lane == 0) s_red[warp] = sq;
        __syncthreads();
        float var = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) var += s_red[w];
        const float rstd = rsqrtf(var / (float)d + 1e-5f);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int k = tid + 256 * i;
            if (k < d) s_sc[k] = Op16<T>::to_f32(Op16<T>::from_f32(xv[i] * rstd * __ldg(fq.ln_g + k) + __ldg(fq.ln_b + k)));
        }
        __syncthreads();
        // q_h[j] = Wq[h*64 + j, :] . h + bq : 4 threads per output row, each a contiguous quarter of K
        const int j = tid >> 2, part = tid & 3;
        const int kq = d >> 2;                       // d % 32 == 0
        const T* wr = reinterpret_cast<const T*>(fq.wq) + (int64_t)(h * 64 + j) * d + part * kq;
        const float* hr = s_sc + part * kq;
        float acc = 0.f;
        for (int k = 0; k < kq; k += 8) {
            const uint4 u = ldg_nc_v4(wr + k);
            float2 f;
            f = Op16<T>::unpack2(u.x); acc = fmaf(f.x, hr[k + 0], acc); acc = fmaf(f.y, hr[k + 1], acc);
            f = Op16<T>::unpack2(u.y); acc = fmaf(f.x, hr[k + 2], acc); acc = fmaf(f.y, hr[k + 3], acc);
            f = Op16<T>::unpack2(u.z); acc = fmaf(f.x, hr[k + 4], acc); acc = fmaf(f.y, hr[k + 5], acc);
            f = Op16<T>::unpack2(u.w); acc = fmaf(f.x, hr[k + 6], acc); acc = fmaf(f.y, hr[k + 7], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        __syncthreads();                              // everyone is done reading the activation row in s_sc
        if (part == 0) s_q[j] = Op16<T>::to_f32(Op16<T>::from_f32(acc + __ldg(fq.bq + h * 64 + j))) * 0.125f;
    }
    for (int tI = 0; tI < n_tiles; ++tI) {
        const int buf = tI & 1;
        if (tI + 1 < n_tiles) { load_tile(buf ^ 1, kp, (tI + 1) * kXKeysPerTile); asm volatile("cp.async.wait_group 1;"); }
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        if (tid < kXKeysPerTile) {
            const int key = tI * kXKeysPerTile + tid;
            const uint4* kr = reinterpret_cast<const uint4*>(&s_tile[buf][tid * kXLd]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = kr[c];
                float2 f;
                f = Op16<T>::unpack2(u.x); s = fmaf(s_q[c * 8 + 0], f.x, s); s = fmaf(s_q[c * 8 + 1], f.y, s);
                f = Op16<T>::unpack2(u.y); s = fmaf(s_q[c * 8 + 2], f.x, s); s = fmaf(s_q[c * 8 + 3], f.y, s);
                f = Op16<T>::unpack2(u.z); s = fmaf(s_q[c * 8 + 4], f.x, s); s = fmaf(s_q[c * 8 + 5], f.y, s);
                f = Op16<T>::unpack2(u.w); s = fmaf(s_q[c * 8 + 6], f.x, s); s = fmaf(s_q[c * 8 + 7], f.y, s);
            }
            if (key < n_ctx) s_sc[key] = s;
        }
        __syncthreads();
    }
    // prefetch the first V tile while the softmax statistics are reduced
    load_tile(0, vp, 0);
    float mx = -INFINITY;
    for (int k = tid; k < n_ctx; k += 256) mx = fmaxf(mx, s_sc[k]);
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int k = tid; k < n_ctx; k += 256) { const float p = __expf(s_sc[k] - mx); s_sc[k] = p; sum += p; }
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[w];
    const float inv = 1.0f / sum;
    // ---- pass 2: O = P V ; thread -> (dim pair dp = tid % 32, key group kg = tid / 32) ----
    float o0 = 0.f, o1 = 0.f;
    for (int tI = 0; tI < n_tiles; ++tI) {
        const int buf = tI & 1;
        if (tI + 1 < n_tiles) { load_tile(buf ^ 1, vp, (tI + 1) * kXKeysPerTile); asm volatile("cp.async.wait_group 1;"); }
        else asm volatile("cp.async.wait_group 0;");
        __syncthreads();
        const int key0 = tI * kXKeysPerTile;
#pragma unroll 4
        for (int r = warp; r < kXKeysPerTile; r += 8) {
            const int key = key0 + r;
            if (key >= n_ctx) break;
            const float p = Op16<T>::to_f32(Op16<T>::from_f32(s_sc[key] * inv));
            const float2 f = Op16<T>::unpack2(reinterpret_cast<const uint32_t*>(&s_tile[buf][r * kXLd])[lane]);
            o0 = fmaf(p, f.x, o0); o1 = fmaf(p, f.y, o1);
        }
        __syncthreads();
    }
    pdl_trigger();
    // reduce the 8 key groups
    if (warp >= 4) { s_o[warp - 4][2 * lane] = o0; s_o[warp - 4][2 * lane + 1] = o1; }
    __syncthreads();
    if (warp < 4) { o0 += s_o[warp][2 * lane]; o1 += s_o[warp][2 * lane + 1]; }
    __syncthreads();
    if (warp >= 1 && warp < 4) { s_o[warp][2 * lane] = o0; s_o[warp][2 * lane + 1] = o1; }
    __syncthreads();
    if (warp == 0) {
        o0 += s_o[1][2 * lane] + s_o[2][2 * lane] + s_o[3][2 * lane];
        o1 += s_o[1][2 * lane + 1] + s_o[2][2 * lane + 1] + s_o[3][2 * lane + 1];
        reinterpret_cast<uint32_t*>(out + (int64_t)b * d + h * 64)[lane] = Op16<T>::pack2(o0, o1);
    }
}

// ------------------------------------------------------------------------------------------
// logits filter + greedy sampler + bookkeeping.  One CTA (1024 threads) per sequence; the
// vocabulary row (207 KB, L2-resident: it was just written by the logits GEMM) is swept three
// times: maxima, exp-sums, arg-max with runner-up.
// ------------------------------------------------------------------------------------------
constexpr int kSampThreads = 1024;

struct BlockRed {
    float* sf; int* si;
    __device__ float max_f(float v) {
        v = warp_max(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sf[threadIdx.x >> 5] = v;
        __syncthreads();
        float r = sf[0];
        for (int i = 1; i < kSampThreads / 32; ++i) r = fmaxf(r, sf[i]);
        return r;
    }
    __device__ float sum_f(float v) {
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sf[threadIdx.x >> 5] = v;
        __syncthreads();
        float r = 0.f;
        for (int i = 0; i < kSampThreads / 32; ++i) r += sf[i];
        return r;
    }
    // argmax with lowest-index tie break
    __device__ void argmax(float v, int idx, float& ov, int& oi) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { sf[threadIdx.x >> 5] = v; si[threadIdx.x >> 5] = idx; }
        __syncthreads();
        ov = sf[0]; oi = si[0];
        for (int i = 1; i < kSampThreads / 32; ++i)
            if (sf[i] > ov || (sf[i] == ov && si[i] < oi)) { ov = sf[i]; oi = si[i]; }
    }
};

struct LogitMask {
    SpecialIds sp;
    bool is_initial, last_ts, pen_ts, has_ts, suppress_blank, no_timestamps;
    int init_lim, mono_lim, max_initial_tid;
    __device__ __forceinline__ bool suppressed(int id) const {
        if (suppress_blank && is_initial && (id == sp.eot || id == sp.blank)) return true;
        if (id == sp.not_ || id == sp.sot || id == sp.nosp || id == sp.solm || id == sp.translate ||
            id == sp.transcribe || id == sp.prev) return true;
        if (id >= sp.lang_first && id < sp.lang_first + sp.num_languages) return true;
        if (no_timestamps && id >= sp.beg) return true;
        if (last_ts) {
            if (pen_ts) { if (id >= sp.beg) return true; }
            else { if (id < sp.eot) return true; }
        }
        if (is_initial && max_initial_tid >= 0 && id >= init_lim) return true;
        if (has_ts && id >= sp.beg && id < mono_lim) return true;
        return false;
    }
};

// online log-sum-exp accumulator
struct Lse {
    float m, s;
    __device__ __forceinline__ void add(float x) {
        if (x > m) { s = s * __expf(m - x) + 1.0f; m = x; }
        else s += __expf(x - m);
    }
    __device__ __forceinline__ void merge(float m2, float s2) {
        const float M = fmaxf(m, m2);
        if (M == -INFINITY) return;
        s = s * __expf(m - M) + s2 * __expf(m2 - M);
        m = M;
    }
};

__global__ void __launch_bounds__(kSampThreads) k_logits_filter_argmax(const float* __restrict__ logits, int ld,
                                                                       SamplerArgs a) {
    __shared__ float sf[32];
    __shared__ float sg[32];
    __shared__ int si[32];
    BlockRed red{sf, si};
    const int b = blockIdx.x;
    pdl_wait();
    pdl_trigger();
    SeqState st;
    {
        static_assert(sizeof(SeqState) == 48, "SeqState is read as three int4");
        const int4* sp4 = reinterpret_cast<const int4*>(a.state + b);
        int4* dp4 = reinterpret_cast<int4*>(&st);
        dp4[0] = __ldcg(sp4); dp4[1] = __ldcg(sp4 + 1); dp4[2] = __ldcg(sp4 + 2);
    }
    const int step = __ldcg(a.step_ptr);   // index of the token being sampled (i in whisper_full)
    const int pos = __ldcg(a.pos_ptr);
    if (pos < a.n_prompt - 1) {            // still feeding the prompt: queue its next token
        if (threadIdx.x == 0) a.next_tokens[b] = a.prompt[pos + 1];
        return;
    }
    if (st.done) return;
    const SpecialIds sp = a.sp;
    const int V = a.n_vocab;
    const float* lg = logits + (int64_t)b * ld;
    LogitMask mk;
    mk.sp = sp;
    mk.is_initial = st.n_tok == 0;
    mk.last_ts = st.n_tok > 0 && st.last >= sp.beg;
    mk.pen_ts = st.n_tok < 2 || st.prev >= sp.beg;
    mk.has_ts = st.has_ts != 0;
    mk.suppress_blank = a.suppress_blank != 0;
    mk.no_timestamps = a.no_timestamps != 0;
    mk.max_initial_tid = a.max_initial_tid;
    mk.init_lim = sp.beg + a.max_initial_tid + 1;     // tokens >= this are suppressed at step 0
    mk.mono_lim = sp.beg + st.seek_delta / 2;
    // plain text tokens [0, eot) only see two rules: the pairing rule and suppress_blank
    const bool text_off = mk.last_ts && !mk.pen_ts;
    const int blank_off = (mk.suppress_blank && mk.is_initial) ? sp.blank : -1;
    // sweep 1: online log-sum-exp over all / timestamp tokens, text maximum
    Lse all{-INFINITY, 0.f}, ts{-INFINITY, 0.f};
    float mx_text = -INFINITY;
    if (!text_off)
        for (int id = threadIdx.x; id < sp.eot; id += kSampThreads) {
            if (id == blank_off) continue;
            const float x = __ldcg(lg + id);
            all.add(x);
            mx_text = fmaxf(mx_text, x);
        }
    for (int id = sp.eot + threadIdx.x; id < V; id += kSampThreads) {
        if (mk.suppressed(id)) continue;
        const float x = __ldcg(lg + id);
        all.add(x);
        if (id >= sp.beg) ts.add(x); else mx_text = fmaxf(mx_text, x);
    }
    // block-combine the two accumulators and the text maximum
    auto combine = [&](Lse& v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, v.m, o), s2 = __shfl_xor_sync(0xffffffffu, v.s, o);
            v.merge(m2, s2);
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { sf[threadIdx.x >> 5] = v.m; sg[threadIdx.x >> 5] = v.s; }
        __syncthreads();
        Lse r{sf[0], sg[0]};
        for (int i = 1; i < kSampThreads / 32; ++i) r.merge(sf[i], sg[i]);
        v = r;
    };
    combine(all);
    combine(ts);
    mx_text = red.max_f(mx_text);
    const float lse = logf(all.s) + all.m;
    // timestamp_logprob = logsumexp(logprobs[beg:]) ; max_text_token_logprob = max(logprobs[:beg])
    const float ts_logprob = (ts.m > -INFINITY && ts.s > 0.f) ? (logf(ts.s) + ts.m - lse) : -INFINITY;
    const float text_logprob = mx_text - lse;
    const bool force_ts = ts_logprob > text_logprob;
    // sweep 2: arg-max (first maximum in ascending id) and the runner-up value
    float bv = -INFINITY, b2 = -INFINITY; int bi = 0x7fffffff;
    if (!text_off && !force_ts)
        for (int id = threadIdx.x; id < sp.eot; id += kSampThreads) {
            if (id == blank_off) continue;
            const float x = __ldcg(lg + id);
            if (x > bv) { b2 = bv; bv = x; bi = id; }
            else if (x > b2) b2 = x;
        }
    for (int id = sp.eot + threadIdx.x; id < V; id += kSampThreads) {
        if (mk.suppressed(id) || (force_ts && id < sp.beg)) continue;
        const float x = __ldcg(lg + id);
        if (x > bv) { b2 = bv; bv = x; bi = id; }
        else if (x > b2) b2 = x;
    }
    float gv; int gi;
    red.argmax(bv, bi, gv, gi);
    const float sv = red.max_f(bi == gi ? b2 : bv);
    if (threadIdx.x != 0) return;

    int tok = gi;
    if (a.forced) {
        const int f = __ldcg(a.forced + (int64_t)b * a.n_max + step);
        if (f >= 0) tok = f;
    }
    a.tokens_out[(int64_t)b * a.n_max + step] = tok;
    if (a.margins_out) a.margins_out[(int64_t)b * a.n_max + step] = gv - sv;
    a.next_tokens[b] = tok;
    st.prev = st.last; st.last = tok; st.n_tok += 1;
    st.sum_logprob += (gv - lse);
    // ---- whisper_full bookkeeping (App. C.4) ----
    bool stop = false;
    if (tok > sp.beg) {
        const int sd_new = 2 * (tok - sp.beg);
        if (st.has_ts && st.seek_delta > sd_new && st.result_len < step) { st.failed = 1; stop = true; }
        else { st.seek_delta = sd_new; st.result_len = step + 1; st.has_ts = 1; }
    }
    if (!stop && (tok == sp.eot || (st.has_ts && st.seek + st.seek_delta + 100 >= st.seek_end))) {
        if (st.result_len == 0 && !a.no_timestamps) {
            if (st.seek + st.seek_delta + 100 >= st.seek_end) st.result_len = step + 1;
            else st.failed = 1;
        }
        if (!st.failed && (a.single_segment || a.no_timestamps)) { st.result_len = step + 1; st.seek_delta = 3000; }
        stop = true;
    }
    if (!stop && step == a.n_max - 1 && (st.result_len == 0 || st.seek_delta < 1500)) { st.failed = 1; stop = true; }
    if (!stop && step == a.n_max - 1) stop = true;
    if (stop) { st.done = 1; atomicAdd(a.n_done, 1); }
    a.state[b] = st;
}

// advance the shared position / step counters (single thread) after a step
__global__ void k_dec_advance(int* pos_ptr, int* step_ptr, int n_prompt) {
    pdl_wait();
    pdl_trigger();
    const int p = __ldcg(pos_ptr);
    if (p >= n_prompt - 1) *step_ptr = __ldcg(step_ptr) + 1;
    *pos_ptr = p + 1;
}

// ---- launchers -------------------------------------------------------------------------
template <typename T>
int dec_embed(const T* tok_emb, const float* pos_emb, const int* tokens, const int* pos_ptr, float* x, int Bn, int d,
              cudaStream_t st) {
    launch_pdl(k_dec_embed<T>, dim3(Bn), dim3(256), 0, st, tok_emb, pos_emb, tokens, pos_ptr, x, d);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int skinny_gemm(const T* X, int ldx, const T* W, int ldw, int Bn, int N, int K, const SkinnyEpilogue& ep, cudaStream_t st) {
    SB_CHECK_ARG(K % 32 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny gemm: K % 32 and 16-byte row alignment required");
    dim3 grid(ceil_div(N, 16), ceil_div(Bn, 64));
    launch_pdl(k_skinny_gemm<T>, grid, dim3(256), 0, st, X, ldx, W, ldw, Bn, N, K, ep);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int dec_self_attn(const T* qkv, T* kc, T* vc, T* out, const int* pos_ptr, int Bn, int n_head, int d, int n_text_ctx,
                  cudaStream_t st) {
    SB_CHECK_ARG(n_text_ctx <= 448 && d == n_head * 64, "self attention: n_text_ctx <= 448, d_head 64");
    launch_pdl(k_dec_self_attn<T>, dim3(ceil_div(Bn * n_head, 4)), dim3(128), 0, st, qkv, kc, vc, out, pos_ptr, Bn, n_head, d, n_text_ctx);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int dec_cross_attn(const T* q, int ldq, const T* kbase, const T* vbase, int64_t ld_kv, int64_t win_stride, T* out, int Bn,
                   int n_head, int d, int n_ctx, const FusedQ& fq, cudaStream_t st) {
    SB_CHECK_ARG(n_ctx <= 1504 && d == n_head * 64 && d <= 1504 && d % 32 == 0, "cross attention: n_audio_ctx, d <= 1504, d_head 64");
    dim3 grid(n_head, Bn);
    launch_pdl(k_dec_cross_attn<T>, grid, dim3(256), 0, st, q, ldq, kbase, vbase, ld_kv, win_stride, out, d, n_ctx, fq);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
int sample_step(const float* logits, int ld, const SamplerArgs& a, int Bn, cudaStream_t st) {
        launch_pdl(k_logits_filter_argmax, dim3(Bn), dim3(kSampThreads), 0, st, logits, ld, a);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
int dec_advance(int* pos_ptr, int* step_ptr, int n_prompt, cudaStream_t st) {
    launch_pdl(k_dec_advance, dim3(1), dim3(1), 0, st, pos_ptr, step_ptr, n_prompt);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

#define SB_INST_D(T)                                                                                              \
    template int dec_embed<T>(const T*, const float*, const int*, const int*, float*, int, int, cudaStream_t);    \
    template int skinny_gemm<T>(const T*, int, const T*, int, int, int, int, const SkinnyEpilogue&, cudaStream_t); \
    template int dec_self_attn<T>(const T*, T*, T*, T*, const int*, int, int, int, int, cudaStream_t);             \
    template int dec_cross_attn<T>(const T*, int, const T*, const T*, int64_t, int64_t, T*, int, int, int, int, const FusedQ&, cudaStream_t);
SB_INST_D(__nv_bfloat16)
SB_INST_D(__half)

}  // namespace sb

// stage entry for parity tests / micro-benchmarks
extern "C" int sb_skinny_gemm_dev(int dtype, const void* X, int64_t ldx, const void* W, int64_t ldw, int Bn, int N, int K,
                                  const float* bias, int act, const float* residual, int64_t ldr, float* out32,
                                  int64_t ldo32, void* out16, int64_t ldo16, void* stream) {
    SB_CHECK_ARG(X && W && (out32 || out16), "null pointer");
    sb::SkinnyEpilogue e;
    e.bias = bias; e.act = act; e.residual = residual; e.ldr = (int)ldr; e.out32 = out32; e.ldo32 = (int)ldo32;
    e.out16 = out16; e.ldo16 = (int)ldo16;
    if (dtype == SB_DTYPE_F16)
        return sb::skinny_gemm<__half>((const __half*)X, (int)ldx, (const __half*)W, (int)ldw, Bn, N, K, e, (cudaStream_t)stream);
    return sb::skinny_gemm<__nv_bfloat16>((const __nv_bfloat16*)X, (int)ldx, (const __nv_bfloat16*)W, (int)ldw, Bn, N, K, e,
                                          (cudaStream_t)stream);
}
