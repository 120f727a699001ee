// Device bodies of the decoder-step stages (the kernels of decoder_kernels.cu are thin wrappers).
// Every body works on one block's share with its shared memory passed as a pointer, and calls
// sync.wait() exactly where the data of the preceding stage is first needed: everything above that
// call only touches immutable operands (weights, the cross-KV cache written by the encoder), so it
// overlaps the predecessor's tail under programmatic dependent launch.  (A persistent per-step
// megakernel driving the same bodies through grid barriers was measured slower and removed:
// profiles/r1_mega_stage_trace.md.)
#pragma once
#include "common.cuh"
#include "decoder.cuh"

namespace sb {

struct PdlSync {
    __device__ __forceinline__ void wait() { pdl_wait(); }
    __device__ __forceinline__ void trigger() { pdl_trigger(); }
};

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
__device__ __forceinline__ void cp_async16_d(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz));
}

// ------------------------------------------------------------------------------------------
// skinny GEMM tile: Y[b0..b0+64, row0..row0+16) = X W^T (+bias, GELU, +residual).
// Fragment trick: both operands are read with 16-byte vector loads of 8 consecutive k; the
// k-permutation is the same for A and B so the products pair up correctly.
// ------------------------------------------------------------------------------------------
// MT = number of 16-row weight tiles per block (1, or 2 for the wide projections).
constexpr int kSkinnySmem = 8 * 32 * 65 * 4;       // MT = 2; MT = 1 needs half

template <typename T, int MT, typename Sync>
__device__ __forceinline__ void skinny_body(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw, int Bn, int N,
                                            int K, const SkinnyEpilogue& ep, int tile, int chunk, unsigned char* smem, Sync& sync) {
    const int kb0 = 0, kb1 = K / 32;        // K % 32 == 0 enforced by the host
    float (*s_red)[16 * MT][65] = reinterpret_cast<float (*)[16 * MT][65]>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = tile * 16 * MT;
    const int b0 = chunk * 64;
    const int nb = min(64, Bn - b0);
    float acc[MT][8][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f; }
    const T* w_lo[MT];
    const T* w_hi[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        w_lo[m] = W + (int64_t)min(row0 + 16 * m + g, N - 1) * ldw + t * 8;
        w_hi[m] = W + (int64_t)min(row0 + 16 * m + g + 8, N - 1) * ldw + t * 8;
    }
    const T* xb = X + (int64_t)b0 * ldx + t * 8;
    // this warp's k-blocks: kb0 + warp, kb0 + warp + 8, ...  Weight loads are issued kWB blocks ahead of
    // their use (the HBM stream must be in flight before anything waits on it); the activation
    // rows come from L2 and are kept kXB blocks ahead.
    constexpr int kWB = MT == 1 ? 4 : 2;
    const int n_it = (kb1 - kb0 - warp + 7) / 8;       // iterations of this warp (may be <= 0)
    uint4 wlo[MT][kWB], whi[MT][kWB];
#pragma unroll
    for (int i = 0; i < kWB; ++i)
        if (i < n_it) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                wlo[m][i] = ldg_nc_v4(w_lo[m] + (kb0 + warp + 8 * i) * 32);
                whi[m][i] = ldg_nc_v4(w_hi[m] + (kb0 + warp + 8 * i) * 32);
            }
        }
    sync.wait();     // weights are immutable: only the activations depend on the previous stage
    // epilogue operands do not depend on the main loop: fetch them now so their L2 round trip is hidden
    const int e_n = tid >> 2, e_rq = (tid & 3) * 4 * MT;
    float e_res[4 * MT], e_bias[4 * MT];
#pragma unroll
    for (int i = 0; i < 4 * MT; ++i) { e_res[i] = 0.f; e_bias[i] = 0.f; }
    if (e_n < nb) {
#pragma unroll
        for (int i = 0; i < 4 * MT; ++i) {
            const int row = row0 + e_rq + i;
            if (row < N) {
                if (ep.bias) e_bias[i] = __ldg(ep.bias + row);
                if (ep.residual) e_res[i] = __ldcg(ep.residual + (int64_t)(b0 + e_n) * ep.ldr + row);
            }
        }
    }
    constexpr int kXB = MT == 1 ? 3 : 2;
    uint4 xq[kXB][8];
#pragma unroll
    for (int i = 0; i < kXB; ++i)
        if (i < n_it) {
            const int k1 = (kb0 + warp + 8 * i) * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = j * 8 + g;
                xq[i][j] = n < nb ? __ldcg(reinterpret_cast<const uint4*>(xb + (int64_t)n * ldx + k1)) : make_uint4(0, 0, 0, 0);
            }
        }
    // rotate the two register rings with fully unrolled bodies of lcm(kWB, kXB) iterations so
    // every ring index is a compile-time constant
    constexpr int kRot = MT == 1 ? 12 : 2;
    for (int it0 = 0; it0 < n_it; it0 += kRot) {
#pragma unroll
        for (int i = 0; i < kRot; ++i) {
            const int it = it0 + i;
            if (it >= n_it) break;
            uint4 alo[MT], ahi[MT];
#pragma unroll
            for (int m = 0; m < MT; ++m) { alo[m] = wlo[m][i % kWB]; ahi[m] = whi[m][i % kWB]; }
            if (it + kWB < n_it) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    wlo[m][i % kWB] = ldg_nc_v4(w_lo[m] + (kb0 + warp + 8 * (it + kWB)) * 32);
                    whi[m][i % kWB] = ldg_nc_v4(w_hi[m] + (kb0 + warp + 8 * (it + kWB)) * 32);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    MmaOpD<T>::mma(acc[m][j], alo[m].x, ahi[m].x, alo[m].y, ahi[m].y, xq[i % kXB][j].x, xq[i % kXB][j].y);
                    MmaOpD<T>::mma(acc[m][j], alo[m].z, ahi[m].z, alo[m].w, ahi[m].w, xq[i % kXB][j].z, xq[i % kXB][j].w);
                }
            }
            if (it + kXB < n_it) {
                const int k1 = (kb0 + warp + 8 * (it + kXB)) * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = j * 8 + g;
                    xq[i % kXB][j] = n < nb ? __ldcg(reinterpret_cast<const uint4*>(xb + (int64_t)n * ldx + k1)) : make_uint4(0, 0, 0, 0);
                }
            }
        }
    }
    sync.trigger();   // main loop done: let the next kernel get scheduled and prefetch its weights
    // acc[m][j]: c0,c1 = (row 16m + g, batch j*8+2t, +1), c2,c3 = (row 16m + g+8, ...)
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s_red[warp][16 * m + g][j * 8 + 2 * t] = acc[m][j][0];
            s_red[warp][16 * m + g][j * 8 + 2 * t + 1] = acc[m][j][1];
            s_red[warp][16 * m + g + 8][j * 8 + 2 * t] = acc[m][j][2];
            s_red[warp][16 * m + g + 8][j * 8 + 2 * t + 1] = acc[m][j][3];
        }
    __syncthreads();
    // epilogue: thread -> (batch n = tid / 4, 4*MT consecutive rows r = (tid % 4) * 4 * MT)
    const int n = e_n, rq = e_rq;
    if (n < nb) {
        const int b = b0 + n;
#pragma unroll
        for (int i = 0; i < 4 * MT; ++i) {
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

// ---- decoder LayerNorm of one row by one warp (two-pass in registers); with tok_emb != null the row is first formed
//      as token_embedding[tok] + positional_embedding[pos] and stored to x
template <typename T, int VPL, typename Sync>
__device__ __forceinline__ void ln_row_dec(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           T* __restrict__ out16, int row, int d, const T* __restrict__ tok_emb,
                                           const float* __restrict__ pos_emb, const int* __restrict__ next_tokens,
                                           const int* __restrict__ pos_ptr, Sync& sync) {
    const int lane = threadIdx.x & 31;
    const int n4 = d >> 2;
    // immutable operands first: they are in flight while the predecessor is still finishing
    float4 g[VPL], bt[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) { g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + idx); bt[i] = __ldg(reinterpret_cast<const float4*>(beta) + idx); }
    }
    sync.wait();
    float4 v[VPL];
    float4* xr = reinterpret_cast<float4*>(x + (int64_t)row * d);
    if (tok_emb) {
        const int tok = __ldcg(next_tokens + row);
        const int pos = __ldcg(pos_ptr);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int idx = lane + 32 * i;
            if (idx < n4) {
                const uint2 u = __ldg(reinterpret_cast<const uint2*>(tok_emb + (int64_t)tok * d) + idx);
                const float4 p = __ldg(reinterpret_cast<const float4*>(pos_emb + (int64_t)pos * d) + idx);
                const float2 a = Op16<T>::unpack2(u.x), b = Op16<T>::unpack2(u.y);
                v[i] = make_float4(a.x + p.x, a.y + p.y, b.x + p.z, b.y + p.w);
                xr[idx] = v[i];
            } else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int idx = lane + 32 * i;
            v[i] = idx < n4 ? __ldcg(xr + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
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
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int idx = lane + 32 * i;
        if (idx < n4) {
            uint2 u;
            u.x = Op16<T>::pack2(v[i].x * rstd * g[i].x + bt[i].x, v[i].y * rstd * g[i].y + bt[i].y);
            u.y = Op16<T>::pack2(v[i].z * rstd * g[i].z + bt[i].z, v[i].w * rstd * g[i].w + bt[i].w);
            reinterpret_cast<uint2*>(out16 + (int64_t)row * d)[idx] = u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// cross attention of one (sequence b, head h) by one 256-thread block.  Kc/Vc rows are strided
// (ld_kv) inside the fused cross-KV buffer [W*1500, L*2*d]; K then V are each streamed exactly
// once through a cp.async double-buffered shared-memory tile.
// ------------------------------------------------------------------------------------------
constexpr int kXKeysPerTile = 128;
constexpr int kXLd = 72;   // padded row (elements)
constexpr int kCrossSmem = 2 * kXKeysPerTile * kXLd * 2 + 1504 * 4 + 64 * 4 + 8 * 4 + 4 * 64 * 4;

template <typename T, typename Sync>
__device__ __forceinline__ void cross_attn_body(const T* __restrict__ q, int ldq, const T* __restrict__ kbase,
                                                const T* __restrict__ vbase, int64_t ld_kv, int64_t win_stride,
                                                T* __restrict__ out, int d, int n_ctx, int h, int b,
                                                unsigned char* smem, Sync& sync) {
    T (*s_tile)[kXKeysPerTile * kXLd] = reinterpret_cast<T (*)[kXKeysPerTile * kXLd]>(smem);
    float* s_sc = reinterpret_cast<float*>(smem + 2 * kXKeysPerTile * kXLd * 2);   // scores
    float* s_q = s_sc + 1504;
    float* s_red = s_q + 64;
    float (*s_o)[64] = reinterpret_cast<float (*)[64]>(s_red + 8);
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
    sync.wait();
    if (tid < 32) {
        const float2 f = Op16<T>::unpack2(__ldcg(reinterpret_cast<const uint32_t*>(q + (int64_t)b * ldq + h * 64) + tid));
        s_q[2 * tid] = f.x * 0.125f; s_q[2 * tid + 1] = f.y * 0.125f;
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
    sync.trigger();
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

}  // namespace sb
