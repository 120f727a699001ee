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

// NJ = 8-sequence column tiles per block (8: 64 sequences; 4: 32 sequences -- half the accumulators and activation
// ring, 96 registers, so a block can share an SM with two resident cross-attention blocks of the other decode lane).
template <typename T, int MT, int NJ, typename Sync>
__device__ __forceinline__ void skinny_body(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw, int Bn, int N,
                                            int K, const SkinnyEpilogue& ep, int tile, int chunk, unsigned char* smem, Sync& sync) {
    trace_begin(ep.trace);
    const int kb0 = 0, kb1 = K / 32;        // K % 32 == 0 enforced by the host
    float (*s_red)[16 * MT][65] = reinterpret_cast<float (*)[16 * MT][65]>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = tile * 16 * MT;
    const int b0 = chunk * (NJ * 8);
    const int nb = min(NJ * 8, Bn - b0);
    float acc[MT][NJ][4];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f; }
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
    constexpr int kWB = (MT == 1 && NJ == 8) ? 4 : 2;
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
    // the ring is only kWB k-blocks deep: the weight fragments of the later k-blocks of this warp are requested into L2
    // now (one prefetch per future load address of lanes t == 0 / 2: a 64-byte k-block row is two 32-byte sectors), so
    // that they do not pay the HBM latency one after the other inside the loop
    if (!(t & 1)) {
        for (int i = kWB; i < n_it; ++i) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w_lo[m] + (kb0 + warp + 8 * i) * 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(w_hi[m] + (kb0 + warp + 8 * i) * 32));
            }
        }
    }
    sync.wait();     // weights are immutable: only the activations depend on the previous stage
    // epilogue operands do not depend on the main loop: fetch them now so their L2 round trip is hidden
    constexpr int kTpb = 256 / (NJ * 8);            // threads per sequence in the epilogue
    constexpr int kRpt = 16 * MT / kTpb;            // output rows per thread
    const int e_n = tid / kTpb, e_rq = (tid % kTpb) * kRpt;
    float e_res[kRpt], e_bias[kRpt];
#pragma unroll
    for (int i = 0; i < kRpt; ++i) { e_res[i] = 0.f; e_bias[i] = 0.f; }
    if (e_n < nb) {
#pragma unroll
        for (int i = 0; i < kRpt; ++i) {
            const int row = row0 + e_rq + i;
            if (row < N) {
                if (ep.bias) e_bias[i] = __ldg(ep.bias + row);
                if (ep.residual) e_res[i] = __ldcg(ep.residual + (int64_t)(b0 + e_n) * ep.ldr + row);
            }
        }
    }
    constexpr int kXB = (MT == 1 && NJ == 8) ? 3 : 2;
    uint4 xq[kXB][NJ];
#pragma unroll
    for (int i = 0; i < kXB; ++i)
        if (i < n_it) {
            const int k1 = (kb0 + warp + 8 * i) * 32;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int n = j * 8 + g;
                xq[i][j] = n < nb ? __ldcg(reinterpret_cast<const uint4*>(xb + (int64_t)n * ldx + k1)) : make_uint4(0, 0, 0, 0);
            }
        }
    // rotate the two register rings with fully unrolled bodies of lcm(kWB, kXB) iterations so
    // every ring index is a compile-time constant
    constexpr int kRot = (MT == 1 && NJ == 8) ? 12 : 2;
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
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    MmaOpD<T>::mma(acc[m][j], alo[m].x, ahi[m].x, alo[m].y, ahi[m].y, xq[i % kXB][j].x, xq[i % kXB][j].y);
                    MmaOpD<T>::mma(acc[m][j], alo[m].z, ahi[m].z, alo[m].w, ahi[m].w, xq[i % kXB][j].z, xq[i % kXB][j].w);
                }
            }
            if (it + kXB < n_it) {
                const int k1 = (kb0 + warp + 8 * (it + kXB)) * 32;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
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
        for (int j = 0; j < NJ; ++j) {
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
        for (int i = 0; i < kRpt; ++i) {
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
    trace_end(ep.trace);
}

// ---- decoder LayerNorm of one row by one warp (two-pass in registers); with tok_emb != null the row is first formed
//      as token_embedding[tok] + positional_embedding[pos] and stored to x
template <typename T, int VPL, typename Sync>
__device__ __forceinline__ void ln_row_dec(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           T* __restrict__ out16, int row, int d, const T* __restrict__ tok_emb,
                                           const float* __restrict__ pos_emb, const int* __restrict__ next_tokens,
                                           const SeqState* __restrict__ state, Sync& sync) {
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
        const int pos = __ldcg(&state[row].pos);         // every sequence has its own position
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
// cross attention of one (sequence b, head h) by one 256-thread block.  Kc/Vc rows are strided (ld_kv) inside the
// fused cross-KV buffer [W*1500, L*2*d]; K then V are each streamed exactly once, straight into registers: the
// streaming loops contain no shared-memory staging and no block barrier, and every warp keeps 4 KB of independent
// 16-byte loads in flight (32 KB per block, 96 KB per SM at 3 blocks).  The K stream starts before the dependency
// wait.  History (profiles/r1_cross_attn_stream.md): the first version staged 128-key cp.async tiles through shared
// memory between two __syncthreads (one 16 KB tile in flight per block): 80 us per launch at 64 live sequences
// (3.7 TB/s) in the decode chain; this form 49.6 us (5.95 TB/s = 91 % of the measured HBM peak).  A fused
// "cross_attn_ln + query projection" prologue (two launches fewer per layer) was measured slower: streaming the
// 98 KB weight slice through one block costs more than the two launches.
// ------------------------------------------------------------------------------------------
constexpr int kXU = 8;
constexpr int kCrossSmem = 1504 * 4 + 8 * 4 + 8 * 64 * 4;

template <typename T, typename Sync>
__device__ __forceinline__ void cross_attn_body(const T* __restrict__ q, int ldq, const T* __restrict__ kbase,
                                                       const T* __restrict__ vbase, int64_t ld_kv, int64_t win_stride,
                                                       T* __restrict__ out, int d, int n_ctx, int h, int b, int bkv,
                                                       unsigned char* smem, Sync& sync, const TraceSlot& ts) {
    trace_begin(ts);
    float* s_sc = reinterpret_cast<float*>(smem);        // scores / probabilities
    float* s_red = s_sc + 1504;
    float (*s_o)[64] = reinterpret_cast<float (*)[64]>(s_red + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3, ch = lane & 7;
    // this lane's keys: key(it, u) = it * 256 + u * 32 + warp * 4 + sub
    const int key_l = warp * 4 + sub;
    const T* vp = vbase + (int64_t)bkv * win_stride + h * 64 + ch * 8 + (int64_t)key_l * ld_kv;      // bkv: whose cross-KV (b: query / output row)
    const int64_t u_stride = 32 * ld_kv;
    // ---- pass 1: scores on the tensor cores.  A 16-key x 64-dim slab is the A operand of four m16n8k16 MMAs whose
    // B operand carries q in column 0 (the other seven columns are zero): lane (g = lane / 4, t = lane % 4) loads
    // 16 bytes = dims [32 hf + 8 t, +8) of rows g and g + 8; q is laid out with the same k permutation, so the
    // products pair up (the skinny GEMM's fragment trick).  Per 2 KB of keys a warp issues 4 loads + 4 MMAs instead
    // of ~200 unpack / FFMA / shuffle instructions: the CUDA-core form of this pass was issue-bound (59 % of the SM's
    // issue slots at HBM speed).  Warp w takes slabs w, w + 8, ...; the ring keeps two slabs (4 KB per warp) in flight.
    const int g = lane >> 2, t = lane & 3;
    const int n_slab = (n_ctx + 15) >> 4;
    // load cursor: rows g / g + 8 of the next slab to request; slabs are requested in the order they are consumed
    // (warp, warp + 8, warp + 16, ...), so the cursor only ever advances by 128 rows
    const T* kcur = kbase + (int64_t)bkv * win_stride + h * 64 + t * 8 + (int64_t)(warp * 16 + g) * ld_kv;
    const int64_t row8 = 8 * ld_kv, slab_step = 128 * ld_kv;
    int krow = warp * 16 + g;
    auto slab_load = [&](uint4 (&dst)[4]) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        const bool v0 = krow < n_ctx, v1 = krow + 8 < n_ctx;
        dst[0] = v0 ? ldg_nc_v4(kcur) : z;
        dst[1] = v0 ? ldg_nc_v4(kcur + 32) : z;
        dst[2] = v1 ? ldg_nc_v4(kcur + row8) : z;
        dst[3] = v1 ? ldg_nc_v4(kcur + row8 + 32) : z;
        kcur += slab_step; krow += 128;
    };
    uint4 ka[4], kb2[4];
    {
        slab_load(ka);          // slab warp (rows past n_ctx load zeros)
        slab_load(kb2);         // slab warp + 8
    }
    sync.wait();           // (the K stream has started before the dependency wait: the encoder wrote K/V long ago)
    // q fragment, read straight from the query projection's output: only column 0 (g == 0) is non-zero
    uint32_t qb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qb[i] = 0u;
    if (g == 0) {
        const uint4* qg = reinterpret_cast<const uint4*>(q + (int64_t)b * ldq + h * 64 + t * 8);
        const uint4 q0 = __ldcg(qg), q1 = __ldcg(qg + 4);        // dims [8 t, +8) and [32 + 8 t, +8)
        qb[0] = q0.x; qb[1] = q0.y; qb[2] = q0.z; qb[3] = q0.w;
        qb[4] = q1.x; qb[5] = q1.y; qb[6] = q1.z; qb[7] = q1.w;
    }
    auto slab_scores = [&](const uint4 (&a)[4], int slab) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        MmaOpD<T>::mma(c, a[0].x, a[2].x, a[0].y, a[2].y, qb[0], qb[1]);
        MmaOpD<T>::mma(c, a[0].z, a[2].z, a[0].w, a[2].w, qb[2], qb[3]);
        MmaOpD<T>::mma(c, a[1].x, a[3].x, a[1].y, a[3].y, qb[4], qb[5]);
        MmaOpD<T>::mma(c, a[1].z, a[3].z, a[1].w, a[3].w, qb[6], qb[7]);
        if (t == 0) {
            const int r0 = slab * 16 + g;
            if (r0 < n_ctx) s_sc[r0] = c[0] * 0.125f;
            if (r0 + 8 < n_ctx) s_sc[r0 + 8] = c[2] * 0.125f;
        }
    };
#pragma unroll 1
    for (int slab = warp; slab < n_slab; slab += 16) {
        uint4 cur[4] = {ka[0], ka[1], ka[2], ka[3]};
        slab_load(ka);                                            // slab + 16
        slab_scores(cur, slab);
        if (slab + 8 < n_slab) {
            uint4 cur2[4] = {kb2[0], kb2[1], kb2[2], kb2[3]};
            slab_load(kb2);                                       // slab + 24
            slab_scores(cur2, slab + 8);
        }
    }
    uint4 ring[kXU];
    const T* ld_p;
    const int n_steps = ((n_ctx + 255) >> 8) * kXU;
    // the first V rows do not depend on the softmax: request them before the block-wide reductions
    ld_p = vp;
#pragma unroll
    for (int u = 0; u < kXU; ++u) { ring[u] = u * 32 + key_l < n_ctx ? ldg_nc_v4(ld_p) : make_uint4(0, 0, 0, 0); ld_p += u_stride; }
    __syncthreads();
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
    // ---- pass 2: O = P V; lane accumulates its 8 dims over its keys ----
    float oacc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) oacc[i] = 0.f;
#pragma unroll 1
    for (int s0 = 0; s0 < n_steps; s0 += kXU) {
#pragma unroll
        for (int u = 0; u < kXU; ++u) {
            const uint4 vv = ring[u];
            const int key = (s0 + u) * 32 + key_l;
            if (key + 256 < n_ctx) ring[u] = ldg_nc_v4(ld_p);
            ld_p += u_stride;
            if (key < n_ctx) {
                const float p = Op16<T>::to_f32(Op16<T>::from_f32(s_sc[key] * inv));
                float2 f;
                f = Op16<T>::unpack2(vv.x); oacc[0] = fmaf(p, f.x, oacc[0]); oacc[1] = fmaf(p, f.y, oacc[1]);
                f = Op16<T>::unpack2(vv.y); oacc[2] = fmaf(p, f.x, oacc[2]); oacc[3] = fmaf(p, f.y, oacc[3]);
                f = Op16<T>::unpack2(vv.z); oacc[4] = fmaf(p, f.x, oacc[4]); oacc[5] = fmaf(p, f.y, oacc[5]);
                f = Op16<T>::unpack2(vv.w); oacc[6] = fmaf(p, f.x, oacc[6]); oacc[7] = fmaf(p, f.y, oacc[7]);
            }
        }
    }
    sync.trigger();
    // reduce the four key sub-lanes of the warp, then the eight warps
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        oacc[i] += __shfl_xor_sync(0xffffffffu, oacc[i], 8);
        oacc[i] += __shfl_xor_sync(0xffffffffu, oacc[i], 16);
    }
    if (sub == 0) {
        *reinterpret_cast<float4*>(&s_o[warp][ch * 8]) = make_float4(oacc[0], oacc[1], oacc[2], oacc[3]);
        *reinterpret_cast<float4*>(&s_o[warp][ch * 8 + 4]) = make_float4(oacc[4], oacc[5], oacc[6], oacc[7]);
    }
    __syncthreads();
    if (tid < 32) {
        float o0 = 0.f, o1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { o0 += s_o[w][2 * tid]; o1 += s_o[w][2 * tid + 1]; }
        reinterpret_cast<uint32_t*>(out + (int64_t)b * d + h * 64)[tid] = Op16<T>::pack2(o0, o1);
    }
    trace_end(ts);
}

}  // namespace sb
