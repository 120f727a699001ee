// Decoder-step kernels (whisper.cpp decoder graph + sampler, SURVEY.md App. C.3 / C.4, rows
// a7 / a8 of 8(a)).  One "step" advances every live sequence of a decode lane by one token; every
// sequence has its own position and prompt (SeqState), finished sequences / empty slots are masked.
//
//   k_dec_ln              LayerNorm of the residual stream (+ token / positional embedding at layer 0)
//   k_skinny_gemm         Y[B,N] = X[B,K] W[N,K]^T (+bias, GELU, +residual); B <= 64 per pass,
//                         HBM-bound weight streaming; tensor cores via mma.sync with the weight
//                         rows as the M operand, 8 warps split K, smem reduction, fused epilogue
//   k_dec_self_attn       append K/V to the cache, causal attention over <= 448 positions
//   k_dec_cross_attn      attention over the 1500 cached encoder keys (the dominant HBM stream)
//   k_logits_filter_argmax  whisper_process_logits + whisper_sample_token(best) + the per-token
//                         bookkeeping of whisper_full (seek_delta / result_len / has_ts / stop)
#include "common.cuh"
#include "decoder.cuh"
#include "decoder_bodies.cuh"
#include <cstdlib>
#include <string>

namespace sb {
extern std::atomic<uint64_t> g_launches;

// ------------------------------------------------------------------------------------------
// stand-alone launches of the shared stage bodies (decoder_bodies.cuh), chained with PDL
// ------------------------------------------------------------------------------------------
// skinny GEMM.  grid.x = ceil(N/16) row tiles, grid.y = batch chunks of 8 NJ sequences.  256 threads.
template <typename T, int NJ>
__global__ void __launch_bounds__(256) __maxnreg__(NJ == 4 ? 96 : 232) k_skinny_gemm(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw,
                                                     int Bn, int N, int K, SkinnyEpilogue ep) {
    __shared__ __align__(16) unsigned char smem[kSkinnySmem / 2];
    PdlSync sync;
    skinny_body<T, 1, NJ>(X, ldx, W, ldw, Bn, N, K, ep, blockIdx.x, blockIdx.y, smem, sync);
}
// 32 weight rows per block: half as many blocks each ingest the shared [B, K] activation matrix (the L2 -> SM traffic of
// a launch is blocks x 98 KB), and a wide projection (QKV, FC1) leaves SMs free for the other decode lane
template <typename T>
__global__ void __launch_bounds__(256) k_skinny_gemm_w32(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw,
                                                         int Bn, int N, int K, SkinnyEpilogue ep) {
    extern __shared__ __align__(16) unsigned char smem_w32[];
    PdlSync sync;
    skinny_body<T, 2, 8>(X, ldx, W, ldw, Bn, N, K, ep, blockIdx.x, blockIdx.y, smem_w32, sync);
}

// 32 weight rows x 32 sequences at <= 128 registers: half the blocks of the narrow kernel for the wide projections of the d >= 1024
// models (Large-v3-Turbo: decode 274 -> 265 ms; Small, d = 768: 420 -> 430 ms, so it keeps the 16-row blocks)
template <typename T>
__global__ void __launch_bounds__(256, 2) k_skinny_gemm_w32n(const T* __restrict__ X, int ldx, const T* __restrict__ W, int ldw,
                                                             int Bn, int N, int K, SkinnyEpilogue ep) {
    extern __shared__ __align__(16) unsigned char smem_w32[];
    PdlSync sync;
    skinny_body<T, 2, 4>(X, ldx, W, ldw, Bn, N, K, ep, blockIdx.x, blockIdx.y, smem_w32, sync);
}

// decoder LayerNorm, one warp (= one block) per row so the rows spread over as many SMs:
// optional embedding (x = tok_emb[tok] + pos_emb[pos]) first, then LN -> 16-bit h.
template <typename T, int VPL>
__global__ void __launch_bounds__(32) k_dec_ln(float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                               T* __restrict__ out16, int d, const T* __restrict__ tok_emb,
                                               const float* __restrict__ pos_emb, const int* __restrict__ next_tokens,
                                               const SeqState* __restrict__ state, TraceSlot ts) {
    trace_begin(ts);
    struct S { __device__ __forceinline__ void wait() { pdl_wait(); pdl_trigger(); } } sync;
    ln_row_dec<T, VPL>(x, gamma, beta, out16, blockIdx.x, d, tok_emb, pos_emb, next_tokens, state, sync);
    trace_end(ts);
}

// self attention for one new token per sequence.  grid = (n_head, B), 128 threads: the four warps of a block split the
// <= 448 cached keys of one (sequence, head) -- scores with lane <-> key, P V with lane <-> dim pair and the keys
// interleaved over the warps -- so the two dependent sweeps over the cache are a quarter as long as with one warp.
template <typename T>
__global__ void __launch_bounds__(128) k_dec_self_attn(const T* __restrict__ qkv, T* __restrict__ kc, T* __restrict__ vc,
                                                       T* __restrict__ out, const SeqState* __restrict__ state, int honor_done,
                                                       int n_head, int d, int n_text_ctx, TraceSlot ts,
                                                       const int* __restrict__ row_slot, const int* __restrict__ row_pos) {
    trace_begin(ts);
    __shared__ float s_p[448];
    __shared__ float s_red[4];
    __shared__ float s_o[4][64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    // `done` and `pos` were written by the sampler of the previous step (many launches ago) and the cache
    // rows [0, pos) by earlier steps: all of it may be touched before the dependency wait, so the rows this block is
    // about to sweep are requested into L2 while the QKV projection in front of it is still finishing
    // prompt-prefill mode (row_slot != null): block row b is prompt token row_pos[b] of decode slot row_slot[b]; every row's
    // K / V was scattered into the cache by the launch in front of this one (a plain launch, no PDL overlap), nothing is
    // appended and nothing may be touched early
    const bool rows = row_slot != nullptr;
    const int sl = rows ? __ldg(row_slot + b) : b;
    const bool skip = !rows && honor_done && __ldcg(&state[b].done);   // finished sequences / empty slots are skipped
    const int pos = rows ? __ldg(row_pos + b) : min(__ldcg(&state[b].pos), n_text_ctx - 1);   // this sequence's new token; attends to [0, pos]
    const T* q = qkv + (int64_t)b * 3 * d + h * 64;
    T* kb = kc + ((int64_t)sl * n_text_ctx) * d + h * 64;
    T* vb = vc + ((int64_t)sl * n_text_ctx) * d + h * 64;
    if (!skip && !rows) {
        for (int k = tid; k < pos; k += 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (int64_t)k * d));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (int64_t)k * d + 32));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (int64_t)k * d));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (int64_t)k * d + 32));
        }
    }
    // this thread's first key row (k = tid < pos) is an old cache row: it is loaded into registers before the wait
    uint4 kpre[8];
    const bool has_pre = !skip && !rows && tid < pos;
    if (has_pre) {
        const uint4* kr = reinterpret_cast<const uint4*>(kb + (int64_t)tid * d);
#pragma unroll
        for (int c = 0; c < 8; ++c) kpre[c] = __ldcg(kr + c);
    }
    pdl_wait();
    pdl_trigger();
    if (skip) return;
    if (warp == 0 && !rows) {                            // append this step's K, V (each lane moves 2 elements)
        reinterpret_cast<uint32_t*>(kb + (int64_t)pos * d)[lane] = __ldcg(reinterpret_cast<const uint32_t*>(q + d) + lane);
        reinterpret_cast<uint32_t*>(vb + (int64_t)pos * d)[lane] = __ldcg(reinterpret_cast<const uint32_t*>(q + 2 * d) + lane);
    }
    float qf[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float2 f = Op16<T>::unpack2(__ldcg(reinterpret_cast<const uint32_t*>(q) + i));
        qf[2 * i] = f.x; qf[2 * i + 1] = f.y;
    }
    __syncthreads();                                     // the appended row is visible to the whole block
    const int n_keys = pos + 1;
    float mx = -INFINITY;
    auto score = [&](const uint4 (&u8)[8]) {
        float sc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = u8[c];
            float2 f;
            f = Op16<T>::unpack2(u.x); sc = fmaf(qf[c * 8 + 0], f.x, sc); sc = fmaf(qf[c * 8 + 1], f.y, sc);
            f = Op16<T>::unpack2(u.y); sc = fmaf(qf[c * 8 + 2], f.x, sc); sc = fmaf(qf[c * 8 + 3], f.y, sc);
            f = Op16<T>::unpack2(u.z); sc = fmaf(qf[c * 8 + 4], f.x, sc); sc = fmaf(qf[c * 8 + 5], f.y, sc);
            f = Op16<T>::unpack2(u.w); sc = fmaf(qf[c * 8 + 6], f.x, sc); sc = fmaf(qf[c * 8 + 7], f.y, sc);
        }
        return sc * 0.125f;
    };
    for (int k = tid; k < n_keys; k += 128) {
        uint4 u8[8];
        if (k == tid && has_pre) {
#pragma unroll
            for (int c = 0; c < 8; ++c) u8[c] = kpre[c];
        } else {
            const uint4* kr = reinterpret_cast<const uint4*>(kb + (int64_t)k * d);
#pragma unroll
            for (int c = 0; c < 8; ++c) u8[c] = __ldcg(kr + c);
        }
        const float sc = score(u8);
        s_p[k] = sc;
        mx = fmaxf(mx, sc);
    }
    // the first batch of value rows does not depend on the softmax: request it before the block-wide reductions
    // P V: warp w takes keys w, w + 4, ...; lane <-> 2 output dims; 16 value loads are issued before their FMAs
    uint32_t vv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int k = min(warp + 4 * i, n_keys - 1);
        vv[i] = __ldcg(reinterpret_cast<const uint32_t*>(vb + (int64_t)k * d) + lane);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int k = tid; k < n_keys; k += 128) {
        const float p = __expf(s_p[k] - mx);
        s_p[k] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    const float inv = 1.0f / (((s_red[0] + s_red[1]) + s_red[2]) + s_red[3]);
    // probabilities rounded to the operand type like ggml's mul_mat
    float o0 = 0.f, o1 = 0.f;
    for (int k0 = warp; k0 < n_keys; k0 += 64) {
        if (k0 != warp) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = min(k0 + 4 * i, n_keys - 1);
                vv[i] = __ldcg(reinterpret_cast<const uint32_t*>(vb + (int64_t)k * d) + lane);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = k0 + 4 * i;
            if (k < n_keys) {
                const float p = Op16<T>::to_f32(Op16<T>::from_f32(s_p[k] * inv));
                const float2 f = Op16<T>::unpack2(vv[i]);
                o0 = fmaf(p, f.x, o0); o1 = fmaf(p, f.y, o1);
            }
        }
    }
    s_o[warp][2 * lane] = o0; s_o[warp][2 * lane + 1] = o1;
    __syncthreads();
    if (warp == 0) {
        const float a0 = ((s_o[0][2 * lane] + s_o[1][2 * lane]) + s_o[2][2 * lane]) + s_o[3][2 * lane];
        const float a1 = ((s_o[0][2 * lane + 1] + s_o[1][2 * lane + 1]) + s_o[2][2 * lane + 1]) + s_o[3][2 * lane + 1];
        reinterpret_cast<uint32_t*>(out + (int64_t)b * d + h * 64)[lane] = Op16<T>::pack2(a0, a1);
    }
    trace_end(ts);
}

// cross attention.  grid = (n_head, B), 256 threads, 3 blocks per SM (80 registers: the K / V register rings).
template <typename T>
__global__ void __launch_bounds__(256, 3) k_dec_cross_attn(const T* __restrict__ q, int ldq, const T* __restrict__ kbase,
                                                           const T* __restrict__ vbase, int64_t ld_kv, int64_t win_stride,
                                                           T* __restrict__ out, const SeqState* __restrict__ state, int d,
                                                           int n_ctx, TraceSlot ts, const int* __restrict__ row_slot) {
    __shared__ __align__(16) unsigned char smem[kCrossSmem];
    // a finished sequence no longer needs its 2 x 1500 x 64 keys/values streamed: `done` was written by the
    // sampler of an earlier step (many launches ago), so it may be read before the dependency wait
    if (state && __ldcg(&state[blockIdx.y].done)) { pdl_wait(); pdl_trigger(); return; }
    PdlSync sync;
    // prompt-prefill mode: query row blockIdx.y attends over the cross-KV of decode slot row_slot[blockIdx.y]
    const int kv = row_slot ? __ldg(row_slot + blockIdx.y) : blockIdx.y;
    cross_attn_body<T>(q, ldq, kbase, vbase, ld_kv, win_stride, out, d, n_ctx, blockIdx.x, blockIdx.y, kv, smem, sync, ts);
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
        if (x == -INFINITY) return;                       // a suppressed (non-speech) text token
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
    if (b == 0 && threadIdx.x == 0 && a.tick) *a.tick = __ldcg(a.tick) + 1;      // lane step counter (launch trace only)
    SeqState st;
    {
        static_assert(sizeof(SeqState) == 64, "SeqState is read as four int4");
        const int4* sp4 = reinterpret_cast<const int4*>(a.state + b);
        int4* dp4 = reinterpret_cast<int4*>(&st);
        dp4[0] = __ldcg(sp4); dp4[1] = __ldcg(sp4 + 1); dp4[2] = __ldcg(sp4 + 2); dp4[3] = __ldcg(sp4 + 3);
    }
    if (st.done) return;                   // finished sequence or empty slot
    if (st.idx < st.n_prompt - 1) {        // still feeding the prompt: queue its next token
        if (threadIdx.x == 0) {
            a.next_tokens[b] = __ldcg(a.prompt + (int64_t)b * kMaxPrompt + st.idx + 1);
            // after the language-detect [sot] the real prompt starts again at position 0
            a.state[b].pos = (st.restart && st.idx == 0) ? 0 : st.pos + 1;
            a.state[b].idx = st.idx + 1;
        }
        return;
    }
    const int step = st.n_tok;             // index of the token being sampled (i in whisper_full)
    const SpecialIds sp = a.sp;
    const int V = a.n_vocab;
    // Stage the whole vocabulary row (207 KB, L2-resident: the logits GEMM just wrote it) in shared memory with one
    // deep burst of 16-byte cp.async (13 in flight per thread).  The sweeps below then run out of shared memory; reading
    // the row three times through L2 with one dependent 4-byte load per iteration cost 50-67 us per launch.
    extern __shared__ __align__(16) float s_row[];
    {
        const float* src = logits + (int64_t)b * ld;
        const int n16 = (V + 3) >> 2;                     // ld is a multiple of 8 floats: the padded tail is readable
        for (int i = threadIdx.x; i < n16; i += kSampThreads)
            cp_async16_d((uint32_t)__cvta_generic_to_shared(s_row + 4 * i), src + 4 * i, true);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
        __syncthreads();
    }
    const float temp = a.temperature ? __ldcg(a.temperature + b) : 0.f;
    if (temp > 0.f) {                  // whisper_process_logits: logits[i] /= temperature, before every rule
        for (int i = threadIdx.x; i < V; i += kSampThreads) s_row[i] = s_row[i] / temp;
        __syncthreads();
    }
    if (a.nst_ids) {                   // whisper_process_logits, suppress_nst: logits[id] = -INFINITY
        for (int i = threadIdx.x; i < a.n_nst; i += kSampThreads) s_row[__ldg(a.nst_ids + i)] = -INFINITY;
        __syncthreads();
    }
    const float* lg = s_row;
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
            const float x = lg[id];
            all.add(x);
            mx_text = fmaxf(mx_text, x);
        }
    for (int id = sp.eot + threadIdx.x; id < V; id += kSampThreads) {
        if (mk.suppressed(id)) continue;
        const float x = lg[id];
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
            const float x = lg[id];
            if (x > bv) { b2 = bv; bv = x; bi = id; }
            else if (x > b2) b2 = x;
        }
    for (int id = sp.eot + threadIdx.x; id < V; id += kSampThreads) {
        if (mk.suppressed(id) || (force_ts && id < sp.beg)) continue;
        const float x = lg[id];
        if (x > bv) { b2 = bv; bv = x; bi = id; }
        else if (x > b2) b2 = x;
    }
    float gv; int gi;
    red.argmax(bv, bi, gv, gi);
    const float sv = red.max_f(bi == gi ? b2 : bv);
    // whisper_token_data.tid (whisper_sample_token): the most probable timestamp token of this step, first maximum in
    // ascending id; probabilities that underflow to 0 never win and leave tid = 0.  Segment start times come from it.
    float tv = -INFINITY; int ti = 0x7fffffff;
    if (!a.no_timestamps)
        for (int id = sp.beg + threadIdx.x; id < V; id += kSampThreads) {
            if (mk.suppressed(id)) continue;
            const float x = lg[id];
            if (x > tv) { tv = x; ti = id; }
        }
    float gtv; int gti;
    red.argmax(tv, ti, gtv, gti);
    // ---- temperature > 0: draw the token (whisper_sample_token(best = false)) ----
    // probs[i] = expf(logprobs[i]) in f32 for the allowed ids, 0 otherwise; std::discrete_distribution normalises them in
    // double, forms the cumulative sums and returns the first index whose cumulative probability is >= u.  Each thread
    // owns a contiguous run of ids, the runs are combined with a block scan in double.
    __shared__ double sd[kSampThreads / 32];
    __shared__ int s_pick;
    int drawn = -1;
    if (temp > 0.f && a.rng_u) {
        const int C = (V + kSampThreads - 1) / kSampThreads;
        const int i0 = threadIdx.x * C, i1 = min(V, i0 + C);
        auto weight = [&](int id) -> float {
            if (id < sp.eot) { if (text_off || force_ts || id == blank_off) return 0.f; }
            else if (mk.suppressed(id) || (force_ts && id < sp.beg)) return 0.f;
            return expf(lg[id] - lse);
        };
        double part = 0.0;
        for (int id = i0; id < i1; ++id) part += (double)weight(id);
        auto block_sum_d = [&](double v) -> double {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = v;
            __syncthreads();
            double r = 0.0;
            for (int i = 0; i < kSampThreads / 32; ++i) r += sd[i];
            return r;
        };
        const double S = block_sum_d(part);
        double q = 0.0;
        for (int id = i0; id < i1; ++id) q += (double)weight(id) / S;
        // exclusive scan of q over the block
        double incl = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double n = __shfl_up_sync(0xffffffffu, incl, o); if ((threadIdx.x & 31) >= o) incl += n; }
        __syncthreads();
        if ((threadIdx.x & 31) == 31) sd[threadIdx.x >> 5] = incl;
        if (threadIdx.x == 0) s_pick = V - 1;            // libstdc++ pins the last cumulative value to 1.0
        __syncthreads();
        double base = incl - q;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += sd[w];
        const double u = __ldcg(a.rng_u + (int64_t)b * a.n_max + step);
        if (base < u && u <= base + q) {
            double c = base;
            int pick = i1 - 1;
            for (int id = i0; id < i1; ++id) { c += (double)weight(id) / S; if (c >= u) { pick = id; break; } }
            atomicMin(&s_pick, pick);
        }
        __syncthreads();
        drawn = s_pick;
    }
    if (threadIdx.x != 0) return;
    // after the timestamp-mass rule the text tokens are gone and the distribution is renormalised over the timestamps
    const float lse_eff = force_ts ? (logf(ts.s) + ts.m) : lse;
    const int tid = (gtv > -INFINITY && expf(gtv - lse_eff) > 0.f) ? gti : 0;

    int tok = drawn >= 0 ? drawn : gi;
    if (a.forced) {
        const int f = __ldcg(a.forced + (int64_t)b * a.n_max + step);
        if (f >= 0) tok = f;
    }
    a.tokens_out[(int64_t)b * a.n_max + step] = tok;
    if (a.margins_out) a.margins_out[(int64_t)b * a.n_max + step] = gv - sv;
    if (a.tids_out) a.tids_out[(int64_t)b * a.n_max + step] = tid;
    const float plog = lg[tok] - lse;                 // logprobs[id] of the token that was taken
    if (a.plogs_out) a.plogs_out[(int64_t)b * a.n_max + step] = plog;
    a.next_tokens[b] = tok;
    st.prev = st.last; st.last = tok; st.n_tok += 1;
    st.sum_logprob += plog;
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
    if (!stop && st.pos + 1 >= a.n_text_ctx) { if (st.result_len == 0) st.failed = 1; stop = true; }   // text context exhausted
    if (stop) st.done = 1;
    st.pos += 1;
    a.state[b] = st;
}

// language auto-detect (whisper.cpp whisper_lang_auto_detect_with_state): the decoder has seen [sot] only;
// among the language tokens the one with the largest raw logit (= largest softmax probability; lowest id on
// ties) becomes prompt token 1 of the sequence.  One warp per sequence.
__global__ void __launch_bounds__(32) k_lang_detect(const float* __restrict__ logits, int ld, int* __restrict__ prompt,
                                                    const SeqState* __restrict__ state, int* __restrict__ lang_out, SpecialIds sp) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x, lane = threadIdx.x;
    const int idx = __ldcg(&state[b].idx), slot = __ldcg(&state[b].lang_slot);
    if (__ldcg(&state[b].done) || idx != 0 || slot < 0) return;
    if (__ldcg(prompt + (int64_t)b * kMaxPrompt + slot) >= 0) return;       // language given (or detected on an earlier window)
    const float* lg = logits + (int64_t)b * ld + sp.lang_first;
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = lane; i < sp.num_languages; i += 32) {
        const float x = __ldcg(lg + i);
        if (x > bv) { bv = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
    }
    if (lane == 0) {
        prompt[(int64_t)b * kMaxPrompt + slot] = sp.lang_first + bi;
        if (lang_out) lang_out[b] = bi;
    }
}

// freshly assigned windows -> their decode slots.  One block per item; runs on the lane's stream between two steps.
__global__ void __launch_bounds__(64) k_slot_init(const SlotInit* __restrict__ items, SeqState* __restrict__ state,
                                                  int* __restrict__ next_tokens, int* __restrict__ prompt, int* __restrict__ lang_out,
                                                  float* __restrict__ temperature) {
    const SlotInit& it = items[blockIdx.x];
    const int slot = it.slot;
    for (int i = threadIdx.x; i < kMaxPrompt; i += 64) prompt[(int64_t)slot * kMaxPrompt + i] = it.prompt[i];
    if (threadIdx.x == 0) {
        state[slot] = it.state; next_tokens[slot] = it.next_token;
        if (lang_out) lang_out[slot] = -1;
        if (temperature) temperature[slot] = it.temperature;
    }
}

// ---- prompt prefill: all prompt tokens of the freshly assigned windows in one pass ------------------------------
// x[r] = token_embedding[tok[r]] + positional_embedding[pos[r]]   (f32 rows for the tcgen05 GEMM chain)
template <typename T>
__global__ void __launch_bounds__(128) k_prefill_embed(const T* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                                       const int* __restrict__ row_tok, const int* __restrict__ row_pos,
                                                       float* __restrict__ x, int d) {
    const int r = blockIdx.x;
    const int tok = __ldg(row_tok + r), pos = __ldg(row_pos + r);
    for (int i = threadIdx.x; i < d / 4; i += 128) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(tok_emb + (int64_t)tok * d) + i);
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos_emb + (int64_t)pos * d) + i);
        const float2 a = Op16<T>::unpack2(u.x), b = Op16<T>::unpack2(u.y);
        reinterpret_cast<float4*>(x + (int64_t)r * d)[i] = make_float4(a.x + p.x, a.y + p.y, b.x + p.z, b.y + p.w);
    }
}
// K, V columns of qkv[r] -> self-KV cache row (slot[r], pos[r]) of one layer
template <typename T>
__global__ void __launch_bounds__(128) k_prefill_kv_scatter(const T* __restrict__ qkv, const int* __restrict__ row_slot,
                                                            const int* __restrict__ row_pos, T* __restrict__ kc, T* __restrict__ vc,
                                                            int d, int n_text_ctx) {
    const int r = blockIdx.x;
    const int64_t dst = ((int64_t)__ldg(row_slot + r) * n_text_ctx + __ldg(row_pos + r)) * d;
    const uint4* src = reinterpret_cast<const uint4*>(qkv + (int64_t)r * 3 * d);
    const int n16 = d / 8;
    for (int i = threadIdx.x; i < n16; i += 128) {
        reinterpret_cast<uint4*>(kc + dst)[i] = src[n16 + i];
        reinterpret_cast<uint4*>(vc + dst)[i] = src[2 * n16 + i];
    }
}

// ---- launchers -------------------------------------------------------------------------
template <typename T>
int prefill_embed(const T* tok_emb, const float* pos_emb, const int* row_tok, const int* row_pos, float* x, int rows, int d, cudaStream_t st) {
    k_prefill_embed<T><<<rows, 128, 0, st>>>(tok_emb, pos_emb, row_tok, row_pos, x, d);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int prefill_kv_scatter(const T* qkv, const int* row_slot, const int* row_pos, T* kc, T* vc, int rows, int d, int n_text_ctx, cudaStream_t st) {
    k_prefill_kv_scatter<T><<<rows, 128, 0, st>>>(qkv, row_slot, row_pos, kc, vc, d, n_text_ctx);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int skinny_gemm(const T* X, int ldx, const T* W, int ldw, int Bn, int N, int K, const SkinnyEpilogue& ep, cudaStream_t st) {
    SB_CHECK_ARG(K % 32 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "skinny gemm: K % 32 and 16-byte row alignment required");
    constexpr int w32_min_n = 2048;                 // 32-row weight tiles for the wide projections (QKV, FC1, logits)
    SkinnyEpilogue ept = ep; ept.trace = g_trace_next; g_trace_next = TraceSlot();
    const bool narrow = Bn <= 32;      // 32-sequence blocks (half the registers) for a decode lane of <= 32
    const int chunks = ceil_div(Bn, narrow ? 32 : 64);
    const bool narrow_w32 = K >= 1024;
    if (N >= w32_min_n && narrow && narrow_w32) {
        SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_skinny_gemm_w32n<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkinnySmem)); });
        launch_pdl(k_skinny_gemm_w32n<T>, dim3(ceil_div(N, 32), chunks), dim3(256), (size_t)kSkinnySmem, st, X, ldx, W, ldw, Bn, N, K, ept);
        g_launches += 1;
        SB_CUDA_CHECK(cudaGetLastError());
        return SB_OK;
    }
    if (N >= w32_min_n && !narrow) {
        SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_skinny_gemm_w32<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSkinnySmem)); });
        launch_pdl(k_skinny_gemm_w32<T>, dim3(ceil_div(N, 32), chunks), dim3(256), (size_t)kSkinnySmem, st, X, ldx, W, ldw, Bn, N, K, ept);
        g_launches += 1;
        SB_CUDA_CHECK(cudaGetLastError());
        return SB_OK;
    }
    dim3 grid(ceil_div(N, 16), chunks);
    if (narrow) launch_pdl(k_skinny_gemm<T, 4>, grid, dim3(256), 0, st, X, ldx, W, ldw, Bn, N, K, ept);
    else launch_pdl(k_skinny_gemm<T, 8>, grid, dim3(256), 0, st, X, ldx, W, ldw, Bn, N, K, ept);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int dec_ln(float* x, const float* gamma, const float* beta, T* out16, int rows, int d, const T* tok_emb, const float* pos_emb,
           const int* next_tokens, const SeqState* state, cudaStream_t st) {
    SB_CHECK_ARG(d % 4 == 0 && d <= 1536, "decoder layernorm: d % 4, d <= 1536");
    const TraceSlot ts = g_trace_next; g_trace_next = TraceSlot();
    if (d <= 768) launch_pdl(k_dec_ln<T, 6>, dim3(rows), dim3(32), 0, st, x, gamma, beta, out16, d, tok_emb, pos_emb, next_tokens, state, ts);
    else if (d <= 1280) launch_pdl(k_dec_ln<T, 10>, dim3(rows), dim3(32), 0, st, x, gamma, beta, out16, d, tok_emb, pos_emb, next_tokens, state, ts);
    else launch_pdl(k_dec_ln<T, 12>, dim3(rows), dim3(32), 0, st, x, gamma, beta, out16, d, tok_emb, pos_emb, next_tokens, state, ts);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int dec_self_attn(const T* qkv, T* kc, T* vc, T* out, const SeqState* state, int honor_done, int Bn, int n_head, int d,
                  int n_text_ctx, cudaStream_t st, const int* row_slot, const int* row_pos) {
    SB_CHECK_ARG(n_text_ctx <= 448 && d == n_head * 64, "self attention: n_text_ctx <= 448, d_head 64");
    const TraceSlot ts = g_trace_next; g_trace_next = TraceSlot();
    if (row_slot) k_dec_self_attn<T><<<dim3(n_head, Bn), dim3(128), 0, st>>>(qkv, kc, vc, out, state, honor_done, n_head, d, n_text_ctx, ts, row_slot, row_pos);
    else launch_pdl(k_dec_self_attn<T>, dim3(n_head, Bn), dim3(128), 0, st, qkv, kc, vc, out, state, honor_done, n_head, d, n_text_ctx, ts, row_slot, row_pos);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
template <typename T>
int dec_cross_attn(const T* q, int ldq, const T* kbase, const T* vbase, int64_t ld_kv, int64_t win_stride, T* out,
                   const SeqState* state, int Bn, int n_head, int d, int n_ctx, cudaStream_t st, const int* row_slot) {
    SB_CHECK_ARG(n_ctx <= 1504 && d == n_head * 64 && d <= 1504 && d % 32 == 0 && ldq % 8 == 0 && ld_kv % 8 == 0,
                 "cross attention: n_audio_ctx, d <= 1504, d_head 64, 16-byte aligned rows");
    const TraceSlot ts = g_trace_next; g_trace_next = TraceSlot();
    dim3 grid(n_head, Bn);
    if (row_slot) k_dec_cross_attn<T><<<grid, dim3(256), 0, st>>>(q, ldq, kbase, vbase, ld_kv, win_stride, out, state, d, n_ctx, ts, row_slot);
    else launch_pdl(k_dec_cross_attn<T>, grid, dim3(256), 0, st, q, ldq, kbase, vbase, ld_kv, win_stride, out, state, d, n_ctx, ts, row_slot);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
int sample_step(const float* logits, int ld, const SamplerArgs& a, int Bn, cudaStream_t st) {
    const size_t smem = (size_t)((a.n_vocab + 3) / 4) * 16;
    SB_CHECK_ARG(smem <= 226 * 1024 && ld % 4 == 0, "sampler: vocabulary row must fit shared memory (<= 57856 tokens)");
    SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_logits_filter_argmax, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); });
    launch_pdl(k_logits_filter_argmax, dim3(Bn), dim3(kSampThreads), smem, st, logits, ld, a);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
int lang_detect_step(const float* logits, int ld, int* prompt, const SeqState* state, int* lang_out, SpecialIds sp, int Bn,
                     cudaStream_t st) {
    SB_CHECK_ARG(sp.num_languages > 0, "language auto-detect needs a multilingual model");
    launch_pdl(k_lang_detect, dim3(Bn), dim3(32), 0, st, logits, ld, prompt, state, lang_out, sp);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
int slot_init(const SlotInit* items, int n, SeqState* state, int* next_tokens, int* prompt, int* lang_out, float* temperature,
              cudaStream_t st) {
    if (n <= 0) return SB_OK;
    k_slot_init<<<n, 64, 0, st>>>(items, state, next_tokens, prompt, lang_out, temperature);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

#define SB_INST_D(T)                                                                                              \
    template int skinny_gemm<T>(const T*, int, const T*, int, int, int, int, const SkinnyEpilogue&, cudaStream_t); \
    template int dec_ln<T>(float*, const float*, const float*, T*, int, int, const T*, const float*, const int*, const SeqState*, cudaStream_t); \
    template int dec_self_attn<T>(const T*, T*, T*, T*, const SeqState*, int, int, int, int, int, cudaStream_t, const int*, const int*); \
    template int dec_cross_attn<T>(const T*, int, const T*, const T*, int64_t, int64_t, T*, const SeqState*, int, int, int, int, cudaStream_t, const int*); \
    template int prefill_embed<T>(const T*, const float*, const int*, const int*, float*, int, int, cudaStream_t);       \
    template int prefill_kv_scatter<T>(const T*, const int*, const int*, T*, T*, int, int, int, cudaStream_t);
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
