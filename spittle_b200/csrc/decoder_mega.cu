// One decoder step (whisper.cpp decoder graph, SURVEY.md App. C.3, row a7 of 8(a)) as ONE persistent
// cooperative kernel: embedding -> L x { LN, QKV, self-attention, O, LN, Q, cross-attention, O, LN,
// MLP } -> final LN.  The stand-alone path issues ~11 launches per layer whose cost is launch /
// dependency latency, not work (a projection streams 1-5 MB of weights in < 1 us of HBM time).
// Here one CTA per SM stays resident for the whole step and the stages are separated by grid
// barriers; each stage's prefetch of immutable operands (weight fragments, the first cross-K
// tile) is issued BEFORE the barrier wait, so the barrier latency hides behind the HBM fetch.
//
// The stage bodies are the ones of the stand-alone kernels (decoder_bodies.cuh): results are
// bit-identical between the two schedulers.
//
// Finished sequences (SeqState.done) are compacted out of the attention stages: the cross
// attention streams 2 x 1500 x d x 2 B per sequence and layer -- the dominant HBM traffic of a
// step -- only for sequences that are still decoding.
#include "common.cuh"
#include "decoder.cuh"
#include "decoder_bodies.cuh"

namespace sb {
extern std::atomic<uint64_t> g_launches;

// optional per-stage timeline of CTA 0 (env SB_MEGA_TRACE=1; read back with sb_debug_mega_trace)
__device__ unsigned long long g_mega_trace[1024];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct GridSync {
    unsigned* ctr;
    unsigned target;     // counter value that completes the barrier closing the previous stage
    bool waited;
    int trace_idx;       // >= 0: CTA 0 records [2*stage] = dependency wait done, [2*stage+1] = own work done
    __device__ __forceinline__ void wait() {
        if (waited) return;
        waited = true;
        if (threadIdx.x == 0) {
            unsigned v;
            long long t0 = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                if (v >= target) break;
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 8000000000LL) __trap();    // ~4 s: a lost CTA must not hang the GPU
            }
            if (trace_idx >= 0 && trace_idx < 511) g_mega_trace[2 * trace_idx] = gtimer();
        }
        __syncthreads();
    }
    __device__ __forceinline__ void trigger() {}
    // close the current stage: publish this CTA's stores, then open the next stage
    __device__ __forceinline__ void arrive() {
        wait();                 // a CTA without work in this stage still has to pass the previous barrier
        __syncthreads();
        if (threadIdx.x == 0) {
            if (trace_idx >= 0 && trace_idx < 511) g_mega_trace[2 * trace_idx + 1] = gtimer();
            __threadfence();
            atomicAdd(ctr, 1u);
        }
        target += gridDim.x;
        waited = false;
        if (trace_idx >= 0) ++trace_idx;
    }
};

constexpr int kMegaMaxB = 256;
constexpr int kMegaWork = kCrossSmem > kSkinnySmem / 2 ? kCrossSmem : kSkinnySmem / 2;    // stage scratch (stages overlay each other)
constexpr int kMegaSmem = 1040 + kMegaWork;

template <typename T, int VPL>
__global__ void __launch_bounds__(256, 1) k_dec_step_mega(const DecLayerDev* __restrict__ layers, DecStepArgs a) {
    extern __shared__ __align__(128) unsigned char mega_smem[];
    int* s_active = reinterpret_cast<int*>(mega_smem);            // [kMegaMaxB] compacted live sequences
    int* s_nact = s_active + kMegaMaxB;                           // [1]
    unsigned char* smem = mega_smem + 1040 + 112;                 // 128-byte aligned work area
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const int d = a.d, H = a.n_head, Bn = a.Bn;
    T* const h16 = reinterpret_cast<T*>(a.h);
    T* const qkv = reinterpret_cast<T*>(a.qkv);
    T* const att = reinterpret_cast<T*>(a.att);
    T* const qb = reinterpret_cast<T*>(a.q);
    T* const mlp = reinterpret_cast<T*>(a.mlp);

    GridSync sync{a.barrier, 0u, true, (a.trace && blockIdx.x == 0) ? 0 : -1};      // stage 0 only reads what earlier launches wrote
    if (sync.trace_idx == 0 && threadIdx.x == 0) g_mega_trace[0] = gtimer();
    const int pos = __ldcg(a.pos_ptr);
    // live-sequence list (identical in every CTA: `done` was written by the sampler of the previous step)
    if (warp == 0) {
        int n = 0;
        for (int b0 = 0; b0 < Bn; b0 += 32) {
            const int b = b0 + lane;
            const bool live = b < Bn && !(a.state && __ldcg(&a.state[b].done));
            const unsigned m = __ballot_sync(0xffffffffu, live);
            if (live) s_active[n + __popc(m & ((1u << lane) - 1u))] = b;
            n += __popc(m);
        }
        if (lane == 0) *s_nact = n;
    }
    __syncthreads();
    const int n_act = *s_nact;

    const int chunks = (Bn + 63) / 64;
    // Stage table: every stage kind has ONE call site of its body (the bodies are large; inlining a
    // body per projection would spill), the per-stage operands are selected by the switch below.
    //   0 LN1(+embed)  1 QKV  2 self-attn  3 O  4 LN2  5 Q  6 cross-attn  7 O  8 LN3  9 FC1  10 FC2
    // Projections whose consumer is a LayerNorm (3, 7, 10) may be
    // split along K (a.ks[kind] > 1): they store f32 slices to a.part and the consumer adds them.
    const int n_stages = a.n_layer * 11 + 1;        // + final LayerNorm
#pragma unroll 1
    for (int sidx = 0; sidx < n_stages; ++sidx) {
        const int l = sidx / 11, kind = sidx - l * 11;
        const bool last = sidx == n_stages - 1;
        const DecLayerDev& L = layers[last ? l - 1 : l];
        if (last || kind == 0 || kind == 4 || kind == 8) {
            const float* g = last ? a.lnf_g : kind == 0 ? L.ln1_g : kind == 4 ? L.ln2_g : L.ln3_g;
            const float* bt = last ? a.lnf_b : kind == 0 ? L.ln1_b : kind == 4 ? L.ln2_b : L.ln3_b;
            const bool embed = sidx == 0;
            // producer of the residual stream this LayerNorm reads
            int np = 0; const float* pbias = nullptr;
            if (last) { np = a.ks[10]; pbias = L.fc2_b; }
            else if (kind == 0 && l > 0) { np = a.ks[10]; pbias = layers[l - 1].fc2_b; }
            else if (kind == 4) { np = a.ks[3]; pbias = L.o_b; }
            else if (kind == 8) { np = a.ks[7]; pbias = L.co_b; }
            if (np <= 1) np = 0;
            bool any = false;
            for (int row = cta + G * warp; row < Bn; row += G * 8) {
                ln_row_mega<T, VPL>(a.x, g, bt, h16, row, d, embed ? reinterpret_cast<const T*>(a.tok_emb) : nullptr, a.pos_emb,
                                    a.next_tokens, a.pos_ptr, a.part, np, a.part_stride, pbias, sync);
                any = true;
            }
            (void)any;
        } else if (kind == 2) {
            // self attention: one warp per live (sequence, head)
            sync.wait();
            float* s_p = reinterpret_cast<float*>(smem) + warp * 448;
            T* kc = reinterpret_cast<T*>(L.kself);
            T* vc = reinterpret_cast<T*>(L.vself);
            for (int it = cta + G * warp; it < n_act * H; it += G * 8) {
                const int b = s_active[it / H], h = it % H;
                self_attn_warp<T>(qkv, kc, vc, att, pos, b, h, d, a.n_text_ctx, s_p);
            }
        } else if (kind == 6) {
            // cross attention: one CTA per live (sequence, head), heads fastest so that concurrently
            // running CTAs stream neighbouring 128-byte column slices of the same cross-KV rows
            const FusedQ fq{};
            const T* kb = reinterpret_cast<const T*>(L.cross_k);
            const T* vb = reinterpret_cast<const T*>(L.cross_v);
            for (int it = cta; it < n_act * H; it += G) {
                const int b = s_active[it / H], h = it % H;
                cross_attn_body<T>(qb, d, kb, vb, a.ld_kv, a.win_stride, att, d, a.n_ctx, fq, h, b, smem, sync);
                __syncthreads();
            }
        } else {
            const T* X; const void* W; int N, K;
            SkinnyEpilogue e{};
            switch (kind) {
                case 1: X = h16; W = L.qkv_w; N = 3 * d; K = d; e.bias = L.qkv_b; e.out16 = qkv; e.ldo16 = 3 * d; break;
                case 3: X = att; W = L.o_w; N = d; K = d; e.bias = L.o_b; e.residual = a.x; e.ldr = d; e.out32 = a.x; e.ldo32 = d; break;
                case 5: X = h16; W = L.cq_w; N = d; K = d; e.bias = L.cq_b; e.out16 = qb; e.ldo16 = d; break;
                case 7: X = att; W = L.co_w; N = d; K = d; e.bias = L.co_b; e.residual = a.x; e.ldr = d; e.out32 = a.x; e.ldo32 = d; break;
                case 9: X = h16; W = L.fc1_w; N = 4 * d; K = d; e.bias = L.fc1_b; e.act = 1; e.out16 = mlp; e.ldo16 = 4 * d; break;
                default: X = mlp; W = L.fc2_w; N = d; K = 4 * d; e.bias = L.fc2_b; e.residual = a.x; e.ldr = d; e.out32 = a.x; e.ldo32 = d; break;
            }
            const int mt = a.mt[kind], ks = a.ks[kind];
            const int tiles = (N + 16 * mt - 1) / (16 * mt);
            const int kb = K / 32;
            const int per = tiles * chunks;
            for (int vb = cta; vb < per * ks; vb += G) {
                const int sI = vb / per, r = vb - sI * per;
                const int kb0 = (int)((int64_t)kb * sI / ks), kb1 = (int)((int64_t)kb * (sI + 1) / ks);
                float* part = ks > 1 ? a.part + (int64_t)sI * a.part_stride : nullptr;
                // (a 32-row MT = 2 variant of the body halves the FC1 block count but spills inside this kernel and
                //  slowed every stage by 1.7x when inlined here -- measured, profiles/r1_mega_stage_trace.md)
                skinny_body<T, 1>(X, K, reinterpret_cast<const T*>(W), K, Bn, N, kb0, kb1, e, part, N, r % tiles, r / tiles, smem, sync);
                __syncthreads();
            }
        }
        if (!last) sync.arrive();
    }
    if (sync.trace_idx >= 0 && threadIdx.x == 0) g_mega_trace[2 * sync.trace_idx + 1] = gtimer();
}

template <typename T>
int dec_step_mega(const DecLayerDev* layers, const DecStepArgs& a, cudaStream_t st) {
    SB_CHECK_ARG(a.Bn >= 1 && a.Bn <= kMegaMaxB, "decoder megakernel: batch must be in [1, 256]");
    SB_CHECK_ARG(a.d % 32 == 0 && a.d <= 1536 && a.d == a.n_head * 64, "decoder megakernel: d % 32 == 0, d <= 1536, d_head 64");
    SB_CHECK_ARG(a.n_ctx <= 1504 && a.n_text_ctx <= 448, "decoder megakernel: n_audio_ctx <= 1504, n_text_ctx <= 448");
    void (*kern)(const DecLayerDev*, DecStepArgs) = a.d <= 768 ? k_dec_step_mega<T, 6> : a.d <= 1280 ? k_dec_step_mega<T, 10> : k_dec_step_mega<T, 12>;
    static bool attr_done[3] = {false, false, false};
    const int ki = a.d <= 768 ? 0 : a.d <= 1280 ? 1 : 2;
    if (!attr_done[ki]) {
        SB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMegaSmem + 128));
        attr_done[ki] = true;
    }
    for (int k : {1, 3, 5, 7, 9, 10}) SB_CHECK_ARG(a.mt[k] == 1 && a.ks[k] >= 1 && a.ks[k] <= kMegaMaxSplit, "decoder megakernel: bad stage plan");
    SB_CHECK_ARG(a.ks[1] == 1 && a.ks[5] == 1 && a.ks[9] == 1, "decoder megakernel: QKV, Q and FC1 cannot be split along K");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms()); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = kMegaSmem + 128; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // co-residency of all CTAs is what makes the grid barrier legal
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, layers, a));
    g_launches += 1;
    return SB_OK;
}
// Pick (16-row tiles per block, K splits) for a [64 x K] x [K x N] projection on n_cta resident CTAs.
// A block ingests (K / ks) * 2 B * (64 activation rows + 16 mt weight rows) through its ~81 GB/s share
// of the L2 crossbar (6300 B/clk chip-wide), which is what bounds these stages -- not the weights' HBM
// time; every K slice adds one partial row the consumer has to read back.
void mega_plan(int N, int K, bool allow_split, int n_cta, int* mt_out, int* ks_out) {
    double best = 1e30;
    for (int mt = 1; mt <= 1; ++mt)
        for (int ks = 1; ks <= (allow_split ? kMegaMaxSplit : 1); ++ks) {
            const int kb = K / 32;
            if (kb / ks < 4) continue;
            const int tiles = (N + 16 * mt - 1) / (16 * mt);
            const int rounds = (tiles * ks + n_cta - 1) / n_cta;
            const double bytes = (double)((kb + ks - 1) / ks) * 64.0 * (64 + 16 * mt);
            const double us = rounds * (bytes / 81e3 + 0.8) + (ks > 1 ? 0.15 * ks : 0.0);
            if (us < best) { best = us; *mt_out = mt; *ks_out = ks; }
        }
}

int mega_trace_read(unsigned long long* out, int n) {
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    SB_CUDA_CHECK(cudaMemcpyFromSymbol(out, g_mega_trace, sizeof(unsigned long long) * (n < 1024 ? n : 1024)));
    return SB_OK;
}
template int dec_step_mega<__half>(const DecLayerDev*, const DecStepArgs&, cudaStream_t);
template int dec_step_mega<__nv_bfloat16>(const DecLayerDev*, const DecStepArgs&, cudaStream_t);

}  // namespace sb

// debug: per-stage timeline (ns, %globaltimer) of CTA 0 for the last megakernel launch made with SB_MEGA_TRACE=1
extern "C" __attribute__((visibility("default"))) int sb_debug_mega_trace(unsigned long long* out, int n) {
    SB_CHECK_ARG(out && n > 0, "null pointer");
    return sb::mega_trace_read(out, n);
}
