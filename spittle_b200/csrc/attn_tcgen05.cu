// k_attn_enc_ts: encoder self-attention (non-causal, d_head 64, T = 1500) on tcgen05 / TMEM.
//
// Replaces the KQ / softmax / KQV part of the whisper.cpp encoder graph.
// One CTA = one (window, head, 128-query tile).  S = Q K^T and O = P V are tcgen05.mma with the
// accumulators in TMEM; the softmax runs on 128 threads, one per query row (a TMEM lane), so row
// maxima and sums need no shuffles.
//
// Two passes over the 12 key tiles of 128 keys:
//   pass 1  S_j = Q K_j^T (UMMA 128x128x16 x4)  ->  running row maximum only
//   pass 2  S_j again, p = exp2((s - m) * scale) with the FINAL maximum, P_j (16-bit) written to shared
//           memory in the 128B-swizzled K-major layout, O += P_j V_j (UMMA 128x64x16 x8, V is the
//           MN-major B operand straight from its [key][64] rows)
// Knowing the final maximum up front removes the accumulator rescaling of the one-pass online softmax
// (TMEM round trips + an extra dependency between softmax and MMA); it costs the QK^T MMAs twice, which
// is free here: the kernel is bound by the exp2 (MUFU) rate, not by the tensor pipe.
// Rounding points of the reference are kept: Q, K, V are 16-bit, S is f32, probabilities are rounded to
// 16 bits before P V, accumulation in f32; the normaliser is the f32 sum of the un-rounded probabilities.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 softmax + epilogue
// (two warps per TMEM lane quarter, each owning 64 of the 128 key columns of a tile, so every SM
// sub-partition has two warps to overlap the MUFU exp2 latency with the FMA/convert work).
// TMEM: S double buffer 2 x 128 columns, O 64 columns (512 allocated).
#include "common.cuh"
#include "decoder.cuh"
#include <cuda.h>
#include <type_traits>
#include <cstdlib>

namespace sb {
extern std::atomic<uint64_t> g_launches;
int make_tmap_2d(CUtensorMap* map, const void* ptr, int is_f16, int64_t rows, int64_t cols, int64_t ld, int box_rows);

constexpr int kAtBM = 128;            // queries per CTA
constexpr int kAtD = 64;
constexpr int kAtThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two per TMEM lane quarter)

__device__ __forceinline__ uint32_t at_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void at_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void at_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void at_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "AT_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra AT_WAIT_DONE;\n"
        "bra AT_WAIT_LOOP;\n"
        "AT_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void at_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void at_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void at_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void at_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ float at_ex2(float x) {     // single MUFU.EX2 (exp2f adds denormal range fix-ups)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// back-off wait for the single-thread roles: a tight try_wait loop steals issue slots from the softmax warps
__device__ __forceinline__ void at_mbar_wait_relaxed(uint32_t bar, uint32_t parity, unsigned ns = 64) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (ns) __nanosleep(ns);
    }
}
__device__ __forceinline__ void at_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void at_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void at_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B operand (rows of 128 B, 8-row groups 1024 B apart); also valid for the MN-major
// V tile whose 64-element rows are exactly one swizzle atom wide (K direction = consecutive 128 B rows).
__device__ __forceinline__ uint64_t at_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                    // LBO (unused: the MN / K extent per instruction fits one atom)
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // Blackwell descriptor version
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// ==========================================================================================
// k_attn_enc_ts: Q and P are tcgen05.mma A operands read from TMEM.
//
// The first tcgen05 version (round 1, `k_attn_enc_tc`, removed) moved Q, K, V and P through shared memory;
// ncu on it (profiles/r1_full_attn_enc_tc_bn64.md): XU 43 %, tensor 33 %, issue 49 % -- nothing
// saturated, yet no pipeline change moved the time.  What is saturated is shared-memory bandwidth: an
// M128 x N64 x K16 UMMA reads 4 KB of A and 2 KB of B for 32 tensor-clocks of math (192 B/clk against the
// SM's 128 B/clk), and per 64-key tile the CTA moved Q (16 KB, re-read for every tile), P (16 KB written by
// the softmax warps, 16 KB read back), K and V (8 KB each, written by TMA and read by the MMA): ~88 KB.
// Here Q is staged once into 32 TMEM columns and P is stored with tcgen05.st over the S columns it was
// computed from, so per tile only K and V cross shared memory (32 KB incl. the TMA writes).
//
// TMEM (256 columns, 2 CTAs/SM): S0 [0,64) S1 [64,128) O [128,192) Q [192,224).  P_j lives inside S buffer
// j%2: the warp that owns key columns [32h, 32h+32) of a row writes its 32 probabilities as 16 packed words
// to columns [32h, 32h+16) -- over S values it has already loaded itself, so no cross-warp hazard.
// The MMA thread issues S_0 S_1 | PV_0 S_2 | PV_1 S_3 | ...: the tensor pipe executes in issue order, so
// S_{j+2} cannot overwrite buffer j%2 before PV_j has read P_j, and the softmax warps always find S_{j+1}
// complete when they finish tile j.  Pass 1 (row maxima) rings over the three 64-column buffers below Q.
// ==========================================================================================
static std::atomic<unsigned long long*> g_attn_fallback_ctr[64];       // one device counter per GPU for all instantiations
static int attn_fallback_counter(unsigned long long** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<unsigned long long*>& slot = g_attn_fallback_ctr[dev & 63];
    unsigned long long* p = slot.load(std::memory_order_acquire);
    if (!p) {
        SB_CUDA_CHECK(cudaMalloc(&p, 8));
        SB_CUDA_CHECK(cudaMemset(p, 0, 8));
        unsigned long long* expected = nullptr;
        if (!slot.compare_exchange_strong(expected, p)) { cudaFree(p); p = expected; }
    }
    *out = p;
    return SB_OK;
}
__device__ unsigned long long g_attn_trace[4096 * 6];     // debug (SB_ATTN_TRACE=1): per-CTA phase timestamps
constexpr int kTsKvStages = 6;
constexpr int kTsTileBytes = 64 * 64 * 2;
constexpr int kTsSmem = 1024 + kTsKvStages * kTsTileBytes + 256 + 2 * 128 * 4;

__device__ __forceinline__ void at_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void at_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void at_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Fast path (attempt 0): ONE sweep.  The softmax needs a reference m with  max - 14 <= m  (no f16 overflow
// of 2^(s - m)) and m <= max (the largest probabilities keep full 16-bit precision): it does not need the exact
// row maximum.  m is taken from the first key tile alone, every later score is checked against m + 14, and the
// result O / l is mathematically independent of m.  Only if some row's scores climb more than 2^14 above its
// first-tile maximum (flagged per CTA) is the q-tile recomputed by the exact two-pass algorithm (attempt 1) --
// measured per-CTA time of the two-pass kernel: pass 1 9.6 us of 25 us (profiles/r1_attn_phase_trace.md).
//
// Measured and rejected (round 2, tools/kbench.py attn): `ex2.approx.f16x2` on packed exponent arguments with the row sums taken
// from the tensor pipe (P x ones).  ptxas lowers the packed ex2 to TWO MUFU.EX2.F16, so the MUFU count does not drop: 530 vs
// 641 TFLOP/s (Large-v3 shape) and twice the error (5.3e-4 vs 2.9e-4 rel-RMS); mixing an f16 P with a bf16 V in one kind::f16
// MMA is an illegal instruction.  The sweep stays one MUFU.EX2 per score: 128 x 64 per tile against 256 tensor clocks.
// Also measured and rejected (end of round 2): every fourth exp2 on the FMA / ALU pipes instead (round-to-nearest split, degree-4
// polynomial, exponent patched in: 10 instructions, 4e-5 relative, parity unchanged) -- 631 -> 559 TFLOP/s on the Turbo shape, 579 -> 517
// on Small: the sweep's issue slots are as loaded as its MUFU, so moving work from one to the other loses.
template <typename T>
__global__ void __launch_bounds__(kAtThreads, 2)
k_attn_enc_ts(const __grid_constant__ CUtensorMap tm_kv, const T* __restrict__ qkv, T* __restrict__ out, int n_ctx, int d_model,
              float scale_log2e, int trace, int force_two_pass, unsigned long long* __restrict__ n_fallback) {
    constexpr int BN = 64;
    const int lin_cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    auto stamp = [&](int k) {
        if (trace && lin_cta < 4096 && threadIdx.x == 64) {
            unsigned long long t; unsigned smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_attn_trace[lin_cta * 6 + k] = t;
            if (k == 0) g_attn_trace[lin_cta * 6 + 5] = smid;
        }
    };
    stamp(0);
    extern __shared__ unsigned char at_smem_raw[];
    const uint32_t raw = at_smem_u32(at_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = at_smem_raw + (base - raw);
    const uint32_t sKV = base;
    constexpr int kCtl = kTsKvStages * kTsTileBytes;      // barriers / flags / exchange follow
    const uint32_t bar0 = sKV + kCtl;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + kCtl + 192);
    int* redo_flag = reinterpret_cast<int*>(base_ptr + kCtl + 196);
    float* xch = reinterpret_cast<float*>(base_ptr + kCtl + 256);
    auto kv_full = [&](int s) { return bar0 + 8u * s; };
    auto kv_empty = [&](int s) { return bar0 + 8u * (kTsKvStages + s); };
    auto s_full = [&](int b) { return bar0 + 8u * (2 * kTsKvStages + b); };           // exp sweep, 2 buffers
    auto p_full = [&](int b) { return bar0 + 8u * (2 * kTsKvStages + 2 + b); };
    auto s1_full = [&](int b) { return bar0 + 8u * (2 * kTsKvStages + 4 + b); };      // maximum sweep, 3 buffers
    auto s1_empty = [&](int b) { return bar0 + 8u * (2 * kTsKvStages + 7 + b); };
    const uint32_t q_full = bar0 + 8u * (2 * kTsKvStages + 10);
    const uint32_t o_full = bar0 + 8u * (2 * kTsKvStages + 11);
    constexpr int kS1 = 3;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, win = blockIdx.z;
    const int n_kt = (n_ctx + BN - 1) / BN;
    const int col_k = d_model + head * kAtD, col_v = 2 * d_model + head * kAtD;

    auto init_barriers = [&]() {
        for (int s = 0; s < kTsKvStages; ++s) { at_mbar_init(kv_full(s), 1); at_mbar_init(kv_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { at_mbar_init(s_full(b), 1); at_mbar_init(p_full(b), 8); }
        for (int b = 0; b < kS1; ++b) { at_mbar_init(s1_full(b), 1); at_mbar_init(s1_empty(b), 8); }
        at_mbar_init(q_full, 8);
        at_mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    };
    if (threadIdx.x == 0) {
        init_barriers();
        *redo_flag = 0;
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_kv) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    at_fence_before();
    __syncthreads();
    at_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tS0 = tmem, tO = tmem + 128, tQ = tmem + 192;

    // softmax-warp coordinates
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int t_q = qt * kAtBM + row;

    for (int attempt = force_two_pass ? 1 : 0; attempt < 2; ++attempt) {
        const bool two_pass = attempt == 1;
        if (two_pass && !force_two_pass) {
            // the fast sweep is complete (every TMA load consumed, every MMA committed and observed): start over
            if (*redo_flag == 0) break;
            __syncthreads();
            if (threadIdx.x == 0) { init_barriers(); if (n_fallback) atomicAdd(n_fallback, 1ULL); }
            at_fence_before();
            __syncthreads();
            at_fence_after();
        }
        const bool load_q = attempt == 0 || force_two_pass;

        if (warp == 0) {
            if (lane == 0) {
                int stage = 0; uint32_t phase = 0;
                for (int pass = two_pass ? 0 : 1; pass < 2; ++pass)
                    for (int j = 0; j < n_kt; ++j)
                        for (int which = 0; which <= pass; ++which) {
                            at_mbar_wait_relaxed(kv_empty(stage), phase ^ 1);
                            at_mbar_expect_tx(kv_full(stage), kTsTileBytes);
                            at_tma_2d(sKV + stage * kTsTileBytes, &tm_kv, which == 0 ? col_k : col_v, win * n_ctx + j * BN, kv_full(stage));
                            if (++stage == kTsKvStages) { stage = 0; phase ^= 1; }
                        }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t fmt = (uint32_t)Op16<T>::kUmmaFormat;
                const uint32_t idesc_s = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kAtBM >> 4) << 24);
                const uint32_t idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)(kAtD >> 3) << 17) | ((uint32_t)(kAtBM >> 4) << 24);
                int stage = 0; uint32_t phase = 0;
                if (load_q) { at_mbar_wait(q_full, 0); at_fence_after(); }
                if (two_pass) {
                    // maximum sweep: S_j into ring buffer j % 3
                    for (int j = 0; j < n_kt; ++j) {
                        const int b1 = j % kS1;
                        at_mbar_wait_relaxed(s1_empty(b1), ((uint32_t)(j / kS1) & 1u) ^ 1u, 0);
                        at_mbar_wait_relaxed(kv_full(stage), phase, 0);
                        at_fence_after();
                        const uint64_t kdesc = at_desc(sKV + stage * kTsTileBytes);
#pragma unroll
                        for (int k = 0; k < kAtD / 16; ++k) at_mma_ts(tS0 + b1 * BN, tQ + 8 * k, kdesc + 2 * k, idesc_s, k != 0);
                        at_commit(kv_empty(stage));
                        at_commit(s1_full(b1));
                        if (++stage == kTsKvStages) { stage = 0; phase ^= 1; }
                    }
                    for (int b1 = 0; b1 < kS1; ++b1) {
                        const int uses = b1 < n_kt ? (n_kt - b1 + kS1 - 1) / kS1 : 0;
                        if (uses > 0) at_mbar_wait_relaxed(s1_empty(b1), (uint32_t)(uses - 1) & 1u, 0);
                    }
                    at_fence_after();
                }
                // exp sweep.  Stage order produced by the TMA warp: K_0 V_0 K_1 V_1 ...; S_{j+2} needs K_{j+2}, two tiles
                // ahead of V_j, so the K stages are consumed out of ring order: (stage, phase) of tile t's K are explicit.
                const int st0 = stage; const uint32_t ph0 = phase;          // ring position of K_0
                auto ring_at = [&](int idx, int& st, uint32_t& ph) {         // idx-th tile (K_0 = 0, V_0 = 1, K_1 = 2, ...)
                    const int lin = st0 + idx;
                    st = lin % kTsKvStages;
                    ph = ph0 ^ ((uint32_t)(lin / kTsKvStages) & 1u);
                };
                auto mma_s2 = [&](int t) {
                    int st; uint32_t ph;
                    ring_at(2 * t, st, ph);
                    at_mbar_wait_relaxed(kv_full(st), ph, 0);
                    at_fence_after();
                    const uint64_t kdesc = at_desc(sKV + st * kTsTileBytes);
#pragma unroll
                    for (int k = 0; k < kAtD / 16; ++k) at_mma_ts(tS0 + (t & 1) * BN, tQ + 8 * k, kdesc + 2 * k, idesc_s, k != 0);
                    at_commit(kv_empty(st));
                    at_commit(s_full(t & 1));
                };
                mma_s2(0);
                if (n_kt > 1) mma_s2(1);
                for (int j = 0; j < n_kt; ++j) {
                    const int pb = j & 1; const uint32_t pphase = (uint32_t)(j >> 1) & 1u;
                    int st; uint32_t ph;
                    ring_at(2 * j + 1, st, ph);
                    at_mbar_wait_relaxed(p_full(pb), pphase, 0);
                    at_mbar_wait_relaxed(kv_full(st), ph, 0);
                    at_fence_after();
                    const uint64_t vdesc = at_desc(sKV + st * kTsTileBytes);
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k) {
                        // keys 16k..16k+15: packed P columns [8k, 8k+8) of the owning warp's 16-column block (block h at column 32h)
                        const uint32_t pa = tS0 + pb * BN + (k >> 1) * 32 + (k & 1) * 8;
                        at_mma_ts(tO, pa, vdesc + (uint64_t)(k * 2048 >> 4), idesc_o, (j | k) != 0);
                    }
                    at_commit(kv_empty(st));
                    if (j + 2 < n_kt) mma_s2(j + 2);
                }
                at_commit(o_full);
            }
        } else {
            // ---- Q row -> TMEM: this warp stores elements [32 half, 32 half + 32) of its rows as 16 packed words ----
            if (load_q) {
                uint32_t w[16];
                if (t_q < n_ctx) {
                    const uint4* src = reinterpret_cast<const uint4*>(qkv + ((int64_t)win * n_ctx + t_q) * 3 * d_model + head * kAtD + half * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 u = __ldg(src + i);
                        w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = 0u;
                }
                at_st16(tQ + lane_addr + half * 16, w);
                at_wait_st();
                at_fence_before();
                __syncwarp();
                if (lane == 0) at_mbar_arrive(q_full);
            }
            float m = -INFINITY;
            stamp(1);
            if (two_pass) {
                // ---- maximum sweep: row maxima over this warp's 32 columns of every tile ----
                for (int j = 0; j < n_kt; ++j) {
                    const int b1 = j % kS1;
                    at_mbar_wait(s1_full(b1), (uint32_t)(j / kS1) & 1u);
                    at_fence_after();
                    uint32_t v0[32];
                    at_ld32(tS0 + lane_addr + b1 * BN + half * 32, v0);
                    at_wait_ld();
                    at_fence_before();
                    __syncwarp();
                    if (lane == 0) at_mbar_arrive(s1_empty(b1));
                    const int k0 = j * BN + half * 32;
                    if (k0 + 32 <= n_ctx) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(v0[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (k0 + i < n_ctx) m = fmaxf(m, __uint_as_float(v0[i]));
                    }
                }
            } else {
                // ---- reference = maximum of the first key tile (peek: the sweep below loads S_0 again) ----
                at_mbar_wait(s_full(0), 0);
                at_fence_after();
                uint32_t v0[32];
                at_ld32(tS0 + lane_addr + half * 32, v0);
                at_wait_ld();
                const int k0 = half * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (k0 + i < n_ctx) m = fmaxf(m, __uint_as_float(v0[i]));
            }
            xch[half * 128 + row] = m;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            m = fmaxf(xch[row], xch[128 + row]);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- exp sweep ----
            stamp(2);
            const float ms = m * scale_log2e;
            float l = 0.f;
            float smax = -INFINITY;      // fast path: largest score seen (overflow guard)
            for (int j = 0; j < n_kt; ++j) {
                const int sb = j & 1;
                at_mbar_wait(s_full(sb), (uint32_t)(j >> 1) & 1u);
                at_fence_after();
                uint32_t v[32];
                at_ld32(tS0 + lane_addr + sb * BN + half * 32, v);
                at_wait_ld();
                const int k0 = j * BN + half * 32;
                uint32_t pk[16];
                if (k0 + 32 <= n_ctx) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float p0 = at_ex2(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -ms));
                        const float p1 = at_ex2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -ms));
                        l += p0 + p1;
                        pk[i] = Op16<T>::pack2(p0, p1);
                    }
                    if (!two_pass) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) smax = fmaxf(smax, __uint_as_float(v[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = at_ex2(fmaf(__uint_as_float(v[2 * i]), scale_log2e, -ms));
                        float p1 = at_ex2(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2e, -ms));
                        if (k0 + 2 * i >= n_ctx) p0 = 0.f; else smax = fmaxf(smax, __uint_as_float(v[2 * i]));
                        if (k0 + 2 * i + 1 >= n_ctx) p1 = 0.f; else smax = fmaxf(smax, __uint_as_float(v[2 * i + 1]));
                        l += p0 + p1;
                        pk[i] = Op16<T>::pack2(p0, p1);
                    }
                }
                at_st16(tS0 + lane_addr + sb * BN + half * 32, pk);
                at_wait_st();
                at_fence_before();
                __syncwarp();
                if (lane == 0) at_mbar_arrive(p_full(sb));
            }
            stamp(3);
            // a score more than 2^14 above the reference would overflow the 16-bit probability: redo exactly
            if (!two_pass && !(fmaf(smax, scale_log2e, -ms) <= 14.0f)) *redo_flag = 1;
            xch[half * 128 + row] = l;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            l = xch[row] + xch[128 + row];
            const bool redo = !two_pass && *redo_flag != 0;
            // ---- epilogue: O / l, each warp stores 32 of the 64 head dims ----
            at_mbar_wait(o_full, 0);
            at_fence_after();
            if (!redo) {
                const float inv = 1.0f / l;
                T* orow = out + ((int64_t)win * n_ctx + t_q) * d_model + head * kAtD + half * 32;
                uint32_t v[32];
                at_ld32(tO + lane_addr + half * 32, v);
                at_wait_ld();
                if (t_q < n_ctx) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 u;
                        u.x = Op16<T>::pack2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
                        u.y = Op16<T>::pack2(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv);
                        u.z = Op16<T>::pack2(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv);
                        u.w = Op16<T>::pack2(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + i) = u;
                    }
                }
            }
            at_fence_before();
            stamp(4);
        }
        __syncthreads();
    }
    if (warp == 2) {
        at_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
    }
}

template <typename T>
static int attn_enc_ts_launch(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st) {
    CUtensorMap tkv;
    const int is_f16 = std::is_same<T, __half>::value ? 1 : 0;
    int rc = make_tmap_2d(&tkv, qkv, is_f16, (int64_t)n_windows * n_ctx, 3 * (int64_t)d_model, 3 * (int64_t)d_model, 64);
    if (rc) return rc;
    SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_attn_enc_ts<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmem)); });
    dim3 grid(ceil_div(n_ctx, kAtBM), n_head, n_windows);
    const float scale_log2e = (1.0f / 8.0f) * 1.4426950408889634f;
    static int trace = [] { const char* e = getenv("SB_ATTN_TRACE"); return e ? atoi(e) : 0; }();
    static int two_pass = [] { const char* e = getenv("SB_ATTN_TWO_PASS"); return e ? atoi(e) : 0; }();     // 1: always the exact two-pass sweep
    unsigned long long* n_fallback = nullptr;             // q-tiles the fast sweep had to redo (sb_debug_attn_fallbacks)
    if ((rc = attn_fallback_counter(&n_fallback))) return rc;
    k_attn_enc_ts<T><<<grid, kAtThreads, kTsSmem, st>>>(tkv, qkv, out, n_ctx, d_model, scale_log2e, trace, two_pass, n_fallback);
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

template <typename T>
int attn_enc_tc(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st) {
    SB_CHECK_ARG(d_model == n_head * kAtD, "attention: d_head must be 64");
    return attn_enc_ts_launch<T>(qkv, out, n_windows, n_ctx, d_model, n_head, st);
}
template int attn_enc_tc<__half>(const __half*, __half*, int, int, int, int, cudaStream_t);
template int attn_enc_tc<__nv_bfloat16>(const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, cudaStream_t);

}  // namespace sb

extern "C" __attribute__((visibility("default"))) int sb_debug_attn_trace(unsigned long long* out, int n) {
    SB_CHECK_ARG(out && n > 0, "null pointer");
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    SB_CUDA_CHECK(cudaMemcpyFromSymbol(out, sb::g_attn_trace, sizeof(unsigned long long) * (n < 4096 * 6 ? n : 4096 * 6)));
    return SB_OK;
}

// number of 128-query tiles the single-sweep attention had to recompute with the exact two-pass algorithm
extern "C" __attribute__((visibility("default"))) int sb_debug_attn_fallbacks(unsigned long long* out) {
    SB_CHECK_ARG(out, "null pointer");
    *out = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    unsigned long long* p = sb::g_attn_fallback_ctr[dev & 63].load();
    if (!p) return SB_OK;
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    SB_CUDA_CHECK(cudaMemcpy(out, p, 8, cudaMemcpyDeviceToHost));
    return SB_OK;
}
