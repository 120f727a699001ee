// gemm_tcgen05: C[M,N] = epilogue(A[M,K] * W[N,K]^T) on the 5th-gen tensor cores (sm_100a).
//
// Replaces the ggml CPU mul_mat calls of the whisper.cpp encoder graph (SURVEY.md App. C.2,
// rows a5/a6 of 8(a)): conv stem (implicit GEMM), QKV / out / MLP projections, cross-KV.
// Reference rounding points are kept: 16-bit operands (bf16, or f16 exactly like ggml),
// fp32 accumulation.
//
// Structure (one CTA per SM, persistent over output tiles, 320 threads):
//   warp 0      TMA producer : cp.async.bulk.tensor 2D loads of A (128x64) and W (256x64)
//                              tiles, 128B-swizzled, into a 4-stage shared-memory ring
//   warp 1      MMA issuer   : one thread issues tcgen05.mma.cta_group::1.kind::f16
//                              (UMMA 128x256x16), accumulators in TMEM, double-buffered
//                              (2 x 256 columns) so the epilogue of tile i overlaps tile i+1
//   warps 2..9  epilogue     : tcgen05.ld 32x32b.x32 -> registers -> XOR-swizzled smem transpose ->
//                              (+bias, GELU, +residual / positional embedding) -> fully coalesced
//                              128-bit global stores (each warp instruction writes whole lines)
// Synchronisation is mbarrier-only: full/empty per smem stage, tmem_full/tmem_empty per
// accumulator stage; tcgen05.commit releases smem stages and publishes accumulators.
// Tile order is n-fastest so the CTAs running together share each A tile through L2 and A
// streams from HBM once; W (a few MB) stays L2-resident.
//
// kCtas = 2 (large M): the two CTAs of a cluster (one TPC) work on one 256x256 tile with
// tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 columns),
// the leader CTA's MMA thread issues UMMA 256x256x16 for the pair, and each CTA's TMEM holds the accumulator rows
// of its own A half.  Per k-block a CTA moves 32 KB through shared memory instead of 48 KB (the 1-CTA tile was
// bound by shared-memory bandwidth: profiles/r1_full_gemm_tn.md), and the ring is 6 stages deep instead of 4.
// Both producers signal the LEADER's full barrier (the peer through its cluster address), tcgen05.commit multicasts
// the "stage free" / "accumulator ready" arrivals to both CTAs, and the peer's epilogue warps release the
// accumulator stage on the leader's barrier.
#include "common.cuh"
#include "decoder.cuh"
#include <cuda.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace sb {
extern std::atomic<uint64_t> g_launches;

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;          // 64 x 2 B = 128 B = swizzle span
constexpr int kUmmaK = 16;
constexpr int kStagingBytes = 32 * 32 * 4;          // per epilogue warp: one 32x32 f32 chunk, XOR-swizzled
constexpr int kStageBytesA = kBM * kBK * 2;
// kEW epilogue warps (warps 2 .. 2 + kEW): kEW / 4 per TMEM lane quarter, each owning 256 / (kEW / 4) accumulator columns.
// ncu on the 8-warp epilogue (profiles/r2_full_gemm_tn.md): FC1 + GELU tensor pipe 53 % active with XU 21 %, FMA 15 %, issue 30 %
// -- nothing saturated: two warps per scheduler cannot hide the tcgen05.ld / shared-memory / MUFU latencies of the chunk
// loop, and for K <= 1280 the epilogue of a tile is as long as its mainloop.  16 warps (four per scheduler, <= 113 registers)
// on the CTA-pair tiles (5-stage ring instead of 6 to make room for the staging buffers); the residual rows of a chunk are
// requested while its TMEM load is in flight.
template <int kCtas, int kEW> struct GemmCfg {
    static_assert(kEW == 8 || (kEW == 16 && kCtas == 2), "16 epilogue warps only with the CTA-pair tile (shared-memory budget)");
    static constexpr int kStages = kCtas == 1 ? 4 : (kEW == 16 ? 5 : 6);
    static constexpr int kThreads = 64 + 32 * kEW;                  // warp 0 TMA, warp 1 MMA, the rest epilogue
    static constexpr int kRowsB = kBN / kCtas;                      // W rows (output columns) staged by one CTA
    static constexpr int kStageBytesB = kRowsB * kBK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEW * kStagingBytes;
    static_assert(kSmem <= 227 * 1024, "shared memory");
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// 2-CTA form: the data lands in this CTA's shared memory, the transaction bytes are counted on `bar`, a
// shared::cluster address that may belong to the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on `bar` in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                         // LBO (unused for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;               // SBO = 1024 B   [32,46)
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}

template <typename T, int kCtas, int kEW, bool kRes>
__global__ void __launch_bounds__(GemmCfg<kCtas, kEW>::kThreads, 1)
k_gemm_tn(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
          GemmEpilogue ep, int M, int N, int K) {
    using Cfg = GemmCfg<kCtas, kEW>;
    constexpr int kEpiWarps = kEW;
    constexpr int kStages = Cfg::kStages;
    constexpr int kStageBytes = Cfg::kStageBytes;
    constexpr int kTileM = kBM * kCtas;                    // rows of one output tile (of the CTA pair)
    extern __shared__ unsigned char smem_raw_g[];
    const uint32_t raw = smem_u32(smem_raw_g);
    const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024 B alignment
    unsigned char* base_ptr = smem_raw_g + (base - raw);
    const uint32_t bar_base = base + kStages * kStageBytes;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + kStages * kStageBytes + 192);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = kCtas == 1 ? 0u : cluster_ctarank();
    const int tile0 = blockIdx.x / kCtas, tile_step = gridDim.x / kCtas;    // tiles are dealt to clusters
    const int num_m = (M + kTileM - 1) / kTileM, num_n = (N + kBN - 1) / kBN;
    const int num_tiles = num_m * num_n;
    const int num_kb = (K + kBK - 1) / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps * kCtas); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    }
    if (warp == 2) {
        if constexpr (kCtas == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * kBN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * kBN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();    // the peer's barriers are initialised too
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int n_blk = tile % num_n, m_blk = tile / num_n;
                const int row_a = m_blk * kTileM + (int)cta_rank * kBM;
                const int row_b = n_blk * kBN + (int)cta_rank * Cfg::kRowsB;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = base + stage * kStageBytes;
                    if constexpr (kCtas == 1) {
                        mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
                        tma_load_2d(sa, &tma_a, kb * kBK, row_a, full_bar(stage));
                        tma_load_2d(sa + kStageBytesA, &tma_b, kb * kBK, row_b, full_bar(stage));
                    } else {
                        // the leader's barrier counts the bytes of both CTAs; a peer load that completes before the
                        // leader's expect_tx only drives the transaction count negative for a moment
                        if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * kStageBytes);
                        const uint32_t fb = mapa_u32(full_bar(stage), 0);
                        tma_load_2d_pair(sa, &tma_a, kb * kBK, row_a, fb);
                        tma_load_2d_pair(sa + kStageBytesA, &tma_b, kb * kBK, row_b, fb);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // instruction descriptor: D=f32, A=B=T, K-major both, N=256, M=128 (256 for the CTA pair)
            const uint32_t idesc = (1u << 4) | ((uint32_t)Op16<T>::kUmmaFormat << 7) |
                                   ((uint32_t)Op16<T>::kUmmaFormat << 10) | ((uint32_t)(kBN >> 3) << 17) |
                                   ((uint32_t)(kTileM >> 4) << 24);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                mbar_wait(tempty_bar(as), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * kBN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + stage * kStageBytes;
                    const uint64_t adesc = make_smem_desc(sa);
                    const uint64_t bdesc = make_smem_desc(sa + kStageBytesA);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // advance 16 elements = 32 B along K inside the swizzle atom: +2 in 16 B units
                        if constexpr (kCtas == 1) tc_mma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                        else tc_mma_f16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
                    }
                    if constexpr (kCtas == 1) tc_commit(empty_bar(stage)); else tc_commit_pair(empty_bar(stage));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if constexpr (kCtas == 1) tc_commit(tfull_bar(as)); else tc_commit_pair(tfull_bar(as));
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        const int ew = warp - 2;              // 0 .. kEW - 1
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        constexpr int kChunks = 8 / (kEW / 4);     // 32-column chunks per warp: 4 (8 warps) or 2 (16 warps)
        const int part = ew >> 2;             // which kChunks x 32 columns of the 256-wide accumulator
        float4* stg = reinterpret_cast<float4*>(base_ptr + kStages * kStageBytes + 256 + ew * kStagingBytes);
        const int rsub = lane >> 3;           // read-back: 4 rows per instruction, 8 lanes x float4 per row
        const int c4 = lane & 7;
        int as = 0; uint32_t aphase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            const int n_blk = tile % num_n, m_blk = tile / num_n;
            const int row_base = m_blk * kTileM + (int)cta_rank * kBM + q * 32;
            const int col_l = c4 * 4;
            const bool has_res = kRes && ep.residual != nullptr;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < kChunks; ++cc) {
                const int c = part * kChunks + cc;
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kBN + c * 32);
                tc_ld32(taddr, v);
                // the residual rows of this chunk are requested while the TMEM load is in flight (four warps per scheduler
                // cover the rest of the HBM latency)
                float4 rcur[kRes ? 8 : 1];
                if constexpr (kRes) {
                    const int colr = n_blk * kBN + c * 32 + col_l;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = row_base + it * 4 + rsub;
                        rcur[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_res && row < M && colr < N) {
                            const int rrow = ep.res_row_mod > 0 ? row % ep.res_row_mod : row;
                            rcur[it] = *reinterpret_cast<const float4*>(ep.residual + (int64_t)rrow * ep.ldr + colr);
                        }
                    }
                }
                tc_wait_ld();
                const int col0 = n_blk * kBN + c * 32;
                if (col0 >= N || row_base >= M) continue;       // warp-uniform
                // (measured and rejected, round 2: storing straight from the registers -- lane = row, 16-byte stores, no
                // shared-memory round trip -- is 15-45 % SLOWER on every K <= 1280 shape: profiles/r2_gemm_epilogue_ab.md)
                // transpose through smem: lane = row writes its 32 values as 8 float4, chunk index XOR (row & 7)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    stg[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                __syncwarp();
                const int col = col0 + col_l;
                const bool col_ok = col < N;                     // N % 8 == 0: a 4-column group is all in or all out
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ep.bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
                float4 f[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + rsub;
                    f[it] = stg[rr * 8 + (c4 ^ (rr & 7))];
                    f[it].x += bias4.x; f[it].y += bias4.y; f[it].z += bias4.z; f[it].w += bias4.w;
                }
                if (ep.act == 1) {      // 32 independent GELUs: straight-line so the MUFU latencies overlap
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        f[it].x = gelu_tanh(f[it].x); f[it].y = gelu_tanh(f[it].y);
                        f[it].z = gelu_tanh(f[it].z); f[it].w = gelu_tanh(f[it].w);
                    }
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = row_base + it * 4 + rsub;
                    if (row < M && col_ok) {
                        float4 o = f[it];
                        if constexpr (kRes) {
                            if (has_res) { o.x += rcur[it].x; o.y += rcur[it].y; o.z += rcur[it].z; o.w += rcur[it].w; }
                        }
                        if (ep.out_f32) {
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (int64_t)row * ep.ldo + col) = o;
                        } else {
                            uint2 u;
                            u.x = Op16<T>::pack2(o.x, o.y); u.y = Op16<T>::pack2(o.z, o.w);
                            *reinterpret_cast<uint2*>(reinterpret_cast<T*>(ep.out) + (int64_t)row * ep.ldo + col) = u;
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {      // the accumulator stage is released on the barrier the MMA thread waits on: the leader's
                if (kCtas == 1 || cta_rank == 0) mbar_arrive(tempty_bar(as));
                else mbar_arrive_cluster(mapa_u32(tempty_bar(as), 0));
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();    // neither CTA of a pair leaves while the other works
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kCtas == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kBN));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kBN));
    }
}

// ---- host side ------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

// 2D row-major [rows, cols] 16-bit tensor, row stride ld elements; box = [box_rows, 64 cols], 128B swizzle
int make_tmap_2d(CUtensorMap* map, const void* ptr, int is_f16, int64_t rows, int64_t cols, int64_t ld,
                 int box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return SB_ERR_CUDA; }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) {
        set_error("gemm operand must be 16-byte aligned with a 16-byte multiple row stride");
        return SB_ERR_INVALID;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: " + std::to_string((int)r)); return SB_ERR_CUDA; }
    return SB_OK;
}

// SM count of the CURRENT device (one process may drive several GPUs)
int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

// A [M,K] (row stride lda), W [N,K] (row stride ldw); dtype 0 bf16 / 1 f16.
int gemm_tn(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
            const GemmEpilogue& ep, cudaStream_t st) {
    SB_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: empty problem");
    SB_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm: K and N must be multiples of 8");
    SB_CHECK_ARG(ep.out_f32 ? (ep.ldo % 4 == 0) : (ep.ldo % 8 == 0), "gemm: output row stride alignment");
    // CTA pairs (cta_group::2, 256x256 tiles) when there are enough rows to fill the machine with them
    constexpr int pair_min_m = 4096;
    const bool pair = M >= pair_min_m;
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, A, dtype, M, K, lda, kBM);
    if (rc != SB_OK) return rc;
    rc = make_tmap_2d(&tb, W, dtype, N, K, ldw, pair ? GemmCfg<2, 8>::kRowsB : GemmCfg<1, 8>::kRowsB);
    if (rc != SB_OK) return rc;
    using C1 = GemmCfg<1, 8>; using C2 = GemmCfg<2, 16>;
    SB_ONCE_PER_DEVICE({
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__nv_bfloat16, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1::kSmem));
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__half, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1::kSmem));
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__nv_bfloat16, 2, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmem));
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__half, 2, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmem));
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__nv_bfloat16, 2, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmem));
        SB_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tn<__half, 2, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::kSmem)); });
    if (pair) {
        const bool res = ep.residual != nullptr;
        const int num_tiles = ceil_div(M, 2 * kBM) * ceil_div(N, kBN);
        const int clusters = std::min(num_tiles, num_sms() / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(C2::kThreads);
        cfg.dynamicSmemBytes = C2::kSmem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (dtype == SB_DTYPE_F16) {
            if (res) SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tn<__half, 2, 16, true>, ta, tb, ep, M, N, K));
            else SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tn<__half, 2, 16, false>, ta, tb, ep, M, N, K));
        } else {
            if (res) SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tn<__nv_bfloat16, 2, 16, true>, ta, tb, ep, M, N, K));
            else SB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_gemm_tn<__nv_bfloat16, 2, 16, false>, ta, tb, ep, M, N, K));
        }
    } else {
        const int num_tiles = ceil_div(M, kBM) * ceil_div(N, kBN);
        const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
        if (dtype == SB_DTYPE_F16)
            k_gemm_tn<__half, 1, 8, true><<<grid, C1::kThreads, C1::kSmem, st>>>(ta, tb, ep, M, N, K);
        else
            k_gemm_tn<__nv_bfloat16, 1, 8, true><<<grid, C1::kThreads, C1::kSmem, st>>>(ta, tb, ep, M, N, K);
    }
    g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

}  // namespace sb

extern "C" int sb_gemm_tn_dev(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                              void* out, int64_t ldo, int out_f32, const float* bias, int act,
                              const float* residual, int64_t ldr, int res_row_mod, void* stream) {
    SB_CHECK_ARG(A && W && out, "null pointer");
    SB_CHECK_ARG(dtype == SB_DTYPE_BF16 || dtype == SB_DTYPE_F16, "dtype must be SB_DTYPE_BF16 or SB_DTYPE_F16");
    sb::GemmEpilogue ep;
    ep.out = out; ep.ldo = (int)ldo; ep.out_f32 = out_f32; ep.bias = bias; ep.act = act;
    ep.residual = residual; ep.ldr = (int)ldr; ep.res_row_mod = res_row_mod;
    return sb::gemm_tn(dtype, A, lda, W, ldw, M, N, K, ep, (cudaStream_t)stream);
}
