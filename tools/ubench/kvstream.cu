// Does the cross-KV layout matter to HBM?  Streams the same bytes three ways with the same per-thread load depth:
//   A: (seq, head) block reads 1500 rows of 128 B at a 36 KB stride (the engine's [win*1500][L*2*d] layout, Small)
//   B: (seq, head) block reads 192 KB contiguous ([layer][kv][win][head][1500][64] layout)
//   C: (seq, key-slice) block reads 1536 B contiguous per key (all 12 heads of a key row), 36 KB stride between keys
// nvcc -arch=sm_100a -O3 -o kvstream kvstream.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
constexpr int kU = 8;
// mode 0: A, mode 1: B.  grid (12, B), 256 threads; lane -> (sub = lane/8, ch = lane%8)
__global__ void __launch_bounds__(256, 3) k_head(const char* base, int64_t row_stride, int64_t head_off, int64_t seq_stride, int n_keys, uint32_t* sink) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane >> 3, ch = lane & 7;
    const int key_l = warp * 4 + sub;
    uint32_t acc = 0;
    for (int pass = 0; pass < 2; ++pass) {     // K then V (V = second half of the buffer for simplicity: + pass * half)
        const char* p = base + blockIdx.y * seq_stride + blockIdx.x * head_off + ch * 16 + (int64_t)key_l * row_stride + pass * (row_stride > 128 ? 1536 : (int64_t)gridDim.y * seq_stride);
        uint4 ring[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) ring[u] = u * 32 + key_l < n_keys ? ldg_nc_v4(p + u * 32 * row_stride) : make_uint4(0, 0, 0, 0);
        for (int it = 0; it * 256 < n_keys; ++it) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint4 v = ring[u];
                const int key = it * 256 + u * 32 + key_l;
                if (key + 256 < n_keys) ring[u] = ldg_nc_v4(p + (int64_t)((it + 1) * 256 + u * 32) * row_stride);
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
// mode C: grid (8, B): CTA j of a sequence takes keys [j*188, ...); a warp instruction reads 512 contiguous bytes of one key row
__global__ void __launch_bounds__(256, 3) k_rows(const char* base, int64_t row_stride, int64_t seq_stride, int n_keys, int row_bytes, uint32_t* sink) {
    const int tid = threadIdx.x;
    const int per = (n_keys + gridDim.x - 1) / gridDim.x;
    const int k0 = blockIdx.x * per, k1 = min(n_keys, k0 + per);
    const int chunks_per_row = row_bytes / 16;                 // 96 for 1536 B
    const int total = (k1 - k0) * chunks_per_row;
    uint32_t acc = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const char* p = base + blockIdx.y * seq_stride + pass * row_bytes;
        uint4 ring[kU];
        auto addr = [&](int idx) { const int r = idx / chunks_per_row, c = idx - r * chunks_per_row; return p + (int64_t)(k0 + r) * row_stride + c * 16; };
#pragma unroll
        for (int u = 0; u < kU; ++u) ring[u] = tid + u * 256 < total ? ldg_nc_v4(addr(tid + u * 256)) : make_uint4(0, 0, 0, 0);
        for (int i0 = 0; i0 < total; i0 += kU * 256) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const uint4 v = ring[u];
                const int nxt = i0 + kU * 256 + u * 256 + tid;
                if (nxt < total) ring[u] = ldg_nc_v4(addr(nxt));
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
int main() {
    const int B = 64, H = 12, T = 1500, L = 12, d = 768;
    const int64_t row_stride = (int64_t)L * 2 * d * 2;                  // 36864 B
    const int64_t bytes = (int64_t)B * T * row_stride;                  // 3.5 GB
    char* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    uint32_t* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double moved = (double)B * H * T * 128 * 2;                   // K + V of one layer
    for (int mode = 0; mode < 3; ++mode) {
        for (int Bn : {64, 32, 16}) {
            float best = 1e9f;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                for (int l = 0; l < L; ++l) {       // 12 "layers": different addresses each launch, like the decode step
                    if (mode == 0) k_head<<<dim3(H, Bn), 256>>>(buf + l * 2 * d * 2, row_stride, 128, T * row_stride, T, sink);
                    else if (mode == 1) k_head<<<dim3(H, Bn), 256>>>(buf + (int64_t)l * 2 * B * H * T * 128, 128, (int64_t)T * 128, (int64_t)H * T * 128, T, sink);
                    else k_rows<<<dim3(8, Bn), 256>>>(buf + l * 2 * d * 2, row_stride, T * row_stride, T, d * 2, sink);
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            printf("mode %c  seqs %2d: %.1f us per launch, %.2f TB/s\n", 'A' + mode, Bn, best * 1000 / L, moved * Bn / B * L / (best * 1e-3) / 1e12);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
