// Capture front-end on sm_100a (rows a9-a13 of SURVEY.md 8(a)):
//   k_resample_poly<D> rubato FftFixedIn (96 / 64 / 48 / 32 -> 16 kHz) as a decimating FIR with the same 1026
//                      Blackman-Harris^2 sinc taps: D polyphase Toeplitz GEMMs on the f16 tensor cores (3-pass split)
//                      (reference: audio_toolkit/audio/resampler.rs:24, 51-56; equivalence of the FFT and
//                      FIR forms: SURVEY.md App. B, ~1e-9; the CUDA-core and 3xTF32 forms were removed)
//   k_resample_dense_* every other rate: rubato's block operator as a dense split-precision GEMM (tcgen05)
//   k_silero_features_fft / _direct   Silero v4 per-frame front: reflect pad, STFT conv, magnitude, log, adaptive
//                      normalisation, 4 separable conv blocks (reference: vad/silero.rs:41-44 ->
//                      vad-rs -> onnxruntime; graph first-hand from silero_vad_v4.onnx, App. A).  The shipped
//                      model's 258 x 256 STFT basis is a periodic-Hann-windowed DFT (checked at sb_vad_create to
//                      4e-7), so the convolution runs as 256-point FFTs in registers (25x fewer FLOPs); any other
//                      basis takes the direct-convolution kernel
//   k_silero_lstm      the two LSTM(64) layers + decoder + sigmoid, state carried across frames,
//                      gate weights resident in registers for the whole sequence
//   k_vad_plan / k_vad_compact   SmoothedVad onset / hangover / prefill FSM and the capture consumer's
//                      concatenation of kept frames (reference: vad/smoothed.rs:41-96,
//                      audio/recorder.rs:284-314)
// All fp32 (Silero decisions are thresholded: SURVEY 7.3 item 6).
#include "common.cuh"
#include <cuda_fp16.h>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace sb {
extern std::atomic<uint64_t> g_launches;

// ------------------------------------------------------------------------------------------
// mono down-mix (row a14): out[f] = (sum_c to_f32(in[f * C + c])) / C, f32 sum in channel order like the
// reference's iterator sum.  HBM-bound: C * sizeof(S) + 4 bytes per frame; consecutive threads take
// consecutive frames, so a warp reads one contiguous span of 32 * C samples.
// ------------------------------------------------------------------------------------------
template <typename S> __device__ __forceinline__ float to_sample_f32(S v);
template <> __device__ __forceinline__ float to_sample_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_sample_f32<int16_t>(int16_t v) { return (float)v * (1.0f / 32768.0f); }
template <> __device__ __forceinline__ float to_sample_f32<uint16_t>(uint16_t v) { return (float)((int)v - 32768) * (1.0f / 32768.0f); }

template <typename S>
__global__ void __launch_bounds__(256) k_downmix_mono(const S* __restrict__ in, int channels, int64_t in_stride, int64_t n_frames,
                                                      float* __restrict__ out, int64_t out_stride) {
    const S* src = in + (int64_t)blockIdx.y * in_stride;
    float* dst = out + (int64_t)blockIdx.y * out_stride;
    const float inv = 1.0f / (float)channels;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += (int64_t)gridDim.x * blockDim.x) {
        const S* fr = src + f * channels;
        if (channels == 1) { dst[f] = to_sample_f32<S>(__ldg(fr)); continue; }
        float a = 0.f;
        for (int c = 0; c < channels; ++c) a += to_sample_f32<S>(__ldg(fr + c));
        dst[f] = a / (float)channels;
        (void)inv;
    }
}

// ------------------------------------------------------------------------------------------
// TF32 mma.sync helpers (block-1 pointwise product of the Silero front)
// ------------------------------------------------------------------------------------------
// cheap split for operands whose partner is split exactly: hi = the top 10 mantissa bits (truncated), lo = the exact
// remainder, of which the tensor core reads the top 10 bits again: 2^-21 relative, two instructions (cvt.rna.tf32 is
// emulated with four on sm_100)
__device__ __forceinline__ void split_tf32_trunc(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// The decimating FIR as D polyphase branches on the f16 tensor cores (the 3xTF32 Toeplitz form it replaced was removed: 26.1 ms per 1000 streams, 3.1e-6).
//   y[m] = sum_u h[u] x[D m - u] = sum_p sum_q h[D q + p] x_p[m - q],   x_p[r] = x[D r - p],  q < Q = ceil(T / D)
// Each branch is a stride-1 convolution, i.e. a Toeplitz GEMM over 32 consecutive outputs per column:
//   A_p[i][j] = h_p[Q - 1 + i - j]  (32 x (Q + 31), constant),   B_p[j][n] = x_p[m0 + 32 n - (Q - 1) + j]  (a sliding window).
// mma.sync.m16n8k16 f16 with the 3-pass split a_hi b_hi + a_hi b_lo + a_lo b_hi (hi = f16(v), lo = f16(v - hi); the taps
// are scaled by 2^12 first so that their lo parts stay normal): products carry 22 bits, accumulation is f32 -- 1.7e-6
// against the f64 rubato restatement (3xTF32 form: 3.1e-6), at twice the MACs per instruction.  Operand traffic is what
// bounds this kernel (the first f16 version sat at 96 % of the L1 / shared-memory pipe with the tensor pipe at 65 %), so:
//   * rows 16..31 of A_p are rows 0..15 shifted by one k-step (Toeplitz), so the second row tile reuses the fragments the
//     first one used a step earlier: one fragment load (hi | lo, pre-arranged on the host, 2 x 16 bytes per lane) feeds
//     two row tiles x eight column tiles = 48 MMAs;
//   * B fragments by ldmatrix.x4 (hi k 0-7, hi k 8-15, lo k 0-7, lo k 8-15), each shared by the two row tiles.  The eight
//     windows of a tile start 32 halves = four 16-byte chunks apart; chunk c is stored at c ^ ((c >> 3) & 3), which puts
//     the eight rows of every 8 x 8 matrix into eight different bank groups with a single copy of the signal;
//   * no operand splitting in the loop (the TF32 form spent 16 emulated cvt.rna per 6 MMAs: it was issue-bound).
// CTA = 64 threads = 2 warps x (2 row tiles x 8 column tiles) = 4096 outputs of one stream; 55 KB of shared memory.
// ------------------------------------------------------------------------------------------
constexpr int kRpTile = 4096;
constexpr int kRpThreads = 64;
constexpr int kRpNT = 8;                 // column tiles (of 8 x 32 outputs) per warp
constexpr int kRpScaleLog2 = 12;

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// smem: f16 arrays [part hi|lo][branch p][L], L a multiple of 64 halves (every array starts on bank 0).  SP = k-steps of one
// 16-row tile; the 32-row product takes SP + 1
__host__ __device__ constexpr int rp_array_len(int SP) { return (kRpTile + 16 * (SP + 1) + 63) & ~63; }
__device__ __forceinline__ int rp_swz(int chunk) { return chunk ^ ((chunk >> 3) & 3); }

template <int D>
__global__ void __launch_bounds__(kRpThreads) k_resample_poly(const float* __restrict__ x, int64_t x_stride, int n_in,
                                                             float* __restrict__ y, int64_t y_stride, int n_out,
                                                             const uint4* __restrict__ afrag, int Q, int SP) {
    extern __shared__ __align__(128) unsigned char s_rp[];
    __half* xh = reinterpret_cast<__half*>(s_rp);
    const int L = rp_array_len(SP);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int stream = blockIdx.y;
    const int m0 = blockIdx.x * kRpTile;
    const float* xin = x + (int64_t)stream * x_stride;
    // A fragments: [D][SP + 1][32 lanes][hi | lo]; the last step of every branch is all zero (it is the second row tile's
    // last step, and what that tile sees as "previous step" when the next branch starts)
    const int n_steps = D * (SP + 1);
    const uint4* ap = afrag + lane * 2;
    uint4 ah1 = __ldg(ap), al1 = __ldg(ap + 1);                            // step 0
    uint4 ah2 = __ldg(ap + 64), al2 = __ldg(ap + 65);                      // step 1
    const int rbase = m0 - (Q - 1);
    const int nload = kRpTile + 16 * (SP + 1) - 32;                        // last window starts at 32 * 127
    // branch signals x_p[r] = x[D r - p]: the 2 D consecutive samples x[D r - D + 1 .. D r + D] give the pair (r, r + 1) of
    // every branch
    for (int r2 = 2 * tid; r2 < nload; r2 += 2 * kRpThreads) {
        const int s0 = D * (rbase + r2);                                   // D n_out < 2^31: sb_resample_dev checks
        float v[2 * D];
#pragma unroll
        for (int e = 0; e < 2 * D; ++e) {
            const int si = s0 - (D - 1) + e;
            v[e] = (unsigned)si < (unsigned)n_in ? __ldg(xin + si) : 0.0f;
        }
        const int off = (rp_swz(r2 >> 3) << 3) | (r2 & 7);
#pragma unroll
        for (int p = 0; p < D; ++p) {
            const float v0 = v[D - 1 - p], v1 = v[2 * D - 1 - p];          // x[D r - p], x[D (r + 1) - p]
            const __half2 hi = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            *reinterpret_cast<__half2*>(xh + p * L + off) = hi;
            *reinterpret_cast<__half2*>(xh + (D + p) * L + off) = lo;
        }
    }
    __syncthreads();
    // ldmatrix row of this lane: matrix lane >> 3 = (part, k half), row lane & 7 = column of the tile.  Chunk (16 bytes)
    // of the row at k-step s, tile nt: cb + 2 s + 32 nt, swizzled (the swizzle bits do not depend on nt)
    const int part = lane >> 4, khalf = (lane >> 3) & 1, r = lane & 7;
    const int cb = 4 * (8 * kRpNT * warp + r) + khalf;
    uint32_t bbase = (uint32_t)__cvta_generic_to_shared(xh + (part * D) * L);
    float acc[2][kRpNT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < kRpNT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    uint4 ahp = make_uint4(0u, 0u, 0u, 0u), alp = ahp;                     // previous step's fragments = the second row tile's
    int sI = 0;
    for (int it = 0; it < n_steps; ++it) {
        const uint4 ah = ah1, al = al1;
        ah1 = ah2; al1 = al2;
        const int nx = min(it + 2, n_steps - 1);
        ah2 = __ldg(ap + (size_t)nx * 64);
        al2 = __ldg(ap + (size_t)nx * 64 + 1);
        const uint32_t baddr = bbase + ((uint32_t)rp_swz(cb + 2 * sI) << 4);
#pragma unroll
        for (int nt = 0; nt < kRpNT; nt += 2) {          // two column tiles per pass: a dependent MMA follows three others
            uint32_t bq[2][4];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(bq[u][0]), "=r"(bq[u][1]), "=r"(bq[u][2]), "=r"(bq[u][3]) : "r"(baddr + (nt + u) * 512));
#pragma unroll
            for (int u = 0; u < 2; ++u) { mma_f16_16816(acc[0][nt + u], ah, bq[u][0], bq[u][1]); mma_f16_16816(acc[1][nt + u], ahp, bq[u][0], bq[u][1]); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { mma_f16_16816(acc[0][nt + u], ah, bq[u][2], bq[u][3]); mma_f16_16816(acc[1][nt + u], ahp, bq[u][2], bq[u][3]); }
#pragma unroll
            for (int u = 0; u < 2; ++u) { mma_f16_16816(acc[0][nt + u], al, bq[u][0], bq[u][1]); mma_f16_16816(acc[1][nt + u], alp, bq[u][0], bq[u][1]); }
        }
        ahp = ah; alp = al;
        if (++sI == SP + 1) { sI = 0; bbase += (uint32_t)(L * 2); }
    }
    // C[i][n]: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)  ->  output m0 + 32 n + 16 mt + i
    float* yo = y + (int64_t)stream * y_stride;
    const int g = lane >> 2, t = lane & 3;
    constexpr float kInv = 1.0f / (float)(1 << kRpScaleLog2);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < kRpNT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = 8 * (kRpNT * warp + nt) + 2 * t + (e & 1);
                const int m = m0 + 32 * n + 16 * mt + g + ((e >> 1) << 3);
                if (m < n_out) yo[m] = acc[mt][nt][e] * kInv;
            }
}

// ------------------------------------------------------------------------------------------
// Rational ratios (44.1 kHz, 22.05 kHz, 8 kHz ... -> 16 kHz): rubato's FftFixedIn as the dense linear operator it is.
// One rubato block maps N1 = fft_size_in input samples to N2 = fft_size_out output samples through rfft(2 N1) -> spectrum
// x filter, truncated to L bins -> irfft(2 N2) -> overlap-add of the two output halves.  Every step is linear, so
//   out_b[m] = sum_j A[m][j] x_b[j] + sum_j A[N2 + m][j] x_{b-1}[j],
//   A[m][j]  = filt[0] + 2 sum_{k=1}^{L-1} Re(filt[k] e^{2 pi i k (m / 2 N2 - j / 2 N1)}),   m < 2 N2, j < N1
// i.e. C[block][m] = X[block][.] W[m][.]^T with X[block] = the 2 N1 consecutive samples x_{b-1} | x_b (materialised as f16
// hi | lo rows, K padded to a multiple of 32) and W = [A_bottom | A_top] (built once per resampler, on the device, in f64).
// The product runs on the tcgen05 GEMM of the encoder (sb_gemm_tn_dev) in three passes hi hi + hi lo + lo hi with the f32
// running sum as the residual input: 22-bit products like the polyphase kernel.  (The integer ratios keep that kernel: a
// Toeplitz operator needs K = 1026 + 45 per output there, the dense block operator 2 N1 = 2646.)
// ------------------------------------------------------------------------------------------
constexpr float kRdScale = 1024.0f;

// W[m][kk], m < N2, kk < Kp: kk < N1 pairs with x_{b-1}[kk] (A row N2 + m), N1 <= kk < 2 N1 with x_b[kk - N1] (A row m)
__global__ void k_resample_dense_matrix(const double2* __restrict__ filt, int N1, int N2, int L, int Kp,
                                        __half* __restrict__ whi, __half* __restrict__ wlo) {
    const int kk = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (kk >= Kp) return;
    double a = 0.0;
    if (kk < 2 * N1 && m < N2) {                      // rows N2 .. N2p-1 pad the GEMM's N to a multiple of 8
        const int row = kk < N1 ? N2 + m : m, j = kk < N1 ? kk : kk - N1;
        // phase of bin k: 2 pi k (row N1 - j N2) / (2 N1 N2), reduced exactly in integers
        const long long P = 2LL * N1 * N2;
        long long t = ((long long)row * N1 - (long long)j * N2) % P;
        if (t < 0) t += P;
        a = filt[0].x;
        // e^{i k theta} by recurrence, re-seeded exactly every 64 bins (the error of the recurrence grows like k x 1e-16)
        double wr, wi;
        sincospi(2.0 * (double)t / (double)P, &wi, &wr);
        double zr = 1.0, zi = 0.0;
        for (int k = 1; k < L; ++k) {
            if ((k & 63) == 0) {
                const long long u = ((long long)k * t) % P;
                sincospi(2.0 * (double)u / (double)P, &zi, &zr);
            } else {
                const double nr = zr * wr - zi * wi;
                zi = zr * wi + zi * wr; zr = nr;
            }
            a += 2.0 * (filt[k].x * zr - filt[k].y * zi);
        }
    }
    const float v = (float)(a * (double)kRdScale);
    const __half hi = __float2half_rn(v);
    whi[(size_t)m * Kp + kk] = hi;
    wlo[(size_t)m * Kp + kk] = __float2half_rn((float)(a * (double)kRdScale - (double)__half2float(hi)));
}

// X rows: row (stream, b) = samples [(b - 1) N1, (b + 1) N1) of the stream (zero outside [0, n_in)), f16 hi | lo
__global__ void k_resample_dense_rows(const float* __restrict__ x, int64_t x_stride, int n_in, int n_blocks, int N1, int Kp,
                                      __half* __restrict__ xhi, __half* __restrict__ xlo) {
    const int row = blockIdx.x;                       // stream * n_blocks + b
    const int stream = row / n_blocks, b = row - stream * n_blocks;
    const float* src = x + (int64_t)stream * x_stride;
    const int64_t base = (int64_t)(b - 1) * N1;
    for (int kk = threadIdx.x; kk < Kp; kk += blockDim.x) {
        const int64_t si = base + kk;
        const float v = (kk < 2 * N1 && si >= 0 && si < n_in) ? __ldg(src + si) : 0.0f;
        const __half hi = __float2half_rn(v);
        xhi[(size_t)row * Kp + kk] = hi;
        xlo[(size_t)row * Kp + kk] = __float2half_rn(v - __half2float(hi));
    }
}

// out[stream][b N2 + m] = C[(stream, b)][m] / scale
__global__ void k_resample_dense_store(const float* __restrict__ c, int n_blocks, int N2, int N2p, float* __restrict__ out, int64_t out_stride) {
    const int row = blockIdx.x;
    const int stream = row / n_blocks, b = row - stream * n_blocks;
    float* dst = out + (int64_t)stream * out_stride + (int64_t)b * N2;
    for (int m = threadIdx.x; m < N2; m += blockDim.x) dst[m] = c[(size_t)row * N2p + m] * (1.0f / kRdScale);
}

// ------------------------------------------------------------------------------------------
// Silero v4 (16 kHz) weights, device resident
// ------------------------------------------------------------------------------------------
struct SileroDev {
    const float* basis_t;      // [256][258]  (transposed STFT basis: coalesced over channels)
    const float* norm_filter;  // [7]
    const float *b1_dw_w, *b1_dw_b, *b1_pw_w, *b1_pw_b, *b1_proj_w, *b1_proj_b, *b1_down_w, *b1_down_b;
    const float *b2_dw_w, *b2_dw_b, *b2_pw_w, *b2_pw_b, *b2_proj_w, *b2_proj_b, *b2_down_w, *b2_down_b;
    const float *b3_dw_w, *b3_dw_b, *b3_pw_w, *b3_pw_b, *b3_down_w, *b3_down_b;
    const float *b4_dw_w, *b4_dw_b, *b4_pw_w, *b4_pw_b, *b4_proj_w, *b4_proj_b, *b4_down_w, *b4_down_b;
    const float *lstm_w[2], *lstm_r[2], *lstm_b[2];
    const float *dec_w, *dec_b;
    // FFT form (sb_vad_create): analysis window = basis row 0, W128^{n2 k1} laid out [k1][n2] and e^{-2 pi i k / 256} in f64, the
    // residual of the stored basis against the exact windowed DFT (x 2^24, bf16, mma A-fragment order), block-1 pointwise
    // weights transposed to [ci][16]
    const float* win;
    const double2 *tw, *tw2;
    const uint4* delta_frag;
    const float *b1_pw_t, *b1_proj_t;
    const uint4* b1_frag;   // block-1 pointwise weights as TF32 hi | lo mma A fragments: [66 k-steps][32 lanes][2]
    // [ci][co] transposes of the k1 convolutions of blocks 1-4 (mix_relu)
    const float *b1_down_t, *b2_pw_t, *b2_proj_t, *b2_down_t, *b3_pw_t, *b3_down_t, *b4_pw_t, *b4_proj_t, *b4_down_t;
};

constexpr int kSfFrames = 4;                 // frames per CTA (72 KB of shared memory: three CTAs per SM)
constexpr int kSfCols = kSfFrames * 7;       // 56 STFT columns
constexpr int kSfThreads = 288;
constexpr int kSfPadLen = 672;

// generic small helpers operating on [C][kSfFrames][T] tiles in shared memory -------------------
// depthwise conv k5 pad 2 (zero pad inside each frame), + bias, ReLU
__device__ void dw5_relu(const float* in, float* out, const float* w, const float* b, int C, int T) {
    const int total = C * kSfFrames * T;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / (kSfFrames * T), rem = i - c * kSfFrames * T;
        const int f = rem / T, t = rem - f * T;
        const float* row = in + (c * kSfFrames + f) * T;
        float a = __ldg(b + c);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int tt = t + k - 2;
            if (tt >= 0 && tt < T) a = fmaf(__ldg(w + c * 5 + k), row[tt], a);
        }
        out[i] = fmaxf(a, 0.f);
    }
}
// pointwise: out[co][f][t] = relu( W1[co,:] . in1[:,f,t] + b1 (+ W2[co,:] . in2[:,f,t] + b2) (+ res[co][f][t]) )
__device__ void pw_relu(const float* in1, const float* w1, const float* b1, const float* in2, const float* w2,
                        const float* b2, const float* res, float* out, int Cin, int Cout, int T) {
    const int FT = kSfFrames * T;
    const int total = Cout * FT;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int co = i / FT, ft = i - co * FT;
        float a = __ldg(b1 + co);
        for (int ci = 0; ci < Cin; ++ci) a = fmaf(__ldg(w1 + co * Cin + ci), in1[ci * FT + ft], a);
        if (in2) {
            a += __ldg(b2 + co);
            for (int ci = 0; ci < Cin; ++ci) a = fmaf(__ldg(w2 + co * Cin + ci), in2[ci * FT + ft], a);
        }
        if (res) a += res[i];
        out[i] = fmaxf(a, 0.f);
    }
}
// k1 conv with stride s over T: out[co][f][to] = relu(W[co,:] . in[:, f, s*to] + b)
__device__ void down_relu(const float* in, const float* w, const float* b, float* out, int C_in, int C_out, int T_in,
                          int T_out, int stride) {
    const int total = C_out * kSfFrames * T_out;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int co = i / (kSfFrames * T_out), rem = i - co * kSfFrames * T_out;
        const int f = rem / T_out, to = rem - f * T_out;
        float a = __ldg(b + co);
        for (int ci = 0; ci < C_in; ++ci) a = fmaf(__ldg(w + co * C_in + ci), in[(ci * kSfFrames + f) * T_in + to * stride], a);
        out[i] = fmaxf(a, 0.f);
    }
}

struct SileroSmem {
    union {
        float xs[kSfFrames * kSfPadLen];  // reflect-padded frames: dead once the STFT is done ...
        struct {                          // ... so the small late-stage buffers live in the same bytes (3 CTAs per SM)
            float y2[16 * kSfFrames * 4];         // T = 4
            float r2[16 * kSfFrames * 4];
            float z1[32 * kSfFrames * 4];
            float z2[32 * kSfFrames * 2];         // T = 2
            float r3[32 * kSfFrames * 2];
            float u1[32 * kSfFrames * 2];
            float u2[32 * kSfFrames];             // T = 1
            float r4[32 * kSfFrames];
            float v1[64 * kSfFrames];
        };
    };
    float x1[258 * kSfCols];              // [258][frames][7]: magnitude (0..128) | norm (129..257)
    float r1[258 * kSfCols];              // depthwise output / scratch
    float y1[16 * kSfCols];               // block-1 pre-downsample
    float mean[kSfFrames * 7];
    float mm[kSfFrames];
};
static_assert(sizeof(SileroSmem) * 3 <= 227 * 1024 - 3 * 1024, "three CTAs per SM");

// pcm: [n_streams][n_frames*480]; out: [n_streams][n_frames][64]
__global__ void __launch_bounds__(kSfThreads, 3) k_silero_features_direct(const float* __restrict__ pcm, int64_t pcm_stride, int n_frames,
                                                                       float* __restrict__ out, SileroDev wts) {
    extern __shared__ __align__(16) unsigned char smem_raw_s[];
    SileroSmem& s = *reinterpret_cast<SileroSmem*>(smem_raw_s);
    const int stream = blockIdx.y;
    const int f0 = blockIdx.x * kSfFrames;
    const float* src = pcm + (int64_t)stream * pcm_stride;
    // reflect pad 96 on both sides of every 480-sample frame
    for (int i = threadIdx.x; i < kSfFrames * kSfPadLen; i += blockDim.x) {
        const int f = i / kSfPadLen, j = i - f * kSfPadLen;
        int k = j - 96;
        if (k < 0) k = -k;
        if (k >= 480) k = 2 * 479 - k;
        s.xs[i] = (f0 + f < n_frames) ? __ldg(src + (int64_t)(f0 + f) * 480 + k) : 0.0f;
    }
    __syncthreads();
    // STFT conv: item = (frame pair hh, channel pair i): re/im of the 2 x 7 columns of two frames.
    // Blocked summation (8 blocks of 32 taps) keeps the fp32 round-off of the 256-term dot products
    // well below the 2^-20 scale that log(1 + 2^20 |X|) magnifies on quiet bins.
    // The window samples are read four taps at a time (one 128-bit shared-memory load feeds 8 FMAs): the loop is
    // bound by the FMA pipe, not by the load/store unit.
    {
        constexpr int kPairs = kSfFrames / 2;
        for (int item = threadIdx.x; item < kPairs * 129; item += blockDim.x) {
            const int hh = item / 129, i = item - hh * 129;
            float re[14], im[14];
#pragma unroll
            for (int c = 0; c < 14; ++c) { re[c] = 0.f; im[c] = 0.f; }
            const float* xb = s.xs + (hh * 2) * kSfPadLen;
#pragma unroll 1
            for (int kb = 0; kb < 256; kb += 32) {
                float tr[14], ti[14];
#pragma unroll
                for (int c = 0; c < 14; ++c) { tr[c] = 0.f; ti[c] = 0.f; }
#pragma unroll 2
                for (int k = kb; k < kb + 32; k += 4) {
                    float br[4], bi[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        br[e] = __ldg(wts.basis_t + (k + e) * 258 + i);
                        bi[e] = __ldg(wts.basis_t + (k + e) * 258 + 129 + i);
                    }
#pragma unroll
                    for (int f = 0; f < 2; ++f)
#pragma unroll
                        for (int t = 0; t < 7; ++t) {
                            const float4 xv = *reinterpret_cast<const float4*>(xb + f * kSfPadLen + 64 * t + k);
                            float a = tr[f * 7 + t], b = ti[f * 7 + t];
                            a = fmaf(br[0], xv.x, a); b = fmaf(bi[0], xv.x, b);
                            a = fmaf(br[1], xv.y, a); b = fmaf(bi[1], xv.y, b);
                            a = fmaf(br[2], xv.z, a); b = fmaf(bi[2], xv.z, b);
                            a = fmaf(br[3], xv.w, a); b = fmaf(bi[3], xv.w, b);
                            tr[f * 7 + t] = a; ti[f * 7 + t] = b;
                        }
                }
#pragma unroll
                for (int c = 0; c < 14; ++c) { re[c] += tr[c]; im[c] += ti[c]; }
            }
#pragma unroll
            for (int c = 0; c < 14; ++c) {
                const float mag = sqrtf(re[c] * re[c] + im[c] * im[c]);
                s.x1[i * kSfCols + hh * 14 + c] = mag;
                s.r1[i * kSfCols + hh * 14 + c] = log1pf(1048576.0f * mag);   // spect (scratch)
            }
        }
    }
    __syncthreads();
    // adaptive normalisation: mean over the 129 bins, reflect pad 3, 7-tap filter, mean over T
    if (threadIdx.x < kSfCols) {
        float a = 0.f;
        for (int i = 0; i < 129; ++i) a += s.r1[i * kSfCols + threadIdx.x];
        s.mean[threadIdx.x] = a / 129.0f;
    }
    __syncthreads();
    if (threadIdx.x < kSfFrames) {
        const float* m = s.mean + threadIdx.x * 7;
        float p[13];
        p[0] = m[3]; p[1] = m[2]; p[2] = m[1];
#pragma unroll
        for (int t = 0; t < 7; ++t) p[3 + t] = m[t];
        p[10] = m[5]; p[11] = m[4]; p[12] = m[3];
        float acc = 0.f;
        for (int t = 0; t < 7; ++t) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) a = fmaf(__ldg(wts.norm_filter + k), p[t + k], a);
            acc += a;
        }
        s.mm[threadIdx.x] = acc / 7.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 129 * kSfCols; i += blockDim.x) {
        const int col = i % kSfCols;
        s.x1[129 * kSfCols + i] = s.r1[i] - s.mm[col / 7];
    }
    __syncthreads();
    // block 1 (258 -> 16, T 7 -> 4)
    dw5_relu(s.x1, s.r1, wts.b1_dw_w, wts.b1_dw_b, 258, 7);
    __syncthreads();
    pw_relu(s.r1, wts.b1_pw_w, wts.b1_pw_b, s.x1, wts.b1_proj_w, wts.b1_proj_b, nullptr, s.y1, 258, 16, 7);
    __syncthreads();
    down_relu(s.y1, wts.b1_down_w, wts.b1_down_b, s.y2, 16, 16, 7, 4, 2);
    __syncthreads();
    // block 2 (16 -> 32, T 4 -> 2)
    dw5_relu(s.y2, s.r2, wts.b2_dw_w, wts.b2_dw_b, 16, 4);
    __syncthreads();
    pw_relu(s.r2, wts.b2_pw_w, wts.b2_pw_b, s.y2, wts.b2_proj_w, wts.b2_proj_b, nullptr, s.z1, 16, 32, 4);
    __syncthreads();
    down_relu(s.z1, wts.b2_down_w, wts.b2_down_b, s.z2, 32, 32, 4, 2, 2);
    __syncthreads();
    // block 3 (32 -> 32 with identity residual, T 2 -> 1)
    dw5_relu(s.z2, s.r3, wts.b3_dw_w, wts.b3_dw_b, 32, 2);
    __syncthreads();
    pw_relu(s.r3, wts.b3_pw_w, wts.b3_pw_b, nullptr, nullptr, nullptr, s.z2, s.u1, 32, 32, 2);
    __syncthreads();
    down_relu(s.u1, wts.b3_down_w, wts.b3_down_b, s.u2, 32, 32, 2, 1, 2);
    __syncthreads();
    // block 4 (32 -> 64, T 1)
    dw5_relu(s.u2, s.r4, wts.b4_dw_w, wts.b4_dw_b, 32, 1);
    __syncthreads();
    pw_relu(s.r4, wts.b4_pw_w, wts.b4_pw_b, s.u2, wts.b4_proj_w, wts.b4_proj_b, nullptr, s.v1, 32, 64, 1);
    __syncthreads();
    // final 64 -> 64 (stride 1), ReLU, write [frame][64]
    for (int i = threadIdx.x; i < 64 * kSfFrames; i += blockDim.x) {
        const int f = i / 64, co = i - f * 64;
        if (f0 + f >= n_frames) continue;
        float a = __ldg(wts.b4_down_b + co);
        for (int ci = 0; ci < 64; ++ci) a = fmaf(__ldg(wts.b4_down_w + co * 64 + ci), s.v1[ci * kSfFrames + f], a);
        out[((int64_t)stream * n_frames + f0 + f) * 64 + co] = fmaxf(a, 0.f);
    }
}

// ------------------------------------------------------------------------------------------
// FFT form of the same front.  The STFT basis of the shipped model is row c: win[n] cos(2 pi c n / 256), row 129 + c:
// -win[n] sin(2 pi c n / 256) (win = periodic Hann) ROUNDED TO f32: B = fl32(W D).  The kernel evaluates the reference's
// operator exactly as  B x = D (W x) + delta x,  delta = B - W D  (|delta| <= 7.7e-8, known to f64 on the host):
//   * D (W x): one 128-point complex FFT per real column, z[n] = x[2n] + i x[2n+1], then the real-input split
//     X[k] = E[k] + e^{-2 pi i k / 256} O[k], E = (Z[k] + conj Z[128-k]) / 2, O = (Z[k] - conj Z[128-k]) / 2i -- in f64:
//     log(1 + 2^20 |X|) magnifies the round-off that loud bins leave on quiet ones, and an f32 FFT alone moved the
//     probabilities by up to 3.5e-4 (the direct f32 convolution needs blocked summation for the same reason).  One FFT =
//     8 lanes: radix-16 over n1 (n = 8 n1 + n2) in registers, twiddle W128^{n2 k1}, 16 x 8 exchange through shared memory
//     (row stride 9 double2: conflict-free both ways, __syncwarp only), two radix-8 over n2 per lane, partner bins by
//     shuffle.  25x fewer FLOPs than the convolution.
//   * delta x: needs 1 % relative accuracy only (it is a 2e-4 effect on the probabilities) -> one bf16 mma.sync pass
//     on the tensor cores: delta * 2^24 in bf16, pre-arranged on the host in m16n8k16 A-fragment order (256 rows: 129
//     cosine + the 127 non-zero sine rows), the columns read as overlapping windows of a bf16 copy of the frames.
// Without the delta term the FFT differs from the reference's stored operator by 2.4e-4 on the probabilities.
// 4 frames = 28 columns = 28 FFTs = 224 threads per CTA, 107 KB of shared memory (two CTAs per SM).
// ------------------------------------------------------------------------------------------
// out[co][p] = relu(b1[co] + sum_ci w1t[ci][co] in1[ci][off(p)] (+ b2[co] + sum_ci w2t[ci][co] in2[ci][off(p)]) (+ res[co][p])):
// the pointwise and the strided k1 convolutions of blocks 1-4 on [C][frames x T] tiles.  Weights are transposed to
// [ci][co] (sb_vad_create) and co is the fastest thread index, so one warp-wide weight load is one contiguous row; the
// activation is a shared-memory broadcast; PPT outputs per thread share the weight.
template <int CIN, int COUT, int P, int PPT, bool TWO, bool RES, typename Off>
__device__ __forceinline__ void mix_relu(const float* __restrict__ in1, const float* __restrict__ w1t, const float* __restrict__ b1,
                                         const float* __restrict__ in2, const float* __restrict__ w2t, const float* __restrict__ b2,
                                         const float* __restrict__ res, float* __restrict__ out, int in_stride, Off off) {
    constexpr int PG = P / PPT;
    static_assert(P % PPT == 0, "P must be a multiple of PPT");
    for (int item = threadIdx.x; item < COUT * PG; item += blockDim.x) {
        const int co = item % COUT, pg = item / COUT;
        int o[PPT];
        float a[PPT];
        const float bias = TWO ? b1[co] + b2[co] : b1[co];
#pragma unroll
        for (int e = 0; e < PPT; ++e) { o[e] = off(pg + PG * e); a[e] = bias; }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float w = w1t[ci * COUT + co];
#pragma unroll
            for (int e = 0; e < PPT; ++e) a[e] = fmaf(w, in1[ci * in_stride + o[e]], a[e]);
            if (TWO) {
                const float w2 = w2t[ci * COUT + co];
#pragma unroll
                for (int e = 0; e < PPT; ++e) a[e] = fmaf(w2, in2[ci * in_stride + o[e]], a[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < PPT; ++e) {
            const int pp = pg + PG * e;
            float v = a[e];
            if (RES) v += res[co * P + pp];
            out[co * P + pp] = fmaxf(v, 0.f);
        }
    }
}

// the same with the identity position map: PPT consecutive positions x NCO output channels (co, co + COUT / NCO, ...) per thread;
// the activations are read as float4 broadcasts and shared by the NCO channels (the stage is bound by these loads)
template <int CIN, int COUT, int P, int PPT, int NCO, bool TWO, bool RES>
__device__ __forceinline__ void mix4_relu(const float* __restrict__ in1, const float* __restrict__ w1t, const float* __restrict__ b1,
                                          const float* __restrict__ in2, const float* __restrict__ w2t, const float* __restrict__ b2,
                                          const float* __restrict__ res, float* __restrict__ out) {
    static_assert(PPT % 4 == 0 && P % PPT == 0 && COUT % NCO == 0, "PPT a multiple of 4 dividing P, NCO dividing COUT");
    constexpr int CG = COUT / NCO;
    for (int item = threadIdx.x; item < CG * (P / PPT); item += blockDim.x) {
        const int co = item % CG, p0 = (item / CG) * PPT;
        float a[NCO][PPT];
#pragma unroll
        for (int c = 0; c < NCO; ++c) {
            const float bias = TWO ? b1[co + c * CG] + b2[co + c * CG] : b1[co + c * CG];
#pragma unroll
            for (int e = 0; e < PPT; ++e) a[c][e] = bias;
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            float w[NCO];
#pragma unroll
            for (int c = 0; c < NCO; ++c) w[c] = w1t[ci * COUT + co + c * CG];
#pragma unroll
            for (int q = 0; q < PPT / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(in1 + ci * P + p0 + 4 * q);
#pragma unroll
                for (int c = 0; c < NCO; ++c) {
                    a[c][4 * q] = fmaf(w[c], v.x, a[c][4 * q]); a[c][4 * q + 1] = fmaf(w[c], v.y, a[c][4 * q + 1]);
                    a[c][4 * q + 2] = fmaf(w[c], v.z, a[c][4 * q + 2]); a[c][4 * q + 3] = fmaf(w[c], v.w, a[c][4 * q + 3]);
                }
            }
            if (TWO) {
#pragma unroll
                for (int c = 0; c < NCO; ++c) w[c] = w2t[ci * COUT + co + c * CG];
#pragma unroll
                for (int q = 0; q < PPT / 4; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(in2 + ci * P + p0 + 4 * q);
#pragma unroll
                    for (int c = 0; c < NCO; ++c) {
                        a[c][4 * q] = fmaf(w[c], v.x, a[c][4 * q]); a[c][4 * q + 1] = fmaf(w[c], v.y, a[c][4 * q + 1]);
                        a[c][4 * q + 2] = fmaf(w[c], v.z, a[c][4 * q + 2]); a[c][4 * q + 3] = fmaf(w[c], v.w, a[c][4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCO; ++c)
#pragma unroll
            for (int q = 0; q < PPT / 4; ++q) {
                float4 v = make_float4(a[c][4 * q], a[c][4 * q + 1], a[c][4 * q + 2], a[c][4 * q + 3]);
                if (RES) {
                    const float4 r = *reinterpret_cast<const float4*>(res + (co + c * CG) * P + p0 + 4 * q);
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                *reinterpret_cast<float4*>(out + (co + c * CG) * P + p0 + 4 * q) = v;
            }
    }
}

template <typename T>
__device__ __forceinline__ void dft4r(T& r0, T& i0, T& r1, T& i1, T& r2, T& i2, T& r3, T& i3) {
    const T t0r = r0 + r2, t0i = i0 + i2, t1r = r0 - r2, t1i = i0 - i2;
    const T t2r = r1 + r3, t2i = i1 + i3, t3r = r1 - r3, t3i = i1 - i3;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r1 = t1r + t3i; i1 = t1i - t3r;   // t1 - i t3
    r3 = t1r - t3i; i3 = t1i + t3r;   // t1 + i t3
}
// 16-point forward DFT in registers: in x[n], out X[k] (natural order, in place)
template <typename T>
__device__ __forceinline__ void dft16(T (&xr)[16], T (&xi)[16]) {
    constexpr double c1 = 0.92387953251128673848, c2 = 0.70710678118654752440, c3 = 0.38268343236508977173;
    constexpr double kC[10] = {1.0, c1, c2, c3, 0.0, -c3, -c2, -c1, -1.0, -c1};
    constexpr double kS[10] = {0.0, c3, c2, c1, 1.0, c1, c2, c3, 0.0, -c3};
    // n = 4 n1 + n2: radix-4 over n1 for each n2 -> element 4 k1 + n2
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
        dft4r(xr[n2], xi[n2], xr[4 + n2], xi[4 + n2], xr[8 + n2], xi[8 + n2], xr[12 + n2], xi[12 + n2]);
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
        for (int n2 = 1; n2 < 4; ++n2) {          // W16^{n2 k1} = c - i s
            const T c = (T)kC[n2 * k1], sn = (T)kS[n2 * k1];
            const T r = xr[4 * k1 + n2], im = xi[4 * k1 + n2];
            xr[4 * k1 + n2] = r * c + im * sn;
            xi[4 * k1 + n2] = im * c - r * sn;
        }
    // radix-4 over n2 for each k1: element 4 k1 + k2 = X[k1 + 4 k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft4r(xr[4 * k1], xi[4 * k1], xr[4 * k1 + 1], xi[4 * k1 + 1], xr[4 * k1 + 2], xi[4 * k1 + 2], xr[4 * k1 + 3], xi[4 * k1 + 3]);
    T tr[16], ti[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) { tr[k1 + 4 * k2] = xr[4 * k1 + k2]; ti[k1 + 4 * k2] = xi[4 * k1 + k2]; }
#pragma unroll
    for (int k = 0; k < 16; ++k) { xr[k] = tr[k]; xi[k] = ti[k]; }
}
// 8-point forward DFT in registers (natural order, in place)
template <typename T>
__device__ __forceinline__ void dft8(T (&xr)[8], T (&xi)[8]) {
    constexpr T h = (T)0.70710678118654752440;
    dft4r(xr[0], xi[0], xr[2], xi[2], xr[4], xi[4], xr[6], xi[6]);   // E[k] at 2k
    dft4r(xr[1], xi[1], xr[3], xi[3], xr[5], xi[5], xr[7], xi[7]);   // O[k] at 2k + 1
    const T o0r = xr[1], o0i = xi[1];
    const T o1r = h * (xr[3] + xi[3]), o1i = h * (xi[3] - xr[3]);     // W8^1 = h (1 - i)
    const T o2r = xi[5], o2i = -xr[5];                                // W8^2 = -i
    const T o3r = h * (xi[7] - xr[7]), o3i = -h * (xr[7] + xi[7]);    // W8^3 = -h (1 + i)
    const T e0r = xr[0], e0i = xi[0], e1r = xr[2], e1i = xi[2], e2r = xr[4], e2i = xi[4], e3r = xr[6], e3i = xi[6];
    xr[0] = e0r + o0r; xi[0] = e0i + o0i; xr[4] = e0r - o0r; xi[4] = e0i - o0i;
    xr[1] = e1r + o1r; xi[1] = e1i + o1i; xr[5] = e1r - o1r; xi[5] = e1i - o1i;
    xr[2] = e2r + o2r; xi[2] = e2i + o2i; xr[6] = e2r - o2r; xi[6] = e2i - o2i;
    xr[3] = e3r + o3r; xi[3] = e3i + o3i; xr[7] = e3r - o3r; xi[7] = e3i - o3i;
}

constexpr int kFfFfts = kSfCols;             // one 128-point complex FFT per real column
constexpr int kFfThreads = kFfFfts * 8;      // 224
constexpr int kFfZRow = 9;                   // double2 row stride of the 16 x 8 exchange
constexpr int kFfZFft = 16 * kFfZRow;        // double2 per FFT

constexpr int kFfR1Pitch = 24;
constexpr int kFfXbFrame = kSfPadLen + 8;     // bf16 frame stride: 340 words = 20 mod 32
constexpr int kFfXbCopy = kSfFrames * kFfXbFrame;   // second copy starts 1360 words = 16 mod 32 further: see the delta term

struct SileroFftSmem {
    union {
        float xs[kSfFrames * kSfPadLen];  // reflect-padded frames: dead once the FFTs are done ...
        struct {                          // ... so the small late-stage buffers live in the same bytes
            float y2[16 * kSfFrames * 4];
            float r2[16 * kSfFrames * 4];
            float z1[32 * kSfFrames * 4];
            float z2[32 * kSfFrames * 2];
            float r3[32 * kSfFrames * 2];
            float u1[32 * kSfFrames * 2];
            float u2[32 * kSfFrames];
            float r4[32 * kSfFrames];
            float v1[64 * kSfFrames];
        };
    };
    union {
        __nv_bfloat16 xb[2 * kFfXbCopy];       // two bf16 copies of the frames (B operand of the delta term), then ...
        double2 z[kFfFfts * kFfZFft];          // ... the FFT exchange, then ...
        struct {
            float r1[264 * kFfR1Pitch];        // ... the depthwise output of block 1 at the even STFT columns (the only ones
                                               //     the stride-2 convolution behind it reads), row pitch 24: conflict-free B loads
            float part1[7 * 256];              //     and the seven partial tiles of its pointwise product
        };
    };
    float x1[264 * kSfCols];              // [258][frames][7]: re | im, then magnitude | log spectrum, then magnitude | norm;
                                          // rows 258..263 stay zero (K padding of the block-1 product)
    float y1[16 * kSfFrames * 4];
    float part[8 * kSfCols];
    float mm[kSfFrames];
};
static_assert(sizeof(SileroFftSmem) * 2 <= 227 * 1024 - 2 * 1024, "two CTAs per SM");
static_assert(kSfPadLen == 3 * kFfThreads, "the frame load assumes three samples per thread");

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kFfThreads, 2) k_silero_features_fft(const float* __restrict__ pcm, int64_t pcm_stride, int n_frames,
                                                                    float* __restrict__ y2g, SileroDev wts) {
    extern __shared__ __align__(16) unsigned char smem_raw_f[];
    SileroFftSmem& s = *reinterpret_cast<SileroFftSmem*>(smem_raw_f);
    const int tid = threadIdx.x;
    const int stream = blockIdx.y;
    const int f0 = blockIdx.x * kSfFrames;
    const float* src = pcm + (int64_t)stream * pcm_stride;
    if (tid < 6 * kSfCols) s.x1[258 * kSfCols + tid] = 0.0f;          // K padding rows of the block-1 product
    // reflect pad 96 on both sides of every 480-sample frame; two bf16 copies for the delta term
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int j = tid + r * kFfThreads;
        int k = j - 96;
        if (k < 0) k = -k;
        if (k >= 480) k = 2 * 479 - k;
#pragma unroll
        for (int f = 0; f < kSfFrames; ++f) {
            const float v = (f0 + f < n_frames) ? __ldg(src + (int64_t)(f0 + f) * 480 + k) : 0.0f;
            s.xs[f * kSfPadLen + j] = v;
            const __nv_bfloat16 vb = __float2bfloat16_rn(v);
            s.xb[f * kFfXbFrame + j] = vb;
            s.xb[kFfXbCopy + f * kFfXbFrame + j] = vb;
        }
    }
    __syncthreads();
    {
        // delta term on the tensor cores: C[256 rows][28 columns] = (2^24 delta)[256][256] . x[256][28].  Warp w takes the
        // row tiles w, w + 7, w + 14; per k-step one coalesced 16-byte load of the pre-arranged A fragment and, per column
        // tile, the B fragment straight from the bf16 frames.  MMA column (tile nt, g) = STFT column t = 2 nt + (g >> 2) of
        // frame g & 3, read from copy g >> 2: the eight windows of one B load then start in eight different 4-bank groups
        // (frame stride 20, copy offset 16 banks), where consecutive windows of ONE frame would all share one (64 samples
        // = 32 words apart).  Tile 3 has t = 6 only; its upper half repeats it and is dropped.
        const int wq = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
        // ldmatrix.x4 = the B fragments of two column tiles at one k-step: lanes 8 i .. 8 i + 7 address the eight rows
        // (= windows) of matrix i = (tile pair's tile i >> 1, k half i & 1)
        uint32_t brow[2];
        {
            const int i = lane >> 3, r = lane & 7;
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                const int t = min(2 * (2 * pr + (i >> 1)) + (r >> 2), 6);
                brow[pr] = (uint32_t)__cvta_generic_to_shared(s.xb + (r >> 2) * kFfXbCopy + (r & 3) * kFfXbFrame + 64 * t + 8 * (i & 1));
            }
        }
        for (int mt = wq; mt < 16; mt += 7) {
            float acc[4][4];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
            const uint4* ap = wts.delta_frag + (size_t)mt * 16 * 32 + lane;
            uint4 af[16];                                // all 16 fragments of the row tile in flight together (they come from L2)
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) af[ks] = __ldg(ap + ks * 32);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const uint4 a = af[ks];
#pragma unroll
                for (int pr = 0; pr < 2; ++pr) {
                    uint32_t b[4];
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(brow[pr] + ks * 32));
                    mma_bf16_16816(acc[2 * pr], a, b[0], b[1]);
                    mma_bf16_16816(acc[2 * pr + 1], a, b[2], b[3]);
                }
            }
            const int r0 = mt * 16 + g, r1 = r0 + 8;
            const int R0 = r0 < 129 ? r0 : r0 + 1, R1 = r1 < 129 ? r1 : r1 + 1;   // sine rows 1..127 live at 130..256
            constexpr float kScale = 1.0f / 16777216.0f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gc = 2 * t4 + e, t = 2 * nt + (gc >> 2);           // C fragment: MMA columns 2 t4, 2 t4 + 1
                    if (t < 7) {
                        const int c = (gc & 3) * 7 + t;
                        s.x1[R0 * kSfCols + c] = kScale * acc[nt][e];
                        s.x1[R1 * kSfCols + c] = kScale * acc[nt][2 + e];
                    }
                }
        }
    }
    __syncthreads();
    {
        const int j = tid >> 3, l = tid & 7;             // column j = frame * 7 + t
        const int f = j / 7, t = j - f * 7;
        const float2* xa = reinterpret_cast<const float2*>(s.xs + f * kSfPadLen + 64 * t);
        const float2* wn = reinterpret_cast<const float2*>(wts.win);
        double xr[16], xi[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {                // z[8 n1 + l] = (x[2n], x[2n + 1]) win  (exact in f64)
            const float2 v = xa[8 * n1 + l], w = __ldg(wn + 8 * n1 + l);
            xr[n1] = (double)v.x * (double)w.x;
            xi[n1] = (double)v.y * (double)w.y;
        }
        dft16(xr, xi);                                   // Y[k1] of column n2 = l
        double2* zj = s.z + j * kFfZFft;
        zj[l] = make_double2(xr[0], xi[0]);
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
            const double2 w = __ldg(wts.tw + k1 * 8 + l);   // W128^{l k1}
            zj[k1 * kFfZRow + l] = make_double2(xr[k1] * w.x - xi[k1] * w.y, xr[k1] * w.y + xi[k1] * w.x);
        }
        __syncwarp();
        {
            // rows k1 = l and l + 8: Z[k1 + 16 k2] = Z[l + 8 m], m = 2 k2 (+ 1 for the second row)
            double ar[8], ai[8];
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) { const double2 v = zj[l * kFfZRow + n2]; ar[n2] = v.x; ai[n2] = v.y; }
            dft8(ar, ai);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { xr[2 * k2] = ar[k2]; xi[2 * k2] = ai[k2]; }
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) { const double2 v = zj[(l + 8) * kFfZRow + n2]; ar[n2] = v.x; ai[n2] = v.y; }
            dft8(ar, ai);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) { xr[2 * k2 + 1] = ar[k2]; xi[2 * k2 + 1] = ai[k2]; }
        }
        // bin k = l + 8 m needs Z[128 - k]: lane (8 - l) & 7 holds it at index 15 - m (l > 0), lane 0 its own at (16 - m) & 15
        const int srcl = (tid & 24) | ((8 - l) & 7);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            double pr = __shfl_sync(0xffffffffu, xr[15 - m], srcl);
            double pi = __shfl_sync(0xffffffffu, xi[15 - m], srcl);
            if (l == 0) { pr = xr[(16 - m) & 15]; pi = xi[(16 - m) & 15]; }
            const int k = l + 8 * m;
            const double er = 0.5 * (xr[m] + pr), ei = 0.5 * (xi[m] - pi);
            const double qr = 0.5 * (xi[m] + pi), qi = 0.5 * (pr - xr[m]);       // O[k]
            const double2 w = __ldg(wts.tw2 + k);                                 // e^{-2 pi i k / 256}
            // + the delta term already in x1, then magnitude | log(1 + 2^20 magnitude) in place (the sine rows of bins 0 and
            // 128 have no delta row: their slots hold nothing yet)
            const float re = (float)(er + (qr * w.x - qi * w.y)) + s.x1[k * kSfCols + j];
            const float im = k == 0 ? 0.0f : (float)(ei + (qr * w.y + qi * w.x)) + s.x1[(129 + k) * kSfCols + j];
            const float mag = sqrtf(re * re + im * im);
            s.x1[k * kSfCols + j] = mag;
            s.x1[(129 + k) * kSfCols + j] = __logf(fmaf(1048576.0f, mag, 1.0f));   // abs error < 5e-6 on values up to 18: see the parity test
        }
        if (l == 0) {                                    // X[128] = E[0] - O[0] = Re Z[0] - Im Z[0], real
            const float mag = fabsf((float)(xr[0] - xi[0]) + s.x1[128 * kSfCols + j]);
            s.x1[128 * kSfCols + j] = mag;
            s.x1[257 * kSfCols + j] = __logf(fmaf(1048576.0f, mag, 1.0f));
        }
    }
    __syncthreads();
    // adaptive normalisation: mean over the 129 bins, reflect pad 3, 7-tap filter, mean over T
    {
        const int p = tid / kSfCols, col = tid - p * kSfCols;     // 8 partial sums per column
        float a = 0.f;
        for (int i = p; i < 129; i += 8) a += s.x1[(129 + i) * kSfCols + col];
        s.part[p * kSfCols + col] = a;
    }
    __syncthreads();
    if (tid < kSfFrames) {
        float m[7];
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            float a = 0.f;
#pragma unroll
            for (int p = 0; p < 8; ++p) a += s.part[p * kSfCols + tid * 7 + t];
            m[t] = a / 129.0f;
        }
        float p[13];
        p[0] = m[3]; p[1] = m[2]; p[2] = m[1];
#pragma unroll
        for (int t = 0; t < 7; ++t) p[3 + t] = m[t];
        p[10] = m[5]; p[11] = m[4]; p[12] = m[3];
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 7; ++k) a = fmaf(__ldg(wts.norm_filter + k), p[t + k], a);
            acc += a;
        }
        s.mm[tid] = acc / 7.0f;
    }
    __syncthreads();
    if (tid < 6 * kFfR1Pitch) s.r1[258 * kFfR1Pitch + tid] = 0.0f;    // K padding rows (the exchange buffer is dead by now)
    // block 1 (258 -> 16, T 7 -> 4).  Depthwise k5: one (channel, frame) row of 7 per item; the normalisation
    // (norm = log spectrum - mm[frame]) is applied on the way through and written back for the projection.
#pragma unroll
    for (int pass = 0; pass < (258 * kSfFrames + kFfThreads - 1) / kFfThreads; ++pass) {   // fixed trip count: the weight loads of all passes overlap
        const int item = tid + pass * kFfThreads;
        if (item >= 258 * kSfFrames) break;
        const int c = item >> 2, f = item & 3;
        float* row = s.x1 + item * 7;
        float v[11];
        v[0] = v[1] = v[9] = v[10] = 0.f;
        const float sub = c >= 129 ? s.mm[f] : 0.f;
#pragma unroll
        for (int t = 0; t < 7; ++t) { v[2 + t] = row[t] - sub; }
        if (c >= 129) {
#pragma unroll
            for (int t = 0; t < 7; ++t) row[t] = v[2 + t];
        }
        float w[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = __ldg(wts.b1_dw_w + c * 5 + k);
        const float b = __ldg(wts.b1_dw_b + c);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                    // t = 2 j: the stride-2 k1 convolution after the pointwise one reads no other
            float a = b;
#pragma unroll
            for (int k = 0; k < 5; ++k) a = fmaf(w[k], v[2 * j + k], a);
            o[j] = fmaxf(a, 0.f);
        }
        *reinterpret_cast<float4*>(s.r1 + c * kFfR1Pitch + f * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();
    {
        // pointwise 2 x (258 -> 16) at the 16 even STFT columns (4 frames x t = 0, 2, 4, 6; the stride-2 convolution that
        // follows reads no others) = one [16 x 528] x [528 x 16] product on the tensor cores: K = the 258 depthwise outputs
        // (padded to 264) | the 258 block inputs (padded to 264), 66 k-steps of mma.m16n8k8 TF32 with the 3-pass split
        // (weights pre-split and pre-arranged as A fragments on the host, activations split here): f32-level products, f32
        // accumulation.  Warp w takes the k-steps w, w + 7, ...; the seven partial tiles are summed in a fixed order.
        const int wq = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
        // MMA column n = frame * 4 + j: r1 holds it at n, x1 at frame * 7 + 2 j
        int xcol[2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { const int n = nt * 8 + g; xcol[nt] = (n >> 2) * 7 + (n & 3) * 2; }
        // rows 258..263 of both operands are zero (K padding): no predicates in the loops
        // A fragments (hi | lo) of this warp's k-steps, all in flight together (they come from L2): steps wq + 7 i of each half
        uint4 fa[2][5][2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int ks = min(wq + 7 * i, 32) + 33 * hf;
                fa[hf][i][0] = __ldg(wts.b1_frag + (ks * 32 + lane) * 2);
                fa[hf][i][1] = __ldg(wts.b1_frag + (ks * 32 + lane) * 2 + 1);
            }
        auto kstep = [&](const uint4& ahv, const uint4& alv, const float* p0, const float* p1, int rs) {
            const uint32_t ah[4] = {ahv.x, ahv.y, ahv.z, ahv.w}, al[4] = {alv.x, alv.y, alv.z, alv.w};
            const float* pp[2] = {p0, p1};
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                uint32_t h0, l0, h1, l1;
                split_tf32_trunc(pp[nt][0], h0, l0);
                split_tf32_trunc(pp[nt][rs], h1, l1);
                mma_tf32(acc[nt], ah, h0, h1);
                mma_tf32(acc[nt], ah, l0, l1);
                mma_tf32(acc[nt], al, h0, h1);
            }
        };
#pragma unroll
        for (int i = 0; i < 5; ++i) {                                  // depthwise outputs: channel 8 ks + t4 (+ 4)
            const int ks = wq + 7 * i;
            if (ks < 33) {
                const float* p = s.r1 + (ks * 8 + t4) * kFfR1Pitch + g;
                kstep(fa[0][i][0], fa[0][i][1], p, p + 8, 4 * kFfR1Pitch);
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {                                  // block inputs
            const int ks = wq + 7 * i;
            if (ks < 33) {
                const float* p = s.x1 + (ks * 8 + t4) * kSfCols;
                kstep(fa[1][i][0], fa[1][i][1], p + xcol[0], p + xcol[1], 4 * kSfCols);
            }
        }
        float* part = s.part1 + wq * 256;                              // [16 co][16 columns] per warp
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            *reinterpret_cast<float2*>(part + g * 16 + nt * 8 + 2 * t4) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2*>(part + (g + 8) * 16 + nt * 8 + 2 * t4) = make_float2(acc[nt][2], acc[nt][3]);
        }
    }
    __syncthreads();
    for (int o = tid; o < 16 * 16; o += kFfThreads) {
        const int co = o >> 4;
        float a = __ldg(wts.b1_pw_b + co) + __ldg(wts.b1_proj_b + co);
#pragma unroll
        for (int w = 0; w < 7; ++w) a += s.part1[w * 256 + o];
        s.y1[o] = fmaxf(a, 0.f);
    }
    __syncthreads();
    constexpr int F = kSfFrames;
    mix_relu<16, 16, F * 4, 2, false, false>(s.y1, wts.b1_down_t, wts.b1_down_b, nullptr, nullptr, nullptr, nullptr, s.y2, F * 4,
                                             [](int p) { return p; });                                       // T 7 -> 4, already decimated
    __syncthreads();
    // block-1 output [frame][16][4] for k_silero_blocks
    for (int i = tid; i < F * 64; i += kFfThreads) {
        const int f = i >> 6, ch = (i >> 2) & 15, t = i & 3;
        if (f0 + f < n_frames) y2g[((int64_t)stream * n_frames + f0 + f) * 64 + (i & 63)] = s.y2[ch * (F * 4) + f * 4 + t];
    }
}

// ------------------------------------------------------------------------------------------
// Blocks 2-4 of the Silero front for 32 frames per CTA: the k1 convolutions' weights (48 KB, [ci][co]) are staged in shared
// memory once per CTA, so the nine small stages run at shared-memory latency instead of one L2 round trip each (inside the
// per-4-frame kernel they were 27 % of its time).  in: block-1 output [frame][16][4]; out: [frame][64].
// ------------------------------------------------------------------------------------------
constexpr int kS2Frames = 32;
constexpr int kS2Weights = 2 * 512 + 1024 + 1024 + 1024 + 2 * 2048 + 4096;   // b2_pw_t .. b4_down_t, contiguous in the table

struct Silero2Smem {
    float w[kS2Weights];
    float a[4096];                       // y2 | y2e | r2, then z2 | r3 | z2e, then v1
    float b[4096];                       // z1, then u1 | u2 | r4
};
static_assert(kS2Frames == 32, "the buffer offsets in k_silero_blocks assume 32 frames");
static_assert(sizeof(Silero2Smem) * 2 <= 227 * 1024 - 2 * 1024, "two CTAs per SM");

__global__ void __launch_bounds__(256, 2) k_silero_blocks(const float* __restrict__ y2g, int n_frames, float* __restrict__ out, SileroDev wts) {
    extern __shared__ __align__(16) unsigned char smem_raw_2[];
    Silero2Smem& s = *reinterpret_cast<Silero2Smem*>(smem_raw_2);
    constexpr int F = kS2Frames;
    const int tid = threadIdx.x;
    const int stream = blockIdx.y;
    const int f0 = blockIdx.x * F;
    {
        const float4* src = reinterpret_cast<const float4*>(wts.b2_pw_t);
        float4* dst = reinterpret_cast<float4*>(s.w);
        for (int i = tid; i < kS2Weights / 4; i += 256) dst[i] = __ldg(src + i);
        // [frame][16][4] -> [16][frame * 4 + t]
        const float4* yin = reinterpret_cast<const float4*>(y2g + ((int64_t)stream * n_frames + f0) * 64);
        for (int i = tid; i < F * 16; i += 256) {
            const int f = i >> 4, ch = i & 15;
            const float4 v = f0 + f < n_frames ? __ldg(yin + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(s.a + ch * (F * 4) + f * 4) = v;
        }
    }
    __syncthreads();
    const float* w_b2_pw = s.w, *w_b2_proj = s.w + 512, *w_b2_down = s.w + 1024, *w_b3_pw = s.w + 2048, *w_b3_down = s.w + 3072;
    const float* w_b4_pw = s.w + 4096, *w_b4_proj = s.w + 6144, *w_b4_down = s.w + 8192;
    // Every block ends in a k1 convolution of stride 2, which reads the even positions only: the depthwise and pointwise
    // convolutions in front of it are evaluated at those positions and stored compactly ([C][frame][T / 2]).
    float* y2 = s.a;                   // [16][F][4]
    float* y2e = s.a + 2048;           // [16][F][2]  (t = 0, 2)
    float* r2 = s.a + 3072;            // [16][F][2]
    float* z1 = s.b;                   // [32][F][2]
    float* z2 = s.a;                   // [32][F][2]
    float* r3 = s.a + 2048;            // [32][F]     (t = 0)
    float* z2e = s.a + 3072;           // [32][F]
    float* u1 = s.b;                   // [32][F]
    float* u2 = s.b + 1024;            // [32][F]
    float* r4 = s.b + 2048;            // [32][F]
    float* v1 = s.a;                   // [64][F]
    // block 2 (16 -> 32, T 4 -> 2): depthwise k5 (zero pad inside the frame) at t = 0, 2
    for (int i = tid; i < 16 * F; i += 256) {
        const int c = i / F;
        const float4 v = *reinterpret_cast<const float4*>(y2 + i * 4);
        float w[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = __ldg(wts.b2_dw_w + c * 5 + k);
        const float b = __ldg(wts.b2_dw_b + c);
        const float o0 = fmaf(w[4], v.z, fmaf(w[3], v.y, fmaf(w[2], v.x, b)));                       // t = 0: taps 2..4
        const float o2 = fmaf(w[3], v.w, fmaf(w[2], v.z, fmaf(w[1], v.y, fmaf(w[0], v.x, b))));      // t = 2: taps 0..3
        *reinterpret_cast<float2*>(r2 + i * 2) = make_float2(fmaxf(o0, 0.f), fmaxf(o2, 0.f));
        *reinterpret_cast<float2*>(y2e + i * 2) = make_float2(v.x, v.z);
    }
    __syncthreads();
    mix4_relu<16, 32, F * 2, 8, 2, true, false>(r2, w_b2_pw, wts.b2_pw_b, y2e, w_b2_proj, wts.b2_proj_b, nullptr, z1);
    __syncthreads();
    mix4_relu<32, 32, F * 2, 8, 2, false, false>(z1, w_b2_down, wts.b2_down_b, nullptr, nullptr, nullptr, nullptr, z2);
    __syncthreads();
    // block 3 (32 -> 32 with identity residual, T 2 -> 1): depthwise at t = 0
    for (int i = tid; i < 32 * F; i += 256) {
        const int c = i / F;
        const float2 v = *reinterpret_cast<const float2*>(z2 + i * 2);
        r3[i] = fmaxf(fmaf(__ldg(wts.b3_dw_w + c * 5 + 3), v.y, fmaf(__ldg(wts.b3_dw_w + c * 5 + 2), v.x, __ldg(wts.b3_dw_b + c))), 0.f);
        z2e[i] = v.x;
    }
    __syncthreads();
    mix4_relu<32, 32, F, 4, 2, false, true>(r3, w_b3_pw, wts.b3_pw_b, nullptr, nullptr, nullptr, z2e, u1);
    __syncthreads();
    mix4_relu<32, 32, F, 4, 2, false, false>(u1, w_b3_down, wts.b3_down_b, nullptr, nullptr, nullptr, nullptr, u2);
    __syncthreads();
    // block 4 (32 -> 64, T 1): the depthwise convolution sees its centre tap only
    for (int i = tid; i < 32 * F; i += 256) {
        const int c = i / F;
        r4[i] = fmaxf(fmaf(__ldg(wts.b4_dw_w + c * 5 + 2), u2[i], __ldg(wts.b4_dw_b + c)), 0.f);
    }
    __syncthreads();
    mix4_relu<32, 64, F, 8, 2, true, false>(r4, w_b4_pw, wts.b4_pw_b, u2, w_b4_proj, wts.b4_proj_b, nullptr, v1);
    __syncthreads();
    // final 64 -> 64, ReLU, write [frame][64]: two output channels x eight consecutive frames per thread (128 threads)
    if (tid < 128) {
        const int co = tid & 31, f8 = (tid >> 5) * 8;
        float acc[2][8];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float bias = __ldg(wts.b4_down_b + co + 32 * c);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[c][e] = bias;
        }
#pragma unroll 8
        for (int ci = 0; ci < 64; ++ci) {
            const float w0 = w_b4_down[ci * 64 + co], w1 = w_b4_down[ci * 64 + co + 32];
            const float4 va = *reinterpret_cast<const float4*>(v1 + ci * F + f8), vb = *reinterpret_cast<const float4*>(v1 + ci * F + f8 + 4);
            const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc[0][e] = fmaf(w0, v[e], acc[0][e]); acc[1][e] = fmaf(w1, v[e], acc[1][e]); }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int f = f0 + f8 + e;
            if (f < n_frames) {
                out[((int64_t)stream * n_frames + f) * 64 + co] = fmaxf(acc[0][e], 0.f);
                out[((int64_t)stream * n_frames + f) * 64 + co + 32] = fmaxf(acc[1][e], 0.f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// LSTM layer (ONNX gate order i, o, f, c) on the f16 tensor cores.  Per time step the 256 gate pre-activations of the
// CTA's 8 streams are one [256 x 128] x [128 x 8] product: A = [W | R] (constant: each warp keeps its two 16-row tiles
// as m16n8k16 fragments in registers for the whole sequence, split hi = f16(256 w) | lo = f16(256 w - hi)), B = [x_t ;
// h_{t-1}] of the 8 streams, kept in shared memory as f16 hi | lo (22 bits) and read with ldmatrix; three MMAs per
// product (hi hi + hi lo + lo hi), f32 accumulation.  The CUDA-core form spent 1024 FMAs per thread and step on this.
//   xin  [n_streams][n_frames][64]   hout [n_streams][n_frames][64] (layer 1) or probs (layer 2)
//   h, c [n_streams][64] in/out state of this layer
// ------------------------------------------------------------------------------------------
constexpr int kLsStreams = 8;
constexpr int kLsPitch = 136;            // halves per stream row of [x ; h]: 17 16-byte chunks, conflict-free for ldmatrix
constexpr int kLsGPitch = 260;           // floats per stream row of the gate buffer (4 mod 32: conflict-free fragment stores)
constexpr float kLsWScale = 256.0f;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

template <bool LAST>
__global__ void __launch_bounds__(256, 1) k_silero_lstm(const float* __restrict__ xin, int n_streams, int n_frames,
                                                        const float* __restrict__ W, const float* __restrict__ R,
                                                        const float* __restrict__ B, float* __restrict__ h_state,
                                                        float* __restrict__ c_state, float* __restrict__ hout,
                                                        const float* __restrict__ dec_w, const float* __restrict__ dec_b,
                                                        float* __restrict__ probs) {
    __shared__ __align__(16) __half xh[2][kLsStreams][kLsPitch];   // [hi | lo][stream][x_t (64) ; h_{t-1} (64)]
    __shared__ float gates[kLsStreams][kLsGPitch];
    __shared__ float cst[kLsStreams][64];
    __shared__ float hf[kLsStreams][64];                           // h_t in f32 (state out, decoder)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int s0 = blockIdx.x * kLsStreams;
    // A fragments of the warp's row tiles 2 warp, 2 warp + 1: a0a1 (row g, k 2t..), a2a3 (row g + 8), a4a5 (row g, k 2t + 8..),
    // a6a7 (row g + 8, k 2t + 8..)
    uint4 ah[2][8], al[2][8];
    float bias[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int row0 = 32 * warp + 16 * mt + g;
        bias[mt][0] = __ldg(B + row0) + __ldg(B + 256 + row0);
        bias[mt][1] = __ldg(B + row0 + 8) + __ldg(B + 256 + row0 + 8);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t h4[4], l4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = row0 + 8 * (e & 1), k = 16 * ks + 2 * t4 + 8 * (e >> 1);
                const float* src = k < 64 ? W + row * 64 + k : R + row * 64 + (k - 64);
                __half h0, l0, h1, l1;
                split_f16(__ldg(src) * kLsWScale, h0, l0);
                split_f16(__ldg(src + 1) * kLsWScale, h1, l1);
                h4[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                l4[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            ah[mt][ks] = make_uint4(h4[0], h4[1], h4[2], h4[3]);
            al[mt][ks] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
        }
    }
    // element pairs (stream, 2 j): 256 threads x one pair of the 8 x 64 tile
    const int es = tid >> 5, ej = (tid & 31) * 2;
    const bool s_ok = s0 + es < n_streams;
    auto put = [&](int col, float v0, float v1) {
        const __half2 hi = __floats2half2_rn(v0, v1);
        const float2 f = __half22float2(hi);
        *reinterpret_cast<__half2*>(&xh[0][es][col]) = hi;
        *reinterpret_cast<__half2*>(&xh[1][es][col]) = __floats2half2_rn(v0 - f.x, v1 - f.y);
    };
    {
        float2 h0 = make_float2(0.f, 0.f), c0 = h0;
        if (s_ok) {
            h0 = *reinterpret_cast<const float2*>(h_state + (int64_t)(s0 + es) * 64 + ej);
            c0 = *reinterpret_cast<const float2*>(c_state + (int64_t)(s0 + es) * 64 + ej);
        }
        put(64 + ej, h0.x, h0.y);
        hf[es][ej] = h0.x; hf[es][ej + 1] = h0.y;
        cst[es][ej] = c0.x; cst[es][ej + 1] = c0.y;
    }
    const float* xrow = xin + (int64_t)(s0 + es) * n_frames * 64 + ej;
    float2 xn = s_ok ? __ldg(reinterpret_cast<const float2*>(xrow)) : make_float2(0.f, 0.f);
    put(ej, xn.x, xn.y);
    // ldmatrix row of this lane: matrix lane >> 3 = (hi | lo, k half), row lane & 7 = stream
    const uint32_t brow = (uint32_t)__cvta_generic_to_shared(&xh[lane >> 4][lane & 7][8 * ((lane >> 3) & 1)]);
    const float db = LAST ? __ldg(dec_b) : 0.f;
    const float dw0 = LAST ? __ldg(dec_w + ej) : 0.f, dw1 = LAST ? __ldg(dec_w + ej + 1) : 0.f;
    __syncthreads();
    for (int t = 0; t < n_frames; ++t) {
        // next step's input: in flight during this step's product
        if (t + 1 < n_frames && s_ok) xn = __ldg(reinterpret_cast<const float2*>(xrow + (int64_t)(t + 1) * 64));
        float acc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            uint32_t b[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(brow + ks * 32));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                mma_f16_16816(acc[mt], ah[mt][ks], b[0], b[1]);
                mma_f16_16816(acc[mt], ah[mt][ks], b[2], b[3]);
                mma_f16_16816(acc[mt], al[mt][ks], b[0], b[1]);
            }
        }
        // C fragment: c0 c1 (row g, streams 2 t4, 2 t4 + 1), c2 c3 (row g + 8)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const int row = 32 * warp + 16 * mt + g;
            gates[2 * t4][row] = fmaf(acc[mt][0], 1.0f / kLsWScale, bias[mt][0]);
            gates[2 * t4 + 1][row] = fmaf(acc[mt][1], 1.0f / kLsWScale, bias[mt][0]);
            gates[2 * t4][row + 8] = fmaf(acc[mt][2], 1.0f / kLsWScale, bias[mt][1]);
            gates[2 * t4 + 1][row + 8] = fmaf(acc[mt][3], 1.0f / kLsWScale, bias[mt][1]);
        }
        __syncthreads();
        {
            float hv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = ej + e;
                const float ig = sigmoidf_(gates[es][j]), og = sigmoidf_(gates[es][64 + j]);
                const float fg = sigmoidf_(gates[es][128 + j]), cg = tanhf(gates[es][192 + j]);
                const float c = fg * cst[es][j] + ig * cg;
                cst[es][j] = c;
                hv[e] = og * tanhf(c);
            }
            put(64 + ej, hv[0], hv[1]);
            put(ej, xn.x, xn.y);
            if (!LAST) {
                if (s_ok) *reinterpret_cast<float2*>(hout + ((int64_t)(s0 + es) * n_frames + t) * 64 + ej) = make_float2(hv[0], hv[1]);
                if (t + 1 == n_frames) { hf[es][ej] = hv[0]; hf[es][ej + 1] = hv[1]; }
            } else {
                hf[es][ej] = hv[0]; hf[es][ej + 1] = hv[1];
                // decoder: warp = stream (es == warp), lane = the two units it just produced
                float a = fmaf(dw0, fmaxf(hv[0], 0.f), dw1 * fmaxf(hv[1], 0.f));
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                if (lane == 0 && s_ok) probs[(int64_t)(s0 + es) * n_frames + t] = sigmoidf_(a + db);
            }
        }
        __syncthreads();
    }
    if (s_ok) {
        *reinterpret_cast<float2*>(h_state + (int64_t)(s0 + es) * 64 + ej) = make_float2(hf[es][ej], hf[es][ej + 1]);
        *reinterpret_cast<float2*>(c_state + (int64_t)(s0 + es) * 64 + ej) = make_float2(cst[es][ej], cst[es][ej + 1]);
    }
}

// ------------------------------------------------------------------------------------------
// SmoothedVad FSM (one thread per stream) and compaction.
//   plan[stream][t] = {first source frame, n frames emitted, output frame offset}
// ------------------------------------------------------------------------------------------
struct GatePlan { int src, n, off; };

__global__ void k_vad_plan(const float* __restrict__ probs, int n_streams, int n_frames, float threshold, int prefill,
                           int hangover, int onset, GatePlan* __restrict__ plan, int* __restrict__ out_frames) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    bool in_speech = false;
    int onset_c = 0, hang_c = 0, off = 0;
    for (int t = 0; t < n_frames; ++t) {
        const bool v = probs[(int64_t)s * n_frames + t] > threshold;
        GatePlan p{t, 0, off};
        if (!in_speech && v) {
            if (++onset_c >= onset) {
                in_speech = true; hang_c = hangover; onset_c = 0;
                const int buffered = min(t + 1, prefill + 1);
                p.src = t + 1 - buffered; p.n = buffered;
            }
        } else if (in_speech && v) {
            hang_c = hangover; p.n = 1;
        } else if (in_speech && !v) {
            if (hang_c > 0) { --hang_c; p.n = 1; }
            else in_speech = false;
        } else {
            onset_c = 0;
        }
        plan[(int64_t)s * n_frames + t] = p;
        off += p.n;
    }
    out_frames[s] = off;
}

__global__ void __launch_bounds__(128) k_vad_compact(const float* __restrict__ pcm, int64_t pcm_stride, int n_frames,
                                                     const GatePlan* __restrict__ plan, float* __restrict__ out,
                                                     int64_t out_stride, int max_out_frames) {
    const int s = blockIdx.y, t = blockIdx.x;
    const GatePlan p = plan[(int64_t)s * n_frames + t];
    if (p.n == 0) return;
    const float* src = pcm + (int64_t)s * pcm_stride + (int64_t)p.src * 480;
    float* dst = out + (int64_t)s * out_stride + (int64_t)p.off * 480;
    const int n = min(p.n, max_out_frames - p.off) * 480;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
}

}  // namespace sb

// ------------------------------------------------------------------------------------------
// host side / C ABI
// ------------------------------------------------------------------------------------------
struct sb_resampler {
    int fs_in = 0, fs_out = 0, decim = 1, n_taps = 0, fft_in = 0, fft_out = 0;
    uint4* d_afrag = nullptr;      // polyphase A fragments, f16 hi | lo: [D][SP + 1][32 lanes][2]
    int Q = 0, SP = 0;             // taps per branch, k-steps per branch
    // rational ratios: the block operator as a dense GEMM (k_resample_dense_*)
    bool dense = false;
    int N1 = 0, N2 = 0, N2p = 0, Kp = 0;   // N2p: N2 rounded up to a multiple of 8 (GEMM N)
    __half *d_whi = nullptr, *d_wlo = nullptr;
    mutable void* d_ws = nullptr;  // X hi | X lo | C of one chunk of streams, grown on demand
    mutable size_t ws_bytes = 0;
    mutable std::mutex ws_mu;      // the workspace is shared by the calls on this object: they enqueue one at a time
};

struct sb_vad {
    float* d_blob = nullptr;
    float* d_basis_t = nullptr;
    void* d_aux = nullptr;         // FFT-form tables, see sb_vad_create
    bool basis_is_dft = false;     // the STFT basis is a windowed 256-point DFT: k_silero_features_fft applies
    sb::SileroDev dev{};
};

namespace {
int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
}

extern "C" {

int sb_resampler_create(int fs_in, int fs_out, sb_resampler** out) {
    SB_CHECK_ARG(out && fs_in > 0 && fs_out > 0, "bad arguments");
    const int g = gcd_i(fs_in, fs_out);
    const int a = fs_in / g, b = fs_out / g;
    const bool integer_ratio = b == 1 && (a == 1 || a == 2 || a == 3 || a == 4 || a == 6);
    sb_resampler* r = new sb_resampler();
    r->fs_in = fs_in; r->fs_out = fs_out; r->decim = integer_ratio ? a : 0; r->dense = !integer_ratio;
    const int fft_chunks = (1024 + a - 1) / a;               // rubato: ceil(chunk_size_in / (fs_in / gcd))
    r->fft_in = fft_chunks * a; r->fft_out = fft_chunks * b; r->n_taps = r->fft_in;
    if (r->dense) {
        // any other ratio (44.1 / 22.05 / 11.025 / 8 kHz ...): the block operator as a dense matrix, see k_resample_dense_matrix
        const int N1 = r->fft_in, N2 = r->fft_out, L = N1 < N2 ? N1 + 1 : N2;
        if ((size_t)N2 * 2 * N1 > (size_t)64 << 20) {
            delete r;
            sb::set_error("resampler: this ratio gives a block operator of more than 64 M entries (fft_size_in x fft_size_out); not taken");
            return SB_ERR_UNSUPPORTED;
        }
        const float cutoff_f = N1 > N2 ? powf(0.4f, 16.0f / (float)N2) * (float)N2 / (float)N1 : powf(0.4f, 16.0f / (float)N1);
        const double cutoff = (double)cutoff_f;
        std::vector<double> h(N1);
        double sum = 0.0;
        for (int x = 0; x < N1; ++x) {
            const double ph = (double)x / N1;
            double w = 0.35875 - 0.48829 * cos(2 * M_PI * ph) + 0.14128 * cos(4 * M_PI * ph) - 0.01168 * cos(6 * M_PI * ph);
            w *= w;
            const double t = ((double)x - (double)(N1 / 2)) * cutoff;
            h[x] = w * (t == 0.0 ? 1.0 : sin(M_PI * t) / (M_PI * t));
            sum += h[x];
        }
        // filter spectrum: rfft([h / sum / (2 N1), zeros(N1)]), bins 0 .. L-1 (exact phase reduction in integers)
        std::vector<double> filt(2 * (size_t)L);
        for (int k = 0; k < L; ++k) {
            double re = 0.0, im = 0.0;
            for (int x = 0; x < N1; ++x) {
                const double th = -2.0 * M_PI * (double)(((long long)k * x) % (2 * N1)) / (double)(2 * N1);
                re += h[x] * cos(th); im += h[x] * sin(th);
            }
            filt[2 * k] = re / sum / (2.0 * N1); filt[2 * k + 1] = im / sum / (2.0 * N1);
        }
        r->N1 = N1; r->N2 = N2; r->N2p = (N2 + 7) & ~7; r->Kp = (2 * N1 + 31) & ~31;
        double2* d_filt = nullptr;
        SB_CUDA_CHECK(cudaMalloc(&d_filt, filt.size() * sizeof(double)));
        SB_CUDA_CHECK(cudaMemcpy(d_filt, filt.data(), filt.size() * sizeof(double), cudaMemcpyHostToDevice));
        SB_CUDA_CHECK(cudaMalloc(&r->d_whi, (size_t)r->N2p * r->Kp * sizeof(__half)));
        SB_CUDA_CHECK(cudaMalloc(&r->d_wlo, (size_t)r->N2p * r->Kp * sizeof(__half)));
        dim3 grid((r->Kp + 127) / 128, r->N2p);
        sb::k_resample_dense_matrix<<<grid, 128>>>(d_filt, N1, N2, L, r->Kp, r->d_whi, r->d_wlo);
        SB_CUDA_CHECK(cudaGetLastError());
        SB_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(d_filt);
        *out = r;
        return SB_OK;
    }
    if (a > 1) {
        // make_sincs(npoints = fft_in, factor 1, cutoff, BlackmanHarris2) -- SURVEY App. B; f32 cutoff like rubato
        const float cutoff_f = powf(0.4f, 16.0f / (float)r->fft_out) * (float)r->fft_out / (float)r->fft_in;
        const double cutoff = (double)cutoff_f;
        std::vector<double> h(r->n_taps);
        double sum = 0.0;
        const int N = r->n_taps;
        for (int x = 0; x < N; ++x) {
            const double ph = (double)x / N;
            double w = 0.35875 - 0.48829 * cos(2 * M_PI * ph) + 0.14128 * cos(4 * M_PI * ph) - 0.01168 * cos(6 * M_PI * ph);
            w *= w;
            const double t = ((double)x - (double)(N / 2)) * cutoff;
            const double sc = t == 0.0 ? 1.0 : sin(M_PI * t) / (M_PI * t);
            h[x] = w * sc; sum += h[x];
        }
        std::vector<float> hf(N);
        for (int x = 0; x < N; ++x) hf[x] = (float)(h[x] / sum);
        // polyphase A fragments: A_p[i][j] = 2^12 h[D (Q - 1 + i - j) + p], m16n8k16 layout (a0a1: row g, k 2t..; a2a3: row g + 8;
        // a4a5: row g, k 2t + 8..; a6a7: row g + 8), split hi = f16(v) | lo = f16(v - hi)
        const int D = a, Q = (N + D - 1) / D, SP = (Q + 15 + 15) / 16;
        r->Q = Q; r->SP = SP;
        auto tap = [&](int p, int i, int j) -> float {
            const int q = Q - 1 + i - j;
            if (q < 0 || q >= Q || D * q + p >= N) return 0.0f;
            return hf[D * q + p] * (float)(1 << sb::kRpScaleLog2);
        };
        std::vector<uint32_t> fr((size_t)D * (SP + 1) * 32 * 8, 0u);      // one all-zero step closes every branch
        for (int p = 0; p < D; ++p)
            for (int sI = 0; sI < SP; ++sI)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3;
                    const int rows[4] = {g, g + 8, g, g + 8}, cols[4] = {2 * t, 2 * t, 2 * t + 8, 2 * t + 8};
                    uint32_t* o = fr.data() + (((size_t)p * (SP + 1) + sI) * 32 + lane) * 8;
                    for (int e = 0; e < 4; ++e) {
                        uint32_t hi2 = 0, lo2 = 0;
                        for (int w = 0; w < 2; ++w) {
                            const float v = tap(p, rows[e], 16 * sI + cols[e] + w);
                            const __half hi = __float2half_rn(v);
                            const __half lo = __float2half_rn(v - __half2float(hi));
                            hi2 |= (uint32_t)__half_as_ushort(hi) << (16 * w);
                            lo2 |= (uint32_t)__half_as_ushort(lo) << (16 * w);
                        }
                        o[e] = hi2; o[4 + e] = lo2;
                    }
                }
        SB_CUDA_CHECK(cudaMalloc(&r->d_afrag, fr.size() * sizeof(uint32_t)));
        SB_CUDA_CHECK(cudaMemcpy(r->d_afrag, fr.data(), fr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    *out = r;
    return SB_OK;
}

int sb_resampler_destroy(sb_resampler* r) {
    if (!r) return SB_OK;
    cudaFree(r->d_afrag); cudaFree(r->d_whi); cudaFree(r->d_wlo); cudaFree(r->d_ws);
    delete r;
    return SB_OK;
}

int sb_resample_geometry(const sb_resampler* r, size_t n_in, size_t* n_fed, size_t* n_out, size_t* n_frames) {
    SB_CHECK_ARG(r, "null resampler");
    const size_t fed = (n_in + 1023) / 1024 * 1024;          // push() chunks + finish() zero pad
    size_t out = fed;
    if (r->decim > 1 || r->dense) out = fed / (size_t)r->fft_in * (size_t)r->fft_out;   // whole rubato blocks only
    else out = n_in;                                          // pass-through (resampler.rs:38-41)
    if (n_fed) *n_fed = fed;
    if (n_out) *n_out = out;
    if (n_frames) *n_frames = (out + 479) / 480;
    return SB_OK;
}

int sb_resample_dev(const sb_resampler* r, const float* in, int64_t in_stride, size_t n_in, int n_streams, float* out,
                    int64_t out_stride, void* stream) {
    SB_CHECK_ARG(r && in && out && n_streams > 0 && n_streams <= 65535, "bad arguments");
    size_t fed, n_out, n_frames;
    sb_resample_geometry(r, n_in, &fed, &n_out, &n_frames);
    SB_CHECK_ARG((size_t)out_stride >= n_frames * 480, "out_stride must hold n_frames * 480 samples");
    cudaStream_t st = (cudaStream_t)stream;
    // the last frame is zero padded by FrameResampler::finish
    // the kernels write [0, n_out); only the tail of the last frame needs zeroing
    if (n_frames * 480 > n_out)
        SB_CUDA_CHECK(cudaMemset2DAsync(out + n_out, out_stride * sizeof(float), 0, (n_frames * 480 - n_out) * sizeof(float), n_streams, st));
    if (r->decim == 1) {
        SB_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride * sizeof(float), in, in_stride * sizeof(float), n_in * sizeof(float),
                                        n_streams, cudaMemcpyDeviceToDevice, st));
        return SB_OK;
    }
    if (n_out == 0) return SB_OK;
    if (r->dense) {
        std::lock_guard<std::mutex> lock(r->ws_mu);
        const int N1 = r->N1, N2 = r->N2, N2p = r->N2p, Kp = r->Kp;
        const int n_blocks = (int)(n_out / (size_t)N2);
        // chunks of streams sized for <= ~1.5 GB of workspace: X hi | X lo (rows x Kp f16 each) | C (rows x N2 f32)
        const size_t row_bytes = (size_t)Kp * 2 * sizeof(__half) + (size_t)N2p * sizeof(float);
        int per_chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_streams, ((size_t)1536 << 20) / (row_bytes * (size_t)n_blocks)));
        const size_t need = row_bytes * (size_t)n_blocks * per_chunk;
        if (need > r->ws_bytes) {
            SB_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(r->d_ws); r->d_ws = nullptr; r->ws_bytes = 0;
            SB_CUDA_CHECK(cudaMalloc(&r->d_ws, need));
            r->ws_bytes = need;
        }
        for (int s0 = 0; s0 < n_streams; s0 += per_chunk) {
            const int ns = std::min(per_chunk, n_streams - s0);
            const int rows = ns * n_blocks;
            __half* xhi = (__half*)r->d_ws;
            __half* xlo = xhi + (size_t)rows * Kp;
            float* c = (float*)(xlo + (size_t)rows * Kp);
            sb::k_resample_dense_rows<<<rows, 256, 0, st>>>(in + (int64_t)s0 * in_stride, in_stride, (int)n_in, n_blocks, N1, Kp, xhi, xlo);
            int rc;
            if ((rc = sb_gemm_tn_dev(SB_DTYPE_F16, xhi, Kp, r->d_whi, Kp, rows, N2p, Kp, c, N2p, 1, nullptr, 0, nullptr, 0, 0, st))) return rc;
            if ((rc = sb_gemm_tn_dev(SB_DTYPE_F16, xhi, Kp, r->d_wlo, Kp, rows, N2p, Kp, c, N2p, 1, nullptr, 0, c, N2p, 0, st))) return rc;
            if ((rc = sb_gemm_tn_dev(SB_DTYPE_F16, xlo, Kp, r->d_whi, Kp, rows, N2p, Kp, c, N2p, 1, nullptr, 0, c, N2p, 0, st))) return rc;
            sb::k_resample_dense_store<<<rows, 256, 0, st>>>(c, n_blocks, N2, N2p, out + (int64_t)s0 * out_stride, out_stride);
            sb::g_launches += 2;
        }
        SB_CUDA_CHECK(cudaGetLastError());
        return SB_OK;
    }
    const int D = r->decim;
    {
        const size_t smem = (size_t)2 * D * sb::rp_array_len(r->SP) * sizeof(__half);
        SB_CHECK_ARG((uint64_t)D * (n_out + sb::kRpTile + 16 * r->SP) < (1ull << 31), "resampler: stream too long for 32-bit sample indices");
        SB_CHECK_ARG(smem <= 200 * 1024, "resampler: filter too long for the shared-memory segment");
        SB_ONCE_PER_DEVICE({
            SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_resample_poly<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_resample_poly<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_resample_poly<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_resample_poly<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        });
        dim3 grid((unsigned)((n_out + sb::kRpTile - 1) / sb::kRpTile), n_streams);
#define SB_RP_LAUNCH(DD) sb::k_resample_poly<DD><<<grid, sb::kRpThreads, smem, st>>>(in, in_stride, (int)n_in, out, out_stride, (int)n_out, r->d_afrag, r->Q, r->SP)
        switch (D) {
            case 2: SB_RP_LAUNCH(2); break;
            case 3: SB_RP_LAUNCH(3); break;
            case 4: SB_RP_LAUNCH(4); break;
            default: SB_RP_LAUNCH(6); break;     // sb_resampler_create admits 2, 3, 4, 6 only
        }
#undef SB_RP_LAUNCH
        sb::g_launches += 1;
        SB_CUDA_CHECK(cudaGetLastError());
        return SB_OK;
    }
}

// blob layout: spittle_b200/silero_weights.py BLOB_LAYOUT
int sb_vad_create(const float* blob, size_t n_floats, sb_vad** out) {
    SB_CHECK_ARG(blob && out, "null pointer");
    static const int sizes[] = {258 * 256, 7, 258 * 5, 258, 16 * 258, 16, 16 * 258, 16, 16 * 16, 16,
                                16 * 5, 16, 32 * 16, 32, 32 * 16, 32, 32 * 32, 32,
                                32 * 5, 32, 32 * 32, 32, 32 * 32, 32,
                                32 * 5, 32, 64 * 32, 64, 64 * 32, 64, 64 * 64, 64,
                                256 * 64, 256 * 64, 512, 256 * 64, 256 * 64, 512, 64, 1};
    size_t total = 0;
    for (int s : sizes) total += s;
    SB_CHECK_ARG(n_floats == total, "Silero weight blob has the wrong size");
    sb_vad* v = new sb_vad();
    SB_CUDA_CHECK(cudaMalloc(&v->d_blob, total * sizeof(float)));
    SB_CUDA_CHECK(cudaMemcpy(v->d_blob, blob, total * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<float> bt(256 * 258);
    for (int c = 0; c < 258; ++c) for (int k = 0; k < 256; ++k) bt[k * 258 + c] = blob[c * 256 + k];
    SB_CUDA_CHECK(cudaMalloc(&v->d_basis_t, bt.size() * sizeof(float)));
    SB_CUDA_CHECK(cudaMemcpy(v->d_basis_t, bt.data(), bt.size() * sizeof(float), cudaMemcpyHostToDevice));
    const float* p = v->d_blob;
    const float* ptr[40];
    for (int i = 0; i < 40; ++i) { ptr[i] = p; p += sizes[i]; }
    sb::SileroDev& d = v->dev;
    d.basis_t = v->d_basis_t; d.norm_filter = ptr[1];
    d.b1_dw_w = ptr[2]; d.b1_dw_b = ptr[3]; d.b1_pw_w = ptr[4]; d.b1_pw_b = ptr[5]; d.b1_proj_w = ptr[6]; d.b1_proj_b = ptr[7];
    d.b1_down_w = ptr[8]; d.b1_down_b = ptr[9];
    d.b2_dw_w = ptr[10]; d.b2_dw_b = ptr[11]; d.b2_pw_w = ptr[12]; d.b2_pw_b = ptr[13]; d.b2_proj_w = ptr[14]; d.b2_proj_b = ptr[15];
    d.b2_down_w = ptr[16]; d.b2_down_b = ptr[17];
    d.b3_dw_w = ptr[18]; d.b3_dw_b = ptr[19]; d.b3_pw_w = ptr[20]; d.b3_pw_b = ptr[21]; d.b3_down_w = ptr[22]; d.b3_down_b = ptr[23];
    d.b4_dw_w = ptr[24]; d.b4_dw_b = ptr[25]; d.b4_pw_w = ptr[26]; d.b4_pw_b = ptr[27]; d.b4_proj_w = ptr[28]; d.b4_proj_b = ptr[29];
    d.b4_down_w = ptr[30]; d.b4_down_b = ptr[31];
    d.lstm_w[0] = ptr[32]; d.lstm_r[0] = ptr[33]; d.lstm_b[0] = ptr[34];
    d.lstm_w[1] = ptr[35]; d.lstm_r[1] = ptr[36]; d.lstm_b[1] = ptr[37];
    d.dec_w = ptr[38]; d.dec_b = ptr[39];
    // is the basis row c = win[n] cos(2 pi c n / 256), row 129 + c = -win[n] sin(...), win = row 0?  (the shipped model:
    // yes, to f32 rounding = 7.7e-8)
    {
        double worst = 0.0;
        for (int c = 0; c < 129; ++c)
            for (int n = 0; n < 256; ++n) {
                const double w = blob[n], th = 2.0 * M_PI * (double)((c * n) & 255) / 256.0;
                worst = std::max(worst, std::fabs((double)blob[c * 256 + n] - w * cos(th)));
                worst = std::max(worst, std::fabs((double)blob[(129 + c) * 256 + n] + w * sin(th)));
            }
        const char* force = getenv("SB_SILERO_DIRECT");
        v->basis_is_dft = worst <= 4e-7 && !(force && force[0] == '1');
    }
    // FFT-form tables (one device buffer): win f32[256] | W128^{n2 k1} double2[16][8] | e^{-2 pi i k / 256} double2[128] |
    // delta fragments uint4[16][16][32] | b1_pw_t f32[258][16] | b1_proj_t f32[258][16] | the nine [ci][co] transposes of the
    // later k1 convolutions
    constexpr size_t kOffTw = 1024, kOffTw2 = kOffTw + 2048, kOffDelta = kOffTw2 + 2048, kOffPw = kOffDelta + 16 * 16 * 32 * 16,
                     kOffProj = kOffPw + 258 * 16 * 4, kOffLate = kOffProj + 258 * 16 * 4, kOffB1Frag = kOffLate + 12544 * 4,
                     kAuxBytes = kOffB1Frag + 66 * 32 * 32;
    std::vector<unsigned char> aux(kAuxBytes);
    float* a_win = reinterpret_cast<float*>(aux.data());
    double* a_tw = reinterpret_cast<double*>(aux.data() + kOffTw);
    double* a_tw2 = reinterpret_cast<double*>(aux.data() + kOffTw2);
    uint32_t* a_delta = reinterpret_cast<uint32_t*>(aux.data() + kOffDelta);
    float* a_pw = reinterpret_cast<float*>(aux.data() + kOffPw);
    float* a_proj = reinterpret_cast<float*>(aux.data() + kOffProj);
    for (int n = 0; n < 256; ++n) a_win[n] = blob[n];
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n2 = 0; n2 < 8; ++n2) {
            const double th = 2.0 * M_PI * (double)(k1 * n2) / 128.0;
            a_tw[2 * (k1 * 8 + n2)] = cos(th);
            a_tw[2 * (k1 * 8 + n2) + 1] = -sin(th);
        }
    for (int k = 0; k < 128; ++k) {
        const double th = 2.0 * M_PI * (double)k / 256.0;
        a_tw2[2 * k] = cos(th);
        a_tw2[2 * k + 1] = -sin(th);
    }
    {
        // residual of the stored basis against the exact windowed DFT, x 2^24, bf16 (round to nearest even), arranged as
        // mma.m16n8k16 A fragments.  MMA row r: cosine bin r (r <= 128) or sine bin r - 128 (r >= 129) = basis row r + 1;
        // the sine rows of bins 0 and 128 are (numerically) zero and are not corrected.
        auto delta_bf16 = [&](int r, int n) -> uint32_t {
            const int R = r < 129 ? r : r + 1, k = r < 129 ? r : r - 128;
            const double th = 2.0 * M_PI * (double)((k * n) & 255) / 256.0;
            const double exact = (double)blob[n] * (r < 129 ? cos(th) : -sin(th));
            const float d = (float)(((double)blob[R * 256 + n] - exact) * 16777216.0);
            uint32_t u;
            memcpy(&u, &d, 4);
            u += 0x7FFFu + ((u >> 16) & 1u);
            return u >> 16;
        };
        for (int mt = 0; mt < 16; ++mt)
            for (int ks = 0; ks < 16; ++ks)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3, r = mt * 16 + g, c = ks * 16 + 2 * t;
                    uint32_t* o = a_delta + ((size_t)(mt * 16 + ks) * 32 + lane) * 4;
                    o[0] = delta_bf16(r, c) | (delta_bf16(r, c + 1) << 16);
                    o[1] = delta_bf16(r + 8, c) | (delta_bf16(r + 8, c + 1) << 16);
                    o[2] = delta_bf16(r, c + 8) | (delta_bf16(r, c + 9) << 16);
                    o[3] = delta_bf16(r + 8, c + 8) | (delta_bf16(r + 8, c + 9) << 16);
                }
    }
    {
        size_t o = 0;
        for (int i = 0; i < 4; ++i) o += sizes[i];
        const float* pw = blob + o;                         // b1_pw_w [16][258]
        const float* pj = blob + o + sizes[4] + sizes[5];   // b1_proj_w [16][258]
        for (int ci = 0; ci < 258; ++ci)
            for (int co = 0; co < 16; ++co) {
                a_pw[ci * 16 + co] = pw[co * 258 + ci];
                a_proj[ci * 16 + co] = pj[co * 258 + ci];
            }
    }
    {
        // block-1 pointwise weights [16][264 | 264] as m16n8k8 TF32 A fragments, split hi (round to nearest, 10-bit
        // mantissa) | lo (the remainder, rounded the same way)
        auto tf32 = [](float x) -> float {
            uint32_t u;
            memcpy(&u, &x, 4);
            u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
            float r;
            memcpy(&r, &u, 4);
            return r;
        };
        auto wsrc = [&](int co, int k) -> float {
            const int ci = k < 264 ? k : k - 264;
            if (ci >= 258) return 0.f;
            return k < 264 ? a_pw[ci * 16 + co] : a_proj[ci * 16 + co];
        };
        float* fr = reinterpret_cast<float*>(aux.data() + kOffB1Frag);
        for (int ks = 0; ks < 66; ++ks)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3;
                const float w[4] = {wsrc(g, ks * 8 + t), wsrc(g + 8, ks * 8 + t), wsrc(g, ks * 8 + t + 4), wsrc(g + 8, ks * 8 + t + 4)};
                for (int e = 0; e < 4; ++e) {
                    const float hi = tf32(w[e]);
                    fr[(size_t)(ks * 32 + lane) * 8 + e] = hi;
                    fr[(size_t)(ks * 32 + lane) * 8 + 4 + e] = tf32(w[e] - hi);
                }
            }
    }
    size_t late_off[9];
    {
        static const struct { int idx, cout, cin; } kLate[9] = {{8, 16, 16}, {12, 32, 16}, {14, 32, 16}, {16, 32, 32}, {20, 32, 32},
                                                                {22, 32, 32}, {26, 64, 32}, {28, 64, 32}, {30, 64, 64}};
        float* dst = reinterpret_cast<float*>(aux.data() + kOffLate);
        size_t used = 0;
        for (int i = 0; i < 9; ++i) {
            size_t o = 0;
            for (int j = 0; j < kLate[i].idx; ++j) o += sizes[j];
            const float* w = blob + o;
            for (int ci = 0; ci < kLate[i].cin; ++ci)
                for (int co = 0; co < kLate[i].cout; ++co) dst[used + (size_t)ci * kLate[i].cout + co] = w[(size_t)co * kLate[i].cin + ci];
            late_off[i] = kOffLate + used * 4;
            used += (size_t)kLate[i].cout * kLate[i].cin;
        }
    }
    SB_CUDA_CHECK(cudaMalloc(&v->d_aux, kAuxBytes));
    SB_CUDA_CHECK(cudaMemcpy(v->d_aux, aux.data(), kAuxBytes, cudaMemcpyHostToDevice));
    {
        const unsigned char* base = reinterpret_cast<const unsigned char*>(v->d_aux);
        d.win = reinterpret_cast<const float*>(base);
        d.tw = reinterpret_cast<const double2*>(base + kOffTw);
        d.tw2 = reinterpret_cast<const double2*>(base + kOffTw2);
        d.delta_frag = reinterpret_cast<const uint4*>(base + kOffDelta);
        d.b1_pw_t = reinterpret_cast<const float*>(base + kOffPw);
        d.b1_proj_t = reinterpret_cast<const float*>(base + kOffProj);
        d.b1_frag = reinterpret_cast<const uint4*>(base + kOffB1Frag);
        const float** slots[9] = {&d.b1_down_t, &d.b2_pw_t, &d.b2_proj_t, &d.b2_down_t, &d.b3_pw_t, &d.b3_down_t, &d.b4_pw_t, &d.b4_proj_t, &d.b4_down_t};
        for (int i = 0; i < 9; ++i) *slots[i] = reinterpret_cast<const float*>(base + late_off[i]);
    }
    *out = v;
    return SB_OK;
}

int sb_vad_destroy(sb_vad* v) {
    if (!v) return SB_OK;
    cudaFree(v->d_blob); cudaFree(v->d_basis_t); cudaFree(v->d_aux);
    delete v;
    return SB_OK;
}

size_t sb_vad_workspace_bytes(int n_streams, int n_frames) {
    return (size_t)n_streams * n_frames * 64 * sizeof(float) * 2;
}

int sb_vad_score_dev(const sb_vad* v, const float* pcm16k, int64_t pcm_stride, int n_streams, int n_frames, float* h_state,
                     float* c_state, float* probs, void* workspace, void* stream) {
    SB_CHECK_ARG(v && pcm16k && h_state && c_state && probs && workspace, "null pointer");
    SB_CHECK_ARG(n_streams > 0 && n_streams <= 65535 && n_frames > 0, "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    float* feat = (float*)workspace;
    float* h1 = feat + (size_t)n_streams * n_frames * 64;
    SB_ONCE_PER_DEVICE({
        SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_silero_features_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sb::SileroSmem)));
        SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_silero_features_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sb::SileroFftSmem)));
        SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_silero_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sb::Silero2Smem)));
    });
    dim3 grid((n_frames + sb::kSfFrames - 1) / sb::kSfFrames, n_streams);
    if (v->basis_is_dft) {
        // blocks 2-4 run in a second kernel on 32 frames per CTA; its input aliases h1 (written only later, by LSTM layer 1)
        sb::k_silero_features_fft<<<grid, sb::kFfThreads, sizeof(sb::SileroFftSmem), st>>>(pcm16k, pcm_stride, n_frames, h1, v->dev);
        dim3 grid2((n_frames + sb::kS2Frames - 1) / sb::kS2Frames, n_streams);
        sb::k_silero_blocks<<<grid2, 256, sizeof(sb::Silero2Smem), st>>>(h1, n_frames, feat, v->dev);
        sb::g_launches += 1;
    } else
        sb::k_silero_features_direct<<<grid, sb::kSfThreads, sizeof(sb::SileroSmem), st>>>(pcm16k, pcm_stride, n_frames, feat, v->dev);
    const int nb = (n_streams + sb::kLsStreams - 1) / sb::kLsStreams;
    // state layout [2][n_streams][64]: layer-major like vad-rs' h,c [2,1,64] per stream
    sb::k_silero_lstm<false><<<nb, 256, 0, st>>>(feat, n_streams, n_frames, v->dev.lstm_w[0], v->dev.lstm_r[0], v->dev.lstm_b[0],
                                               h_state, c_state, h1, nullptr, nullptr, nullptr);
    sb::k_silero_lstm<true><<<nb, 256, 0, st>>>(h1, n_streams, n_frames, v->dev.lstm_w[1], v->dev.lstm_r[1], v->dev.lstm_b[1],
                                              h_state + (size_t)n_streams * 64, c_state + (size_t)n_streams * 64, nullptr,
                                              v->dev.dec_w, v->dev.dec_b, probs);
    sb::g_launches += 3;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

size_t sb_vad_gate_workspace_bytes(int n_streams, int n_frames) { return (size_t)n_streams * n_frames * sizeof(sb::GatePlan); }

int sb_vad_gate_dev(const float* probs, const float* pcm16k, int64_t pcm_stride, int n_streams, int n_frames, float threshold,
                    int prefill, int hangover, int onset, float* out, int64_t out_stride, int32_t* out_frames,
                    void* workspace, void* stream) {
    SB_CHECK_ARG(probs && pcm16k && out && out_frames && workspace, "null pointer");
    SB_CHECK_ARG(n_streams > 0 && n_streams <= 65535 && n_frames > 0 && n_frames <= 65535 * 32, "bad shape");
    SB_CHECK_ARG(prefill >= 0 && hangover >= 0 && onset >= 1, "bad gate parameters");
    cudaStream_t st = (cudaStream_t)stream;
    sb::GatePlan* plan = (sb::GatePlan*)workspace;
    sb::k_vad_plan<<<(n_streams + 127) / 128, 128, 0, st>>>(probs, n_streams, n_frames, threshold, prefill, hangover, onset, plan, out_frames);
    dim3 grid(n_frames, n_streams);
    sb::k_vad_compact<<<grid, 128, 0, st>>>(pcm16k, pcm_stride, n_frames, plan, out, out_stride, (int)(out_stride / 480));
    sb::g_launches += 2;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

}  // extern "C"

extern "C" int sb_downmix_mono_dev(const void* in, int sample_format, int channels, int64_t in_stride, size_t n_frames, int n_streams,
                                   float* out, int64_t out_stride, void* stream) {
    SB_CHECK_ARG(in && out, "null pointer");
    SB_CHECK_ARG(channels >= 1 && channels <= 64 && n_streams >= 1 && n_streams <= 65535, "channels in [1, 64], n_streams in [1, 65535]");
    SB_CHECK_ARG(in_stride >= (int64_t)n_frames * channels && out_stride >= (int64_t)n_frames, "strides too small");
    if (n_frames == 0) return SB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int bx = (int)std::min<size_t>((n_frames + 255) / 256, 148 * 8);
    dim3 grid(bx, n_streams);
    switch (sample_format) {
        case SB_SAMPLE_F32: sb::k_downmix_mono<float><<<grid, 256, 0, st>>>((const float*)in, channels, in_stride, (int64_t)n_frames, out, out_stride); break;
        case SB_SAMPLE_I16: sb::k_downmix_mono<int16_t><<<grid, 256, 0, st>>>((const int16_t*)in, channels, in_stride, (int64_t)n_frames, out, out_stride); break;
        case SB_SAMPLE_U16: sb::k_downmix_mono<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, channels, in_stride, (int64_t)n_frames, out, out_stride); break;
        default: sb::set_error("invalid argument: sample_format"); return SB_ERR_INVALID;
    }
    sb::g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
