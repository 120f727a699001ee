// Capture front-end on sm_100a (rows a9-a13 of SURVEY.md 8(a)):
//   k_resample_mma     rubato FftFixedIn (48 -> 16 kHz) as a decimating FIR with the same 1026
//                      Blackman-Harris^2 sinc taps, run as a Toeplitz GEMM on the tensor cores (3xTF32)
//                      (reference: audio_toolkit/audio/resampler.rs:24, 51-56; equivalence of the FFT and
//                      FIR forms: SURVEY.md App. B, ~1e-9; the CUDA-core form of round 1 was removed)
//   k_silero_features  Silero v4 per-frame front: reflect pad, STFT conv, magnitude, log, adaptive
//                      normalisation, 4 separable conv blocks (reference: vad/silero.rs:41-44 ->
//                      vad-rs -> onnxruntime; graph first-hand from silero_vad_v4.onnx, App. A)
//   k_silero_lstm      the two LSTM(64) layers + decoder + sigmoid, state carried across frames,
//                      gate weights resident in registers for the whole sequence
//   k_vad_plan / k_vad_compact   SmoothedVad onset / hangover / prefill FSM and the capture consumer's
//                      concatenation of kept frames (reference: vad/smoothed.rs:41-96,
//                      audio/recorder.rs:284-314)
// All fp32 (Silero decisions are thresholded: SURVEY 7.3 item 6); first, correctness-oriented
// version on the CUDA cores -- the tensor-core (split-precision) formulation of the FIR and of the
// STFT convolution is the next step.
#include "common.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>

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
// The same FIR on the tensor cores.  16 consecutive outputs m0+16n .. m0+16n+15 are one GEMM column block:
//   y[m0 + 16 n + i] = sum_j A[i][j] * B[j][n],   A[i][j] = h[(T-1) + D i - j]  (0 outside [0, T)),
//                                                B[j][n] = x[D (m0 + 16 n) - (T-1) + j],   j in [0, T + 15 D)
// A is a constant 16 x (T + 15 D) Toeplitz expansion of the taps (4 % more MACs than the direct form at D = 3), B is a
// sliding window over ONE contiguous input segment (column n starts 16 D samples after column n-1), so both operands
// are built from shared memory with index arithmetic only.  mma.sync.m16n8k8 TF32 with the 3-pass split
// (a_hi b_hi + a_hi b_lo + a_lo b_hi, hi = cvt.rna.tf32, lo = the exact f32 remainder): products are exact to ~2^-21,
// accumulation is f32 -- the result stays within the 1e-5 parity bound of the f64 rubato restatement.
// CTA = 256 threads, 2048 outputs (128 column blocks = 16 n-tiles, two per warp) of one stream.
// ------------------------------------------------------------------------------------------
constexpr int kRmBlocks = 128;                    // 16-output column blocks per CTA
constexpr int kRmTile = 16 * kRmBlocks;           // outputs per CTA

__device__ __forceinline__ uint32_t tf32_hi(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_hi(x);
    lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) k_resample_mma(const float* __restrict__ x, int64_t x_stride, int n_in,
                                                      float* __restrict__ y, int64_t y_stride, int n_out,
                                                      const float* __restrict__ h, int T, int D, int Kp) {
    extern __shared__ float s_rm[];
    const int hp_pad = Kp;                                  // hp[idx + hp_pad] = h[idx], zero outside [0, T)
    const int hp_len = T + 15 * D + Kp + 8;
    float* hp = s_rm;
    float* xs = s_rm + ((hp_len + 3) & ~3);                 // input segment: Kp + 16 D (kRmBlocks - 1) samples
    const int seg = Kp + 16 * D * (kRmBlocks - 1);
    const int stream = blockIdx.y;
    const int m0 = blockIdx.x * kRmTile;
    const float* xin = x + (int64_t)stream * x_stride;
    for (int i = threadIdx.x; i < hp_len; i += 256) {
        const int idx = i - hp_pad;
        hp[i] = (idx >= 0 && idx < T) ? __ldg(h + idx) : 0.0f;
    }
    const int64_t base0 = (int64_t)D * m0 - (T - 1);
    for (int i = threadIdx.x; i < seg; i += 256) {
        const int64_t src = base0 + i;
        xs[i] = (src >= 0 && src < n_in) ? __ldg(xin + src) : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float acc[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
    // A[i][j] = hp[hp_pad + (T-1) + D i - j]; rows g and g + 8, columns 8 s + t and 8 s + t + 4
    const float* ha = hp + hp_pad + (T - 1) + D * g - t;
    const float* hb = ha + 8 * D;
    // B[j][n] = xs[j + 16 D n]; this warp's n-tiles: 2 warp and 2 warp + 1 (columns 8 nt + g)
    const float* xb0 = xs + 16 * D * (8 * (2 * warp) + g) + t;
    const float* xb1 = xb0 + 16 * D * 8;
    const int n_steps = Kp >> 3;
#pragma unroll 2
    for (int sI = 0; sI < n_steps; ++sI) {
        const int j = 8 * sI;
        uint32_t ah[4], al[4];
        split_tf32(ha[-j], ah[0], al[0]);
        split_tf32(hb[-j], ah[1], al[1]);
        split_tf32(ha[-j - 4], ah[2], al[2]);
        split_tf32(hb[-j - 4], ah[3], al[3]);
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(xb0[j], bh0, bl0);
        split_tf32(xb0[j + 4], bh1, bl1);
        mma_tf32(acc[0], ah, bh0, bh1);
        mma_tf32(acc[0], ah, bl0, bl1);
        mma_tf32(acc[0], al, bh0, bh1);
        split_tf32(xb1[j], bh0, bl0);
        split_tf32(xb1[j + 4], bh1, bl1);
        mma_tf32(acc[1], ah, bh0, bh1);
        mma_tf32(acc[1], ah, bl0, bl1);
        mma_tf32(acc[1], al, bh0, bh1);
    }
    // C[i][n]: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)  ->  output m0 + 16 n + i
    float* yo = y + (int64_t)stream * y_stride;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int nt = 2 * warp + u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int n = 8 * nt + 2 * t + (e & 1);
            const int i = g + ((e >> 1) << 3);
            const int m = m0 + 16 * n + i;
            if (m < n_out) yo[m] = acc[u][e];
        }
    }
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
__global__ void __launch_bounds__(kSfThreads, 3) k_silero_features(const float* __restrict__ pcm, int64_t pcm_stride, int n_frames,
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
// LSTM layer (ONNX gate order i, o, f, c).  CTA = 256 threads (one gate row each, its 128 weights
// in registers for the whole sequence) x kLsStreams streams processed in lock step.
//   xin  [n_streams][n_frames][64]   hout [n_streams][n_frames][64] (layer 1) or probs (layer 2)
//   h, c [n_streams][64] in/out state of this layer
// ------------------------------------------------------------------------------------------
constexpr int kLsStreams = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <bool LAST>
__global__ void __launch_bounds__(256, 1) k_silero_lstm(const float* __restrict__ xin, int n_streams, int n_frames,
                                                        const float* __restrict__ W, const float* __restrict__ R,
                                                        const float* __restrict__ B, float* __restrict__ h_state,
                                                        float* __restrict__ c_state, float* __restrict__ hout,
                                                        const float* __restrict__ dec_w, const float* __restrict__ dec_b,
                                                        float* __restrict__ probs) {
    __shared__ __align__(16) float xh[kLsStreams][128];
    __shared__ float gates[kLsStreams][256];
    __shared__ float cst[kLsStreams][64];
    const int g = threadIdx.x;
    const int s0 = blockIdx.x * kLsStreams;
    float w[128];
#pragma unroll
    for (int j = 0; j < 64; ++j) { w[j] = __ldg(W + g * 64 + j); w[64 + j] = __ldg(R + g * 64 + j); }
    const float bias = __ldg(B + g) + __ldg(B + 256 + g);
    for (int i = threadIdx.x; i < kLsStreams * 64; i += 256) {
        const int s = i >> 6, j = i & 63;
        const bool ok = s0 + s < n_streams;
        xh[s][64 + j] = ok ? h_state[(int64_t)(s0 + s) * 64 + j] : 0.f;
        cst[s][j] = ok ? c_state[(int64_t)(s0 + s) * 64 + j] : 0.f;
    }
    for (int t = 0; t < n_frames; ++t) {
        for (int i = threadIdx.x; i < kLsStreams * 64; i += 256) {
            const int s = i >> 6, j = i & 63;
            xh[s][j] = (s0 + s < n_streams) ? __ldg(xin + ((int64_t)(s0 + s) * n_frames + t) * 64 + j) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int s = 0; s < kLsStreams; ++s) {
            float a = bias;
            const float4* v = reinterpret_cast<const float4*>(xh[s]);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 q = v[j];
                a = fmaf(w[4 * j], q.x, a); a = fmaf(w[4 * j + 1], q.y, a);
                a = fmaf(w[4 * j + 2], q.z, a); a = fmaf(w[4 * j + 3], q.w, a);
            }
            gates[s][g] = a;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < kLsStreams * 64; i += 256) {
            const int s = i >> 6, j = i & 63;
            const float ig = sigmoidf_(gates[s][j]), og = sigmoidf_(gates[s][64 + j]);
            const float fg = sigmoidf_(gates[s][128 + j]), cg = tanhf(gates[s][192 + j]);
            const float c = fg * cst[s][j] + ig * cg;
            const float h = og * tanhf(c);
            cst[s][j] = c;
            xh[s][64 + j] = h;
            if (!LAST && s0 + s < n_streams) hout[((int64_t)(s0 + s) * n_frames + t) * 64 + j] = h;
        }
        __syncthreads();
        if (LAST && threadIdx.x < kLsStreams && s0 + threadIdx.x < n_streams) {
            const int s = threadIdx.x;
            float a = __ldg(dec_b);
            for (int j = 0; j < 64; ++j) a = fmaf(__ldg(dec_w + j), fmaxf(xh[s][64 + j], 0.f), a);
            probs[(int64_t)(s0 + s) * n_frames + t] = sigmoidf_(a);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kLsStreams * 64; i += 256) {
        const int s = i >> 6, j = i & 63;
        if (s0 + s < n_streams) { h_state[(int64_t)(s0 + s) * 64 + j] = xh[s][64 + j]; c_state[(int64_t)(s0 + s) * 64 + j] = cst[s][j]; }
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
    float* d_h = nullptr;
};

struct sb_vad {
    float* d_blob = nullptr;
    float* d_basis_t = nullptr;
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
    if (b != 1 || (a != 1 && a != 2 && a != 3 && a != 4 && a != 6)) {
        sb::set_error("resampler: only integer decimation ratios (in/out in {1,2,3,4,6}) are implemented; "
                      "rubato's rational ratios (e.g. 44100 -> 16000) are not yet");
        return SB_ERR_UNSUPPORTED;
    }
    sb_resampler* r = new sb_resampler();
    r->fs_in = fs_in; r->fs_out = fs_out; r->decim = a;
    const int fft_chunks = (1024 + a - 1) / a;               // rubato: ceil(chunk_size_in / (fs_in / gcd))
    r->fft_in = fft_chunks * a; r->fft_out = fft_chunks * b; r->n_taps = r->fft_in;
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
        SB_CUDA_CHECK(cudaMalloc(&r->d_h, N * sizeof(float)));
        SB_CUDA_CHECK(cudaMemcpy(r->d_h, hf.data(), N * sizeof(float), cudaMemcpyHostToDevice));
    }
    *out = r;
    return SB_OK;
}

int sb_resampler_destroy(sb_resampler* r) {
    if (!r) return SB_OK;
    cudaFree(r->d_h);
    delete r;
    return SB_OK;
}

int sb_resample_geometry(const sb_resampler* r, size_t n_in, size_t* n_fed, size_t* n_out, size_t* n_frames) {
    SB_CHECK_ARG(r, "null resampler");
    const size_t fed = (n_in + 1023) / 1024 * 1024;          // push() chunks + finish() zero pad
    size_t out = fed;
    if (r->decim > 1) out = fed / (size_t)r->fft_in * (size_t)r->fft_out;   // whole rubato blocks only
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
    SB_CUDA_CHECK(cudaMemset2DAsync(out, out_stride * sizeof(float), 0, n_frames * 480 * sizeof(float), n_streams, st));
    if (r->decim == 1) {
        SB_CUDA_CHECK(cudaMemcpy2DAsync(out, out_stride * sizeof(float), in, in_stride * sizeof(float), n_in * sizeof(float),
                                        n_streams, cudaMemcpyDeviceToDevice, st));
        return SB_OK;
    }
    if (n_out == 0) return SB_OK;
    const int D = r->decim;
    {
        // tensor-core Toeplitz form (k_resample_mma)
        const int T = r->n_taps;
        const int Kp = (T + 15 * D + 7) & ~7;
        const size_t smem = (size_t)(((T + 15 * D + Kp + 8 + 3) & ~3) + Kp + 16 * D * (sb::kRmBlocks - 1)) * sizeof(float);
        SB_CHECK_ARG(smem <= 200 * 1024, "resampler: filter too long for the shared-memory segment");
        SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_resample_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); });
        dim3 grid((unsigned)((n_out + sb::kRmTile - 1) / sb::kRmTile), n_streams);
        sb::k_resample_mma<<<grid, 256, smem, st>>>(in, in_stride, (int)n_in, out, out_stride, (int)n_out, r->d_h, T, D, Kp);
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
    *out = v;
    return SB_OK;
}

int sb_vad_destroy(sb_vad* v) {
    if (!v) return SB_OK;
    cudaFree(v->d_blob); cudaFree(v->d_basis_t);
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
    SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(sb::k_silero_features, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(sb::SileroSmem))); });
    dim3 grid((n_frames + sb::kSfFrames - 1) / sb::kSfFrames, n_streams);
    sb::k_silero_features<<<grid, sb::kSfThreads, sizeof(sb::SileroSmem), st>>>(pcm16k, pcm_stride, n_frames, feat, v->dev);
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
