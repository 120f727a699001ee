// k_logmel: whisper.cpp log_mel_spectrogram on sm_100a (SURVEY.md App. C.1, row a4 of 8(a)).
//
// Replaces whisper.cpp `log_mel_spectrogram` + `fft`/`dft` (reached from the reference at
// src-tauri/src/managers/transcription.rs:501-503).  HBM-bound stage: algorithmic bytes per
// 30 s clip = 1.92 MB PCM in + n_mel*3002*4 B out.
//
// Design
//  * one CTA = 32 consecutive frames of one clip (grid = [ceil(n_calc/32), n_clips]);
//    the 32*160+240 samples the tile touches are staged ONCE in shared memory with coalesced
//    loads (the 2.5x frame overlap never goes back to L2/HBM); the reflect pad at the front
//    is an index map, the 30 s zero tail is never read.
//  * two real frames ride one complex 400-point FFT (z = a + i b); 400 = 20 x 20, each thread
//    owns one 20-point DFT held entirely in registers (20 = 4 x 5, radix-4 / radix-5
//    butterflies with compile-time twiddles), two passes through shared memory.
//  * power spectrum by Hermitian pair separation, sparse slaney filterbank (each mel row is
//    a contiguous band), log10, coalesced mel-major store, per-clip max via one atomicMax
//    per CTA.
//  * second kernel applies the clip-global  max-8  clamp and (x+4)/4 in place (the mel of a
//    64-clip batch is L2-resident when it runs, so DRAM sees the output once).
#include "common.cuh"
#include <vector>
#include <cmath>

namespace sb {

constexpr int kNFft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFpb = 32;                 // frames per CTA
constexpr int kFftPerCta = kFpb / 2;     // complex FFTs per CTA
constexpr int kThreads = kFftPerCta * 20;  // 320
constexpr int kTileSamples = (kFpb - 1) * kHop + kNFft;  // 5360
constexpr int kZStride = 21;             // float2 row stride of the 20x20 scratch (bank-conflict-free)
constexpr int kZPerFft = 20 * kZStride;  // 420 float2
constexpr int kPStride = 201;            // odd -> conflict-free when lanes index frames
constexpr int kMaxMel = 128;
constexpr int kMaxBandW = 1024;          // slaney triangles touch <= 2 rows per bin: 402 taps, every band padded to a multiple of 4

struct MelPlanDev {
    const float* tab;        // [400] hann | [800] twiddle (cos, sin) of W400^i, computed once on the host in f64
    int n_mel;
    int n_w;                 // packed weights in w
    const int* band_start;   // [n_mel] first non-zero bin
    const int* band_len;     // [n_mel]
    const int* band_off;     // [n_mel] offset into w
    const float* w;          // packed band weights
};

struct LogmelArgs {
    const float* pcm;
    float* mel;
    int32_t* clip_max;
    // uniform geometry
    int64_t pcm_clip_stride;   // floats between clips
    int n_samples;
    int n_calc;
    int64_t mel_clip_stride;   // floats between clips
    int mel_stride;            // floats between mel rows
    // ragged batch in one launch (engine path): per-clip sample pointers / lengths / frame counts (device arrays); the
    // grid is sized for the longest clip, blocks past a clip's last frame exit.  null -> the uniform geometry above
    const float* const* pcm_ptrs;
    const int* n_samples_v;
    const int* n_calc_v;
};

__constant__ float c_cos20[20];
__constant__ float c_sin20[20];

// ---- register DFTs (forward, e^{-2 pi i nk/N}) ------------------------------------------
__device__ __forceinline__ void dft4(float& r0, float& i0, float& r1, float& i1,
                                     float& r2, float& i2, float& r3, float& i3) {
    float t0r = r0 + r2, t0i = i0 + i2;
    float t1r = r0 - r2, t1i = i0 - i2;
    float t2r = r1 + r3, t2i = i1 + i3;
    float t3r = r1 - r3, t3i = i1 - i3;
    r0 = t0r + t2r; i0 = t0i + t2i;
    r2 = t0r - t2r; i2 = t0i - t2i;
    r1 = t1r + t3i; i1 = t1i - t3r;   // t1 - i t3
    r3 = t1r - t3i; i3 = t1i + t3r;   // t1 + i t3
}

__device__ __forceinline__ void dft5(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2,
                                     float& r3, float& i3, float& r4, float& i4) {
    const float c1 = 0.30901699437494742f;   // cos(2pi/5)
    const float c2 = -0.80901699437494742f;  // cos(4pi/5)
    const float s1 = 0.95105651629515357f;   // sin(2pi/5)
    const float s2 = 0.58778525229247313f;   // sin(4pi/5)
    float t1r = r1 + r4, t1i = i1 + i4;
    float t2r = r2 + r3, t2i = i2 + i3;
    float t3r = r1 - r4, t3i = i1 - i4;
    float t4r = r2 - r3, t4i = i2 - i3;
    float m1r = fmaf(c2, t2r, fmaf(c1, t1r, r0)), m1i = fmaf(c2, t2i, fmaf(c1, t1i, i0));
    float m2r = fmaf(c1, t2r, fmaf(c2, t1r, r0)), m2i = fmaf(c1, t2i, fmaf(c2, t1i, i0));
    float u1r = fmaf(s2, t4r, s1 * t3r), u1i = fmaf(s2, t4i, s1 * t3i);
    float u2r = fmaf(-s1, t4r, s2 * t3r), u2i = fmaf(-s1, t4i, s2 * t3i);
    r0 = r0 + t1r + t2r; i0 = i0 + t1i + t2i;
    r1 = m1r + u1i; i1 = m1i - u1r;   // m1 - i u1
    r4 = m1r - u1i; i4 = m1i + u1r;
    r2 = m2r + u2i; i2 = m2i - u2r;
    r3 = m2r - u2i; i3 = m2i + u2r;
}

// 20-point DFT in registers.  in: x[n] n = 5*n1 + n2 ; out: X[k] k = k1 + 4*k2 (in place,
// output written back so that element index == k).
__device__ __forceinline__ void dft20(float (&xr)[20], float (&xi)[20]) {
    // 5 radix-4 over n1 for each n2: elements n2, 5+n2, 10+n2, 15+n2  -> index k1 at 5*k1+n2
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2)
        dft4(xr[n2], xi[n2], xr[5 + n2], xi[5 + n2], xr[10 + n2], xi[10 + n2], xr[15 + n2], xi[15 + n2]);
    // twiddle W20^{n2*k1}
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
        for (int n2 = 1; n2 < 5; ++n2) {
            const float c = c_cos20[(n2 * k1) % 20], s = c_sin20[(n2 * k1) % 20];  // W = c - i s
            float r = xr[5 * k1 + n2], im = xi[5 * k1 + n2];
            xr[5 * k1 + n2] = fmaf(r, c, im * s);
            xi[5 * k1 + n2] = fmaf(im, c, -r * s);
        }
    // 4 radix-5 over n2 for each k1: elements 5*k1 + n2 -> k2 at 5*k1 + k2
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft5(xr[5 * k1], xi[5 * k1], xr[5 * k1 + 1], xi[5 * k1 + 1], xr[5 * k1 + 2], xi[5 * k1 + 2],
             xr[5 * k1 + 3], xi[5 * k1 + 3], xr[5 * k1 + 4], xi[5 * k1 + 4]);
    // now element 5*k1 + k2 holds X[k1 + 4*k2]; permute to natural order
    float tr[20], ti[20];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) { tr[k1 + 4 * k2] = xr[5 * k1 + k2]; ti[k1 + 4 * k2] = xi[5 * k1 + k2]; }
#pragma unroll
    for (int k = 0; k < 20; ++k) { xr[k] = tr[k]; xi[k] = ti[k]; }
}

struct __align__(16) LogmelSmem {
    float x[kTileSamples + 16];
    float hann[kNFft];
    float2 tw[kNFft];
    float2 z[kFftPerCta * kZPerFft];
    float p[kFpb * kPStride + 4];   // + 4: a padded band of the last frame reads (zero-weighted) past bin 200
    float red[16];
    float bw[kMaxBandW];          // packed band weights (staged once per CTA)
    int bstart[kMaxMel], blen[kMaxMel], boff[kMaxMel];
};

__global__ void __launch_bounds__(kThreads, 2)
k_logmel(LogmelArgs a, MelPlanDev plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LogmelSmem& s = *reinterpret_cast<LogmelSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int clip = blockIdx.y;
    const int frame0 = blockIdx.x * kFpb;
    const int n_calc_c = a.n_calc_v ? __ldg(a.n_calc_v + clip) : a.n_calc;
    if (frame0 >= n_calc_c) return;
    const float* pcm = a.pcm_ptrs ? a.pcm_ptrs[clip] : a.pcm + (int64_t)clip * a.pcm_clip_stride;
    const int n = a.n_samples_v ? __ldg(a.n_samples_v + clip) : a.n_samples;

    // ---- stage samples (padded coordinates p = frame*160 + j ; original index = p - 200) ----
    // p0 - 200 is a multiple of 8, so the interior of the tile is read with aligned 128-bit loads
    const int p0 = frame0 * kHop;
    const int src0 = p0 - 200;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;   // clip bases need not be 16 B aligned
    // Everything this CTA reads from global memory goes out as ONE batch of cp.async copies (tables first, then the tile):
    // the stage was 30 % of the kernel's time as four or five dependent L2 round trips (branchy per-element loads).
    auto cp16 = [](void* dst, const void* src) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
    };
    auto cp4 = [](void* dst, const void* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src));
    };
    // hann[400] and tw[400] are adjacent in LogmelSmem: one coalesced copy of the 1200-float host table
    for (int i = tid; i < 3 * kNFft / 4; i += kThreads) cp16(reinterpret_cast<float4*>(s.hann) + i, reinterpret_cast<const float4*>(plan.tab) + i);
    for (int i = tid; i < plan.n_w; i += kThreads) cp4(s.bw + i, plan.w + i);
    for (int i = tid; i < plan.n_mel; i += kThreads) {
        cp4(s.bstart + i, plan.band_start + i); cp4(s.blen + i, plan.band_len + i); cp4(s.boff + i, plan.band_off + i);
    }
    const bool interior = base_aligned && src0 >= 0 && src0 + kTileSamples <= n;      // CTA-uniform
    if (interior) {
        for (int i4 = tid; i4 < kTileSamples / 4; i4 += kThreads) cp16(s.x + 4 * i4, pcm + src0 + 4 * i4);
    } else {
        for (int i4 = tid; i4 < kTileSamples / 4; i4 += kThreads) {
            const int src = src0 + 4 * i4;
            if (base_aligned && src >= 0 && src + 3 < n) {
                cp16(s.x + 4 * i4, pcm + src);
            } else {
                float t[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int sidx = src + e;
                    if (sidx < 0) sidx = -sidx;          // reflect: padded[p] = samples[200 - p]
                    t[e] = sidx < n ? __ldg(pcm + sidx) : 0.0f;
                }
                *reinterpret_cast<float4*>(s.x + 4 * i4) = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int f = tid / 20;        // complex FFT index (frames 2f, 2f+1)
    const int q = tid % 20;        // n2 in pass 1, k1 in pass 2
    float xr[20], xi[20];
    float2* z = s.z + f * kZPerFft;

    // ---- pass 1: DFT over n1 of z[20*n1 + n2], twiddle W400^{n2*k1}, store S[k1][n2] ----
    {
        const float* xa = s.x + (2 * f) * kHop;
        const float* xb = xa + kHop;
#pragma unroll
        for (int n1 = 0; n1 < 20; ++n1) {
            const int nn = 20 * n1 + q;
            const float w = s.hann[nn];
            // dft20 expects element index 5*a + b  <->  n1 = 5*a + b  (natural order)
            xr[n1] = xa[nn] * w;
            xi[n1] = xb[nn] * w;
        }
        dft20(xr, xi);
#pragma unroll
        for (int k1 = 0; k1 < 20; ++k1) {
            const float2 w = s.tw[(q * k1) % kNFft];
            float r = xr[k1], im = xi[k1];
            z[k1 * kZStride + q] = make_float2(fmaf(r, w.x, im * w.y), fmaf(im, w.x, -r * w.y));
        }
    }
    __syncthreads();
    // ---- pass 2: DFT over n2 of S[k1][n2] -> X[k1 + 20*k2] ----
    {
#pragma unroll
        for (int n2 = 0; n2 < 20; ++n2) {
            float2 v = z[q * kZStride + n2];
            xr[n2] = v.x; xi[n2] = v.y;
        }
        dft20(xr, xi);
    }
    __syncthreads();
    {
        // natural order, stride-1 float2 (420 slots per FFT hold 400)
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) z[q + 20 * k2] = make_float2(xr[k2], xi[k2]);
    }
    __syncthreads();
    // ---- Hermitian pair separation -> power spectra of both frames: thread (f, q) takes the bins q + 20 j ----
    {
        const float2* zf = s.z + f * kZPerFft;
        float* pa = s.p + (2 * f) * kPStride;
        float* pb = pa + kPStride;
#pragma unroll
        for (int j = 0; j < 11; ++j) {
            const int k = q + 20 * j;
            if (k < kBins) {
                const float2 zk = zf[k];
                const float2 zy = zf[k == 0 ? 0 : kNFft - k];
                const float ar = zk.x + zy.x, ai = zk.y - zy.y;     // 2*A[k]
                const float br = zk.y + zy.y, bi = zy.x - zk.x;     // 2*B[k]
                pa[k] = 0.25f * fmaf(ar, ar, ai * ai);
                pb[k] = 0.25f * fmaf(br, br, bi * bi);
            }
        }
        if (tid < 4) s.p[kFpb * kPStride + tid] = 0.0f;
    }
    __syncthreads();
    // ---- mel bands: lane = frame, warp strides over mel rows ----
    const int lane = tid & 31, warp = tid >> 5;
    const int frame = frame0 + lane;
    const bool valid = frame < n_calc_c;
    float* out = a.mel + (int64_t)clip * a.mel_clip_stride;
    const float* prow = s.p + lane * kPStride;
    float vmax = -10.0f;
    for (int j = warp; j < plan.n_mel; j += kThreads / 32) {
        const int b0 = s.bstart[j];
        const int bl = s.blen[j];
        const float* w = s.bw + s.boff[j];
        float acc = 0.0f;
        // bands are padded to a multiple of 4 taps with zero weights (16-byte aligned offsets): one broadcast 128-bit load
        // per four taps, same summation order (a zero-weighted term leaves the sum unchanged)
        for (int k = 0; k < bl; k += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(w + k);
            acc = fmaf(prow[b0 + k], wv.x, acc);
            acc = fmaf(prow[b0 + k + 1], wv.y, acc);
            acc = fmaf(prow[b0 + k + 2], wv.z, acc);
            acc = fmaf(prow[b0 + k + 3], wv.w, acc);
        }
        // lg2.approx (abs error < 2^-22 on the mantissa range) * log10(2): 2 instructions instead of log10f's ~20
        float v = __log2f(fmaxf(acc, 1e-10f)) * 0.30102999566398120f;
        if (valid) {
            out[(int64_t)j * a.mel_stride + frame] = v;
            vmax = fmaxf(vmax, v);
        }
    }
    vmax = warp_max(vmax);
    if (lane == 0) s.red[warp] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = s.red[0];
        for (int w2 = 1; w2 < kThreads / 32; ++w2) m = fmaxf(m, s.red[w2]);
        atomicMax(a.clip_max + clip, float_key(m));
    }
}

__global__ void k_logmel_init(int32_t* clip_max, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) clip_max[i] = float_key(-10.0f);   // the zero-padded tail frames are log10(1e-10)
}

// clamp to (clip max - 8), (x + 4) / 4, in place; also emits the per-clip floor value.
__global__ void k_logmel_norm(float* mel, const int32_t* clip_max, float* floor_val, int n_mel,
                              int n_calc, int64_t mel_clip_stride, int mel_stride) {
    const int clip = blockIdx.y;
    const float mmax = key_float(clip_max[clip]) - 8.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && floor_val) floor_val[clip] = (fmaxf(-10.0f, mmax) + 4.0f) * 0.25f;
    float* base = mel + (int64_t)clip * mel_clip_stride;
    const int per_row4 = mel_stride >> 2;            // mel_stride % 4 == 0 enforced by the host
    const int total4 = n_mel * per_row4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
        const int row = i / per_row4;
        const int c4 = (i - row * per_row4) << 2;
        if (c4 >= n_calc) continue;
        float4* p = reinterpret_cast<float4*>(base + (int64_t)row * mel_stride + c4);
        float4 v = *p;
        v.x = (fmaxf(v.x, mmax) + 4.0f) * 0.25f;
        v.y = (fmaxf(v.y, mmax) + 4.0f) * 0.25f;
        v.z = (fmaxf(v.z, mmax) + 4.0f) * 0.25f;
        v.w = (fmaxf(v.w, mmax) + 4.0f) * 0.25f;
        *p = v;   // columns in [n_calc, mel_stride) are padding; harmless
    }
}

}  // namespace sb

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct sb_melplan {
    int n_mel = 0;
    int* d_start = nullptr;
    int* d_len = nullptr;
    int* d_off = nullptr;
    float* d_w = nullptr;
    float* d_tab = nullptr;
    int n_w = 0;
    bool consts_ready = false;
};

namespace sb {
extern std::atomic<uint64_t> g_launches;

static int upload_twiddles() {
    float c[20], s[20];
    for (int i = 0; i < 20; ++i) {
        c[i] = (float)std::cos(2.0 * M_PI * i / 20.0);
        s[i] = (float)std::sin(2.0 * M_PI * i / 20.0);
    }
    SB_CUDA_CHECK(cudaMemcpyToSymbol(c_cos20, c, sizeof(c)));
    SB_CUDA_CHECK(cudaMemcpyToSymbol(c_sin20, s, sizeof(s)));
    return SB_OK;
}

// Ragged batch, one launch, RAW output (engine path): clip c = pcm_ptrs[c][0 .. n_samples_v[c]) (device pointers: the
// caller's own buffers when they already live in device memory), frames [0, n_calc_v[c]) of mel row-block c receive
// log10(max(mel energy, 1e-10)) and clip_max[c] the monotone key of the clip's maximum.  The global-max clamp and the
// (x + 4) / 4 of whisper.cpp are applied by the consumer (k_im2col_conv1) from clip_max: no second pass over the mel.
int logmel_launch_ragged(const sb_melplan* plan, const float* const* pcm_ptrs, const int* n_samples_v, const int* n_calc_v,
                         int n_clips, int max_n_calc, float* mel, int64_t mel_clip_stride, int mel_stride, int32_t* clip_max,
                         cudaStream_t st) {
    SB_CHECK_ARG(mel_stride >= max_n_calc && mel_stride % 4 == 0, "mel_stride must be >= n_calc and a multiple of 4");
    SB_CHECK_ARG(n_clips > 0 && n_clips <= 65535, "n_clips out of range");
    SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_logmel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(LogmelSmem))); });
    LogmelArgs a{};
    a.mel = mel; a.clip_max = clip_max; a.mel_clip_stride = mel_clip_stride; a.mel_stride = mel_stride;
    a.pcm_ptrs = pcm_ptrs; a.n_samples_v = n_samples_v; a.n_calc_v = n_calc_v; a.n_calc = max_n_calc;
    MelPlanDev pd{plan->d_tab, plan->n_mel, plan->n_w, plan->d_start, plan->d_len, plan->d_off, plan->d_w};
    k_logmel_init<<<ceil_div(n_clips, 256), 256, 0, st>>>(clip_max, n_clips);
    dim3 grid(ceil_div(max_n_calc, kFpb), n_clips);
    k_logmel<<<grid, kThreads, sizeof(LogmelSmem), st>>>(a, pd);
    g_launches += 2;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

int logmel_launch(const sb_melplan* plan, const float* pcm, int n_clips, size_t n_samples,
                  int64_t pcm_clip_stride, float* mel, int64_t mel_clip_stride, int mel_stride,
                  int32_t* clip_max, float* floor_val, cudaStream_t st) {
    int n_len, n_len_org, n_calc;
    sb_logmel_geometry(n_samples, &n_len, &n_len_org, &n_calc);
    SB_CHECK_ARG(n_samples >= 201, "log-mel needs more than 200 samples (whisper.cpp reflect pad)");
    SB_CHECK_ARG(mel_stride >= n_calc && mel_stride % 4 == 0, "mel_stride must be >= n_calc and a multiple of 4");
    SB_CHECK_ARG(n_clips > 0 && n_clips <= 65535, "n_clips out of range");
    SB_ONCE_PER_DEVICE({ SB_CUDA_CHECK(cudaFuncSetAttribute(k_logmel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sizeof(LogmelSmem))); });
    LogmelArgs a{};
    a.pcm = pcm; a.mel = mel; a.clip_max = clip_max;
    a.pcm_clip_stride = pcm_clip_stride; a.n_samples = (int)n_samples; a.n_calc = n_calc;
    a.mel_clip_stride = mel_clip_stride; a.mel_stride = mel_stride;
    MelPlanDev pd{plan->d_tab, plan->n_mel, plan->n_w, plan->d_start, plan->d_len, plan->d_off, plan->d_w};
    k_logmel_init<<<ceil_div(n_clips, 256), 256, 0, st>>>(clip_max, n_clips);
    dim3 grid(ceil_div(n_calc, kFpb), n_clips);
    k_logmel<<<grid, kThreads, sizeof(LogmelSmem), st>>>(a, pd);
    dim3 g2(ceil_div(plan->n_mel * (mel_stride / 4), 256 * 4), n_clips);
    k_logmel_norm<<<g2, 256, 0, st>>>(mel, clip_max, floor_val, plan->n_mel, n_calc, mel_clip_stride, mel_stride);
    g_launches += 3;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}
}  // namespace sb

extern "C" {

int sb_logmel_geometry(size_t n_samples, int* n_len, int* n_len_org, int* n_calc) {
    const int64_t n = (int64_t)n_samples;
    const int64_t nl = (n + 16000 * 30 + 400 - 400) / 160;
    const int64_t no = n + 200 >= 400 ? 1 + (n + 200 - 400) / 160 : 0;
    int64_t nc = (n + 200) / 160 + 1;
    if (nc > nl) nc = nl;
    if (n_len) *n_len = (int)nl;
    if (n_len_org) *n_len_org = (int)no;
    if (n_calc) *n_calc = (int)nc;
    return SB_OK;
}

int sb_melplan_create(const float* filters, int n_mel, sb_melplan** out) {
    SB_CHECK_ARG(filters && out && n_mel > 0 && n_mel <= 128, "filters/out null or n_mel out of range");
    std::vector<int> start(n_mel), len(n_mel), off(n_mel);
    std::vector<float> w;
    for (int j = 0; j < n_mel; ++j) {
        int lo = sb::kBins, hi = -1;
        for (int k = 0; k < sb::kBins; ++k)
            if (filters[j * sb::kBins + k] != 0.0f) { lo = lo < k ? lo : k; hi = k; }
        if (hi < 0) { lo = 0; hi = -1; }
        start[j] = lo; len[j] = (hi - lo + 1 + 3) & ~3; off[j] = (int)w.size();          // padded to four taps, zero weights
        for (int k = lo; k < lo + len[j]; ++k) w.push_back(k <= hi ? filters[j * sb::kBins + k] : 0.0f);
    }
    if (w.empty()) w.push_back(0.0f);
    SB_CHECK_ARG(n_mel <= sb::kMaxMel && (int)w.size() <= sb::kMaxBandW,
                 "mel filterbank too dense for the shared-memory band table (n_mel <= 128, <= 1024 taps after padding each band to four)");
    sb_melplan* p = new sb_melplan();
    p->n_mel = n_mel;
    p->n_w = (int)w.size();
    SB_CUDA_CHECK(cudaMalloc(&p->d_start, n_mel * sizeof(int)));
    SB_CUDA_CHECK(cudaMalloc(&p->d_len, n_mel * sizeof(int)));
    SB_CUDA_CHECK(cudaMalloc(&p->d_off, n_mel * sizeof(int)));
    SB_CUDA_CHECK(cudaMalloc(&p->d_w, w.size() * sizeof(float)));
    SB_CUDA_CHECK(cudaMemcpy(p->d_start, start.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    SB_CUDA_CHECK(cudaMemcpy(p->d_len, len.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    SB_CUDA_CHECK(cudaMemcpy(p->d_off, off.data(), n_mel * sizeof(int), cudaMemcpyHostToDevice));
    SB_CUDA_CHECK(cudaMemcpy(p->d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    {
        std::vector<float> tab(3 * sb::kNFft);
        for (int i = 0; i < sb::kNFft; ++i) {
            const double a = 2.0 * M_PI * i / sb::kNFft;
            tab[i] = (float)(0.5 * (1.0 - std::cos(a)));           // whisper.cpp fill_hann_window (periodic)
            tab[sb::kNFft + 2 * i] = (float)std::cos(a);            // W400^i = cos - i sin
            tab[sb::kNFft + 2 * i + 1] = (float)std::sin(a);
        }
        SB_CUDA_CHECK(cudaMalloc(&p->d_tab, tab.size() * sizeof(float)));
        SB_CUDA_CHECK(cudaMemcpy(p->d_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    int rc = sb::upload_twiddles();
    if (rc != SB_OK) return rc;
    *out = p;
    return SB_OK;
}

int sb_melplan_destroy(sb_melplan* p) {
    if (!p) return SB_OK;
    cudaFree(p->d_start); cudaFree(p->d_len); cudaFree(p->d_off); cudaFree(p->d_w); cudaFree(p->d_tab);
    delete p;
    return SB_OK;
}

int sb_logmel_batch_dev(const sb_melplan* plan, const float* pcm, int n_clips, size_t n_samples,
                        float* mel, int mel_stride, int32_t* clip_max, float* floor_val, void* stream) {
    SB_CHECK_ARG(plan && pcm && mel && clip_max, "null pointer");
    return sb::logmel_launch(plan, pcm, n_clips, n_samples, (int64_t)n_samples, mel,
                             (int64_t)plan->n_mel * mel_stride, mel_stride, clip_max, floor_val,
                             (cudaStream_t)stream);
}

int sb_logmel(const sb_melplan* plan, const float* pcm16k, size_t n_samples, float* out, int* n_len_out,
              int* n_len_org_out) {
    SB_CHECK_ARG(plan && pcm16k && out, "null pointer");
    int n_len, n_len_org, n_calc;
    sb_logmel_geometry(n_samples, &n_len, &n_len_org, &n_calc);
    const int stride = (int)sb::round_up(n_calc, 32);
    float *d_pcm = nullptr, *d_mel = nullptr, *d_floor = nullptr;
    int32_t* d_max = nullptr;
    SB_CUDA_CHECK(cudaMalloc(&d_pcm, n_samples * sizeof(float)));
    SB_CUDA_CHECK(cudaMalloc(&d_mel, (size_t)plan->n_mel * stride * sizeof(float)));
    SB_CUDA_CHECK(cudaMalloc(&d_max, sizeof(int32_t)));
    SB_CUDA_CHECK(cudaMalloc(&d_floor, sizeof(float)));
    SB_CUDA_CHECK(cudaMemcpy(d_pcm, pcm16k, n_samples * sizeof(float), cudaMemcpyHostToDevice));
    int rc = sb::logmel_launch(plan, d_pcm, 1, n_samples, (int64_t)n_samples, d_mel,
                               (int64_t)plan->n_mel * stride, stride, d_max, d_floor, 0);
    if (rc == SB_OK) {
        float floor_v = 0.f;
        cudaError_t e = cudaMemcpy(&floor_v, d_floor, sizeof(float), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy2D(out, (size_t)n_len * sizeof(float), d_mel, (size_t)stride * sizeof(float),
                             (size_t)n_calc * sizeof(float), plan->n_mel, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { sb::set_error(std::string("sb_logmel: ") + cudaGetErrorString(e)); rc = SB_ERR_CUDA; }
        else
            for (int j = 0; j < plan->n_mel; ++j)
                for (int i = n_calc; i < n_len; ++i) out[(size_t)j * n_len + i] = floor_v;
    }
    cudaFree(d_pcm); cudaFree(d_mel); cudaFree(d_max); cudaFree(d_floor);
    if (n_len_out) *n_len_out = n_len;
    if (n_len_org_out) *n_len_org_out = n_len_org;
    return rc;
}

}  // extern "C"
