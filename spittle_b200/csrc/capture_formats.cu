// Capture-side data formats, batched over streams:
//   k_pcm_f32_to_i16      history WAV payload: (sample * 32767) as i16          audio_toolkit/audio/utils.rs:17-20
//   k_visualiser_levels   mic-level visualiser: 16 bucket levels per chunk      audio_toolkit/audio/visualizer.rs:84-149
// Both are HBM-bound: 6 B and ~2.1 KB of algorithmic traffic per sample / chunk.
#include "common.cuh"
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace sb {
extern std::atomic<uint64_t> g_launches;

// Rust `as i16` on an f32: truncate toward zero, saturate, NaN -> 0.  cvt.rzi.s32.f32 already saturates and maps NaN to 0.
__device__ __forceinline__ int16_t f32_to_i16_rust(float s) {
    const int v = __float2int_rz(s * 32767.0f);
    return (int16_t)max(-32768, min(32767, v));
}

// 8 samples per thread: two 16-byte loads, one 16-byte store (n8 = n / 8 vector groups; the tail is scalar)
__global__ void __launch_bounds__(256) k_pcm_f32_to_i16(const float* __restrict__ in, int16_t* __restrict__ out, size_t n) {
    const size_t n8 = n >> 3;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i);
        const float4 b = __ldcs(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        uint4 o;
        o.x = (uint16_t)f32_to_i16_rust(a.x) | ((uint32_t)(uint16_t)f32_to_i16_rust(a.y) << 16);
        o.y = (uint16_t)f32_to_i16_rust(a.z) | ((uint32_t)(uint16_t)f32_to_i16_rust(a.w) << 16);
        o.z = (uint16_t)f32_to_i16_rust(b.x) | ((uint32_t)(uint16_t)f32_to_i16_rust(b.y) << 16);
        o.w = (uint16_t)f32_to_i16_rust(b.z) | ((uint32_t)(uint16_t)f32_to_i16_rust(b.w) << 16);
        __stcs(reinterpret_cast<uint4*>(out) + i, o);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const size_t i = (n8 << 3) + threadIdx.x;
        out[i] = f32_to_i16_rust(in[i]);
    }
}

// ---- visualiser ------------------------------------------------------------------------------------------
constexpr int kVisN = 512, kVisBuckets = 16;
struct VisPlan { int start[kVisBuckets], end[kVisBuckets]; int bin_lo, bin_hi; };

// One warp per (stream, chunk): the first 512 samples of the chunk (AudioVisualiser::feed analyses the head of its buffer
// and clears the rest), DC removal, Hann window, and a direct DFT of only the bins the 16 buckets cover (400-4000 Hz:
// bins 4..42 at 48 kHz, 12..128 at 16 kHz) -- lane l holds samples l, l + 32, ..., each bin is a 16-term partial sum per
// lane and a warp reduction.  The twiddle of sample n advances from bin k to k + 1 by a rotation with e^{2 pi i n / 512}
// (exact start value per lane from sincospi; <= 128 rotations, error ~1e-5 relative), so the bin loop has no table
// look-ups: a shared-memory table indexed by (k n) mod 512 was bank-conflict-bound (20 ms per 1.4 M chunks).
__global__ void __launch_bounds__(256) k_visualiser_levels(const float* __restrict__ pcm, int64_t stream_stride, int n_chunks,
                                                           int chunk_len, int n_items, VisPlan plan, float* __restrict__ out) {
    __shared__ float s_pow[8][kVisN / 2 + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + warp;
    if (item >= n_items) return;
    const int stream = item / n_chunks, chunk = item - stream * n_chunks;
    const float* src = pcm + (int64_t)stream * stream_stride + (int64_t)chunk * chunk_len;
    float x[16];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { x[j] = __ldcs(src + lane + 32 * j); sum += x[j]; }
    const float mean = warp_sum(sum) / (float)kVisN;
    float2 rot[16], w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = lane + 32 * j;
        float sn, cs;
        sincospif(2.0f * (float)n / (float)kVisN, &sn, &cs);
        rot[j] = make_float2(cs, sn);
        // Hann exactly as the reference builds it: 0.5 (1 - cos(2 pi i / N)) in f32
        x[j] = (x[j] - mean) * (0.5f * (1.0f - cs));
        sincospif(2.0f * (float)((plan.bin_lo * n) & (kVisN - 1)) / (float)kVisN, &sn, &cs);
        w[j] = make_float2(cs, sn);
    }
    for (int k = plan.bin_lo; k < plan.bin_hi; ++k) {
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            re = fmaf(x[j], w[j].x, re);
            im = fmaf(x[j], w[j].y, im);
            const float wx = w[j].x * rot[j].x - w[j].y * rot[j].y;
            w[j].y = w[j].x * rot[j].y + w[j].y * rot[j].x;
            w[j].x = wx;
        }
        re = warp_sum(re); im = warp_sum(im);
        if (lane == 0) s_pow[warp][k - plan.bin_lo] = re * re + im * im;
    }
    __syncwarp();
    float level = 0.f;
    if (lane < kVisBuckets) {
        const int s = plan.start[lane], e = plan.end[lane];
        if (s < e && e <= kVisN / 2) {
            float p = 0.f;
            for (int k = s; k < e; ++k) p += s_pow[warp][k - plan.bin_lo];
            const float avg = p / (float)(e - s);
            const float db = avg > 1e-12f ? 20.0f * log10f(sqrtf(avg) / (float)kVisN) : -80.0f;
            const float norm = fminf(fmaxf((db - (-55.0f)) / (-8.0f - (-55.0f)), 0.0f), 1.0f);
            level = fminf(fmaxf(powf(norm * 1.3f, 0.7f), 0.0f), 1.0f);
        }
    }
    // the reference smooths in place, left to right: bucket i uses the already smoothed i - 1 and the raw i + 1
    const float right = __shfl_down_sync(0xffffffffu, level, 1);
    float prev = __shfl_sync(0xffffffffu, level, 0);
    float mine = level;
    for (int i = 1; i < kVisBuckets - 1; ++i) {
        const float sm = __shfl_sync(0xffffffffu, level, i) * 0.7f + prev * 0.15f + __shfl_sync(0xffffffffu, right, i) * 0.15f;
        if (lane == i) mine = sm;
        prev = sm;
    }
    if (lane < kVisBuckets) out[(int64_t)item * kVisBuckets + lane] = mine;
}

}  // namespace sb

extern "C" int sb_pcm_f32_to_i16_dev(const float* in, int16_t* out, size_t n, void* stream) {
    SB_CHECK_ARG(in && out, "null pointer");
    SB_CHECK_ARG(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "pcm conversion: 16-byte aligned buffers required");
    if (n == 0) return SB_OK;
    const size_t n8 = std::max<size_t>(n >> 3, 1);
    const int grid = (int)std::min<size_t>((n8 + 255) / 256, (size_t)148 * 16);
    sb::k_pcm_f32_to_i16<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, n);
    sb::g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_visualiser_levels_dev(const float* pcm, int64_t stream_stride, int n_streams, int n_chunks, int chunk_len,
                                        int sample_rate, float* out, void* stream) {
    SB_CHECK_ARG(pcm && out, "null pointer");
    SB_CHECK_ARG(n_streams >= 1 && n_chunks >= 1 && sample_rate > 0, "n_streams, n_chunks >= 1");
    SB_CHECK_ARG(chunk_len >= sb::kVisN, "visualiser: a chunk must hold the 512-sample analysis window");
    SB_CHECK_ARG(stream_stride >= (int64_t)n_chunks * chunk_len && (int64_t)n_streams * n_chunks < (1ll << 31), "visualiser: stride / size");
    // bucket edges exactly as AudioVisualiser::new computes them (visualizer.rs:38-66), in f32
    sb::VisPlan plan;
    const float nyq = (float)sample_rate / 2.0f;
    const float fmin = std::min(400.0f, nyq), fmax = std::min(4000.0f, nyq);
    plan.bin_lo = sb::kVisN; plan.bin_hi = 0;
    for (int b = 0; b < sb::kVisBuckets; ++b) {
        const float r0 = (float)b / (float)sb::kVisBuckets, r1 = (float)(b + 1) / (float)sb::kVisBuckets;
        const float log_start = r0 * r0, log_end = r1 * r1;
        const float start_hz = fmin + (fmax - fmin) * log_start, end_hz = fmin + (fmax - fmin) * log_end;
        int sb_ = (int)((start_hz * (float)sb::kVisN) / (float)sample_rate);
        int eb = (int)((end_hz * (float)sb::kVisN) / (float)sample_rate);
        if (eb <= sb_) eb = sb_ + 1;
        sb_ = std::min(sb_, sb::kVisN / 2); eb = std::min(eb, sb::kVisN / 2);
        plan.start[b] = sb_; plan.end[b] = eb;
        if (sb_ < eb) { plan.bin_lo = std::min(plan.bin_lo, sb_); plan.bin_hi = std::max(plan.bin_hi, eb); }
    }
    if (plan.bin_hi <= plan.bin_lo) { plan.bin_lo = 0; plan.bin_hi = 0; }
    const int n_items = n_streams * n_chunks;
    sb::k_visualiser_levels<<<(n_items + 7) / 8, 256, 0, (cudaStream_t)stream>>>(pcm, stream_stride, n_chunks, chunk_len, n_items, plan, out);
    sb::g_launches += 1;
    SB_CUDA_CHECK(cudaGetLastError());
    return SB_OK;
}

// ---- host-pointer forms: what the reference's (CUDA-free) Rust side calls -- save_wav_file hands over a &[f32], the
//      recorder a captured chunk; staging buffers live for the duration of the call like sb_logmel's ----------------
extern "C" int sb_pcm_f32_to_i16(const float* samples, int16_t* out, size_t n) {
    SB_CHECK_ARG(samples && out, "null pointer");
    if (n == 0) return SB_OK;
    float* d_in = nullptr; int16_t* d_out = nullptr;
    SB_CUDA_CHECK(cudaMalloc(&d_in, n * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, n * sizeof(int16_t));
    int rc = SB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_in, samples, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = sb_pcm_f32_to_i16_dev(d_in, d_out, n, nullptr);
        if (rc == SB_OK) e = cudaMemcpy(out, d_out, n * sizeof(int16_t), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { sb::set_error(std::string("sb_pcm_f32_to_i16: ") + cudaGetErrorString(e)); return SB_ERR_CUDA; }
    return rc;
}

extern "C" int sb_visualiser_levels(const float* pcm, size_t n_samples, int chunk_len, int sample_rate, float* out, int* n_chunks_out) {
    SB_CHECK_ARG(pcm && out, "null pointer");
    SB_CHECK_ARG(chunk_len >= sb::kVisN, "visualiser: a chunk must hold the 512-sample analysis window");
    const int n_chunks = (int)(n_samples / (size_t)chunk_len);
    if (n_chunks_out) *n_chunks_out = n_chunks;
    if (n_chunks == 0) return SB_OK;
    const size_t n = (size_t)n_chunks * chunk_len;
    float *d_in = nullptr, *d_out = nullptr;
    SB_CUDA_CHECK(cudaMalloc(&d_in, n * sizeof(float)));
    cudaError_t e = cudaMalloc(&d_out, (size_t)n_chunks * sb::kVisBuckets * sizeof(float));
    int rc = SB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_in, pcm, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = sb_visualiser_levels_dev(d_in, (int64_t)n, 1, n_chunks, chunk_len, sample_rate, d_out, nullptr);
        if (rc == SB_OK) e = cudaMemcpy(out, d_out, (size_t)n_chunks * sb::kVisBuckets * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { sb::set_error(std::string("sb_visualiser_levels: ") + cudaGetErrorString(e)); return SB_ERR_CUDA; }
    return rc;
}

// ------------------------------------------------------------------------------------------------------------------
// SURVEY.md 8(b) stage entry points by their blueprint names, host pointers, one stream, synchronous: thin staging
// wrappers over the batched device forms in frontend.cu (for a CUDA-free caller such as the reference's Rust side).
// ------------------------------------------------------------------------------------------------------------------
namespace {
struct DevBufs {                       // frees whatever was allocated when the call leaves
    std::vector<void*> p;
    ~DevBufs() { for (void* q : p) cudaFree(q); }
    template <typename V> cudaError_t alloc(V** out, size_t bytes) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
        if (e == cudaSuccess) { p.push_back(q); *out = (V*)q; }
        return e;
    }
};
int cuda_fail(const char* what, cudaError_t e) { sb::set_error(std::string(what) + ": " + cudaGetErrorString(e)); return SB_ERR_CUDA; }
}  // namespace

extern "C" int sb_resample_48k_16k(const float* pcm48k, size_t n_in, float* out16k, size_t out_cap, size_t* n_out) {
    SB_CHECK_ARG(pcm48k && out16k && n_out, "null pointer");
    sb_resampler* r = nullptr;
    int rc = sb_resampler_create(48000, 16000, &r);
    if (rc != SB_OK) return rc;
    size_t n_fed = 0, n_res = 0, n_frames = 0;
    rc = sb_resample_geometry(r, n_in, &n_fed, &n_res, &n_frames);
    const size_t need = n_frames * 480;
    *n_out = need;
    if (rc == SB_OK && need > out_cap) { sb::set_error("sb_resample_48k_16k: output buffer too small"); rc = SB_ERR_INVALID; }
    if (rc == SB_OK && need > 0) {
        DevBufs b;
        float *d_in = nullptr, *d_out = nullptr;
        cudaError_t e = b.alloc(&d_in, n_in * sizeof(float));
        if (e == cudaSuccess) e = b.alloc(&d_out, need * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(d_in, pcm48k, n_in * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            rc = sb_resample_dev(r, d_in, (int64_t)n_in, n_in, 1, d_out, (int64_t)need, nullptr);
            if (rc == SB_OK) e = cudaMemcpy(out16k, d_out, need * sizeof(float), cudaMemcpyDeviceToHost);
        }
        if (e != cudaSuccess) rc = cuda_fail("sb_resample_48k_16k", e);
    }
    sb_resampler_destroy(r);
    return rc;
}

extern "C" int sb_silero_v4(const sb_vad* v, const float* pcm16k, int n_frames, float* h_state, float* c_state, float* probs) {
    SB_CHECK_ARG(v && pcm16k && h_state && c_state && probs, "null pointer");
    if (n_frames <= 0) return SB_OK;
    DevBufs b;
    float *d_pcm = nullptr, *d_h = nullptr, *d_c = nullptr, *d_p = nullptr; void* d_ws = nullptr;
    const size_t n = (size_t)n_frames * 480;
    cudaError_t e = b.alloc(&d_pcm, n * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_h, 2 * 64 * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_c, 2 * 64 * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_p, (size_t)n_frames * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_ws, sb_vad_workspace_bytes(1, n_frames));
    if (e == cudaSuccess) e = cudaMemcpy(d_pcm, pcm16k, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_h, h_state, 2 * 64 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_c, c_state, 2 * 64 * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail("sb_silero_v4", e);
    int rc = sb_vad_score_dev(v, d_pcm, (int64_t)n, 1, n_frames, d_h, d_c, d_p, d_ws, nullptr);
    if (rc != SB_OK) return rc;
    e = cudaMemcpy(probs, d_p, (size_t)n_frames * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_state, d_h, 2 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(c_state, d_c, 2 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? SB_OK : cuda_fail("sb_silero_v4", e);
}

extern "C" int sb_vad_gate(const float* probs, const float* pcm16k, int n_frames, float threshold, int prefill, int hangover,
                           int onset, float* out, size_t out_cap, int* out_frames) {
    SB_CHECK_ARG(probs && pcm16k && out && out_frames, "null pointer");
    SB_CHECK_ARG(prefill >= 0 && hangover >= 0 && onset >= 1, "bad gate parameters");
    *out_frames = 0;
    if (n_frames <= 0) return SB_OK;
    // every onset may re-emit the prefill ring: worst case (prefill + 1) extra frames per `onset` voiced frames
    const size_t max_frames = (size_t)n_frames * (prefill + 1 + onset) / onset + prefill + 1;
    DevBufs b;
    float *d_pr = nullptr, *d_pcm = nullptr, *d_out = nullptr; int32_t* d_cnt = nullptr; void* d_ws = nullptr;
    const size_t n = (size_t)n_frames * 480;
    cudaError_t e = b.alloc(&d_pr, (size_t)n_frames * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_pcm, n * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_out, max_frames * 480 * sizeof(float));
    if (e == cudaSuccess) e = b.alloc(&d_cnt, sizeof(int32_t));
    if (e == cudaSuccess) e = b.alloc(&d_ws, sb_vad_gate_workspace_bytes(1, n_frames));
    if (e == cudaSuccess) e = cudaMemcpy(d_pr, probs, (size_t)n_frames * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_pcm, pcm16k, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_cnt, 0, sizeof(int32_t));
    if (e != cudaSuccess) return cuda_fail("sb_vad_gate", e);
    int rc = sb_vad_gate_dev(d_pr, d_pcm, (int64_t)n, 1, n_frames, threshold, prefill, hangover, onset, d_out, (int64_t)(max_frames * 480),
                             d_cnt, d_ws, nullptr);
    if (rc != SB_OK) return rc;
    int32_t cnt = 0;
    e = cudaMemcpy(&cnt, d_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail("sb_vad_gate", e);
    *out_frames = cnt;
    if ((size_t)cnt * 480 > out_cap) { sb::set_error("sb_vad_gate: output buffer too small"); return SB_ERR_INVALID; }
    e = cudaMemcpy(out, d_out, (size_t)cnt * 480 * sizeof(float), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? SB_OK : cuda_fail("sb_vad_gate", e);
}
