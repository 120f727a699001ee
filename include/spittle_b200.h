/*
 * spittle_b200.h -- C ABI of libspittle_b200.so (B200 / sm_100a only).
 *
 * The drop-in boundary for the Whisper transcription hot path of tchamp1912/Spittle.
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference repository root).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *  - every function returns SB_OK (0) or a negative sb_status; nothing throws or aborts
 *    across the ABI; the message for the calling thread is available from
 *    sb_last_error() (reference convention: anyhow::Result, transcription.rs:398,503).
 *  - "_dev" entry points take DEVICE pointers and a CUDA stream (cudaStream_t passed as
 *    void*; NULL = default stream) and are asynchronous; all others take HOST pointers
 *    and are synchronous.
 *  - there is no CPU fallback: without a CUDA device every compute entry fails with
 *    SB_ERR_CUDA.
 */
#ifndef SPITTLE_B200_H
#define SPITTLE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_API __attribute__((visibility("default")))

typedef enum sb_status {
    SB_OK = 0,
    SB_ERR_INVALID = -1,     /* bad argument */
    SB_ERR_CUDA = -2,        /* CUDA runtime/driver error (incl. no device) */
    SB_ERR_IO = -3,          /* model file could not be read */
    SB_ERR_FORMAT = -4,      /* not a GGML legacy whisper file / unsupported tensor type */
    SB_ERR_NOT_LOADED = -5,  /* reference: "Model is not loaded for transcription." transcription.rs:427-429 */
    SB_ERR_NOMEM = -6,
    SB_ERR_UNSUPPORTED = -7
} sb_status;

typedef enum sb_dtype {
    SB_DTYPE_BF16 = 0,       /* north-star operand type */
    SB_DTYPE_F16 = 1         /* the reference's own rounding points (ggml f16 x f16 -> f32) */
} sb_dtype;

/* thread-local message of the last failing call on this thread ("" if none). */
SB_API const char* sb_last_error(void);
SB_API const char* sb_version(void);
/* number of kernel launches issued by this library since load (bench.py "gpu_launches"). */
SB_API uint64_t sb_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Log-mel front-end.  Replaces whisper.cpp log_mel_spectrogram, reached from the reference
 * at managers/transcription.rs:501-503 (whisper_engine.transcribe_samples) via
 * transcribe-rs -> whisper-rs WhisperState::full -> whisper_pcm_to_mel.
 * ---------------------------------------------------------------------------------- */
typedef struct sb_melplan sb_melplan;

/* filters: host [n_mel][201] f32 exactly as stored in the GGML model file. */
SB_API int sb_melplan_create(const float* filters, int n_mel, sb_melplan** out);
SB_API int sb_melplan_destroy(sb_melplan* plan);

/* Frame geometry of whisper.cpp for an n-sample clip:
 *   n_len     = (n + 480000 + 400 - 400) / 160   (frames in whisper_mel, 30 s zero pad)
 *   n_len_org = 1 + (n + 200 - 400) / 160        (seek_end)
 *   n_calc    = min((n + 200) / 160 + 1, n_len)  (frames that see audio; the rest are the floor)
 */
SB_API int sb_logmel_geometry(size_t n_samples, int* n_len, int* n_len_org, int* n_calc);

/* One clip, host buffers.  out: [n_mel][n_len] f32 mel-major (whisper_mel layout). */
SB_API int sb_logmel(const sb_melplan* plan, const float* pcm16k, size_t n_samples,
                     float* out, int* n_len, int* n_len_org);

/* Batch of equal-length clips, device buffers, asynchronous.
 *   pcm      [n_clips][n_samples] f32
 *   mel      [n_clips][n_mel][mel_stride] f32; frames [0, n_calc) of each row are written,
 *            mel_stride >= n_calc (use a multiple of 32 for aligned rows)
 *   clip_max [n_clips] i32 scratch (monotone key of the per-clip raw log10 maximum)
 * After the call frames >= n_calc of clip c hold nothing; their value is the per-clip floor
 * returned in floor_val[c] (device, f32, may be NULL). */
SB_API int sb_logmel_batch_dev(const sb_melplan* plan, const float* pcm, int n_clips,
                               size_t n_samples, float* mel, int mel_stride, int32_t* clip_max,
                               float* floor_val, void* stream);

/* ------------------------------------------------------------------------------------
 * Tensor-core GEMM stage entry (parity tests / benchmarks).  C[M,N] = epi(A[M,K] * W[N,K]^T)
 * with tcgen05.mma, TMEM accumulators, TMA-fed.  Replaces ggml's CPU mul_mat (f16 x f16 ->
 * f32) inside the whisper.cpp encoder graph.
 *   dtype      SB_DTYPE_BF16 / SB_DTYPE_F16: element type of A, W and of a 16-bit output
 *   out_f32    1: out is f32 [M, ldo]; 0: out is 16-bit [M, ldo]
 *   bias       f32 [N] or NULL;  act: 0 none, 1 tanh-GELU
 *   residual   f32 rows added after the activation (may alias an f32 out), or NULL;
 *              res_row_mod > 0 indexes it with (row % res_row_mod) (positional embedding)
 * ---------------------------------------------------------------------------------- */
SB_API int sb_gemm_tn_dev(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                          int M, int N, int K, void* out, int64_t ldo, int out_f32,
                          const float* bias, int act, const float* residual, int64_t ldr,
                          int res_row_mod, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPITTLE_B200_H */
