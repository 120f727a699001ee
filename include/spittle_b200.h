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
 * Capture front-end (batched over independent streams).
 *
 * Resampler: replaces rubato::FftFixedIn::<f32>::new(in_hz, out_hz, 1024, 1, 1) + process() as
 * wrapped by FrameResampler (audio_toolkit/audio/resampler.rs:16-98).  Output semantics are those
 * of push(everything) + finish(): the input is zero-padded to whole 1024-sample chunks, only whole
 * rubato blocks (1026 -> 342 at 48 -> 16 kHz) are produced, the last 480-sample frame is zero padded.
 * Integer decimation ratios (96 / 64 / 48 / 32 kHz) run the polyphase tensor-core kernel, 16 kHz passes through, every other
 * rate (44.1 / 24 / 22.05 / 12 / 11.025 / 8 kHz ...) applies rubato's block operator as a dense split-precision GEMM (rates whose
 * operator would exceed 64 M entries, e.g. a prime input rate, return SB_ERR_UNSUPPORTED); that path keeps a
 * device workspace inside the resampler object, so concurrent calls on ONE object must be issued on one stream.
 * Samples are audio in [-1, 1]; the tensor-core paths split them into two f16 halves, so |x| must stay below 65504.
 * ---------------------------------------------------------------------------------- */
typedef struct sb_resampler sb_resampler;
SB_API int sb_resampler_create(int fs_in, int fs_out, sb_resampler** out);
SB_API int sb_resampler_destroy(sb_resampler* r);
/* n_fed: samples handed to rubato; n_out: resampled samples; n_frames: 480-sample frames emitted */
SB_API int sb_resample_geometry(const sb_resampler* r, size_t n_in, size_t* n_fed, size_t* n_out,
                                size_t* n_frames);
/* in [n_streams][in_stride] (n_in valid), out [n_streams][out_stride >= n_frames*480], device, async */
SB_API int sb_resample_dev(const sb_resampler* r, const float* in, int64_t in_stride, size_t n_in,
                           int n_streams, float* out, int64_t out_stride, void* stream);

/* Mono down-mix of interleaved capture frames.  Replaces the cpal input callback of
 * AudioRecorder::build_stream (audio_toolkit/audio/recorder.rs:182-201): every sample goes through cpal's
 * to_sample::<f32>() (i16: x / 32768, u16: (x - 32768) / 32768, f32: identity), a frame becomes the f32 sum
 * of its channels in channel order divided by the channel count; channels == 1 is a plain conversion.
 *   in  [n_streams][in_stride elements] interleaved, n_frames * channels valid per stream
 *   out [n_streams][out_stride] f32, n_frames valid; device pointers, async on `stream` */
typedef enum sb_sample_format { SB_SAMPLE_F32 = 0, SB_SAMPLE_I16 = 1, SB_SAMPLE_U16 = 2 } sb_sample_format;
SB_API int sb_downmix_mono_dev(const void* in, int sample_format, int channels, int64_t in_stride, size_t n_frames,
                               int n_streams, float* out, int64_t out_stride, void* stream);

/* History WAV payload.  Replaces the sample loop of save_wav_file (audio_toolkit/audio/utils.rs:17-20):
 * out[i] = (in[i] * 32767) as i16 with Rust's cast semantics (truncate toward zero, saturate, NaN -> 0).
 * Device pointers, 16-byte aligned, async on `stream`. */
SB_API int sb_pcm_f32_to_i16_dev(const float* in, int16_t* out, size_t n, void* stream);
/* same with host pointers (what save_wav_file's caller has: a &[f32]); synchronous */
SB_API int sb_pcm_f32_to_i16(const float* samples, int16_t* out, size_t n);

/* Mic-level visualiser, batched.  Replaces AudioVisualiser::{new, feed} (audio_toolkit/audio/visualizer.rs:20-149,
 * constructed at audio/recorder.rs:276-282 with window 512, 16 buckets, 400-4000 Hz; fed once per captured chunk,
 * recorder.rs:323): for every chunk the first 512 samples are analysed (DC removal, Hann, power per bucket, dB, range
 * map, gain / curve, left-to-right smoothing).  pcm [n_streams][stream_stride] with n_chunks * chunk_len valid samples,
 * chunk_len >= 512; out [n_streams][n_chunks][16] f32.  Device pointers, async on `stream`. */
SB_API int sb_visualiser_levels_dev(const float* pcm, int64_t stream_stride, int n_streams, int n_chunks, int chunk_len,
                                    int sample_rate, float* out, void* stream);
/* one stream with host pointers: pcm[n_samples] cut into chunks of chunk_len (a trailing partial chunk is ignored);
 * out [n_chunks][16], *n_chunks_out = n_samples / chunk_len; synchronous */
SB_API int sb_visualiser_levels(const float* pcm, size_t n_samples, int chunk_len, int sample_rate, float* out, int* n_chunks_out);

/* Silero VAD v4 (16 kHz branch).  Replaces vad_rs::Vad::{new, compute} over onnxruntime
 * (audio_toolkit/vad/silero.rs:25,41-44).  blob: the f32 tensors of the model in the order of
 * spittle_b200/silero_weights.py BLOB_LAYOUT (read from the reference's silero_vad_v4.onnx).
 * State h, c: [2][n_streams][64] f32, carried across calls like vad-rs carries h/c across frames. */
typedef struct sb_vad sb_vad;
SB_API int sb_vad_create(const float* blob, size_t n_floats, sb_vad** out);
SB_API int sb_vad_destroy(sb_vad* v);
SB_API size_t sb_vad_workspace_bytes(int n_streams, int n_frames);
/* pcm16k [n_streams][pcm_stride] (n_frames*480 valid) -> probs [n_streams][n_frames]; device, async */
SB_API int sb_vad_score_dev(const sb_vad* v, const float* pcm16k, int64_t pcm_stride, int n_streams,
                            int n_frames, float* h_state, float* c_state, float* probs,
                            void* workspace, void* stream);

/* SmoothedVad gate + concatenation of kept frames.  Replaces SmoothedVad::push_frame
 * (audio_toolkit/vad/smoothed.rs:41-96) and handle_frame's out_buf.extend (audio/recorder.rs:284-314).
 * is_voice = prob > threshold (vad/silero.rs:46).  out [n_streams][out_stride]: kept samples,
 * out_frames[s] = number of 480-sample frames written (may exceed n_frames: an onset re-emits the
 * prefill ring exactly like the reference). */
SB_API size_t sb_vad_gate_workspace_bytes(int n_streams, int n_frames);
SB_API int sb_vad_gate_dev(const float* probs, const float* pcm16k, int64_t pcm_stride, int n_streams,
                           int n_frames, float threshold, int prefill, int hangover, int onset,
                           float* out, int64_t out_stride, int32_t* out_frames, void* workspace,
                           void* stream);

/* The capture stages by their blueprint names, host pointers, ONE stream, synchronous: staging
 * wrappers over the batched device forms above for a caller without CUDA code (the reference's Rust side).
 *   sb_resample_48k_16k  FrameResampler::{push(all), finish} at 48 -> 16 kHz (audio/resampler.rs:16-98): *n_out =
 *                        480 * frames emitted (also set when out_cap is too small, which fails with SB_ERR_INVALID)
 *   sb_silero_v4         SileroVad scoring of n_frames 480-sample frames (vad/silero.rs:41-44); h, c: [2][64] f32,
 *                        read and updated like vad-rs carries them from frame to frame
 *   sb_vad_gate          SmoothedVad::push_frame over the frames + concatenation of the kept ones (vad/smoothed.rs:41-96,
 *                        recorder.rs:284-314); *out_frames may exceed n_frames (an onset re-emits the prefill ring) */
SB_API int sb_resample_48k_16k(const float* pcm48k, size_t n_in, float* out16k, size_t out_cap, size_t* n_out);
SB_API int sb_silero_v4(const sb_vad* v, const float* pcm16k, int n_frames, float* h_state, float* c_state, float* probs);
SB_API int sb_vad_gate(const float* probs, const float* pcm16k, int n_frames, float threshold, int prefill, int hangover,
                       int onset, float* out, size_t out_cap, int* out_frames);

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

/* Stage entries used by the parity tests (device pointers). */
SB_API int sb_layernorm_dev(int dtype, const float* x, const float* gamma, const float* beta,
                            void* out16, float* out32, int rows, int d, void* stream);
/* qkv [n_windows*n_ctx, 3*d_model] 16-bit (columns Q|K|V, heads of 64) -> out [n_windows*n_ctx, d_model] */
SB_API int sb_attn_enc_dev(int dtype, const void* qkv, void* out, int n_windows, int n_ctx,
                           int d_model, int n_head, void* stream);

/* Decoder-step weight-streaming GEMM: Y[Bn,N] = epi(X[Bn,K] W[N,K]^T), Bn small (HBM-bound on W).
 * out32 (f32 [Bn, ldo32]) and/or out16 (16-bit [Bn, ldo16]); residual f32 may alias out32. */
SB_API int sb_skinny_gemm_dev(int dtype, const void* X, int64_t ldx, const void* W, int64_t ldw,
                              int Bn, int N, int K, const float* bias, int act, const float* residual,
                              int64_t ldr, float* out32, int64_t ldo32, void* out16, int64_t ldo16,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * Engine: the native equivalent of transcribe-rs' WhisperEngine as the reference uses it
 *   new() + load_model(&path)      managers/transcription.rs:262-263  -> sb_engine_create
 *   unload_model()                 managers/transcription.rs:183-189  -> sb_engine_destroy
 *   transcribe_samples(audio, Some(WhisperInferenceParams{language, translate,
 *       initial_prompt, ..}))      managers/transcription.rs:494-503  -> sb_transcribe
 * An engine may be created / used / destroyed from different threads; calls on one engine
 * must be serialised by the caller (the reference holds its engine mutex across the whole
 * inference, transcription.rs:437).  An engine drives one CUDA device, or several when
 * sb_config.devices is set (one replica per device; clips are independent, no collective).
 * ---------------------------------------------------------------------------------- */
typedef struct sb_engine sb_engine;

typedef struct sb_config {
    const char* model_path;   /* GGML legacy ggml-*.bin (f32 / f16 / q4_0 / q4_1 / q5_0 / q5_1 / q8_0 / q5_K tensors) */
    int device;               /* CUDA device ordinal */
    int max_batch;            /* decode slots: windows decoded together (0 -> 64) */
    int dtype;                /* sb_dtype: operand type of the tensor-core GEMMs */
    int use_cuda_graph;       /* 1: capture the decoder step into a CUDA graph (default), 0: plain launches */
    const int* devices;       /* NULL, or n_devices distinct CUDA ordinals: one replica of the model per device; then     */
    int n_devices;            /* sb_transcribe_batch sends clip i to devices[i % n_devices], one worker thread per device  */
} sb_config;

typedef struct sb_model_info {
    int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
    int32_t token_eot, token_sot, token_beg, token_blank;
} sb_model_info;

/* Decode policy.  sb_params_default() gives the configuration pinned for parity in
 * language "en", transcribe, timestamps on, suppress_blank, no_context (nothing is carried
 * from one CALL to the next; inside a call whisper_full conditions every window on the previous windows' text,
 * n_max_text_ctx 16384), max_initial_ts 1.0, greedy (temperature 0, no fallback), n_max = n_text_ctx/2 - 4. */
typedef struct sb_params {
    const char* language;        /* "en", "de", ...; NULL, "" or "auto" = the reference's default "auto": the language is
                                    detected on the clip's first window like whisper_full (whisper_lang_auto_detect) */
    int translate;               /* WhisperInferenceParams.translate: task token <|translate|> instead of <|transcribe|> */
    const char* initial_prompt;  /* UTF-8 or NULL.  WhisperInferenceParams.initial_prompt (the reference sets it from the jargon
                                    dictionary, transcription.rs:461-499): tokenised with sb_tokenize and prepended as
                                    [prev] + tokens to the decoder prompt of every window, like whisper.cpp's prompt_past */
    int no_timestamps;
    int suppress_blank;
    int single_segment;
    float max_initial_ts;
    int n_max_tokens;            /* 0 -> n_text_ctx/2 - 4 (whisper.cpp); tests may cap it */
    int max_windows;             /* 0 -> unlimited; safety cap on the seek loop */
    int n_max_text_ctx;          /* whisper_full_params.n_max_text_ctx (default 16384): tokens of text context carried from one
                                    window of a clip to the next ([prev] + last min(this, n_text_ctx/2) tokens); <= 0 disables */
    /* Temperature fallback (whisper_full): a window is decoded at `temperature`; when the result fails (timestamps went
     * backwards, no timestamp before n_max, token entropy of the last 32 tokens < entropy_thold) or its mean token
     * log-probability is below logprob_thold, it is decoded again at temperature + temperature_inc, ... up to 1.0.  At a
     * temperature > 0 tokens are DRAWN (best_of = 1) with a per-clip std::mt19937(0) stream like whisper.cpp; from 0.5 up the
     * text context is dropped from the prompt.  temperature_inc = 0 (the pinned parity configuration, sb_params_default)
     * disables the fallback; whisper.cpp's own defaults are 0.0 / 0.2 / -1.0 / 2.4. */
    float temperature;
    float temperature_inc;
    float logprob_thold;
    float entropy_thold;
    int suppress_nst;            /* whisper_full_params.suppress_nst (transcribe-rs: suppress_non_speech_tokens): the non-speech symbol
                                    tokens of whisper.cpp's list (quotes, brackets, music notes ..., with and without a leading
                                    space, plus " -" and " '") get -inf in whisper_process_logits.  whisper.cpp's default is off */
} sb_params;

typedef struct sb_window_info {
    int32_t seek;            /* window start, mel frames (10 ms) */
    int32_t n_tokens;        /* tokens sampled in this window */
    int32_t result_len;      /* tokens kept (whisper.cpp result_len) */
    int32_t seek_delta;
    int32_t failed;
    int32_t token_offset;    /* offset of this window's sampled tokens in sb_result.sampled */
    int32_t n_prompt;        /* decoder prompt length of this window: [prev + text context] + [sot, lang, task, ...] */
    float temperature;       /* temperature of the accepted decode of this window */
    int32_t n_attempts;      /* decodes of this window (1 + temperature fallbacks) */
    float avg_logprob;       /* mean log-probability of the kept tokens (whisper_sequence_score) */
} sb_window_info;

/* One segment of the transcript (transcribe-rs TranscriptionResult.segments = whisper_full_get_segment_{t0,t1,text}):
 * the text between two timestamp tokens. */
typedef struct sb_segment {
    int64_t t0, t1;          /* start / end in 10 ms units from the start of the clip */
    const char* text;        /* UTF-8, NUL-terminated, untrimmed; points into sb_result.segment_text */
    size_t text_len;
    int32_t token_offset;    /* this segment's tokens inside sb_result.tokens */
    int32_t n_tokens;
} sb_segment;

typedef struct sb_result {
    char* text;  size_t text_len;          /* UTF-8 bytes, trimmed like transcribe-rs; NUL-terminated */
    int32_t* tokens; size_t n_tokens;      /* kept tokens (incl. timestamps / EOT), all windows */
    int32_t* sampled; size_t n_sampled;    /* every sampled token, all windows */
    float* margins;                        /* [n_sampled] top1-top2 of the filtered logits */
    int32_t* tids;                         /* [n_sampled] whisper_token_data.tid: most probable timestamp token of each step */
    float* logprobs;                       /* [n_sampled] whisper_token_data.plog: log-probability of each sampled token */
    sb_window_info* windows; size_t n_windows;
    sb_segment* segments; size_t n_segments;
    char* segment_text;                    /* storage of the segment texts */
    float ms_mel, ms_encode, ms_decode;    /* device time of the batch this clip was part of */
    int status;                            /* per-clip sb_status (batch API) */
    int lang_id;                           /* whisper language id used for the prompt (detected when params.language is NULL /
                                              "auto", like whisper_full); -1 for English-only models */
} sb_result;

/* Cumulative engine counters (bench.py): device times are CUDA-event brackets on the engine's
 * stream.  gemm_* / attn_* are only filled while profiling is enabled (one event pair per launch). */
typedef struct sb_stats {
    double clips, windows, rounds, decoder_steps, tokens_sampled;
    double pcm_bytes;                 /* clip bytes copied to the device (H2D when the caller passes host memory) */
    double h2d_bytes, d2h_bytes;      /* control traffic: decode state up, tokens / state down */
    double mel_ms, encode_ms, decode_ms;
    double gemm_ms, gemm_flops, gemm_launches;     /* tcgen05 GEMM launches of the encoder + cross-KV */
    double attn_ms, attn_flops, attn_launches;     /* encoder attention launches */
    /* sb_engine_set_profile(e, 2) only: device-side launch trace of the decoder step (first block start -> last block end
     * of every launch, %globaltimer; the step keeps its CUDA graph and PDL overlap): projections (weight bytes),
     * cross-attention (K/V bytes of the live sequences), LayerNorm, self-attention; dstep = first start -> last end of the
     * traced launches of one lane-step */
    double skinny_ms, skinny_bytes, skinny_launches;
    double xattn_ms, xattn_bytes, xattn_launches;
    double dln_ms, dln_launches, dself_ms, dself_launches, dstep_ms, dstep_count;
    double prefill_rows;              /* prompt tokens that went through the batched prefill pass */
    double fallbacks;                 /* windows decoded again at a higher temperature */
} sb_stats;

SB_API void sb_params_default(sb_params* p);
/* the CUDA stream (cudaStream_t) every kernel of this engine is launched on */
SB_API void* sb_engine_stream(sb_engine* e);
SB_API int sb_engine_set_profile(sb_engine* e, int enable);
SB_API int sb_engine_stats(sb_engine* e, sb_stats* out, int reset);
SB_API int sb_engine_create(const sb_config* cfg, sb_engine** out);
SB_API int sb_engine_destroy(sb_engine* e);
SB_API int sb_engine_info(const sb_engine* e, sb_model_info* info);
/* number of devices (model replicas) this engine drives */
SB_API int sb_engine_device_count(const sb_engine* e);
/* id -> token bytes (whisper_token_to_str); returns length, copies at most cap bytes */
SB_API int sb_token_text(const sb_engine* e, int32_t id, char* buf, int cap);
/* whisper.cpp's tokeniser (whisper_tokenize, what turns WhisperInferenceParams::initial_prompt into prompt tokens,
 * managers/transcription.rs:461-499): regex word split, then greedy longest vocabulary match per word.  Writes at most
 * `cap` ids and returns the number of tokens of the text (>= 0), or a negative sb_status. */
SB_API int sb_tokenize(const sb_engine* e, const char* text, int32_t* tokens, int cap);

/* One clip: 16 kHz mono f32 host samples -> text.  Empty input returns SB_OK with empty text
 * (reference: transcription.rs:412-416); < 1 s of audio returns empty text like whisper.cpp
 * (the reference pads such clips upstream, managers/audio.rs:466-475). */
SB_API int sb_transcribe(sb_engine* e, const float* pcm16k, size_t n_samples, const sb_params* p,
                         sb_result* out);
/* Independent clips batched on this engine's GPU(s).  out: array of `count` results.  All clips of the call share the
 * engine's decode slots: every clip's seek loop runs window by window, a finished window's slot is refilled with the
 * next ready window of any clip.  With sb_config.devices the clips are split over the devices (no collective: clips do
 * not interact) and decoded concurrently by one worker thread per device. */
SB_API int sb_transcribe_batch(sb_engine* e, const float* const* pcm16k, const size_t* n_samples,
                               size_t count, const sb_params* p, sb_result* out);
SB_API void sb_result_free(sb_result* r);

/* Parity hooks (host buffers).
 * sb_encode: mel windows [n_windows][n_mel][3000] f32 -> encoder output [n_windows][1500][d] f32
 *            (after ln_post; computed in the engine dtype, returned widened to f32). */
SB_API int sb_encode(sb_engine* e, const float* mel_windows, int n_windows, float* enc_out);
/* sb_decode_trace: encode + cross-KV + n_steps decoder steps with optional teacher forcing.
 *   forced      [n_windows][n_steps] token ids, < 0 = use the sampled token; may be NULL
 *   logits_out  [n_windows][n_steps][n_vocab] raw f32 logits before filtering; may be NULL
 *   tokens_out  [n_windows][n_steps] sampled tokens;  margins_out same shape or NULL
 * seek = 0 and seek_end = seek_end[w] for every window. */
/* (the blueprint's `sb_decode_step` is this entry with n_steps = 1) */
SB_API int sb_decode_trace(sb_engine* e, const float* mel_windows, int n_windows, const int32_t* seek_end,
                           const sb_params* p, const int32_t* forced, int n_steps, float* logits_out,
                           int32_t* tokens_out, float* margins_out);

/* ABI self-description: one row per field of the structs above (struct name, field name, sizeof(struct),
 * offsetof(struct, field)).  Writes at most `cap` rows, returns the number of rows.  A binding checks its own layout
 * against this table (tests/test_abi.py does it for ctypes; rust/spittle-b200-sys asserts the same numbers). */
typedef struct sb_abi_field { const char* struct_name; const char* field; int struct_size; int offset; } sb_abi_field;
SB_API int sb_abi_layout(sb_abi_field* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* SPITTLE_B200_H */
