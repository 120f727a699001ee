// Shared declarations of the decoder-step kernels and the encoder-side launchers.
#pragma once
#include "common.cuh"

namespace sb {

struct SpecialIds {
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang_first, num_languages, blank;
};

// longest decoder prompt: [prev] + n_text_ctx/2 = 224 context tokens + [sot, lang, task, notimestamps] (+ the extra leading
// [sot] of a language-detect restart), rounded up to a multiple of 8
constexpr int kMaxPrompt = 232;

// per-sequence decode state, device resident (whisper_full's per-decoder bookkeeping, App. C.4).  Every sequence of a
// batch has its OWN position: prompts differ in length (text context carried between the windows of one call,
// initial_prompt), and a decode slot is refilled with the next window as soon as its sequence ends.
struct __align__(16) SeqState {
    int n_tok;        // tokens sampled so far in this window (tokens_cur.size()) = index of the token sampled next
    int last, prev;   // last / penultimate sampled token
    int has_ts;
    int seek_delta;
    int result_len;
    int failed;
    int done;         // 1: finished (or an empty slot): attention kernels skip it, the sampler leaves it alone
    int seek, seek_end;
    float sum_logprob;
    int pos;          // KV-cache position of the token fed in the current step (n_past)
    int idx;          // prompt index of the token fed in the current step; sampling starts at idx == n_prompt - 1
    int n_prompt;
    int lang_slot;    // prompt index of the language token (holds -1 until k_lang_detect fills it), or -1
    int restart;      // 1: prompt[0] is an extra [sot] fed at position 0 only to detect the language (whisper_lang_auto_detect
                      //    decodes [sot] alone); the real prompt then starts again at position 0
};   // 64 bytes, 16-byte aligned array elements (read with four 128-bit loads)

struct SamplerArgs {
    SeqState* state;          // [B]
    int* tokens_out;          // [B][n_max]
    float* margins_out;       // [B][n_max] top1 - top2 of the filtered logits, or null
    int* tids_out;            // [B][n_max] whisper_token_data.tid: most probable timestamp token of the step (0: none)
    int* next_tokens;         // [B] input of the next decoder step
    const int* forced;        // [B][n_max] teacher-forced tokens (<0 = free) or null
    int* tick;                // device: steps this lane has run (launch trace index), or null
    SpecialIds sp;
    int n_vocab;
    int n_max;
    int n_text_ctx;
    int suppress_blank;
    int no_timestamps;
    int single_segment;
    int max_initial_tid;      // round(max_initial_ts / 0.02), < 0 disables the rule
    int* prompt;              // device: prompt tokens [B][kMaxPrompt]
    float* plogs_out;         // [B][n_max] log-probability of every sampled token (whisper_token_data.plog)
    // temperature fallback (whisper_full's temperature schedule): temperature[b] > 0 -> the logits are divided by it and the
    // token is DRAWN from the filtered distribution with the uniform number rng_u[b][step] (inverse CDF, like
    // std::discrete_distribution over std::mt19937); null / 0 -> greedy
    const float* temperature; // [B] or null
    const double* rng_u;      // [B][n_max] or null
    // suppress_nst: ids of the non-speech tokens (device) that get -inf before every other rule, or null
    const int* nst_ids;
    int n_nst;
};

struct SkinnyEpilogue {
    const float* bias = nullptr;
    int act = 0;
    const float* residual = nullptr; int ldr = 0;
    float* out32 = nullptr; int ldo32 = 0;
    void* out16 = nullptr; int ldo16 = 0;
    TraceSlot trace;       // filled by the launcher from g_trace_next
};

// conv1 im2col: window w reads clip clip_of[w] starting at mel frame seek[w]
struct Im2col1Args {
    const float* mel;          // [n_clips][n_mel][mel_stride]
    const float* floor_val;    // [n_clips]; used when clip_max == null (mel already normalised)
    const int* clip_of;        // [n_windows]
    const int* seek;           // [n_windows]
    const int* n_calc;         // [n_clips]
    const int* n_len;          // [n_clips]
    int64_t mel_clip_stride;
    int mel_stride;
    int n_mel;
    int n_frames;              // 3000
    const int32_t* clip_max = nullptr;   // [n_clips] or null.  Not null: `mel` holds RAW log10 values and the kernel applies whisper.cpp's
                               // normalisation itself: max(v, clip maximum - 8), then (v + 4) / 4; the floor follows from it
};

struct GemmEpilogue {
    void* out;            // [M, ldo] T or f32
    int ldo;
    int out_f32;          // 1: f32 output, 0: 16-bit output
    const float* bias;    // [N] or null
    int act;              // 0 none, 1 tanh-GELU
    const float* residual;  // f32 [*, ldr] or null; added after activation
    int ldr;
    int res_row_mod;      // 0: residual row = row; >0: row % res_row_mod (positional embedding)
};

int gemm_tn(int dtype, const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
            const GemmEpilogue& ep, cudaStream_t st);
int num_sms();

// encoder-side launchers (encoder_kernels.cu, gemm_tcgen05.cu)
template <typename T> int im2col_conv1(const Im2col1Args& a, T* out, int n_windows, cudaStream_t st);
template <typename T> int im2col_conv2(const T* in, T* out, int n_windows, int n_in, int n_out, int d, cudaStream_t st);
template <typename T> int layernorm(const float* x, const float* g, const float* b, T* out16, float* out32, int rows, int d, cudaStream_t st);
template <typename T> int attn_enc_tc(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st);   // tcgen05 (k_attn_enc_ts)

// decoder-side launchers (decoder_kernels.cu)
template <typename T> int skinny_gemm(const T* X, int ldx, const T* W, int ldw, int Bn, int N, int K, const SkinnyEpilogue& ep, cudaStream_t st);
template <typename T> int dec_ln(float* x, const float* gamma, const float* beta, T* out16, int rows, int d, const T* tok_emb, const float* pos_emb, const int* next_tokens, const SeqState* state, cudaStream_t st);
// honor_done = 0: teacher-forced traces keep every sequence alive
// row_slot / row_pos != null: prompt-prefill mode -- "sequence" b is prompt row b = (decode slot row_slot[b], position
// row_pos[b]); plain launches (no PDL), nothing appended to the cache, `state` unused
template <typename T> int dec_self_attn(const T* qkv, T* kc, T* vc, T* out, const SeqState* state, int honor_done, int Bn, int n_head, int d, int n_text_ctx, cudaStream_t st, const int* row_slot = nullptr, const int* row_pos = nullptr);
template <typename T> int dec_cross_attn(const T* q, int ldq, const T* kbase, const T* vbase, int64_t ld_kv, int64_t win_stride, T* out, const SeqState* state, int Bn, int n_head, int d, int n_ctx, cudaStream_t st, const int* row_slot = nullptr);
template <typename T> int prefill_embed(const T* tok_emb, const float* pos_emb, const int* row_tok, const int* row_pos, float* x, int rows, int d, cudaStream_t st);
template <typename T> int prefill_kv_scatter(const T* qkv, const int* row_slot, const int* row_pos, T* kc, T* vc, int rows, int d, int n_text_ctx, cudaStream_t st);
int sample_step(const float* logits, int ld, const SamplerArgs& a, int Bn, cudaStream_t st);
// whisper_lang_auto_detect: at prompt index 0 ([sot] alone at position 0) pick the language token with the largest logit
// and write it into the language slot of every sequence whose slot holds the sentinel -1 (no-op for every other sequence)
int lang_detect_step(const float* logits, int ld, int* prompt, const SeqState* state, int* lang_out, SpecialIds sp, int Bn, cudaStream_t st);
// scatter freshly assigned windows into their decode slots (state, first token, prompt); items: device copy of SlotInit[n]
struct SlotInit { int slot; int next_token; float temperature; int pad_; SeqState state; int prompt[kMaxPrompt]; };
int slot_init(const SlotInit* items, int n, SeqState* state, int* next_tokens, int* prompt, int* lang_out, float* temperature, cudaStream_t st);

}  // namespace sb
