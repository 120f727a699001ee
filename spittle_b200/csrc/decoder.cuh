// Shared declarations of the decoder-step kernels and the encoder-side launchers.
#pragma once
#include "common.cuh"

namespace sb {

struct SpecialIds {
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang_first, num_languages, blank;
};

// per-sequence decode state, device resident (whisper_full's per-decoder bookkeeping, App. C.4)
struct __align__(16) SeqState {
    int n_tok;        // tokens sampled so far in this window (tokens_cur.size())
    int last, prev;   // last / penultimate sampled token
    int has_ts;
    int seek_delta;
    int result_len;
    int failed;
    int done;
    int seek, seek_end;
    float sum_logprob;
    int pad_;
};   // 48 bytes, 16-byte aligned array elements (read with three 128-bit loads)

struct SamplerArgs {
    SeqState* state;          // [B]
    const int* step_ptr;      // device: index of the token being sampled
    int* tokens_out;          // [B][n_max]
    float* margins_out;       // [B][n_max] top1 - top2 of the filtered logits, or null
    int* next_tokens;         // [B] input of the next decoder step
    const int* forced;        // [B][n_max] teacher-forced tokens (<0 = free) or null
    int* n_done;              // device counter of finished sequences
    SpecialIds sp;
    int n_vocab;
    int n_max;
    int suppress_blank;
    int no_timestamps;
    int single_segment;
    int max_initial_tid;      // round(max_initial_ts / 0.02), < 0 disables the rule
    const int* pos_ptr;       // device: position of the token just fed to the decoder
    const int* prompt;        // device: prompt tokens [B][n_prompt] (per sequence: the language token may differ)
    int n_prompt;
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
    const float* floor_val;    // [n_clips]
    const int* clip_of;        // [n_windows]
    const int* seek;           // [n_windows]
    const int* n_calc;         // [n_clips]
    const int* n_len;          // [n_clips]
    int64_t mel_clip_stride;
    int mel_stride;
    int n_mel;
    int n_frames;              // 3000
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
template <typename T> int attn_enc(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st);   // mma.sync version
template <typename T> int attn_enc_tc(const T* qkv, T* out, int n_windows, int n_ctx, int d_model, int n_head, cudaStream_t st);  // tcgen05 version
bool use_tc_attention();   // env SB_ATTN=mma selects the legacy mma.sync kernel

// decoder-side launchers (decoder_kernels.cu)
template <typename T> int skinny_gemm(const T* X, int ldx, const T* W, int ldw, int Bn, int N, int K, const SkinnyEpilogue& ep, cudaStream_t st);
template <typename T> int dec_ln(float* x, const float* gamma, const float* beta, T* out16, int rows, int d, const T* tok_emb, const float* pos_emb, const int* next_tokens, const int* pos_ptr, cudaStream_t st);
template <typename T> int dec_self_attn(const T* qkv, T* kc, T* vc, T* out, const int* pos_ptr, const SeqState* state, int Bn, int n_head, int d, int n_text_ctx, cudaStream_t st);
template <typename T> int dec_cross_attn(const T* q, int ldq, const T* kbase, const T* vbase, int64_t ld_kv, int64_t win_stride, T* out, const SeqState* state, int Bn, int n_head, int d, int n_ctx, cudaStream_t st);
int sample_step(const float* logits, int ld, const SamplerArgs& a, int Bn, cudaStream_t st);
// whisper_lang_auto_detect: at decode position 0 ([sot] only) pick the language token with the largest logit and
// write it into prompt slot 1 of every sequence whose slot holds the sentinel -1 (no-op at any other position)
int lang_detect_step(const float* logits, int ld, int* prompt, int n_prompt, const int* pos_ptr, int* lang_out, SpecialIds sp, int Bn, cudaStream_t st);
int dec_advance(int* pos_ptr, int* step_ptr, int n_prompt, cudaStream_t st);

}  // namespace sb
