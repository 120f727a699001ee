// Engine: model upload, batched encode, batched greedy decode, the whisper_full seek loop.
//
// Native equivalent of transcribe-rs' WhisperEngine as used by the reference
// (src-tauri/src/managers/transcription.rs:262-263 load_model, :494-503 transcribe_samples,
// :183-189 unload_model) and of whisper.cpp's whisper_full_with_state.
#include "common.cuh"
#include <cstdio>
#include "decoder.cuh"
#include "ggml_loader.h"
#include <algorithm>
#include <memory>
#include <mutex>
#include <regex>
#include <unordered_map>
#include <vector>

struct sb_melplan;
namespace sb {
extern std::atomic<uint64_t> g_launches;
int logmel_launch(const sb_melplan* plan, const float* pcm, int n_clips, size_t n_samples, int64_t pcm_clip_stride,
                  float* mel, int64_t mel_clip_stride, int mel_stride, int32_t* clip_max, float* floor_val,
                  cudaStream_t st);

static const char* kLangs[] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi",
    "vi", "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml",
    "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs",
    "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am",
    "yi", "lb", "my", "bo", "tl", "mt", "sa", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "ba", "as", "tt", "ln", "ha",
    "mg", "jw", "su", "haw", "yue"};

static int lang_id(const char* s) {
    for (int i = 0; i < (int)(sizeof(kLangs) / sizeof(kLangs[0])); ++i)
        if (strcmp(kLangs[i], s) == 0) return i;
    return -1;
}

// ---- small device buffer helper -----------------------------------------------------------
static std::atomic<int> g_ws_gen{0};   // bumped on every (re)allocation: invalidates captured graphs
struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return SB_OK;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc(") + std::to_string(need) + "): " + cudaGetErrorString(e)); return SB_ERR_NOMEM; }
        bytes = need;
        g_ws_gen++;
        return SB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

__global__ void k_fill_i32(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
template <typename T>
__global__ void k_widen(const T* in, float* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = Op16<T>::to_f32(in[i]);
}

struct LnW { float* g = nullptr; float* b = nullptr; };
template <typename T> struct LinW { T* w = nullptr; float* b = nullptr; };
template <typename T> struct EncLayer { LnW ln1, ln2; LinW<T> qkv, o, fc1, fc2; };
template <typename T> struct DecLayer { LnW ln1, ln2, ln3; LinW<T> qkv, o, cq, co, fc1, fc2; };

struct EngineBase {
    virtual ~EngineBase() {}
    WhisperHParams hp{};
    SpecialIds sp{};
    std::vector<std::string> vocab;
    int device = 0, max_batch = 64, dtype = 0, use_graph = 1;
    sb_stats stats{};
    int profile = 0;
    virtual void* stream_handle() = 0;
    virtual int transcribe_batch(const float* const* pcm, const size_t* ns, size_t count, const sb_params& p, sb_result* out) = 0;
    virtual int encode_host(const float* mel_windows, int n_windows, float* enc_out) = 0;
    virtual int decode_trace(const float* mel_windows, int n_windows, const int32_t* seek_end, const sb_params& p,
                             const int32_t* forced, int n_steps, float* logits_out, int32_t* tokens_out,
                             float* margins_out) = 0;
    // whisper.cpp `tokenize` (whisper_tokenize; used for initial_prompt) [MEM]: GPT-2 style word split by regex over the
    // bytes of the text, then each word is cut greedily into the LONGEST vocabulary entries; bytes no entry covers are
    // skipped.  token_to_id keeps the last id of a duplicated entry, like the loader's `token_to_id[word] = i`.
    mutable std::unordered_map<std::string, int> token_to_id;
    mutable std::once_flag token_to_id_once;
    std::vector<int> tokenize(const std::string& text) const {
        std::call_once(token_to_id_once, [this] {
            for (int i = 0; i < (int)vocab.size(); ++i) token_to_id[vocab[i]] = i;
        });
        static const std::regex re(R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)");
        std::vector<int> out;
        for (std::sregex_iterator it(text.begin(), text.end(), re), end; it != end; ++it) {
            const std::string word = it->str();
            const int n = (int)word.size();
            int i = 0;
            while (i < n) {
                int j = n;
                bool found = false;
                while (j > i) {
                    auto f = token_to_id.find(word.substr(i, j - i));
                    if (f != token_to_id.end()) { out.push_back(f->second); i = j; found = true; break; }
                    --j;
                }
                if (!found) ++i;          // whisper.cpp logs "unknown token" and moves on
            }
        }
        return out;
    }
    std::string token_text(int id) const {
        if (id >= 0 && id < (int)vocab.size()) return vocab[id];
        if (id == sp.eot) return "[_EOT_]";
        if (id == sp.sot) return "[_SOT_]";
        if (id == sp.translate) return "[_TRANSLATE_]";
        if (id == sp.transcribe) return "[_TRANSCRIBE_]";
        if (id == sp.solm) return "[_SOLM_]";
        if (id == sp.prev) return "[_PREV_]";
        if (id == sp.nosp) return "[_NOSP_]";
        if (id == sp.not_) return "[_NOT_]";
        if (id == sp.beg) return "[_BEG_]";
        if (id > sp.beg) return "[_TT_" + std::to_string(id - sp.beg) + "]";
        if (id >= sp.lang_first && id < sp.lang_first + sp.num_languages) return "[_LANG_" + std::to_string(id - sp.lang_first) + "]";
        return "[_extra_token_" + std::to_string(id) + "]";
    }
};

static SpecialIds special_from_vocab(int n_vocab, const std::vector<std::string>& vocab) {
    SpecialIds s{50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362, 50363, 0, 0, 220};
    if (n_vocab >= 51865) {
        s.num_languages = n_vocab - 51765 - 1;
        s.eot++; s.sot++;
        const int dt = s.num_languages - 98;
        s.translate += dt; s.transcribe += dt; s.solm += dt; s.prev += dt; s.nosp += dt; s.not_ += dt; s.beg += dt;
    }
    s.lang_first = s.sot + 1;
    for (size_t i = 0; i < vocab.size(); ++i)
        if (vocab[i] == " ") { s.blank = (int)i; break; }
    return s;
}

template <typename T>
struct Engine : EngineBase {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    sb_melplan* melplan = nullptr;
    std::vector<void*> owned;   // weight allocations

    // weights
    T* conv1_w = nullptr; float* conv1_b = nullptr;
    T* conv2_w = nullptr; float* conv2_b = nullptr;
    float* enc_pos = nullptr;
    std::vector<EncLayer<T>> enc;
    LnW ln_post;
    T* tok_emb = nullptr; float* dec_pos = nullptr;
    LinW<T> cross_kv;
    std::vector<DecLayer<T>> dec;
    LnW ln_f;

    // workspaces
    DevBuf b_pcm, b_mel, b_cmax, b_floor, b_clipmeta, b_winmeta;
    DevBuf b_col1, b_c1, b_x, b_h, b_qkv, b_att, b_mlp, b_enc32;
    DevBuf b_ckv, b_kself, b_vself, b_dx, b_dh, b_dqkv, b_datt, b_dq, b_dmlp, b_logits;
    DevBuf b_state, b_tokens, b_margins, b_next, b_forced, b_ctr, b_prompt;
    int* h_ctr = nullptr;   // pinned: [pos, step, n_done]
    // profile == 2: device-side launch trace of the decoder step (TraceSlot, common.cuh)
    DevBuf b_trace;
    std::vector<int> trace_cls;          // class of launch idx inside a step: 0 projection, 1 LayerNorm, 2 self-attn, 3 cross-attn
    std::vector<double> trace_work;      // algorithmic bytes of launch idx (projections: weight bytes)
    int trace_per_step = 0, trace_max_steps = 0;
    std::vector<double> trace_sum_dur, trace_sum_t0, trace_sum_t1, trace_cnt;   // per launch index, for SB_TRACE_DUMP

    // per-launch CUDA-event brackets (only when profile != 0): class 0 = tcgen05 GEMM, 1 = encoder attention
    struct ProfRec { cudaEvent_t a, b; int cls; double work; };
    std::vector<ProfRec> prof_pool; size_t prof_used = 0;
    // (profile == 2 adds the device-side launch trace of the decoder step, see b_trace)
    int prof_begin(int cls, double work) { return prof_begin_on(cls, work, st); }
    int prof_end() { return prof_end_on(st); }
    int prof_begin_on(int cls, double work, cudaStream_t s_) {
        if (!profile) return SB_OK;
        if (prof_used == prof_pool.size()) {
            ProfRec r{}; r.cls = cls;
            SB_CUDA_CHECK(cudaEventCreate(&r.a)); SB_CUDA_CHECK(cudaEventCreate(&r.b));
            prof_pool.push_back(r);
        }
        prof_pool[prof_used].cls = cls; prof_pool[prof_used].work = work;
        SB_CUDA_CHECK(cudaEventRecord(prof_pool[prof_used].a, s_));
        return SB_OK;
    }
    int prof_end_on(cudaStream_t s_) {
        if (!profile) return SB_OK;
        SB_CUDA_CHECK(cudaEventRecord(prof_pool[prof_used].b, s_));
        ++prof_used;
        return SB_OK;
    }
    void prof_collect() {   // stream must be idle
        for (size_t i = 0; i < prof_used; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, prof_pool[i].a, prof_pool[i].b);
            if (prof_pool[i].cls == 0) { stats.gemm_ms += ms; stats.gemm_flops += prof_pool[i].work; stats.gemm_launches += 1; }
            else { stats.attn_ms += ms; stats.attn_flops += prof_pool[i].work; stats.attn_launches += 1; }
        }
        prof_used = 0;
    }
    int gemm_p(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, const GemmEpilogue& ep) {
        int rc = prof_begin(0, 2.0 * M * N * K);
        if (rc) return rc;
        if ((rc = gemm_tn(dtype, A, lda, W, ldw, M, N, K, ep, st))) return rc;
        return prof_end();
    }

    // The decoder step is a chain of ~140 small, latency-bound launches plus one HBM-bound stream
    // (cross-attention).  The batch is therefore cut into up to kMaxLanes independent sub-batches
    // ("lanes"), each with its own CUDA stream, counters and captured step graph, so the chains of
    // different lanes overlap and the cross-attention of one lane streams while the others wait on latency.
    static constexpr int kMaxLanes = 4;
    struct GraphKey {
        int W, w0, Wl, n_max, flags, max_init, n_prompt, has_forced, gen;
        bool operator==(const GraphKey& o) const {
            return W == o.W && w0 == o.w0 && Wl == o.Wl && n_max == o.n_max && flags == o.flags && max_init == o.max_init &&
                   n_prompt == o.n_prompt && has_forced == o.has_forced && gen == o.gen;
        }
    };
    struct Lane {
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        cudaGraphExec_t gexec = nullptr;
        GraphKey key{0, 0, 0, 0, 0, 0, 0, 0, -1};
        int graph_nodes = 0;
    };
    Lane lanes[kMaxLanes];
    cudaEvent_t ev_fork = nullptr;
    int n_lanes_cfg = 2;

    ~Engine() override {
        for (auto& l : lanes) {
            if (l.gexec) cudaGraphExecDestroy(l.gexec);
            if (l.done) cudaEventDestroy(l.done);
            if (l.st) cudaStreamDestroy(l.st);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& r : prof_pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (void* p : owned) cudaFree(p);
        DevBuf* bufs[] = {&b_pcm, &b_mel, &b_cmax, &b_floor, &b_clipmeta, &b_winmeta, &b_col1, &b_c1, &b_x, &b_h, &b_qkv,
                          &b_att, &b_mlp, &b_enc32, &b_ckv, &b_kself, &b_vself, &b_dx, &b_dh, &b_dqkv, &b_datt, &b_dq,
                          &b_dmlp, &b_logits, &b_state, &b_tokens, &b_margins, &b_next, &b_forced, &b_ctr, &b_prompt};
        for (DevBuf* b : bufs) b->release();
        if (melplan) sb_melplan_destroy(melplan);
        if (h_ctr) cudaFreeHost(h_ctr);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
    }

    void* stream_handle() override { return (void*)st; }

    // ---- upload helpers ----
    int dev_alloc(void** p, size_t bytes) {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc weights: ") + cudaGetErrorString(e)); return SB_ERR_NOMEM; }
        owned.push_back(*p);
        return SB_OK;
    }
    static void to_f32(const HostTensor& t, std::vector<float>& out) {
        const int64_t n = t.numel();
        out.resize(n);
        if (t.ttype == 0) memcpy(out.data(), t.data, n * 4);
        else { const uint16_t* h = reinterpret_cast<const uint16_t*>(t.data); for (int64_t i = 0; i < n; ++i) out[i] = f16_bits_to_f32(h[i]); }
    }
    static uint16_t f32_to_t(float f) {
        if (std::is_same<T, __half>::value) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
        __nv_bfloat16 b = __float2bfloat16_rn(f); uint16_t u; memcpy(&u, &b, 2); return u;
    }
    // rows of 16-bit data appended to a host staging vector
    static void append16(const HostTensor& t, std::vector<uint16_t>& dst) {
        const int64_t n = t.numel();
        const size_t o = dst.size();
        dst.resize(o + n);
        if (t.ttype == 1 && std::is_same<T, __half>::value) { memcpy(dst.data() + o, t.data, n * 2); return; }
        if (t.ttype == 1) { const uint16_t* h = reinterpret_cast<const uint16_t*>(t.data); for (int64_t i = 0; i < n; ++i) dst[o + i] = f32_to_t(f16_bits_to_f32(h[i])); return; }
        const float* f = reinterpret_cast<const float*>(t.data);
        for (int64_t i = 0; i < n; ++i) dst[o + i] = f32_to_t(f[i]);
    }
    int up16(const std::vector<uint16_t>& h, T** out) {
        int rc = dev_alloc((void**)out, h.size() * 2);
        if (rc != SB_OK) return rc;
        SB_CUDA_CHECK(cudaMemcpy(*out, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
        return SB_OK;
    }
    int up32(const std::vector<float>& h, float** out) {
        int rc = dev_alloc((void**)out, h.size() * 4);
        if (rc != SB_OK) return rc;
        SB_CUDA_CHECK(cudaMemcpy(*out, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        return SB_OK;
    }
    const HostTensor* find(const GgmlFile& f, const std::string& name, std::vector<int64_t> shape) {
        auto it = f.tensors.find(name);
        if (it == f.tensors.end()) { set_error("model file lacks tensor " + name); return nullptr; }
        if (it->second.numel() != [&] { int64_t n = 1; for (auto s : shape) n *= s; return n; }()) {
            set_error("tensor " + name + " has an unexpected shape"); return nullptr;
        }
        return &it->second;
    }
#define SB_FIND(var, file, name, ...)                                   \
    const HostTensor* var = find(file, name, {__VA_ARGS__});            \
    if (!var) return SB_ERR_FORMAT;

    int up_ln(const GgmlFile& f, const std::string& prefix, int d, LnW& out) {
        SB_FIND(g, f, prefix + ".weight", d);
        SB_FIND(b, f, prefix + ".bias", d);
        std::vector<float> v;
        to_f32(*g, v); int rc = up32(v, &out.g); if (rc) return rc;
        to_f32(*b, v); return up32(v, &out.b);
    }
    // concatenated linear layers: names[i] weight [n_i, k] (+ optional bias; absent -> zeros)
    int up_lin(const GgmlFile& f, const std::vector<std::string>& prefixes, const std::vector<int>& outs, int k, LinW<T>& out) {
        std::vector<uint16_t> w;
        std::vector<float> b;
        for (size_t i = 0; i < prefixes.size(); ++i) {
            SB_FIND(t, f, prefixes[i] + ".weight", outs[i], k);
            append16(*t, w);
            auto it = f.tensors.find(prefixes[i] + ".bias");
            if (it != f.tensors.end()) { std::vector<float> v; to_f32(it->second, v); b.insert(b.end(), v.begin(), v.end()); }
            else b.insert(b.end(), outs[i], 0.0f);
        }
        int rc = up16(w, &out.w); if (rc) return rc;
        return up32(b, &out.b);
    }
    int up_conv(const GgmlFile& f, const std::string& prefix, int co, int ci, T** w_out, float** b_out) {
        SB_FIND(t, f, prefix + ".weight", co, ci, 3);
        std::vector<float> v; to_f32(*t, v);
        std::vector<uint16_t> w((size_t)co * 3 * ci);
        for (int o = 0; o < co; ++o)
            for (int c = 0; c < ci; ++c)
                for (int k = 0; k < 3; ++k) w[((size_t)o * 3 + k) * ci + c] = f32_to_t(v[((size_t)o * ci + c) * 3 + k]);
        int rc = up16(w, w_out); if (rc) return rc;
        SB_FIND(bt, f, prefix + ".bias", co);
        to_f32(*bt, v);
        return up32(v, b_out);
    }

    int load(const GgmlFile& f) {
        hp = f.hp;
        vocab = f.vocab;
        sp = special_from_vocab(hp.n_vocab, vocab);
        const int d = hp.n_audio_state, dt = hp.n_text_state;
        SB_CHECK_ARG(d == dt, "n_audio_state != n_text_state is not supported");
        SB_CHECK_ARG(d % 64 == 0 && hp.n_audio_head * 64 == d && hp.n_text_head * 64 == dt, "d_head must be 64");
        SB_CHECK_ARG(hp.n_audio_ctx == 1500, "n_audio_ctx must be 1500");
        SB_CHECK_ARG(hp.n_text_ctx <= 448, "n_text_ctx must be <= 448");
        SB_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (auto& e : ev) SB_CUDA_CHECK(cudaEventCreate(&e));
        SB_CUDA_CHECK(cudaMallocHost(&h_ctr, 16 * kMaxLanes * sizeof(int)));
        for (auto& l : lanes) {
            SB_CUDA_CHECK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
            SB_CUDA_CHECK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        }
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        if (const char* e = getenv("SB_DECODE_LANES")) { n_lanes_cfg = atoi(e); if (n_lanes_cfg < 1) n_lanes_cfg = 1; if (n_lanes_cfg > kMaxLanes) n_lanes_cfg = kMaxLanes; }
        int rc = sb_melplan_create(f.mel_filters.data(), hp.n_mels, &melplan);
        if (rc) return rc;
        rc = up_conv(f, "encoder.conv1", d, hp.n_mels, &conv1_w, &conv1_b); if (rc) return rc;
        rc = up_conv(f, "encoder.conv2", d, d, &conv2_w, &conv2_b); if (rc) return rc;
        {
            SB_FIND(t, f, "encoder.positional_embedding", hp.n_audio_ctx, d);
            std::vector<float> v; to_f32(*t, v); rc = up32(v, &enc_pos); if (rc) return rc;
        }
        enc.resize(hp.n_audio_layer);
        for (int i = 0; i < hp.n_audio_layer; ++i) {
            const std::string p = "encoder.blocks." + std::to_string(i);
            EncLayer<T>& L = enc[i];
            if ((rc = up_ln(f, p + ".attn_ln", d, L.ln1))) return rc;
            if ((rc = up_lin(f, {p + ".attn.query", p + ".attn.key", p + ".attn.value"}, {d, d, d}, d, L.qkv))) return rc;
            if ((rc = up_lin(f, {p + ".attn.out"}, {d}, d, L.o))) return rc;
            if ((rc = up_ln(f, p + ".mlp_ln", d, L.ln2))) return rc;
            if ((rc = up_lin(f, {p + ".mlp.0"}, {4 * d}, d, L.fc1))) return rc;
            if ((rc = up_lin(f, {p + ".mlp.2"}, {d}, 4 * d, L.fc2))) return rc;
        }
        if ((rc = up_ln(f, "encoder.ln_post", d, ln_post))) return rc;
        {
            SB_FIND(t, f, "decoder.token_embedding.weight", hp.n_vocab, dt);
            std::vector<uint16_t> w; append16(*t, w);
            w.resize((size_t)round_up(hp.n_vocab, 8) * dt, 0);      // zero rows: the logits GEMM runs on N rounded up to 8
            rc = up16(w, &tok_emb); if (rc) return rc;
            SB_FIND(pe, f, "decoder.positional_embedding", hp.n_text_ctx, dt);
            std::vector<float> v; to_f32(*pe, v); rc = up32(v, &dec_pos); if (rc) return rc;
        }
        dec.resize(hp.n_text_layer);
        std::vector<std::string> ckv_names; std::vector<int> ckv_outs;
        for (int i = 0; i < hp.n_text_layer; ++i) {
            const std::string p = "decoder.blocks." + std::to_string(i);
            DecLayer<T>& L = dec[i];
            if ((rc = up_ln(f, p + ".attn_ln", dt, L.ln1))) return rc;
            if ((rc = up_lin(f, {p + ".attn.query", p + ".attn.key", p + ".attn.value"}, {dt, dt, dt}, dt, L.qkv))) return rc;
            if ((rc = up_lin(f, {p + ".attn.out"}, {dt}, dt, L.o))) return rc;
            if ((rc = up_ln(f, p + ".cross_attn_ln", dt, L.ln2))) return rc;
            if ((rc = up_lin(f, {p + ".cross_attn.query"}, {dt}, dt, L.cq))) return rc;
            if ((rc = up_lin(f, {p + ".cross_attn.out"}, {dt}, dt, L.co))) return rc;
            if ((rc = up_ln(f, p + ".mlp_ln", dt, L.ln3))) return rc;
            if ((rc = up_lin(f, {p + ".mlp.0"}, {4 * dt}, dt, L.fc1))) return rc;
            if ((rc = up_lin(f, {p + ".mlp.2"}, {dt}, 4 * dt, L.fc2))) return rc;
            ckv_names.push_back(p + ".cross_attn.key"); ckv_outs.push_back(dt);
            ckv_names.push_back(p + ".cross_attn.value"); ckv_outs.push_back(dt);
        }
        if ((rc = up_lin(f, ckv_names, ckv_outs, dt, cross_kv))) return rc;
        if ((rc = up_ln(f, "decoder.ln", dt, ln_f))) return rc;
        return SB_OK;
    }

    // ---- encoder over `W` windows whose im2col rows are already in b_col1 -------------------
    // produces cross-KV rows [w0*1500 .. (w0+W)*1500) of b_ckv and (optionally) f32 encoder output
    int encode_chunk(int W, int w0, float* enc32_out) {
        const int d = hp.n_audio_state, nctx = hp.n_audio_ctx, nfr = 2 * nctx;
        const int M2 = W * nfr, M = W * nctx;
        int rc;
        GemmEpilogue ep{};
        // conv1 + GELU -> c1 [M2, d]
        ep = GemmEpilogue{b_c1.p, d, 0, conv1_b, 1, nullptr, 0, 0};
        if ((rc = gemm_p(b_col1.p, 3 * hp.n_mels, conv1_w, 3 * hp.n_mels, M2, d, 3 * hp.n_mels, ep))) return rc;
        // conv2 (stride 2) + GELU + positional embedding -> x f32 [M, d]
        if ((rc = im2col_conv2<T>(b_c1.as<T>(), b_mlp.as<T>(), W, nfr, nctx, d, st))) return rc;
        ep = GemmEpilogue{b_x.p, d, 1, conv2_b, 1, enc_pos, d, nctx};
        if ((rc = gemm_p(b_mlp.p, 3 * d, conv2_w, 3 * d, M, d, 3 * d, ep))) return rc;
        for (int l = 0; l < hp.n_audio_layer; ++l) {
            const EncLayer<T>& L = enc[l];
            if ((rc = layernorm<T>(b_x.as<float>(), L.ln1.g, L.ln1.b, b_h.as<T>(), nullptr, M, d, st))) return rc;
            ep = GemmEpilogue{b_qkv.p, 3 * d, 0, L.qkv.b, 0, nullptr, 0, 0};
            if ((rc = gemm_p(b_h.p, d, L.qkv.w, d, M, 3 * d, d, ep))) return rc;
            if ((rc = prof_begin(1, 4.0 * W * (double)nctx * nctx * d))) return rc;
            if (use_tc_attention()) { if ((rc = attn_enc_tc<T>(b_qkv.as<T>(), b_att.as<T>(), W, nctx, d, hp.n_audio_head, st))) return rc; }
            else if ((rc = attn_enc<T>(b_qkv.as<T>(), b_att.as<T>(), W, nctx, d, hp.n_audio_head, st))) return rc;
            if ((rc = prof_end())) return rc;
            ep = GemmEpilogue{b_x.p, d, 1, L.o.b, 0, b_x.as<float>(), d, 0};
            if ((rc = gemm_p(b_att.p, d, L.o.w, d, M, d, d, ep))) return rc;
            if ((rc = layernorm<T>(b_x.as<float>(), L.ln2.g, L.ln2.b, b_h.as<T>(), nullptr, M, d, st))) return rc;
            ep = GemmEpilogue{b_mlp.p, 4 * d, 0, L.fc1.b, 1, nullptr, 0, 0};
            if ((rc = gemm_p(b_h.p, d, L.fc1.w, d, M, 4 * d, d, ep))) return rc;
            ep = GemmEpilogue{b_x.p, d, 1, L.fc2.b, 0, b_x.as<float>(), d, 0};
            if ((rc = gemm_p(b_mlp.p, 4 * d, L.fc2.w, 4 * d, M, d, 4 * d, ep))) return rc;
        }
        if ((rc = layernorm<T>(b_x.as<float>(), ln_post.g, ln_post.b, b_h.as<T>(), enc32_out, M, d, st))) return rc;
        // cross-KV for every decoder layer in one GEMM: [M, Ld*2*d]
        const int nkv = hp.n_text_layer * 2 * d;
        ep = GemmEpilogue{b_ckv.as<T>() + (int64_t)w0 * nctx * nkv, nkv, 0, cross_kv.b, 0, nullptr, 0, 0};
        if ((rc = gemm_p(b_h.p, d, cross_kv.w, d, M, nkv, d, ep))) return rc;
        return SB_OK;
    }

    int ensure_encoder_ws(int Wc, int Wtot, bool want32) {
        const size_t d = hp.n_audio_state, nctx = hp.n_audio_ctx;
        const size_t M = (size_t)Wc * nctx, M2 = 2 * M;
        int rc;
        if ((rc = b_col1.ensure(M2 * 3 * hp.n_mels * 2))) return rc;
        if ((rc = b_c1.ensure(M2 * d * 2))) return rc;
        if ((rc = b_x.ensure(M * d * 4))) return rc;
        if ((rc = b_h.ensure(M * d * 2))) return rc;
        if ((rc = b_qkv.ensure(M * 3 * d * 2))) return rc;
        if ((rc = b_att.ensure(M * d * 2))) return rc;
        if ((rc = b_mlp.ensure(M * 4 * d * 2))) return rc;
        if (want32 && (rc = b_enc32.ensure(M * d * 4))) return rc;
        if ((rc = b_ckv.ensure((size_t)Wtot * nctx * hp.n_text_layer * 2 * d * 2))) return rc;
        return SB_OK;
    }

    int ensure_decoder_ws(int W, int n_max) {
        const size_t d = hp.n_text_state;
        int rc;
        const size_t kvb = (size_t)hp.n_text_layer * W * hp.n_text_ctx * d * 2;
        if ((rc = b_kself.ensure(kvb))) return rc;
        if ((rc = b_vself.ensure(kvb))) return rc;
        if ((rc = b_dx.ensure(W * d * 4))) return rc;
        if ((rc = b_dh.ensure(W * d * 2))) return rc;
        if ((rc = b_dqkv.ensure(W * 3 * d * 2))) return rc;
        if ((rc = b_datt.ensure(W * d * 2))) return rc;
        if ((rc = b_dq.ensure(W * d * 2))) return rc;
        if ((rc = b_dmlp.ensure(W * 4 * d * 2))) return rc;
        if ((rc = b_logits.ensure((size_t)W * round_up(hp.n_vocab, 8) * 4))) return rc;
        if ((rc = b_state.ensure(W * sizeof(SeqState)))) return rc;
        if ((rc = b_tokens.ensure((size_t)W * n_max * 4))) return rc;
        if ((rc = b_margins.ensure((size_t)W * n_max * 4))) return rc;
        if ((rc = b_forced.ensure((size_t)W * n_max * 4))) return rc;
        if ((rc = b_next.ensure(W * 4))) return rc;
        if ((rc = b_ctr.ensure(64 * kMaxLanes))) return rc;
        if ((rc = b_prompt.ensure(64))) return rc;
        return SB_OK;
    }

    // ---- one decoder step for sequences [w0, w0 + Wl) of a W-sequence batch, all launches on `sl` ----
    int enqueue_step(int W, int w0, int Wl, int* ctr, cudaStream_t sl, const SamplerArgs& sa) {
        const int d = hp.n_text_state, nctx = hp.n_audio_ctx;
        const int nkv = hp.n_text_layer * 2 * d;
        int* pos_ptr = ctr;
        int* step_ptr = ctr + 1;
        float* dx = b_dx.as<float>() + (int64_t)w0 * d;
        T* dh = b_dh.as<T>() + (int64_t)w0 * d;
        T* dqkv = b_dqkv.as<T>() + (int64_t)w0 * 3 * d;
        T* datt = b_datt.as<T>() + (int64_t)w0 * d;
        T* dq = b_dq.as<T>() + (int64_t)w0 * d;
        T* dmlp = b_dmlp.as<T>() + (int64_t)w0 * 4 * d;
        const int vpad = (int)round_up(hp.n_vocab, 8);
        float* logits = b_logits.as<float>() + (int64_t)w0 * vpad;
        // teacher-forced traces keep every sequence alive (their `done` flag only marks where free-running would stop)
        const SeqState* seq_state = sa.forced ? nullptr : sa.state;
        int rc;
        // PDL chain, one launch per stage (what was measured against it and lost -- a persistent per-step megakernel,
        // split-K / cluster projections, LayerNorm fused into the projections -- is in profiles/r1_mega_stage_trace.md)
        // profile == 2: every stage launch of the step gets a trace slot (class, algorithmic bytes); the step keeps its graph
        const bool tracing = profile == 2 && trace_per_step > 0;
        const int lane_idx = (int)((ctr - b_ctr.as<int>()) / 16);
        int tidx = 0;
        auto tr = [&](int cls, double work) {
            if (!tracing) return;
            if ((int)trace_cls.size() <= tidx) { trace_cls.resize(tidx + 1); trace_work.resize(tidx + 1); }
            trace_cls[tidx] = cls; trace_work[tidx] = work;
            const size_t n = (size_t)trace_max_steps * trace_per_step;
            TraceSlot ts;
            ts.t0 = b_trace.as<unsigned long long>() + (size_t)lane_idx * 2 * n;
            ts.t1 = ts.t0 + n;
            ts.pos = pos_ptr; ts.idx = tidx; ts.per_step = trace_per_step;
            g_trace_next = ts;
            ++tidx;
        };
        auto sk = [&](const T* X, int ldx, const T* Wt, int ldw, int N, int K, const SkinnyEpilogue& ep) -> int {
            tr(0, 2.0 * N * K);
            return skinny_gemm<T>(X, ldx, Wt, ldw, Wl, N, K, ep, sl);
        };
        const int* next_tok = b_next.as<int>() + w0;
        for (int l = 0; l < hp.n_text_layer; ++l) {
            const DecLayer<T>& L = dec[l];
            T* kc = b_kself.as<T>() + ((int64_t)l * W + w0) * hp.n_text_ctx * d;
            T* vc = b_vself.as<T>() + ((int64_t)l * W + w0) * hp.n_text_ctx * d;
            SkinnyEpilogue e{};
            // attn_ln; layer 0 forms x = token_embedding[tok] + positional_embedding[pos] first
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln1.g, L.ln1.b, dh, Wl, d, l == 0 ? tok_emb : nullptr, dec_pos, next_tok, pos_ptr, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.qkv.b; e.out16 = dqkv; e.ldo16 = 3 * d;
            if ((rc = sk(dh, d, L.qkv.w, d, 3 * d, d, e))) return rc;
            tr(2, 0);
            if ((rc = dec_self_attn<T>(dqkv, kc, vc, datt, pos_ptr, seq_state, Wl, hp.n_text_head, d, hp.n_text_ctx, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.o.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(datt, d, L.o.w, d, d, d, e))) return rc;
            const T* kb = b_ckv.as<T>() + (int64_t)w0 * nctx * nkv + (int64_t)l * 2 * d;
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln2.g, L.ln2.b, dh, Wl, d, nullptr, nullptr, nullptr, pos_ptr, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.cq.b; e.out16 = dq; e.ldo16 = d;
            if ((rc = sk(dh, d, L.cq.w, d, d, d, e))) return rc;
            tr(3, 0);
            if ((rc = dec_cross_attn<T>(dq, d, kb, kb + d, nkv, (int64_t)nctx * nkv, datt, seq_state, Wl, hp.n_text_head, d, nctx, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.co.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(datt, d, L.co.w, d, d, d, e))) return rc;
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln3.g, L.ln3.b, dh, Wl, d, nullptr, nullptr, nullptr, pos_ptr, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.fc1.b; e.act = 1; e.out16 = dmlp; e.ldo16 = 4 * d;
            if ((rc = sk(dh, d, L.fc1.w, d, 4 * d, d, e))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.fc2.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(dmlp, 4 * d, L.fc2.w, 4 * d, d, 4 * d, e))) return rc;
        }
        tr(1, 0);
        if ((rc = dec_ln<T>(dx, ln_f.g, ln_f.b, dh, Wl, d, nullptr, nullptr, nullptr, pos_ptr, sl))) return rc;
        // tied-embedding logits: 80-130 MB of weights per step -> the TMA-fed tcgen05 GEMM streams them
        // (one 128-row tile of sequences, ~200 column tiles) instead of the small-N weight-streaming kernel
        GemmEpilogue ge{logits, vpad, 1, nullptr, 0, nullptr, 0, 0};
        if ((rc = gemm_tn(dtype, dh, d, tok_emb, d, Wl, vpad, d, ge, sl))) return rc;
        if (detect_lang && (rc = lang_detect_step(logits, vpad, const_cast<int*>(sa.prompt), sa.n_prompt, pos_ptr, detect_out + w0, sp, Wl, sl))) return rc;
        if ((rc = sample_step(logits, vpad, sa, Wl, sl))) return rc;
        if ((rc = dec_advance(pos_ptr, step_ptr, sa.n_prompt, sl))) return rc;
        return SB_OK;
    }

    struct DecodeOut { std::vector<SeqState> state; std::vector<int> tokens; std::vector<float> margins; std::vector<int> langs; int n_max = 0; };
    bool detect_lang = false; int* detect_out = nullptr;     // set by decode() for enqueue_step
    bool auto_mode = false;                                  // current transcribe_batch call asked for language auto-detect

    // decode W windows whose cross-KV occupies rows [0, W*1500) of b_ckv
    // langs[w]: language id of window w's prompt, or < 0: detect it at decode position 0 (reference default "auto")
    int decode(int W, const std::vector<int>& seek, const std::vector<int>& seek_end, const sb_params& p, const std::vector<int>& langs,
               const int32_t* forced_host, int n_steps_cap, float* logits_out, DecodeOut& out) {
        int n_max = hp.n_text_ctx / 2 - 4;
        if (p.n_max_tokens > 0) n_max = std::min(n_max, p.n_max_tokens);
        if (n_steps_cap > 0) n_max = std::min(n_max, n_steps_cap);
        out.n_max = n_max;
        int rc = ensure_decoder_ws(W, n_max);
        if (rc) return rc;
        std::vector<int> prompt = {sp.sot};
        if (hp.n_vocab >= 51865) { prompt.push_back(sp.lang_first); prompt.push_back(p.translate ? sp.translate : sp.transcribe); }
        if (p.no_timestamps) prompt.push_back(sp.not_);
        const int n_prompt = (int)prompt.size();
        bool detect = false;
        std::vector<int> prompts((size_t)W * n_prompt);
        for (int w = 0; w < W; ++w) {
            for (int i = 0; i < n_prompt; ++i) prompts[(size_t)w * n_prompt + i] = prompt[i];
            if (n_prompt >= 2) {
                if (langs[w] < 0) { detect = true; prompts[(size_t)w * n_prompt + 1] = -1; }     // sentinel: k_lang_detect fills it
                else prompts[(size_t)w * n_prompt + 1] = sp.lang_first + langs[w];
            }
        }
        if ((rc = b_prompt.ensure((size_t)W * n_prompt * 4 + (size_t)W * 4))) return rc;
        int* d_lang = b_prompt.as<int>() + (size_t)W * n_prompt;      // detected language ids [W]
        std::vector<SeqState> hs(W);
        for (int w = 0; w < W; ++w) {
            SeqState s{}; s.seek_delta = 3000; s.seek = seek[w]; s.seek_end = seek_end[w];
            hs[w] = s;
        }
        SB_CUDA_CHECK(cudaMemcpyAsync(b_state.p, hs.data(), W * sizeof(SeqState), cudaMemcpyHostToDevice, st));
        SB_CUDA_CHECK(cudaMemcpyAsync(b_prompt.p, prompts.data(), prompts.size() * 4, cudaMemcpyHostToDevice, st));
        SB_CUDA_CHECK(cudaMemsetAsync(d_lang, 0xff, (size_t)W * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_ctr.p, 0, 64 * kMaxLanes, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_tokens.p, 0xff, (size_t)W * n_max * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_margins.p, 0, (size_t)W * n_max * 4, st));
        k_fill_i32<<<ceil_div(W, 256), 256, 0, st>>>(b_next.as<int>(), W, prompt[0]);
        g_launches += 1;
        if (forced_host) SB_CUDA_CHECK(cudaMemcpyAsync(b_forced.p, forced_host, (size_t)W * n_max * 4, cudaMemcpyHostToDevice, st));
        SamplerArgs sa0{};
        sa0.forced = nullptr;
        sa0.sp = sp; sa0.n_vocab = hp.n_vocab; sa0.n_max = n_max;
        sa0.suppress_blank = p.suppress_blank; sa0.no_timestamps = p.no_timestamps; sa0.single_segment = p.single_segment;
        sa0.max_initial_tid = p.max_initial_ts > 0.f ? (int)lroundf(p.max_initial_ts / (30.0f / hp.n_audio_ctx)) : -1;
        sa0.prompt = b_prompt.as<int>(); sa0.n_prompt = n_prompt;
        // the detect launch stays in the step graph for the whole call once auto-detect was requested (it is a no-op for
        // sequences whose language is known), so later seek-loop rounds reuse the captured graph
        if (detect) auto_mode = true;
        detect = auto_mode && n_prompt >= 2;
        detect_lang = detect; detect_out = d_lang;

        const int total_steps = n_prompt - 1 + n_max;
        const bool tracing = profile == 2 && !logits_out;
        const bool graph = use_graph && !logits_out;
        // launch trace: [lane][start | end][step][launch] globaltimer stamps, reset per decode() call
        trace_per_step = 0;
        if (tracing) {
            trace_per_step = 11 * hp.n_text_layer + 1;
            trace_max_steps = total_steps;
            const size_t n = (size_t)trace_max_steps * trace_per_step;
            if ((rc = b_trace.ensure((size_t)kMaxLanes * 2 * n * 8))) return rc;
            for (int i = 0; i < kMaxLanes; ++i) {
                SB_CUDA_CHECK(cudaMemsetAsync(b_trace.as<unsigned long long>() + (size_t)i * 2 * n, 0xff, n * 8, st));
                SB_CUDA_CHECK(cudaMemsetAsync(b_trace.as<unsigned long long>() + (size_t)i * 2 * n + n, 0, n * 8, st));
            }
        }
        // lanes: independent sub-batches on their own streams (one lane when tracing logits)
        int n_lanes = logits_out ? 1 : std::min(n_lanes_cfg, std::max(1, W / 8));
        struct LaneRun { int w0, Wl; bool finished; int steps; };
        std::vector<LaneRun> lr(n_lanes);
        for (int i = 0; i < n_lanes; ++i) {
            const int w0 = (int)((int64_t)W * i / n_lanes), w1 = (int)((int64_t)W * (i + 1) / n_lanes);
            lr[i] = LaneRun{w0, w1 - w0, false, 0};
        }
        auto lane_args = [&](int i) {
            SamplerArgs sa = sa0;
            const int w0 = lr[i].w0;
            int* ctr = b_ctr.as<int>() + 16 * i;
            sa.state = b_state.as<SeqState>() + w0;
            sa.step_ptr = ctr + 1;
            sa.tokens_out = b_tokens.as<int>() + (size_t)w0 * n_max;
            sa.margins_out = b_margins.as<float>() + (size_t)w0 * n_max;
            sa.next_tokens = b_next.as<int>() + w0;
            sa.forced = forced_host ? b_forced.as<int>() + (size_t)w0 * n_max : nullptr;
            sa.n_done = ctr + 2;
            sa.pos_ptr = ctr;
            sa.prompt = b_prompt.as<int>() + (size_t)w0 * n_prompt;
            return sa;
        };
        // fork: every lane stream waits for the setup work queued on the main stream
        SB_CUDA_CHECK(cudaEventRecord(ev_fork, st));
        for (int i = 0; i < n_lanes; ++i) SB_CUDA_CHECK(cudaStreamWaitEvent(lanes[i].st, ev_fork, 0));
        if (graph) {
            for (int i = 0; i < n_lanes; ++i) {
                Lane& L = lanes[i];
                GraphKey k{W, lr[i].w0, lr[i].Wl, n_max,
                           (p.suppress_blank ? 1 : 0) | (p.no_timestamps ? 2 : 0) | (p.single_segment ? 4 : 0) | (detect ? 8 : 0) | (tracing ? 16 : 0),
                           sa0.max_initial_tid, n_prompt, forced_host ? 1 : 0, g_ws_gen.load()};
                if (L.gexec && k == L.key) continue;
                if (L.gexec) { cudaGraphExecDestroy(L.gexec); L.gexec = nullptr; }
                cudaGraph_t g = nullptr;
                const uint64_t l0 = g_launches.load();
                SB_CUDA_CHECK(cudaStreamBeginCapture(L.st, cudaStreamCaptureModeThreadLocal));
                rc = enqueue_step(W, lr[i].w0, lr[i].Wl, b_ctr.as<int>() + 16 * i, L.st, lane_args(i));
                cudaError_t ce = cudaStreamEndCapture(L.st, &g);
                L.graph_nodes = (int)(g_launches.load() - l0);
                g_launches -= (uint64_t)L.graph_nodes;      // captured, not executed
                if (rc) { if (g) cudaGraphDestroy(g); return rc; }
                if (ce != cudaSuccess) { set_error(std::string("graph capture: ") + cudaGetErrorString(ce)); return SB_ERR_CUDA; }
                ce = cudaGraphInstantiate(&L.gexec, g, 0);
                cudaGraphDestroy(g);
                if (ce != cudaSuccess) { L.gexec = nullptr; set_error(std::string("graph instantiate: ") + cudaGetErrorString(ce)); return SB_ERR_CUDA; }
                L.key = k;
            }
        }
        int active = n_lanes;
        while (active > 0) {
            for (int i = 0; i < n_lanes; ++i) {
                if (lr[i].finished) continue;
                Lane& L = lanes[i];
                const int burst = logits_out ? 1 : std::min(8, total_steps - lr[i].steps);
                for (int b = 0; b < burst; ++b) {
                    if (graph) { SB_CUDA_CHECK(cudaGraphLaunch(L.gexec, L.st)); g_launches += (uint64_t)L.graph_nodes; }
                    else if ((rc = enqueue_step(W, lr[i].w0, lr[i].Wl, b_ctr.as<int>() + 16 * i, L.st, lane_args(i)))) return rc;
                }
                if (logits_out && lr[i].steps >= n_prompt - 1) {
                    const int sidx = lr[i].steps - (n_prompt - 1);
                    for (int w = 0; w < W; ++w)
                        SB_CUDA_CHECK(cudaMemcpyAsync(logits_out + ((size_t)w * n_max + sidx) * hp.n_vocab,
                                                      b_logits.as<float>() + (size_t)w * round_up(hp.n_vocab, 8), (size_t)hp.n_vocab * 4,
                                                      cudaMemcpyDeviceToHost, L.st));
                }
                lr[i].steps += burst;
                stats.decoder_steps += (double)burst * lr[i].Wl / W; stats.d2h_bytes += 16;
                SB_CUDA_CHECK(cudaMemcpyAsync(h_ctr + 16 * i, b_ctr.as<int>() + 16 * i, 16, cudaMemcpyDeviceToHost, L.st));
            }
            for (int i = 0; i < n_lanes; ++i) {
                if (lr[i].finished) continue;
                SB_CUDA_CHECK(cudaStreamSynchronize(lanes[i].st));
                if (h_ctr[16 * i + 2] >= lr[i].Wl || lr[i].steps >= total_steps) { lr[i].finished = true; --active; }
            }
        }
        // join: the main stream continues after every lane
        for (int i = 0; i < n_lanes; ++i) {
            SB_CUDA_CHECK(cudaEventRecord(lanes[i].done, lanes[i].st));
            SB_CUDA_CHECK(cudaStreamWaitEvent(st, lanes[i].done, 0));
        }
        out.state.resize(W); out.tokens.resize((size_t)W * n_max); out.margins.resize((size_t)W * n_max);
        out.langs.assign(W, -1);
        if (detect) SB_CUDA_CHECK(cudaMemcpyAsync(out.langs.data(), d_lang, (size_t)W * 4, cudaMemcpyDeviceToHost, st));     // -1 where nothing was detected
        SB_CUDA_CHECK(cudaMemcpyAsync(out.state.data(), b_state.p, W * sizeof(SeqState), cudaMemcpyDeviceToHost, st));
        SB_CUDA_CHECK(cudaMemcpyAsync(out.tokens.data(), b_tokens.p, (size_t)W * n_max * 4, cudaMemcpyDeviceToHost, st));
        SB_CUDA_CHECK(cudaMemcpyAsync(out.margins.data(), b_margins.p, (size_t)W * n_max * 4, cudaMemcpyDeviceToHost, st));
        SB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (tracing) {
            const size_t n = (size_t)trace_max_steps * trace_per_step;
            std::vector<unsigned long long> ht((size_t)n_lanes * 2 * n);
            SB_CUDA_CHECK(cudaMemcpy(ht.data(), b_trace.p, ht.size() * 8, cudaMemcpyDeviceToHost));
            for (int i = 0; i < n_lanes; ++i) {
                const unsigned long long* t0 = ht.data() + (size_t)i * 2 * n;
                const unsigned long long* t1 = t0 + n;
                for (int sI = 0; sI < lr[i].steps && sI < trace_max_steps; ++sI) {
                    unsigned long long first = ~0ull, last = 0;
                    for (int k = 0; k < trace_per_step && k < (int)trace_cls.size(); ++k) {
                        const unsigned long long a = t0[(size_t)sI * trace_per_step + k], b = t1[(size_t)sI * trace_per_step + k];
                        if (a == ~0ull || b <= a) continue;          // never ran (every block skipped) or clock wrap
                        const double ms = (double)(b - a) * 1e-6;
                        first = std::min(first, a); last = std::max(last, b);
                        switch (trace_cls[k]) {
                            case 0: stats.skinny_ms += ms; stats.skinny_bytes += trace_work[k]; stats.skinny_launches += 1; break;
                            case 1: stats.dln_ms += ms; stats.dln_launches += 1; break;
                            case 2: stats.dself_ms += ms; stats.dself_launches += 1; break;
                            default: stats.xattn_ms += ms; stats.xattn_launches += 1; break;
                        }
                    }
                    if (last > first) { stats.dstep_ms += (double)(last - first) * 1e-6; stats.dstep_count += 1; }
                    if (last > first && getenv("SB_TRACE_DUMP")) {    // per-index timeline relative to the lane-step's first start
                        if ((int)trace_cnt.size() < trace_per_step) {
                            trace_sum_dur.assign(trace_per_step, 0.0); trace_sum_t0.assign(trace_per_step, 0.0);
                            trace_sum_t1.assign(trace_per_step, 0.0); trace_cnt.assign(trace_per_step, 0.0);
                        }
                        for (int k = 0; k < trace_per_step && k < (int)trace_cls.size(); ++k) {
                            const unsigned long long a = t0[(size_t)sI * trace_per_step + k], b = t1[(size_t)sI * trace_per_step + k];
                            if (a == ~0ull || b <= a) continue;
                            trace_sum_dur[k] += (double)(b - a); trace_sum_t0[k] += (double)(a - first);
                            trace_sum_t1[k] += (double)(b - first); trace_cnt[k] += 1;
                        }
                    }
                }
            }
            if (const char* dump = getenv("SB_TRACE_DUMP")) {
                if (FILE* f = fopen(dump, "w")) {
                    fprintf(f, "idx,class,work_bytes,count,avg_us,avg_start_us,avg_end_us\n");
                    for (int k = 0; k < (int)trace_cnt.size() && k < (int)trace_cls.size(); ++k)
                        if (trace_cnt[k] > 0)
                            fprintf(f, "%d,%d,%.0f,%.0f,%.3f,%.3f,%.3f\n", k, trace_cls[k], trace_work[k], trace_cnt[k],
                                    trace_sum_dur[k] / trace_cnt[k] * 1e-3, trace_sum_t0[k] / trace_cnt[k] * 1e-3, trace_sum_t1[k] / trace_cnt[k] * 1e-3);
                    fclose(f);
                }
            }
            // cross-attention bytes: a sequence with n_tok sampled tokens was live for n_prompt + n_tok - 1 steps
            for (int w = 0; w < W; ++w)
                stats.xattn_bytes += (double)(n_prompt + std::max(out.state[w].n_tok, 1) - 1) * hp.n_text_layer * 4.0 * hp.n_audio_ctx * hp.n_text_state;
        }
        stats.d2h_bytes += (double)W * (sizeof(SeqState) + 8.0 * n_max);
        stats.h2d_bytes += (double)W * sizeof(SeqState) + 4.0 * n_prompt;
        return SB_OK;
    }

    int resolve_language(const sb_params& p, int* lang) {
        if (p.initial_prompt && p.initial_prompt[0]) { set_error("initial_prompt is not implemented yet (needs the BPE encoder)"); return SB_ERR_UNSUPPORTED; }
        // NULL / "" / "auto": whisper_full detects the language on the first window (multilingual models);
        // English-only models have no language token at all
        if (!p.language || !p.language[0] || std::string(p.language) == "auto") { *lang = hp.n_vocab >= 51865 ? -1 : 0; return SB_OK; }
        std::string l = p.language;
        if (l == "zh-Hans" || l == "zh-Hant") l = "zh";   // reference: transcription.rs:448-459
        const int id = lang_id(l.c_str());
        if (id < 0 || (hp.n_vocab >= 51865 && id >= sp.num_languages)) { set_error("unknown language code: " + l); return SB_ERR_INVALID; }
        *lang = id;
        return SB_OK;
    }

    // im2col of the mel windows given as a host array [W][n_mel][3000] (parity hooks)
    int stage_mel_windows(const float* mel_windows, int W) {
        const int n_mel = hp.n_mels, nfr = 2 * hp.n_audio_ctx;
        int rc;
        if ((rc = b_mel.ensure((size_t)W * n_mel * nfr * 4))) return rc;
        if ((rc = b_floor.ensure(W * 4))) return rc;
        if ((rc = b_clipmeta.ensure(W * 2 * 4))) return rc;
        if ((rc = b_winmeta.ensure(W * 2 * 4))) return rc;
        SB_CUDA_CHECK(cudaMemcpyAsync(b_mel.p, mel_windows, (size_t)W * n_mel * nfr * 4, cudaMemcpyHostToDevice, st));
        std::vector<int> meta(4 * W);
        for (int w = 0; w < W; ++w) { meta[w] = nfr; meta[W + w] = nfr; meta[2 * W + w] = w; meta[3 * W + w] = 0; }
        SB_CUDA_CHECK(cudaMemcpyAsync(b_clipmeta.p, meta.data(), 2 * W * 4, cudaMemcpyHostToDevice, st));
        SB_CUDA_CHECK(cudaMemcpyAsync(b_winmeta.p, meta.data() + 2 * W, 2 * W * 4, cudaMemcpyHostToDevice, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_floor.p, 0, W * 4, st));
        SB_CUDA_CHECK(cudaStreamSynchronize(st));   // meta is a stack vector
        Im2col1Args a{b_mel.as<float>(), b_floor.as<float>(), b_winmeta.as<int>(), b_winmeta.as<int>() + W,
                      b_clipmeta.as<int>(), b_clipmeta.as<int>() + W, (int64_t)n_mel * nfr, nfr, n_mel, nfr};
        return im2col_conv1<T>(a, b_col1.as<T>(), W, st);
    }

    int encode_host(const float* mel_windows, int W, float* enc_out) override {
        SB_CHECK_ARG(mel_windows && enc_out && W > 0 && W <= max_batch, "sb_encode: bad arguments (n_windows <= max_batch)");
        SB_CUDA_CHECK(cudaSetDevice(device));
        int rc = ensure_encoder_ws(W, W, true);
        if (rc) return rc;
        if ((rc = stage_mel_windows(mel_windows, W))) return rc;
        SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
        if ((rc = encode_chunk(W, 0, b_enc32.as<float>()))) return rc;
        SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
        SB_CUDA_CHECK(cudaMemcpyAsync(enc_out, b_enc32.p, (size_t)W * hp.n_audio_ctx * hp.n_audio_state * 4, cudaMemcpyDeviceToHost, st));
        SB_CUDA_CHECK(cudaStreamSynchronize(st));
        { float t = 0.f; cudaEventElapsedTime(&t, ev[0], ev[1]); stats.encode_ms += t; stats.windows += W; }   // conv stem .. ln_post + cross-KV
        prof_collect();
        return SB_OK;
    }

    int decode_trace(const float* mel_windows, int W, const int32_t* seek_end, const sb_params& p, const int32_t* forced,
                     int n_steps, float* logits_out, int32_t* tokens_out, float* margins_out) override {
        SB_CHECK_ARG(mel_windows && seek_end && tokens_out && W > 0 && W <= max_batch && n_steps > 0, "sb_decode_trace: bad arguments");
        SB_CUDA_CHECK(cudaSetDevice(device));
        int lang = 0;
        int rc = resolve_language(p, &lang);
        if (rc) return rc;
        auto_mode = lang < 0;
        if ((rc = ensure_encoder_ws(W, W, false))) return rc;
        if ((rc = stage_mel_windows(mel_windows, W))) return rc;
        if ((rc = encode_chunk(W, 0, nullptr))) return rc;
        std::vector<int> seek(W, 0), se(seek_end, seek_end + W);
        DecodeOut out;
        if ((rc = decode(W, seek, se, p, std::vector<int>(W, lang), forced, n_steps, logits_out, out))) return rc;
        prof_collect();
        if (out.n_max != n_steps) { set_error("n_steps exceeds n_text_ctx/2 - 4"); return SB_ERR_INVALID; }
        memcpy(tokens_out, out.tokens.data(), (size_t)W * n_steps * 4);
        if (margins_out) memcpy(margins_out, out.margins.data(), (size_t)W * n_steps * 4);
        return SB_OK;
    }

    // ---- whisper_full over a group of <= max_batch clips -------------------------------------
    struct ClipRun {
        size_t n = 0; int n_len = 0, n_len_org = 0, n_calc = 0;
        int seek = 0; bool active = false; int lang = 0;
        std::vector<int32_t> kept, sampled; std::vector<float> margins; std::vector<sb_window_info> windows;
        std::string text;
    };

    int run_group(const float* const* pcm, const size_t* ns, int G, const sb_params& p, int lang, sb_result* out) {
        const int n_mel = hp.n_mels;
        std::vector<ClipRun> clips(G);
        size_t max_n = 0; int max_calc = 0; bool uniform = true;
        for (int c = 0; c < G; ++c) {
            ClipRun& r = clips[c];
            r.n = ns[c];
            r.lang = lang;
            if (r.n == 0) continue;
            sb_logmel_geometry(r.n, &r.n_len, &r.n_len_org, &r.n_calc);
            // whisper.cpp: "input is too short" below 1 s -> no segments
            r.active = r.n >= 201 && r.n_len_org >= 100;
            if (!r.active) continue;
            max_n = std::max(max_n, r.n); max_calc = std::max(max_calc, r.n_calc);
        }
        for (int c = 0; c < G; ++c) if (clips[c].active && clips[c].n != max_n) uniform = false;
        float ms_mel = 0.f, ms_enc = 0.f, ms_dec = 0.f;
        int rc;
        if (max_n > 0) {
            const int stride = (int)round_up(max_calc, 32);
            if ((rc = b_pcm.ensure((size_t)G * max_n * 4))) return rc;
            if ((rc = b_mel.ensure((size_t)G * n_mel * stride * 4))) return rc;
            if ((rc = b_cmax.ensure(G * 4))) return rc;
            if ((rc = b_floor.ensure(G * 4))) return rc;
            if ((rc = b_clipmeta.ensure(G * 2 * 4))) return rc;
            if ((rc = b_winmeta.ensure(G * 2 * 4))) return rc;
            SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
            std::vector<int> meta(2 * G, 0);
            for (int c = 0; c < G; ++c) {
                if (!clips[c].active) continue;
                // cudaMemcpyDefault: the clip may live in host (pinned or pageable) or device memory
                SB_CUDA_CHECK(cudaMemcpyAsync(b_pcm.as<float>() + (size_t)c * max_n, pcm[c], clips[c].n * 4, cudaMemcpyDefault, st));
                stats.pcm_bytes += clips[c].n * 4.0;
                meta[c] = clips[c].n_calc; meta[G + c] = clips[c].n_len;
            }
            SB_CUDA_CHECK(cudaMemcpyAsync(b_clipmeta.p, meta.data(), 2 * G * 4, cudaMemcpyHostToDevice, st));
            bool all_active = true;
            for (int c = 0; c < G; ++c) all_active = all_active && clips[c].active;
            if (uniform && all_active) {
                if ((rc = logmel_launch(melplan, b_pcm.as<float>(), G, max_n, (int64_t)max_n, b_mel.as<float>(),
                                        (int64_t)n_mel * stride, stride, b_cmax.as<int32_t>(), b_floor.as<float>(), st))) return rc;
            } else {
                for (int c = 0; c < G; ++c) {
                    if (!clips[c].active) continue;
                    if ((rc = logmel_launch(melplan, b_pcm.as<float>() + (size_t)c * max_n, 1, clips[c].n, (int64_t)max_n,
                                            b_mel.as<float>() + (size_t)c * n_mel * stride, (int64_t)n_mel * stride, stride,
                                            b_cmax.as<int32_t>() + c, b_floor.as<float>() + c, st))) return rc;
                }
            }
            SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
            SB_CUDA_CHECK(cudaStreamSynchronize(st));   // meta vector lifetime + mel timing
            { float t; cudaEventElapsedTime(&t, ev[0], ev[1]); ms_mel += t; }

            // seek loop: every round encodes + decodes one window of every clip that still has audio
            const int max_windows = p.max_windows > 0 ? p.max_windows : 1 << 20;
            for (;;) {
                std::vector<int> wclip, wseek, wend, wlang;
                for (int c = 0; c < G; ++c) {
                    ClipRun& r = clips[c];
                    if (!r.active) continue;
                    if (r.seek + 100 >= r.n_len_org || (int)r.windows.size() >= max_windows) { r.active = false; continue; }
                    wclip.push_back(c); wseek.push_back(r.seek); wend.push_back(r.n_len_org); wlang.push_back(r.lang);
                }
                const int W = (int)wclip.size();
                if (W == 0) break;
                if ((rc = ensure_encoder_ws(W, W, false))) return rc;
                std::vector<int> wm(2 * W);
                for (int w = 0; w < W; ++w) { wm[w] = wclip[w]; wm[W + w] = wseek[w]; }
                SB_CUDA_CHECK(cudaMemcpyAsync(b_winmeta.p, wm.data(), 2 * W * 4, cudaMemcpyHostToDevice, st));
                SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
                Im2col1Args a{b_mel.as<float>(), b_floor.as<float>(), b_winmeta.as<int>(), b_winmeta.as<int>() + W,
                              b_clipmeta.as<int>(), b_clipmeta.as<int>() + G, (int64_t)n_mel * stride, stride, n_mel,
                              2 * hp.n_audio_ctx};
                if ((rc = im2col_conv1<T>(a, b_col1.as<T>(), W, st))) return rc;
                if ((rc = encode_chunk(W, 0, nullptr))) return rc;
                SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
                DecodeOut d;
                if ((rc = decode(W, wseek, wend, p, wlang, nullptr, 0, nullptr, d))) return rc;
                SB_CUDA_CHECK(cudaEventRecord(ev[2], st));
                SB_CUDA_CHECK(cudaStreamSynchronize(st));
                { float t; cudaEventElapsedTime(&t, ev[0], ev[1]); ms_enc += t; cudaEventElapsedTime(&t, ev[1], ev[2]); ms_dec += t; }
                prof_collect();
                stats.windows += W; stats.rounds += 1;
                for (int w = 0; w < W; ++w) {
                    ClipRun& r = clips[wclip[w]];
                    if (r.lang < 0) r.lang = d.langs[w] >= 0 ? d.langs[w] : 0;     // detected on the clip's first window
                    const SeqState& s = d.state[w];
                    sb_window_info wi{};
                    wi.seek = r.seek; wi.n_tokens = s.n_tok; wi.result_len = s.result_len; wi.seek_delta = s.seek_delta;
                    wi.failed = s.failed; wi.token_offset = (int)r.sampled.size();
                    for (int i = 0; i < s.n_tok; ++i) {
                        const int t = d.tokens[(size_t)w * d.n_max + i];
                        r.sampled.push_back(t); r.margins.push_back(d.margins[(size_t)w * d.n_max + i]);
                        if (i < s.result_len) { r.kept.push_back(t); if (t < sp.eot) r.text += token_text(t); }
                    }
                    r.windows.push_back(wi);
                    r.seek += s.seek_delta;
                }
            }
        }
        for (int c = 0; c < G; ++c) {
            ClipRun& r = clips[c];
            sb_result& o = out[c];
            memset(&o, 0, sizeof(o));
            // transcribe-rs: full text trimmed
            size_t b = 0, e = r.text.size();
            auto ws = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r' || ch == '\v' || ch == '\f'; };
            while (b < e && ws(r.text[b])) ++b;
            while (e > b && ws(r.text[e - 1])) --e;
            o.text_len = e - b;
            o.text = (char*)malloc(o.text_len + 1);
            memcpy(o.text, r.text.data() + b, o.text_len); o.text[o.text_len] = 0;
            auto dup = [](const void* src, size_t bytes) -> void* { void* p = malloc(bytes ? bytes : 1); if (bytes) memcpy(p, src, bytes); return p; };
            o.n_tokens = r.kept.size(); o.tokens = (int32_t*)dup(r.kept.data(), r.kept.size() * 4);
            o.n_sampled = r.sampled.size(); o.sampled = (int32_t*)dup(r.sampled.data(), r.sampled.size() * 4);
            o.margins = (float*)dup(r.margins.data(), r.margins.size() * 4);
            o.n_windows = r.windows.size(); o.windows = (sb_window_info*)dup(r.windows.data(), r.windows.size() * sizeof(sb_window_info));
            o.ms_mel = ms_mel; o.ms_encode = ms_enc; o.ms_decode = ms_dec; o.status = SB_OK;
            o.lang_id = hp.n_vocab >= 51865 ? r.lang : -1;
            stats.tokens_sampled += (double)r.sampled.size();
        }
        stats.mel_ms += ms_mel; stats.encode_ms += ms_enc; stats.decode_ms += ms_dec; stats.clips += G;
        return SB_OK;
    }

    int transcribe_batch(const float* const* pcm, const size_t* ns, size_t count, const sb_params& p, sb_result* out) override {
        SB_CUDA_CHECK(cudaSetDevice(device));
        int lang = 0;
        int rc = resolve_language(p, &lang);
        if (rc) return rc;
        auto_mode = lang < 0;
        for (size_t c0 = 0; c0 < count; c0 += max_batch) {
            const int G = (int)std::min<size_t>(max_batch, count - c0);
            if ((rc = run_group(pcm + c0, ns + c0, G, p, lang, out + c0))) return rc;
        }
        return SB_OK;
    }
};

}  // namespace sb

struct sb_engine { std::unique_ptr<sb::EngineBase> impl; };

extern "C" {

void sb_params_default(sb_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->language = "en";
    p->suppress_blank = 1;
    p->max_initial_ts = 1.0f;
}

int sb_engine_create(const sb_config* cfg, sb_engine** out) {
    SB_CHECK_ARG(cfg && out && cfg->model_path, "cfg/out/model_path is null");
    SB_CHECK_ARG(cfg->dtype == SB_DTYPE_BF16 || cfg->dtype == SB_DTYPE_F16, "cfg.dtype");
    int ndev = 0;
    SB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    SB_CHECK_ARG(cfg->device >= 0 && cfg->device < ndev, "cfg.device out of range");
    SB_CUDA_CHECK(cudaSetDevice(cfg->device));
    sb::GgmlFile file;
    int rc = sb::load_ggml_file(cfg->model_path, file);
    if (rc) return rc;
    std::unique_ptr<sb::EngineBase> e;
    if (cfg->dtype == SB_DTYPE_F16) e.reset(new sb::Engine<__half>());
    else e.reset(new sb::Engine<__nv_bfloat16>());
    e->device = cfg->device;
    e->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 64;
    e->dtype = cfg->dtype;
    e->use_graph = cfg->use_cuda_graph;
    if (cfg->dtype == SB_DTYPE_F16) rc = static_cast<sb::Engine<__half>*>(e.get())->load(file);
    else rc = static_cast<sb::Engine<__nv_bfloat16>*>(e.get())->load(file);
    if (rc) return rc;
    SB_CUDA_CHECK(cudaDeviceSynchronize());
    sb_engine* h = new sb_engine();
    h->impl = std::move(e);
    *out = h;
    return SB_OK;
}

int sb_engine_destroy(sb_engine* e) {
    if (!e) return SB_OK;
    cudaSetDevice(e->impl->device);
    cudaDeviceSynchronize();
    delete e;
    return SB_OK;
}

int sb_engine_info(const sb_engine* e, sb_model_info* info) {
    SB_CHECK_ARG(e && info, "null pointer");
    const sb::WhisperHParams& h = e->impl->hp;
    *info = sb_model_info{h.n_vocab, h.n_audio_ctx, h.n_audio_state, h.n_audio_head, h.n_audio_layer, h.n_text_ctx,
                          h.n_text_state, h.n_text_head, h.n_text_layer, h.n_mels, h.ftype,
                          e->impl->sp.eot, e->impl->sp.sot, e->impl->sp.beg, e->impl->sp.blank};
    return SB_OK;
}

void* sb_engine_stream(sb_engine* e) { return e ? e->impl->stream_handle() : nullptr; }

int sb_engine_set_profile(sb_engine* e, int enable) {
    SB_CHECK_ARG(e, "null engine");
    e->impl->profile = enable;
    return SB_OK;
}

int sb_engine_stats(sb_engine* e, sb_stats* out, int reset) {
    SB_CHECK_ARG(e && out, "null pointer");
    *out = e->impl->stats;
    if (reset) memset(&e->impl->stats, 0, sizeof(sb_stats));
    return SB_OK;
}

int sb_token_text(const sb_engine* e, int32_t id, char* buf, int cap) {
    if (!e) return 0;
    const std::string s = e->impl->token_text(id);
    if (buf && cap > 0) memcpy(buf, s.data(), std::min<size_t>(s.size(), (size_t)cap));
    return (int)s.size();
}

int sb_tokenize(const sb_engine* e, const char* text, int32_t* tokens, int cap) {
    if (!e) { sb::set_error("Model is not loaded for transcription."); return SB_ERR_NOT_LOADED; }
    SB_CHECK_ARG(text && (tokens || cap == 0), "null pointer");
    const std::vector<int> t = e->impl->tokenize(text);
    for (int i = 0; i < (int)t.size() && i < cap; ++i) tokens[i] = t[i];
    return (int)t.size();
}

int sb_transcribe_batch(sb_engine* e, const float* const* pcm16k, const size_t* n_samples, size_t count,
                        const sb_params* p, sb_result* out) {
    if (!e) { sb::set_error("Model is not loaded for transcription."); return SB_ERR_NOT_LOADED; }
    SB_CHECK_ARG(out && (count == 0 || (pcm16k && n_samples)), "null pointer");
    memset(out, 0, count * sizeof(sb_result));
    sb_params dp;
    if (!p) { sb_params_default(&dp); p = &dp; }
    for (size_t i = 0; i < count; ++i) SB_CHECK_ARG(n_samples[i] == 0 || pcm16k[i], "null clip pointer");
    return e->impl->transcribe_batch(pcm16k, n_samples, count, *p, out);
}

int sb_transcribe(sb_engine* e, const float* pcm16k, size_t n_samples, const sb_params* p, sb_result* out) {
    const float* ptrs[1] = {pcm16k};
    size_t ns[1] = {n_samples};
    return sb_transcribe_batch(e, ptrs, ns, 1, p, out);
}

void sb_result_free(sb_result* r) {
    if (!r) return;
    free(r->text); free(r->tokens); free(r->sampled); free(r->margins); free(r->windows);
    memset(r, 0, sizeof(*r));
}

int sb_encode(sb_engine* e, const float* mel_windows, int n_windows, float* enc_out) {
    if (!e) { sb::set_error("Model is not loaded for transcription."); return SB_ERR_NOT_LOADED; }
    return e->impl->encode_host(mel_windows, n_windows, enc_out);
}

int sb_decode_trace(sb_engine* e, const float* mel_windows, int n_windows, const int32_t* seek_end, const sb_params* p,
                    const int32_t* forced, int n_steps, float* logits_out, int32_t* tokens_out, float* margins_out) {
    if (!e) { sb::set_error("Model is not loaded for transcription."); return SB_ERR_NOT_LOADED; }
    sb_params dp;
    if (!p) { sb_params_default(&dp); p = &dp; }
    return e->impl->decode_trace(mel_windows, n_windows, seek_end, *p, forced, n_steps, logits_out, tokens_out, margins_out);
}

}  // extern "C"
