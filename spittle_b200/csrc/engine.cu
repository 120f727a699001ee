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
#include <cmath>
#include <thread>
#include <cstddef>
#include <unordered_map>
#include <vector>

struct sb_melplan;
namespace sb {
extern std::atomic<uint64_t> g_launches;
int logmel_launch_ragged(const sb_melplan* plan, const float* const* pcm_ptrs, const int* n_samples_v, const int* n_calc_v,
                         int n_clips, int max_n_calc, float* mel, int64_t mel_clip_stride, int mel_stride, int32_t* clip_max,
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

struct PinBuf {        // pinned host staging
    void* p = nullptr; size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return SB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMallocHost(&p, need);
        if (e != cudaSuccess) { set_error(std::string("cudaMallocHost(") + std::to_string(need) + "): " + cudaGetErrorString(e)); return SB_ERR_NOMEM; }
        bytes = need;
        return SB_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

// std::mt19937 restated (whisper.cpp gives every decoder a std::mt19937(0)); uniform() is libstdc++'s
// std::generate_canonical<double, 53>: two 32-bit draws, (x1 + x2 * 2^32) / 2^64 -- what std::discrete_distribution consumes
struct Mt19937 {
    uint32_t mt[624]; int idx = 624;
    explicit Mt19937(uint32_t seed = 0) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    double uniform() {
        const double x1 = (double)next(), x2 = (double)next();
        double u = (x1 + x2 * 4294967296.0) / 18446744073709551616.0;
        if (u >= 1.0) u = std::nextafter(1.0, 0.0);
        return u;
    }
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

// whisper.cpp's `non_speech_tokens` (the list of OpenAI's tokenizer.non_speech_tokens) as ids of THIS vocabulary: every listed
// string with and without a leading space, plus " -" and " '" ("allow hyphens and single quotes between words, but not at the
// beginning of a word"); whisper_process_logits sets them to -inf when params.suppress_nst is on.  Ascending, unique.
static std::vector<int> non_speech_token_ids(const std::vector<std::string>& vocab) {
    static const char* const kList[] = {
        "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`", "{", "|", "}", "~",
        "\xe3\x80\x8c", "\xe3\x80\x8d", "\xe3\x80\x8e", "\xe3\x80\x8f",                       // corner brackets
        "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[", "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}",
        "\xe2\x99\xaa\xe2\x99\xaa", "\xe2\x99\xaa\xe2\x99\xaa\xe2\x99\xaa",                 // two / three eighth notes
        "\xe2\x99\xa9", "\xe2\x99\xaa", "\xe2\x99\xab", "\xe2\x99\xac", "\xe2\x99\xad", "\xe2\x99\xae", "\xe2\x99\xaf"};
    std::unordered_map<std::string, int> t2i;
    for (int i = 0; i < (int)vocab.size(); ++i) t2i[vocab[i]] = i;        // last id of a duplicate, like the loader
    std::vector<int> ids;
    auto add = [&](const std::string& s) { auto f = t2i.find(s); if (f != t2i.end()) ids.push_back(f->second); };
    for (const char* t : kList) { add(t); add(std::string(" ") + t); }
    add(" -"); add(" '");
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    return ids;
}

static SpecialIds special_from_vocab(int n_vocab, const std::vector<std::string>& vocab) {
    SpecialIds s{50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362, 50363, 0, 0, 220};
    // whisper.cpp: num_languages = n_vocab - 51765 - (multilingual ? 1 : 0); English-only vocabularies keep the 99 language
    // ids 50258..50356 (always suppressed by the logits filter) even though their prompt carries no language token
    s.num_languages = std::max(0, n_vocab - 51765);
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
    DevBuf b_pcm, b_mel, b_cmax, b_floor, b_clipmeta, b_winmeta, b_pcmptr;
    DevBuf b_col1, b_c1, b_x, b_h, b_qkv, b_att, b_mlp, b_enc32;
    DevBuf b_ckv, b_kself, b_vself, b_dx, b_dh, b_dqkv, b_datt, b_dq, b_dmlp, b_logits;
    DevBuf b_state, b_tokens, b_margins, b_tids, b_plogs, b_next, b_forced, b_tick, b_prompt, b_lang, b_init, b_prow, b_temp, b_rng;
    DevBuf b_nst;                 // ids of whisper.cpp's non-speech tokens present in this vocabulary (suppress_nst)
    int n_nst = 0;
    PinBuf h_state, h_tokens, h_margins, h_tids, h_plogs, h_lang, h_init, h_winmeta, h_prow, h_rng;   // host mirrors polled once per burst / slot-init staging
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
        int S, s0, n, n_max, flags, max_init, has_forced, gen;
        bool operator==(const GraphKey& o) const {
            return S == o.S && s0 == o.s0 && n == o.n && n_max == o.n_max && flags == o.flags && max_init == o.max_init &&
                   has_forced == o.has_forced && gen == o.gen;
        }
    };
    // A decode SLOT holds one window being decoded (its cross-KV rows, self-KV cache, SeqState, prompt, token row); a
    // LANE is a contiguous range of slots stepped together on its own stream with its own captured step graph.
    struct Lane {
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        cudaGraphExec_t gexec = nullptr;
        GraphKey key{0, 0, 0, 0, 0, 0, 0, -1};
        int graph_nodes = 0;
        int s0 = 0, n = 0;        // slots [s0, s0 + n)
        int active = 0;           // slots of this lane holding a live sequence
        int steps = 0;            // steps run in the current call
    };
    Lane lanes[kMaxLanes];
    cudaEvent_t ev_fork = nullptr, ev_enc = nullptr;
    int n_lanes_cfg = 2;
    int n_slots = 0, n_lanes = 0, n_max_cur = 0;
    int refill_min_cfg = 0;       // free slots needed before a refill encode is started (0: max(1, slots / 8))
    int enc_batch_max = 0;        // windows per encoder batch (0: as many as there are free slots): smaller first batches let the
                                  // decode lanes start while the rest of the group is still being encoded
    int prefill_min = 8;          // prompts of at least this many tokens are prefilled in one pass (0: token-by-token feed)

    ~Engine() override {
        for (auto& l : lanes) {
            if (l.gexec) cudaGraphExecDestroy(l.gexec);
            if (l.done) cudaEventDestroy(l.done);
            if (l.st) cudaStreamDestroy(l.st);
        }
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_enc) cudaEventDestroy(ev_enc);
        for (auto& r : prof_pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        for (void* p : owned) cudaFree(p);
        DevBuf* bufs[] = {&b_pcm, &b_mel, &b_cmax, &b_floor, &b_clipmeta, &b_winmeta, &b_col1, &b_c1, &b_x, &b_h, &b_qkv,
                          &b_att, &b_mlp, &b_enc32, &b_ckv, &b_kself, &b_vself, &b_dx, &b_dh, &b_dqkv, &b_datt, &b_dq,
                          &b_dmlp, &b_logits, &b_state, &b_tokens, &b_margins, &b_tids, &b_next, &b_forced, &b_tick, &b_prompt,
                          &b_lang, &b_init, &b_trace, &b_prow, &b_plogs, &b_temp, &b_rng, &b_pcmptr, &b_nst};
        for (DevBuf* b : bufs) b->release();
        PinBuf* pins[] = {&h_state, &h_tokens, &h_margins, &h_tids, &h_lang, &h_init, &h_winmeta, &h_prow, &h_plogs, &h_rng};
        for (PinBuf* b : pins) b->release();
        if (melplan) sb_melplan_destroy(melplan);
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
        {
            const std::vector<int> ids = non_speech_token_ids(vocab);
            n_nst = (int)ids.size();
            if (n_nst) {
                int rc = b_nst.ensure(ids.size() * sizeof(int));
                if (rc) return rc;
                SB_CUDA_CHECK(cudaMemcpy(b_nst.p, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
        }
        for (auto& e : ev) SB_CUDA_CHECK(cudaEventCreate(&e));
        for (auto& l : lanes) {
            SB_CUDA_CHECK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
            SB_CUDA_CHECK(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
        }
        SB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        SB_CUDA_CHECK(cudaEventCreate(&ev_enc));
        if (const char* e = getenv("SB_REFILL_MIN")) refill_min_cfg = std::max(0, atoi(e));
        if (const char* e = getenv("SB_PREFILL_MIN")) prefill_min = std::max(0, atoi(e));
        if (const char* e = getenv("SB_ENC_BATCH_MAX")) enc_batch_max = std::max(0, atoi(e));
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
    // writes the cross-KV of window i into decode slot slots[i] (identity when slots == nullptr) and (optionally) the
    // f32 encoder output
    int encode_chunk(int W, const int* slots, float* enc32_out) {
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
            if ((rc = attn_enc_tc<T>(b_qkv.as<T>(), b_att.as<T>(), W, nctx, d, hp.n_audio_head, st))) return rc;
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
        // cross-KV for every decoder layer in one GEMM per run of consecutive slots: [rows, Ld*2*d]
        const int nkv = hp.n_text_layer * 2 * d;
        for (int i = 0; i < W;) {
            int j = i + 1;
            while (slots && j < W && slots[j] == slots[j - 1] + 1) ++j;
            if (!slots) j = W;
            const int slot0 = slots ? slots[i] : 0;
            ep = GemmEpilogue{b_ckv.as<T>() + (int64_t)slot0 * nctx * nkv, nkv, 0, cross_kv.b, 0, nullptr, 0, 0};
            if ((rc = gemm_p(b_h.as<T>() + (int64_t)i * nctx * d, d, cross_kv.w, d, (j - i) * nctx, nkv, d, ep))) return rc;
            i = j;
        }
        return SB_OK;
    }

    int ensure_encoder_ws(int Wc, bool want32) {
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
        return SB_OK;
    }

    int ensure_decoder_ws(int S, int n_max) {
        const size_t d = hp.n_text_state;
        int rc;
        const size_t kvb = (size_t)hp.n_text_layer * S * hp.n_text_ctx * d * 2;
        if ((rc = b_ckv.ensure((size_t)S * hp.n_audio_ctx * hp.n_text_layer * 2 * d * 2))) return rc;
        if ((rc = b_kself.ensure(kvb))) return rc;
        if ((rc = b_vself.ensure(kvb))) return rc;
        if ((rc = b_dx.ensure(S * d * 4))) return rc;
        if ((rc = b_dh.ensure(S * d * 2))) return rc;
        if ((rc = b_dqkv.ensure(S * 3 * d * 2))) return rc;
        if ((rc = b_datt.ensure(S * d * 2))) return rc;
        if ((rc = b_dq.ensure(S * d * 2))) return rc;
        if ((rc = b_dmlp.ensure(S * 4 * d * 2))) return rc;
        if ((rc = b_logits.ensure((size_t)S * round_up(hp.n_vocab, 8) * 4))) return rc;
        if ((rc = b_state.ensure(S * sizeof(SeqState)))) return rc;
        if ((rc = b_tokens.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = b_margins.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = b_tids.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = b_plogs.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = b_temp.ensure((size_t)S * 4))) return rc;
        if ((rc = b_rng.ensure((size_t)S * n_max * 8))) return rc;
        if ((rc = h_plogs.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = h_rng.ensure((size_t)S * n_max * 8))) return rc;
        if ((rc = b_forced.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = b_next.ensure(S * 4))) return rc;
        if ((rc = b_lang.ensure(S * 4))) return rc;
        if ((rc = b_tick.ensure(64 * kMaxLanes))) return rc;
        if ((rc = b_prompt.ensure((size_t)S * kMaxPrompt * 4))) return rc;
        if ((rc = b_init.ensure((size_t)S * sizeof(SlotInit)))) return rc;
        if ((rc = h_state.ensure(S * sizeof(SeqState)))) return rc;
        if ((rc = h_tokens.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = h_margins.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = h_tids.ensure((size_t)S * n_max * 4))) return rc;
        if ((rc = h_lang.ensure(S * 4))) return rc;
        if ((rc = h_init.ensure((size_t)S * sizeof(SlotInit)))) return rc;
        return SB_OK;
    }

    // ---- one decoder step for the slots of lane `li`, all launches on the lane's stream ----
    int enqueue_step(int li, const SamplerArgs& sa, bool honor_done) {
        const Lane& Ln = lanes[li];
        const int S = n_slots, w0 = Ln.s0, Wl = Ln.n;
        cudaStream_t sl = Ln.st;
        const int d = hp.n_text_state, nctx = hp.n_audio_ctx;
        const int nkv = hp.n_text_layer * 2 * d;
        float* dx = b_dx.as<float>() + (int64_t)w0 * d;
        T* dh = b_dh.as<T>() + (int64_t)w0 * d;
        T* dqkv = b_dqkv.as<T>() + (int64_t)w0 * 3 * d;
        T* datt = b_datt.as<T>() + (int64_t)w0 * d;
        T* dq = b_dq.as<T>() + (int64_t)w0 * d;
        T* dmlp = b_dmlp.as<T>() + (int64_t)w0 * 4 * d;
        const int vpad = (int)round_up(hp.n_vocab, 8);
        float* logits = b_logits.as<float>() + (int64_t)w0 * vpad;
        const SeqState* seq_state = sa.state;
        int rc;
        // PDL chain, one launch per stage (what was measured against it and lost -- a persistent per-step megakernel,
        // split-K / cluster projections, LayerNorm fused into the projections -- is in profiles/r1_mega_stage_trace.md)
        // profile == 2: every stage launch of the step gets a trace slot (class, algorithmic bytes); the step keeps its graph
        const bool tracing = profile == 2 && trace_per_step > 0;
        int tidx = 0;
        auto tr = [&](int cls, double work) {
            if (!tracing) return;
            if ((int)trace_cls.size() <= tidx) { trace_cls.resize(tidx + 1); trace_work.resize(tidx + 1); }
            trace_cls[tidx] = cls; trace_work[tidx] = work;
            const size_t n = (size_t)trace_max_steps * trace_per_step;
            TraceSlot ts;
            ts.t0 = b_trace.as<unsigned long long>() + (size_t)li * 2 * n;
            ts.t1 = ts.t0 + n;
            ts.tick = sa.tick; ts.idx = tidx; ts.per_step = trace_per_step; ts.max_steps = trace_max_steps;
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
            T* kc = b_kself.as<T>() + ((int64_t)l * S + w0) * hp.n_text_ctx * d;
            T* vc = b_vself.as<T>() + ((int64_t)l * S + w0) * hp.n_text_ctx * d;
            SkinnyEpilogue e{};
            // attn_ln; layer 0 forms x = token_embedding[tok] + positional_embedding[pos] first
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln1.g, L.ln1.b, dh, Wl, d, l == 0 ? tok_emb : nullptr, dec_pos, next_tok, seq_state, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.qkv.b; e.out16 = dqkv; e.ldo16 = 3 * d;
            if ((rc = sk(dh, d, L.qkv.w, d, 3 * d, d, e))) return rc;
            tr(2, 0);
            if ((rc = dec_self_attn<T>(dqkv, kc, vc, datt, seq_state, honor_done ? 1 : 0, Wl, hp.n_text_head, d, hp.n_text_ctx, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.o.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(datt, d, L.o.w, d, d, d, e))) return rc;
            const T* kb = b_ckv.as<T>() + (int64_t)w0 * nctx * nkv + (int64_t)l * 2 * d;
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln2.g, L.ln2.b, dh, Wl, d, nullptr, nullptr, nullptr, seq_state, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.cq.b; e.out16 = dq; e.ldo16 = d;
            if ((rc = sk(dh, d, L.cq.w, d, d, d, e))) return rc;
            tr(3, 0);
            if ((rc = dec_cross_attn<T>(dq, d, kb, kb + d, nkv, (int64_t)nctx * nkv, datt, honor_done ? seq_state : nullptr, Wl, hp.n_text_head, d, nctx, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.co.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(datt, d, L.co.w, d, d, d, e))) return rc;
            tr(1, 0);
            if ((rc = dec_ln<T>(dx, L.ln3.g, L.ln3.b, dh, Wl, d, nullptr, nullptr, nullptr, seq_state, sl))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.fc1.b; e.act = 1; e.out16 = dmlp; e.ldo16 = 4 * d;
            if ((rc = sk(dh, d, L.fc1.w, d, 4 * d, d, e))) return rc;
            e = SkinnyEpilogue{}; e.bias = L.fc2.b; e.residual = dx; e.ldr = d; e.out32 = dx; e.ldo32 = d;
            if ((rc = sk(dmlp, 4 * d, L.fc2.w, 4 * d, d, 4 * d, e))) return rc;
        }
        tr(1, 0);
        if ((rc = dec_ln<T>(dx, ln_f.g, ln_f.b, dh, Wl, d, nullptr, nullptr, nullptr, seq_state, sl))) return rc;
        // tied-embedding logits: 80-130 MB of weights per step -> the TMA-fed tcgen05 GEMM streams them
        // (one 128-row tile of sequences, ~200 column tiles) instead of the small-N weight-streaming kernel
        GemmEpilogue ge{logits, vpad, 1, nullptr, 0, nullptr, 0, 0};
        if ((rc = gemm_tn(dtype, dh, d, tok_emb, d, Wl, vpad, d, ge, sl))) return rc;
        if (detect_lang && (rc = lang_detect_step(logits, vpad, sa.prompt, sa.state, b_lang.as<int>() + w0, sp, Wl, sl))) return rc;
        if ((rc = sample_step(logits, vpad, sa, Wl, sl))) return rc;
        return SB_OK;
    }

    bool detect_lang = false;       // the current call asked for language auto-detect: k_lang_detect is part of the step

    // ---- slots / lanes of one call ----------------------------------------------------------
    struct DecodeCfg { SamplerArgs sa0; bool has_forced = false, graph = false, honor_done = true; int flags = 0; };
    DecodeCfg dcfg;

    int setup_slots(int S, const sb_params& p, int n_steps_cap, bool single_lane, bool has_forced, bool want_graph) {
        int n_max = hp.n_text_ctx / 2 - 4;
        if (p.n_max_tokens > 0) n_max = std::min(n_max, p.n_max_tokens);
        if (n_steps_cap > 0) n_max = std::min(n_max, n_steps_cap);
        int rc = ensure_decoder_ws(S, n_max);
        if (rc) return rc;
        n_slots = S; n_max_cur = n_max;
        n_lanes = single_lane ? 1 : std::min(n_lanes_cfg, std::max(1, S / 8));
        for (int i = 0; i < n_lanes; ++i) {
            Lane& L = lanes[i];
            L.s0 = (int)((int64_t)S * i / n_lanes);
            L.n = (int)((int64_t)S * (i + 1) / n_lanes) - L.s0;
            L.active = 0; L.steps = 0;
        }
        // every slot starts empty: done = 1, and a valid token / position so that the row-wise stages stay in bounds
        SeqState* hs = h_state.as<SeqState>();
        for (int s_ = 0; s_ < S; ++s_) { SeqState z{}; z.done = 1; z.n_prompt = 1; z.lang_slot = -1; hs[s_] = z; }
        SB_CUDA_CHECK(cudaMemcpyAsync(b_state.p, hs, S * sizeof(SeqState), cudaMemcpyHostToDevice, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_tick.p, 0, 64 * kMaxLanes, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_tokens.p, 0xff, (size_t)S * n_max * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_margins.p, 0, (size_t)S * n_max * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_tids.p, 0, (size_t)S * n_max * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_plogs.p, 0, (size_t)S * n_max * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_temp.p, 0, (size_t)S * 4, st));
        SB_CUDA_CHECK(cudaMemsetAsync(b_lang.p, 0xff, (size_t)S * 4, st));
        k_fill_i32<<<ceil_div(S, 256), 256, 0, st>>>(b_next.as<int>(), S, sp.sot);
        g_launches += 1;
        SB_CUDA_CHECK(cudaStreamSynchronize(st));        // h_state is reused as the polling mirror
        SamplerArgs sa0{};
        sa0.sp = sp; sa0.n_vocab = hp.n_vocab; sa0.n_max = n_max; sa0.n_text_ctx = hp.n_text_ctx;
        sa0.suppress_blank = p.suppress_blank; sa0.no_timestamps = p.no_timestamps; sa0.single_segment = p.single_segment;
        sa0.max_initial_tid = p.max_initial_ts > 0.f ? (int)lroundf(p.max_initial_ts / (30.0f / hp.n_audio_ctx)) : -1;
        sa0.nst_ids = (p.suppress_nst && n_nst > 0) ? b_nst.as<int>() : nullptr;
        sa0.n_nst = n_nst;
        dcfg.sa0 = sa0; dcfg.has_forced = has_forced; dcfg.graph = want_graph; dcfg.honor_done = !has_forced;
        const bool tracing = profile == 2 && want_graph;
        dcfg.flags = (p.suppress_blank ? 1 : 0) | (p.no_timestamps ? 2 : 0) | (p.single_segment ? 4 : 0) | (detect_lang ? 8 : 0) | (tracing ? 16 : 0) |
                     (p.suppress_nst ? 32 : 0);
        // launch trace: [lane][start | end][step][launch] globaltimer stamps, reset per call
        trace_per_step = 0;
        if (tracing) {
            trace_per_step = 11 * hp.n_text_layer + 1;
            trace_max_steps = 4096;
            const size_t n = (size_t)trace_max_steps * trace_per_step;
            if ((rc = b_trace.ensure((size_t)kMaxLanes * 2 * n * 8))) return rc;
            for (int i = 0; i < kMaxLanes; ++i) {
                SB_CUDA_CHECK(cudaMemsetAsync(b_trace.as<unsigned long long>() + (size_t)i * 2 * n, 0xff, n * 8, st));
                SB_CUDA_CHECK(cudaMemsetAsync(b_trace.as<unsigned long long>() + (size_t)i * 2 * n + n, 0, n * 8, st));
            }
        }
        // fork: every lane stream waits for the setup work queued on the main stream
        SB_CUDA_CHECK(cudaEventRecord(ev_fork, st));
        for (int i = 0; i < n_lanes; ++i) SB_CUDA_CHECK(cudaStreamWaitEvent(lanes[i].st, ev_fork, 0));
        if (want_graph)
            for (int i = 0; i < n_lanes; ++i) if ((rc = ensure_graph(i))) return rc;
        return SB_OK;
    }

    SamplerArgs lane_args(int li) const {
        SamplerArgs sa = dcfg.sa0;
        const int w0 = lanes[li].s0;
        sa.state = b_state.as<SeqState>() + w0;
        sa.tokens_out = b_tokens.as<int>() + (size_t)w0 * n_max_cur;
        sa.margins_out = b_margins.as<float>() + (size_t)w0 * n_max_cur;
        sa.tids_out = b_tids.as<int>() + (size_t)w0 * n_max_cur;
        sa.plogs_out = b_plogs.as<float>() + (size_t)w0 * n_max_cur;
        sa.temperature = b_temp.as<float>() + w0;
        sa.rng_u = b_rng.as<double>() + (size_t)w0 * n_max_cur;
        sa.next_tokens = b_next.as<int>() + w0;
        sa.forced = dcfg.has_forced ? b_forced.as<int>() + (size_t)w0 * n_max_cur : nullptr;
        sa.tick = b_tick.as<int>() + 16 * li;
        sa.prompt = b_prompt.as<int>() + (size_t)w0 * kMaxPrompt;
        return sa;
    }

    int ensure_graph(int li) {
        Lane& L = lanes[li];
        GraphKey k{n_slots, L.s0, L.n, n_max_cur, dcfg.flags, dcfg.sa0.max_initial_tid, dcfg.has_forced ? 1 : 0, g_ws_gen.load()};
        if (L.gexec && k == L.key) return SB_OK;
        if (L.gexec) { cudaGraphExecDestroy(L.gexec); L.gexec = nullptr; }
        cudaGraph_t g = nullptr;
        const uint64_t l0 = g_launches.load();
        SB_CUDA_CHECK(cudaStreamBeginCapture(L.st, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_step(li, lane_args(li), dcfg.honor_done);
        cudaError_t ce = cudaStreamEndCapture(L.st, &g);
        L.graph_nodes = (int)(g_launches.load() - l0);
        g_launches -= (uint64_t)L.graph_nodes;      // captured, not executed
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) { set_error(std::string("graph capture: ") + cudaGetErrorString(ce)); return SB_ERR_CUDA; }
        ce = cudaGraphInstantiate(&L.gexec, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { L.gexec = nullptr; set_error(std::string("graph instantiate: ") + cudaGetErrorString(ce)); return SB_ERR_CUDA; }
        L.key = k;
        return SB_OK;
    }

    int run_steps(int li, int k) {
        Lane& L = lanes[li];
        int rc;
        for (int b = 0; b < k; ++b) {
            if (dcfg.graph) { SB_CUDA_CHECK(cudaGraphLaunch(L.gexec, L.st)); g_launches += (uint64_t)L.graph_nodes; }
            else if ((rc = enqueue_step(li, lane_args(li), dcfg.honor_done))) return rc;
        }
        L.steps += k;
        stats.decoder_steps += (double)k * L.n / n_slots;
        return SB_OK;
    }

    // D2H of the lane's state / token rows into the pinned mirrors (async on the lane's stream)
    int poll_lane(int li) {
        const Lane& L = lanes[li];
        const size_t o = (size_t)L.s0 * n_max_cur, nb = (size_t)L.n * n_max_cur * 4;
        SB_CUDA_CHECK(cudaMemcpyAsync(h_state.as<SeqState>() + L.s0, b_state.as<SeqState>() + L.s0, L.n * sizeof(SeqState), cudaMemcpyDeviceToHost, L.st));
        SB_CUDA_CHECK(cudaMemcpyAsync(h_tokens.as<int>() + o, b_tokens.as<int>() + o, nb, cudaMemcpyDeviceToHost, L.st));
        SB_CUDA_CHECK(cudaMemcpyAsync(h_margins.as<float>() + o, b_margins.as<float>() + o, nb, cudaMemcpyDeviceToHost, L.st));
        SB_CUDA_CHECK(cudaMemcpyAsync(h_tids.as<int>() + o, b_tids.as<int>() + o, nb, cudaMemcpyDeviceToHost, L.st));
        SB_CUDA_CHECK(cudaMemcpyAsync(h_plogs.as<float>() + o, b_plogs.as<float>() + o, nb, cudaMemcpyDeviceToHost, L.st));
        SB_CUDA_CHECK(cudaMemcpyAsync(h_lang.as<int>() + L.s0, b_lang.as<int>() + L.s0, L.n * 4, cudaMemcpyDeviceToHost, L.st));
        stats.d2h_bytes += (double)L.n * (sizeof(SeqState) + 4) + 4.0 * nb;
        return SB_OK;
    }

    // one window assigned to a decode slot
    struct WinJob {
        int clip = 0, slot = 0, seek = 0, seek_end = 0;
        std::vector<int> prompt;      // full decoder prompt (language slot = -1 when it is to be detected)
        int lang_slot = -1, restart = 0;
        int prefilled = 0;            // prompt tokens [0, prefilled) went through the batched prefill pass
        float temperature = 0.f;      // > 0: tokens are drawn with `rng_u` (n_max uniform numbers of the clip's stream)
        int attempt = 0;
        std::vector<double> rng_u;
    };

    // whisper_full's prompt of one window: [prev] + the last <= n_text_ctx/2 tokens of prompt_past (text context of this
    // call: initial_prompt tokens, then the kept tokens of the previous windows) + [sot, lang, task] (+ [notimestamps])
    void build_prompt(WinJob& j, const std::vector<int>& prompt_past, int lang, const sb_params& p) const {
        j.prompt.clear();
        const int n_ctx_max = p.n_max_text_ctx;
        if (!prompt_past.empty() && n_ctx_max > 0 && j.temperature < 0.5f) {       // whisper.cpp: t_cur < 0.5
            const int n_take = std::min(std::min(n_ctx_max, hp.n_text_ctx / 2), (int)prompt_past.size());
            j.prompt.push_back(sp.prev);
            j.prompt.insert(j.prompt.end(), prompt_past.end() - n_take, prompt_past.end());
        }
        const bool multilingual = hp.n_vocab >= 51865;
        const bool prefixed = !j.prompt.empty();
        j.prompt.push_back(sp.sot);
        j.lang_slot = -1; j.restart = 0;
        if (multilingual) {
            j.lang_slot = (int)j.prompt.size();
            j.prompt.push_back(lang >= 0 ? sp.lang_first + lang : -1);
            j.prompt.push_back(p.translate ? sp.translate : sp.transcribe);
        }
        if (p.no_timestamps) j.prompt.push_back(sp.not_);
        if (multilingual && lang < 0 && prefixed) {        // detect on [sot] alone, then restart at position 0
            j.prompt.insert(j.prompt.begin(), sp.sot);
            j.lang_slot += 1; j.restart = 1;
        }
    }

    // Prompt prefill.  A window that carries text context has a prompt of up to 229 tokens; fed one token per decoder step
    // it would hold its slot for as many steps before the first token is sampled.  Instead, all prompt tokens but the last
    // of every job of a batch go through the decoder ONCE as rows of the encoder's tcgen05 GEMMs (M = total prompt tokens),
    // with the decoder-step attention kernels in row mode (row -> (slot, position)); their K / V land in the slot's self-KV
    // cache and the step loop starts at the last prompt token.  Runs on the main stream right after the encoder batch.
    int prefill(std::vector<WinJob>& jobs) {
        if (prefill_min <= 0) return SB_OK;
        int R = 0;
        for (const WinJob& j : jobs)
            if (!j.restart && (int)j.prompt.size() - 1 >= prefill_min) R += (int)j.prompt.size() - 1;
        if (R == 0) return SB_OK;
        int rc;
        if ((rc = b_prow.ensure((size_t)3 * R * 4))) return rc;
        if ((rc = h_prow.ensure((size_t)3 * n_slots * kMaxPrompt * 4))) return rc;
        int* hr = h_prow.as<int>();
        int r = 0;
        for (WinJob& j : jobs) {
            if (j.restart || (int)j.prompt.size() - 1 < prefill_min) continue;
            j.prefilled = (int)j.prompt.size() - 1;
            for (int p_ = 0; p_ < j.prefilled; ++p_, ++r) { hr[r] = j.slot; hr[R + r] = p_; hr[2 * R + r] = j.prompt[p_]; }
        }
        SB_CUDA_CHECK(cudaMemcpyAsync(b_prow.p, hr, (size_t)3 * R * 4, cudaMemcpyHostToDevice, st));
        stats.h2d_bytes += 12.0 * R;
        const int* row_slot = b_prow.as<int>();
        const int* row_pos = row_slot + R;
        const int* row_tok = row_slot + 2 * R;
        const int d = hp.n_text_state, nctx = hp.n_audio_ctx, S = n_slots;
        const int nkv = hp.n_text_layer * 2 * d;
        float* x = b_x.as<float>();
        T* h = b_h.as<T>(); T* qkv = b_qkv.as<T>(); T* att = b_att.as<T>(); T* mlp = b_mlp.as<T>();
        if ((rc = prefill_embed<T>(tok_emb, dec_pos, row_tok, row_pos, x, R, d, st))) return rc;
        for (int l = 0; l < hp.n_text_layer; ++l) {
            const DecLayer<T>& L = dec[l];
            T* kc = b_kself.as<T>() + (int64_t)l * S * hp.n_text_ctx * d;
            T* vc = b_vself.as<T>() + (int64_t)l * S * hp.n_text_ctx * d;
            if ((rc = layernorm<T>(x, L.ln1.g, L.ln1.b, h, nullptr, R, d, st))) return rc;
            GemmEpilogue ep{qkv, 3 * d, 0, L.qkv.b, 0, nullptr, 0, 0};
            if ((rc = gemm_tn(dtype, h, d, L.qkv.w, d, R, 3 * d, d, ep, st))) return rc;
            if ((rc = prefill_kv_scatter<T>(qkv, row_slot, row_pos, kc, vc, R, d, hp.n_text_ctx, st))) return rc;
            if ((rc = dec_self_attn<T>(qkv, kc, vc, att, nullptr, 0, R, hp.n_text_head, d, hp.n_text_ctx, st, row_slot, row_pos))) return rc;
            ep = GemmEpilogue{x, d, 1, L.o.b, 0, x, d, 0};
            if ((rc = gemm_tn(dtype, att, d, L.o.w, d, R, d, d, ep, st))) return rc;
            if ((rc = layernorm<T>(x, L.ln2.g, L.ln2.b, h, nullptr, R, d, st))) return rc;
            ep = GemmEpilogue{qkv, d, 0, L.cq.b, 0, nullptr, 0, 0};
            if ((rc = gemm_tn(dtype, h, d, L.cq.w, d, R, d, d, ep, st))) return rc;
            const T* kb = b_ckv.as<T>() + (int64_t)l * 2 * d;
            if ((rc = dec_cross_attn<T>(qkv, d, kb, kb + d, nkv, (int64_t)nctx * nkv, att, nullptr, R, hp.n_text_head, d, nctx, st, row_slot))) return rc;
            ep = GemmEpilogue{x, d, 1, L.co.b, 0, x, d, 0};
            if ((rc = gemm_tn(dtype, att, d, L.co.w, d, R, d, d, ep, st))) return rc;
            if ((rc = layernorm<T>(x, L.ln3.g, L.ln3.b, h, nullptr, R, d, st))) return rc;
            ep = GemmEpilogue{mlp, 4 * d, 0, L.fc1.b, 1, nullptr, 0, 0};
            if ((rc = gemm_tn(dtype, h, d, L.fc1.w, d, R, 4 * d, d, ep, st))) return rc;
            ep = GemmEpilogue{x, d, 1, L.fc2.b, 0, x, d, 0};
            if ((rc = gemm_tn(dtype, mlp, 4 * d, L.fc2.w, 4 * d, R, d, 4 * d, ep, st))) return rc;
        }
        stats.prefill_rows += R;
        return SB_OK;
    }

    // scatter the jobs into their slots: staged per lane in pinned memory, copied and applied on the lane's own stream
    // (in order with the lane's steps, after the encoder's cross-KV for these slots is complete: ev_enc)
    int init_slots(const std::vector<WinJob>& jobs, cudaEvent_t wait_ev) {
        SlotInit* hi = h_init.as<SlotInit>();
        int rc;
        for (int li = 0; li < n_lanes; ++li) {
            Lane& L = lanes[li];
            int cnt = 0;
            for (const WinJob& j : jobs) {
                if (j.slot < L.s0 || j.slot >= L.s0 + L.n) continue;
                SlotInit& it = hi[L.s0 + cnt];
                memset(&it, 0, sizeof(it));
                it.slot = j.slot;
                it.temperature = j.temperature;
                if (j.temperature > 0.f) {          // this slot's uniform numbers: staged and copied on the lane's stream
                    double* hr = h_rng.as<double>() + (size_t)j.slot * n_max_cur;
                    for (int k = 0; k < n_max_cur; ++k) hr[k] = k < (int)j.rng_u.size() ? j.rng_u[k] : 0.5;
                    SB_CUDA_CHECK(cudaMemcpyAsync(b_rng.as<double>() + (size_t)j.slot * n_max_cur, hr, (size_t)n_max_cur * 8,
                                                  cudaMemcpyHostToDevice, L.st));
                }
                it.next_token = j.prompt[0];
                SeqState s{};
                s.seek_delta = 3000; s.seek = j.seek; s.seek_end = j.seek_end;
                s.n_prompt = (int)j.prompt.size(); s.lang_slot = j.lang_slot; s.restart = j.restart;
                s.idx = s.pos = j.prefilled;           // the step loop continues after the prefilled part of the prompt
                it.next_token = j.prompt[j.prefilled];
                it.state = s;
                for (size_t k = 0; k < j.prompt.size(); ++k) it.prompt[k] = j.prompt[k];
                ++cnt;
            }
            if (!cnt) continue;
            if (wait_ev) SB_CUDA_CHECK(cudaStreamWaitEvent(L.st, wait_ev, 0));
            SB_CUDA_CHECK(cudaMemcpyAsync(b_init.as<SlotInit>() + L.s0, hi + L.s0, cnt * sizeof(SlotInit), cudaMemcpyHostToDevice, L.st));
            if ((rc = slot_init(b_init.as<SlotInit>() + L.s0, cnt, b_state.as<SeqState>(), b_next.as<int>(), b_prompt.as<int>(),
                                b_lang.as<int>(), b_temp.as<float>(), L.st))) return rc;
            L.active += cnt;
            stats.h2d_bytes += (double)cnt * sizeof(SlotInit);
        }
        return SB_OK;
    }

    // join: the main stream continues after every lane
    int join_lanes() {
        for (int i = 0; i < n_lanes; ++i) {
            SB_CUDA_CHECK(cudaEventRecord(lanes[i].done, lanes[i].st));
            SB_CUDA_CHECK(cudaStreamWaitEvent(st, lanes[i].done, 0));
        }
        return SB_OK;
    }

    // profile == 2: fold the device-side launch trace of this call into the stats (stream idle)
    int collect_trace() {
        if (!(profile == 2 && trace_per_step > 0)) return SB_OK;
        const size_t n = (size_t)trace_max_steps * trace_per_step;
        std::vector<unsigned long long> ht((size_t)n_lanes * 2 * n);
        SB_CUDA_CHECK(cudaMemcpy(ht.data(), b_trace.p, ht.size() * 8, cudaMemcpyDeviceToHost));
        const bool dump = getenv("SB_TRACE_DUMP") != nullptr;
        for (int i = 0; i < n_lanes; ++i) {
            const unsigned long long* t0 = ht.data() + (size_t)i * 2 * n;
            const unsigned long long* t1 = t0 + n;
            for (int sI = 0; sI < lanes[i].steps && sI < trace_max_steps; ++sI) {
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
                if (last > first && dump) {    // per-index timeline relative to the lane-step's first start
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
        if (const char* path = getenv("SB_TRACE_DUMP")) {
            if (FILE* f = fopen(path, "w")) {
                fprintf(f, "idx,class,work_bytes,count,avg_us,avg_start_us,avg_end_us\n");
                for (int k = 0; k < (int)trace_cnt.size() && k < (int)trace_cls.size(); ++k)
                    if (trace_cnt[k] > 0)
                        fprintf(f, "%d,%d,%.0f,%.0f,%.3f,%.3f,%.3f\n", k, trace_cls[k], trace_work[k], trace_cnt[k],
                                trace_sum_dur[k] / trace_cnt[k] * 1e-3, trace_sum_t0[k] / trace_cnt[k] * 1e-3, trace_sum_t1[k] / trace_cnt[k] * 1e-3);
                fclose(f);
            }
        }
        return SB_OK;
    }

    // language / initial prompt of a call.  *lang = -1: detect (reference default "auto")
    int resolve_language(const sb_params& p, int* lang) {
        // NULL / "" / "auto": whisper_full detects the language on the first window (multilingual models);
        // English-only models have no language token at all
        if (!p.language || !p.language[0] || std::string(p.language) == "auto") { *lang = hp.n_vocab >= 51865 ? -1 : 0; return SB_OK; }
        std::string l = p.language;
        if (l == "zh-Hans" || l == "zh-Hant") l = "zh";   // reference: transcription.rs:448-459
        const int id = lang_id(l.c_str());
        if (id < 0 || (hp.n_vocab >= 51865 && id >= sp.num_languages)) { set_error("unknown language code: " + l); return SB_ERR_INVALID; }
        if (hp.n_vocab < 51865 && id != 0) { set_error("English-only model: language must be \"en\" or auto, got " + l); return SB_ERR_INVALID; }
        *lang = id;
        return SB_OK;
    }
    // WhisperInferenceParams::initial_prompt (transcription.rs:461-499) -> whisper.cpp tokenises it into prompt_past
    std::vector<int> initial_prompt_tokens(const sb_params& p) const {
        if (!p.initial_prompt || !p.initial_prompt[0]) return {};
        return tokenize(p.initial_prompt);
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
        int rc = ensure_encoder_ws(W, true);
        if (rc) return rc;
        if ((rc = b_ckv.ensure((size_t)W * hp.n_audio_ctx * hp.n_text_layer * 2 * hp.n_text_state * 2))) return rc;
        if ((rc = stage_mel_windows(mel_windows, W))) return rc;
        SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
        if ((rc = encode_chunk(W, nullptr, b_enc32.as<float>()))) return rc;
        SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
        SB_CUDA_CHECK(cudaMemcpyAsync(enc_out, b_enc32.p, (size_t)W * hp.n_audio_ctx * hp.n_audio_state * 4, cudaMemcpyDeviceToHost, st));
        SB_CUDA_CHECK(cudaStreamSynchronize(st));
        { float t = 0.f; cudaEventElapsedTime(&t, ev[0], ev[1]); stats.encode_ms += t; stats.windows += W; }   // conv stem .. ln_post + cross-KV
        prof_collect();
        return SB_OK;
    }

    // parity hook: encode W windows, then n_steps decoder steps from the window prompt, optionally teacher-forced, with
    // the raw logits of every sampling step copied out
    int decode_trace(const float* mel_windows, int W, const int32_t* seek_end, const sb_params& p, const int32_t* forced,
                     int n_steps, float* logits_out, int32_t* tokens_out, float* margins_out) override {
        SB_CHECK_ARG(mel_windows && seek_end && tokens_out && W > 0 && W <= max_batch && n_steps > 0, "sb_decode_trace: bad arguments");
        SB_CHECK_ARG(n_steps <= hp.n_text_ctx / 2 - 4 && (p.n_max_tokens <= 0 || p.n_max_tokens >= n_steps), "n_steps exceeds n_text_ctx/2 - 4");
        SB_CUDA_CHECK(cudaSetDevice(device));
        int lang = 0;
        int rc = resolve_language(p, &lang);
        if (rc) return rc;
        detect_lang = lang < 0;
        const std::vector<int> past = initial_prompt_tokens(p);
        if ((rc = ensure_encoder_ws(W, false))) return rc;
        if ((rc = setup_slots(W, p, n_steps, logits_out != nullptr, forced != nullptr, use_graph && !logits_out))) return rc;
        if ((rc = stage_mel_windows(mel_windows, W))) return rc;
        if ((rc = encode_chunk(W, nullptr, nullptr))) return rc;
        if (forced) SB_CUDA_CHECK(cudaMemcpyAsync(b_forced.p, forced, (size_t)W * n_steps * 4, cudaMemcpyHostToDevice, st));
        std::vector<WinJob> jobs(W);
        for (int w = 0; w < W; ++w) {
            jobs[w].clip = w; jobs[w].slot = w; jobs[w].seek = 0; jobs[w].seek_end = seek_end[w];
            build_prompt(jobs[w], past, lang, p);
        }
        if ((rc = prefill(jobs))) return rc;
        SB_CUDA_CHECK(cudaEventRecord(ev_enc, st));
        if ((rc = init_slots(jobs, ev_enc))) return rc;
        const int feed = (int)jobs[0].prompt.size() - 1 - jobs[0].prefilled;         // the prompt is the same for every window here
        const int total_steps = feed + n_steps;
        const int vpad = (int)round_up(hp.n_vocab, 8);
        if (logits_out) {
            for (int s_ = 0; s_ < total_steps; ++s_) {
                if ((rc = run_steps(0, 1))) return rc;
                if (s_ >= feed)
                    for (int w = 0; w < W; ++w)
                        SB_CUDA_CHECK(cudaMemcpyAsync(logits_out + ((size_t)w * n_steps + (s_ - feed)) * hp.n_vocab,
                                                      b_logits.as<float>() + (size_t)w * vpad, (size_t)hp.n_vocab * 4,
                                                      cudaMemcpyDeviceToHost, lanes[0].st));
            }
        } else {
            for (int li = 0; li < n_lanes; ++li) if ((rc = run_steps(li, total_steps))) return rc;
        }
        for (int li = 0; li < n_lanes; ++li) if ((rc = poll_lane(li))) return rc;
        if ((rc = join_lanes())) return rc;
        SB_CUDA_CHECK(cudaStreamSynchronize(st));
        prof_collect();
        memcpy(tokens_out, h_tokens.p, (size_t)W * n_steps * 4);
        if (margins_out) memcpy(margins_out, h_margins.p, (size_t)W * n_steps * 4);
        return SB_OK;
    }

    // ---- whisper_full over a group of clips ---------------------------------------------------
    struct Segment { int64_t t0, t1; std::string text; int tok_off, n_tok; };
    struct ClipRun {
        size_t n = 0; int n_len = 0, n_len_org = 0, n_calc = 0;
        int seek = 0; bool active = false, running = false; int lang = 0;
        std::vector<int> prompt_past;         // whisper_full's prompt_past of this call
        Mt19937 rng{0};                       // the decoder's std::mt19937(0): consumed by the draws at temperature > 0
        std::vector<int32_t> kept, sampled, tids; std::vector<float> margins, plogs; std::vector<sb_window_info> windows;
        std::vector<Segment> segments;
        std::string text;
    };

    // one finished window: whisper_full's per-window epilogue (tokens_cur.resize(result_len), segments at timestamp
    // tokens, prompt_past update, seek += seek_delta)
    void finish_window(ClipRun& r, const WinJob& j, const SeqState& s, const int* toks, const float* margs, const int* tids,
                       const float* plogs, float avg_logprob, const sb_params& p) {
        sb_window_info wi{};
        wi.temperature = j.temperature; wi.n_attempts = j.attempt + 1; wi.avg_logprob = avg_logprob;
        wi.seek = r.seek; wi.n_tokens = s.n_tok; wi.result_len = s.result_len; wi.seek_delta = s.seek_delta;
        wi.failed = s.failed; wi.token_offset = (int)r.sampled.size();
        wi.n_prompt = (int)j.prompt.size() - j.restart;
        const int kept0 = (int)r.kept.size();
        for (int i = 0; i < s.n_tok; ++i) {
            r.sampled.push_back(toks[i]); r.margins.push_back(margs[i]); r.tids.push_back(tids[i]); r.plogs.push_back(plogs[i]);
            if (i < s.result_len) { r.kept.push_back(toks[i]); if (toks[i] < sp.eot) r.text += token_text(toks[i]); }
        }
        r.windows.push_back(wi);
        // segments (whisper_full_with_state): text between timestamp tokens; the first segment starts at the most
        // probable timestamp of the first token (tid), an unterminated tail ends at seek + seek_delta
        const int n = std::min(s.result_len, s.n_tok);
        if (n > 0) {
            int i0 = 0;
            int64_t t0 = r.seek + 2 * (int64_t)(tids[0] - sp.beg);
            std::string text;
            for (int i = 0; i < n; ++i) {
                if (toks[i] < sp.eot) text += token_text(toks[i]);
                if (toks[i] > sp.beg && !p.single_segment) {
                    const int64_t t1 = r.seek + 2 * (int64_t)(tids[i] - sp.beg);
                    if (!text.empty()) r.segments.push_back(Segment{t0, t1, text, kept0 + i0, i - i0 + 1});
                    text.clear();
                    while (i < n && toks[i] > sp.beg) ++i;
                    --i;
                    t0 = t1; i0 = i + 1;
                }
            }
            if (!text.empty()) r.segments.push_back(Segment{t0, (int64_t)r.seek + s.seek_delta, text, kept0 + i0, n - i0});
        }
        // prompt_past = the context this window actually used + its kept tokens
        std::vector<int> np;
        const int lead = j.restart;
        if ((int)j.prompt.size() > lead && j.prompt[lead] == sp.prev) {
            const int n_init = 1 + (hp.n_vocab >= 51865 ? 2 : 0) + (p.no_timestamps ? 1 : 0);
            np.assign(j.prompt.begin() + lead + 1, j.prompt.end() - n_init);
        }
        for (int i = 0; i < n; ++i) np.push_back(toks[i]);
        r.prompt_past.swap(np);
        r.seek += s.seek_delta;
        stats.tokens_sampled += (double)s.n_tok;
        if (profile == 2)   // cross-attention bytes: the sequence was live for every fed prompt token and every sampled token
            stats.xattn_bytes += (double)((int)j.prompt.size() + std::max(s.n_tok, 1) - 1) * hp.n_text_layer * 4.0 * hp.n_audio_ctx * hp.n_text_state;
    }

    int run_group(const float* const* pcm, const size_t* ns, int G, const sb_params& p, int lang, sb_result* out) {
        const int n_mel = hp.n_mels;
        std::vector<ClipRun> clips(G);
        const std::vector<int> init_past = initial_prompt_tokens(p);
        size_t max_n = 0; int max_calc = 0; int n_active = 0;
        for (int c = 0; c < G; ++c) {
            ClipRun& r = clips[c];
            r.n = ns[c];
            r.lang = lang;
            r.prompt_past = init_past;
            if (r.n == 0) continue;
            sb_logmel_geometry(r.n, &r.n_len, &r.n_len_org, &r.n_calc);
            // whisper.cpp: "input is too short" below 1 s -> no segments
            r.active = r.n >= 201 && r.n_len_org >= 100;
            if (!r.active) continue;
            ++n_active;
            max_n = std::max(max_n, r.n); max_calc = std::max(max_calc, r.n_calc);
        }
        float ms_mel = 0.f, ms_enc = 0.f, ms_total = 0.f;
        int rc;
        if (max_n > 0) {
            const int stride = (int)round_up(max_calc, 32);
            if ((rc = b_mel.ensure((size_t)G * n_mel * stride * 4))) return rc;
            if ((rc = b_cmax.ensure(G * 4))) return rc;
            if ((rc = b_floor.ensure(G * 4))) return rc;
            if ((rc = b_clipmeta.ensure(G * 3 * 4))) return rc;
            if ((rc = b_pcmptr.ensure(G * sizeof(float*)))) return rc;
            const int S = std::min(max_batch, n_active);
            if ((rc = b_winmeta.ensure(std::max(G, S) * 2 * 4))) return rc;
            if ((rc = h_winmeta.ensure(S * 2 * 4))) return rc;
            if ((rc = ensure_encoder_ws(S, false))) return rc;
            // clips that already live in this device's memory are read in place; host clips are staged (H2D)
            std::vector<char> on_device(G, 0);
            bool any_host = false;
            for (int c = 0; c < G; ++c) {
                if (!clips[c].active) continue;
                cudaPointerAttributes attr{};
                if (cudaPointerGetAttributes(&attr, pcm[c]) == cudaSuccess)
                    on_device[c] = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) && attr.device == device;
                else cudaGetLastError();
                any_host = any_host || !on_device[c];
            }
            if (any_host && (rc = b_pcm.ensure((size_t)G * max_n * 4))) return rc;
            SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
            std::vector<int> meta(3 * G, 0);             // n_calc | n_len | n_samples
            std::vector<const float*> ptrs(G, nullptr);
            for (int c = 0; c < G; ++c) {
                if (!clips[c].active) continue;
                if (on_device[c]) ptrs[c] = pcm[c];
                else {
                    ptrs[c] = b_pcm.as<float>() + (size_t)c * max_n;
                    SB_CUDA_CHECK(cudaMemcpyAsync(b_pcm.as<float>() + (size_t)c * max_n, pcm[c], clips[c].n * 4, cudaMemcpyHostToDevice, st));
                }
                stats.pcm_bytes += clips[c].n * 4.0;
                meta[c] = clips[c].n_calc; meta[G + c] = clips[c].n_len; meta[2 * G + c] = (int)clips[c].n;
            }
            SB_CUDA_CHECK(cudaMemcpyAsync(b_clipmeta.p, meta.data(), 3 * G * 4, cudaMemcpyHostToDevice, st));
            SB_CUDA_CHECK(cudaMemcpyAsync(b_pcmptr.p, ptrs.data(), G * sizeof(float*), cudaMemcpyHostToDevice, st));
            // ONE log-mel launch for the whole (ragged) group; raw log10 values, normalised by the conv1 im2col
            if ((rc = logmel_launch_ragged(melplan, b_pcmptr.as<const float*>(), b_clipmeta.as<int>() + 2 * G, b_clipmeta.as<int>(), G,
                                           max_calc, b_mel.as<float>(), (int64_t)n_mel * stride, stride, b_cmax.as<int32_t>(), st))) return rc;
            SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
            SB_CUDA_CHECK(cudaStreamSynchronize(st));   // meta vector lifetime + mel timing
            { float t; cudaEventElapsedTime(&t, ev[0], ev[1]); ms_mel += t; }

            // ---- the seek loop of every clip, scheduled over S decode slots -------------------------------------
            // A clip's next window becomes READY when its previous one ends (seek and text context depend on it).  Ready
            // windows are encoded in batches on the main stream while the lanes keep stepping the live slots; a batch
            // is started once `refill_min` slots are free (or nothing is running), and its windows enter their slots
            // as soon as the encoder is done.  Sequences are independent, so the schedule does not change any result.
            if ((rc = setup_slots(S, p, 0, false, false, use_graph != 0))) return rc;
            const int refill_min = refill_min_cfg > 0 ? refill_min_cfg : std::max(1, S / 8);
            const int max_windows = p.max_windows > 0 ? p.max_windows : 1 << 20;
            std::vector<int> slot_job(S, -1);             // index into `live` jobs, or -1
            std::vector<WinJob> live(S);
            std::vector<WinJob> pending;                  // encoded (or encoding) on the main stream, not yet in their slots
            bool pend_active = false;
            int running = 0;
            SB_CUDA_CHECK(cudaEventRecord(ev[2], st));     // (ev[1] / ev_enc bracket one encoder batch at a time)
            for (long iter = 0; iter < (1L << 24); ++iter) {
                // (1) windows whose encoder pass is complete enter their slots
                if (pend_active && (running == 0 || cudaEventQuery(ev_enc) == cudaSuccess)) {
                    if (running == 0) SB_CUDA_CHECK(cudaEventSynchronize(ev_enc));      // nothing to step meanwhile
                    { float t = 0.f; if (cudaEventElapsedTime(&t, ev[1], ev_enc) == cudaSuccess) ms_enc += t; }
                    if ((rc = init_slots(pending, ev_enc))) return rc;
                    for (const WinJob& j : pending) { live[j.slot] = j; slot_job[j.slot] = 1; }
                    running += (int)pending.size();
                    pending.clear(); pend_active = false;
                }
                // (2) start encoding the next batch of ready windows
                if (!pend_active) {
                    std::vector<int> ready, freeslots;
                    for (int c = 0; c < G; ++c) {
                        ClipRun& r = clips[c];
                        if (!r.active || r.running) continue;
                        if (r.seek + 100 >= r.n_len_org || (int)r.windows.size() >= max_windows) { r.active = false; continue; }
                        ready.push_back(c);
                    }
                    for (int s_ = 0; s_ < S; ++s_) if (slot_job[s_] < 0) freeslots.push_back(s_);
                    int k = (int)std::min(ready.size(), freeslots.size());
                    if (enc_batch_max > 0) k = std::min(k, enc_batch_max);
                    if (k > 0 && (running == 0 || k >= std::min(refill_min, enc_batch_max > 0 ? enc_batch_max : refill_min))) {
                        // spread the new windows over the lanes with the fewest live sequences
                        std::vector<int> lane_load(n_lanes);
                        for (int li = 0; li < n_lanes; ++li) lane_load[li] = lanes[li].active;
                        std::vector<int> chosen;
                        std::vector<char> used(S, 0);
                        for (int i = 0; i < k; ++i) {
                            int best = -1, best_lane = -1;
                            for (int s_ : freeslots) {
                                if (used[s_]) continue;
                                int li = 0;
                                while (s_ >= lanes[li].s0 + lanes[li].n) ++li;
                                if (best < 0 || lane_load[li] < lane_load[best_lane]) { best = s_; best_lane = li; }
                            }
                            used[best] = 1; lane_load[best_lane] += 1; chosen.push_back(best);
                        }
                        std::sort(chosen.begin(), chosen.end());
                        // pinned: the previous batch (the only other user) has left the encoder before a new one is staged
                        int* wm = h_winmeta.as<int>();
                        std::vector<int> slots(k);
                        for (int i = 0; i < k; ++i) {
                            ClipRun& r = clips[ready[i]];
                            // whisper.cpp: a very short tail drops the text context ("it tends to confuse the decoder")
                            if (r.seek > 0 && r.seek + 500 >= r.n_len_org) r.prompt_past.clear();
                            WinJob j;
                            j.clip = ready[i]; j.slot = chosen[i]; j.seek = r.seek; j.seek_end = r.n_len_org;
                            j.temperature = std::max(0.f, p.temperature);
                            if (j.temperature > 0.f) {
                                Mt19937 peek = r.rng;
                                j.rng_u.resize(n_max_cur);
                                for (double& u : j.rng_u) u = peek.uniform();
                            }
                            build_prompt(j, r.prompt_past, r.lang, p);
                            r.running = true;
                            pending.push_back(std::move(j));
                            wm[i] = ready[i]; wm[k + i] = r.seek; slots[i] = chosen[i];
                        }
                        SB_CUDA_CHECK(cudaEventRecord(ev[1], st));
                        SB_CUDA_CHECK(cudaMemcpyAsync(b_winmeta.p, wm, 2 * k * 4, cudaMemcpyHostToDevice, st));
                        Im2col1Args a{b_mel.as<float>(), b_floor.as<float>(), b_winmeta.as<int>(), b_winmeta.as<int>() + k,
                                      b_clipmeta.as<int>(), b_clipmeta.as<int>() + G, (int64_t)n_mel * stride, stride, n_mel,
                                      2 * hp.n_audio_ctx, b_cmax.as<int32_t>()};        // raw mel: normalised on the fly
                        if ((rc = im2col_conv1<T>(a, b_col1.as<T>(), k, st))) return rc;
                        if ((rc = encode_chunk(k, slots.data(), nullptr))) return rc;
                        if ((rc = prefill(pending))) return rc;
                        SB_CUDA_CHECK(cudaEventRecord(ev_enc, st));
                        pend_active = true;
                        stats.windows += k; stats.rounds += 1;
                        if (running == 0) continue;            // nothing to step meanwhile: enter the slots right away
                    }
                }
                if (running == 0 && !pend_active) break;
                if (running == 0) continue;
                // (3) a burst of steps on every lane with live sequences, then poll
                for (int li = 0; li < n_lanes; ++li) {
                    if (lanes[li].active == 0) continue;
                    if ((rc = run_steps(li, 8))) return rc;
                    if ((rc = poll_lane(li))) return rc;
                }
                for (int li = 0; li < n_lanes; ++li) {
                    Lane& L = lanes[li];
                    if (L.active == 0) continue;
                    SB_CUDA_CHECK(cudaStreamSynchronize(L.st));
                    const SeqState* hs = h_state.as<SeqState>();
                    std::vector<WinJob> retries;
                    for (int s_ = L.s0; s_ < L.s0 + L.n; ++s_) {
                        if (slot_job[s_] < 0 || !hs[s_].done) continue;
                        const WinJob& j = live[s_];
                        ClipRun& r = clips[j.clip];
                        if (r.lang < 0) { const int dl = h_lang.as<int>()[s_]; r.lang = dl >= 0 ? dl : 0; }     // detected on the clip's first window
                        const SeqState& fs = hs[s_];
                        const int* toks = h_tokens.as<int>() + (size_t)s_ * n_max_cur;
                        const float* plogs = h_plogs.as<float>() + (size_t)s_ * n_max_cur;
                        // whisper_sequence_score + the fallback test of whisper_full: mean log-probability of the kept tokens,
                        // entropy of the last 32 of them
                        const int n = std::min(fs.result_len, fs.n_tok);
                        double sum = 0.0;
                        for (int i = 0; i < n; ++i) sum += plogs[i];
                        const float avg = n > 0 ? (float)(sum / n) : NAN;
                        bool failed = fs.failed != 0;
                        if (n > 32) {
                            std::unordered_map<int, int> cnt;
                            for (int i = n - 32; i < n; ++i) cnt[toks[i]]++;
                            double ent = 0.0;
                            for (auto& kv : cnt) { const double pr = kv.second / 32.0; ent -= pr * std::log(pr); }
                            if (ent < p.entropy_thold) failed = true;
                        }
                        if (j.temperature > 0.f) for (int i = 0; i < fs.n_tok; ++i) r.rng.uniform();      // the draws this attempt consumed
                        const bool success = !(failed || avg < p.logprob_thold);
                        const float t_next = j.temperature + p.temperature_inc;
                        L.active -= 1; running -= 1;
                        if (!success && p.temperature_inc > 0.f && t_next < 1.0f + 1e-6f) {
                            // decode the same window again, one temperature up, in the same slot (its cross-KV is still there)
                            WinJob nj;
                            nj.clip = j.clip; nj.slot = s_; nj.seek = j.seek; nj.seek_end = j.seek_end;
                            nj.temperature = t_next; nj.attempt = j.attempt + 1;
                            build_prompt(nj, r.prompt_past, r.lang, p);
                            Mt19937 peek = r.rng;
                            nj.rng_u.resize(n_max_cur);
                            for (double& u : nj.rng_u) u = peek.uniform();
                            retries.push_back(std::move(nj));
                            stats.fallbacks += 1;
                            continue;
                        }
                        finish_window(r, j, fs, toks, h_margins.as<float>() + (size_t)s_ * n_max_cur, h_tids.as<int>() + (size_t)s_ * n_max_cur,
                                      plogs, avg, p);
                        r.running = false;
                        slot_job[s_] = -1;
                    }
                    if (!retries.empty()) {
                        cudaEvent_t wait_ev = nullptr;
                        size_t rows0 = (size_t)stats.prefill_rows;
                        if ((rc = prefill(retries))) return rc;
                        if ((size_t)stats.prefill_rows != rows0) { SB_CUDA_CHECK(cudaEventRecord(ev_fork, st)); wait_ev = ev_fork; }
                        if ((rc = init_slots(retries, wait_ev))) return rc;
                        for (WinJob& nj : retries) live[nj.slot] = std::move(nj);
                        running += (int)retries.size();
                    }
                }
            }
            if ((rc = join_lanes())) return rc;
            SB_CUDA_CHECK(cudaEventRecord(ev[0], st));
            SB_CUDA_CHECK(cudaStreamSynchronize(st));
            { float t; cudaEventElapsedTime(&t, ev[2], ev[0]); ms_total = t; }
            prof_collect();
            if ((rc = collect_trace())) return rc;
        }
        const float ms_dec = std::max(0.f, ms_total - ms_enc);
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
            auto dup = [](const void* src, size_t bytes) -> void* { void* q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, src, bytes); return q; };
            o.n_tokens = r.kept.size(); o.tokens = (int32_t*)dup(r.kept.data(), r.kept.size() * 4);
            o.n_sampled = r.sampled.size(); o.sampled = (int32_t*)dup(r.sampled.data(), r.sampled.size() * 4);
            o.margins = (float*)dup(r.margins.data(), r.margins.size() * 4);
            o.tids = (int32_t*)dup(r.tids.data(), r.tids.size() * 4);
            o.logprobs = (float*)dup(r.plogs.data(), r.plogs.size() * 4);
            o.n_windows = r.windows.size(); o.windows = (sb_window_info*)dup(r.windows.data(), r.windows.size() * sizeof(sb_window_info));
            // segments: one blob holds every segment text (NUL-terminated), the records point into it
            size_t blob = 0;
            for (const Segment& sg : r.segments) blob += sg.text.size() + 1;
            o.n_segments = r.segments.size();
            o.segments = (sb_segment*)malloc(std::max<size_t>(1, r.segments.size()) * sizeof(sb_segment));
            o.segment_text = (char*)malloc(std::max<size_t>(1, blob));
            size_t off = 0;
            for (size_t i = 0; i < r.segments.size(); ++i) {
                const Segment& sg = r.segments[i];
                memcpy(o.segment_text + off, sg.text.data(), sg.text.size());
                o.segment_text[off + sg.text.size()] = 0;
                o.segments[i] = sb_segment{sg.t0, sg.t1, o.segment_text + off, sg.text.size(), sg.tok_off, sg.n_tok};
                off += sg.text.size() + 1;
            }
            o.ms_mel = ms_mel; o.ms_encode = ms_enc; o.ms_decode = ms_dec; o.status = SB_OK;
            o.lang_id = hp.n_vocab >= 51865 ? r.lang : -1;
        }
        stats.mel_ms += ms_mel; stats.encode_ms += ms_enc; stats.decode_ms += ms_dec; stats.clips += G;
        return SB_OK;
    }

    int transcribe_batch(const float* const* pcm, const size_t* ns, size_t count, const sb_params& p, sb_result* out) override {
        SB_CUDA_CHECK(cudaSetDevice(device));
        int lang = 0;
        int rc = resolve_language(p, &lang);
        if (rc) return rc;
        detect_lang = lang < 0;
        // all clips of the call share the decode slots: a finished clip's slot goes to the next waiting clip
        return run_group(pcm, ns, (int)count, p, lang, out);
    }
};

}  // namespace sb

// One replica per CUDA device.  replicas[0] answers the single-device queries (info, stream, stats, tokenize).
struct sb_engine {
    std::vector<std::unique_ptr<sb::EngineBase>> replicas;
    sb::EngineBase* impl() const { return replicas[0].get(); }
};

// nothing may throw across the C ABI (the reference's release profile is panic = "abort"): host-side allocation
// failures and library exceptions (vector, string, regex) become an sb_status
template <typename F>
static int sb_guarded(F&& f) noexcept {
    try { return f(); }
    catch (const std::bad_alloc&) { sb::set_error("out of host memory"); return SB_ERR_NOMEM; }
    catch (const std::exception& e) { sb::set_error(std::string("internal error: ") + e.what()); return SB_ERR_INVALID; }
    catch (...) { sb::set_error("internal error"); return SB_ERR_INVALID; }
}

static const char* kNotLoaded = "Model is not loaded for transcription.";   // transcription.rs:427-429

extern "C" {

void sb_params_default(sb_params* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->language = "en";
    p->suppress_blank = 1;
    p->max_initial_ts = 1.0f;
    p->n_max_text_ctx = 16384;
    p->temperature = 0.0f;
    p->temperature_inc = 0.0f;          // pinned parity configuration: no fallback (whisper.cpp's default is 0.2)
    p->logprob_thold = -1.0f;
    p->entropy_thold = 2.4f;
}

int sb_engine_create(const sb_config* cfg, sb_engine** out) {
    return sb_guarded([&]() -> int {
        SB_CHECK_ARG(cfg && out && cfg->model_path, "cfg/out/model_path is null");
        SB_CHECK_ARG(cfg->dtype == SB_DTYPE_BF16 || cfg->dtype == SB_DTYPE_F16, "cfg.dtype");
        int ndev = 0;
        SB_CUDA_CHECK(cudaGetDeviceCount(&ndev));
        std::vector<int> devs;
        if (cfg->devices && cfg->n_devices > 0) devs.assign(cfg->devices, cfg->devices + cfg->n_devices);
        else devs.push_back(cfg->device);
        SB_CHECK_ARG(devs.size() <= 64, "at most 64 devices");
        for (size_t i = 0; i < devs.size(); ++i) {
            SB_CHECK_ARG(devs[i] >= 0 && devs[i] < ndev, "cfg.device / cfg.devices[] out of range");
            for (size_t j = 0; j < i; ++j) SB_CHECK_ARG(devs[j] != devs[i], "cfg.devices[] lists a device twice");
        }
        sb::GgmlFile file;
        int rc = sb::load_ggml_file(cfg->model_path, file);
        if (rc) return rc;
        std::unique_ptr<sb_engine> h(new sb_engine());
        for (int dev : devs) {
            SB_CUDA_CHECK(cudaSetDevice(dev));
            std::unique_ptr<sb::EngineBase> e;
            if (cfg->dtype == SB_DTYPE_F16) e.reset(new sb::Engine<__half>());
            else e.reset(new sb::Engine<__nv_bfloat16>());
            e->device = dev;
            e->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 64;
            e->dtype = cfg->dtype;
            e->use_graph = cfg->use_cuda_graph;
            if (cfg->dtype == SB_DTYPE_F16) rc = static_cast<sb::Engine<__half>*>(e.get())->load(file);
            else rc = static_cast<sb::Engine<__nv_bfloat16>*>(e.get())->load(file);
            if (rc) return rc;
            SB_CUDA_CHECK(cudaDeviceSynchronize());
            h->replicas.push_back(std::move(e));
        }
        *out = h.release();
        return SB_OK;
    });
}

int sb_engine_destroy(sb_engine* e) {
    if (!e) return SB_OK;
    return sb_guarded([&]() -> int {
        for (auto& r : e->replicas) {
            cudaSetDevice(r->device);
            cudaDeviceSynchronize();
            r.reset();
        }
        delete e;
        return SB_OK;
    });
}

int sb_engine_info(const sb_engine* e, sb_model_info* info) {
    SB_CHECK_ARG(e && info, "null pointer");
    const sb::WhisperHParams& h = e->impl()->hp;
    *info = sb_model_info{h.n_vocab, h.n_audio_ctx, h.n_audio_state, h.n_audio_head, h.n_audio_layer, h.n_text_ctx,
                          h.n_text_state, h.n_text_head, h.n_text_layer, h.n_mels, h.ftype,
                          e->impl()->sp.eot, e->impl()->sp.sot, e->impl()->sp.beg, e->impl()->sp.blank};
    return SB_OK;
}

int sb_engine_device_count(const sb_engine* e) { return e ? (int)e->replicas.size() : 0; }

void* sb_engine_stream(sb_engine* e) { return e ? e->impl()->stream_handle() : nullptr; }

int sb_engine_set_profile(sb_engine* e, int enable) {
    SB_CHECK_ARG(e, "null engine");
    for (auto& r : e->replicas) r->profile = enable;
    return SB_OK;
}

int sb_engine_stats(sb_engine* e, sb_stats* out, int reset) {
    SB_CHECK_ARG(e && out, "null pointer");
    *out = e->impl()->stats;
    if (reset) for (auto& r : e->replicas) memset(&r->stats, 0, sizeof(sb_stats));
    return SB_OK;
}

int sb_token_text(const sb_engine* e, int32_t id, char* buf, int cap) {
    if (!e) return 0;
    int n = 0;
    sb_guarded([&]() -> int {
        const std::string s = e->impl()->token_text(id);
        if (buf && cap > 0) memcpy(buf, s.data(), std::min<size_t>(s.size(), (size_t)cap));
        n = (int)s.size();
        return SB_OK;
    });
    return n;
}

int sb_tokenize(const sb_engine* e, const char* text, int32_t* tokens, int cap) {
    if (!e) { sb::set_error(kNotLoaded); return SB_ERR_NOT_LOADED; }
    return sb_guarded([&]() -> int {
        SB_CHECK_ARG(text && (tokens || cap == 0), "null pointer");
        const std::vector<int> t = e->impl()->tokenize(text);
        for (int i = 0; i < (int)t.size() && i < cap; ++i) tokens[i] = t[i];
        return (int)t.size();
    });
}

int sb_transcribe_batch(sb_engine* e, const float* const* pcm16k, const size_t* n_samples, size_t count,
                        const sb_params* p, sb_result* out) {
    if (!e) { sb::set_error(kNotLoaded); return SB_ERR_NOT_LOADED; }
    return sb_guarded([&]() -> int {
        SB_CHECK_ARG(out && (count == 0 || (pcm16k && n_samples)), "null pointer");
        memset(out, 0, count * sizeof(sb_result));
        sb_params dp;
        if (!p) { sb_params_default(&dp); p = &dp; }
        for (size_t i = 0; i < count; ++i) SB_CHECK_ARG(n_samples[i] == 0 || pcm16k[i], "null clip pointer");
        const size_t R = e->replicas.size();
        if (R == 1 || count <= 1) return e->impl()->transcribe_batch(pcm16k, n_samples, count, *p, out);
        // clips are independent: clip i goes to device i mod R, one worker thread per device, no collective
        std::vector<std::vector<const float*>> ptrs(R);
        std::vector<std::vector<size_t>> ns(R), idx(R);
        for (size_t i = 0; i < count; ++i) { ptrs[i % R].push_back(pcm16k[i]); ns[i % R].push_back(n_samples[i]); idx[i % R].push_back(i); }
        std::vector<std::vector<sb_result>> res(R);
        std::vector<int> rcs(R, SB_OK);
        std::vector<std::string> errs(R);
        std::vector<std::thread> workers;
        for (size_t r = 0; r < R; ++r) {
            if (idx[r].empty()) continue;
            res[r].resize(idx[r].size());
            workers.emplace_back([&, r] {
                rcs[r] = sb_guarded([&]() -> int {
                    return e->replicas[r]->transcribe_batch(ptrs[r].data(), ns[r].data(), idx[r].size(), *p, res[r].data());
                });
                if (rcs[r]) errs[r] = sb::get_error();       // the message is thread-local: hand it to the caller
            });
        }
        for (auto& w : workers) w.join();
        int rc = SB_OK;
        for (size_t r = 0; r < R; ++r) {
            for (size_t k = 0; k < idx[r].size(); ++k) out[idx[r][k]] = res[r][k];
            if (rcs[r] && !rc) { rc = rcs[r]; sb::set_error("device " + std::to_string(e->replicas[r]->device) + ": " + errs[r]); }
        }
        if (rc) for (size_t i = 0; i < count; ++i) sb_result_free(out + i);
        return rc;
    });
}

int sb_transcribe(sb_engine* e, const float* pcm16k, size_t n_samples, const sb_params* p, sb_result* out) {
    const float* ptrs[1] = {pcm16k};
    size_t ns[1] = {n_samples};
    return sb_transcribe_batch(e, ptrs, ns, 1, p, out);
}

void sb_result_free(sb_result* r) {
    if (!r) return;
    free(r->text); free(r->tokens); free(r->sampled); free(r->margins); free(r->tids); free(r->logprobs); free(r->windows);
    free(r->segments); free(r->segment_text);
    memset(r, 0, sizeof(*r));
}

int sb_encode(sb_engine* e, const float* mel_windows, int n_windows, float* enc_out) {
    if (!e) { sb::set_error(kNotLoaded); return SB_ERR_NOT_LOADED; }
    return sb_guarded([&]() -> int { return e->impl()->encode_host(mel_windows, n_windows, enc_out); });
}

int sb_decode_trace(sb_engine* e, const float* mel_windows, int n_windows, const int32_t* seek_end, const sb_params* p,
                    const int32_t* forced, int n_steps, float* logits_out, int32_t* tokens_out, float* margins_out) {
    if (!e) { sb::set_error(kNotLoaded); return SB_ERR_NOT_LOADED; }
    return sb_guarded([&]() -> int {
        sb_params dp;
        if (!p) { sb_params_default(&dp); p = &dp; }
        return e->impl()->decode_trace(mel_windows, n_windows, seek_end, *p, forced, n_steps, logits_out, tokens_out, margins_out);
    });
}

/* ABI layout table: sizeof / offsetof of every struct that crosses the boundary, so that a binding in another language
 * (ctypes here, the Rust crate in rust/spittle-b200-sys) can be checked against the library it loads. */
#define SB_LAYOUT_ROWS(X)                                                                                               \
    X(sb_config, model_path) X(sb_config, device) X(sb_config, max_batch) X(sb_config, dtype) X(sb_config, use_cuda_graph) \
    X(sb_config, devices) X(sb_config, n_devices)                                                                       \
    X(sb_params, language) X(sb_params, translate) X(sb_params, initial_prompt) X(sb_params, no_timestamps)             \
    X(sb_params, suppress_blank) X(sb_params, single_segment) X(sb_params, max_initial_ts) X(sb_params, n_max_tokens)   \
    X(sb_params, max_windows) X(sb_params, n_max_text_ctx) X(sb_params, temperature) X(sb_params, temperature_inc)            \
    X(sb_params, logprob_thold) X(sb_params, entropy_thold) X(sb_params, suppress_nst)                                   \
    X(sb_window_info, seek) X(sb_window_info, n_tokens) X(sb_window_info, result_len) X(sb_window_info, seek_delta)     \
    X(sb_window_info, failed) X(sb_window_info, token_offset) X(sb_window_info, n_prompt) X(sb_window_info, temperature)  \
    X(sb_window_info, n_attempts) X(sb_window_info, avg_logprob)                               \
    X(sb_segment, t0) X(sb_segment, t1) X(sb_segment, text) X(sb_segment, text_len) X(sb_segment, token_offset)         \
    X(sb_segment, n_tokens)                                                                                             \
    X(sb_result, text) X(sb_result, text_len) X(sb_result, tokens) X(sb_result, n_tokens) X(sb_result, sampled)         \
    X(sb_result, n_sampled) X(sb_result, margins) X(sb_result, tids) X(sb_result, logprobs) X(sb_result, windows) X(sb_result, n_windows)      \
    X(sb_result, segments) X(sb_result, n_segments) X(sb_result, segment_text) X(sb_result, ms_mel)                     \
    X(sb_result, ms_encode) X(sb_result, ms_decode) X(sb_result, status) X(sb_result, lang_id)                          \
    X(sb_model_info, n_vocab) X(sb_model_info, token_blank) X(sb_stats, clips) X(sb_stats, dstep_count) X(sb_stats, prefill_rows) X(sb_stats, fallbacks)

int sb_abi_layout(sb_abi_field* out, int cap) {
    static const sb_abi_field rows[] = {
#define X(S, F) {#S, #F, (int)sizeof(S), (int)offsetof(S, F)},
        SB_LAYOUT_ROWS(X)
#undef X
    };
    const int n = (int)(sizeof(rows) / sizeof(rows[0]));
    for (int i = 0; i < n && i < cap && out; ++i) out[i] = rows[i];
    return n;
}

}  // extern "C"
