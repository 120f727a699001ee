// whisper_cpu_ref -- multi-threaded C++ restatement of the reference's CPU path for the Whisper hot path.
//
// TEST INFRASTRUCTURE ONLY (oracle/__init__.py): the checker and the timed CPU baseline of bench.py; never linked or
// called by the product (spittle_b200/, host/, include/).  PARITY UNPINNED: the reference holds no fixture for this path
// and whisper.cpp is not vendored -- this file restates, from SURVEY.md App. C / D and oracle/ASSUMPTIONS.md, what the
// reference executes at src-tauri/src/managers/transcription.rs:501-503 (`whisper_engine.transcribe_samples`) through
// transcribe-rs 0.2.3 -> whisper-rs 0.13.2 -> whisper.cpp `whisper_full_with_state` on the CPU:
//
//   logmel            App. C.1  log_mel_spectrogram: f32 recursive radix-2 FFT (25-point DFT leaf), double mel sums
//   encode            App. C.2  conv stem + pre-LN blocks + ln_post; f16 weights x f16-rounded activations, f32 accumulate
//   cross_kv          App. C.2  per decoder layer K = Wk enc, V = Wv enc + bv, stored f16
//   decode_step       App. C.3  KV-cached self-attention, cross-attention, MLP, tied-embedding logits
//   process_logits    App. C.4  whisper_process_logits; sample_best = whisper_sample_token(best): lowest index wins ties
//   full              App. C.4  the seek loop with [prev] + prompt_past text conditioning, pinned greedy configuration
//
// It follows the numpy oracle (oracle/whisper_ref.py, oracle/logmel.py) operation by operation and is validated against it
// (tests/test_cpu_ref_cpu.py); like whisper.cpp it runs on n_threads host threads (whisper.cpp's default is
// min(4, hardware_concurrency); bench.py times both 4 threads and all cores).  x86-64 with AVX2 + FMA + F16C required,
// AVX-512 used when the CPU has it (runtime dispatch).
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <unordered_map>

// ------------------------------------------------------------------------------------------------------------
// SIMD kernels, two instruction sets
// ------------------------------------------------------------------------------------------------------------
#define KT __attribute__((target("avx512f,avx512vl,avx512bw,fma,f16c")))
#define KN(x) x##_avx512
#define VEC __m512
#define VW 16
#define MR 12
#define VLOAD(p) _mm512_loadu_ps(p)
#define VSTORE(p, v) _mm512_storeu_ps(p, v)
#define VLOADH(p) _mm512_cvtph_ps(_mm256_loadu_si256((const __m256i*)(p)))
#define VFMA(a, b, c) _mm512_fmadd_ps(a, b, c)
#define VZERO() _mm512_setzero_ps()
#define VSET1(x) _mm512_set1_ps(x)
#define VHSUM(v) _mm512_reduce_add_ps(v)
#define VROUNDH(v) _mm512_cvtph_ps(_mm512_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC))
#define VMAX(a, b) _mm512_max_ps(a, b)
#define VHMAX(v) _mm512_reduce_max_ps(v)
#define VMUL(a, b) _mm512_mul_ps(a, b)
#define VSUB(a, b) _mm512_sub_ps(a, b)
#define VADD(a, b) _mm512_add_ps(a, b)
#define VFNMA(a, b, c) _mm512_fnmadd_ps(a, b, c)
#define VRINT(v) _mm512_roundscale_ps(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC)
#define VSCALE2K(p, k) _mm512_scalef_ps(p, k)
#include "kernels.inc"
#undef KT
#undef KN
#undef VEC
#undef VW
#undef MR
#undef VLOAD
#undef VSTORE
#undef VLOADH
#undef VFMA
#undef VZERO
#undef VSET1
#undef VHSUM
#undef VROUNDH
#undef VMAX
#undef VHMAX
#undef VMUL
#undef VSUB
#undef VADD
#undef VFNMA
#undef VRINT
#undef VSCALE2K

static inline float hsum256(__m256 v) {
    __m128 lo = _mm256_castps256_ps128(v), hi = _mm256_extractf128_ps(v, 1);
    lo = _mm_add_ps(lo, hi);
    lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_add_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}
static inline float hmax256(__m256 v) {
    __m128 lo = _mm_max_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
    lo = _mm_max_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_max_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}
static inline __m256 scale2k256(__m256 p, __m256 k) {      // p * 2^k, k integral in [-126, 127]
    const __m256i e = _mm256_slli_epi32(_mm256_add_epi32(_mm256_cvtps_epi32(k), _mm256_set1_epi32(127)), 23);
    return _mm256_mul_ps(p, _mm256_castsi256_ps(e));
}
#define KT
#define KN(x) x##_avx2
#define VEC __m256
#define VW 8
#define MR 6
#define VLOAD(p) _mm256_loadu_ps(p)
#define VSTORE(p, v) _mm256_storeu_ps(p, v)
#define VLOADH(p) _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(p)))
#define VFMA(a, b, c) _mm256_fmadd_ps(a, b, c)
#define VZERO() _mm256_setzero_ps()
#define VSET1(x) _mm256_set1_ps(x)
#define VHSUM(v) hsum256(v)
#define VROUNDH(v) _mm256_cvtph_ps(_mm256_cvtps_ph(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC))
#define VMAX(a, b) _mm256_max_ps(a, b)
#define VHMAX(v) hmax256(v)
#define VMUL(a, b) _mm256_mul_ps(a, b)
#define VSUB(a, b) _mm256_sub_ps(a, b)
#define VADD(a, b) _mm256_add_ps(a, b)
#define VFNMA(a, b, c) _mm256_fnmadd_ps(a, b, c)
#define VRINT(v) _mm256_round_ps(v, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC)
#define VSCALE2K(p, k) scale2k256(p, k)
#include "kernels.inc"
#undef KT
#undef KN

namespace {

bool g_avx512 = false;
thread_local std::string t_err;

inline void gemm_range(const float* A, int lda, const uint16_t* W, int ldw, float* C, int ldc, int M, int K, int n0, int n1,
                       float* scratch, int panel_rows) {
    if (g_avx512) gemm_range_avx512(A, lda, W, ldw, C, ldc, M, K, n0, n1, scratch, panel_rows);
    else gemm_range_avx2(A, lda, W, ldw, C, ldc, M, K, n0, n1, scratch, panel_rows);
}
inline void gemv_h(const uint16_t* W, int ldw, const float* x, float* y, int K, int n0, int n1) {
    if (g_avx512) gemv_h_avx512(W, ldw, x, y, K, n0, n1); else gemv_h_avx2(W, ldw, x, y, K, n0, n1);
}
inline void axpy_h(float p, const uint16_t* v, float* o, int d) {
    if (g_avx512) axpy_h_avx512(p, v, o, d); else axpy_h_avx2(p, v, o, d);
}
inline void round_h(float* x, size_t n) { if (g_avx512) round_h_avx512(x, n); else round_h_avx2(x, n); }
inline void softmax_round_h(float* s, int n, int n_pad, float scale) { if (g_avx512) softmax_round_h_avx512(s, n, n_pad, scale); else softmax_round_h_avx2(s, n, n_pad, scale); }
inline float h2f(uint16_t h) { return _cvtsh_ss(h); }
inline uint16_t f2h(float f) { return _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }
inline float rh(float f) { return h2f(f2h(f)); }

// ------------------------------------------------------------------------------------------------------------
// thread pool: workers spin briefly on a generation counter (a decoder step is ~100 small parallel regions), then sleep
// ------------------------------------------------------------------------------------------------------------
class Pool {
  public:
    explicit Pool(int n) : n_(std::max(1, n)) {
        for (int i = 1; i < n_; ++i) th_.emplace_back([this, i] { worker(i); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; gen_.fetch_add(1); }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return n_; }
    // fn(thread index, n threads) on every thread of the pool, the caller being thread 0
    void run(const std::function<void(int, int)>& fn) {
        if (n_ == 1) { fn(0, 1); return; }
        fn_ = &fn;
        pending_.store(n_ - 1, std::memory_order_release);
        { std::lock_guard<std::mutex> g(mu_); gen_.fetch_add(1, std::memory_order_release); }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        fn(0, n_);
        while (pending_.load(std::memory_order_acquire) != 0) _mm_pause();
    }
    // contiguous split of [0, n) in units of `unit`
    static void split(int n, int unit, int t, int nt, int* b, int* e) {
        const int units = (n + unit - 1) / unit;
        const int u0 = (int)((int64_t)units * t / nt), u1 = (int)((int64_t)units * (t + 1) / nt);
        *b = std::min(n, u0 * unit); *e = std::min(n, u1 * unit);
    }

  private:
    void worker(int idx) {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < 20000) { _mm_pause(); continue; }
                std::unique_lock<std::mutex> lk(mu_);
                sleepers_.fetch_add(1);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1);
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            (*fn_)(idx, n_);
            pending_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> pending_{0}, sleepers_{0};
    const std::function<void(int, int)>* fn_ = nullptr;
    bool stop_ = false;
};

// ------------------------------------------------------------------------------------------------------------
// GGML legacy model file: f32 / f16 tensors
// ------------------------------------------------------------------------------------------------------------
struct HParams { int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer, n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype; };
struct Special { int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang_first, num_languages, blank; };

struct Tensor { std::vector<int64_t> shape; int ttype = 0; std::vector<uint8_t> data; int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; } };

struct Lin { std::vector<uint16_t> w; std::vector<float> b; int n = 0, k = 0; };      // w [n][k] f16 (k padded to 16), b [n] (zeros if absent)
struct Ln { std::vector<float> g, b; };
struct EncLayer { Ln ln1, ln2; Lin q, k, v, o, fc1, fc2; };
struct DecLayer { Ln ln1, ln2, ln3; Lin q, k, v, o, cq, ck, cv, co, fc1, fc2; };

}  // namespace

struct wcr_model {
    HParams hp{};
    Special sp{};
    std::vector<float> mel_filters;       // [n_mel][201]
    std::vector<std::string> vocab;
    std::vector<int> nst_ids;             // whisper.cpp's non_speech_tokens present in this vocabulary (suppress_nst)
    Lin conv1, conv2;                     // [d][3*n_mel], [d][3*d] (tap-major columns)
    std::vector<float> enc_pos, dec_pos;
    std::vector<EncLayer> enc;
    Ln ln_post, ln_f;
    std::vector<uint16_t> tok_emb;        // [n_vocab][d] f16
    std::vector<DecLayer> dec;
    std::vector<uint16_t> gelu_tab;       // ggml's f16 GELU table: f16 bits -> f16 bits
    std::unique_ptr<Pool> pool;
    int pool_threads = 0;
    Pool& threads(int n) {
        n = std::max(1, n);
        if (!pool || pool_threads != n) { pool.reset(new Pool(n)); pool_threads = n; }
        return *pool;
    }
};

namespace {

struct Reader {
    const uint8_t* p; size_t n, off = 0; bool ok = true;
    template <typename V> V get() { V v{}; if (sizeof(V) > n - off) { ok = false; return v; } memcpy(&v, p + off, sizeof(V)); off += sizeof(V); return v; }
    const uint8_t* take(size_t k) { if (k > n - off) { ok = false; return nullptr; } const uint8_t* r = p + off; off += k; return r; }
};

bool to_f32(const Tensor& t, std::vector<float>& out) {
    const int64_t n = t.numel();
    out.resize(n);
    if (t.ttype == 0) memcpy(out.data(), t.data.data(), n * 4);
    else if (t.ttype == 1) { const uint16_t* h = (const uint16_t*)t.data.data(); for (int64_t i = 0; i < n; ++i) out[i] = h2f(h[i]); }
    else return false;
    return true;
}

bool load_lin(const std::map<std::string, Tensor>& T, const std::string& prefix, int n, int k, Lin& out, std::string& err) {
    auto it = T.find(prefix + ".weight");
    if (it == T.end() || it->second.numel() != (int64_t)n * k) { err = "missing / misshapen tensor " + prefix + ".weight"; return false; }
    std::vector<float> w;
    if (!to_f32(it->second, w)) { err = "unsupported tensor type in " + prefix; return false; }
    out.n = n; out.k = (k + 15) / 16 * 16;
    out.w.assign((size_t)n * out.k, 0);
    for (int r = 0; r < n; ++r) for (int c = 0; c < k; ++c) out.w[(size_t)r * out.k + c] = f2h(w[(size_t)r * k + c]);   // exact for f16 files
    out.b.assign(n, 0.f);
    auto ib = T.find(prefix + ".bias");
    if (ib != T.end()) { std::vector<float> b; to_f32(ib->second, b); if ((int)b.size() == n) out.b = b; }
    return true;
}
bool load_ln(const std::map<std::string, Tensor>& T, const std::string& prefix, int d, Ln& out, std::string& err) {
    auto g = T.find(prefix + ".weight"), b = T.find(prefix + ".bias");
    if (g == T.end() || b == T.end() || g->second.numel() != d) { err = "missing tensor " + prefix; return false; }
    to_f32(g->second, out.g); to_f32(b->second, out.b);
    return true;
}
// conv weight [co][ci][3] -> rows [co][3][ci] (tap-major, the im2col column order)
bool load_conv(const std::map<std::string, Tensor>& T, const std::string& prefix, int co, int ci, Lin& out, std::string& err) {
    auto it = T.find(prefix + ".weight");
    if (it == T.end() || it->second.numel() != (int64_t)co * ci * 3) { err = "missing tensor " + prefix + ".weight"; return false; }
    std::vector<float> w; to_f32(it->second, w);
    out.n = co; out.k = (3 * ci + 15) / 16 * 16;
    out.w.assign((size_t)co * out.k, 0);
    for (int o = 0; o < co; ++o) for (int c = 0; c < ci; ++c) for (int k = 0; k < 3; ++k)
        out.w[(size_t)o * out.k + k * ci + c] = f2h(w[((size_t)o * ci + c) * 3 + k]);
    auto ib = T.find(prefix + ".bias");
    if (ib == T.end()) { err = "missing tensor " + prefix + ".bias"; return false; }
    to_f32(ib->second, out.b);
    return (int)out.b.size() == co;
}

float gelu_tanh_f32(float x) {
    const float c = 0.79788456080286535587989211986876f;
    return 0.5f * x * (1.0f + tanhf(c * x * (1.0f + 0.044715f * x * x)));
}

// ---- building blocks (rows x d row-major f32) ----
void layer_norm(const float* x, const Ln& w, float* out, int rows, int d, Pool& P) {
    P.run([&](int t, int nt) {
        int b, e; Pool::split(rows, 1, t, nt, &b, &e);
        for (int r = b; r < e; ++r) {
            const float* xr = x + (size_t)r * d;
            double s = 0.0;
            for (int i = 0; i < d; ++i) s += xr[i];
            const float mu = (float)(s / d);
            double q = 0.0;
            for (int i = 0; i < d; ++i) { const float c = xr[i] - mu; q += (double)c * (double)c; }
            const float var = (float)(q / d);
            const float inv = 1.0f / sqrtf(var + 1e-5f);
            float* o = out + (size_t)r * d;
            for (int i = 0; i < d; ++i) o[i] = (xr[i] - mu) * inv * w.g[i] + w.b[i];
        }
    });
}

// y[M, n] = r(x)[M, k] W^T + b ; x must already be f16-rounded when round_in is false.  act: 0 none, 1 GELU (f16 table)
struct Scratch { std::vector<float> panel; };
void linear(const wcr_model& m, const float* x, int ldx, int M, const Lin& L, float* y, int ldy, int act, Pool& P,
            std::vector<Scratch>& scr) {
    const int K = L.k;
    if (M == 1) {
        P.run([&](int t, int nt) {
            int b, e; Pool::split(L.n, 4, t, nt, &b, &e);
            gemv_h(L.w.data(), K, x, y, K, b, e);
            for (int i = b; i < e; ++i) y[i] += L.b[i];
        });
    } else {
        const int panel_rows = std::max(4, (196608 / K) & ~3);      // (scratch sizing only: >= 512 x 32 floats)
        P.run([&](int t, int nt) {
            int b, e; Pool::split(L.n, 4, t, nt, &b, &e);
            if (b >= e) return;
            std::vector<float>& pan = scr[t].panel;
            if (pan.size() < std::max<size_t>((size_t)panel_rows * K, 512 * 32)) pan.resize(std::max<size_t>((size_t)panel_rows * K, 512 * 32));
            gemm_range(x, ldx, L.w.data(), K, y, ldy, M, K, b, e, pan.data(), panel_rows);
            for (int r = 0; r < M; ++r) { float* yr = y + (size_t)r * ldy; for (int i = b; i < e; ++i) yr[i] += L.b[i]; }
        });
    }
    if (act == 1) {
        P.run([&](int t, int nt) {
            int b, e; Pool::split(M, 1, t, nt, &b, &e);
            for (int r = b; r < e; ++r) {
                float* yr = y + (size_t)r * ldy;
                for (int i = 0; i < L.n; ++i) {
                    const float v = yr[i];
                    // ggml_vec_gelu_f32: f16 table lookup inside (-10, 10)
                    yr[i] = v <= -10.0f ? 0.0f : (v >= 10.0f ? v : h2f(m.gelu_tab[f2h(v)]));
                }
            }
        });
    }
}

void round_rows(float* x, int rows, int d, int ld, Pool& P) {
    P.run([&](int t, int nt) {
        int b, e; Pool::split(rows, 1, t, nt, &b, &e);
        for (int r = b; r < e; ++r) round_h(x + (size_t)r * ld, d);
    });
}

// ---- encoder ----
struct EncWork {
    std::vector<float> cols, y1, x, h, q, k, v, att, mlp;
    std::vector<Scratch> scr;
};

// softmax(q k^T / 8) v for one head over T keys; q/k/v f16-rounded f32 [T][d] views with head offset; out [T][d]
void enc_attention(const wcr_model& m, const float* q, const float* k, const float* v, float* out, int T, int d, int n_head, Pool& P,
                   std::vector<Scratch>& scr) {
    const int dh = 64, Tp = (T + 15) / 16 * 16;
    // per head: K as f16 [T][64], V^T as f16 [64][Tp]; scores block of rows at a time
    P.run([&](int t, int nt) {
        std::vector<uint16_t> kh((size_t)T * dh), vt((size_t)dh * Tp, 0);
        std::vector<float> qh((size_t)T * dh), s((size_t)64 * Tp), o((size_t)64 * dh);
        std::vector<float>& pan = scr[t].panel;
        for (int h = t; h < n_head; h += nt) {
            for (int i = 0; i < T; ++i)
                for (int c = 0; c < dh; ++c) {
                    kh[(size_t)i * dh + c] = f2h(k[(size_t)i * d + h * dh + c]);
                    vt[(size_t)c * Tp + i] = f2h(v[(size_t)i * d + h * dh + c]);
                    qh[(size_t)i * dh + c] = q[(size_t)i * d + h * dh + c];
                }
            const int pr_k = std::max(4, (196608 / dh) & ~3), pr_v = std::max(4, (196608 / Tp) & ~3);
            if (pan.size() < 512 * 32) pan.resize(512 * 32);
            for (int i0 = 0; i0 < T; i0 += 64) {
                const int mi = std::min(64, T - i0);
                gemm_range(qh.data() + (size_t)i0 * dh, dh, kh.data(), dh, s.data(), Tp, mi, dh, 0, T, pan.data(), pr_k);
                // softmax(s / 8) rows, probabilities rounded to f16 before P V (ggml mul_mat)
                for (int i = 0; i < mi; ++i) softmax_round_h(s.data() + (size_t)i * Tp, T, Tp, 0.125f);
                gemm_range(s.data(), Tp, vt.data(), Tp, o.data(), dh, mi, Tp, 0, dh, pan.data(), pr_v);
                for (int i = 0; i < mi; ++i) memcpy(out + (size_t)(i0 + i) * d + h * dh, o.data() + (size_t)i * dh, dh * 4);
            }
        }
    });
}

struct PhaseTimer {
    bool on = getenv("WCR_PROFILE") != nullptr;
    std::map<std::string, double> acc;
    std::chrono::steady_clock::time_point t0;
    void start() { if (on) t0 = std::chrono::steady_clock::now(); }
    void stop(const char* name) { if (on) acc[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    ~PhaseTimer() { if (on) for (auto& kv : acc) fprintf(stderr, "  [cpu_ref] %-10s %.3f s\n", kv.first.c_str(), kv.second); }
};

void encode(wcr_model& m, const float* mel_win, int n_threads, float* enc_out) {
    PhaseTimer pt;
    const HParams& hp = m.hp;
    const int d = hp.n_audio_state, nm = hp.n_mels, T = hp.n_audio_ctx, T2 = 2 * T;
    Pool& P = m.threads(n_threads);
    EncWork w;
    w.scr.resize(P.size());
    // conv1: x = r(mel^T) [3000][n_mel]; im2col, tap-major columns
    const int K1 = m.conv1.k, K2 = m.conv2.k;
    w.cols.assign((size_t)T2 * K1, 0.f);
    for (int i = 0; i < T2; ++i)
        for (int tap = 0; tap < 3; ++tap) {
            const int src = i + tap - 1;
            if (src < 0 || src >= T2) continue;
            for (int c = 0; c < nm; ++c) w.cols[(size_t)i * K1 + tap * nm + c] = rh(mel_win[(size_t)c * T2 + src]);
        }
    w.y1.resize((size_t)T2 * d);
    linear(m, w.cols.data(), K1, T2, m.conv1, w.y1.data(), d, 1, P, w.scr);
    round_rows(w.y1.data(), T2, d, d, P);
    // conv2: stride 2
    w.cols.assign((size_t)T * K2, 0.f);
    for (int i = 0; i < T; ++i)
        for (int tap = 0; tap < 3; ++tap) {
            const int src = 2 * i + tap - 1;
            if (src < 0 || src >= T2) continue;
            memcpy(&w.cols[(size_t)i * K2 + (size_t)tap * d], &w.y1[(size_t)src * d], (size_t)d * 4);
        }
    w.x.resize((size_t)T * d);
    linear(m, w.cols.data(), K2, T, m.conv2, w.x.data(), d, 1, P, w.scr);
    for (size_t i = 0; i < (size_t)T * d; ++i) w.x[i] += m.enc_pos[i];
    w.h.resize((size_t)T * d); w.q.resize((size_t)T * d); w.k.resize((size_t)T * d); w.v.resize((size_t)T * d);
    w.att.resize((size_t)T * d); w.mlp.resize((size_t)T * 4 * d);
    std::vector<float> tmp((size_t)T * d);
    for (int l = 0; l < hp.n_audio_layer; ++l) {
        const EncLayer& L = m.enc[l];
        pt.start();
        layer_norm(w.x.data(), L.ln1, w.h.data(), T, d, P);
        round_rows(w.h.data(), T, d, d, P);
        pt.stop("ln+round"); pt.start();
        linear(m, w.h.data(), d, T, L.q, w.q.data(), d, 0, P, w.scr);
        linear(m, w.h.data(), d, T, L.k, w.k.data(), d, 0, P, w.scr);
        linear(m, w.h.data(), d, T, L.v, w.v.data(), d, 0, P, w.scr);
        pt.stop("qkv"); pt.start();
        round_rows(w.q.data(), T, d, d, P);        // q rounded as the activation operand of Q K^T; K, V stored f16
        enc_attention(m, w.q.data(), w.k.data(), w.v.data(), w.att.data(), T, d, hp.n_audio_head, P, w.scr);
        round_rows(w.att.data(), T, d, d, P);
        pt.stop("attention"); pt.start();
        linear(m, w.att.data(), d, T, L.o, tmp.data(), d, 0, P, w.scr);
        for (size_t i = 0; i < (size_t)T * d; ++i) w.x[i] += tmp[i];
        pt.stop("o"); pt.start();
        layer_norm(w.x.data(), L.ln2, w.h.data(), T, d, P);
        round_rows(w.h.data(), T, d, d, P);
        pt.stop("ln+round"); pt.start();
        linear(m, w.h.data(), d, T, L.fc1, w.mlp.data(), 4 * d, 1, P, w.scr);
        round_rows(w.mlp.data(), T, 4 * d, 4 * d, P);
        pt.stop("fc1+gelu"); pt.start();
        linear(m, w.mlp.data(), 4 * d, T, L.fc2, tmp.data(), d, 0, P, w.scr);
        for (size_t i = 0; i < (size_t)T * d; ++i) w.x[i] += tmp[i];
        pt.stop("fc2");
    }
    layer_norm(w.x.data(), m.ln_post, enc_out, T, d, P);
}

// ---- decoder ----
struct DecState {
    // cross cache per layer: K [H][T][64] f16, V [H][T][64] f16 ; self cache per layer: K, V [n_ctx][d] f16
    std::vector<std::vector<uint16_t>> ck, cv, sk, sv;
    std::vector<Scratch> scr;
    std::vector<float> x, h, q, k, v, att, mlp, tmp, sc, logits;
};

void cross_kv(wcr_model& m, const float* enc, int n_threads, DecState& s) {
    const HParams& hp = m.hp;
    const int d = hp.n_text_state, T = hp.n_audio_ctx, H = hp.n_text_head;
    Pool& P = m.threads(n_threads);
    s.scr.resize(P.size());
    std::vector<float> e((size_t)T * d), k((size_t)T * d), v((size_t)T * d);
    memcpy(e.data(), enc, e.size() * 4);
    round_rows(e.data(), T, d, d, P);
    s.ck.assign(hp.n_text_layer, {}); s.cv.assign(hp.n_text_layer, {});
    s.sk.assign(hp.n_text_layer, std::vector<uint16_t>((size_t)hp.n_text_ctx * d));
    s.sv.assign(hp.n_text_layer, std::vector<uint16_t>((size_t)hp.n_text_ctx * d));
    for (int l = 0; l < hp.n_text_layer; ++l) {
        linear(m, e.data(), d, T, m.dec[l].ck, k.data(), d, 0, P, s.scr);
        linear(m, e.data(), d, T, m.dec[l].cv, v.data(), d, 0, P, s.scr);
        s.ck[l].resize((size_t)T * d); s.cv[l].resize((size_t)T * d);
        for (int h = 0; h < H; ++h)
            for (int t = 0; t < T; ++t)
                for (int c = 0; c < 64; ++c) {
                    s.ck[l][((size_t)h * T + t) * 64 + c] = f2h(k[(size_t)t * d + h * 64 + c]);
                    s.cv[l][((size_t)h * T + t) * 64 + c] = f2h(v[(size_t)t * d + h * 64 + c]);
                }
    }
    s.x.resize(d); s.h.resize(d); s.q.resize(d); s.k.resize(d); s.v.resize(d); s.att.resize(d); s.mlp.resize(4 * d); s.tmp.resize(d);
    s.sc.resize((size_t)P.size() * 1536);
    s.logits.resize(hp.n_vocab);
}

// attention of one query over n keys: K, V f16 rows (stride ld); per head by thread
void dec_attention(const float* q, const uint16_t* K, const uint16_t* V, int ld, int64_t head_stride, int n, int H, float* out,
                   float* sc_all, Pool& P) {
    P.run([&](int t, int nt) {
        float* sc = sc_all + (size_t)t * 1536;
        for (int h = t; h < H; h += nt) {
            const uint16_t* kh = K + h * head_stride;
            const uint16_t* vh = V + h * head_stride;
            gemv_h(kh, ld, q + h * 64, sc, 64, 0, n);
            float mx = -INFINITY;
            for (int j = 0; j < n; ++j) { sc[j] *= 0.125f; mx = std::max(mx, sc[j]); }
            float sum = 0.f;
            for (int j = 0; j < n; ++j) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
            float* o = out + h * 64;
            for (int c = 0; c < 64; ++c) o[c] = 0.f;
            for (int j = 0; j < n; ++j) axpy_h(rh(sc[j] / sum), vh + (size_t)j * ld, o, 64);
        }
    });
}

// one token at position n_past -> logits [n_vocab]
void decode_step(wcr_model& m, DecState& s, int token, int n_past, int n_threads) {
    const HParams& hp = m.hp;
    const int d = hp.n_text_state, T = hp.n_audio_ctx, H = hp.n_text_head;
    Pool& P = m.threads(n_threads);
    for (int i = 0; i < d; ++i) s.x[i] = h2f(m.tok_emb[(size_t)token * d + i]) + m.dec_pos[(size_t)n_past * d + i];
    auto lin1 = [&](const float* in, const Lin& L, float* out, int act) { linear(m, in, L.k, 1, L, out, L.n, act, P, s.scr); };
    for (int l = 0; l < hp.n_text_layer; ++l) {
        const DecLayer& L = m.dec[l];
        layer_norm(s.x.data(), L.ln1, s.h.data(), 1, d, P); round_h(s.h.data(), d);
        lin1(s.h.data(), L.q, s.q.data(), 0); lin1(s.h.data(), L.k, s.k.data(), 0); lin1(s.h.data(), L.v, s.v.data(), 0);
        for (int i = 0; i < d; ++i) { s.sk[l][(size_t)n_past * d + i] = f2h(s.k[i]); s.sv[l][(size_t)n_past * d + i] = f2h(s.v[i]); }
        round_h(s.q.data(), d);
        dec_attention(s.q.data(), s.sk[l].data(), s.sv[l].data(), d, 64, n_past + 1, H, s.att.data(), s.sc.data(), P);
        round_h(s.att.data(), d);
        lin1(s.att.data(), L.o, s.tmp.data(), 0);
        for (int i = 0; i < d; ++i) s.x[i] += s.tmp[i];
        layer_norm(s.x.data(), L.ln2, s.h.data(), 1, d, P); round_h(s.h.data(), d);
        lin1(s.h.data(), L.cq, s.q.data(), 0);
        round_h(s.q.data(), d);
        dec_attention(s.q.data(), s.ck[l].data(), s.cv[l].data(), 64, (int64_t)T * 64, T, H, s.att.data(), s.sc.data(), P);
        round_h(s.att.data(), d);
        lin1(s.att.data(), L.co, s.tmp.data(), 0);
        for (int i = 0; i < d; ++i) s.x[i] += s.tmp[i];
        layer_norm(s.x.data(), L.ln3, s.h.data(), 1, d, P); round_h(s.h.data(), d);
        lin1(s.h.data(), L.fc1, s.mlp.data(), 1);
        round_h(s.mlp.data(), 4 * d);
        lin1(s.mlp.data(), L.fc2, s.tmp.data(), 0);
        for (int i = 0; i < d; ++i) s.x[i] += s.tmp[i];
    }
    layer_norm(s.x.data(), m.ln_f, s.h.data(), 1, d, P); round_h(s.h.data(), d);
    const int K = d;           // d % 16 == 0 (d_head 64)
    P.run([&](int t, int nt) {
        int b, e; Pool::split(hp.n_vocab, 4, t, nt, &b, &e);
        gemv_h(m.tok_emb.data(), K, s.h.data(), s.logits.data(), K, b, e);
    });
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C interface (ctypes: oracle/cpu_ref/__init__.py)
// ------------------------------------------------------------------------------------------------------------
extern "C" {

struct wcr_cfg {
    int language_id;          // >= 0; -1: detect on the first window
    int translate, no_timestamps, suppress_blank, single_segment;
    float max_initial_ts;
    int n_max_override;       // <= 0: n_text_ctx/2 - 4
    int n_max_text_ctx;       // whisper_full_params.n_max_text_ctx (16384); <= 0: no text context
    const int32_t* initial_prompt; int n_initial_prompt;
    int suppress_nst;         // whisper_full_params.suppress_nst: the non-speech tokens get -inf
};
struct wcr_window { int32_t seek, n_tokens, result_len, seek_delta, failed, token_offset, n_prompt; };

const char* wcr_last_error(void) { return t_err.c_str(); }
int wcr_isa(void) { return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") ? 512 : 256; }

void wcr_free(wcr_model* m) { delete m; }

int wcr_load(const char* path, wcr_model** out) {
    if (!__builtin_cpu_supports("avx2") || !__builtin_cpu_supports("fma") || !__builtin_cpu_supports("f16c")) { t_err = "cpu_ref needs AVX2 + FMA + F16C"; return -1; }
    g_avx512 = wcr_isa() == 512 && !getenv("WCR_NO_AVX512");
    FILE* f = fopen(path, "rb");
    if (!f) { t_err = std::string("cannot open ") + path; return -1; }
    fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> blob((size_t)sz);
    if (fread(blob.data(), 1, blob.size(), f) != blob.size()) { fclose(f); t_err = "short read"; return -1; }
    fclose(f);
    Reader r{blob.data(), blob.size()};
    if (r.get<uint32_t>() != 0x67676d6cu) { t_err = "not a GGML legacy file"; return -1; }
    std::unique_ptr<wcr_model> m(new wcr_model());
    int32_t* hp = (int32_t*)&m->hp;
    for (int i = 0; i < 11; ++i) hp[i] = r.get<int32_t>();
    const int n_mel = r.get<int32_t>(), n_fft = r.get<int32_t>();
    if (!r.ok || n_mel != m->hp.n_mels || n_fft != 201) { t_err = "bad mel header"; return -1; }
    const uint8_t* mf = r.take((size_t)n_mel * n_fft * 4);
    if (!mf) { t_err = "truncated"; return -1; }
    m->mel_filters.resize((size_t)n_mel * n_fft); memcpy(m->mel_filters.data(), mf, m->mel_filters.size() * 4);
    const int nv = r.get<int32_t>();
    m->vocab.resize(nv);
    for (int i = 0; i < nv; ++i) { const uint32_t len = r.get<uint32_t>(); const uint8_t* w = r.take(len); if (!r.ok) { t_err = "truncated vocab"; return -1; } m->vocab[i].assign((const char*)w, len); }
    std::map<std::string, Tensor> T;
    while (r.off < r.n) {
        const int n_dims = r.get<int32_t>(), name_len = r.get<int32_t>(), ttype = r.get<int32_t>();
        if (!r.ok || n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 256) { t_err = "corrupt tensor header"; return -1; }
        int64_t ne[4] = {1, 1, 1, 1};
        for (int i = 0; i < n_dims; ++i) ne[i] = r.get<int32_t>();
        const uint8_t* nm = r.take(name_len);
        if (!r.ok) { t_err = "corrupt tensor header"; return -1; }
        Tensor t; t.ttype = ttype;
        for (int i = n_dims - 1; i >= 0; --i) t.shape.push_back(ne[i]);
        if (ttype != 0 && ttype != 1) { t_err = "cpu_ref reads f32 / f16 tensors only"; return -1; }
        const size_t nb = (size_t)t.numel() * (ttype == 0 ? 4 : 2);
        const uint8_t* d = r.take(nb);
        if (!d) { t_err = "truncated tensor"; return -1; }
        t.data.assign(d, d + nb);
        T[std::string((const char*)nm, name_len)] = std::move(t);
    }
    const HParams& h = m->hp;
    const int d = h.n_audio_state, dt = h.n_text_state;
    std::string err;
    // special ids (App. C.5)
    Special s{50256, 50257, 50357, 50358, 50359, 50360, 50361, 50362, 50363, 0, std::max(0, h.n_vocab - 51765), 220};
    if (h.n_vocab >= 51865) {
        s.num_languages = h.n_vocab - 51765 - 1; s.eot++; s.sot++;
        const int dd = s.num_languages - 98;
        s.translate += dd; s.transcribe += dd; s.solm += dd; s.prev += dd; s.nosp += dd; s.not_ += dd; s.beg += dd;
    }
    s.lang_first = s.sot + 1;
    for (size_t i = 0; i < m->vocab.size(); ++i) if (m->vocab[i] == " ") { s.blank = (int)i; break; }
    {
        // whisper.cpp non_speech_tokens (= OpenAI tokenizer.non_speech_tokens), with and without a leading space, plus " -" and " '"
        static const char* const kList[] = {
            "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^", "_", "`", "{", "|", "}", "~",
            "\xe3\x80\x8c", "\xe3\x80\x8d", "\xe3\x80\x8e", "\xe3\x80\x8f",
            "<<", ">>", "<<<", ">>>", "--", "---", "-(", "-[", "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}",
            "\xe2\x99\xaa\xe2\x99\xaa", "\xe2\x99\xaa\xe2\x99\xaa\xe2\x99\xaa",
            "\xe2\x99\xa9", "\xe2\x99\xaa", "\xe2\x99\xab", "\xe2\x99\xac", "\xe2\x99\xad", "\xe2\x99\xae", "\xe2\x99\xaf"};
        std::unordered_map<std::string, int> t2i;
        for (int i = 0; i < (int)m->vocab.size(); ++i) t2i[m->vocab[i]] = i;
        auto add = [&](const std::string& w) { auto f = t2i.find(w); if (f != t2i.end()) m->nst_ids.push_back(f->second); };
        for (const char* t : kList) { add(t); add(std::string(" ") + t); }
        add(" -"); add(" '");
    }
    m->sp = s;
    bool ok = load_conv(T, "encoder.conv1", d, h.n_mels, m->conv1, err) && load_conv(T, "encoder.conv2", d, d, m->conv2, err);
    if (ok) { auto it = T.find("encoder.positional_embedding"); ok = it != T.end() && to_f32(it->second, m->enc_pos); if (!ok) err = "encoder.positional_embedding"; }
    m->enc.resize(h.n_audio_layer);
    for (int i = 0; ok && i < h.n_audio_layer; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i);
        EncLayer& L = m->enc[i];
        ok = load_ln(T, p + ".attn_ln", d, L.ln1, err) && load_lin(T, p + ".attn.query", d, d, L.q, err) && load_lin(T, p + ".attn.key", d, d, L.k, err) &&
             load_lin(T, p + ".attn.value", d, d, L.v, err) && load_lin(T, p + ".attn.out", d, d, L.o, err) && load_ln(T, p + ".mlp_ln", d, L.ln2, err) &&
             load_lin(T, p + ".mlp.0", 4 * d, d, L.fc1, err) && load_lin(T, p + ".mlp.2", d, 4 * d, L.fc2, err);
    }
    ok = ok && load_ln(T, "encoder.ln_post", d, m->ln_post, err) && load_ln(T, "decoder.ln", dt, m->ln_f, err);
    if (ok) {
        auto it = T.find("decoder.token_embedding.weight");
        std::vector<float> e;
        ok = it != T.end() && it->second.numel() == (int64_t)h.n_vocab * dt && to_f32(it->second, e);
        if (ok) { m->tok_emb.resize(e.size()); for (size_t i = 0; i < e.size(); ++i) m->tok_emb[i] = f2h(e[i]); } else err = "decoder.token_embedding.weight";
        auto ip = T.find("decoder.positional_embedding");
        ok = ok && ip != T.end() && to_f32(ip->second, m->dec_pos);
    }
    m->dec.resize(h.n_text_layer);
    for (int i = 0; ok && i < h.n_text_layer; ++i) {
        const std::string p = "decoder.blocks." + std::to_string(i);
        DecLayer& L = m->dec[i];
        ok = load_ln(T, p + ".attn_ln", dt, L.ln1, err) && load_lin(T, p + ".attn.query", dt, dt, L.q, err) && load_lin(T, p + ".attn.key", dt, dt, L.k, err) &&
             load_lin(T, p + ".attn.value", dt, dt, L.v, err) && load_lin(T, p + ".attn.out", dt, dt, L.o, err) &&
             load_ln(T, p + ".cross_attn_ln", dt, L.ln2, err) && load_lin(T, p + ".cross_attn.query", dt, dt, L.cq, err) &&
             load_lin(T, p + ".cross_attn.key", dt, dt, L.ck, err) && load_lin(T, p + ".cross_attn.value", dt, dt, L.cv, err) &&
             load_lin(T, p + ".cross_attn.out", dt, dt, L.co, err) && load_ln(T, p + ".mlp_ln", dt, L.ln3, err) &&
             load_lin(T, p + ".mlp.0", 4 * dt, dt, L.fc1, err) && load_lin(T, p + ".mlp.2", dt, 4 * dt, L.fc2, err);
    }
    if (!ok) { t_err = "model file: " + err; return -1; }
    if (d % 64 || dt % 64 || h.n_audio_ctx > 1536) { t_err = "unsupported shape"; return -1; }
    m->gelu_tab.resize(65536);
    for (int i = 0; i < 65536; ++i) m->gelu_tab[i] = f2h(gelu_tanh_f32(h2f((uint16_t)i)));
    *out = m.release();
    return 0;
}

void wcr_hparams(const wcr_model* m, int32_t* out11) { memcpy(out11, &m->hp, 44); }

// log_mel_spectrogram, f32-faithful (oracle/logmel.py logmel_f32_faithful).  out [n_mel][n_len]
int wcr_logmel_geometry(size_t n, int* n_len, int* n_len_org) {
    *n_len = (int)((n + 480000 + 400 - 400) / 160);
    *n_len_org = 1 + (int)((n + 200 - 400) / 160);
    return 0;
}

static void fft_rec(const float* in, int N, float* out_re, float* out_im, const float* SIN, const float* COS, std::vector<float>& work, size_t woff) {
    if (N == 1) { out_re[0] = in[0]; out_im[0] = 0.f; return; }
    const int half = N / 2;
    if (N - half * 2 == 1) {           // dft(): sequential in n, all f32
        const int step = 400 / N;
        for (int k = 0; k < N; ++k) {
            float re = 0.f, im = 0.f;
            for (int n = 0; n < N; ++n) {
                const int idx = (k * n * step) % 400;
                re += in[n] * COS[idx];
                im -= in[n] * SIN[idx];
            }
            out_re[k] = re; out_im[k] = im;
        }
        return;
    }
    // even / odd halves
    if (work.size() < woff + 6 * (size_t)half) work.resize(woff + 6 * (size_t)half);
    const size_t e_in = woff, o_in = woff + half, er = woff + 2 * half, ei = woff + 3 * half, orr = woff + 4 * half, oi = woff + 5 * half;
    for (int i = 0; i < half; ++i) { work[e_in + i] = in[2 * i]; work[o_in + i] = in[2 * i + 1]; }
    const size_t next = woff + 6 * (size_t)half;
    {
        std::vector<float> tmp(half);
        memcpy(tmp.data(), &work[e_in], half * 4);
        std::vector<float> re(half), im(half);
        fft_rec(tmp.data(), half, re.data(), im.data(), SIN, COS, work, next);
        if (work.size() < woff + 6 * (size_t)half) work.resize(woff + 6 * (size_t)half);
        memcpy(&work[er], re.data(), half * 4); memcpy(&work[ei], im.data(), half * 4);
        memcpy(tmp.data(), &work[o_in], half * 4);
        fft_rec(tmp.data(), half, re.data(), im.data(), SIN, COS, work, next);
        memcpy(&work[orr], re.data(), half * 4); memcpy(&work[oi], im.data(), half * 4);
    }
    const int step = 400 / N;
    for (int k = 0; k < half; ++k) {
        const float re = COS[k * step], im = -SIN[k * step];
        const float re_odd = work[orr + k], im_odd = work[oi + k];
        volatile float p1 = re * re_odd, p2 = im * im_odd, p3 = re * im_odd, p4 = im * re_odd;   // separate roundings, no FMA contraction
        out_re[k] = (work[er + k] + p1) - p2;
        out_im[k] = (work[ei + k] + p3) + p4;
        out_re[k + half] = (work[er + k] - p1) + p2;
        out_im[k + half] = (work[ei + k] - p3) - p4;
    }
}

int wcr_logmel(wcr_model* m, const float* pcm, size_t n, int n_threads, float* out, int* n_len_out, int* n_len_org_out) {
    if (n < 201) { t_err = "log-mel needs more than 200 samples"; return -1; }
    int n_len, n_len_org;
    wcr_logmel_geometry(n, &n_len, &n_len_org);
    const int n_mel = m->hp.n_mels;
    std::vector<float> padded(n + 480000 + 400, 0.f);
    memcpy(padded.data() + 200, pcm, n * 4);
    for (int i = 0; i < 200; ++i) padded[199 - i] = pcm[1 + i];           // reverse_copy(samples + 1, samples + 201, padded.begin())
    float hann[400], SIN[400], COS[400];
    for (int i = 0; i < 400; ++i) {
        hann[i] = (float)(0.5 * (1.0 - cos((2.0 * M_PI * i) / 400)));
        SIN[i] = (float)sin(2.0 * M_PI * i / 400); COS[i] = (float)cos(2.0 * M_PI * i / 400);
    }
    const size_t n_eff = n + 200;
    const int n_frames = (int)std::min<size_t>(n_eff / 160 + 1, (size_t)n_len);
    const float floor_v = (float)log10(1e-10);
    for (size_t i = 0; i < (size_t)n_mel * n_len; ++i) out[i] = floor_v;
    Pool& P = m->threads(n_threads);
    P.run([&](int t, int nt) {
        std::vector<float> work, fr(400), re(400), im(400), power(201);
        for (int f = t; f < n_frames; f += nt) {
            const size_t off = (size_t)f * 160;
            for (int j = 0; j < 400; ++j) fr[j] = off + j < n_eff ? hann[j] * padded[off + j] : 0.f;
            fft_rec(fr.data(), 400, re.data(), im.data(), SIN, COS, work, 0);
            for (int k = 0; k < 201; ++k) { volatile float a = re[k] * re[k], b = im[k] * im[k]; power[k] = a + b; }
            for (int j = 0; j < n_mel; ++j) {
                double sum = 0.0;
                const float* fl = m->mel_filters.data() + (size_t)j * 201;
                for (int k = 0; k < 201; ++k) { volatile float pr = power[k] * fl[k]; sum += (double)pr; }
                out[(size_t)j * n_len + f] = (float)log10(std::max(sum, 1e-10));
            }
        }
    });
    float mx = -1e30f;
    for (size_t i = 0; i < (size_t)n_mel * n_len; ++i) mx = std::max(mx, out[i]);
    const double mmax = (double)mx - 8.0;
    for (size_t i = 0; i < (size_t)n_mel * n_len; ++i) out[i] = (float)((std::max((double)out[i], mmax) + 4.0) / 4.0);
    *n_len_out = n_len; *n_len_org_out = n_len_org;
    return 0;
}

int wcr_encode(wcr_model* m, const float* mel_win, int n_threads, float* enc_out) {
    encode(*m, mel_win, n_threads, enc_out);
    return 0;
}

// ---- logits filter + greedy sampler (oracle/whisper_ref.py process_logits / sample_best) ----
static int process_and_sample(const wcr_model& m, const wcr_cfg& cfg, std::vector<float>& lg, const std::vector<int>& tokens_cur, bool has_ts,
                              int seek_delta, float* margin) {
    const Special& sp = m.sp;
    const int n = m.hp.n_vocab;
    const float NEG = -INFINITY;
    const bool is_initial = tokens_cur.empty();
    if (cfg.suppress_blank && is_initial) { lg[sp.eot] = NEG; lg[sp.blank] = NEG; }
    if (cfg.suppress_nst) for (int id : m.nst_ids) lg[id] = NEG;
    lg[sp.not_] = NEG;
    if (cfg.no_timestamps) for (int i = sp.beg; i < n; ++i) lg[i] = NEG;
    lg[sp.sot] = NEG; lg[sp.nosp] = NEG; lg[sp.solm] = NEG; lg[sp.translate] = NEG; lg[sp.transcribe] = NEG; lg[sp.prev] = NEG;
    for (int i = 0; i < sp.num_languages; ++i) lg[sp.lang_first + i] = NEG;
    const bool last_ts = !tokens_cur.empty() && tokens_cur.back() >= sp.beg;
    const bool pen_ts = tokens_cur.size() < 2 || tokens_cur[tokens_cur.size() - 2] >= sp.beg;
    if (last_ts) {
        if (pen_ts) for (int i = sp.beg; i < n; ++i) lg[i] = NEG;
        else for (int i = 0; i < sp.eot; ++i) lg[i] = NEG;
    }
    if (is_initial && cfg.max_initial_ts > 0.f) {
        const double precision = 30.0 / m.hp.n_audio_ctx;
        const int tid0 = (int)std::lrint(cfg.max_initial_ts / precision);
        for (int i = sp.beg + tid0 + 1; i < n; ++i) lg[i] = NEG;
    }
    if (has_ts) { const int tid0 = seek_delta / 2; for (int i = sp.beg; i < sp.beg + tid0 && i < n; ++i) lg[i] = NEG; }
    float logit_max = NEG;
    for (int i = 0; i < n; ++i) logit_max = std::max(logit_max, lg[i]);
    float se = 0.f;
    for (int i = 0; i < n; ++i) if (lg[i] > NEG) se += expf(lg[i] - logit_max);
    const float lse = logf(se) + logit_max;
    // timestamp-mass rule
    float ts_max = NEG;
    for (int i = sp.beg; i < n; ++i) if (lg[i] > NEG) ts_max = std::max(ts_max, lg[i] - lse);
    float ts_lp = NEG;
    if (ts_max > NEG) {
        float s = 0.f;
        for (int i = sp.beg; i < n; ++i) if (lg[i] > NEG) s += expf((lg[i] - lse) - ts_max);
        if (s > 0.f) ts_lp = logf(s) + ts_max;
    }
    float max_text = NEG;
    for (int i = 0; i < sp.beg; ++i) if (lg[i] > NEG) max_text = std::max(max_text, lg[i] - lse);
    if (ts_lp > max_text) for (int i = 0; i < sp.beg; ++i) lg[i] = NEG;
    // greedy: first (lowest-index) maximum of probs = expf(logprob); margin = top1 - top2 of the filtered logits
    int best = 0; float best_p = -1.f, top1 = NEG, top2 = NEG;
    for (int i = 0; i < n; ++i) {
        const float p = lg[i] == NEG ? 0.f : expf(lg[i] - lse);
        if (p > best_p) { best_p = p; best = i; }
        if (lg[i] > top1) { top2 = top1; top1 = lg[i]; } else if (lg[i] > top2) top2 = lg[i];
    }
    if (margin) *margin = top1 - top2;
    return best;
}

// decode one window whose encoder output is `enc`; tokens -> tokens_out (cap n_max), optional logits trace [n_max][n_vocab]
int wcr_decode_window(wcr_model* m, const float* enc, int seek, int seek_end, const wcr_cfg* cfg, const int32_t* prompt_past, int n_past_ctx,
                      const int32_t* forced, int n_forced, int n_threads, int32_t* tokens_out, float* margins_out, wcr_window* win,
                      float* logits_trace) {
    const HParams& hp = m->hp; const Special& sp = m->sp;
    DecState s;
    cross_kv(*m, enc, n_threads, s);
    std::vector<int> prompt;
    if (n_past_ctx > 0 && cfg->n_max_text_ctx > 0) {
        const int n_take = std::min(std::min(cfg->n_max_text_ctx, hp.n_text_ctx / 2), n_past_ctx);
        prompt.push_back(sp.prev);
        for (int i = n_past_ctx - n_take; i < n_past_ctx; ++i) prompt.push_back(prompt_past[i]);
    }
    prompt.push_back(sp.sot);
    if (hp.n_vocab >= 51865) { prompt.push_back(sp.lang_first + cfg->language_id); prompt.push_back(cfg->translate ? sp.translate : sp.transcribe); }
    if (cfg->no_timestamps) prompt.push_back(sp.not_);
    int n_past = 0;
    for (int tok : prompt) { decode_step(*m, s, tok, n_past, n_threads); ++n_past; }
    int n_max = hp.n_text_ctx / 2 - 4;
    if (cfg->n_max_override > 0) n_max = std::min(n_max, cfg->n_max_override);
    std::vector<int> tokens;
    int seek_delta = 3000, result_len = 0; bool has_ts = false, failed = false;
    std::vector<float> lg(hp.n_vocab);
    for (int i = 0; i < n_max; ++i) {
        if (logits_trace) memcpy(logits_trace + (size_t)i * hp.n_vocab, s.logits.data(), (size_t)hp.n_vocab * 4);
        lg = s.logits;
        float margin = 0.f;
        int tid = process_and_sample(*m, *cfg, lg, tokens, has_ts, seek_delta, &margin);
        if (forced && i < n_forced && forced[i] >= 0) tid = forced[i];
        tokens.push_back(tid);
        if (margins_out) margins_out[i] = margin;
        bool completed = false;
        if (tid > sp.beg) {
            const int sd_new = 2 * (tid - sp.beg);
            if (has_ts && seek_delta > sd_new && result_len < i) { failed = true; break; }
            seek_delta = sd_new; result_len = i + 1; has_ts = true;
        }
        if (tid == sp.eot || (has_ts && seek + seek_delta + 100 >= seek_end)) {
            if (result_len == 0 && !cfg->no_timestamps) {
                if (seek + seek_delta + 100 >= seek_end) result_len = i + 1;
                else { failed = true; break; }
            }
            if (cfg->single_segment || cfg->no_timestamps) { result_len = i + 1; seek_delta = 3000; }
            completed = true;
        }
        if (completed) break;
        if (i == n_max - 1 && (result_len == 0 || seek_delta < 1500)) { failed = true; break; }
        if (n_past >= hp.n_text_ctx) break;
        decode_step(*m, s, tid, n_past, n_threads);
        ++n_past;
    }
    for (size_t i = 0; i < tokens.size(); ++i) tokens_out[i] = tokens[i];
    win->seek = seek; win->n_tokens = (int)tokens.size(); win->result_len = result_len; win->seek_delta = seek_delta; win->failed = failed ? 1 : 0;
    win->n_prompt = (int)prompt.size();
    return 0;
}

// whisper_lang_auto_detect: decode [sot] alone, arg-max over the language logits
int wcr_detect_language(wcr_model* m, const float* enc, int n_threads) {
    DecState s;
    cross_kv(*m, enc, n_threads, s);
    decode_step(*m, s, m->sp.sot, 0, n_threads);
    int best = 0; float bv = -INFINITY;
    for (int i = 0; i < m->sp.num_languages; ++i) { const float v = s.logits[m->sp.lang_first + i]; if (v > bv) { bv = v; best = i; } }
    return best;
}

// whisper_full: log-mel, then the seek loop.  tokens_out: every sampled token of every window (cap entries).
int wcr_full(wcr_model* m, const float* pcm, size_t n, const wcr_cfg* cfg_in, int n_threads, int max_windows, int32_t* tokens_out,
             float* margins_out, int cap, wcr_window* wins, int* n_wins, int* lang_out, double* t_mel, double* t_enc, double* t_dec) {
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    wcr_cfg cfg = *cfg_in;
    *n_wins = 0; *t_mel = *t_enc = *t_dec = 0.0;
    if (n < 201) return 0;
    int n_len, n_len_org;
    wcr_logmel_geometry(n, &n_len, &n_len_org);
    const HParams& hp = m->hp;
    std::vector<float> mel((size_t)hp.n_mels * n_len);
    auto t0 = clk::now();
    if (wcr_logmel(m, pcm, n, n_threads, mel.data(), &n_len, &n_len_org)) return -1;
    *t_mel = secs(t0, clk::now());
    int seek = 0; const int seek_end = n_len_org;
    std::vector<int32_t> past(cfg.initial_prompt, cfg.initial_prompt + (cfg.initial_prompt ? cfg.n_initial_prompt : 0));
    if (seek_end < seek + 100) return 0;
    const int T2 = 2 * hp.n_audio_ctx;
    std::vector<float> win((size_t)hp.n_mels * T2), enc((size_t)hp.n_audio_ctx * hp.n_audio_state);
    int used = 0;
    std::vector<int32_t> wtok(hp.n_text_ctx); std::vector<float> wmar(hp.n_text_ctx);
    while (*n_wins < max_windows) {
        if (seek + 100 >= seek_end) break;
        std::fill(win.begin(), win.end(), 0.f);
        const int i0 = std::min(seek, n_len), i1 = std::min(seek + T2, n_len);
        for (int j = 0; j < hp.n_mels; ++j) memcpy(&win[(size_t)j * T2], &mel[(size_t)j * n_len + i0], (size_t)(i1 - i0) * 4);
        t0 = clk::now();
        encode(*m, win.data(), n_threads, enc.data());
        *t_enc += secs(t0, clk::now());
        t0 = clk::now();
        if (cfg.language_id < 0) cfg.language_id = hp.n_vocab >= 51865 ? wcr_detect_language(m, enc.data(), n_threads) : 0;
        if (seek > 0 && seek + 500 >= seek_end) past.clear();
        wcr_window w{};
        if (wcr_decode_window(m, enc.data(), seek, seek_end, &cfg, past.data(), (int)past.size(), nullptr, 0, n_threads, wtok.data(), wmar.data(), &w, nullptr)) return -1;
        *t_dec += secs(t0, clk::now());
        w.token_offset = used;
        if (used + w.n_tokens > cap) { t_err = "token buffer too small"; return -1; }
        memcpy(tokens_out + used, wtok.data(), (size_t)w.n_tokens * 4);
        if (margins_out) memcpy(margins_out + used, wmar.data(), (size_t)w.n_tokens * 4);
        used += w.n_tokens;
        wins[(*n_wins)++] = w;
        const int n_take = cfg.n_max_text_ctx > 0 ? std::min(std::min(cfg.n_max_text_ctx, hp.n_text_ctx / 2), (int)past.size()) : 0;
        std::vector<int32_t> np(past.end() - n_take, past.end());
        for (int i = 0; i < w.result_len; ++i) np.push_back(wtok[i]);
        past.swap(np);
        seek += w.seek_delta;
    }
    if (lang_out) *lang_out = cfg.language_id;
    return 0;
}

}  // extern "C"
