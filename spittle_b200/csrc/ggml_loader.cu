// GGML legacy Whisper model file reader (see ggml_loader.h).
#include "common.cuh"
#include "ggml_loader.h"
#include <cstdio>
#include <cstring>

namespace sb {

float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ff;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {  // subnormal
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

namespace {
// ggml block-quantised types, 32 weights per block (layouts per ggml [MEM]):
//   q4_0 {f16 d; u8 qs[16]}            x = (q - 8) d       q4_1 {f16 d, m; u8 qs[16]}            x = q d + m
//   q5_0 {f16 d; u8 qh[4]; u8 qs[16]}  x = (q - 16) d      q5_1 {f16 d, m; u8 qh[4]; u8 qs[16]}  x = q d + m
//   q8_0 {f16 d; i8 qs[32]}            x = q d
// element j < 16 = low nibble of qs[j], element j + 16 = high nibble; q5's fifth bit is bit j / j + 16 of qh.
// The weights are expanded to f32 here and then stored in the engine's 16-bit operand type; ggml's CPU kernels
// instead quantise the ACTIVATIONS to 8 bits for these dot products -- that extra noise is not reproduced.
int quant_block_bytes(int ttype) {
    switch (ttype) { case 2: return 18; case 3: return 20; case 6: return 22; case 7: return 24; case 8: return 34; default: return 0; }
}
void dequantize(const uint8_t* src, int ttype, int64_t count, float* dst) {
    const int bs = quant_block_bytes(ttype);
    for (int64_t b = 0; b < count / 32; ++b, src += bs, dst += 32) {
        uint16_t dh; memcpy(&dh, src, 2);
        const float d = f16_bits_to_f32(dh);
        const uint8_t* p = src + 2;
        float m = 0.f;
        if (ttype == 3 || ttype == 7) { uint16_t mh; memcpy(&mh, p, 2); m = f16_bits_to_f32(mh); p += 2; }
        if (ttype == 8) { for (int j = 0; j < 32; ++j) dst[j] = (float)(int8_t)p[j] * d; continue; }
        uint32_t qh = 0;
        if (ttype == 6 || ttype == 7) { memcpy(&qh, p, 4); p += 4; }
        for (int j = 0; j < 16; ++j) {
            int lo = p[j] & 0x0F, hi = p[j] >> 4;
            if (ttype == 6 || ttype == 7) { lo |= (int)((qh >> j) & 1u) << 4; hi |= (int)((qh >> (j + 16)) & 1u) << 4; }
            if (ttype == 2) { dst[j] = (float)(lo - 8) * d; dst[j + 16] = (float)(hi - 8) * d; }
            else if (ttype == 6) { dst[j] = (float)(lo - 16) * d; dst[j + 16] = (float)(hi - 16) * d; }
            else { dst[j] = (float)lo * d + m; dst[j + 16] = (float)hi * d + m; }
        }
    }
}
// k-quant q5_K (the catalog's breeze-asr-q5_k.bin): super-blocks of 256 weights, 176 bytes
//   {f16 d, dmin; u8 scales[12]; u8 qh[32]; u8 qs[128]}, x = d sc_j q - dmin mn_j per 32-weight sub-block j
// (ggml dequantize_row_q5_K / get_scale_min_k4 [MEM])
constexpr int kQkK = 256, kQ5KBytes = 176;
void dequantize_q5_k(const uint8_t* src, int64_t count, float* dst) {
    for (int64_t b = 0; b < count / kQkK; ++b, src += kQ5KBytes, dst += kQkK) {
        uint16_t h; memcpy(&h, src, 2);
        const float d = f16_bits_to_f32(h);
        memcpy(&h, src + 2, 2);
        const float dmin = f16_bits_to_f32(h);
        const uint8_t* sc = src + 4;
        const uint8_t* qh = src + 16;
        const uint8_t* ql = src + 48;
        for (int j = 0; j < 8; ++j) {
            int s, m;
            if (j < 4) { s = sc[j] & 63; m = sc[j + 4] & 63; }
            else { s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
            const float dj = d * (float)s, mj = dmin * (float)m;
            const uint8_t* q = ql + 32 * (j / 2);
            for (int l = 0; l < 32; ++l) {
                const int nib = (j & 1) ? (q[l] >> 4) : (q[l] & 0xF);
                dst[32 * j + l] = dj * (float)(nib + (((qh[l] >> j) & 1) << 4)) - mj;
            }
        }
    }
}
struct Reader {
    const uint8_t* p; size_t n; size_t off = 0; bool ok = true;
    template <typename V> V get() {
        V v{};
        if (sizeof(V) > n - off) { ok = false; return v; }
        memcpy(&v, p + off, sizeof(V));
        off += sizeof(V);
        return v;
    }
    const uint8_t* take(size_t k) {
        if (k > n - off) { ok = false; return nullptr; }      // (off <= n always: no wrap-around)
        const uint8_t* r = p + off;
        off += k;
        return r;
    }
};
}  // namespace

int load_ggml_file(const char* path, GgmlFile& out) {
    if (!path) { set_error("model path is null"); return SB_ERR_INVALID; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error(std::string("cannot open model file: ") + path); return SB_ERR_IO; }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 64) { fclose(f); set_error("model file too small"); return SB_ERR_FORMAT; }
    out.blob.resize((size_t)sz);
    const size_t rd = fread(out.blob.data(), 1, (size_t)sz, f);
    fclose(f);
    if (rd != (size_t)sz) { set_error("short read on model file"); return SB_ERR_IO; }

    Reader r{out.blob.data(), out.blob.size()};
    const uint32_t magic = r.get<uint32_t>();
    if (magic != 0x67676d6cu) {
        char buf[96];
        snprintf(buf, sizeof buf, "bad magic 0x%08x: not a GGML legacy whisper file", magic);
        set_error(buf);
        return SB_ERR_FORMAT;
    }
    int32_t* hp = reinterpret_cast<int32_t*>(&out.hp);
    for (int i = 0; i < 11; ++i) hp[i] = r.get<int32_t>();
    out.n_mel = r.get<int32_t>();
    out.n_fft = r.get<int32_t>();
    if (!r.ok || out.n_mel != out.hp.n_mels || out.n_fft != 201 || out.n_mel <= 0 || out.n_mel > 512) {
        set_error("unexpected mel filter header"); return SB_ERR_FORMAT;
    }
    const uint8_t* mf = r.take((size_t)out.n_mel * out.n_fft * 4);
    if (!mf) { set_error("truncated mel filters"); return SB_ERR_FORMAT; }
    out.mel_filters.resize((size_t)out.n_mel * out.n_fft);
    memcpy(out.mel_filters.data(), mf, out.mel_filters.size() * 4);
    const int32_t nv = r.get<int32_t>();
    if (!r.ok || nv < 0 || nv > out.hp.n_vocab + 1024) { set_error("bad vocab size"); return SB_ERR_FORMAT; }
    out.vocab.resize(nv);
    for (int i = 0; i < nv; ++i) {
        const uint32_t len = r.get<uint32_t>();
        const uint8_t* w = r.take(len);
        if (!r.ok || (len && !w)) { set_error("truncated vocab"); return SB_ERR_FORMAT; }
        out.vocab[i].assign(reinterpret_cast<const char*>(w), len);
    }
    while (r.off < r.n) {
        const int32_t n_dims = r.get<int32_t>();
        const int32_t name_len = r.get<int32_t>();
        const int32_t ttype = r.get<int32_t>();
        if (!r.ok || n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 256) {
            set_error("corrupt tensor header"); return SB_ERR_FORMAT;
        }
        int64_t ne[4] = {1, 1, 1, 1};
        for (int i = 0; i < n_dims; ++i) ne[i] = r.get<int32_t>();
        const uint8_t* nm = r.take(name_len);
        if (!r.ok) { set_error("corrupt tensor header"); return SB_ERR_FORMAT; }
        // a corrupt file must end in SB_ERR_FORMAT, never in a wrapped size, a backwards cursor or std::bad_alloc
        int64_t numel = 1;
        bool dims_ok = true;
        for (int i = 0; i < n_dims; ++i) {
            if (ne[i] <= 0 || ne[i] > (int64_t)1 << 31) { dims_ok = false; break; }
            numel *= ne[i];
            if (numel > 2 * (int64_t)r.n) { dims_ok = false; break; }    // every supported type stores > 0.5 byte per element
        }
        if (!dims_ok) { set_error("corrupt tensor header: bad dimensions"); return SB_ERR_FORMAT; }
        std::string name(reinterpret_cast<const char*>(nm), name_len);
        HostTensor t;
        t.ttype = ttype;
        for (int i = n_dims - 1; i >= 0; --i) t.shape.push_back(ne[i]);   // fastest-first -> torch order
        if (ttype == 0 || ttype == 1) {
            t.nbytes = (size_t)t.numel() * (ttype == 0 ? 4 : 2);
            t.data = r.take(t.nbytes);
            if (!t.data) { set_error("truncated tensor data: " + name); return SB_ERR_FORMAT; }
        } else if (quant_block_bytes(ttype) > 0) {
            if (ne[0] % 32 != 0) { set_error("quantised tensor '" + name + "': row length is not a multiple of 32"); return SB_ERR_FORMAT; }
            const size_t qbytes = (size_t)(t.numel() / 32) * quant_block_bytes(ttype);
            const uint8_t* q = r.take(qbytes);
            if (!q) { set_error("truncated tensor data: " + name); return SB_ERR_FORMAT; }
            out.dequant.emplace_back((size_t)t.numel());
            dequantize(q, ttype, t.numel(), out.dequant.back().data());
            t.ttype = 0;
            t.nbytes = (size_t)t.numel() * 4;
            t.data = reinterpret_cast<const uint8_t*>(out.dequant.back().data());
        } else if (ttype == 13) {
            if (ne[0] % kQkK != 0) { set_error("q5_K tensor '" + name + "': row length is not a multiple of 256"); return SB_ERR_FORMAT; }
            const size_t qbytes = (size_t)(t.numel() / kQkK) * kQ5KBytes;
            const uint8_t* q = r.take(qbytes);
            if (!q) { set_error("truncated tensor data: " + name); return SB_ERR_FORMAT; }
            out.dequant.emplace_back((size_t)t.numel());
            dequantize_q5_k(q, t.numel(), out.dequant.back().data());
            t.ttype = 0;
            t.nbytes = (size_t)t.numel() * 4;
            t.data = reinterpret_cast<const uint8_t*>(out.dequant.back().data());
        } else {
            set_error("tensor '" + name + "' has ggml type " + std::to_string(ttype) +
                      " (supported: f32, f16, q4_0, q4_1, q5_0, q5_1, q8_0, q5_K; the other k-quants are not)");
            return SB_ERR_FORMAT;
        }
        out.tensors[name] = t;
    }
    return SB_OK;
}

}  // namespace sb
