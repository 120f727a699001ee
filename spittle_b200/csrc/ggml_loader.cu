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
struct Reader {
    const uint8_t* p; size_t n; size_t off = 0; bool ok = true;
    template <typename V> V get() {
        V v{};
        if (off + sizeof(V) > n) { ok = false; return v; }
        memcpy(&v, p + off, sizeof(V));
        off += sizeof(V);
        return v;
    }
    const uint8_t* take(size_t k) {
        if (off + k > n) { ok = false; return nullptr; }
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
        std::string name(reinterpret_cast<const char*>(nm), name_len);
        HostTensor t;
        t.ttype = ttype;
        for (int i = n_dims - 1; i >= 0; --i) t.shape.push_back(ne[i]);   // fastest-first -> torch order
        size_t esz;
        if (ttype == 0) esz = 4;
        else if (ttype == 1) esz = 2;
        else {
            set_error("tensor '" + name + "' has quantised type " + std::to_string(ttype) +
                      " (only f32/f16 GGML files are supported so far)");
            return SB_ERR_FORMAT;
        }
        t.nbytes = (size_t)t.numel() * esz;
        t.data = r.take(t.nbytes);
        if (!t.data) { set_error("truncated tensor data: " + name); return SB_ERR_FORMAT; }
        out.tensors[name] = t;
    }
    return SB_OK;
}

}  // namespace sb
