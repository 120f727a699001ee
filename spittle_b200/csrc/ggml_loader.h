// GGML legacy Whisper model file ("ggml-*.bin", magic 0x67676d6c) -- host-side reader.
// The format whisper.cpp defines and the reference loads via
// WhisperEngine::load_model(&path) (src-tauri/src/managers/transcription.rs:262-263);
// restated in SURVEY.md Appendix D.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace sb {

struct HostTensor {
    int ttype = 0;                    // 0 f32, 1 f16 (block-quantised tensors are dequantised to f32 on load)
    std::vector<int64_t> shape;       // torch order (slowest first)
    const uint8_t* data = nullptr;    // points into GgmlFile::blob
    size_t nbytes = 0;
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

struct WhisperHParams {
    int32_t n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_text_ctx, n_text_state, n_text_head, n_text_layer, n_mels, ftype;
};

struct GgmlFile {
    WhisperHParams hp{};
    int n_mel = 0, n_fft = 0;
    std::vector<float> mel_filters;             // [n_mel][n_fft]
    std::vector<std::string> vocab;             // id -> bytes, entries present in the file
    std::map<std::string, HostTensor> tensors;
    std::vector<uint8_t> blob;                  // whole file
    std::vector<std::vector<float>> dequant;    // owned f32 copies of block-quantised tensors
};

// returns SB_OK / SB_ERR_IO / SB_ERR_FORMAT (message via sb::set_error)
int load_ggml_file(const char* path, GgmlFile& out);

float f16_bits_to_f32(uint16_t h);

}  // namespace sb
