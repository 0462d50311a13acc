// analyze_batch — the reference's batch example (examples/analyze_batch.rs:196-405) as a plain C++ program over the C ABI
// (include/stratum_b200.h): the same call a Rust / Go / Java binding would make.  No CUDA, torch or Python on this side.
//
//   make -C stratum_dsp_b200 example      (g++ -std=c++17 -Iinclude examples/analyze_batch.cpp -Lstratum_dsp_b200/_build -lstratum_b200 ...)
//   ./analyze_batch [--json] [--devices 0,1,...] a.wav b.wav ...
//
// RIFF/WAVE files (PCM 8/16/24/32 bit, IEEE float 32/64 bit, WAVE_FORMAT_EXTENSIBLE, any channel count) are read with <cstdio> and
// handed to stratum_b200_analyze_batch_pcm undecoded: the per-format conversion and the mono mixdown of the reference's decoder loop
// (analyze_batch.rs:70-165) run on the device.  Output as the
// reference: one JSON object per line with --json (analyze_batch.rs:331-351), else "[i/n] path: BPM=.. Key=..".
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "stratum_b200.h"

struct Wav {
    std::vector<uint8_t> pcm;  // interleaved sample frames as stored in the file
    uint32_t sr = 0, channels = 0, format = 0;  // StratumPcmFormat
    std::string error;
};

static Wav read_wav(const char* path) {
    Wav w;
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        w.error = std::string("cannot open ") + path;
        return w;
    }
    auto rd = [&](void* p, size_t n) { return std::fread(p, 1, n, f) == n; };
    char id[4];
    uint32_t sz = 0;
    if (!rd(id, 4) || std::memcmp(id, "RIFF", 4) || !rd(&sz, 4) || !rd(id, 4) || std::memcmp(id, "WAVE", 4)) w.error = "not a RIFF/WAVE file";
    uint16_t fmt = 0, ch = 0, bits = 0;
    while (w.error.empty() && rd(id, 4) && rd(&sz, 4)) {
        if (!std::memcmp(id, "fmt ", 4)) {
            uint8_t b[40] = {};
            const size_t take = sz < 40 ? sz : 40;
            if (sz < 16 || !rd(b, take)) { w.error = "bad fmt chunk"; break; }
            std::memcpy(&fmt, b, 2);
            std::memcpy(&ch, b + 2, 2);
            std::memcpy(&w.sr, b + 4, 4);
            std::memcpy(&bits, b + 14, 2);
            if (fmt == 0xFFFE && take >= 26) std::memcpy(&fmt, b + 24, 2);  // WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real tag
            std::fseek(f, (long)(sz - take + (sz & 1)), SEEK_CUR);
        } else if (!std::memcmp(id, "data", 4)) {
            if (fmt == 1 && bits == 8) w.format = STRATUM_PCM_U8;
            else if (fmt == 1 && bits == 16) w.format = STRATUM_PCM_S16;
            else if (fmt == 1 && bits == 24) w.format = STRATUM_PCM_S24;
            else if (fmt == 1 && bits == 32) w.format = STRATUM_PCM_S32;
            else if (fmt == 3 && bits == 32) w.format = STRATUM_PCM_F32;
            else if (fmt == 3 && bits == 64) w.format = STRATUM_PCM_F64;
            if (w.format == 0 || ch == 0) { w.error = "Unsupported sample format"; break; }
            w.channels = ch;
            const size_t frame = (size_t)(bits / 8) * ch;
            w.pcm.resize(sz);
            const size_t got = std::fread(w.pcm.data(), 1, sz, f);  // a streamed file may carry a bogus length: keep what is there
            w.pcm.resize(got / frame * frame);
            break;
        } else {
            std::fseek(f, (long)(sz + (sz & 1)), SEEK_CUR);
        }
    }
    if (w.error.empty() && w.pcm.empty()) w.error = "no data chunk";
    std::fclose(f);
    return w;
}

int main(int argc, char** argv) {
    bool json = false;
    std::vector<int32_t> devices;
    std::vector<const char*> paths;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--json")) json = true;
        else if (!std::strcmp(argv[i], "--jobs") && i + 1 < argc) ++i;  // accepted, unused: parallelism is across tracks on the device(s)
        else if (!std::strcmp(argv[i], "--devices") && i + 1 < argc) {
            for (char* t = std::strtok(argv[++i], ","); t; t = std::strtok(nullptr, ",")) devices.push_back(std::atoi(t));
        } else if (!std::strcmp(argv[i], "--help") || !std::strcmp(argv[i], "-h")) {
            std::fprintf(stderr, "usage: %s [--json] [--devices 0,1,...] <file.wav>...\n", argv[0]);
            return 0;
        } else paths.push_back(argv[i]);
    }
    if (paths.empty()) {
        std::fprintf(stderr, "ERROR: Provide at least one audio file path. Use --help for usage.\n");
        return 2;
    }
    if (stratum_b200_sizeof(0) != sizeof(StratumConfig) || stratum_b200_sizeof(1) != sizeof(StratumResult)) {
        std::fprintf(stderr, "ERROR: libstratum_b200 does not match include/stratum_b200.h\n");
        return 1;
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Wav> wavs;
    for (const char* p : paths) wavs.push_back(read_wav(p));
    // tracks that decoded, concatenated: the batch surface takes one buffer + offsets (include/stratum_b200.h)
    std::vector<uint8_t> cat;
    std::vector<uint64_t> offsets(1, 0);  // byte offsets
    std::vector<uint32_t> srs, chans, fmts, index;
    for (size_t i = 0; i < wavs.size(); ++i)
        if (wavs[i].error.empty()) {
            cat.insert(cat.end(), wavs[i].pcm.begin(), wavs[i].pcm.end());
            offsets.push_back(cat.size());
            srs.push_back(wavs[i].sr);
            chans.push_back(wavs[i].channels);
            fmts.push_back(wavs[i].format);
            index.push_back((uint32_t)i);
        }
    std::vector<StratumResult> res(index.size());
    if (!index.empty()) {
        StratumConfig cfg;
        stratum_b200_config_default(&cfg);
        const int32_t st = stratum_b200_analyze_batch_pcm(cat.data(), offsets.data(), srs.data(), chans.data(), fmts.data(), (uint32_t)index.size(), &cfg,
                                                          devices.empty() ? nullptr : devices.data(), (uint32_t)devices.size(), res.data());
        if (st != STRATUM_OK) {
            char msg[512];
            stratum_b200_last_error(msg, sizeof msg);
            std::fprintf(stderr, "ERROR: %s\n", msg);
            return 1;
        }
    }
    // JSON string escaping of serde_json::to_string for the characters paths and messages can hold
    auto jstr = [](const std::string& v) {
        std::string o = "\"";
        for (char ch : v) {
            if (ch == '"' || ch == '\\') { o += '\\'; o += ch; }
            else if (ch == '\n') o += "\\n";
            else if (ch == '\t') o += "\\t";
            else o += ch;
        }
        return o + "\"";
    };
    auto variant = [](int32_t st) {  // Display of AnalysisError (src/error.rs:24-34)
        switch (st) {
            case STRATUM_INVALID_INPUT: return "Invalid input";
            case STRATUM_DECODING_ERROR: return "Decoding error";
            case STRATUM_NOT_IMPLEMENTED: return "Not implemented";
            case STRATUM_NUMERICAL_ERROR: return "Numerical error";
            default: return "Processing error";
        }
    };
    size_t ok = 0, r = 0;
    std::vector<float> times;
    for (size_t i = 0; i < wavs.size(); ++i) {
        const bool decoded = wavs[i].error.empty();
        const StratumResult* R = decoded ? &res[r++] : nullptr;
        std::string err;
        if (!decoded) err = "decode failed: " + wavs[i].error;                                   // analyze_batch.rs:322
        else if (R->status) err = std::string("analysis failed: ") + variant(R->status) + ": " + R->error;  // analyze_batch.rs:306
        if (!err.empty()) {
            if (json) std::printf("{\"file\":%s,\"error\":%s}\n", jstr(paths[i]).c_str(), jstr(err).c_str());
            else std::printf("[%zu/%zu] %s: ERROR: %s\n", i + 1, wavs.size(), paths[i], err.c_str());
            continue;
        }
        ++ok;
        times.push_back(R->processing_time_ms);
        StratumConfidence c;
        stratum_b200_compute_confidence(R, &c);
        char key[16];
        stratum_b200_key_name(R->key_is_minor, R->key_index, 0, key, sizeof key);
        auto opt = [](int32_t v) { return v < 0 ? "null" : (v ? "true" : "false"); };
        if (json)  // analyze_batch.rs:331-345: same keys, order and precision
            std::printf("{\"file\":%s,\"bpm\":%.2f,\"bpm_confidence\":%.4f,\"key\":%s,\"key_confidence\":%.4f,\"processing_time_ms\":%.2f,"
                        "\"tempogram_multi_res_triggered\":%s,\"tempogram_multi_res_used\":%s,\"tempogram_percussive_triggered\":%s,"
                        "\"tempogram_percussive_used\":%s}\n",
                        jstr(paths[i]).c_str(), R->bpm, c.bpm_confidence, jstr(key).c_str(), c.key_confidence, R->processing_time_ms,
                        opt(R->tempogram_multi_res_triggered), opt(R->tempogram_multi_res_used), opt(R->tempogram_percussive_triggered),
                        opt(R->tempogram_percussive_used));
        else  // analyze_batch.rs:355-365
            std::printf("[%zu/%zu] %s: BPM=%.2f (conf=%.3f) Key=%s (conf=%.3f) time=%.2fms\n", i + 1, wavs.size(), paths[i], R->bpm, c.bpm_confidence, key,
                        c.key_confidence, R->processing_time_ms);
    }
    stratum_b200_result_free(res.data(), (uint32_t)res.size());
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "Done: ok=%zu/%zu wall=%.0fms\n", ok, wavs.size(), wall * 1000.0);  // analyze_batch.rs:386-391
    if (!times.empty()) {
        float sum = 0.0f, mn = times[0], mx = 0.0f;
        for (float t : times) {
            sum += t;
            mn = t < mn ? t : mn;
            mx = t > mx ? t : mx;
        }
        std::fprintf(stderr, "processing_time_ms: mean=%.2f min=%.2f max=%.2f\n", sum / (float)times.size(), mn, mx);
    }
    stratum_b200_shutdown();
    return ok == wavs.size() ? 0 : 1;
}
