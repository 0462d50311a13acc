// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates reference src/preprocessing/normalization.rs and silence.rs.
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;

// KWeightingFilter — normalization.rs:112-175 (single RBJ high-pass biquad, DF-II-T).
struct KWeight {
    float x1 = 0, x2 = 0, b0, b1, b2, a1, a2;
    explicit KWeight(float sr) {
        float w0 = 2.0f * PI_F * 1681.9745f / sr;  // :137
        float cw = cosf(w0), sw = sinf(w0);
        float alpha = sw / 2.0f * sqrtf(1.0f / 0.707f);  // :140
        float B0 = (1.0f + cw) / 2.0f, B1 = -(1.0f + cw), B2 = (1.0f + cw) / 2.0f;
        float a0 = 1.0f + alpha, A1 = -2.0f * cw, A2 = 1.0f - alpha;
        b0 = B0 / a0; b1 = B1 / a0; b2 = B2 / a0; a1 = A1 / a0; a2 = A2 / a0;
    }
    float process(float s) {  // :161-167
        float out = b0 * s + x1;
        x1 = b1 * s + x2 - a1 * out;
        x2 = b2 * s - a2 * out;
        return out;
    }
};

// calculate_lufs — normalization.rs:185-259.  *lufs = -inf when all blocks gated.
Error calculate_lufs(const std::vector<float>& s, float sr, float* lufs) {
    if (s.empty()) return Error{INVALID_INPUT, "Empty audio samples"};
    if (sr <= 0.0f) return Error{INVALID_INPUT, "Invalid sample rate"};
    size_t block = as_usize(sr * 400.0f / 1000.0f);  // :198
    if (block == 0) return Error{INVALID_INPUT, "Sample rate too low for LUFS calculation"};
    KWeight f(sr);
    std::vector<float> filt(s.size());
    for (size_t i = 0; i < s.size(); ++i) filt[i] = f.process(s[i]);
    size_t nb = div_ceil(filt.size(), block);
    std::vector<float> be;
    be.reserve(nb);
    for (size_t i = 0; i < nb; ++i) {
        size_t st = i * block, en = std::min(st + block, filt.size());
        float sum = 0.0f;
        for (size_t j = st; j < en; ++j) sum += filt[j] * filt[j];
        be.push_back(sum / (float)(en - st));
    }
    if (be.empty()) return Error{PROCESSING_ERROR, "No blocks computed for LUFS"};
    float gate = powf(10.0f, (-70.0f + 0.691f) / 10.0f);  // :232
    float acc = 0.0f;
    size_t cnt = 0;
    for (float e : be)
        if (e > gate) { acc += e; ++cnt; }
    if (cnt == 0) { *lufs = -INFINITY; return Error{}; }
    float mean = acc / (float)cnt;
    if (mean <= EPSILON) return Error{NUMERICAL_ERROR, "Mean square too small for LUFS calculation"};
    *lufs = -0.691f + 10.0f * log10f(mean);
    return Error{};
}

static float peak_of(const std::vector<float>& s) {
    float p = 0.0f;
    for (float x : s) p = fmax_rs(p, fabsf(x));
    return p;
}

// normalize_peak — normalization.rs:262-322
static Error normalize_peak(std::vector<float>& s, float headroom_db, float* gain) {
    if (s.empty()) return Error{INVALID_INPUT, "Empty audio samples"};
    float peak = peak_of(s);
    if (peak <= EPSILON) { *gain = 1.0f; return Error{}; }  // no-op (:275-283)
    float target = powf(10.0f, (0.0f - headroom_db) / 20.0f);
    float g = target / peak;
    g = fmin_rs(g, 1.0f / peak);  // :295
    for (float& x : s) x *= g;
    *gain = g;
    return Error{};
}

// normalize_rms — normalization.rs:325-398
static Error normalize_rms(std::vector<float>& s, float target_rms_db, float headroom_db, float* gain) {
    if (s.empty()) return Error{INVALID_INPUT, "Empty audio samples"};
    float sum = 0.0f;
    for (float x : s) sum += x * x;
    float rms = sqrtf(sum / (float)s.size());
    if (rms <= EPSILON) { *gain = 1.0f; return Error{}; }
    float peak = peak_of(s);
    float target = powf(10.0f, (target_rms_db - headroom_db) / 20.0f);
    float g = target / rms;
    if (peak * g > 1.0f) g = 1.0f / peak;  // :362-379
    for (float& x : s) x *= g;
    *gain = g;
    return Error{};
}

// normalize_lufs — normalization.rs:401-484
static Error normalize_lufs(std::vector<float>& s, float target_lufs, float headroom_db, float sr, float* gain) {
    if (s.empty()) return Error{INVALID_INPUT, "Empty audio samples"};
    float measured;
    if (Error e = calculate_lufs(s, sr, &measured)) return e;
    if (measured == -INFINITY) return normalize_peak(s, headroom_db, gain);
    float g = powf(10.0f, (target_lufs - measured) / 20.0f);
    float peak = peak_of(s);
    float target_peak = powf(10.0f, (0.0f - headroom_db) / 20.0f);
    if (peak * g > target_peak) g = target_peak / peak;  // :432-456
    for (float& x : s) x *= g;
    *gain = g;
    return Error{};
}

// normalize — normalization.rs:520-547
Error normalize(std::vector<float>& s, int method, float target_lufs, float headroom_db, float sr, float* gain_out) {
    float g = 1.0f;
    Error e;
    switch (method) {
        case NORM_PEAK: e = normalize_peak(s, headroom_db, &g); break;
        case NORM_RMS: e = normalize_rms(s, target_lufs + 3.0f, headroom_db, &g); break;
        default: e = normalize_lufs(s, target_lufs, headroom_db, sr, &g); break;
    }
    if (gain_out) *gain_out = g;
    return e;
}

// detect_and_trim — silence.rs:102-279.  Returns the [trim_start, trim_end) slice bounds.
Error detect_and_trim(const std::vector<float>& s, uint32_t sr, float threshold_db, uint32_t min_duration_ms, size_t frame_size,
                      size_t* trim_start_o, size_t* trim_end_o, std::vector<std::pair<size_t, size_t>>* regions_o) {
    *trim_start_o = 0;
    *trim_end_o = 0;
    if (s.empty()) return Error{};
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (frame_size == 0) return Error{INVALID_INPUT, "Frame size must be > 0"};
    const size_t n = s.size();
    float thr = powf(10.0f, threshold_db / 20.0f);  // :141
    size_t hop = frame_size / 2;                      // :144
    size_t nf = (n >= frame_size) ? (n - frame_size) / hop + 1 : 1;
    std::vector<char> silent(nf);
    std::vector<size_t> starts(nf);
    for (size_t i = 0; i < nf; ++i) {
        size_t st = i * hop, en = std::min(st + frame_size, n);
        float sum = 0.0f;
        for (size_t j = st; j < en; ++j) sum += s[j] * s[j];  // :159 strict left fold
        float rms = (en > st) ? sqrtf(sum / (float)(en - st)) : 0.0f;
        silent[i] = rms <= thr;
        starts[i] = st;
    }
    size_t min_samples = as_usize((float)min_duration_ms / 1000.0f * (float)sr);  // :179-181
    size_t min_frames = div_ceil(min_samples, hop);
    std::vector<std::pair<size_t, size_t>> regions;
    bool in_sil = false;
    size_t sil_start = 0;
    for (size_t fi = 0; fi < nf; ++fi) {
        if (silent[fi] && !in_sil) {
            in_sil = true;
            sil_start = fi;
        } else if (!silent[fi] && in_sil) {
            in_sil = false;
            size_t sil_end = fi;
            if (sil_end - sil_start >= min_frames || sil_start == 0 || sil_end == nf) {
                size_t a = starts[sil_start];
                size_t b = sil_end < nf ? starts[sil_end] : n;
                regions.emplace_back(a, b);
            }
        }
    }
    if (in_sil) {  // :221-231
        if (nf - sil_start >= min_frames || sil_start == 0) regions.emplace_back(starts[sil_start], n);
    }
    size_t ts = 0, te = n;
    if (!regions.empty() && regions.front().first == 0) ts = regions.front().second;
    if (!regions.empty() && regions.back().second == n) te = regions.back().first;
    ts = std::min(ts, te);  // :255-256
    te = std::max(te, ts);
    if (!(ts < te && te <= n)) { ts = 0; te = 0; }  // empty result
    *trim_start_o = ts;
    *trim_end_o = te;
    if (regions_o) *regions_o = regions;
    return Error{};
}

}  // namespace so
