// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Unit-level C entry points: one per reference function whose own #[test] known answers tests/test_oracle_ref_units.py
// re-runs on this restatement (entry points that need file-local helpers live at the end of so_beat.cpp / so_legacy.cpp).
#include <algorithm>
#include <cstring>

#include "so_common.hpp"

using namespace so;

static Spec spec_of(const float* d, uint64_t frames, uint64_t bins) {
    Spec S;
    S.frames = frames;
    S.bins = bins;
    S.d.assign(d, d + frames * bins);
    return S;
}

extern "C" {

// detect_spectral_flux_onsets (kind 0, onset/spectral_flux.rs:82-221) / detect_hfc_onsets (kind 1, onset/hfc.rs:84-215): frame indices
int so_u_spec_onsets(int kind, const float* spec, uint64_t frames, uint64_t bins, uint32_t sr, float pct, int64_t* out, int cap) {
    Spec S = spec_of(spec, frames, bins);
    std::vector<size_t> on;
    Error e = kind == 0 ? detect_spectral_flux_onsets(S, pct, on) : detect_hfc_onsets(S, sr, pct, on);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)on.size()); ++i) out[i] = (int64_t)on[i];
    return (int)on.size();
}
// detect_energy_flux_onsets — onset/energy_flux.rs:67-243: sample positions
int so_u_energy_onsets(const float* s, uint64_t n, uint64_t frame, uint64_t hop, float thr_db, int64_t* out, int cap) {
    std::vector<size_t> on;
    Error e = detect_energy_flux_onsets(s, (size_t)n, (size_t)frame, (size_t)hop, thr_db, on);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)on.size()); ++i) out[i] = (int64_t)on[i];
    return (int)on.size();
}
// novelty curves of period/novelty.rs: kind 0 superflux (max filter k), 1 energy flux, 2 HFC
int so_u_novelty(int kind, const float* spec, uint64_t frames, uint64_t bins, uint64_t k, float* out, int cap) {
    Spec S = spec_of(spec, frames, bins);
    std::vector<float> v = kind == 0 ? superflux_novelty(S, (size_t)k) : (kind == 1 ? energy_flux_novelty(S) : hfc_novelty(S));
    for (int i = 0; i < std::min<int>(cap, (int)v.size()); ++i) out[i] = v[i];
    return (int)v.size();
}
int so_u_combined_novelty(const float* s, int ns, const float* e, int ne, const float* h, int nh, float ws, float we, float wh, uint64_t lmw, uint64_t smw,
                          float* out, int cap) {
    std::vector<float> v = combined_novelty_with_params(std::vector<float>(s, s + ns), std::vector<float>(e, e + ne), std::vector<float>(h, h + nh), ws, we, wh,
                                                        (size_t)lmw, (size_t)smw);
    for (int i = 0; i < std::min<int>(cap, (int)v.size()); ++i) out[i] = v[i];
    return (int)v.size();
}
// fft_tempogram (kind 0, period/tempogram_fft.rs:80-236) / autocorrelation_tempogram (kind 1, period/tempogram_autocorr.rs:74-222): (bpm, value) by value desc
int so_u_tempogram(int kind, const float* nov, int n, uint32_t sr, uint32_t hop, float min_bpm, float max_bpm, float res, float* bpm, float* val, int cap) {
    Tempogram t;
    std::vector<float> v(nov, nov + n);
    Error e = kind == 0 ? fft_tempogram(v, sr, hop, min_bpm, max_bpm, t) : autocorrelation_tempogram(v, sr, hop, min_bpm, max_bpm, res, t);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)t.size()); ++i) {
        bpm[i] = t[i].first;
        val[i] = t[i].second;
    }
    return (int)t.size();
}
// estimate_bpm_tempogram on a hand-built spectrogram (period/tempogram.rs; the band-fusion form analyze_audio calls)
int so_u_estimate_tempogram(const float* spec, uint64_t frames, uint64_t bins, uint32_t sr, uint32_t hop, void* cfg, float* bpm, float* conf, uint32_t* agree) {
    Spec S = spec_of(spec, frames, bins);
    Config def;
    const Config& c = cfg ? *(Config*)cfg : def;
    BpmEstimate est;
    std::vector<TempoCand> cands;
    Error e = estimate_bpm_tempogram(S, sr, hop, c, 0, est, cands);
    if (e) return e.kind;
    *bpm = est.bpm;
    *conf = est.confidence;
    *agree = est.method_agreement;
    return 0;
}
// extract_chroma_with_options (chroma/extractor.rs:186-299) = compute_stft + frame_to_chroma per frame
int64_t so_u_extract_chroma(const float* s, uint64_t n, uint32_t sr, uint64_t frame, uint64_t hop, int soft, float sigma, float* out, uint64_t cap_frames) {
    Spec K = compute_stft(s, (size_t)n, (size_t)frame, (size_t)hop);
    std::vector<float> chroma, energy;
    extract_chroma(K, sr, (size_t)frame, soft != 0, sigma, chroma, energy);
    const uint64_t nf = chroma.size() / 12;
    if (out) memcpy(out, chroma.data(), sizeof(float) * 12 * std::min<uint64_t>(nf, cap_frames));
    return (int64_t)nf;
}
// frame_to_chroma on one magnitude frame (extractor.rs:393-487)
void so_u_frame_to_chroma(const float* mag, uint64_t bins, uint32_t sr, uint64_t fft_size, int soft, float sigma, float* out12) {
    Spec K = spec_of(mag, 1, bins);
    std::vector<float> chroma, energy;
    extract_chroma(K, sr, (size_t)fft_size, soft != 0, sigma, chroma, energy);
    memcpy(out12, chroma.data(), sizeof(float) * 12);
}
void so_u_smooth_chroma(float* chroma, uint64_t frames, uint64_t window) {  // chroma/smoothing.rs:37-94, in place
    std::vector<float> v(chroma, chroma + frames * 12);
    smooth_chroma(v, (size_t)frames, (size_t)window);
    memcpy(chroma, v.data(), sizeof(float) * v.size());
}
void so_u_sharpen_chroma(float* ch12, float power) { sharpen_chroma(ch12, power); }  // chroma/normalization.rs:41-65
// normalize (preprocessing/normalization.rs:520-547), in place; method 0 peak, 1 RMS, 2 loudness
int so_u_normalize(float* s, uint64_t n, int method, float target_lufs, float headroom_db, float sr, float* gain) {
    std::vector<float> v(s, s + n);
    Error e = normalize(v, method, target_lufs, headroom_db, sr, gain);
    if (e) return e.kind;
    memcpy(s, v.data(), sizeof(float) * n);
    return 0;
}
// detect_and_trim (preprocessing/silence.rs:107-256): trim range + silence regions (pairs)
int so_u_trim(const float* s, uint64_t n, uint32_t sr, float thr_db, uint32_t min_ms, uint64_t frame_size, uint64_t* ts, uint64_t* te, uint64_t* regions, int cap) {
    size_t a = 0, b = 0;
    std::vector<std::pair<size_t, size_t>> reg;
    Error e = detect_and_trim(std::vector<float>(s, s + n), sr, thr_db, min_ms, (size_t)frame_size, &a, &b, &reg);
    if (e) return -e.kind;
    *ts = a;
    *te = b;
    for (int i = 0; i < std::min<int>(cap, (int)reg.size()); ++i) {
        regions[2 * i] = reg[i].first;
        regions[2 * i + 1] = reg[i].second;
    }
    return (int)reg.size();
}
// harmonic_spectrogram_time_mask / smooth_spectrogram_time (chroma/extractor.rs:1246-1349) on a small spectrogram
void so_u_time_mask(const float* spec, uint64_t frames, uint64_t bins, uint64_t margin, float power, int smooth_only, float* out) {
    Spec K = spec_of(spec, frames, bins);
    Spec R = smooth_only ? smooth_spectrogram_time(K, (size_t)margin) : harmonic_spectrogram_time_mask(K, (size_t)margin, power);
    memcpy(out, R.d.data(), sizeof(float) * R.d.size());
}

}  // extern "C"
