// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Plain C entry points so tests/ and bench.py can drive the oracle through ctypes.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include "so_common.hpp"

using namespace so;

struct SoHandle {
    Result r;
    Dump d;
    Error e;
};

#define SO_FIELDS(X)                                                                                                                     \
    X(min_amplitude_db) X(normalization) X(enable_normalization) X(enable_silence_trimming) X(enable_onset_consensus)                     \
    X(onset_threshold_percentile) X(onset_consensus_tolerance_ms) X(emit_tempogram_candidates) X(enable_hpss_onsets) X(hpss_margin) X(enable_tempogram_percussive_fallback) X(force_legacy_bpm) X(enable_bpm_fusion) X(enable_legacy_bpm_guardrails) \
    X(enable_tempogram_multi_resolution) X(tempogram_multi_res_top_k) X(tempogram_multi_res_w512) X(tempogram_multi_res_w256)             \
    X(tempogram_multi_res_w1024) X(tempogram_multi_res_structural_discount) X(tempogram_multi_res_double_time_512_factor)                 \
    X(tempogram_multi_res_margin_threshold) X(tempogram_multi_res_use_human_prior) X(enable_tempogram_band_fusion)                        \
    X(tempogram_band_low_max_hz) X(tempogram_band_mid_max_hz) X(tempogram_band_high_max_hz) X(tempogram_band_w_full)                      \
    X(tempogram_band_w_low) X(tempogram_band_w_mid) X(tempogram_band_w_high) X(tempogram_band_seed_only)                                  \
    X(tempogram_band_support_threshold) X(tempogram_band_consensus_bonus) X(tempogram_novelty_w_spectral) X(tempogram_novelty_w_energy)   \
    X(tempogram_novelty_w_hfc) X(tempogram_novelty_local_mean_window) X(tempogram_novelty_smooth_window) X(enable_tempogram_mel_novelty)  \
    X(tempogram_mel_n_mels) X(tempogram_mel_fmin_hz) X(tempogram_mel_fmax_hz) X(tempogram_mel_max_filter_bins) X(tempogram_mel_weight)    \
    X(tempogram_superflux_max_filter_bins) X(tempogram_candidates_top_n) X(legacy_bpm_preferred_min) X(legacy_bpm_preferred_max)          \
    X(legacy_bpm_soft_min) X(legacy_bpm_soft_max) X(legacy_bpm_conf_mul_preferred) X(legacy_bpm_conf_mul_soft)                            \
    X(legacy_bpm_conf_mul_extreme) X(min_bpm) X(max_bpm) X(bpm_resolution) X(frame_size) X(hop_size) X(soft_chroma_mapping) X(soft_mapping_sigma) X(chroma_sharpening_power) X(enable_key_spectrogram_time_smoothing)                \
    X(key_spectrogram_smooth_margin) X(enable_key_frame_weighting) X(key_min_tonalness) X(key_tonalness_power) X(key_energy_power)        \
    X(enable_key_harmonic_mask) X(key_harmonic_mask_power) X(enable_key_stft_override) X(key_stft_frame_size) X(key_stft_hop_size)        \
    X(enable_key_segment_voting) X(key_segment_len_frames) X(key_segment_hop_frames) X(key_segment_min_clarity) X(enable_key_hpcp)        \
    X(key_hpcp_peaks_per_frame) X(key_hpcp_num_harmonics) X(key_hpcp_harmonic_decay) X(key_hpcp_mag_power)                                \
    X(key_template_set) X(enable_key_edge_trim) X(key_edge_trim_fraction) X(enable_key_mode_heuristic) X(key_mode_third_ratio_margin)     \
    X(key_mode_flip_min_score_ratio) X(enable_key_minor_harmonic_bonus) X(key_minor_leading_tone_bonus_weight) X(enable_key_ensemble)     \
    X(key_ensemble_kk_weight) X(key_ensemble_temperley_weight) X(enable_key_multi_scale) X(key_multi_scale_n_lengths)                     \
    X(key_multi_scale_hop) X(key_multi_scale_min_clarity) X(key_multi_scale_n_weights) X(enable_key_median)                               \
    X(enable_key_tuning_compensation) X(key_tuning_max_abs_semitones) X(key_tuning_frame_step) X(key_tuning_peak_rel_threshold)           \
    X(enable_key_hpss_harmonic) X(key_hpss_frame_step) X(key_hpss_time_margin) X(key_hpss_freq_margin) X(key_hpss_mask_power)             \
    X(enable_key_log_frequency) X(enable_key_beat_synchronous) X(enable_key_hpcp_whitening) X(key_hpcp_whitening_smooth_bins)             \
    X(enable_key_hpcp_bass_blend) X(key_hpcp_bass_fmin_hz) X(key_hpcp_bass_fmax_hz) X(key_hpcp_bass_weight)

template <class T>
static void assign(T& dst, double v) { dst = (T)v; }
static void assign(bool& dst, double v) { dst = v != 0.0; }

extern "C" {

void* so_config_new() { return new Config(); }
void so_config_free(void* c) { delete (Config*)c; }
int so_config_set(void* cp, const char* name, double v) {
    Config& c = *(Config*)cp;
#define X(f) \
    if (!strcmp(name, #f)) { assign(c.f, v); return 0; }
    SO_FIELDS(X)
#undef X
    // array fields: "name[i]"
    unsigned idx = 0;
    if (sscanf(name, "key_multi_scale_lengths[%u]", &idx) == 1 && idx < 8) { c.key_multi_scale_lengths[idx] = (uint32_t)v; return 0; }
    if (sscanf(name, "key_multi_scale_weights[%u]", &idx) == 1 && idx < 8) { c.key_multi_scale_weights[idx] = (float)v; return 0; }
    if (sscanf(name, "onset_consensus_weights[%u]", &idx) == 1 && idx < 4) { c.onset_consensus_weights[idx] = (float)v; return 0; }
    return -1;
}
double so_config_get(void* cp, const char* name) {
    Config& c = *(Config*)cp;
#define X(f) \
    if (!strcmp(name, #f)) return (double)c.f;
    SO_FIELDS(X)
#undef X
    return NAN;
}

void* so_analyze(const float* samples, uint64_t n, uint32_t sr, void* cfg, int want_dump) {
    SoHandle* h = new SoHandle();
    Config def;
    const Config& c = cfg ? *(Config*)cfg : def;
    h->e = analyze_audio(samples, (size_t)n, sr, c, h->r, want_dump ? &h->d : nullptr);
    const Result& r = h->r;
    h->d.f["result.beats"] = r.beats;
    h->d.f["result.downbeats"] = r.downbeats;
    h->d.f["result.bars"] = r.bars;
    h->d.i["result.onsets"] = r.onsets;
    if (r.has_candidates) {
        std::vector<float> flat;
        for (auto& c : r.tempogram_candidates) {
            flat.push_back(c.bpm);
            flat.push_back(c.score);
            flat.push_back(c.fft_norm);
            flat.push_back(c.autocorr_norm);
            flat.push_back(c.selected ? 1.0f : 0.0f);
        }
        h->d.f["result.candidates"] = flat;
    }
    h->d.i["result.hmm_beat_frames"] = std::vector<int64_t>(r.hmm_beat_frames.begin(), r.hmm_beat_frames.end());
    return h;
}
void so_free(void* h) { delete (SoHandle*)h; }
int so_status(void* h) { return ((SoHandle*)h)->e.kind; }
int so_error_message(void* h, char* buf, int cap) {
    const std::string& m = ((SoHandle*)h)->e.msg;
    if (cap > 0) {
        strncpy(buf, m.c_str(), cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)m.size();
}
double so_scalar(void* hp, const char* name) {
    const Result& r = ((SoHandle*)hp)->r;
#define S(f) \
    if (!strcmp(name, #f)) return (double)r.f;
    S(bpm) S(bpm_confidence) S(key_is_minor) S(key_index) S(key_confidence) S(key_clarity) S(grid_stability) S(duration_seconds)
    S(sample_rate) S(onset_method_consensus) S(warnings) S(flags) S(multi_res_triggered) S(multi_res_used) S(percussive_triggered)
    S(percussive_used) S(trim_start) S(trim_end) S(time_sig_beats_per_bar) S(beats_refined) S(key_hashmap_tie)
#undef S
    return NAN;
}
void so_confidence(void* hp, float out[4], uint32_t* flags) {
    Confidence c = compute_confidence(((SoHandle*)hp)->r);
    out[0] = c.bpm_confidence;
    out[1] = c.key_confidence;
    out[2] = c.grid_stability;
    out[3] = c.overall_confidence;
    *flags = c.flags;
}
// compute_confidence on a hand-built result (reference unit tests confidence.rs:340-422)
void so_confidence_of(float bpm, float bpm_conf, float key_conf, float key_clarity, float grid_stability, uint32_t warnings, uint32_t flags_in,
                      float out[4], uint32_t* flags) {
    Result r;
    r.bpm = bpm;
    r.bpm_confidence = bpm_conf;
    r.key_confidence = key_conf;
    r.key_clarity = key_clarity;
    r.grid_stability = grid_stability;
    r.warnings = warnings;
    r.flags = flags_in;
    Confidence c = compute_confidence(r);
    out[0] = c.bpm_confidence;
    out[1] = c.key_confidence;
    out[2] = c.grid_stability;
    out[3] = c.overall_confidence;
    *flags = c.flags;
}
int so_warning_strings(void* hp, char* buf, int cap) {
    std::string all;
    for (auto& s : warning_strings(((SoHandle*)hp)->r)) all += s + "\n";
    if (cap > 0) {
        strncpy(buf, all.c_str(), cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)all.size();
}
int64_t so_farray_len(void* hp, const char* name) {
    auto& m = ((SoHandle*)hp)->d.f;
    auto it = m.find(name);
    return it == m.end() ? -1 : (int64_t)it->second.size();
}
int64_t so_farray_copy(void* hp, const char* name, float* out, int64_t cap) {
    auto& m = ((SoHandle*)hp)->d.f;
    auto it = m.find(name);
    if (it == m.end()) return -1;
    int64_t n = std::min<int64_t>(cap, (int64_t)it->second.size());
    memcpy(out, it->second.data(), n * sizeof(float));
    return n;
}
int64_t so_iarray_len(void* hp, const char* name) {
    auto& m = ((SoHandle*)hp)->d.i;
    auto it = m.find(name);
    return it == m.end() ? -1 : (int64_t)it->second.size();
}
int64_t so_iarray_copy(void* hp, const char* name, int64_t* out, int64_t cap) {
    auto& m = ((SoHandle*)hp)->d.i;
    auto it = m.find(name);
    if (it == m.end()) return -1;
    int64_t n = std::min<int64_t>(cap, (int64_t)it->second.size());
    memcpy(out, it->second.data(), n * sizeof(int64_t));
    return n;
}
int so_array_names(void* hp, char* buf, int cap) {
    std::string all;
    for (auto& kv : ((SoHandle*)hp)->d.f) all += "f:" + kv.first + "\n";
    for (auto& kv : ((SoHandle*)hp)->d.i) all += "i:" + kv.first + "\n";
    if (cap > 0) {
        strncpy(buf, all.c_str(), cap - 1);
        buf[cap - 1] = 0;
    }
    return (int)all.size();
}

void so_set_fft_variant(int v) { set_fft_variant(v); }
int so_fft_variant() { return fft_variant(); }

// ---- stage-level entry points -------------------------------------------------------------
void so_cfft(float* reim, uint64_t m) {  // interleaved re,im; in place
    std::vector<cpx> x(m);
    memcpy(x.data(), reim, m * sizeof(cpx));
    cfft_forward(x);
    memcpy(reim, x.data(), m * sizeof(cpx));
}
void so_rfft(const float* x, uint64_t n, float* out_reim /* (n/2+1)*2 */) {
    std::vector<cpx> X;
    rfft_forward(x, n, X);
    memcpy(out_reim, X.data(), X.size() * sizeof(cpx));
}
int64_t so_stft(const float* s, uint64_t n, uint64_t frame, uint64_t hop, float* out, uint64_t cap) {
    Spec S = compute_stft(s, n, frame, hop);
    if (out && S.d.size() <= cap) memcpy(out, S.d.data(), S.d.size() * sizeof(float));
    return (int64_t)S.frames;
}
// normalisation gain + trim bounds only
int so_preprocess(const float* s, uint64_t n, uint32_t sr, int method, float* gain, uint64_t* ts, uint64_t* te) {
    std::vector<float> v(s, s + n);
    Error e = normalize(v, method, -14.0f, 1.0f, (float)sr, gain);
    if (e) return e.kind;
    size_t a, b;
    e = detect_and_trim(v, sr, -40.0f, 500, 2048, &a, &b, nullptr);
    *ts = a;
    *te = b;
    return e.kind;
}
int so_lufs(const float* s, uint64_t n, float sr, float* lufs) {
    std::vector<float> v(s, s + n);
    return calculate_lufs(v, sr, lufs).kind;
}
int so_vote_onsets(const int64_t* a, int na, const int64_t* b, int nb, const int64_t* c, int nc, const int64_t* d, int nd, const float* w,
                   uint32_t tol_ms, uint32_t sr, int64_t* out_centre, float* out_conf, uint32_t* out_voted, int cap) {
    std::vector<size_t> lists[4];
    lists[0].assign(a, a + na);
    lists[1].assign(b, b + nb);
    lists[2].assign(c, c + nc);
    lists[3].assign(d, d + nd);
    std::vector<OnsetCand> o;
    Error e = vote_onsets(lists, w, tol_ms, sr, o);
    if (e) return -e.kind;
    int n = std::min<int>(cap, (int)o.size());
    for (int i = 0; i < n; ++i) {
        out_centre[i] = (int64_t)o[i].time_samples;
        out_conf[i] = o[i].confidence;
        out_voted[i] = o[i].voted_by;
    }
    return (int)o.size();
}
int so_detect_key(const float* chroma, uint64_t frames, const float* w, int* key, float* conf, float* clarity, float* scores24, int* order24) {
    KeyScores ks;
    Error e = detect_key_weighted(chroma, frames, w, ks);
    if (e) return e.kind;
    *key = ks.key;
    *conf = ks.confidence;
    *clarity = compute_key_clarity(ks.scores, 24);
    if (scores24) memcpy(scores24, ks.scores, sizeof ks.scores);
    if (order24) memcpy(order24, ks.keys, sizeof ks.keys);
    return 0;
}
float so_key_clarity(const float* scores, int n) { return compute_key_clarity(scores, n); }
void so_key_templates(float* major144, float* minor144) {
    float a[12][12], b[12][12];
    key_templates(a, b);
    memcpy(major144, a, sizeof a);
    memcpy(minor144, b, sizeof b);
}
int so_key_name(int is_minor, uint32_t idx, int numerical, char* buf, int cap) {
    std::string s = numerical ? key_numerical(is_minor, idx) : key_name(is_minor, idx);
    strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
    return (int)s.size();
}
int so_hmm(float bpm, const float* onsets, int n, int32_t* frames, float* times, int cap, int* path, int path_cap, int* path_len) {
    std::vector<float> o(onsets, onsets + n);
    std::vector<BeatPos> b;
    std::vector<int> p;
    Error e = hmm_track_beats(bpm, o, b, &p);
    if (e) return -e.kind;
    int m = std::min<int>(cap, (int)b.size());
    for (int i = 0; i < m; ++i) {
        frames[i] = b[i].frame;
        times[i] = b[i].time_seconds;
    }
    if (path_len) *path_len = (int)p.size();
    if (path)
        for (int i = 0; i < std::min<int>(path_cap, (int)p.size()); ++i) path[i] = p[i];
    return (int)b.size();
}
int so_beat_grid(float bpm, float conf, const float* onsets, int n, uint32_t sr, float* stability, int* n_beats, int* n_down, int* bpb) {
    std::vector<float> o(onsets, onsets + n);
    Result r;
    Error e = generate_beat_grid(bpm, conf, o, sr, r, nullptr);
    if (e) return e.kind;
    *stability = r.grid_stability;
    *n_beats = (int)r.beats.size();
    *n_down = (int)r.downbeats.size();
    *bpb = r.time_sig_beats_per_bar;
    return 0;
}

int so_hpss_decompose(const float* spec, uint64_t frames, uint64_t bins, uint64_t margin, float* h_out, float* p_out) {
    Spec S, H, P;
    S.frames = frames;
    S.bins = bins;
    S.d.assign(spec, spec + frames * bins);
    Error e = hpss_decompose(S, margin, H, P);
    if (e) return e.kind;
    memcpy(h_out, H.d.data(), H.d.size() * sizeof(float));
    memcpy(p_out, P.d.data(), P.d.size() * sizeof(float));
    return 0;
}
int so_hpss_onsets(const float* perc, uint64_t frames, uint64_t bins, float pct, int64_t* out, int cap) {
    Spec P;
    P.frames = frames;
    P.bins = bins;
    P.d.assign(perc, perc + frames * bins);
    std::vector<size_t> on;
    Error e = detect_hpss_onsets(P, pct, on, nullptr);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)on.size()); ++i) out[i] = (int64_t)on[i];
    return (int)on.size();
}

// ---- batch timing (CPU baseline; mirrors examples/analyze_batch.rs:239-326 with a std::thread pool) ----
double so_batch_timed(const float* samples, const uint64_t* offsets, const uint32_t* srs, uint32_t n_tracks, uint32_t jobs, float* bpm_out,
                      int* key_out, double* per_track_ms) {
    if (jobs == 0) jobs = 1;
    std::atomic<uint32_t> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&]() {
        Config c;
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= n_tracks) break;
            Result r;
            auto a = std::chrono::steady_clock::now();
            Error e = analyze_audio(samples + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), srs[i], c, r, nullptr);
            Confidence cf = compute_confidence(r);
            (void)cf;
            auto b = std::chrono::steady_clock::now();
            if (per_track_ms) per_track_ms[i] = std::chrono::duration<double, std::milli>(b - a).count();
            if (bpm_out) bpm_out[i] = e ? -1.0f : r.bpm;
            if (key_out) key_out[i] = e ? -1 : (int)(r.key_is_minor * 12 + r.key_index);
        }
    };
    std::vector<std::thread> th;
    for (uint32_t j = 1; j < jobs; ++j) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
