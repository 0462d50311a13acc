// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, single thread, f32) of stratum-dsp 1.0.0's per-track
// analysis path (`analyze_audio`, reference src/lib.rs:86-1635) used as the
// parity checker for the CUDA implementation in stratum_dsp_b200/.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may build, link, import or run anything in this directory.
// The product path never includes or links it.
//
// PARITY STATUS: the Rust reference cannot be built here (no cargo/rustc) and
// it ships no end-to-end golden vectors, so this oracle is pinned only by the
// reference's own unit-test known answers and integration-test assertions
// (tests/test_oracle_pins.py).  FFT arithmetic lives in the third-party crate
// rustfft 6.2 (Cargo.toml:18, no Cargo.lock) whose rounding is not
// reproducible; FFT-level parity is therefore "unpinned" — see so_fft.cpp.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace so {

// ---- Rust f32 semantics helpers -----------------------
// f32::max / f32::min ignore NaN; for non-NaN inputs they are plain max/min.
static inline float fmax_rs(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
static inline float fmin_rs(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
// f32::powi with a run-time exponent: llvm.powi.f32 -> compiler-rt __powisf2 (square-and-multiply, reciprocal for b < 0)
static inline float powi_rs(float a, int b) {
    const bool recip = b < 0;
    float r = 1.0f;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}
static inline float clamp_rs(float x, float lo, float hi) {
    // f32::clamp: NaN stays NaN; else if x < lo -> lo; if x > hi -> hi.
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
// `x as usize` from f32: truncate toward zero, saturating, NaN -> 0.
static inline size_t as_usize(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 1.8446744e19f) return SIZE_MAX;
    return (size_t)x;
}
static inline long as_isize(float x) {
    if (!(x == x)) return 0;
    if (x >= 9.2233715e18f) return INT64_MAX;
    if (x <= -9.2233715e18f) return INT64_MIN;
    return (long)x;
}
static inline int as_i32(float x) {
    if (!(x == x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
}
static inline size_t next_pow2(size_t n) {
    size_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
static inline size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

constexpr float PI_F = 3.14159265358979323846f;  // std::f32::consts::PI

// ---- spectrogram: frames x bins, flat ------------------------------------
struct Spec {
    size_t frames = 0, bins = 0;
    std::vector<float> d;
    const float* row(size_t t) const { return d.data() + t * bins; }
    float* row(size_t t) { return d.data() + t * bins; }
    bool empty() const { return frames == 0; }
};

// ---- configuration (reference src/config.rs:8-744; Default at :594-744) ---
enum NormMethod { NORM_PEAK = 0, NORM_RMS = 1, NORM_LOUDNESS = 2 };

struct Config {
    float min_amplitude_db = -40.0f;
    int normalization = NORM_PEAK;
    bool enable_normalization = true;
    bool enable_silence_trimming = true;
    bool enable_onset_consensus = true;
    float onset_threshold_percentile = 0.80f;
    uint32_t onset_consensus_tolerance_ms = 50;
    float onset_consensus_weights[4] = {0.25f, 0.25f, 0.25f, 0.25f};
    bool emit_tempogram_candidates = false;
    bool enable_hpss_onsets = false;
    size_t hpss_margin = 10;
    bool enable_tempogram_percussive_fallback = false;
    bool force_legacy_bpm = false;
    bool enable_bpm_fusion = false;
    bool enable_legacy_bpm_guardrails = true;
    bool enable_tempogram_multi_resolution = true;
    size_t tempogram_multi_res_top_k = 25;
    float tempogram_multi_res_w512 = 0.45f;
    float tempogram_multi_res_w256 = 0.35f;
    float tempogram_multi_res_w1024 = 0.20f;
    float tempogram_multi_res_structural_discount = 0.85f;
    float tempogram_multi_res_double_time_512_factor = 0.92f;
    float tempogram_multi_res_margin_threshold = 0.08f;
    bool tempogram_multi_res_use_human_prior = false;
    bool enable_tempogram_band_fusion = true;
    float tempogram_band_low_max_hz = 200.0f;
    float tempogram_band_mid_max_hz = 2000.0f;
    float tempogram_band_high_max_hz = 8000.0f;
    float tempogram_band_w_full = 0.40f;
    float tempogram_band_w_low = 0.25f;
    float tempogram_band_w_mid = 0.20f;
    float tempogram_band_w_high = 0.15f;
    bool tempogram_band_seed_only = true;
    float tempogram_band_support_threshold = 0.25f;
    float tempogram_band_consensus_bonus = 0.08f;
    float tempogram_novelty_w_spectral = 0.30f;
    float tempogram_novelty_w_energy = 0.35f;
    float tempogram_novelty_w_hfc = 0.35f;
    size_t tempogram_novelty_local_mean_window = 16;
    size_t tempogram_novelty_smooth_window = 5;
    bool enable_tempogram_mel_novelty = true;
    size_t tempogram_mel_n_mels = 40;
    float tempogram_mel_fmin_hz = 30.0f;
    float tempogram_mel_fmax_hz = 8000.0f;
    size_t tempogram_mel_max_filter_bins = 2;
    float tempogram_mel_weight = 0.15f;
    size_t tempogram_superflux_max_filter_bins = 4;
    size_t tempogram_candidates_top_n = 10;
    float legacy_bpm_preferred_min = 72.0f;
    float legacy_bpm_preferred_max = 168.0f;
    float legacy_bpm_soft_min = 60.0f;
    float legacy_bpm_soft_max = 210.0f;
    float legacy_bpm_conf_mul_preferred = 1.30f;
    float legacy_bpm_conf_mul_soft = 0.70f;
    float legacy_bpm_conf_mul_extreme = 0.01f;
    float min_bpm = 40.0f;
    float max_bpm = 240.0f;
    float bpm_resolution = 1.0f;
    size_t frame_size = 2048;
    size_t hop_size = 512;
    bool soft_chroma_mapping = true;
    float soft_mapping_sigma = 0.5f;
    float chroma_sharpening_power = 1.0f;
    bool enable_key_spectrogram_time_smoothing = true;
    size_t key_spectrogram_smooth_margin = 12;
    bool enable_key_frame_weighting = true;
    float key_min_tonalness = 0.0f;
    float key_tonalness_power = 2.0f;
    float key_energy_power = 0.50f;
    bool enable_key_harmonic_mask = true;
    float key_harmonic_mask_power = 2.0f;
    bool enable_key_stft_override = true;
    size_t key_stft_frame_size = 8192;
    size_t key_stft_hop_size = 512;
    bool enable_key_segment_voting = true;
    size_t key_segment_len_frames = 1024;
    size_t key_segment_hop_frames = 512;
    float key_segment_min_clarity = 0.20f;
    bool enable_key_hpcp = true;
    size_t key_hpcp_peaks_per_frame = 24;
    size_t key_hpcp_num_harmonics = 4;
    float key_hpcp_harmonic_decay = 0.60f;
    float key_hpcp_mag_power = 0.50f;
    // a39 variants (config.rs:683-741), all off by default
    int key_template_set = 0;  // TemplateSet: 0 = KrumhanslKessler, 1 = Temperley
    bool enable_key_edge_trim = false;
    float key_edge_trim_fraction = 0.15f;
    bool enable_key_mode_heuristic = false;
    float key_mode_third_ratio_margin = 0.0f;
    float key_mode_flip_min_score_ratio = 0.60f;
    bool enable_key_minor_harmonic_bonus = false;
    float key_minor_leading_tone_bonus_weight = 0.2f;
    bool enable_key_ensemble = false;
    float key_ensemble_kk_weight = 0.5f;
    float key_ensemble_temperley_weight = 0.5f;
    bool enable_key_multi_scale = false;
    uint32_t key_multi_scale_n_lengths = 3;
    uint32_t key_multi_scale_lengths[8] = {120, 360, 720, 0, 0, 0, 0, 0};
    size_t key_multi_scale_hop = 60;
    float key_multi_scale_min_clarity = 0.20f;
    uint32_t key_multi_scale_n_weights = 0;
    float key_multi_scale_weights[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool enable_key_median = false;  // read by nothing in analyze_audio (lib.rs never calls detect_key_median)
    bool enable_key_tuning_compensation = false;
    float key_tuning_max_abs_semitones = 0.08f;
    size_t key_tuning_frame_step = 20;
    float key_tuning_peak_rel_threshold = 0.35f;
    bool enable_key_hpss_harmonic = false;
    size_t key_hpss_frame_step = 4;
    size_t key_hpss_time_margin = 8;
    size_t key_hpss_freq_margin = 8;
    float key_hpss_mask_power = 2.0f;
    bool enable_key_log_frequency = false;
    bool enable_key_beat_synchronous = false;
    bool enable_key_hpcp_whitening = false;
    size_t key_hpcp_whitening_smooth_bins = 31;
    bool enable_key_hpcp_bass_blend = false;
    float key_hpcp_bass_fmin_hz = 55.0f;
    float key_hpcp_bass_fmax_hz = 300.0f;
    float key_hpcp_bass_weight = 0.35f;
};

// ---- result (reference src/analysis/result.rs:7-263) ----------------------
enum ErrKind { OK = 0, INVALID_INPUT = 1, DECODING_ERROR = 2, PROCESSING_ERROR = 3, NOT_IMPLEMENTED = 4, NUMERICAL_ERROR = 5 };

struct Error {
    int kind = OK;
    std::string msg;
    explicit operator bool() const { return kind != OK; }
};

struct BpmEstimate {
    float bpm = 0, confidence = 0;
    uint32_t method_agreement = 0;
};
struct BpmCandidate {
    float bpm = 0, confidence = 0;
};
struct TempoCand {  // TempogramCandidateDebug, tempogram.rs:102-114
    float bpm = 0, score = 0, fft_norm = 0, autocorr_norm = 0;
    bool selected = false;
};

enum WarnBits { WARN_BPM_FAILED = 1, WARN_LOW_GRID = 2, WARN_LOW_KEY_CONF = 4, WARN_LOW_KEY_CLARITY = 8 };
enum FlagBits { FLAG_MULTIMODAL_BPM = 1, FLAG_WEAK_TONALITY = 2, FLAG_TEMPO_VARIATION = 4, FLAG_ONSET_AMBIGUOUS = 8 };

struct Result {
    float bpm = 0, bpm_confidence = 0;
    int key_is_minor = 0;  // Key::Major(i) / Key::Minor(i)
    uint32_t key_index = 0;
    float key_confidence = 0, key_clarity = 0;
    std::vector<float> beats, downbeats, bars;
    float grid_stability = 0;
    // metadata
    float duration_seconds = 0;
    uint32_t sample_rate = 0;
    float onset_method_consensus = 0;
    uint32_t warnings = 0;  // WarnBits (exact strings rebuilt by so::warning_strings)
    uint32_t flags = 0;     // FlagBits
    int multi_res_triggered = -1, multi_res_used = -1;  // Option<bool>: -1 None
    int percussive_triggered = -1, percussive_used = -1;
    // extra integer views for parity checks (not part of the reference struct)
    size_t trim_start = 0, trim_end = 0;
    std::vector<int64_t> onsets;           // onsets_for_beat_tracking (samples)
    std::vector<int32_t> hmm_beat_frames;  // t indices kept by the first HMM pass
    int time_sig_beats_per_bar = 4;
    int beats_refined = 0;  // 1 if the Bayesian per-segment refinement replaced the grid
    int key_hashmap_tie = 0;  // 1 if the returned key label came out of a weighted vote tied at the top (undetermined in the reference)
    bool has_candidates = false;  // metadata.tempogram_candidates: Option<Vec<TempoCandidateDebug>> (lib.rs:684-697, 740-752)
    std::vector<TempoCand> tempogram_candidates;
};

struct Confidence {  // analysis/confidence.rs:32-68
    float bpm_confidence = 0, key_confidence = 0, grid_stability = 0, overall_confidence = 0;
    uint32_t flags = 0;
};

// Named intermediate dump used by the kernel-level parity tests.
struct Dump {
    std::map<std::string, std::vector<float>> f;
    std::map<std::string, std::vector<int64_t>> i;
};

// ---- FFT (so_fft.cpp) -------------------------------------------------------
struct cpx {
    float re, im;
};
void set_fft_variant(int v);  // 0 = SFFT (the parity arithmetic); 1 = float64 rounded once; 2 = f32 radix-2 DIT, no fma (so_fft.cpp)
int fft_variant();
void cfft_forward(std::vector<cpx>& x);                         // in place, size power of two
void rfft_forward(const float* x, size_t n, std::vector<cpx>& X);  // n real (pow2, >=4) -> n/2+1 bins

// ---- preprocessing ------------------------------------------------------------
Error normalize(std::vector<float>& s, int method, float target_lufs, float max_headroom_db, float sample_rate, float* gain_out);
Error calculate_lufs(const std::vector<float>& s, float sample_rate, float* lufs);
Error detect_and_trim(const std::vector<float>& s, uint32_t sr, float threshold_db, uint32_t min_duration_ms, size_t frame_size,
                      size_t* trim_start, size_t* trim_end, std::vector<std::pair<size_t, size_t>>* regions);

// ---- features -------------------------------------------------------------------
void hann_window(size_t n, std::vector<float>& w);
Spec compute_stft(const float* s, size_t n, size_t frame_size, size_t hop);
Error detect_energy_flux_onsets(const float* s, size_t n, size_t frame_size, size_t hop, float threshold_db, std::vector<size_t>& out);
Error detect_spectral_flux_onsets(const Spec& S, float pct, std::vector<size_t>& out, std::vector<float>* flux_out = nullptr);
Error detect_hfc_onsets(const Spec& S, uint32_t sr, float pct, std::vector<size_t>& out, std::vector<float>* flux_out = nullptr);
struct OnsetCand {
    size_t time_samples;
    float confidence;
    uint32_t voted_by;
};
Error hpss_decompose(const Spec& S, size_t margin, Spec& harmonic, Spec& percussive);
Error detect_hpss_onsets(const Spec& P, float pct, std::vector<size_t>& out, std::vector<float>* flux_out = nullptr);
Error vote_onsets(const std::vector<size_t> lists[4], const float weights[4], uint32_t tol_ms, uint32_t sr, std::vector<OnsetCand>& out);

// ---- period ---------------------------------------------------------------------
std::vector<float> superflux_novelty(const Spec& S, size_t k);
std::vector<float> superflux_novelty_band(const Spec& S, size_t k, size_t b0, size_t b1);
std::vector<float> energy_flux_novelty(const Spec& S);
std::vector<float> energy_flux_novelty_band(const Spec& S, size_t b0, size_t b1);
std::vector<float> hfc_novelty(const Spec& S);
std::vector<float> hfc_novelty_band(const Spec& S, size_t b0, size_t b1);
Error mel_superflux_novelty(const Spec& S, uint32_t sr, size_t n_mels, float fmin, float fmax, size_t k, std::vector<float>& out);
std::vector<float> combined_novelty_with_params(const std::vector<float>& s, const std::vector<float>& e, const std::vector<float>& h,
                                                float ws, float we, float wh, size_t local_mean_window, size_t smooth_window);
typedef std::vector<std::pair<float, float>> Tempogram;  // (bpm, value) sorted desc by value (stable)
Error fft_tempogram(const std::vector<float>& nov, uint32_t sr, uint32_t hop, float min_bpm, float max_bpm, Tempogram& out);
Error autocorrelation_tempogram(const std::vector<float>& nov, uint32_t sr, uint32_t hop, float min_bpm, float max_bpm, float res, Tempogram& out);
Error estimate_bpm_tempogram(const Spec& S, uint32_t sr, uint32_t hop, const Config& c, size_t top_n, BpmEstimate& est,
                             std::vector<TempoCand>& cands, Dump* dump = nullptr, const char* tag = "");
Error multi_resolution_tempogram(const float* s, size_t n, uint32_t sr, const Config& c, const Spec* S512, BpmEstimate& est,
                                 std::vector<TempoCand>& c512, Dump* dump = nullptr);
Error estimate_bpm_legacy(const std::vector<size_t>& onsets, uint32_t sr, size_t hop, const Config& c, bool guardrails, bool* has,
                          BpmEstimate& est, Dump* dump = nullptr);

// ---- beat tracking ---------------------------------------------------------------
struct BeatPos {
    float time_seconds, confidence;
    int32_t frame;
};
Error hmm_track_beats(float bpm, const std::vector<float>& onsets, std::vector<BeatPos>& beats, std::vector<int>* path = nullptr);
Error generate_beat_grid(float bpm, float bpm_conf, const std::vector<float>& onsets_s, uint32_t sr, Result& r, Dump* dump = nullptr);

// ---- key ---------------------------------------------------------------------------
Spec harmonic_spectrogram_time_mask(const Spec& K, size_t margin, float power);
Spec smooth_spectrogram_time(const Spec& K, size_t margin);
void extract_chroma(const Spec& K, uint32_t sr, size_t fft_size, bool soft, float sigma, std::vector<float>& chroma, std::vector<float>& energy, float tuning = 0.0f);
void extract_hpcp(const Spec& K, uint32_t sr, size_t fft_size, const Config& c, std::vector<float>& chroma /*frames*12*/, std::vector<float>& energy,
                  float tuning = 0.0f);
float estimate_tuning_offset(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, size_t frame_step, float peak_rel_threshold);
Spec linear_to_log_frequency(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, int* semitone_bin_min_out);
void extract_chroma_log_frequency(const Spec& L, int semitone_offset, std::vector<float>& chroma, std::vector<float>& energy);
void extract_beat_synchronous_chroma(const Spec& K, uint32_t sr, size_t fft_size, size_t hop, const std::vector<float>& beats, bool soft, float sigma,
                                     float tuning, std::vector<float>& chroma, std::vector<float>& energy);
Spec harmonic_spectrogram_hpss_median_mask(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, size_t frame_step, size_t time_margin,
                                           size_t freq_margin, float mask_power);
void smooth_chroma(std::vector<float>& chroma, size_t frames, size_t window);
void sharpen_chroma(float* ch12, float power);
struct KeyScores {
    int keys[24];  // 0..11 major, 12..23 minor, in ranked order
    float scores[24];
    int key;
    float confidence;
    int vote_tie = 0;  // 1: the weighted top-3 vote ran with two keys at exactly the best vote (the reference picks by HashMap order, detector.rs:254-275)
};
Error detect_key_weighted(const float* chroma, size_t frames, const float* w /*nullable*/, KeyScores& out, int template_set = 0);
Error detect_key_weighted_mode_heuristic(const float* chroma, size_t frames, const float* w /*nullable*/, int template_set, float third_ratio_margin,
                                         float flip_min_score_ratio, bool minor_bonus, float minor_bonus_weight, KeyScores& out);
float compute_key_clarity(const float* sorted_scores, size_t n);
void key_templates(float major[12][12], float minor[12][12], int template_set = 0);
Error detect_key_path(const float* s, size_t n, uint32_t sr, const Config& c, const Spec& S_base, Result& r, Dump* dump = nullptr,
                      const std::vector<float>* beat_times = nullptr);

// ---- top level -----------------------------------------------------------------------
Error analyze_audio(const float* samples, size_t n, uint32_t sr, const Config& c, Result& r, Dump* dump = nullptr);
Confidence compute_confidence(const Result& r);
std::vector<std::string> warning_strings(const Result& r);
std::string key_name(int is_minor, uint32_t idx);
std::string key_numerical(int is_minor, uint32_t idx);

}  // namespace so
