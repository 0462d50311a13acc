// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates reference src/features/period/tempogram.rs:255-775 and
// multi_resolution.rs:205-901.
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static size_t hz_to_bin(float hz, float res, size_t n_bins) {  // tempogram.rs:279-289
    if (!std::isfinite(hz) || hz <= 0.0f || !std::isfinite(res) || res <= 0.0f) return 0;
    long b = as_isize(roundf(hz / res));
    long hi = (long)n_bins - 1;
    if (hi < 0) hi = 0;
    return (size_t)std::max<long>(0, std::min<long>(b, hi));
}

static float lookup_nearest(const Tempogram& t, float bpm, float tol) {  // tempogram.rs:518-529
    float best_d = INFINITY, best_v = 0.0f;
    for (auto& e : t) {
        float d = fabsf(e.first - bpm);
        if (d <= tol && d < best_d) {
            best_d = d;
            best_v = e.second;
        }
    }
    return best_v;
}

struct Variant {
    const char* name;
    float w;
    Tempogram fft, ac;
    float max_fft, max_ac;
};

static void dump_curve(Dump* d, const std::string& key, const std::vector<float>& v) {
    if (d) d->f[key] = v;
}
static void dump_tempogram(Dump* d, const std::string& key, const Tempogram& t) {
    if (!d) return;
    std::vector<float> b, v;
    for (auto& e : t) {
        b.push_back(e.first);
        v.push_back(e.second);
    }
    d->f[key + ".bpm"] = b;
    d->f[key + ".val"] = v;
}

// estimate_bpm_tempogram_impl with band_cfg = Some(..) built from the config (lib.rs:346-373);
// candidates truncated to top_n as in estimate_bpm_tempogram_with_candidates_band_fusion (:226-251).
Error estimate_bpm_tempogram(const Spec& S, uint32_t sr, uint32_t hop, const Config& c, size_t top_n, BpmEstimate& est,
                             std::vector<TempoCand>& cands, Dump* dump, const char* tag) {
    cands.clear();
    const size_t n_bins = S.frames ? S.bins : 0;
    if (n_bins == 0) return Error{INVALID_INPUT, "Empty magnitude frames"};
    const float min_bpm = c.min_bpm, max_bpm = c.max_bpm, res_bpm = c.bpm_resolution;
    size_t fft_size = std::max<size_t>((n_bins - 1) * 2, 2);
    float freq_res = (float)sr / (float)fft_size;
    const std::string T = tag;

    const size_t sf_k = c.tempogram_superflux_max_filter_bins;
    auto spectral_full = superflux_novelty(S, sf_k);
    auto energy_full = energy_flux_novelty(S);
    auto hfc_full = hfc_novelty(S);
    auto nov_full = combined_novelty_with_params(spectral_full, energy_full, hfc_full, c.tempogram_novelty_w_spectral,
                                                 c.tempogram_novelty_w_energy, c.tempogram_novelty_w_hfc,
                                                 c.tempogram_novelty_local_mean_window, c.tempogram_novelty_smooth_window);
    if (nov_full.empty()) return Error{PROCESSING_ERROR, "Novelty curve is empty after extraction"};
    dump_curve(dump, T + "nov.full.superflux", spectral_full);
    dump_curve(dump, T + "nov.full.energy", energy_full);
    dump_curve(dump, T + "nov.full.hfc", hfc_full);
    dump_curve(dump, T + "nov.full", nov_full);

    std::vector<Variant> seeds;
    {
        Variant v;
        v.name = "full";
        v.w = c.tempogram_band_w_full;
        if (Error e = fft_tempogram(nov_full, sr, hop, min_bpm, max_bpm, v.fft)) return e;
        if (Error e = autocorrelation_tempogram(nov_full, sr, hop, min_bpm, max_bpm, res_bpm, v.ac)) return e;
        v.max_fft = fmax_rs(v.fft.empty() ? 1.0f : v.fft[0].second, 1e-12f);
        v.max_ac = fmax_rs(v.ac.empty() ? 1.0f : v.ac[0].second, 1e-12f);
        seeds.push_back(std::move(v));
    }
    // primaries (find_best_bpm_fft / _autocorr: tempogram_fft.rs:206-236)
    float fft_primary = seeds[0].fft.empty() ? 0.0f : seeds[0].fft[0].first;
    float ac_primary = seeds[0].ac.empty() ? 0.0f : seeds[0].ac[0].first;

    if (c.enable_tempogram_band_fusion) {  // tempogram.rs:356-429
        size_t b0 = std::min<size_t>(1, n_bins - 1);
        size_t b_low = std::max(hz_to_bin(c.tempogram_band_low_max_hz, freq_res, n_bins), b0);
        size_t b_mid = std::max(hz_to_bin(c.tempogram_band_mid_max_hz, freq_res, n_bins), b_low + 1);
        size_t b_hi = c.tempogram_band_high_max_hz > 0.0f ? std::max(hz_to_bin(c.tempogram_band_high_max_hz, freq_res, n_bins), b_mid + 1) : n_bins;
        b_hi = std::min(b_hi, n_bins);
        struct B {
            const char* name;
            size_t s, e;
            float w;
        } bands[3] = {{"low", b0, b_low, c.tempogram_band_w_low}, {"mid", b_low, b_mid, c.tempogram_band_w_mid}, {"high", b_mid, b_hi, c.tempogram_band_w_high}};
        for (auto& b : bands) {
            if (!(std::isfinite(b.w) && b.w > 0.0f)) continue;
            if (b.e <= b.s + 1) continue;
            auto sp = superflux_novelty_band(S, sf_k, b.s, b.e);
            auto en = energy_flux_novelty_band(S, b.s, b.e);
            auto hf = hfc_novelty_band(S, b.s, b.e);
            auto nov = combined_novelty_with_params(sp, en, hf, c.tempogram_novelty_w_spectral, c.tempogram_novelty_w_energy,
                                                    c.tempogram_novelty_w_hfc, c.tempogram_novelty_local_mean_window,
                                                    c.tempogram_novelty_smooth_window);
            if (nov.empty()) continue;
            dump_curve(dump, T + "nov." + b.name, nov);
            Variant v;
            v.name = b.name;
            v.w = b.w;
            if (Error e = fft_tempogram(nov, sr, hop, min_bpm, max_bpm, v.fft)) return e;
            if (Error e = autocorrelation_tempogram(nov, sr, hop, min_bpm, max_bpm, res_bpm, v.ac)) return e;
            v.max_fft = fmax_rs(v.fft.empty() ? 1.0f : v.fft[0].second, 1e-12f);
            v.max_ac = fmax_rs(v.ac.empty() ? 1.0f : v.ac[0].second, 1e-12f);
            seeds.push_back(std::move(v));
        }
    }
    if (c.enable_tempogram_mel_novelty) {  // :432-462
        std::vector<float> mel;
        if (Error e = mel_superflux_novelty(S, sr, c.tempogram_mel_n_mels, c.tempogram_mel_fmin_hz, c.tempogram_mel_fmax_hz,
                                            c.tempogram_mel_max_filter_bins, mel))
            return e;
        if (!mel.empty()) {
            dump_curve(dump, T + "nov.mel", mel);
            Variant v;
            v.name = "mel";
            v.w = c.tempogram_mel_weight;
            if (Error e = fft_tempogram(mel, sr, hop, min_bpm, max_bpm, v.fft)) return e;
            if (Error e = autocorrelation_tempogram(mel, sr, hop, min_bpm, max_bpm, res_bpm, v.ac)) return e;
            v.max_fft = fmax_rs(v.fft.empty() ? 1.0f : v.fft[0].second, 1e-12f);
            v.max_ac = fmax_rs(v.ac.empty() ? 1.0f : v.ac[0].second, 1e-12f);
            seeds.push_back(std::move(v));
        }
    }
    if (dump)
        for (auto& v : seeds) {
            dump_tempogram(dump, T + "tg.fft." + v.name, v.fft);
            dump_tempogram(dump, T + "tg.ac." + v.name, v.ac);
        }

    const bool seed_only = c.tempogram_band_seed_only;
    std::vector<const Variant*> score_v;
    for (auto& v : seeds)
        if (!seed_only || std::string(v.name) == "full") score_v.push_back(&v);
    float support_thr = clamp_rs(c.tempogram_band_support_threshold, 0.0f, 1.0f);
    float bonus = fmax_rs(c.tempogram_band_consensus_bonus, 0.0f);
    float w_sum = 0.0f;
    for (auto* v : score_v) w_sum += fmax_rs(v->w, 0.0f);
    w_sum = fmax_rs(w_sum, 1e-6f);

    bool all_empty = true;
    for (auto& v : seeds)
        if (!(v.fft.empty() && v.ac.empty())) all_empty = false;
    if (all_empty) return Error{PROCESSING_ERROR, "Both FFT and autocorrelation tempograms are empty"};

    std::vector<float> seed_bpms;  // :538-548
    for (auto& v : seeds) {
        for (size_t i = 0; i < std::min<size_t>(8, v.fft.size()); ++i) seed_bpms.push_back(v.fft[i].first);
        for (size_t i = 0; i < std::min<size_t>(8, v.ac.size()); ++i) seed_bpms.push_back(v.ac[i].first);
    }
    if (fft_primary > 0.0f) seed_bpms.push_back(fft_primary);
    if (ac_primary > 0.0f) seed_bpms.push_back(ac_primary);
    const float FACTORS[7] = {1.0f, 0.5f, 2.0f, 1.0f / 3.0f, 3.0f, 2.0f / 3.0f, 3.0f / 2.0f};
    std::vector<float> candidates;
    for (float base : seed_bpms)
        for (float f : FACTORS) {
            float b = base * f;
            if (std::isfinite(b) && b >= min_bpm && b <= max_bpm) candidates.push_back(b);
        }
    std::stable_sort(candidates.begin(), candidates.end());
    std::vector<float> uniq;
    for (float b : candidates) {
        if (!uniq.empty() && fabsf(b - uniq.back()) < 0.75f) continue;
        uniq.push_back(b);
    }

    struct Sc {
        float bpm, score, fft_norm, ac_norm;
    };
    std::vector<Sc> scored;
    const float ac_tol = fmax_rs(res_bpm, 0.5f);
    for (float bpm : uniq) {
        float fft_acc = 0.0f, ac_acc = 0.0f;
        for (auto* v : score_v) {
            if (v->w <= 0.0f) continue;
            float fv = lookup_nearest(v->fft, bpm, 0.75f);
            float av = lookup_nearest(v->ac, bpm, ac_tol);
            fft_acc += v->w * clamp_rs(fv / v->max_fft, 0.0f, 1.0f);
            ac_acc += v->w * clamp_rs(av / v->max_ac, 0.0f, 1.0f);
        }
        float fft_norm = clamp_rs(fft_acc / w_sum, 0.0f, 1.0f);
        float ac_norm = clamp_rs(ac_acc / w_sum, 0.0f, 1.0f);
        float score = 0.55f * ac_norm + 0.45f * fft_norm;
        if (bonus > 0.0f && (c.enable_tempogram_band_fusion || c.enable_tempogram_mel_novelty)) {
            uint32_t support = 0;
            for (auto& v : seeds) {
                if (std::string(v.name) == "full") continue;
                float sf = clamp_rs(lookup_nearest(v.fft, bpm, 0.75f) / v.max_fft, 0.0f, 1.0f);
                float sa = clamp_rs(lookup_nearest(v.ac, bpm, ac_tol) / v.max_ac, 0.0f, 1.0f);
                if (fmax_rs(sf, sa) >= support_thr) ++support;
            }
            if (support >= 2) score *= 1.0f + bonus * ((float)support - 1.0f);
        }
        if (bpm > 180.0f)
            score *= 0.80f;
        else if (bpm < 60.0f)
            score *= 0.90f;
        scored.push_back(Sc{bpm, score, fft_norm, ac_norm});
    }
    std::stable_sort(scored.begin(), scored.end(), [](const Sc& a, const Sc& b) { return a.score > b.score; });
    if (scored.empty()) return Error{PROCESSING_ERROR, "No BPM candidates could be scored"};
    Sc best = scored[0];
    if (best.bpm > 180.0f) {  // :669-699
        float folded = best.bpm / 2.0f;
        if (folded >= min_bpm && folded <= max_bpm) {
            for (auto& s : scored)
                if (fabsf(s.bpm - folded) < 0.75f) {
                    const float eps = 1e-6f;
                    float ar = (best.ac_norm + eps) / (s.ac_norm + eps);
                    float fr = (best.fft_norm + eps) / (s.fft_norm + eps);
                    if (!(ar > 2.0f && fr > 2.0f)) best = s;
                    break;
                }
        }
    }
    float conf = 0.0f;
    if (best.score > 1e-12f) {
        float second = scored.size() > 1 ? scored[1].score : 0.0f;
        conf = clamp_rs(fmax_rs(best.score - second, 0.0f) / best.score, 0.0f, 1.0f);
    }
    uint32_t agree = 0;
    if (fft_primary > 0.0f && fabsf(fft_primary - best.bpm) < 2.0f) ++agree;
    if (ac_primary > 0.0f && fabsf(ac_primary - best.bpm) < 2.0f) ++agree;
    est.bpm = best.bpm;
    est.confidence = conf;
    est.method_agreement = agree;
    for (auto& s : scored) cands.push_back(TempoCand{s.bpm, s.score, s.fft_norm, s.ac_norm, fabsf(s.bpm - best.bpm) < 0.75f});
    if (dump) {
        std::vector<float> b, sc;
        for (auto& s : scored) {
            b.push_back(s.bpm);
            sc.push_back(s.score);
        }
        dump->f[T + "cands.bpm"] = b;
        dump->f[T + "cands.score"] = sc;
    }
    if (top_n == 0)
        cands.clear();
    else if (cands.size() > top_n)
        cands.resize(top_n);
    return Error{};
}

// ---- multi-resolution escalation — multi_resolution.rs:205-901 -----------------------------
static float cand_lookup(const std::vector<TempoCand>& c, float bpm, float tol) {  // :282-293
    float best_d = INFINITY, best_s = 0.0f;
    for (auto& x : c) {
        float d = fabsf(x.bpm - bpm);
        if (d <= tol && d < best_d) {
            best_d = d;
            best_s = x.score;
        }
    }
    return best_s;
}

static float beat_contrast_score(const std::vector<float>& nov, uint32_t sr, uint32_t hop, float bpm) {  // :580-678
    if (nov.size() < 16 || !(std::isfinite(bpm) && bpm > 0.0f) || sr == 0 || hop == 0) return 0.0f;
    float fpb = (60.0f * (float)sr) / (bpm * (float)hop);
    if (!std::isfinite(fpb) || fpb < 3.0f) return 0.0f;
    long per = as_isize(roundf(fpb));
    if (per < 3 || per > 512) return 0.0f;
    const size_t period = (size_t)per, w = 2, n = nov.size();
    float total = 0.0f;
    for (float v : nov) total += v;
    total = fmax_rs(total, 1e-6f);
    auto win_max = [&](size_t c) {
        size_t st = c >= w ? c - w : 0, en = std::min(c + w + 1, n);
        float mx = 0.0f;
        for (size_t j = st; j < en; ++j) mx = fmax_rs(mx, nov[j]);
        return mx;
    };
    float best = -1e9f;
    for (size_t phase = 0; phase < period; ++phase) {
        float bs = 0, hs = 0, ts = 0;
        uint32_t bn = 0, hn = 0, tn = 0;
        for (size_t i = phase; i < n; i += period) {
            bs += win_max(i);
            ++bn;
            if (period >= 6) {
                size_t j = i + period / 2;
                if (j < n) {
                    hs += win_max(j);
                    ++hn;
                }
            }
            if (period >= 9)
                for (size_t frac = 1; frac <= 2; ++frac) {
                    size_t j = i + (period * frac) / 3;
                    if (j < n) {
                        ts += win_max(j);
                        ++tn;
                    }
                }
        }
        float bm = bn ? bs / (float)bn : 0.0f, hm = hn ? hs / (float)hn : 0.0f, tm = tn ? ts / (float)tn : 0.0f;
        float contrast = bm - 0.60f * hm - 0.40f * tm;
        float score = clamp_rs(contrast / fmax_rs(total / (float)n, 1e-6f), -10.0f, 10.0f);
        best = fmax_rs(best, score);
    }
    return best;
}

Error multi_resolution_tempogram(const float* s, size_t n, uint32_t sr, const Config& c, const Spec* S512_in, BpmEstimate& est,
                                 std::vector<TempoCand>& c512, Dump* dump) {
    const size_t frame_size = c.frame_size;
    if (n < frame_size) return Error{INVALID_INPUT, "Audio too short for STFT"};
    const float min_bpm = c.min_bpm, max_bpm = c.max_bpm;
    const size_t top_k = std::max<size_t>(c.tempogram_multi_res_top_k, 1);
    const size_t aux_k = std::min<size_t>(std::max<size_t>(top_k * 4, 25), 200);
    const float tol = fmax_rs(2.0f, c.bpm_resolution);
    const float w512 = c.tempogram_multi_res_w512, w256 = c.tempogram_multi_res_w256, w1024 = c.tempogram_multi_res_w1024;
    const float dt = c.tempogram_multi_res_double_time_512_factor, margin_thr = c.tempogram_multi_res_margin_threshold;

    Spec h256 = compute_stft(s, n, frame_size, 256);
    Spec h512_local;
    if (!S512_in) h512_local = compute_stft(s, n, frame_size, 512);
    const Spec& h512 = S512_in ? *S512_in : h512_local;
    Spec h1024 = compute_stft(s, n, frame_size, 1024);

    std::vector<TempoCand> c256, c1024;
    BpmEstimate e_;
    if (Error e = estimate_bpm_tempogram(h256, sr, 256, c, aux_k, e_, c256, dump, "h256.")) return e;
    if (Error e = estimate_bpm_tempogram(h512, sr, 512, c, top_k, e_, c512, nullptr, "")) return e;
    if (Error e = estimate_bpm_tempogram(h1024, sr, 1024, c, aux_k, e_, c1024, dump, "h1024.")) return e;

    struct Hyp {
        float bpm, score;
    };
    std::vector<Hyp> hyps;
    for (size_t ti = 0; ti < std::min(top_k, c512.size()); ++ti) {  // :407-523
        float t = c512[ti].bpm;
        if (!(std::isfinite(t) && t > 0.0f)) continue;
        float s_t_512 = cand_lookup(c512, t, tol), s_t_256 = cand_lookup(c256, t, tol), s_t_1024 = cand_lookup(c1024, t, tol);
        float s_2_512 = cand_lookup(c512, t * 2.0f, tol), s_2_256 = cand_lookup(c256, t * 2.0f, tol), s_2_1024 = cand_lookup(c1024, t * 2.0f, tol);
        float s_h_512 = cand_lookup(c512, t * 0.5f, tol), s_h_256 = cand_lookup(c256, t * 0.5f, tol), s_h_1024 = cand_lookup(c1024, t * 0.5f, tol);
        float h_t = w512 * s_t_512 + w256 * s_t_256 + w1024 * s_t_1024;
        float h_2t = w512 * (dt * s_t_512 + (1.0f - dt) * s_2_512) + w256 * s_2_256 + w1024 * s_2_1024;
        float h_half = w512 * (dt * s_t_512 + (1.0f - dt) * s_h_512) + w256 * s_h_256 + w1024 * s_h_1024;
        if (s_t_1024 > s_h_1024 * 1.02f) h_half *= 0.90f;
        if (s_t_1024 > s_2_1024 * 1.02f) h_2t *= 0.90f;
        const float eps = 1e-6f;
        float r2 = (s_2_256 + eps) / (s_t_256 + eps);
        if (r2 < 1.10f) h_2t *= 0.75f;
        if (r2 < 1.00f) h_2t *= 0.75f;
        float rh = (s_h_1024 + eps) / (s_t_1024 + eps);
        if (rh < 1.10f) h_half *= 0.75f;
        if (rh < 1.00f) h_half *= 0.75f;
        std::vector<std::pair<float, float>> local = {{t, h_t}, {t * 2.0f, h_2t}, {t * 0.5f, h_half}};
        local.erase(std::remove_if(local.begin(), local.end(), [&](const std::pair<float, float>& p) { return !(p.first >= min_bpm && p.first <= max_bpm); }),
                    local.end());
        for (auto& p : local) {
            if (p.first > 210.0f)
                p.second *= 0.80f;
            else if (p.first > 180.0f)
                p.second *= 0.90f;
            else if (p.first < 60.0f)
                p.second *= 0.92f;
        }
        std::stable_sort(local.begin(), local.end(), [](const std::pair<float, float>& a, const std::pair<float, float>& b) { return a.second > b.second; });
        if (local.empty()) continue;
        float best_bpm = local[0].first, best_score = local[0].second;
        float second = local.size() > 1 ? local[1].second : 0.0f;
        float margin = best_score - second;
        float ch_bpm = best_bpm, ch_score = best_score;
        if (fabsf(ch_bpm - t) > 1e-3f && margin < margin_thr) {
            ch_bpm = t;
            ch_score = h_t;
        }
        if (margin < margin_thr && c.tempogram_multi_res_use_human_prior && ch_bpm >= 70.0f && ch_bpm <= 180.0f && margin < 0.05f) ch_score += 0.05f;
        hyps.push_back(Hyp{ch_bpm, ch_score});
    }
    if (hyps.empty()) return Error{PROCESSING_ERROR, "Multi-resolution fusion produced no hypotheses"};
    std::stable_sort(hyps.begin(), hyps.end(), [](const Hyp& a, const Hyp& b) { return a.score > b.score; });
    std::vector<Hyp> uniq;
    for (auto& h : hyps) {
        bool dup = false;
        for (auto& u : uniq)
            if (fabsf(u.bpm - h.bpm) < 0.75f) dup = true;
        if (dup) continue;
        uniq.push_back(h);
        if (uniq.size() >= 8) break;
    }
    Hyp best = uniq[0];
    auto total_support = [&](float bpm, uint32_t* agree) {
        float a = cand_lookup(c256, bpm, tol), b = cand_lookup(c512, bpm, tol), d = cand_lookup(c1024, bpm, tol);
        *agree = (a > 0.0f) + (b > 0.0f) + (d > 0.0f);
        return a + b + d;
    };
    // novelty_512 recomputed from the hop-512 spectrogram (:681-695)
    std::vector<float> nov512 = combined_novelty_with_params(
        superflux_novelty(h512, c.tempogram_superflux_max_filter_bins), energy_flux_novelty(h512), hfc_novelty(h512), c.tempogram_novelty_w_spectral,
        c.tempogram_novelty_w_energy, c.tempogram_novelty_w_hfc, c.tempogram_novelty_local_mean_window, c.tempogram_novelty_smooth_window);

    if (best.bpm >= 170.0f) {  // fold-down :698-724
        float half = best.bpm * 0.5f;
        if (half >= 70.0f && half <= 120.0f) {
            uint32_t ab, ah;
            float sb = total_support(best.bpm, &ab), sh = total_support(half, &ah);
            float ratio = sb > 0.0f ? sh / sb : 0.0f;
            if (ah >= 3 && sh > 0.0f && sb > 0.0f && ratio >= 0.45f) best = Hyp{half, sh};
        }
    }
    if (best.bpm <= 80.0f) {  // fold-up :727-751
        float dbl = best.bpm * 2.0f;
        if (dbl >= 70.0f && dbl <= 180.0f) {
            uint32_t ab, ad;
            float sb = total_support(best.bpm, &ab), sd = total_support(dbl, &ad);
            float ratio = sb > 0.0f ? sd / sb : 0.0f;
            if (ad >= 2 && sd > 0.0f && sb > 0.0f && ratio >= 0.55f) best = Hyp{dbl, sd};
        }
    }
    if (best.bpm >= 70.0f && best.bpm <= 180.0f && !nov512.empty()) {  // triplet family :764-867
        const float family[5] = {1.0f, 3.0f / 2.0f, 2.0f / 3.0f, 4.0f / 3.0f, 3.0f / 4.0f};
        struct Fam {
            float bpm, support, align;
        };
        std::vector<Fam> fams;
        for (float f : family) {
            float bpm = best.bpm * f;
            if (!(std::isfinite(bpm) && bpm >= min_bpm && bpm <= max_bpm)) continue;
            if (!(bpm >= 70.0f && bpm <= 180.0f)) continue;
            uint32_t ag;
            float sup = total_support(bpm, &ag);
            if (ag < 2 || sup <= 0.0f) continue;
            fams.push_back(Fam{bpm, sup, beat_contrast_score(nov512, sr, 512, bpm)});
        }
        if (fams.size() >= 2) {
            float best_support = 0.0f;
            for (auto& f : fams) best_support = fmax_rs(best_support, f.support);
            best_support = fmax_rs(best_support, 1e-6f);
            float max_alt = 0.0f;
            for (auto& f : fams)
                if (fabsf(f.bpm - best.bpm) > 0.75f) max_alt = fmax_rs(max_alt, f.support / best_support);
            if (max_alt >= 0.45f) {
                Fam chosen = fams[0];
                float chosen_score = -1e9f;
                for (auto& f : fams) {
                    float sn = clamp_rs(f.support / best_support, 0.0f, 1.0f);
                    float sc = f.align + 0.35f * sn;
                    if (sc > chosen_score) {
                        chosen = f;
                        chosen_score = sc;
                    }
                }
                float cur_align = beat_contrast_score(nov512, sr, 512, best.bpm);
                if (fabsf(chosen.bpm - best.bpm) > 0.75f && chosen.align >= cur_align + 0.40f) best = Hyp{chosen.bpm, chosen.support};
            }
        }
    }
    float second = uniq.size() > 1 ? uniq[1].score : 0.0f;
    float conf = best.score > 1e-6f ? clamp_rs(fmax_rs(best.score - second, 0.0f) / best.score, 0.0f, 1.0f) : 0.0f;
    uint32_t agree = 0;
    if (cand_lookup(c256, best.bpm, tol) > 0.0f) ++agree;
    if (cand_lookup(c512, best.bpm, tol) > 0.0f) ++agree;
    if (cand_lookup(c1024, best.bpm, tol) > 0.0f) ++agree;
    for (auto& x : c512) x.selected = fabsf(x.bpm - best.bpm) < 0.75f;
    est.bpm = best.bpm;
    est.confidence = conf;
    est.method_agreement = agree;
    if (dump) {
        std::vector<float> hb, hs;
        for (auto& u : uniq) {
            hb.push_back(u.bpm);
            hs.push_back(u.score);
        }
        dump->f["mr.hyp.bpm"] = hb;
        dump->f["mr.hyp.score"] = hs;
        dump->f["mr.est"] = {est.bpm, est.confidence, (float)est.method_agreement};
    }
    return Error{};
}

}  // namespace so
