// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates analyze_audio (reference src/lib.rs:86-1635, default-config branches) and
// compute_confidence (src/analysis/confidence.rs:121-297).
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "so_common.hpp"

namespace so {

static float cand_support(const std::vector<TempoCand>& c, float bpm, float tol) {  // lib.rs:420-432
    float best = 0.0f;
    for (auto& x : c)
        if (fabsf(x.bpm - bpm) <= tol) best = fmax_rs(best, x.score);
    return best;
}

Error analyze_audio(const float* samples, size_t n, uint32_t sr, const Config& c, Result& r, Dump* dump) {
    r = Result{};
    if (n == 0) return Error{INVALID_INPUT, "Empty audio samples"};
    if (sr == 0) return Error{INVALID_INPUT, "Invalid sample rate"};
    std::vector<float> proc(samples, samples + n);
    float gain = 1.0f;
    if (c.enable_normalization)
        if (Error e = normalize(proc, c.normalization, -14.0f, 1.0f, (float)sr, &gain)) return e;
    size_t ts = 0, te = n;
    if (c.enable_silence_trimming)
        if (Error e = detect_and_trim(proc, sr, c.min_amplitude_db, 500, c.frame_size, &ts, &te, nullptr)) return e;
    r.trim_start = ts;
    r.trim_end = te;
    if (te <= ts) return Error{PROCESSING_ERROR, "Audio is entirely silent after trimming"};
    const float* x = proc.data() + ts;
    const size_t m = te - ts;
    if (dump) dump->f["gain"] = {gain};

    std::vector<size_t> energy_onsets;
    if (Error e = detect_energy_flux_onsets(x, m, c.frame_size, c.hop_size, -20.0f, energy_onsets)) return e;
    Spec S = compute_stft(x, m, c.frame_size, c.hop_size);

    std::vector<size_t> onsets_legacy = energy_onsets, onsets_bt = energy_onsets;
    if (c.enable_onset_consensus && !S.empty()) {  // lib.rs:176-291
        auto to_samples = [&](const std::vector<size_t>& fr) {
            std::vector<size_t> o;
            for (size_t f : fr) {
                size_t smp = f * c.hop_size;
                if (smp < m) o.push_back(smp);
            }
            std::sort(o.begin(), o.end());
            o.erase(std::unique(o.begin(), o.end()), o.end());
            return o;
        };
        std::vector<size_t> sf, hf;
        std::vector<float> sflux, hflux;
        if (detect_spectral_flux_onsets(S, c.onset_threshold_percentile, sf, dump ? &sflux : nullptr)) sf.clear();
        if (detect_hfc_onsets(S, sr, c.onset_threshold_percentile, hf, dump ? &hflux : nullptr)) hf.clear();
        std::vector<size_t> hp;  // lib.rs:222-235: HPSS onsets as the fourth detector
        std::vector<float> pflux;
        if (c.enable_hpss_onsets) {
            Spec Hh, Pp;
            if (hpss_decompose(S, c.hpss_margin, Hh, Pp) || detect_hpss_onsets(Pp, c.onset_threshold_percentile, hp, dump ? &pflux : nullptr)) hp.clear();
            if (dump) {
                dump->f["hpss.flux"] = pflux;
                const size_t head = std::min<size_t>(Pp.frames, 64) * Pp.bins;
                dump->f["hpss.perc_head"] = std::vector<float>(Pp.d.begin(), Pp.d.begin() + head);
            }
        }
        std::vector<size_t> lists[4] = {energy_onsets, to_samples(sf), to_samples(hf), to_samples(hp)};
        if (dump) {
            dump->f["onset.spectral_flux"] = sflux;
            dump->f["onset.hfc_flux"] = hflux;
            dump->i["onset.energy"] = std::vector<int64_t>(energy_onsets.begin(), energy_onsets.end());
            dump->i["onset.spectral"] = std::vector<int64_t>(lists[1].begin(), lists[1].end());
            dump->i["onset.hfc"] = std::vector<int64_t>(lists[2].begin(), lists[2].end());
            dump->i["onset.hpss"] = std::vector<int64_t>(lists[3].begin(), lists[3].end());
        }
        std::vector<OnsetCand> cands;
        if (!vote_onsets(lists, c.onset_consensus_weights, c.onset_consensus_tolerance_ms, sr, cands)) {
            std::vector<size_t> strong, any;
            for (auto& cd : cands) {
                if (cd.voted_by >= 2) strong.push_back(cd.time_samples);
                any.push_back(cd.time_samples);
            }
            std::sort(strong.begin(), strong.end());
            strong.erase(std::unique(strong.begin(), strong.end()), strong.end());
            std::sort(any.begin(), any.end());
            any.erase(std::unique(any.begin(), any.end()), any.end());
            const std::vector<size_t>& chosen = !strong.empty() ? strong : any;
            if (!chosen.empty()) {
                onsets_legacy = chosen;
                onsets_bt = chosen;
            }
        }
    }
    r.onsets.assign(onsets_bt.begin(), onsets_bt.end());

    // legacy estimate (always computed; errors propagate) — lib.rs:294-329
    bool has_legacy = false;
    BpmEstimate legacy;
    if (onsets_legacy.size() >= 2)
        if (Error e = estimate_bpm_legacy(onsets_legacy, sr, c.hop_size, c, c.enable_legacy_bpm_guardrails, &has_legacy, legacy, dump)) return e;
    if (dump) dump->f["legacy.est"] = {has_legacy ? legacy.bpm : 0.0f, has_legacy ? legacy.confidence : 0.0f};

    bool has_tempogram = false;
    BpmEstimate tg;
    if (!c.force_legacy_bpm && !S.empty()) {
        size_t base_top_n = std::max(std::max(c.tempogram_candidates_top_n, c.tempogram_multi_res_top_k), (size_t)10);
        BpmEstimate base;
        std::vector<TempoCand> base_c;
        // lib.rs:378-410 / 714-737: with multi-resolution the base call keeps base_top_n candidates; without it candidates are
        // only kept (top tempogram_candidates_top_n) when emit_tempogram_candidates is set
        const size_t plain_top_n = c.emit_tempogram_candidates ? c.tempogram_candidates_top_n : 0;
        Error te_ = c.enable_tempogram_multi_resolution ? estimate_bpm_tempogram(S, sr, (uint32_t)c.hop_size, c, base_top_n, base, base_c, dump, "base.")
                                                        : estimate_bpm_tempogram(S, sr, (uint32_t)c.hop_size, c, plain_top_n, base, base_c, dump, "base.");
        std::vector<TempoCand> chosen_c = base_c;
        if (!te_) {
            has_tempogram = true;
            tg = base;
            if (dump) dump->f["base.est"] = {base.bpm, base.confidence, (float)base.method_agreement};
            if (c.enable_tempogram_multi_resolution) {  // lib.rs:412-580
                bool trap_low = base.bpm >= 55.0f && base.bpm <= 80.0f;
                bool trap_high = base.bpm >= 170.0f && base.bpm <= 200.0f;
                float tol = fmax_rs(2.0f, c.bpm_resolution);
                float s_base = cand_support(base_c, base.bpm, tol);
                float s_2x = cand_support(base_c, base.bpm * 2.0f, tol);
                float s_half = cand_support(base_c, base.bpm * 0.5f, tol);
                bool family = (s_2x > 0.0f && s_2x >= s_base * 0.90f) || (s_half > 0.0f && s_half >= s_base * 0.90f);
                bool fold_into_trap = base.bpm * 2.0f >= 170.0f && base.bpm * 2.0f <= 200.0f;
                bool weak = base.method_agreement == 0 || base.confidence < 0.06f;
                bool ambiguous = trap_low || trap_high || family || (weak && fold_into_trap);
                r.multi_res_triggered = ambiguous ? 1 : 0;
                bool used = false;
                if (ambiguous) {
                    BpmEstimate mr;
                    std::vector<TempoCand> mc;
                    // multi_resolution_tempogram_from_samples (lib.rs:493-509) recomputes the STFT at hops 256 / 512 / 1024 from the samples; the base
                    // spectrogram is the hop-512 one only when hop_size is 512
                    if (!multi_resolution_tempogram(x, m, sr, c, c.hop_size == 512 ? &S : nullptr, mr, mc, dump)) {
                        float rel = base.bpm > 1e-6f ? fmax_rs(mr.bpm / base.bpm, base.bpm / mr.bpm) : 1.0f;
                        bool fam_rel = fabsf(rel - 2.0f) < 0.05f || fabsf(rel - 1.5f) < 0.05f || fabsf(rel - (4.0f / 3.0f)) < 0.05f;
                        bool forbid = base.bpm <= 180.0f && mr.bpm > 180.0f;
                        bool better = !forbid && (mr.confidence >= (base.confidence + 0.05f) ||
                                                  (mr.method_agreement > base.method_agreement && mr.confidence >= base.confidence * 0.90f) ||
                                                  ((trap_low || trap_high) && fam_rel && mr.confidence >= base.confidence * 0.88f &&
                                                   ((mr.bpm >= 70.0f && mr.bpm <= 180.0f) || base.bpm > 180.0f)));
                        if (better) {
                            tg = mr;
                            chosen_c = mc;
                            used = true;
                        }
                    }
                }
                r.multi_res_used = used ? 1 : 0;
                const bool percussive_needed = ambiguous && trap_low;  // lib.rs:587-588
                r.percussive_triggered = percussive_needed ? 1 : 0;
                if (c.enable_tempogram_percussive_fallback && percussive_needed) {  // lib.rs:590-683
                    r.percussive_used = 0;
                    Spec Hh, Pp;
                    BpmEstimate pe;
                    std::vector<TempoCand> pc;
                    if (!hpss_decompose(S, c.hpss_margin, Hh, Pp) && !estimate_bpm_tempogram(Pp, sr, (uint32_t)c.hop_size, c, base_top_n, pe, pc, dump, "perc.")) {
                        if (dump) dump->f["perc.est"] = {pe.bpm, pe.confidence, (float)pe.method_agreement};
                        const float rel = tg.bpm > 1e-6f ? fmax_rs(pe.bpm / tg.bpm, tg.bpm / pe.bpm) : 1.0f;
                        const bool family_related = fabsf(rel - 2.0f) < 0.05f || fabsf(rel - 1.5f) < 0.05f || fabsf(rel - (4.0f / 3.0f)) < 0.05f ||
                                                    fabsf(rel - (3.0f / 2.0f)) < 0.05f || fabsf(rel - (2.0f / 3.0f)) < 0.05f || fabsf(rel - (3.0f / 4.0f)) < 0.05f;
                        const bool forbid_promote_high = tg.bpm <= 180.0f && pe.bpm > 180.0f;
                        const bool base_low_trap = trap_low || base.bpm < 95.0f;
                        const bool percussive_in_common = pe.bpm >= 70.0f && pe.bpm <= 180.0f;
                        const bool p_better = !forbid_promote_high && family_related && percussive_in_common &&
                                              (pe.confidence >= tg.confidence + 0.04f || (base_low_trap && pe.confidence >= tg.confidence * 0.85f) ||
                                               (pe.method_agreement > tg.method_agreement && pe.confidence >= tg.confidence * 0.92f));
                        if (p_better) {
                            tg = pe;
                            chosen_c = pc;
                            r.percussive_used = 1;
                        }
                    }
                } else if (c.enable_tempogram_percussive_fallback) {
                    r.percussive_used = 0;
                }
            }
            if (c.emit_tempogram_candidates) {  // lib.rs:684-697, 740-752
                r.has_candidates = true;
                r.tempogram_candidates = chosen_c;
            }
        }
    }

    float bpm = 0.0f, bpm_conf = 0.0f;
    if (c.force_legacy_bpm) {
        if (has_legacy) {
            bpm = legacy.bpm;
            bpm_conf = legacy.confidence;
        }
    } else if (c.enable_bpm_fusion) {  // lib.rs:819-892: the legacy estimate only validates the tempogram's confidence
        const float t_bpm = has_tempogram ? tg.bpm : 0.0f, t_conf = has_tempogram ? tg.confidence : 0.0f;
        const float l_bpm = has_legacy ? legacy.bpm : 0.0f, l_conf = clamp_rs(has_legacy ? legacy.confidence : 0.0f, 0.0f, 1.0f);
        if (t_bpm <= 0.0f) {
            if (has_legacy) {
                bpm = legacy.bpm;
                bpm_conf = legacy.confidence;
            }
        } else {
            float conf = clamp_rs(t_conf, 0.0f, 1.0f);
            bool agreement = false;
            if (l_bpm > 0.0f) {
                const float diffs[5] = {fabsf(l_bpm - t_bpm), fabsf(l_bpm - (t_bpm * 0.5f)), fabsf(l_bpm - (t_bpm * 2.0f)),
                                        fabsf(l_bpm - (t_bpm * (2.0f / 3.0f))), fabsf(l_bpm - (t_bpm * (3.0f / 2.0f)))};
                for (float d : diffs) agreement |= d <= 2.0f;
            }
            if (agreement)
                conf = clamp_rs(conf + 0.12f * l_conf, 0.0f, 1.0f);
            else if (l_bpm > 0.0f)
                conf = clamp_rs(conf * 0.90f, 0.0f, 1.0f);
            bpm = t_bpm;
            bpm_conf = conf;
        }
    } else if (has_tempogram) {
        bpm = tg.bpm;
        bpm_conf = tg.confidence;
    } else if (has_legacy) {
        bpm = legacy.bpm;
        bpm_conf = legacy.confidence;
    }
    r.bpm = bpm;
    r.bpm_confidence = bpm_conf;

    if (bpm > 0.0f && onsets_bt.size() >= 2) {  // lib.rs:913-958
        std::vector<float> os;
        for (size_t v : onsets_bt) os.push_back((float)v / (float)sr);
        Result tmp;
        if (!generate_beat_grid(bpm, bpm_conf, os, sr, tmp, dump)) {
            r.beats = tmp.beats;
            r.downbeats = tmp.downbeats;
            r.bars = tmp.bars;
            r.grid_stability = tmp.grid_stability;
            r.hmm_beat_frames = tmp.hmm_beat_frames;
            r.time_sig_beats_per_bar = tmp.time_sig_beats_per_bar;
            r.beats_refined = tmp.beats_refined;
        }
    }

    if (Error e = detect_key_path(x, m, sr, c, S, r, dump, &r.beats)) return e;

    // warnings / flags — lib.rs:1567-1589
    if (bpm == 0.0f) r.warnings |= WARN_BPM_FAILED;
    if (r.grid_stability < 0.5f) r.warnings |= WARN_LOW_GRID;
    if (r.key_confidence < 0.3f) r.warnings |= WARN_LOW_KEY_CONF;
    if (r.key_clarity < 0.2f) {
        r.warnings |= WARN_LOW_KEY_CLARITY;
        r.flags |= FLAG_WEAK_TONALITY;
    }
    r.duration_seconds = (float)m / (float)sr;
    r.sample_rate = sr;
    r.onset_method_consensus = energy_onsets.empty() ? 0.0f : 1.0f;
    return Error{};
}

std::vector<std::string> warning_strings(const Result& r) {  // exact strings of lib.rs:1567-1589
    std::vector<std::string> w;
    char buf[256];
    if (r.warnings & WARN_BPM_FAILED) w.push_back("BPM detection failed: insufficient onsets or estimation error");
    if (r.warnings & WARN_LOW_GRID) {
        snprintf(buf, sizeof buf, "Low beat grid stability: %.2f (may indicate tempo variation)", r.grid_stability);
        w.push_back(buf);
    }
    if (r.warnings & WARN_LOW_KEY_CONF) {
        snprintf(buf, sizeof buf, "Low key detection confidence: %.2f (may indicate ambiguous or atonal music)", r.key_confidence);
        w.push_back(buf);
    }
    if (r.warnings & WARN_LOW_KEY_CLARITY) {
        snprintf(buf, sizeof buf, "Low key clarity: %.2f (track may be atonal or have weak tonality)", r.key_clarity);
        w.push_back(buf);
    }
    return w;
}

// compute_confidence — confidence.rs:121-297
Confidence compute_confidence(const Result& r) {
    Confidence c;
    auto warns = warning_strings(r);
    auto any_contains = [&](const char* needle) {
        for (auto& s : warns)
            if (s.find(needle) != std::string::npos) return true;
        return false;
    };
    float bc = 0.0f;
    if (r.bpm > 0.0f) {
        bc = clamp_rs(r.bpm_confidence, 0.0f, 1.0f);
        if (any_contains("BPM")) bc = bc * 0.7f;
    }
    float kc = 0.0f;
    if (r.key_confidence > 0.0f) {
        float base = clamp_rs(r.key_confidence, 0.0f, 1.0f);
        float ca = r.key_clarity < 0.2f ? 0.6f : (r.key_clarity < 0.5f ? 0.85f : 1.0f);
        float wa = (any_contains("key") || any_contains("Key") || any_contains("tonality")) ? 0.7f : 1.0f;
        kc = base * ca * wa;
    }
    float gs = clamp_rs(r.grid_stability, 0.0f, 1.0f);
    float overall;
    if (bc > 0.0f && kc > 0.0f)
        overall = clamp_rs(bc * 0.4f + kc * 0.3f + gs * 0.3f, 0.0f, 1.0f);
    else if (bc > 0.0f)
        overall = bc * 0.6f;
    else if (kc > 0.0f)
        overall = kc * 0.6f;
    else
        overall = 0.0f;
    uint32_t flags = r.flags;
    if (bc < 0.3f) flags |= FLAG_MULTIMODAL_BPM;
    if (kc < 0.2f) flags |= FLAG_WEAK_TONALITY;
    if (gs < 0.3f) flags |= FLAG_TEMPO_VARIATION;
    c.bpm_confidence = bc;
    c.key_confidence = kc;
    c.grid_stability = gs;
    c.overall_confidence = overall;
    c.flags = flags;
    return c;
}

static const char* NOTE_NAMES[12] = {"C", "C#", "D", "D#", "E", "F", "F#", "G", "G#", "A", "A#", "B"};

std::string key_name(int is_minor, uint32_t idx) {  // result.rs:31-39
    std::string s = NOTE_NAMES[idx % 12];
    if (is_minor) s += "m";
    return s;
}

std::string key_numerical(int is_minor, uint32_t idx) {  // result.rs:60-87
    const int cof_major[12] = {0, 7, 2, 9, 4, 11, 6, 1, 8, 3, 10, 5};
    const int cof_minor[12] = {9, 4, 11, 6, 1, 8, 3, 10, 5, 0, 7, 2};
    const int* t = is_minor ? cof_minor : cof_major;
    int pos = 0;
    for (int i = 0; i < 12; ++i)
        if (t[i] == (int)(idx % 12)) {
            pos = i;
            break;
        }
    return std::to_string(pos + 1) + (is_minor ? "B" : "A");
}

}  // namespace so
