// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates reference src/features/beat_tracking/{hmm,tempo_variation,bayesian,
// time_signature,mod}.rs.
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;
static const int NUM_STATES = 5;
static const float TIMING_TOLERANCE_S = 0.05f;
static const float EMISSION_SIGMA = TIMING_TOLERANCE_S / 2.0f;

// nearest-onset distance, brute force as hmm.rs:273-279 (first strict minimum)
static float nearest_dist(const std::vector<float>& onsets, float t) {
    float md = INFINITY;
    for (float o : onsets) {
        float d = fabsf(o - t);
        if (d < md) md = d;
    }
    return md;
}

// HmmBeatTracker::track_beats — hmm.rs:121-160
Error hmm_track_beats(float bpm, const std::vector<float>& onsets, std::vector<BeatPos>& beats, std::vector<int>* path_out) {
    beats.clear();
    if (bpm <= EPSILON || bpm > 300.0f) return Error{INVALID_INPUT, "Invalid BPM estimate"};
    if (onsets.empty()) return Error{INVALID_INPUT, "Cannot track beats: no onsets provided"};
    // transition matrix — hmm.rs:184-219
    float A[NUM_STATES][NUM_STATES];
    for (int i = 0; i < NUM_STATES; ++i)
        for (int j = 0; j < NUM_STATES; ++j) {
            int d = std::abs(i - j);
            A[i][j] = d == 0 ? 0.7f : (d == 1 ? 0.15f : 0.0f);
        }
    for (int i = 0; i < NUM_STATES; ++i) {
        float sum = 0.0f;
        for (int j = 0; j < NUM_STATES; ++j) sum += A[i][j];
        if (sum > EPSILON)
            for (int j = 0; j < NUM_STATES; ++j) A[i][j] /= sum;
    }
    // emissions — hmm.rs:231-298 (identical for all 5 states: state_bpm is unused)
    float start = onsets.front(), end = onsets.back();
    float interval = 60.0f / bpm;
    size_t T = as_usize(ceilf((end - start) / interval)) + 1;
    if (T == 0) return Error{PROCESSING_ERROR, "Cannot compute emissions: invalid time range"};
    std::vector<float> em(T);
    const float sigma_sq = EMISSION_SIGMA * EMISSION_SIGMA;
    for (size_t t = 0; t < T; ++t) {
        float ft = start + ((float)t * interval);
        float md = nearest_dist(onsets, ft);
        em[t] = expf(-(md * md) / (2.0f * sigma_sq));
    }
    // Viterbi forward pass in the probability domain — hmm.rs:308-375
    std::vector<float> V(T * NUM_STATES, 0.0f);
    std::vector<uint8_t> bp(T * NUM_STATES, 0);
    const float init = 1.0f / (float)NUM_STATES;
    for (int s = 0; s < NUM_STATES; ++s) V[s] = init * em[0];
    for (size_t t = 1; t < T; ++t)
        for (int s = 0; s < NUM_STATES; ++s) {
            float bestp = 0.0f;
            int bprev = 0;
            for (int p = 0; p < NUM_STATES; ++p) {
                float pr = V[(t - 1) * NUM_STATES + p] * A[p][s];
                if (pr > bestp) {
                    bestp = pr;
                    bprev = p;
                }
            }
            V[t * NUM_STATES + s] = bestp * em[t];
            bp[t * NUM_STATES + s] = (uint8_t)bprev;
        }
    std::vector<int> path(T, 0);
    float bf = 0.0f;
    int bs = 0;
    for (int s = 0; s < NUM_STATES; ++s)
        if (V[(T - 1) * NUM_STATES + s] > bf) {
            bf = V[(T - 1) * NUM_STATES + s];
            bs = s;
        }
    path[T - 1] = bs;
    for (size_t t = T - 1; t-- > 0;) path[t] = bp[(t + 1) * NUM_STATES + path[t + 1]];
    if (path_out) *path_out = path;
    // extract_beats_from_path — hmm.rs:383-441
    for (size_t t = 0; t < T; ++t) {
        float e = em[t];  // emission_matrix[t][state] — same for every state
        if (e > 0.1f) {
            float bt = start + ((float)t * interval);
            float md = nearest_dist(onsets, bt);
            float align = md < TIMING_TOLERANCE_S ? 1.0f - (md / TIMING_TOLERANCE_S) : 0.0f;
            float conf = fmin_rs(e * 0.7f + align * 0.3f, 1.0f);
            beats.push_back(BeatPos{bt, conf, (int32_t)t});
        }
    }
    std::stable_sort(beats.begin(), beats.end(), [](const BeatPos& a, const BeatPos& b) { return a.time_seconds < b.time_seconds; });
    return Error{};
}

struct TempoSegment {
    float start, end, bpm, confidence;
    bool variable;
};

// detect_tempo_variations — tempo_variation.rs:95-220
static Error detect_tempo_variations(const std::vector<float>& beats, float nominal, std::vector<TempoSegment>& segs) {
    segs.clear();
    if (beats.size() < 4) {
        segs.push_back(TempoSegment{beats.empty() ? 0.0f : beats.front(), beats.empty() ? 0.0f : beats.back(), nominal, 0.5f, false});
        return Error{};
    }
    if (nominal <= EPSILON) return Error{INVALID_INPUT, "Invalid nominal BPM"};
    float total = beats.back() - beats.front();
    if (total < 2.0f) {
        segs.push_back(TempoSegment{beats.front(), beats.back(), nominal, 0.8f, false});
        return Error{};
    }
    float seg_dur = clamp_rs(total / 4.0f, 4.0f, 8.0f);
    float overlap = seg_dur * 0.5f;
    float cur = beats.front();
    while (cur < beats.back()) {
        float seg_end = fmin_rs(cur + seg_dur, beats.back());
        std::vector<float> sb;
        for (float b : beats)
            if (b >= cur && b <= seg_end) sb.push_back(b);
        if (sb.size() >= 3) {
            std::vector<float> iv;
            for (size_t i = 1; i < sb.size(); ++i) {
                float d = sb[i] - sb[i - 1];
                if (d > 0.0f) iv.push_back(d);
            }
            if (!iv.empty()) {
                float sum = 0.0f;
                for (float d : iv) sum += d;
                float mean = sum / (float)iv.size();
                float vs = 0.0f;
                for (float d : iv) {
                    float df = d - mean;
                    vs += df * df;
                }
                float var = vs / (float)iv.size();
                float sd = sqrtf(var);
                float cv = mean > EPSILON ? sd / mean : 0.0f;
                float sbpm = mean > EPSILON ? 60.0f / mean : nominal;
                float conf = fmax_rs(1.0f - fmin_rs(cv / 0.3f, 1.0f), 0.0f);
                segs.push_back(TempoSegment{cur, seg_end, sbpm, conf, cv > 0.15f});
            }
        }
        cur += seg_dur - overlap;
    }
    if (segs.empty()) segs.push_back(TempoSegment{beats.front(), beats.back(), nominal, 0.8f, false});
    return Error{};
}

// BayesianBeatTracker — bayesian.rs:53-255 (only current_bpm carries across updates)
struct Bayes {
    float current_bpm;
    Error update(const std::vector<float>& onsets, float* out_bpm) {
        if (onsets.empty()) return Error{INVALID_INPUT, "Cannot update: no onsets provided"};
        if (current_bpm <= EPSILON || current_bpm > 300.0f) return Error{INVALID_INPUT, "Invalid current BPM"};
        float lo = fmax_rs(current_bpm - 5.0f, 60.0f), hi = fmin_rs(current_bpm + 5.0f, 180.0f);
        float best_bpm = current_bpm, best_l = 0.0f;
        for (float b = lo; b <= hi; b += 0.5f) {
            float l = likelihood(onsets, b);
            if (l > best_l) {
                best_l = l;
                best_bpm = b;
            }
        }
        current_bpm = best_bpm;
        *out_bpm = best_bpm;
        return Error{};
    }
    static float likelihood(const std::vector<float>& onsets, float bpm) {  // :203-255
        float interval = 60.0f / bpm, start = onsets[0];
        float ll = 0.0f;
        int n = 0;
        const float sigma_sq = 0.05f * 0.05f;
        for (float o : onsets) {
            int bi = as_i32(roundf((o - start) / interval));
            float eb = start + ((float)bi * interval);
            float d = fabsf(o - eb);
            ll += -(d * d) / (2.0f * sigma_sq);
            ++n;
        }
        if (n == 0) return 0.0f;
        return expf(ll / (float)n);
    }
};

// detect_time_signature — time_signature.rs:90-199; returns beats per bar
static int detect_time_signature(const std::vector<float>& beats) {
    if (beats.size() < 8) return 4;
    std::vector<float> iv;
    for (size_t i = 1; i < beats.size(); ++i) {
        float d = beats[i] - beats[i - 1];
        if (d > 0.0f) iv.push_back(d);
    }
    if (iv.empty()) return 4;
    float sum = 0.0f;
    for (float d : iv) sum += d;
    float mean = sum / (float)iv.size();
    auto score = [&](size_t lag) {
        if (iv.size() < lag) return 0.0f;
        float acc = 0.0f;
        size_t cnt = 0;
        for (size_t i = 0; i + lag < iv.size(); ++i) {
            float diff = fabsf(iv[i] - iv[i + lag]);
            acc += 1.0f / (1.0f + diff / mean);
            ++cnt;
        }
        if (cnt == 0) return 0.0f;
        float ac = acc / (float)cnt;
        float vs = 0.0f;
        for (float d : iv) {
            float df = d - mean;
            vs += df * df;
        }
        float var = vs / (float)iv.size();
        float cv = mean > EPSILON ? sqrtf(var) / mean : 1.0f;
        return fmin_rs(ac * 0.7f + (1.0f / (1.0f + cv)) * 0.3f, 1.0f);
    };
    float s44 = score(4), s34 = score(3), s68 = score(6);
    // max_by returns the LAST maximal element (order: 4/4, 3/4, 6/8)
    int best = 4;
    float bsc = s44;
    if (s34 >= bsc) {
        bsc = s34;
        best = 3;
    }
    if (s68 >= bsc) {
        bsc = s68;
        best = 6;
    }
    return best;
}

// detect_downbeats_with_time_sig — mod.rs:363-404 (beats non-empty, bpm > 0 checked by the caller)
static std::vector<float> detect_downbeats_with_time_sig(const std::vector<float>& beats, float bpm, int bpb) {
    std::vector<float> down;
    if (beats.empty()) return down;
    float beat_iv = 60.0f / bpm;
    float bar_iv = beat_iv * (float)bpb;
    float tol = bar_iv * 0.1f;
    down.push_back(beats[0]);
    for (size_t i = 1; i < beats.size(); ++i) {
        float expected = down.back() + bar_iv;
        if (fabsf(beats[i] - expected) <= tol) down.push_back(beats[i]);
    }
    return down;
}

// calculate_grid_stability — mod.rs:425-485, on the beat positions in their given order
static float calculate_grid_stability(const std::vector<float>& times) {
    float stab = 0.0f;
    if (times.size() >= 2) {
        std::vector<float> iv;
        for (size_t i = 1; i < times.size(); ++i) {
            float d = times[i] - times[i - 1];
            if (d > 0.0f) iv.push_back(d);
        }
        if (!iv.empty()) {
            float sum = 0.0f;
            for (float d : iv) sum += d;
            float mean = sum / (float)iv.size();
            if (mean > 1e-10f) {
                float vs = 0.0f;
                for (float d : iv) {
                    float df = d - mean;
                    vs += df * df;
                }
                float var = vs / (float)iv.size();
                float cv = sqrtf(var) / mean;
                stab = 1.0f / (1.0f + cv);
            }
        }
    }
    return stab;
}

// generate_beat_grid — beat_tracking/mod.rs:108-247 (+ downbeats :363-404, stability :425-485)
Error generate_beat_grid(float bpm, float bpm_conf, const std::vector<float>& onsets_s, uint32_t sr, Result& r, Dump* dump) {
    (void)bpm_conf;
    (void)sr;
    if (bpm <= 0.0f || bpm > 300.0f) return Error{INVALID_INPUT, "Invalid BPM estimate"};
    if (onsets_s.empty()) return Error{INVALID_INPUT, "Cannot generate beat grid: no onsets provided"};
    std::vector<float> onsets = onsets_s;
    std::stable_sort(onsets.begin(), onsets.end());
    std::vector<BeatPos> pos;
    std::vector<int> path;
    if (Error e = hmm_track_beats(bpm, onsets, pos, &path)) return e;
    if (pos.empty()) return Error{PROCESSING_ERROR, "HMM beat tracking produced no beats"};
    r.hmm_beat_frames.clear();
    for (auto& p : pos) r.hmm_beat_frames.push_back(p.frame);
    if (dump) {
        std::vector<int64_t> pp(path.begin(), path.end());
        dump->i["hmm.path"] = pp;
    }
    std::vector<float> bt;
    for (auto& p : pos) bt.push_back(p.time_seconds);
    std::vector<TempoSegment> segs;
    if (Error e = detect_tempo_variations(bt, bpm, segs)) return e;
    bool has_var = false;
    for (auto& s : segs) has_var |= s.variable;
    r.beats_refined = 0;
    if (has_var) {  // mod.rs:160-219
        std::vector<BeatPos> refined;
        Bayes bay{bpm};
        for (auto& sg : segs) {
            if (sg.variable) {
                std::vector<float> so_;
                for (float o : onsets)
                    if (o >= sg.start && o <= sg.end) so_.push_back(o);
                if (!so_.empty()) {
                    float ub = bpm;
                    if (Error e = bay.update(so_, &ub)) return e;
                    std::vector<BeatPos> sb;
                    if (!hmm_track_beats(ub, so_, sb)) refined.insert(refined.end(), sb.begin(), sb.end());
                }
            } else {
                for (auto& p : pos)
                    if (p.time_seconds >= sg.start && p.time_seconds <= sg.end) refined.push_back(p);
            }
        }
        if (!refined.empty()) {
            std::stable_sort(refined.begin(), refined.end(), [](const BeatPos& a, const BeatPos& b) { return a.time_seconds < b.time_seconds; });
            pos = refined;
            r.beats_refined = 1;
        }
    }
    bt.clear();
    for (auto& p : pos) bt.push_back(p.time_seconds);
    int bpb = detect_time_signature(bt);
    r.time_sig_beats_per_bar = bpb;
    // grid — mod.rs:293-321, 363-404
    std::vector<float> beats = bt;
    std::stable_sort(beats.begin(), beats.end());
    std::vector<float> down = detect_downbeats_with_time_sig(beats, bpm, bpb);
    r.beats = beats;
    r.downbeats = down;
    r.bars = down;
    std::vector<float> pt;
    for (auto& p : pos) pt.push_back(p.time_seconds);
    const float stab = calculate_grid_stability(pt);  // mod.rs:425-485 (on beat_positions order)
    r.grid_stability = stab;
    return Error{};
}

}  // namespace so

// ---- unit-level entry points (tests/test_oracle_ref_units.py: the reference's own #[test] known answers) ----------------
extern "C" {
int so_u_tempo_variations(const float* beats, int n, float nominal, float* out5, int cap) {  // rows: start, end, bpm, confidence, is_variable
    std::vector<so::TempoSegment> segs;
    so::Error e = so::detect_tempo_variations(std::vector<float>(beats, beats + n), nominal, segs);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)segs.size()); ++i) {
        out5[5 * i + 0] = segs[i].start;
        out5[5 * i + 1] = segs[i].end;
        out5[5 * i + 2] = segs[i].bpm;
        out5[5 * i + 3] = segs[i].confidence;
        out5[5 * i + 4] = segs[i].variable ? 1.0f : 0.0f;
    }
    return (int)segs.size();
}
int so_u_time_signature(const float* beats, int n) { return so::detect_time_signature(std::vector<float>(beats, beats + n)); }
float so_u_bayes_likelihood(const float* onsets, int n, float bpm) {  // compute_likelihood (bayesian.rs:203-255); empty -> 0 (:211-213)
    if (n == 0) return 0.0f;
    return so::Bayes::likelihood(std::vector<float>(onsets, onsets + n), bpm);
}
int so_u_bayes_update(float current_bpm, const float* onsets, int n, float* out_bpm) {
    so::Bayes b{current_bpm};
    so::Error e = b.update(std::vector<float>(onsets, onsets + n), out_bpm);
    return e.kind;
}
int so_u_downbeats(const float* beats, int n, float bpm, int bpb, float* out, int cap) {
    if (n > 0 && bpm <= 0.0f) return -so::INVALID_INPUT;
    std::vector<float> d = so::detect_downbeats_with_time_sig(std::vector<float>(beats, beats + n), bpm, bpb);
    for (int i = 0; i < std::min<int>(cap, (int)d.size()); ++i) out[i] = d[i];
    return (int)d.size();
}
float so_u_grid_stability(const float* times, int n) { return so::calculate_grid_stability(std::vector<float>(times, times + n)); }
// generate_beat_grid with the grid itself: beats / downbeats copied out
int so_u_beat_grid(float bpm, float conf, const float* onsets, int n, uint32_t sr, float* stability, float* beats, int* n_beats, float* down, int* n_down, int cap) {
    so::Result r;
    so::Error e = so::generate_beat_grid(bpm, conf, std::vector<float>(onsets, onsets + n), sr, r, nullptr);
    if (e) return e.kind;
    *stability = r.grid_stability;
    *n_beats = (int)r.beats.size();
    *n_down = (int)r.downbeats.size();
    for (int i = 0; i < std::min<int>(cap, *n_beats); ++i) beats[i] = r.beats[i];
    for (int i = 0; i < std::min<int>(cap, *n_down); ++i) down[i] = r.downbeats[i];
    return 0;
}
int so_u_hmm_full(float bpm, const float* onsets, int n, float* times, float* confs, int32_t* frames, int cap) {
    std::vector<so::BeatPos> b;
    so::Error e = so::hmm_track_beats(bpm, std::vector<float>(onsets, onsets + n), b, nullptr);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)b.size()); ++i) {
        times[i] = b[i].time_seconds;
        confs[i] = b[i].confidence;
        frames[i] = b[i].frame;
    }
    return (int)b.size();
}
}
