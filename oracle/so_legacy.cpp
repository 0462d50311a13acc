// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates the legacy BPM estimator: period/autocorrelation.rs:99-338,
// period/comb_filter.rs:96-215 + 342-397, period/candidate_filter.rs:51-443,
// period/mod.rs:216-404.  Its result is only consumed when the tempogram fails
// (lib.rs:894-899) but its errors propagate (lib.rs:315).
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;

// compute_autocorrelation_fft — autocorrelation.rs:229-268.
// ACF = IFFT(|FFT(x)|^2)/N on the zero-padded signal.  The power spectrum of a real
// signal is real and even, so the inverse transform equals the forward transform of
// it; we run the forward SFFT on the symmetric extension and take the real part.
static std::vector<float> acf_fft(const std::vector<float>& sig) {
    const size_t n = sig.size();
    size_t N = next_pow2(2 * n);
    std::vector<cpx> x(N, cpx{0.0f, 0.0f});
    for (size_t i = 0; i < n; ++i) x[i].re = sig[i];
    cfft_forward(x);
    for (auto& v : x) {  // *x *= x.conj()  (:245-247)
        float re = v.re * v.re - v.im * (-v.im);
        float im = v.re * (-v.im) + v.im * v.re;
        v.re = re;
        v.im = im;
    }
    // inverse via conjugate trick: ifft(y) = conj(fft(conj(y)))
    for (auto& v : x) v.im = -v.im;
    cfft_forward(x);
    float scale = 1.0f / (float)N;
    std::vector<float> acf(n);
    for (size_t i = 0; i < n; ++i) acf[i] = fmax_rs(x[i].re * scale, 0.0f);
    return acf;
}

// find_peaks_in_acf — autocorrelation.rs:282-338
static std::vector<std::pair<size_t, float>> find_peaks_in_acf(const float* a, size_t len, size_t offset) {
    std::vector<std::pair<size_t, float>> peaks;
    if (len == 0) return peaks;
    float mx = 0.0f;
    for (size_t i = 0; i < len; ++i) mx = fmax_rs(mx, a[i]);
    if (mx < EPSILON) return peaks;
    float min_prom = mx * 0.1f;
    for (size_t i = 1; i + 1 < len; ++i) {
        float v = a[i];
        if (v > a[i - 1] && v > a[i + 1]) {
            float prom = v - fmax_rs(a[i - 1], a[i + 1]);
            if (prom >= min_prom) {
                size_t lag = i + offset;
                if (peaks.empty() || std::abs((int)lag - (int)peaks.back().first) >= 2)
                    peaks.emplace_back(lag, v);
                else if (v > peaks.back().second)
                    peaks.back() = {lag, v};
            }
        }
    }
    std::stable_sort(peaks.begin(), peaks.end(), [](const std::pair<size_t, float>& x, const std::pair<size_t, float>& y) { return x.second > y.second; });
    return peaks;
}

// estimate_bpm_from_autocorrelation — autocorrelation.rs:99-216
static Error bpm_from_autocorrelation(const std::vector<size_t>& onsets, uint32_t sr, size_t hop, float min_bpm, float max_bpm,
                                      std::vector<BpmCandidate>& out) {
    out.clear();
    if (onsets.empty()) return Error{INVALID_INPUT, "Empty onset list"};
    if (sr == 0) return Error{INVALID_INPUT, "Invalid sample rate: 0"};
    if (hop == 0) return Error{INVALID_INPUT, "Invalid hop size: 0"};
    if (min_bpm <= 0.0f || max_bpm <= 0.0f || min_bpm >= max_bpm) return Error{INVALID_INPUT, "Invalid BPM range"};
    if (onsets.size() < 2) return Error{};
    size_t max_frame = *std::max_element(onsets.begin(), onsets.end()) / hop;
    size_t len = max_frame + 1;
    if (len < 2) return Error{PROCESSING_ERROR, "Signal too short for autocorrelation"};
    std::vector<float> sig(len, 0.0f);
    for (size_t o : onsets) {
        size_t f = o / hop;
        if (f < len) sig[f] = 1.0f;
    }
    std::vector<float> acf = acf_fft(sig);
    size_t lag_min = as_usize(ceilf((60.0f * (float)sr) / (max_bpm * (float)hop)));
    size_t lag_max = as_usize(floorf((60.0f * (float)sr) / (min_bpm * (float)hop)));
    if (lag_min >= lag_max || lag_min >= acf.size() || lag_max >= acf.size()) return Error{};
    auto peaks = find_peaks_in_acf(acf.data() + lag_min, lag_max - lag_min + 1, lag_min);
    float max_acf = 0.0f;
    for (float v : acf) max_acf = fmax_rs(max_acf, v);
    for (auto& p : peaks) {
        float bpm = (60.0f * (float)sr) / ((float)p.first * (float)hop);
        if (bpm >= min_bpm && bpm <= max_bpm) {
            float conf = max_acf > EPSILON ? fmin_rs(p.second / max_acf, 1.0f) : 0.0f;
            out.push_back(BpmCandidate{bpm, conf});
        }
    }
    std::stable_sort(out.begin(), out.end(), [](const BpmCandidate& a, const BpmCandidate& b) { return a.confidence > b.confidence; });
    return Error{};
}

// score_bpm_candidate — comb_filter.rs:342-397 (brute-force nearest onset, first minimum of
// the truncated distance key).
static Error score_bpm_candidate(const std::vector<size_t>& onsets, uint32_t sr, float bpm, float tol, float* score) {
    *score = 0.0f;
    if (onsets.empty()) return Error{};
    float period = (60.0f * (float)sr) / bpm;
    if (period < 1.0f) return Error{NUMERICAL_ERROR, "Invalid period"};
    float tol_s = period * tol;
    float last = (float)onsets.back();
    size_t num_beats = as_usize(ceilf(last / period)) + 1;
    size_t aligned = 0;
    // The reference scans every onset for every beat (O(beats*onsets)).  Onsets are sorted and
    // the key |onset-beat| truncated to usize is unimodal, so a moving cursor that scans a
    // bounded neighbourhood with the same key and the same first-minimum rule is equivalent.
    size_t cur = 0;
    const size_t m = onsets.size();
    for (size_t bi = 0; bi < num_beats; ++bi) {
        float eb = (float)bi * period;
        while (cur + 1 < m && (float)onsets[cur + 1] <= eb) ++cur;
        size_t lo = cur >= 3 ? cur - 3 : 0, hi = std::min(cur + 4, m);
        size_t best_i = lo;
        size_t best_k = SIZE_MAX;
        for (size_t i = lo; i < hi; ++i) {
            size_t k = as_usize(fabsf((float)onsets[i] - eb));
            if (k < best_k) {
                best_k = k;
                best_i = i;
            }
        }
        float d = fabsf((float)onsets[best_i] - eb);
        if (d <= tol_s) ++aligned;
    }
    *score = num_beats > 0 ? (float)aligned / (float)num_beats : 0.0f;
    return Error{};
}

// estimate_bpm_from_comb_filter — comb_filter.rs:96-215
static Error bpm_from_comb(const std::vector<size_t>& onsets, uint32_t sr, float min_bpm, float max_bpm, float res, std::vector<BpmCandidate>& out,
                           Dump* dump) {
    out.clear();
    if (onsets.empty()) return Error{INVALID_INPUT, "Empty onset list"};
    if (sr == 0) return Error{INVALID_INPUT, "Invalid sample rate: 0"};
    if (min_bpm <= 0.0f || max_bpm <= 0.0f || min_bpm >= max_bpm) return Error{INVALID_INPUT, "Invalid BPM range"};
    if (res <= 0.0f) return Error{INVALID_INPUT, "Invalid BPM resolution"};
    if (onsets.size() < 2) return Error{};
    std::vector<size_t> sorted = onsets;
    std::sort(sorted.begin(), sorted.end());
    std::vector<std::pair<float, float>> cands;
    float max_score = 0.0f;
    float bpm = min_bpm;
    while (bpm <= max_bpm + EPSILON) {
        float tol = clamp_rs(0.1f * (120.0f / bpm), 0.05f, 0.15f);
        float sc;
        if (Error e = score_bpm_candidate(sorted, sr, bpm, tol, &sc)) return e;
        if (sc > max_score) max_score = sc;
        cands.emplace_back(bpm, sc);
        bpm += res;
    }
    if (dump) {
        std::vector<float> raw;
        for (auto& c : cands) raw.push_back(c.second);
        dump->f["legacy.comb.raw"] = raw;
    }
    for (auto& c : cands) out.push_back(BpmCandidate{c.first, max_score > EPSILON ? c.second / max_score : 0.0f});
    std::stable_sort(out.begin(), out.end(), [](const BpmCandidate& a, const BpmCandidate& b) { return a.confidence > b.confidence; });
    out.erase(std::remove_if(out.begin(), out.end(), [](const BpmCandidate& c) { return !(c.confidence >= 0.1f); }), out.end());
    return Error{};
}

// merge_bpm_candidates — candidate_filter.rs:147-443.
// NOTE: the final comparator (:385-433) is not a strict weak order; Rust's stable merge sort
// result is then implementation-defined.  The oracle fixes it as std::stable_sort over the same
// three-way comparator (documented deviation; the legacy result is a fallback only).
static std::vector<BpmEstimate> merge_candidates(std::vector<BpmCandidate> ac, std::vector<BpmCandidate> comb) {
    std::vector<BpmEstimate> est;
    if (ac.empty() && comb.empty()) return est;
    float oct = exp2f(50.0f / 1200.0f);
    size_t top3 = std::min<size_t>(3, comb.size());
    for (auto& a : ac)
        for (size_t i = 0; i < top3; ++i) {
            float ratio = a.bpm / comb[i].bpm;
            if (fabsf(ratio / 2.0f - 1.0f) < (oct - 1.0f)) {
                bool ok = (comb[i].bpm >= 60.0f && comb[i].bpm <= 180.0f) || (a.bpm > 200.0f || a.bpm < 30.0f);
                if (ok) {
                    a.bpm = comb[i].bpm;
                    break;
                }
            }
        }
    for (auto& a : ac)
        for (size_t i = 0; i < top3; ++i) {
            float ratio = comb[i].bpm / a.bpm;
            if (fabsf(ratio / 2.0f - 1.0f) < (oct - 1.0f)) {
                if (comb[i].bpm >= 60.0f && comb[i].bpm <= 180.0f) {
                    a.bpm = comb[i].bpm;
                    break;
                }
            }
        }
    bool disagree = false;
    if (!ac.empty() && !comb.empty()) {
        float d = fabsf(ac[0].bpm - comb[0].bpm);
        disagree = d > 10.0f && d < 50.0f;
    }
    std::vector<BpmCandidate> acl(ac.begin(), ac.begin() + std::min<size_t>(10, ac.size()));
    for (auto& cnd : ac) {
        if (cnd.bpm >= 60.0f && cnd.bpm <= 180.0f) {
            bool near = false;
            for (auto& x : acl)
                if (fabsf(x.bpm - cnd.bpm) < 1.0f) near = true;
            if (!near) acl.push_back(cnd);
        }
    }
    std::vector<BpmCandidate> cl(comb.begin(), comb.begin() + std::min<size_t>(10, comb.size()));
    struct G {
        float bpm, total;
        uint32_t cnt;
        float mx;
    };
    std::vector<G> groups;
    auto add = [&](const BpmCandidate& cnd) {
        for (auto& g : groups)
            if (fabsf(cnd.bpm - g.bpm) <= 2.0f) {
                g.bpm = (g.bpm * (float)g.cnt + cnd.bpm) / (float)(g.cnt + 1);
                g.total += cnd.confidence;
                g.cnt += 1;
                g.mx = fmax_rs(g.mx, cnd.confidence);
                return;
            }
        groups.push_back(G{cnd.bpm, cnd.confidence, 1, cnd.confidence});
    };
    for (auto& x : acl) add(x);
    for (auto& x : cl) add(x);
    for (auto& g : groups) {
        float conf;
        if (g.cnt >= 2) {
            float avg = g.total / (float)g.cnt;
            conf = fmin_rs((avg + g.mx) / 2.0f * 1.2f, 1.0f);
        } else
            conf = fmin_rs(g.total, 1.0f);
        if (disagree && g.cnt == 1) conf *= 0.7f;
        est.push_back(BpmEstimate{g.bpm, conf, g.cnt});
    }
    // boost_consensus_candidates — :51-112
    size_t a5 = std::min<size_t>(5, acl.size()), c5 = std::min<size_t>(5, cl.size());
    auto harm = [](float x, float y) {
        float r = fmax_rs(x / y, y / x);
        return fabsf(r - 2.0f) < 0.1f || fabsf(r - 1.5f) < 0.1f || fabsf(r - 0.75f) < 0.1f;
    };
    for (auto& e : est) {
        bool ad = false, cd = false, ah = false, ch = false;
        for (size_t i = 0; i < a5; ++i) {
            if (fabsf(acl[i].bpm - e.bpm) < 2.5f) ad = true;
            if (harm(acl[i].bpm, e.bpm)) ah = true;
        }
        for (size_t i = 0; i < c5; ++i) {
            if (fabsf(cl[i].bpm - e.bpm) < 2.5f) cd = true;
            if (harm(cl[i].bpm, e.bpm)) ch = true;
        }
        if (ad && cd)
            e.confidence *= 1.5f;
        else if ((ad && ch) || (cd && ah))
            e.confidence *= 1.3f;
        if (cd && e.bpm >= 60.0f && e.bpm <= 180.0f) e.confidence *= 1.4f;
    }
    bool reasonable_top5 = false;
    for (size_t i = 0; i < std::min<size_t>(5, est.size()); ++i)
        if (est[i].bpm >= 60.0f && est[i].bpm <= 180.0f) reasonable_top5 = true;
    if (!reasonable_top5)
        for (auto& e : est)
            if (e.bpm >= 60.0f && e.bpm <= 180.0f) {
                e.confidence *= 2.0f;
                break;
            }
    auto cmp3 = [](const BpmEstimate& a, const BpmEstimate& b) -> int {  // <0: a first
        bool ai = a.bpm >= 60.0f && a.bpm <= 180.0f, bi = b.bpm >= 60.0f && b.bpm <= 180.0f;
        float ae = ai ? a.confidence : a.confidence * 0.5f, be = bi ? b.confidence : b.confidence * 0.5f;
        int ec = (be < ae) ? -1 : ((be > ae) ? 1 : 0);  // b_eff.partial_cmp(a_eff)
        if (fabsf(ae - be) < 0.5f) {
            if (ai && !bi) return -1;
            if (!ai && bi) return 1;
        }
        if (ec != 0) return ec;
        return (b.method_agreement < a.method_agreement) ? -1 : ((b.method_agreement > a.method_agreement) ? 1 : 0);
    };
    // insertion sort = stable, well-defined for any comparator
    for (size_t i = 1; i < est.size(); ++i) {
        BpmEstimate x = est[i];
        size_t j = i;
        while (j > 0 && cmp3(x, est[j - 1]) < 0) {
            est[j] = est[j - 1];
            --j;
        }
        est[j] = x;
    }
    return est;
}

// estimate_bpm_internal — period/mod.rs:216-404
Error estimate_bpm_legacy(const std::vector<size_t>& onsets, uint32_t sr, size_t hop, const Config& c, bool use_guardrails, bool* has,
                          BpmEstimate& out, Dump* dump) {
    *has = false;
    std::vector<BpmCandidate> ac, comb;
    if (Error e = bpm_from_autocorrelation(onsets, sr, hop, c.min_bpm, c.max_bpm, ac)) return e;
    if (Error e = bpm_from_comb(onsets, sr, c.min_bpm, c.max_bpm, c.bpm_resolution, comb, dump)) return e;
    float pmin = 60.0f, pmax = 180.0f, smin = 0, smax = 0, mp = 0, ms = 0, me = 0;
    if (use_guardrails) {  // clamp_sane :119-148
        float a = c.legacy_bpm_preferred_min, b = c.legacy_bpm_preferred_max;
        pmin = fmin_rs(a, b);
        pmax = fmax_rs(a, b);
        smin = fmin_rs(fmin_rs(c.legacy_bpm_soft_min, c.legacy_bpm_soft_max), pmin);
        smax = fmax_rs(fmax_rs(c.legacy_bpm_soft_min, c.legacy_bpm_soft_max), pmax);
        auto sane = [](float m) { return std::isfinite(m) ? fmax_rs(m, 0.0f) : 0.0f; };
        mp = sane(c.legacy_bpm_conf_mul_preferred);
        ms = sane(c.legacy_bpm_conf_mul_soft);
        me = sane(c.legacy_bpm_conf_mul_extreme);
    }
    bool has_pref = false;
    float pref_bpm = 0.0f;
    for (auto& x : ac)
        if (x.bpm >= pmin && x.bpm <= pmax) {
            has_pref = true;
            pref_bpm = x.bpm;
            break;
        }
    std::vector<BpmEstimate> merged = merge_candidates(ac, comb);
    if (use_guardrails) {
        for (auto& e : merged) {
            float mul;
            if (!std::isfinite(e.bpm))
                mul = 0.0f;
            else if (e.bpm >= pmin && e.bpm <= pmax)
                mul = mp;
            else if (e.bpm >= smin && e.bpm <= smax)
                mul = ms;
            else
                mul = me;
            e.confidence *= mul;
        }
        std::stable_sort(merged.begin(), merged.end(), [](const BpmEstimate& a, const BpmEstimate& b) { return a.confidence > b.confidence; });
    }
    if (has_pref)
        for (size_t i = 0; i < merged.size(); ++i)
            if (fabsf(merged[i].bpm - pref_bpm) < 2.0f) {
                BpmEstimate m = merged[i];
                merged.erase(merged.begin() + i);
                merged.insert(merged.begin(), m);
                break;
            }
    if (dump) {
        std::vector<float> a, b;
        for (auto& x : ac) {
            a.push_back(x.bpm);
            b.push_back(x.confidence);
        }
        dump->f["legacy.acf.bpm"] = a;
        dump->f["legacy.acf.conf"] = b;
        a.clear();
        b.clear();
        for (auto& x : comb) {
            a.push_back(x.bpm);
            b.push_back(x.confidence);
        }
        dump->f["legacy.comb.bpm"] = a;
        dump->f["legacy.comb.conf"] = b;
    }
    if (!merged.empty()) {
        *has = true;
        out = merged[0];
    }
    return Error{};
}

}  // namespace so

// ---- unit-level entry points (tests/test_oracle_ref_units.py: the reference's own #[test] known answers) ----------------
extern "C" {
static std::vector<size_t> u_onsets(const int64_t* o, int n) {
    std::vector<size_t> v;
    for (int i = 0; i < n; ++i) v.push_back((size_t)o[i]);
    return v;
}
int so_u_acf_bpm(const int64_t* onsets, int n, uint32_t sr, uint64_t hop, float min_bpm, float max_bpm, float* bpm, float* conf, int cap) {
    std::vector<so::BpmCandidate> out;
    so::Error e = so::bpm_from_autocorrelation(u_onsets(onsets, n), sr, (size_t)hop, min_bpm, max_bpm, out);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)out.size()); ++i) {
        bpm[i] = out[i].bpm;
        conf[i] = out[i].confidence;
    }
    return (int)out.size();
}
int so_u_comb_bpm(const int64_t* onsets, int n, uint32_t sr, float min_bpm, float max_bpm, float res, float* bpm, float* conf, int cap) {
    std::vector<so::BpmCandidate> out;
    so::Error e = so::bpm_from_comb(u_onsets(onsets, n), sr, min_bpm, max_bpm, res, out, nullptr);
    if (e) return -e.kind;
    for (int i = 0; i < std::min<int>(cap, (int)out.size()); ++i) {
        bpm[i] = out[i].bpm;
        conf[i] = out[i].confidence;
    }
    return (int)out.size();
}
int so_u_score_bpm(const int64_t* onsets, int n, uint32_t sr, float bpm, float tol, float* score) {
    *score = 0.0f;
    return so::score_bpm_candidate(u_onsets(onsets, n), sr, bpm, tol, score).kind;
}
int so_u_acf_fft(const float* sig, int n, float* out) {
    std::vector<float> a = so::acf_fft(std::vector<float>(sig, sig + n));
    for (size_t i = 0; i < a.size(); ++i) out[i] = a[i];
    return (int)a.size();
}
int so_u_find_peaks(const float* acf, int n, uint64_t offset, int64_t* idx, float* val, int cap) {
    auto p = so::find_peaks_in_acf(acf, (size_t)n, (size_t)offset);
    for (int i = 0; i < std::min<int>(cap, (int)p.size()); ++i) {
        idx[i] = (int64_t)p[i].first;
        val[i] = p[i].second;
    }
    return (int)p.size();
}
int so_u_merge(const float* ab, const float* ac, int na, const float* cb, const float* cc, int nc, float* bpm, float* conf, uint32_t* agree, int cap) {
    std::vector<so::BpmCandidate> a, c;
    for (int i = 0; i < na; ++i) a.push_back(so::BpmCandidate{ab[i], ac[i]});
    for (int i = 0; i < nc; ++i) c.push_back(so::BpmCandidate{cb[i], cc[i]});
    std::vector<so::BpmEstimate> m = so::merge_candidates(a, c);
    for (int i = 0; i < std::min<int>(cap, (int)m.size()); ++i) {
        bpm[i] = m[i].bpm;
        conf[i] = m[i].confidence;
        agree[i] = m[i].method_agreement;
    }
    return (int)m.size();
}
}
