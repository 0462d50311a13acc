// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates reference src/features/onset/hpss.rs: hpss_decompose (:71-175, iterative median filtering with soft
// masks, 10 iterations, early exit when the largest change is < 1e-6) and detect_hpss_onsets (:277-375).
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static float median_of(std::vector<float>& w) {  // hpss.rs:193-203 / 227-237
    std::stable_sort(w.begin(), w.end());
    if (w.empty()) return 0.0f;
    if (w.size() % 2 == 0) return (w[w.size() / 2 - 1] + w[w.size() / 2]) * 0.5f;
    return w[w.size() / 2];
}

Error hpss_decompose(const Spec& S, size_t margin, Spec& harmonic, Spec& percussive) {
    if (S.frames == 0) return Error{INVALID_INPUT, "Empty spectrogram"};
    if (S.bins == 0) return Error{INVALID_INPUT, "Empty frames"};
    const size_t nf = S.frames, nb = S.bins;
    harmonic = S;
    percussive = S;
    std::vector<float> hf(nf * nb), pf(nf * nb), w;
    for (int iteration = 0; iteration < 10; ++iteration) {
        const std::vector<float> hprev = harmonic.d, pprev = percussive.d;
        for (size_t b = 0; b < nb; ++b)  // horizontal (time) median for the harmonic part, :177-209
            for (size_t t = 0; t < nf; ++t) {
                const size_t st = t >= margin ? t - margin : 0, en = std::min(t + margin + 1, nf);
                w.clear();
                for (size_t f = st; f < en; ++f) w.push_back(harmonic.d[f * nb + b]);
                hf[t * nb + b] = median_of(w);
            }
        for (size_t t = 0; t < nf; ++t)  // vertical (frequency) median for the percussive part, :211-243
            for (size_t b = 0; b < nb; ++b) {
                const size_t st = b >= margin ? b - margin : 0, en = std::min(b + margin + 1, nb);
                w.assign(percussive.d.begin() + t * nb + st, percussive.d.begin() + t * nb + en);
                pf[t * nb + b] = median_of(w);
            }
        for (size_t i = 0; i < nf * nb; ++i) {  // soft masks, :128-149
            const float original = S.d[i], h = hf[i], p = pf[i];
            const float total = h + p;
            if (total > 1e-10f) {
                const float hr = h / total, pr = p / total;
                harmonic.d[i] = original * hr;
                percussive.d[i] = original * pr;
            } else {
                harmonic.d[i] = original * 0.5f;
                percussive.d[i] = original * 0.5f;
            }
        }
        if (iteration > 0) {  // :152-169
            float max_change = 0.0f;
            for (size_t i = 0; i < nf * nb; ++i) {
                max_change = fmax_rs(max_change, fabsf(harmonic.d[i] - hprev[i]));
                max_change = fmax_rs(max_change, fabsf(percussive.d[i] - pprev[i]));
            }
            if (max_change < 1e-6f) break;
        }
    }
    return Error{};
}

static void pick_peaks_(const std::vector<float>& f, float thr, std::vector<size_t>& idx) {
    idx.clear();
    const size_t n = f.size();
    if (n == 0) return;
    for (size_t i = 1; i + 1 < n; ++i)
        if (f[i] > thr && f[i] > f[i - 1] && f[i] >= f[i + 1]) idx.push_back(i);
    if (n > 1 && f[0] > thr && f[0] >= f[1]) idx.push_back(0);
    if (n > 1 && f[n - 1] > thr && f[n - 1] > f[n - 2]) idx.push_back(n - 1);
}

Error detect_hpss_onsets(const Spec& P, float pct, std::vector<size_t>& out, std::vector<float>* flux_out) {
    out.clear();
    if (P.frames == 0) return Error{};
    if (!(pct >= 0.0f && pct <= 1.0f)) return Error{INVALID_INPUT, "Threshold percentile must be in [0, 1]"};
    if (P.frames < 2) return Error{};
    std::vector<float> e(P.frames);
    for (size_t t = 0; t < P.frames; ++t) {
        const float* r = P.row(t);
        float acc = 0.0f;
        for (size_t k = 0; k < P.bins; ++k) acc += r[k] * r[k];  // :305
        e[t] = acc;
    }
    std::vector<float> flux(P.frames - 1);
    for (size_t t = 1; t < P.frames; ++t) flux[t - 1] = fmax_rs(e[t] - e[t - 1], 0.0f);
    std::vector<float> sorted = flux;
    std::stable_sort(sorted.begin(), sorted.end());
    size_t ti = as_usize((float)sorted.size() * pct);
    ti = std::min(ti, sorted.size() - 1);
    const float thr = sorted[ti];
    std::vector<size_t> idx;
    pick_peaks_(flux, thr, idx);
    for (size_t i : idx) out.push_back(i + 1);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    if (flux_out) *flux_out = flux;
    return Error{};
}

}  // namespace so
