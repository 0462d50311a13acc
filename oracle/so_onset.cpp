// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates compute_stft (chroma/extractor.rs:301-359) and the onset detectors
// (onset/energy_flux.rs, spectral_flux.rs, hfc.rs, consensus.rs).
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;

// extractor.rs:318-323 — symmetric Hann in f32.
void hann_window(size_t n, std::vector<float>& w) {
    w.resize(n);
    for (size_t i = 0; i < n; ++i) {
        float x = 2.0f * PI_F * (float)i / (float)(n - 1);
        w[i] = 0.5f * (1.0f - cosf(x));
    }
}

// compute_stft — extractor.rs:301-359 (no padding, no centring; bins 0..=N/2).
Spec compute_stft(const float* s, size_t n, size_t frame_size, size_t hop) {
    Spec S;
    if (n < frame_size) return S;
    size_t nf = (n - frame_size) / hop + 1;
    S.frames = nf;
    S.bins = frame_size / 2 + 1;
    S.d.resize(nf * S.bins);
    std::vector<float> w;
    hann_window(frame_size, w);
    std::vector<float> buf(frame_size);
    std::vector<cpx> X;
    for (size_t t = 0; t < nf; ++t) {
        const float* p = s + t * hop;
        for (size_t i = 0; i < frame_size; ++i) buf[i] = p[i] * w[i];  // :342
        rfft_forward(buf.data(), frame_size, X);
        float* row = S.row(t);
        for (size_t k = 0; k < S.bins; ++k) row[k] = sqrtf(X[k].re * X[k].re + X[k].im * X[k].im);  // :352
    }
    return S;
}

// shared peak-picking rules of the three detectors (energy_flux.rs:176-217 etc.)
static void pick_peaks(const std::vector<float>& f, float thr, std::vector<size_t>& idx) {
    idx.clear();
    const size_t n = f.size();
    if (n == 0) return;
    for (size_t i = 1; i + 1 < n; ++i)
        if (f[i] > thr && f[i] > f[i - 1] && f[i] >= f[i + 1]) idx.push_back(i);
    if (n > 1 && f[0] > thr && f[0] >= f[1]) idx.push_back(0);
    if (n > 1 && f[n - 1] > thr && f[n - 1] > f[n - 2]) idx.push_back(n - 1);
}

// detect_energy_flux_onsets — energy_flux.rs:67-243
Error detect_energy_flux_onsets(const float* s, size_t n, size_t frame_size, size_t hop, float threshold_db, std::vector<size_t>& out) {
    out.clear();
    if (n == 0) return Error{};
    if (frame_size == 0) return Error{INVALID_INPUT, "Frame size must be > 0"};
    if (hop == 0) return Error{INVALID_INPUT, "Hop size must be > 0"};
    if (frame_size > n) return Error{};
    size_t nf = (n - frame_size) / hop + 1;
    if (nf < 2) return Error{};
    std::vector<float> e(nf);
    for (size_t i = 0; i < nf; ++i) {
        size_t st = i * hop, en = std::min(st + frame_size, n);
        float sum = 0.0f;
        for (size_t j = st; j < en; ++j) sum += s[j] * s[j];
        e[i] = sqrtf(sum / (float)(en - st));
    }
    std::vector<float> flux(nf - 1);
    for (size_t i = 1; i < nf; ++i) flux[i - 1] = fmax_rs(e[i] - e[i - 1], 0.0f);
    float mx = 0.0f;
    for (float v : flux) mx = fmax_rs(mx, v);
    if (mx <= EPSILON) return Error{};
    float thr = mx * powf(10.0f, threshold_db / 20.0f);  // :160
    std::vector<size_t> idx;
    pick_peaks(flux, thr, idx);
    std::vector<size_t> on;
    for (size_t i : idx) {
        size_t smp = (i + 1) * hop;
        if (smp < n) on.push_back(smp);
    }
    std::sort(on.begin(), on.end());
    for (size_t v : on)  // :224-238 dedupe within hop/2
        if (out.empty() || v >= out.back() + hop / 2) out.push_back(v);
    return Error{};
}

static float percentile_threshold(const std::vector<float>& flux, float pct) {
    std::vector<float> sorted = flux;
    std::stable_sort(sorted.begin(), sorted.end(), [](float a, float b) { return a < b; });
    size_t idx = as_usize((float)sorted.size() * pct);  // spectral_flux.rs:168
    idx = std::min(idx, sorted.size() - 1);
    return sorted[idx];
}

// detect_spectral_flux_onsets — spectral_flux.rs:69-221 (returns frame indices)
Error detect_spectral_flux_onsets(const Spec& S, float pct, std::vector<size_t>& out, std::vector<float>* flux_out) {
    out.clear();
    if (S.frames == 0) return Error{};
    if (!(pct >= 0.0f && pct <= 1.0f)) return Error{INVALID_INPUT, "Threshold percentile must be in [0, 1]"};
    if (S.bins == 0) return Error{INVALID_INPUT, "Empty magnitude frames"};
    if (S.frames < 2) return Error{};
    const size_t nb = S.bins;
    std::vector<float> prev(nb), cur(nb);
    auto normalise = [&](size_t t, std::vector<float>& dst) {
        const float* r = S.row(t);
        float mx = 0.0f;
        for (size_t k = 0; k < nb; ++k) mx = fmax_rs(mx, r[k]);
        if (mx > EPSILON)
            for (size_t k = 0; k < nb; ++k) dst[k] = r[k] / mx;
        else
            std::fill(dst.begin(), dst.end(), 0.0f);
    };
    std::vector<float> flux(S.frames - 1);
    normalise(0, prev);
    for (size_t t = 1; t < S.frames; ++t) {
        normalise(t, cur);
        float sum = 0.0f;
        for (size_t k = 0; k < nb; ++k) {
            float d = fmax_rs(cur[k] - prev[k], 0.0f);
            sum += d * d;
        }
        flux[t - 1] = sqrtf(sum);
        std::swap(prev, cur);
    }
    float thr = percentile_threshold(flux, pct);
    std::vector<size_t> idx;
    pick_peaks(flux, thr, idx);
    for (size_t i : idx) out.push_back(i + 1);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    if (flux_out) *flux_out = flux;
    return Error{};
}

// detect_hfc_onsets — hfc.rs:76-215 (returns frame indices)
Error detect_hfc_onsets(const Spec& S, uint32_t sr, float pct, std::vector<size_t>& out, std::vector<float>* flux_out) {
    out.clear();
    if (S.frames == 0) return Error{};
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (!(pct >= 0.0f && pct <= 1.0f)) return Error{INVALID_INPUT, "Threshold percentile must be in [0, 1]"};
    if (S.bins == 0) return Error{INVALID_INPUT, "Empty magnitude frames"};
    if (S.frames < 2) return Error{};
    std::vector<float> h(S.frames);
    for (size_t t = 0; t < S.frames; ++t) {
        const float* r = S.row(t);
        float acc = 0.0f;
        for (size_t k = 0; k < S.bins; ++k) acc += (float)k * r[k] * r[k];  // :137 ((k*m)*m)
        h[t] = acc;
    }
    std::vector<float> flux(S.frames - 1);
    for (size_t t = 1; t < S.frames; ++t) flux[t - 1] = fmax_rs(h[t] - h[t - 1], 0.0f);
    float thr = percentile_threshold(flux, pct);
    std::vector<size_t> idx;
    pick_peaks(flux, thr, idx);
    for (size_t i : idx) out.push_back(i + 1);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    if (flux_out) *flux_out = flux;
    return Error{};
}

// vote_onsets — consensus.rs:111-287.  lists = {energy_flux, spectral_flux, hfc, hpss}.
Error vote_onsets(const std::vector<size_t> lists[4], const float weights[4], uint32_t tol_ms, uint32_t sr, std::vector<OnsetCand>& out) {
    out.clear();
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (tol_ms == 0) return Error{INVALID_INPUT, "Tolerance must be > 0"};
    for (int i = 0; i < 4; ++i)
        if (weights[i] < 0.0f) return Error{INVALID_INPUT, "Weights must be non-negative"};
    size_t tol = as_usize((float)tol_ms / 1000.0f * (float)sr);  // :149
    struct OM {
        size_t sample;
        int method;
        float weight;
    };
    std::vector<OM> all;
    for (int m = 0; m < 4; ++m)
        for (size_t v : lists[m]) all.push_back(OM{v, m, weights[m]});
    if (all.empty()) return Error{};
    std::stable_sort(all.begin(), all.end(), [](const OM& a, const OM& b) { return a.sample < b.sample; });
    // greedy clustering exactly as written (:208-233)
    std::vector<std::vector<const OM*>> clusters;
    for (const OM& o : all) {
        bool added = false;
        for (auto& cl : clusters) {
            for (const OM* e : cl) {
                int d = (int)o.sample - (int)e->sample;  // `as i32` casts
                size_t ad = (size_t)(d < 0 ? -(long)d : (long)d);
                if (ad <= tol) {
                    cl.push_back(&o);
                    added = true;
                    break;
                }
            }
            if (added) break;
        }
        if (!added) clusters.push_back({&o});
    }
    float maxw = 0.0f;
    for (int i = 0; i < 4; ++i) maxw += weights[i];
    for (auto& cl : clusters) {
        size_t sum = 0;
        for (const OM* e : cl) sum += e->sample;
        size_t centre = sum / cl.size();
        float tw = 0.0f;
        bool voted[4] = {false, false, false, false};
        for (const OM* e : cl) {
            tw += e->weight;
            voted[e->method] = true;
        }
        uint32_t vb = voted[0] + voted[1] + voted[2] + voted[3];
        float conf = maxw > 0.0f ? clamp_rs(tw / maxw, 0.0f, 1.0f) : 0.0f;
        out.push_back(OnsetCand{centre, conf, vb});
    }
    std::stable_sort(out.begin(), out.end(), [](const OnsetCand& a, const OnsetCand& b) { return a.confidence > b.confidence; });
    return Error{};
}

}  // namespace so
