// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates reference src/features/period/novelty.rs, tempogram_fft.rs and
// tempogram_autocorr.rs.
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;

static void normalize_max(std::vector<float>& v) {  // novelty.rs:935-942
    float mx = 0.0f;
    for (float x : v) mx = fmax_rs(mx, x);
    if (mx > EPSILON)
        for (float& x : v) x /= mx;
}

static void log_frame(const float* r, size_t nb, std::vector<float>& lf) {
    lf.resize(nb);
    for (size_t k = 0; k < nb; ++k) lf[k] = logf(1.0f + fmax_rs(r[k], 0.0f));  // :354
}

// superflux over bins [b0,b1) with the max filter clipped to [lo,hi) — novelty.rs:336-388 / 395-455
static std::vector<float> superflux_range(const Spec& S, size_t kk, size_t b0, size_t b1) {
    std::vector<float> flux;
    if (S.frames < 2 || S.bins == 0) return flux;
    const size_t k = std::max<size_t>(kk, 1);
    flux.resize(S.frames - 1);
    std::vector<float> prev, cur;
    log_frame(S.row(0), S.bins, prev);
    for (size_t t = 1; t < S.frames; ++t) {
        log_frame(S.row(t), S.bins, cur);
        float sum = 0.0f;
        for (size_t b = b0; b < b1; ++b) {
            size_t lo = std::max(b >= k ? b - k : 0, b0);
            size_t hi = std::min(b + k + 1, b1);
            float pm = 0.0f;
            for (size_t j = lo; j < hi; ++j) pm = fmax_rs(pm, prev[j]);
            float d = fmax_rs(cur[b] - pm, 0.0f);
            sum += d * d;
        }
        flux[t - 1] = sqrtf(sum);
        std::swap(prev, cur);
    }
    normalize_max(flux);
    return flux;
}

std::vector<float> superflux_novelty(const Spec& S, size_t k) { return superflux_range(S, k, 0, S.bins); }

std::vector<float> superflux_novelty_band(const Spec& S, size_t k, size_t bs, size_t be) {
    if (S.frames < 2 || S.bins == 0) return {};
    size_t b0 = std::min(bs, S.bins), b1 = std::min(be, S.bins);
    if (b1 <= b0 + 1) return {};
    return superflux_range(S, k, b0, b1);
}

static std::vector<float> diff_pos_norm(const std::vector<float>& v) {
    std::vector<float> flux;
    if (v.size() < 2) return flux;
    flux.resize(v.size() - 1);
    for (size_t i = 1; i < v.size(); ++i) flux[i - 1] = fmax_rs(v[i] - v[i - 1], 0.0f);
    normalize_max(flux);
    return flux;
}

// energy_flux_novelty(_band) — novelty.rs:477-545, 612-665
static std::vector<float> energy_range(const Spec& S, size_t b0, size_t b1) {
    std::vector<float> e(S.frames);
    for (size_t t = 0; t < S.frames; ++t) {
        const float* r = S.row(t);
        float acc = 0.0f;
        for (size_t k = b0; k < b1; ++k) acc += r[k] * r[k];
        e[t] = acc;
    }
    return diff_pos_norm(e);
}
std::vector<float> energy_flux_novelty(const Spec& S) {
    if (S.frames < 2 || S.bins == 0) return {};
    return energy_range(S, 0, S.bins);
}
std::vector<float> energy_flux_novelty_band(const Spec& S, size_t bs, size_t be) {
    if (S.frames < 2 || S.bins == 0) return {};
    size_t b0 = std::min(bs, S.bins), b1 = std::min(be, S.bins);
    if (b1 <= b0 + 1) return {};
    return energy_range(S, b0, b1);
}

// hfc_novelty(_band) — novelty.rs:687-768, 774-836 (absolute bin index weighting)
static std::vector<float> hfc_range(const Spec& S, size_t b0, size_t b1) {
    std::vector<float> h(S.frames);
    for (size_t t = 0; t < S.frames; ++t) {
        const float* r = S.row(t);
        float acc = 0.0f;
        for (size_t k = b0; k < b1; ++k) acc += (float)k * r[k] * r[k];
        h[t] = acc;
    }
    return diff_pos_norm(h);
}
std::vector<float> hfc_novelty(const Spec& S) {
    if (S.frames < 2 || S.bins == 0) return {};
    return hfc_range(S, 0, S.bins);
}
std::vector<float> hfc_novelty_band(const Spec& S, size_t bs, size_t be) {
    if (S.frames < 2 || S.bins == 0) return {};
    size_t b0 = std::min(bs, S.bins), b1 = std::min(be, S.bins);
    if (b1 <= b0 + 1) return {};
    return hfc_range(S, b0, b1);
}

// MelFilterbank — novelty.rs:62-190
static float mel_of(float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); }
static float inv_mel(float m) { return 700.0f * (powf(10.0f, m / 2595.0f) - 1.0f); }

struct MelFB {
    size_t n_mels = 0;
    std::vector<std::vector<std::pair<size_t, float>>> contribs;  // per bin
};

static Error mel_new(uint32_t sr, size_t n_bins, size_t n_mels_in, float fmin_hz, float fmax_hz, MelFB& fb) {
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (n_bins < 2) return Error{INVALID_INPUT, "Not enough FFT bins"};
    size_t n_mels = std::max<size_t>(n_mels_in, 4);
    float nyq = (float)sr * 0.5f;
    float fmin = fmin_rs(fmax_rs(fmin_hz, 0.0f), fmax_rs(nyq, 1.0f));
    float fmax = fmax_hz;
    if (!(std::isfinite(fmax) && fmax > 0.0f)) fmax = nyq;
    fmax = clamp_rs(fmax, fmin + 1.0f, nyq);
    size_t fft_size = (n_bins - 1) * 2;
    float res = (float)sr / (float)fft_size;
    float mmin = mel_of(fmin), mmax = mel_of(fmax);
    float step = (mmax - mmin) / (float)(n_mels + 1);
    std::vector<size_t> pts(n_mels + 2);
    for (size_t i = 0; i < n_mels + 2; ++i) {
        float hz = inv_mel(mmin + step * (float)i);
        long b = as_isize(roundf(hz / res));
        b = std::max<long>(0, std::min<long>(b, (long)n_bins - 1));
        pts[i] = (size_t)b;
    }
    for (size_t i = 1; i < pts.size(); ++i)
        if (pts[i] <= pts[i - 1]) pts[i] = std::min(pts[i - 1] + 1, n_bins - 1);
    fb.n_mels = n_mels;
    fb.contribs.assign(n_bins, {});
    for (size_t m = 0; m < n_mels; ++m) {
        size_t l = pts[m], c = pts[m + 1], r = pts[m + 2];
        if (!(l < c && c < r)) continue;
        for (size_t b = l; b <= c; ++b) {
            float w = (b == l) ? 0.0f : ((float)b - (float)l) / ((float)c - (float)l);
            if (w > 0.0f) fb.contribs[b].emplace_back(m, w);
        }
        for (size_t b = c; b <= r; ++b) {
            float w = (b == r) ? 0.0f : ((float)r - (float)b) / ((float)r - (float)c);
            if (w > 0.0f) fb.contribs[b].emplace_back(m, w);
        }
    }
    return Error{};
}

// mel_superflux_novelty — novelty.rs:553-609
Error mel_superflux_novelty(const Spec& S, uint32_t sr, size_t n_mels, float fmin, float fmax, size_t kk, std::vector<float>& out) {
    out.clear();
    if (S.frames < 2) return Error{};
    if (S.bins == 0) return Error{INVALID_INPUT, "Empty magnitude frames"};
    MelFB fb;
    if (Error e = mel_new(sr, S.bins, n_mels, fmin, fmax, fb)) return e;
    const size_t k = std::max<size_t>(kk, 1), nm = fb.n_mels;
    auto apply = [&](size_t t, std::vector<float>& mel) {
        mel.assign(nm, 0.0f);
        const float* r = S.row(t);
        for (size_t b = 0; b < S.bins; ++b) {
            float v = logf(1.0f + fmax_rs(r[b], 0.0f));
            if (v <= 0.0f) continue;
            for (auto& c : fb.contribs[b]) mel[c.first] += v * c.second;
        }
    };
    std::vector<float> prev, cur;
    apply(0, prev);
    out.resize(S.frames - 1);
    for (size_t t = 1; t < S.frames; ++t) {
        apply(t, cur);
        float sum = 0.0f;
        for (size_t b = 0; b < nm; ++b) {
            size_t lo = b >= k ? b - k : 0, hi = std::min(b + k + 1, nm);
            float pm = 0.0f;
            for (size_t j = lo; j < hi; ++j) pm = fmax_rs(pm, prev[j]);
            float d = fmax_rs(cur[b] - pm, 0.0f);
            sum += d * d;
        }
        out[t - 1] = sqrtf(sum);
        std::swap(prev, cur);
    }
    normalize_max(out);
    return Error{};
}

// combined_novelty_with_params — novelty.rs:874-932 (+ local_mean_subtract :947, smoothing :970)
std::vector<float> combined_novelty_with_params(const std::vector<float>& sp, const std::vector<float>& en, const std::vector<float>& hf,
                                                float w_s, float w_e, float w_h, size_t lmw, size_t smw) {
    size_t n = std::min(sp.size(), std::min(en.size(), hf.size()));
    if (n == 0) return {};
    float ws = fmax_rs(w_s, 0.0f), we = fmax_rs(w_e, 0.0f), wh = fmax_rs(w_h, 0.0f);
    float wsum = fmax_rs(ws + we + wh, EPSILON);
    std::vector<float> c(n);
    for (size_t i = 0; i < n; ++i) c[i] = (sp[i] * ws + en[i] * we + hf[i] * wh) / wsum;
    normalize_max(c);
    if (lmw > 1) {
        size_t half = lmw / 2;
        std::vector<float> o(n);
        for (size_t i = 0; i < n; ++i) {
            size_t st = i >= half ? i - half : 0, e = std::min(i + half + 1, n);
            float sum = 0.0f;
            for (size_t j = st; j < e; ++j) sum += c[j];
            float mean = sum / (float)(e - st);
            o[i] = fmax_rs(c[i] - mean, 0.0f);
        }
        c.swap(o);
    }
    if (smw > 1 && n >= 3) {
        size_t half = smw / 2;
        std::vector<float> orig = c;
        for (size_t i = 0; i < n; ++i) {
            size_t st = i >= half ? i - half : 0, e = std::min(i + half + 1, n);
            float sum = 0.0f;
            for (size_t j = st; j < e; ++j) sum += orig[j];
            c[i] = sum / (float)(e - st);
        }
    }
    normalize_max(c);
    return c;
}

static void sort_desc_stable(Tempogram& t) {
    std::stable_sort(t.begin(), t.end(), [](const std::pair<float, float>& a, const std::pair<float, float>& b) { return a.second > b.second; });
}

// fft_tempogram — tempogram_fft.rs:78-192
Error fft_tempogram(const std::vector<float>& nov, uint32_t sr, uint32_t hop, float min_bpm, float max_bpm, Tempogram& out) {
    out.clear();
    if (nov.empty()) return Error{INVALID_INPUT, "Novelty curve is empty"};
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (hop == 0) return Error{INVALID_INPUT, "Hop size must be > 0"};
    if (min_bpm <= 0.0f || max_bpm <= min_bpm) return Error{INVALID_INPUT, "Invalid BPM range"};
    float frame_rate = (float)sr / (float)hop;
    const size_t n = nov.size();
    float sum = 0.0f;
    for (float x : nov) sum += x;
    float mean = sum / (float)n;
    size_t fft_size = next_pow2(n);
    std::vector<float> power(fft_size / 2 + 1);
    if (fft_size >= 4) {
        std::vector<float> in(fft_size, 0.0f);
        for (size_t i = 0; i < n; ++i) {
            float w = 1.0f;
            if (n > 1) {
                float t = 2.0f * PI_F * (float)i / (float)(n - 1);
                w = 0.5f * (1.0f - cosf(t));
            }
            in[i] = (nov[i] - mean) * w;
        }
        std::vector<cpx> X;
        rfft_forward(in.data(), fft_size, X);
        for (size_t k = 0; k <= fft_size / 2; ++k) power[k] = X[k].re * X[k].re + X[k].im * X[k].im;
    } else {  // sizes 1 and 2: direct DFT
        float a = (nov[0] - mean) * (n > 1 ? 0.0f : 1.0f);
        float b = n > 1 ? (nov[1] - mean) * 0.0f : 0.0f;
        power[0] = (a + b) * (a + b);
        if (fft_size == 2) power[1] = (a - b) * (a - b);
    }
    float res = frame_rate / (float)fft_size;
    for (size_t k = 0; k <= fft_size / 2; ++k) {
        float bpm = ((float)k * res) * 60.0f;
        if (bpm >= min_bpm && bpm <= max_bpm) out.emplace_back(bpm, power[k]);
    }
    sort_desc_stable(out);
    return Error{};
}

// autocorrelation_tempogram — tempogram_autocorr.rs:79-178
Error autocorrelation_tempogram(const std::vector<float>& nov, uint32_t sr, uint32_t hop, float min_bpm, float max_bpm, float res, Tempogram& out) {
    out.clear();
    if (nov.empty()) return Error{INVALID_INPUT, "Novelty curve is empty"};
    if (sr == 0) return Error{INVALID_INPUT, "Sample rate must be > 0"};
    if (hop == 0) return Error{INVALID_INPUT, "Hop size must be > 0"};
    if (min_bpm <= 0.0f || max_bpm <= min_bpm) return Error{INVALID_INPUT, "Invalid BPM range"};
    if (res <= 0.0f) return Error{INVALID_INPUT, "BPM resolution must be > 0"};
    float frame_rate = (float)sr / (float)hop;
    const size_t n = nov.size();
    float bpm = min_bpm;
    while (bpm <= max_bpm) {
        float bps = bpm / 60.0f;
        float fpb = frame_rate / bps;
        size_t lag = as_usize(fpb);
        float sum = 0.0f;
        size_t cnt = 0;
        for (size_t i = 0; i + lag < n; ++i) {
            sum += nov[i] * nov[i + lag];
            ++cnt;
        }
        float strength = cnt > 0 ? sum / (float)cnt : 0.0f;
        out.emplace_back(bpm, strength);
        bpm += res;
    }
    sort_desc_stable(out);
    return Error{};
}

}  // namespace so
