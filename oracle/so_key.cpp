// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
// Restates the default key path of analyze_audio (lib.rs:961-1559): key STFT, soft harmonic
// time mask (chroma/extractor.rs:1246-1349), HPCP (extractor.rs:529-680, 1097-1150), median
// smoothing (chroma/smoothing.rs:37-94), frame weights (lib.rs:1236-1287), segment voting
// (lib.rs:1332-1436), detect_key_weighted (key/detector.rs:68-313), clarity (key_clarity.rs:51-93)
// and the Krumhansl-Kessler templates (key/templates.rs:64-143).
#include <algorithm>
#include <cmath>

#include "so_common.hpp"

namespace so {

static const float EPSILON = 1e-10f;

// smooth_spectrogram_time + harmonic_spectrogram_time_mask — extractor.rs:1246-1349
Spec harmonic_spectrogram_time_mask(const Spec& K, size_t margin, float power) {
    Spec out;
    if (K.frames == 0) return out;
    out.frames = K.frames;
    out.bins = K.bins;
    out.d.assign(K.d.size(), 0.0f);
    const size_t nf = K.frames, nb = K.bins;
    const float p = fmax_rs(power, 1.0f), eps = 1e-12f;
    std::vector<float> prefix(nf + 1);
    for (size_t b = 0; b < nb; ++b) {
        prefix[0] = 0.0f;
        for (size_t t = 0; t < nf; ++t) prefix[t + 1] = prefix[t] + K.d[t * nb + b];  // :1277-1279 sequential f32 prefix
        for (size_t t = 0; t < nf; ++t) {
            float h_est;
            if (margin == 0) {
                h_est = K.d[t * nb + b];  // smooth_spectrogram_time returns the input when margin == 0
            } else {
                size_t st = t >= margin ? t - margin : 0, en = std::min(t + margin + 1, nf);
                float sum = prefix[en] - prefix[st];
                float denom = (float)std::max<size_t>(en - st, 1);
                h_est = sum / denom;
            }
            float x = fmax_rs(K.d[t * nb + b], 0.0f);
            float h = fmax_rs(h_est, 0.0f);
            float r = fmax_rs(x - h, 0.0f);
            float hp = powf(h, p), rp = powf(r, p);
            float m = hp / (hp + rp + eps);
            out.d[t * nb + b] = x * m;
        }
    }
    return out;
}

// smooth_spectrogram_time — extractor.rs:1246-1290 (same sequential prefix as the mask above)
Spec smooth_spectrogram_time(const Spec& K, size_t margin) {
    if (K.frames == 0 || margin == 0) return K;
    Spec out;
    out.frames = K.frames;
    out.bins = K.bins;
    out.d.assign(K.d.size(), 0.0f);
    const size_t nf = K.frames, nb = K.bins;
    std::vector<float> prefix(nf + 1);
    for (size_t b = 0; b < nb; ++b) {
        prefix[0] = 0.0f;
        for (size_t t = 0; t < nf; ++t) prefix[t + 1] = prefix[t] + K.d[t * nb + b];
        for (size_t t = 0; t < nf; ++t) {
            size_t st = t >= margin ? t - margin : 0, en = std::min(t + margin + 1, nf);
            float sum = prefix[en] - prefix[st];
            float denom = (float)std::max<size_t>(en - st, 1);
            out.d[t * nb + b] = sum / denom;
        }
    }
    return out;
}

// rem_euclid for f32 (Rust): r = fmod(a,b); if r < 0 { r += |b| }
static float rem_euclid(float a, float b) {
    float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}

// frame_to_hpcp_tuned_band — extractor.rs:529-680.
// Top-K ordering: `select_nth_unstable_by` leaves the K selected peaks in an unspecified order
// (which only changes f32 accumulation order).  The oracle fixes it as: selected set = the K
// largest by (selection value desc, bin asc); accumulation in ascending bin order.
static void frame_to_hpcp_band(const float* mag, size_t nb, uint32_t sr, size_t fft_size, const Config& c, float tuning, size_t peaks_per_frame,
                               float fmin_hz, float fmax_hz, float pc[12]) {
    for (int i = 0; i < 12; ++i) pc[i] = 0.0f;
    if (nb == 0 || sr == 0 || fft_size == 0) return;
    float res = (float)sr / (float)fft_size;
    float fmin = fmax_rs(fmin_hz, 20.0f), fmax = fmin_rs(fmax_hz, (float)sr / 2.0f);
    if (fmax <= fmin) return;
    // optional per-frame spectral whitening (:558-580): moving-average over bins from a sequential f32 prefix
    std::vector<float> whitened;
    const bool use_whitening = c.enable_key_hpcp_whitening && c.key_hpcp_whitening_smooth_bins >= 3;
    if (use_whitening) {
        const size_t win = std::max<size_t>(c.key_hpcp_whitening_smooth_bins, 3) | 1;
        const size_t half = win / 2;
        std::vector<float> prefix(nb + 1, 0.0f);
        for (size_t i = 0; i < nb; ++i) prefix[i + 1] = prefix[i] + fmax_rs(mag[i], 0.0f);
        whitened.assign(nb, 0.0f);
        for (size_t i = 0; i < nb; ++i) {
            const size_t l = i >= half ? i - half : 0;
            const size_t r = std::min(i + half, nb - 1);
            const float denom = (float)(r + 1 - l);
            const float mean = (prefix[r + 1] - prefix[l]) / fmax_rs(denom, 1.0f);
            const float v = fmax_rs(mag[i], 0.0f) / (mean + 1e-12f);
            whitened[i] = fmin_rs(v, 20.0f);
        }
    }
    const float* selv = use_whitening ? whitened.data() : mag;
    std::vector<std::pair<size_t, float>> peaks;
    for (size_t b = 1; b + 1 < nb; ++b) {
        float f = (float)b * res;
        if (f < fmin) continue;
        if (f > fmax) break;
        float m = selv[b];
        if (m <= selv[b - 1] || m < selv[b + 1]) continue;
        peaks.emplace_back(b, m);
    }
    if (peaks.empty()) return;
    size_t k = std::min(std::max<size_t>(peaks_per_frame, 1), peaks.size());
    if (peaks.size() > k) {
        std::vector<std::pair<size_t, float>> byv = peaks;
        std::stable_sort(byv.begin(), byv.end(), [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) { return a.second > b.second; });
        byv.resize(k);
        std::sort(byv.begin(), byv.end(), [](const std::pair<size_t, float>& a, const std::pair<size_t, float>& b) { return a.first < b.first; });
        peaks.swap(byv);
    }
    float sigma = fmax_rs(c.soft_mapping_sigma, 1e-6f);
    size_t hmax = std::max<size_t>(c.key_hpcp_num_harmonics, 1);
    float decay = clamp_rs(c.key_hpcp_harmonic_decay, 0.0f, 1.0f);
    float p = clamp_rs(c.key_hpcp_mag_power, 0.05f, 1.0f);
    for (auto& pk : peaks) {
        float f0 = (float)pk.first * res;
        if (f0 <= 0.0f) continue;
        float w0 = powf(fmax_rs(mag[pk.first], 0.0f), p);  // original magnitude even when whitening selected the peak (:628-637)
        if (w0 <= 0.0f) continue;
        for (size_t h = 1; h <= hmax; ++h) {
            float fh = f0 * (float)h;
            if (fh > fmax) break;
            if (fh < fmin) continue;
            float semitone = 12.0f * log2f(fh / 440.0f) + 57.0f - tuning;
            float spc = rem_euclid(semitone, 12.0f);
            float ppc = rem_euclid(roundf(spc), 12.0f);
            int primary = as_i32(ppc);
            // decay.powi(h-1): llvm.powi.f32 with a run-time exponent lowers to compiler-rt's __powisf2, square-and-multiply
            // (identical to a left-to-right product up to h = 4, the default; different roundings from h = 5 on)
            float dp = powi_rs(decay, (int)h - 1);
            float hw = dp / (float)h;
            float contrib = w0 * hw;
            for (int off = -1; off <= 1; ++off) {
                int tc = ((primary + off) % 12 + 12) % 12;
                float dist = fabsf(spc - (float)tc);
                dist = fmin_rs(dist, 12.0f - dist);
                float wgt = expf(-dist * dist / (2.0f * sigma * sigma));
                pc[tc] += contrib * wgt;
            }
        }
    }
    float ss = 0.0f;
    for (int i = 0; i < 12; ++i) ss += pc[i] * pc[i];
    float norm = sqrtf(ss);
    if (norm > EPSILON)
        for (int i = 0; i < 12; ++i) pc[i] /= norm;
}

// frame_to_chroma_tuned ("chroma folding") — extractor.rs:393-487
static void frame_to_chroma(const float* mag, size_t nb, uint32_t sr, size_t fft_size, bool soft, float sigma_in, float tuning, float pc[12]) {
    for (int i = 0; i < 12; ++i) pc[i] = 0.0f;
    float res = (float)sr / (float)fft_size;
    for (size_t b = 0; b < nb; ++b) {
        float freq = (float)b * res;
        if (freq < 100.0f) continue;
        if (freq > fmin_rs(5000.0f, (float)sr / 2.0f)) break;
        if (freq >= (float)sr / 2.0f) break;
        float semitone = 12.0f * log2f(freq / 440.0f) + 57.0f - tuning;
        float contrib = powf(fmax_rs(mag[b], 0.0f), 0.6f);
        if (soft) {
            float spc = rem_euclid(semitone, 12.0f);
            float ppc = rem_euclid(roundf(spc), 12.0f);
            int primary = as_i32(ppc);
            for (int off = -1; off <= 1; ++off) {
                int tc = ((primary + off) % 12 + 12) % 12;
                float dist = fabsf(spc - (float)tc);
                dist = fmin_rs(dist, 12.0f - dist);
                float sigma = fmax_rs(sigma_in, 1e-6f);
                float w = expf(-dist * dist / (2.0f * sigma * sigma));
                pc[tc] += contrib * w;
            }
        } else {
            int cls = as_i32(roundf(semitone)) % 12;
            if (cls < 0) cls += 12;
            pc[cls] += contrib;
        }
    }
    float ss = 0.0f;
    for (int i = 0; i < 12; ++i) ss += pc[i] * pc[i];
    float norm = sqrtf(ss);
    if (norm > EPSILON)
        for (int i = 0; i < 12; ++i) pc[i] /= norm;
}

// extract_chroma_from_spectrogram_with_options_and_energy(_tuned) — extractor.rs:1028-1091
void extract_chroma(const Spec& K, uint32_t sr, size_t fft_size, bool soft, float sigma, std::vector<float>& chroma, std::vector<float>& energy, float tuning) {
    chroma.assign(K.frames * 12, 0.0f);
    energy.assign(K.frames, 0.0f);
    for (size_t t = 0; t < K.frames; ++t) {
        const float* r = K.row(t);
        float e = 0.0f;
        for (size_t b = 0; b < K.bins; ++b) e += r[b] * r[b];
        energy[t] = e;
        frame_to_chroma(r, K.bins, sr, fft_size, soft, sigma, tuning, &chroma[t * 12]);
    }
}

// extract_hpcp_from_spectrogram_with_options_and_energy_tuned — extractor.rs:1097-1150;
// extract_hpcp_bass_blend_from_spectrogram_with_options_and_energy_tuned — extractor.rs:1154-1239
void extract_hpcp(const Spec& K, uint32_t sr, size_t fft_size, const Config& c, std::vector<float>& chroma, std::vector<float>& energy, float tuning) {
    chroma.assign(K.frames * 12, 0.0f);
    energy.assign(K.frames, 0.0f);
    const float bw = clamp_rs(c.key_hpcp_bass_weight, 0.0f, 1.0f);
    for (size_t t = 0; t < K.frames; ++t) {
        const float* r = K.row(t);
        float e = 0.0f;
        for (size_t b = 0; b < K.bins; ++b) e += r[b] * r[b];
        energy[t] = e;
        float* out = &chroma[t * 12];
        frame_to_hpcp_band(r, K.bins, sr, fft_size, c, tuning, c.key_hpcp_peaks_per_frame, 100.0f, 5000.0f, out);
        if (c.enable_key_hpcp_bass_blend) {
            float bass[12];
            const size_t kb = std::min<size_t>(std::max<size_t>(c.key_hpcp_peaks_per_frame, 1), 12);  // peaks_per_frame.clamp(1, 12)
            frame_to_hpcp_band(r, K.bins, sr, fft_size, c, tuning, kb, c.key_hpcp_bass_fmin_hz, c.key_hpcp_bass_fmax_hz, bass);
            float ss = 0.0f;
            for (int i = 0; i < 12; ++i) {
                out[i] = (1.0f - bw) * out[i] + bw * bass[i];
                ss += out[i] * out[i];
            }
            const float norm = sqrtf(ss);
            if (norm > 1e-10f)
                for (int i = 0; i < 12; ++i) out[i] /= norm;
        }
    }
}

// estimate_tuning_offset_semitones_from_spectrogram — extractor.rs:66-170
float estimate_tuning_offset(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, size_t frame_step, float peak_rel_threshold) {
    if (K.frames == 0 || sr == 0 || fft_size == 0) return 0.0f;
    const float res = (float)sr / (float)fft_size;
    const float fmin = fmax_rs(fmin_hz, 20.0f);
    const float fmax = clamp_rs(fmax_hz, fmin + 1.0f, (float)sr / 2.0f);
    const size_t step = std::max<size_t>(frame_step, 1);
    const float thr = clamp_rs(peak_rel_threshold, 0.0f, 1.0f);
    float sum_sin = 0.0f, sum_cos = 0.0f, sum_w = 0.0f;
    for (size_t t = 0; t < K.frames; t += step) {
        const float* fr = K.row(t);
        float peak = 0.0f;
        for (size_t b = 0; b < K.bins; ++b) {
            const float f = (float)b * res;
            if (f < fmin) continue;
            if (f > fmax) break;
            peak = fmax_rs(peak, fr[b]);
        }
        if (peak <= 1e-12f) continue;
        const float abs_thr = peak * thr;
        for (size_t b = 0; b < K.bins; ++b) {
            const float mag = fr[b];
            if (mag < abs_thr) continue;
            const float f = (float)b * res;
            if (f < fmin) continue;
            if (f > fmax) break;
            const float semitone = 12.0f * log2f(f / 440.0f) + 57.0f;
            const float residual = semitone - roundf(semitone);
            const float w = powf(fmax_rs(mag, 0.0f), 0.5f);
            if (w <= 0.0f) continue;
            const float angle = 2.0f * PI_F * residual;
            sum_sin += w * sinf(angle);
            sum_cos += w * cosf(angle);
            sum_w += w;
        }
    }
    if (sum_w <= 1e-6f) return 0.0f;
    const float r = sqrtf(sum_sin * sum_sin + sum_cos * sum_cos) / sum_w;
    if (r < 0.05f) return 0.0f;
    return atan2f(sum_sin, sum_cos) / (2.0f * PI_F);
}

// convert_linear_to_log_frequency_spectrogram — extractor.rs:701-807
Spec linear_to_log_frequency(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, int* semitone_bin_min_out) {
    Spec out;
    const float res = (float)sr / (float)fft_size;
    const float nyq = (float)sr / 2.0f;
    const float fmin = fmax_rs(fmin_hz, 20.0f), fmax = fmin_rs(fmax_hz, nyq - 1.0f);
    const float smin = 12.0f * log2f(fmin / 440.0f) + 57.0f, smax = 12.0f * log2f(fmax / 440.0f) + 57.0f;
    const int bmin = as_i32(floorf(smin)), bmax = as_i32(ceilf(smax));
    const long nbl = (long)bmax - (long)bmin + 1;
    *semitone_bin_min_out = bmin;
    if (nbl <= 0) return out;  // (the reference's `as usize` of a negative count would allocate absurdly; unreachable for fmin < fmax)
    const size_t n_semi = (size_t)nbl;
    out.frames = K.frames;
    out.bins = n_semi;
    out.d.assign(K.frames * n_semi, 0.0f);
    for (size_t t = 0; t < K.frames; ++t) {
        const float* fr = K.row(t);
        float* lf = out.row(t);
        for (size_t b = 0; b < K.bins; ++b) {
            const float m = fr[b];
            if (m <= 0.0f) continue;
            const float f = (float)b * res;
            if (f < fmin || f >= fmax || f >= nyq) continue;
            const float semitone = 12.0f * log2f(f / 440.0f) + 57.0f;
            const float x = semitone - (float)bmin;
            const size_t lo = as_usize(floorf(x));
            const size_t hi = std::min(as_usize(ceilf(x)), n_semi - 1);
            if (lo < n_semi) {
                const float wh = x - (float)lo, wl = 1.0f - wh;
                lf[lo] += m * wl;
                if (hi != lo && hi < n_semi) lf[hi] += m * wh;
            }
        }
    }
    return out;
}

// extract_chroma_from_log_frequency_spectrogram — extractor.rs:941-984 (+ energies, lib.rs:1124-1132)
void extract_chroma_log_frequency(const Spec& L, int semitone_offset, std::vector<float>& chroma, std::vector<float>& energy) {
    chroma.assign(L.frames * 12, 0.0f);
    energy.assign(L.frames, 0.0f);
    for (size_t t = 0; t < L.frames; ++t) {
        const float* fr = L.row(t);
        float* ch = &chroma[t * 12];
        for (size_t b = 0; b < L.bins; ++b) {
            if (fr[b] <= 0.0f) continue;
            int pc = (semitone_offset + (int)b) % 12;
            if (pc < 0) pc += 12;
            ch[pc] += fr[b];
        }
        float ss = 0.0f;
        for (int i = 0; i < 12; ++i) ss += ch[i] * ch[i];
        const float norm = sqrtf(ss);
        if (norm > EPSILON)
            for (int i = 0; i < 12; ++i) ch[i] /= norm;
        float e = 0.0f;
        for (size_t b = 0; b < L.bins; ++b) e += fr[b] * fr[b];
        energy[t] = e;
    }
}

// extract_beat_synchronous_chroma — extractor.rs:830-922
void extract_beat_synchronous_chroma(const Spec& K, uint32_t sr, size_t fft_size, size_t hop, const std::vector<float>& beats, bool soft, float sigma,
                                     float tuning, std::vector<float>& chroma, std::vector<float>& energy) {
    chroma.clear();
    energy.clear();
    if (K.frames == 0 || beats.empty()) return;
    const float dur = (float)hop / (float)sr;
    const size_t ni = beats.size() - 1;
    chroma.assign(ni * 12, 0.0f);
    energy.assign(ni, 0.0f);
    for (size_t i = 0; i < ni; ++i) {
        const float b0 = beats[i], b1 = beats[i + 1];
        float acc[12];
        for (int j = 0; j < 12; ++j) acc[j] = 0.0f;
        float e_int = 0.0f;
        size_t cnt = 0;
        for (size_t f = 0; f < K.frames; ++f) {
            const float ft = (float)f * dur;
            if (ft >= b0 && ft < b1) {
                float ch[12];
                frame_to_chroma(K.row(f), K.bins, sr, fft_size, soft, sigma, tuning, ch);
                float e = 0.0f;
                const float* r = K.row(f);
                for (size_t b = 0; b < K.bins; ++b) e += r[b] * r[b];
                for (int j = 0; j < 12; ++j) acc[j] += ch[j];
                e_int += e;
                ++cnt;
            }
        }
        if (cnt > 0) {
            const float n = (float)cnt;
            float ss = 0.0f;
            for (int j = 0; j < 12; ++j) {
                acc[j] /= n;
                ss += acc[j] * acc[j];
            }
            const float norm = sqrtf(ss);
            if (norm > EPSILON)
                for (int j = 0; j < 12; ++j) acc[j] /= norm;
            for (int j = 0; j < 12; ++j) chroma[i * 12 + j] = acc[j];
            energy[i] = e_int;
        }
    }
}

// harmonic_spectrogram_hpss_median_mask — extractor.rs:1369-1501
Spec harmonic_spectrogram_hpss_median_mask(const Spec& K, uint32_t sr, size_t fft_size, float fmin_hz, float fmax_hz, size_t frame_step, size_t time_margin,
                                           size_t freq_margin, float mask_power) {
    if (K.frames == 0 || sr == 0 || fft_size == 0) return K;
    const size_t nf = K.frames, nb = K.bins;
    const float res = (float)sr / (float)fft_size;
    const float fmin = fmax_rs(fmin_hz, 20.0f);
    const float fmax = clamp_rs(fmax_hz, fmin + 1.0f, (float)sr / 2.0f);
    long bs = (long)floorf(fmin / res), be = (long)ceilf(fmax / res);
    bs = std::min<long>(std::max<long>(bs, 0), (long)nb);
    be = std::min<long>(std::max<long>(be, 0), (long)nb);
    if (be <= bs) return K;
    const size_t b0 = (size_t)bs, band = (size_t)(be - bs);
    const size_t step = std::max<size_t>(frame_step, 1);
    const size_t n_ds = std::max<size_t>((nf + step - 1) / step, 1);
    auto san = [](float x) { return std::isfinite(x) ? fmax_rs(x, 0.0f) : 0.0f; };
    std::vector<float> ds(n_ds * band), h_est(n_ds * band), p_est(n_ds * band);
    for (size_t k = 0; k < n_ds; ++k)
        for (size_t b = 0; b < band; ++b) ds[k * band + b] = K.d[(k * step) * nb + b0 + b];
    std::vector<float> scratch;
    auto median = [&](std::vector<float>& v) {  // select_nth_unstable_by(len/2): the value is order-independent
        if (v.empty()) return 0.0f;
        std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
        return v[v.size() / 2];
    };
    for (size_t b = 0; b < band; ++b)
        for (size_t t = 0; t < n_ds; ++t) {
            scratch.clear();
            const size_t st = t >= time_margin ? t - time_margin : 0, en = std::min(t + time_margin + 1, n_ds);
            for (size_t q = st; q < en; ++q) scratch.push_back(san(ds[q * band + b]));
            h_est[t * band + b] = median(scratch);
        }
    for (size_t t = 0; t < n_ds; ++t)
        for (size_t b = 0; b < band; ++b) {
            scratch.clear();
            const size_t st = b >= freq_margin ? b - freq_margin : 0, en = std::min(b + freq_margin + 1, band);
            for (size_t q = st; q < en; ++q) scratch.push_back(san(ds[t * band + q]));
            p_est[t * band + b] = median(scratch);
        }
    const float p = fmax_rs(mask_power, 1.0f);
    Spec out;
    out.frames = nf;
    out.bins = nb;
    out.d.assign(nf * nb, 0.0f);
    for (size_t t = 0; t < nf; ++t) {
        const size_t k = std::min(t / step, n_ds - 1);
        for (size_t b = 0; b < band; ++b) {
            const float h = fmax_rs(h_est[k * band + b], 0.0f), per = fmax_rs(p_est[k * band + b], 0.0f);
            const float hp = powf(h, p), pp = powf(per, p);
            const float m = hp / (hp + pp + 1e-12f);
            out.d[t * nb + b0 + b] = san(K.d[t * nb + b0 + b]) * m;
        }
    }
    return out;
}

// smooth_chroma (median) — smoothing.rs:37-94
// sharpen_chroma — chroma/normalization.rs:41-65: element-wise power, then L2 normalisation (uniform vector when the norm vanishes)
void sharpen_chroma(float* ch, float power) {
    float ss = 0.0f;
    for (int i = 0; i < 12; ++i) {
        ch[i] = powf(ch[i], power);
        ss += ch[i] * ch[i];
    }
    float norm = sqrtf(ss);
    if (norm > EPSILON)
        for (int i = 0; i < 12; ++i) ch[i] /= norm;
    else
        for (int i = 0; i < 12; ++i) ch[i] = 1.0f / sqrtf(12.0f);
}

void smooth_chroma(std::vector<float>& chroma, size_t frames, size_t window) {
    if (frames == 0 || window <= 1) return;
    if (window % 2 == 0) window += 1;
    long half = (long)(window / 2);
    std::vector<float> out(chroma.size());
    for (size_t t = 0; t < frames; ++t)
        for (int s = 0; s < 12; ++s) {
            float vals[16];
            int n = 0;
            for (long off = 0; off < (long)window && n < 16; ++off) {
                long ft = (long)t + (off - half);
                if (ft >= 0 && (size_t)ft < frames) vals[n++] = chroma[(size_t)ft * 12 + s];
            }
            if (n > 0) {
                std::stable_sort(vals, vals + n);
                out[t * 12 + s] = vals[n / 2];
            } else
                out[t * 12 + s] = chroma[t * 12 + s];
        }
    chroma.swap(out);
}

// KeyTemplates::new_krumhansl_kessler — templates.rs:64-143; new_temperley — templates.rs:145-222
void key_templates(float major[12][12], float minor[12][12], int template_set) {
    const float kk_maj[12] = {6.35f, 2.23f, 3.48f, 2.33f, 4.38f, 4.09f, 2.52f, 5.19f, 2.39f, 3.66f, 2.29f, 2.88f};
    const float kk_min[12] = {6.33f, 2.68f, 3.52f, 5.38f, 2.60f, 3.53f, 2.54f, 4.75f, 3.98f, 2.69f, 3.34f, 3.17f};
    const float tp_maj[12] = {5.0f, 2.0f, 3.5f, 2.0f, 4.5f, 4.0f, 2.0f, 4.5f, 2.0f, 3.5f, 1.5f, 4.0f};
    const float tp_min[12] = {5.0f, 2.0f, 3.5f, 5.0f, 2.0f, 3.5f, 2.0f, 4.5f, 3.5f, 2.0f, 4.0f, 3.5f};
    const float* cmaj = template_set == 1 ? tp_maj : kk_maj;
    const float* cmin = template_set == 1 ? tp_min : kk_min;
    for (int k = 0; k < 12; ++k)
        for (int s = 0; s < 12; ++s) {
            major[k][s] = cmaj[(s + 12 - k) % 12];
            minor[k][s] = cmin[(s + 12 - k) % 12];
        }
    auto l2 = [](float* v) {
        float ss = 0.0f;
        for (int i = 0; i < 12; ++i) ss += v[i] * v[i];
        float n = sqrtf(ss);
        if (n > 1e-12f)
            for (int i = 0; i < 12; ++i) v[i] /= n;
    };
    for (int k = 0; k < 12; ++k) {
        l2(major[k]);
        l2(minor[k]);
    }
}

// compute_key_clarity — key_clarity.rs:51-93 (scores in ranked order)
float compute_key_clarity(const float* sc, size_t n) {
    if (n < 2) return 0.0f;
    float best = sc[0], sum = 0.0f;
    for (size_t i = 0; i < n; ++i) sum += sc[i];
    float avg = sum / (float)n;
    float mn = sc[0], mx = sc[0];
    for (size_t i = 1; i < n; ++i) {
        if (sc[i] < mn) mn = sc[i];   // min_by: first minimal
        if (sc[i] >= mx) mx = sc[i];  // max_by: last maximal (same value either way)
    }
    float range = mx - mn;
    if (range > 1e-10f) return clamp_rs((best - avg) / range, 0.0f, 1.0f);
    return 0.0f;
}

// detect_key_weighted — detector.rs:68-313.  Key ids: 0..11 major, 12..23 minor.
Error detect_key_weighted(const float* chroma, size_t frames, const float* w, KeyScores& out, int template_set) {
    if (frames == 0) return Error{INVALID_INPUT, "Empty chroma vectors"};
    float maj[12][12], mnr[12][12];
    key_templates(maj, mnr, template_set);
    float sc[24];
    for (int k = 0; k < 24; ++k) {
        const float* tpl = k < 12 ? maj[k] : mnr[k - 12];
        float acc = 0.0f;
        for (size_t t = 0; t < frames; ++t) {  // weighted_sum_dot :984-1001
            const float* ch = chroma + t * 12;
            if (w) {
                if (w[t] > 0.0f) {
                    float dot = 0.0f;
                    for (int i = 0; i < 12; ++i) dot += ch[i] * tpl[i];
                    acc += w[t] * dot;
                }
            } else {
                float dot = 0.0f;
                for (int i = 0; i < 12; ++i) dot += ch[i] * tpl[i];
                acc += dot;
            }
        }
        sc[k] = acc;
    }
    float mxM = 0.0f, mxm = 0.0f;
    for (int k = 0; k < 12; ++k) mxM = fmax_rs(mxM, sc[k]);
    for (int k = 12; k < 24; ++k) mxm = fmax_rs(mxm, sc[k]);
    if (mxM > 1e-9f && mxm > 1e-9f) {
        for (int k = 0; k < 12; ++k) sc[k] /= mxM;
        for (int k = 12; k < 24; ++k) sc[k] /= mxm;
    }
    const int cof[12] = {0, 7, 2, 9, 4, 11, 6, 1, 8, 3, 10, 5};
    int pos_of[12];
    for (int i = 0; i < 12; ++i) pos_of[cof[i]] = i;
    // top key per mode: max_by -> last maximal (:177-198)
    int topM = 0, topm = 12;
    for (int k = 0; k < 12; ++k)
        if (sc[k] >= sc[topM]) topM = k;
    for (int k = 12; k < 24; ++k)
        if (sc[k] >= sc[topm]) topm = k;
    float refined[24];
    for (int k = 0; k < 24; ++k) {
        refined[k] = sc[k];
        int ref = k < 12 ? topM : topm;
        float ref_score = sc[ref];
        if (ref_score > 1e-9f) {
            int tp = pos_of[k % 12], rp = pos_of[ref % 12];
            int d = std::abs(tp - rp);
            int dist = std::min(d, 12 - d);
            if (dist <= 2) {
                float bonus = 0.20f * (1.0f - (float)dist * 0.5f);
                refined[k] += ref_score * bonus;
            }
        }
    }
    int order[24];
    for (int k = 0; k < 24; ++k) order[k] = k;
    std::stable_sort(order, order + 24, [&](int a, int b) { return refined[a] > refined[b]; });
    for (int i = 0; i < 24; ++i) {
        out.keys[i] = order[i];
        out.scores[i] = refined[order[i]];
    }
    // weighted top-3 voting (:254-275): the three keys are distinct, so each vote is
    // score/best and the best key wins unless scores tie exactly (HashMap order in the
    // reference; first-ranked here).
    int final_key = out.keys[0];
    float final_score = out.scores[0];
    float best_other = out.scores[1];
    {   // which calls enter the vote with an exact tie (instrumentation for the parity report)
        const float thr = final_score * 0.95f;
        const bool use_vote = out.scores[1] >= thr && out.scores[2] >= thr * 0.90f;
        out.vote_tie = (use_vote && final_score > 0.0f && out.scores[1] / final_score == final_score / final_score) ? 1 : 0;
    }
    out.key = final_key;
    out.confidence = final_score > 0.0f ? clamp_rs((final_score - best_other) / final_score, 0.0f, 1.0f) : 0.0f;
    return Error{};
}

// detect_key_weighted_mode_heuristic — detector.rs:326-518.  out.keys/out.scores = the post-bonus score table in
// ranked order (stable re-sort of the base ranking), out.key = chosen key (after the optional parallel-mode flip).
Error detect_key_weighted_mode_heuristic(const float* chroma, size_t frames, const float* w, int template_set, float third_ratio_margin,
                                         float flip_min_score_ratio, bool minor_bonus, float minor_bonus_weight, KeyScores& out) {
    KeyScores base;
    if (Error e = detect_key_weighted(chroma, frames, w, base, template_set)) return e;
    const float flip_ratio = clamp_rs(flip_min_score_ratio, 0.0f, 1.0f);
    const bool enable_flip = flip_ratio > 0.0f;
    out = base;
    if (!minor_bonus && !enable_flip) return Error{};
    float avg[12];
    for (int i = 0; i < 12; ++i) avg[i] = 0.0f;
    float wsum = 0.0f;
    if (!w) {
        for (size_t t = 0; t < frames; ++t)
            for (int i = 0; i < 12; ++i) avg[i] += chroma[t * 12 + i];
        wsum = (float)frames;
    } else {
        for (size_t t = 0; t < frames; ++t) {
            const float wt = w[t];
            if (wt <= 0.0f) continue;
            for (int i = 0; i < 12; ++i) avg[i] += wt * chroma[t * 12 + i];
            wsum += wt;
        }
    }
    if (wsum <= 1e-9f) return Error{};
    float sum = 0.0f;
    for (int i = 0; i < 12; ++i) sum += avg[i];
    if (sum > 1e-9f)
        for (int i = 0; i < 12; ++i) avg[i] /= sum;
    int keys[24];
    float sc[24];
    for (int i = 0; i < 24; ++i) {
        keys[i] = base.keys[i];
        sc[i] = base.scores[i];
    }
    if (minor_bonus) {
        const float bw = fmax_rs(minor_bonus_weight, 0.0f);
        if (bw > 0.0f)
            for (int i = 0; i < 24; ++i)
                if (keys[i] >= 12) {
                    const int tonic = keys[i] - 12;
                    const int lt = (tonic + 11) % 12, b7 = (tonic + 10) % 12;
                    sc[i] += wsum * bw * (avg[lt] - avg[b7]);
                }
    }
    int order[24];
    for (int i = 0; i < 24; ++i) order[i] = i;
    std::stable_sort(order, order + 24, [&](int a, int b) { return sc[a] > sc[b]; });
    float by_key[24];
    for (int i = 0; i < 24; ++i) {
        out.keys[i] = keys[order[i]];
        out.scores[i] = sc[order[i]];
        by_key[out.keys[i]] = out.scores[i];
    }
    const int best_key = out.keys[0];
    const int tonic = best_key % 12;
    const bool best_is_major = best_key < 12;
    const float p_min3 = avg[(tonic + 3) % 12], p_maj3 = avg[(tonic + 4) % 12];
    const float p_min6 = avg[(tonic + 8) % 12], p_maj6 = avg[(tonic + 9) % 12];
    const float p_min7 = avg[(tonic + 10) % 12], p_maj7 = avg[(tonic + 11) % 12];
    const float margin = fmax_rs(third_ratio_margin, 0.0f);
    float minor_score = 0.0f, major_score = 0.0f;
    const float third_diff = fabsf(p_min3 - p_maj3);
    if (p_min3 > (p_maj3 * (1.0f + margin))) minor_score += third_diff * 2.0f;
    else if (p_maj3 > (p_min3 * (1.0f + margin))) major_score += third_diff * 2.0f;
    const float sixth_diff = fabsf(p_min6 - p_maj6);
    if (p_min6 > (p_maj6 * (1.0f + margin))) minor_score += sixth_diff * 1.0f;
    else if (p_maj6 > (p_min6 * (1.0f + margin))) major_score += sixth_diff * 1.0f;
    const float seventh_diff = fabsf(p_min7 - p_maj7);
    if (p_min7 > (p_maj7 * (1.0f + margin))) minor_score += seventh_diff * 1.0f;
    else if (p_maj7 > (p_min7 * (1.0f + margin))) major_score += seventh_diff * 1.0f;
    const float total = minor_score + major_score;
    const bool minor_pref = total > 1e-9f ? minor_score > major_score * (1.0f + margin * 0.5f) : false;
    const bool major_pref = total > 1e-9f ? major_score > minor_score * (1.0f + margin * 0.5f) : false;
    int chosen = best_key;
    if (enable_flip) {
        if (best_is_major && minor_pref) {
            const float s_best = by_key[tonic], s_alt = by_key[12 + tonic];
            if (s_best > 0.0f && s_alt >= s_best * flip_ratio) chosen = 12 + tonic;
        } else if (!best_is_major && major_pref) {
            const float s_best = by_key[12 + tonic], s_alt = by_key[tonic];
            if (s_best > 0.0f && s_alt >= s_best * flip_ratio) chosen = tonic;
        }
    }
    const float chosen_score = by_key[chosen];
    float best_other = 0.0f;
    for (int i = 0; i < 24; ++i)
        if (out.keys[i] != chosen) best_other = fmax_rs(best_other, out.scores[i]);
    out.key = chosen;
    out.confidence = chosen_score > 0.0f ? clamp_rs((chosen_score - best_other) / chosen_score, 0.0f, 1.0f) : 0.0f;
    return Error{};
}

// Segment/whole-track detector selected by the config (lib.rs:1353-1372, 1386-1410, 1437-1453)
static Error detect_one(const float* chroma, size_t frames, const float* w, const Config& c, KeyScores& out) {
    if (c.enable_key_mode_heuristic || c.enable_key_minor_harmonic_bonus)
        return detect_key_weighted_mode_heuristic(chroma, frames, w, c.key_template_set, c.key_mode_third_ratio_margin,
                                                  c.enable_key_mode_heuristic ? c.key_mode_flip_min_score_ratio : 0.0f, c.enable_key_minor_harmonic_bonus,
                                                  c.key_minor_leading_tone_bonus_weight, out);
    return detect_key_weighted(chroma, frames, w, out, c.key_template_set);
}

// Final table from 24 accumulated scores indexed by key id: stable sort, confidence (lib.rs:1412-1435, detector.rs:672-699)
static void finish_accumulated(const float* acc, KeyScores& out) {
    int order[24];
    for (int k = 0; k < 24; ++k) order[k] = k;
    std::stable_sort(order, order + 24, [&](int a, int b) { return acc[a] > acc[b]; });
    for (int i = 0; i < 24; ++i) {
        out.keys[i] = order[i];
        out.scores[i] = acc[order[i]];
    }
    out.key = out.keys[0];
    out.confidence = out.scores[0] > 0.0f ? clamp_rs((out.scores[0] - out.scores[1]) / out.scores[0], 0.0f, 1.0f) : 0.0f;
}

// detect_key_multi_scale — detector.rs:546-700
static Error detect_key_multi_scale(const float* chroma, size_t frames, const float* w, const Config& c, KeyScores& out) {
    if (frames == 0) return Error{INVALID_INPUT, "Empty chroma vectors"};
    if (c.key_multi_scale_n_lengths == 0) return Error{INVALID_INPUT, "No segment lengths provided for multi-scale detection"};
    float acc[24];
    for (int k = 0; k < 24; ++k) acc[k] = 0.0f;
    float total_weight = 0.0f;
    size_t used = 0;
    const float min_cl = clamp_rs(c.key_multi_scale_min_clarity, 0.0f, 1.0f);
    const size_t hop = std::max<size_t>(c.key_multi_scale_hop, 1);
    for (uint32_t si = 0; si < c.key_multi_scale_n_lengths; ++si) {
        const size_t seg_len = c.key_multi_scale_lengths[si];
        if (seg_len == 0 || seg_len > frames) continue;
        const float scale_weight = (c.key_multi_scale_n_weights > 0 && si < c.key_multi_scale_n_weights) ? c.key_multi_scale_weights[si] : 1.0f;
        if (scale_weight <= 0.0f) continue;
        for (size_t st = 0; st + seg_len <= frames; st += hop) {
            KeyScores ks;
            if (Error e = detect_one(chroma + st * 12, seg_len, w ? w + st : nullptr, c, ks)) return e;
            const float cl = compute_key_clarity(ks.scores, 24);
            if (cl >= min_cl) {
                ++used;
                const float cw = cl * scale_weight;
                total_weight += cw;
                for (int i = 0; i < 24; ++i) acc[ks.keys[i]] += ks.scores[i] * cw;
            }
        }
    }
    if (used == 0 || total_weight <= 1e-12f) return detect_one(chroma, frames, w, c, out);
    for (int k = 0; k < 24; ++k) acc[k] /= total_weight;
    finish_accumulated(acc, out);
    return Error{};
}

// detect_key_ensemble — detector.rs:881-976
static Error detect_key_ensemble(const float* chroma, size_t frames, const float* w, float kk_weight, float temperley_weight, KeyScores& out) {
    const float total = kk_weight + temperley_weight;
    const float kk_norm = total > 1e-9f ? kk_weight / total : 0.5f;
    const float tp_norm = total > 1e-9f ? temperley_weight / total : 0.5f;
    KeyScores kk, tp;
    if (Error e = detect_key_weighted(chroma, frames, w, kk, 0)) return e;
    if (Error e = detect_key_weighted(chroma, frames, w, tp, 1)) return e;
    float a[24], b[24], comb[24];
    for (int i = 0; i < 24; ++i) {
        a[kk.keys[i]] = kk.scores[i];
        b[tp.keys[i]] = tp.scores[i];
    }
    for (int k = 0; k < 24; ++k) comb[k] = kk_norm * a[k] + tp_norm * b[k];
    finish_accumulated(comb, out);
    return Error{};
}

static float chroma_tonalness(const float* ch) {  // lib.rs:1236-1251
    float sum = 0.0f;
    for (int i = 0; i < 12; ++i) sum += ch[i];
    if (sum <= 1e-12f) return 0.0f;
    float ent = 0.0f;
    for (int i = 0; i < 12; ++i) {
        float p = ch[i] / sum;
        if (p > 1e-12f) ent -= p * logf(p);
    }
    float t = 1.0f - (ent / logf(12.0f));
    return clamp_rs(t, 0.0f, 1.0f);
}

// Key section of analyze_audio — lib.rs:961-1559, default-config branches only.
Error detect_key_path(const float* s, size_t n, uint32_t sr, const Config& c, const Spec& S_base, Result& r, Dump* dump, const std::vector<float>* beat_times) {
    r.key_is_minor = 0;
    r.key_index = 0;
    r.key_confidence = 0.0f;
    r.key_clarity = 0.0f;
    if (n < c.frame_size) return Error{};
    size_t kfft = c.enable_key_stft_override ? std::max<size_t>(c.key_stft_frame_size, 256) : c.frame_size;
    size_t khop = c.enable_key_stft_override ? std::max<size_t>(c.key_stft_hop_size, 1) : c.hop_size;
    Spec Kown;
    if (c.enable_key_stft_override) Kown = compute_stft(s, n, kfft, khop);
    const Spec& K = c.enable_key_stft_override ? Kown : S_base;
    Spec masked;
    const Spec* forkey = &K;
    if (!K.empty() && c.enable_key_hpss_harmonic) {  // lib.rs:1011-1030
        masked = harmonic_spectrogram_hpss_median_mask(K, sr, kfft, 100.0f, 5000.0f, c.key_hpss_frame_step, c.key_hpss_time_margin, c.key_hpss_freq_margin,
                                                       c.key_hpss_mask_power);
        forkey = &masked;
    } else if (!K.empty() && c.enable_key_harmonic_mask) {  // lib.rs:1031-1042
        masked = harmonic_spectrogram_time_mask(K, c.key_spectrogram_smooth_margin, c.key_harmonic_mask_power);
        forkey = &masked;
    } else if (!K.empty() && c.enable_key_spectrogram_time_smoothing) {
        masked = smooth_spectrogram_time(K, c.key_spectrogram_smooth_margin);
        forkey = &masked;
    }
    // optional log-frequency (semitone-aligned) spectrogram (lib.rs:1064-1094); disables HPCP, tuning and beat-sync
    bool use_log = false;
    Spec logspec;
    int semitone_offset = 0;
    if (c.enable_key_log_frequency && !forkey->empty()) {
        int bmin = 0;
        logspec = linear_to_log_frequency(*forkey, sr, kfft, 100.0f, 5000.0f, &bmin);
        use_log = true;
        // lib.rs:1076-1079 recomputes the offset from fmin = 100 (same value as the converter's)
        semitone_offset = as_i32(floorf(12.0f * log2f(100.0f / 440.0f) + 57.0f));
        (void)bmin;
    }
    float tuning = 0.0f;
    if (c.enable_key_tuning_compensation && !forkey->empty() && !use_log) {  // lib.rs:1098-1121
        const float d = estimate_tuning_offset(*forkey, sr, kfft, 80.0f, 2000.0f, c.key_tuning_frame_step, c.key_tuning_peak_rel_threshold);
        const float lim = fabsf(c.key_tuning_max_abs_semitones);
        tuning = clamp_rs(d, -lim, lim);
    }
    if (dump) dump->f["key.tuning"] = std::vector<float>(1, tuning);
    std::vector<float> chroma, energy;
    size_t nf_chroma = forkey->frames;
    if (c.enable_key_beat_synchronous && beat_times && !beat_times->empty() && !use_log) {  // lib.rs:1124-1136
        if (forkey->empty()) return Error{};
        extract_beat_synchronous_chroma(*forkey, sr, kfft, khop, *beat_times, c.soft_chroma_mapping, c.soft_mapping_sigma, tuning, chroma, energy);
        nf_chroma = energy.size();
    } else if (use_log) {
        extract_chroma_log_frequency(logspec, semitone_offset, chroma, energy);
    } else if (c.enable_key_hpcp) {
        extract_hpcp(*forkey, sr, kfft, c, chroma, energy, tuning);
    } else {  // lib.rs:1178-1196: the tuned variant with |tuning| <= 1e-6 equals the untuned one up to `semitone - 0.0`
        const float tn = (c.enable_key_tuning_compensation && fabsf(tuning) > 1e-6f) ? tuning : 0.0f;
        extract_chroma(*forkey, sr, kfft, c.soft_chroma_mapping, c.soft_mapping_sigma, chroma, energy, tn);
    }
    size_t nf = nf_chroma;
    if (c.chroma_sharpening_power > 1.0f)  // lib.rs:1200-1208
        for (size_t t = 0; t < nf; ++t) sharpen_chroma(&chroma[t * 12], c.chroma_sharpening_power);
    if (dump) {
        dump->f["key.hpcp_raw"] = chroma;
        dump->f["key.energy"] = energy;
    }
    if (nf > 5) smooth_chroma(chroma, nf, 5);
    if (dump) dump->f["key.hpcp_smooth"] = chroma;

    // optional edge trimming (lib.rs:1216-1233): the chroma / energy slices every later step works on
    size_t f0 = 0;
    if (c.enable_key_edge_trim && nf >= 200) {
        const float frac = clamp_rs(c.key_edge_trim_fraction, 0.0f, 0.49f);
        const size_t start = as_usize(roundf((float)nf * frac));
        const size_t end = as_usize(roundf((float)nf * (1.0f - frac)));
        if (end > start + 50 && end <= nf) {
            f0 = start;
            nf = end - start;
        }
    }
    const float* ch = chroma.data() + f0 * 12;
    const float* en = energy.data() + f0;
    std::vector<float> weights;
    bool have_w = false;
    if (c.enable_key_frame_weighting && nf > 0) {  // lib.rs:1253-1287
        std::vector<float> sorted(en, en + nf);
        std::stable_sort(sorted.begin(), sorted.end());
        float median = fmax_rs(sorted[sorted.size() / 2], 1e-12f);
        weights.resize(nf);
        for (size_t t = 0; t < nf; ++t) {
            float tonal = chroma_tonalness(&ch[t * 12]);
            if (tonal < c.key_min_tonalness) tonal = 0.0f;
            float e_norm = fmax_rs(en[t] / median, 0.0f);
            float wt = powf(tonal, fmax_rs(c.key_tonalness_power, 0.0f));
            float we = powf(e_norm, fmax_rs(c.key_energy_power, 0.0f));
            weights[t] = fmax_rs(wt * we, 0.0f);
        }
        float sw = 0.0f;
        size_t used = 0;
        for (float x : weights) {
            sw += x;
            if (x > 0.0f) ++used;
        }
        have_w = !(sw <= 1e-12f || used < 10);
    }
    if (dump && have_w) dump->f["key.weights"] = weights;
    const float* wp = have_w ? weights.data() : nullptr;

    KeyScores ks;
    float all_scores[24];
    int all_keys[24];
    float confidence;
    int key;
    bool done = false;
    size_t ms_min = 0;
    for (uint32_t i = 0; i < c.key_multi_scale_n_lengths; ++i) ms_min = i == 0 ? c.key_multi_scale_lengths[0] : std::min<size_t>(ms_min, c.key_multi_scale_lengths[i]);
    if (c.enable_key_ensemble) {  // lib.rs:1290-1298
        if (detect_key_ensemble(ch, nf, wp, c.key_ensemble_kk_weight, c.key_ensemble_temperley_weight, ks)) return Error{};
        done = true;
    } else if (c.enable_key_multi_scale && c.key_multi_scale_n_lengths > 0 && nf >= ms_min) {  // lib.rs:1304-1330
        if (detect_key_multi_scale(ch, nf, wp, c, ks)) return Error{};
        done = true;
    } else if (c.enable_key_segment_voting && nf >= std::max<size_t>(c.key_segment_len_frames, 1) && c.key_segment_len_frames >= 120 &&
               c.key_segment_hop_frames >= 1) {  // lib.rs:1332-1436
        size_t seg_len = std::min(c.key_segment_len_frames, nf);
        size_t hop = std::max<size_t>(std::min(c.key_segment_hop_frames, seg_len), 1);
        float min_cl = clamp_rs(c.key_segment_min_clarity, 0.0f, 1.0f);
        float acc[24];
        for (int k = 0; k < 24; ++k) acc[k] = 0.0f;
        size_t used = 0;
        std::vector<float> seg_dump;
        for (size_t st = 0; st + seg_len <= nf; st += hop) {
            if (Error e = detect_one(&ch[st * 12], seg_len, wp ? wp + st : nullptr, c, ks)) return e;  // `?` at lib.rs:1369/1371
            float cl = compute_key_clarity(ks.scores, 24);
            if (dump) {
                seg_dump.push_back((float)ks.keys[0]);
                seg_dump.push_back(cl);
            }
            if (cl >= min_cl) {
                ++used;
                for (int i = 0; i < 24; ++i) acc[ks.keys[i]] += ks.scores[i] * cl;
            }
        }
        if (dump) dump->f["key.segments"] = seg_dump;
        if (used > 0) {
            finish_accumulated(acc, ks);
            done = true;
        }
    }
    if (!done) {
        // chroma extraction/key detection failure degrades to the default key (lib.rs:1542-1551)
        if (detect_one(ch, nf, wp, c, ks)) return Error{};
    }
    for (int i = 0; i < 24; ++i) {
        all_keys[i] = ks.keys[i];
        all_scores[i] = ks.scores[i];
    }
    key = ks.key;
    confidence = ks.confidence;
    r.key_hashmap_tie = (!done && ks.vote_tie) ? 1 : 0;  // segment voting / multi-scale / ensemble rank accumulated scores: no HashMap involved
    float clarity = compute_key_clarity(all_scores, 24);
    r.key_is_minor = key >= 12;
    r.key_index = (uint32_t)(key % 12);
    r.key_confidence = confidence;
    r.key_clarity = clarity;
    if (dump) {
        std::vector<float> a(all_scores, all_scores + 24), b;
        for (int i = 0; i < 24; ++i) b.push_back((float)all_keys[i]);
        dump->f["key.scores"] = a;
        dump->f["key.order"] = b;
    }
    return Error{};
}

}  // namespace so
