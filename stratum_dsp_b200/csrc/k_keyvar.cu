// Optional chroma-side variants of the key path (every one is off by default, config.rs:683-741):
//   key_setup_kernel        which front end a track takes (lib.rs:1123-1197), chroma vector count
//   khpss_mask_kernel,
//   khpss_apply_kernel      harmonic_spectrogram_hpss_median_mask      chroma/extractor.rs:1369-1501
//   tuning_kernel           estimate_tuning_offset_semitones_...       chroma/extractor.rs:66-170, lib.rs:1098-1121
//   fold_table_kernel       frame_to_chroma_tuned weights per track    chroma/extractor.rs:393-487
//   whiten_kernel           per-frame spectral whitening               chroma/extractor.rs:558-580
//   logfreq_chroma_kernel   convert_linear_to_log_frequency_spectrogram + extract_chroma_from_log_frequency_spectrogram
//                                                                      chroma/extractor.rs:701-807, 941-984, lib.rs:1124-1132
//   beat_sync_kernel        extract_beat_synchronous_chroma            chroma/extractor.rs:830-922
#include "framed.cuh"
#include "kernels.h"
#include "sortnet21.h"

namespace sb {

constexpr int FOLD_MAX = 1024;   // chroma band bins (k_key.cu)
constexpr int LOG_MAX = 128;     // semitone bins of the log-frequency spectrogram (70 at 44.1 kHz)
constexpr int WHITE_RING = 64;   // whitening window <= 63 bins (the ABI rejects wider ones)

__device__ __forceinline__ float rem_euclid12(float a) {
    float r = fmodf(a, 12.0f);
    return r < 0.0f ? r + 12.0f : r;
}

// ---- per-track setup: one thread per track ----------------------------------------------------------------------------
__global__ void key_setup_kernel(TrackDev* tr, int32_t* ia, int n_tracks, DevCfg cfg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    TrackDev& T = tr[t];
    T.kf = T.Fk;
    T.key_tuning = 0.0f;
    // lib.rs:1123: beat-synchronous chroma needs a non-empty beat grid and is disabled by the log-frequency path
    T.beat_sync = (cfg.key_beat_sync && !cfg.key_log_freq && T.status == 0 && T.n_beats > 0 && T.Fk > 0) ? 1 : 0;
    if (cfg.key_tuning) ia[T.kfold_bin + 12 * FOLD_MAX + 12] = 0;
}

// ---- median-filter HPSS mask on the time-downsampled, band-limited key spectrogram -------------------------------------
// thread per (downsampled frame k, band bin b): harmonic estimate = median over k-tm..k+tm of the same bin, percussive
// estimate = median over b-fm..b+fm of the same frame, windows clipped at the edges; median = element len/2 of the sorted
// window (select_nth_unstable_by, extractor.rs:1434-1441), found with the 21-wire sorting network (+inf padding).
__device__ __forceinline__ float median21_upper(float (&v)[21], int n) {
#define CE(i, j)                          \
    {                                     \
        const float a_ = v[i], b_ = v[j]; \
        v[i] = fminf(a_, b_);             \
        v[j] = fmaxf(a_, b_);             \
    }
    SORTNET21(CE)
#undef CE
    const int mid = n >> 1;
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 21; ++i) m = (i == mid) ? v[i] : m;
    return m;
}

__device__ __forceinline__ float sanitize(float x) { return isfinite(x) ? fmaxf(x, 0.0f) : 0.0f; }

__global__ void __launch_bounds__(128) khpss_mask_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                         float* fa, DevCfg cfg) {
    const int t = blockIdx.z;
    const TrackDev& T = tr[t];
    const SrTables& st = srtab[sr_index[t]];
    const uint32_t band = st.hpss_band, b0 = st.hpss_b0;
    const uint32_t step = max(cfg.khpss_step, 1u);
    const uint32_t nf = T.Fk, KBINS = cfg.key_bins;
    if (T.status != 0 || nf == 0 || band == 0) return;
    const uint32_t n_ds = (nf + step - 1) / step;
    const uint32_t b = blockIdx.y * blockDim.x + threadIdx.x, k = blockIdx.x;
    if (b >= band || k >= n_ds) return;
    const float* K = fa + T.keyspec;
    const int tm = (int)cfg.khpss_tm, fm = (int)cfg.khpss_fm;
    const float inf = __int_as_float(0x7f800000);
    float v[21];
    int n = 0;
#pragma unroll
    for (int q = 0; q < 21; ++q) {
        const int kk = (int)k - tm + q;
        const bool in = q <= 2 * tm && kk >= 0 && kk < (int)n_ds;
        v[q] = in ? sanitize(K[(uint64_t)kk * step * KBINS + b0 + b]) : inf;
        n += in;
    }
    const float h = fmaxf(median21_upper(v, n), 0.0f);
    n = 0;
#pragma unroll
    for (int q = 0; q < 21; ++q) {
        const int bb = (int)b - fm + q;
        const bool in = q <= 2 * fm && bb >= 0 && bb < (int)band;
        v[q] = in ? sanitize(K[(uint64_t)k * step * KBINS + b0 + bb]) : inf;
        n += in;
    }
    const float per = fmaxf(median21_upper(v, n), 0.0f);
    const float p = fmaxf(cfg.khpss_power, 1.0f);
    const float hp = p == 2.0f ? h * h : powf(h, p), pp = p == 2.0f ? per * per : powf(per, p);
    fa[T.khpss_mask + (uint64_t)k * band + b] = hp / (hp + pp + 1e-12f);
}

// out[t][bin] = x * mask[min(t / step, n_ds - 1)][bin - b0] inside the band, 0 outside (extractor.rs:1486-1498); in place
__global__ void __launch_bounds__(256) khpss_apply_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                          float* fa, DevCfg cfg) {
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const SrTables& st = srtab[sr_index[t]];
    const uint32_t band = st.hpss_band, b0 = st.hpss_b0;
    const uint32_t step = max(cfg.khpss_step, 1u);
    const uint32_t nf = T.Fk;
    if (T.status != 0 || nf == 0 || band == 0) return;
    const uint32_t n_ds = (nf + step - 1) / step;
    const uint32_t f = blockIdx.x, KBINS = cfg.key_bins;
    if (f >= nf) return;
    float* row = fa + T.keyspec + (uint64_t)f * KBINS;
    const float* m = fa + T.khpss_mask + (uint64_t)min(f / step, n_ds - 1) * band;
    for (uint32_t bin = threadIdx.x; bin < KBINS; bin += blockDim.x) {
        float o = 0.0f;
        if (bin >= b0 && bin < b0 + band) o = sanitize(row[bin]) * m[bin - b0];
        row[bin] = o;
    }
}

// ---- tuning offset: one CTA per track ----------------------------------------------------------------------------------------
// Circular mean of the semitone residuals of the strong bins of every `step`-th frame.  The three sums (w sin, w cos, w) only
// reach the result through atan2 / a 0.05 concentration gate and the clamp of lib.rs:1109-1113, all tolerance-level, so they
// are tree reductions (double across the CTA) instead of the reference's serial f32 fold.
__global__ void __launch_bounds__(256) tuning_kernel(TrackDev* tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index, const float* fa,
                                                     DevCfg cfg) {
    __shared__ double red[3][8];
    const int t = blockIdx.x;
    TrackDev& T = tr[t];
    if (T.status != 0 || T.Fk == 0) return;
    const SrTables& st = srtab[sr_index[t]];
    const uint32_t lo = st.tune_bin_lo, hi = st.tune_bin_hi;
    const uint32_t step = max(cfg.tune_step, 1u);
    const float thr = clamp_rs(cfg.tune_thr, 0.0f, 1.0f);
    const float res = (float)T.sr / (float)cfg.key_frame;
    const uint32_t KBINS = cfg.key_bins;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ss = 0.0f, sc = 0.0f, sw = 0.0f;
    if (lo <= hi) {
        for (uint32_t f = (uint32_t)w * step; f < T.Fk; f += 8 * step) {
            const float* row = fa + T.keyspec + (uint64_t)f * KBINS;
            float peak = 0.0f;
            for (uint32_t b = lo + lane; b <= hi; b += 32) peak = fmaxf(peak, row[b]);
            for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
            if (peak <= 1e-12f) continue;
            const float abs_thr = peak * thr;
            for (uint32_t b = lo + lane; b <= hi; b += 32) {
                const float mag = row[b];
                if (mag < abs_thr) continue;
                const float semitone = 12.0f * log2f(((float)b * res) / 440.0f) + 57.0f;
                const float residual = semitone - roundf(semitone);
                const float wt = sqrtf(fmaxf(mag, 0.0f));  // powf(0.5)
                if (wt <= 0.0f) continue;
                const float angle = 2.0f * 3.14159265358979323846f * residual;
                ss += wt * sinf(angle);
                sc += wt * cosf(angle);
                sw += wt;
            }
        }
    }
    double ds = ss, dc = sc, dw = sw;
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dc += __shfl_xor_sync(0xffffffffu, dc, o);
        dw += __shfl_xor_sync(0xffffffffu, dw, o);
    }
    if (lane == 0) {
        red[0][w] = ds;
        red[1][w] = dc;
        red[2][w] = dw;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int q = 0; q < 8; ++q) {
            a += red[0][q];
            b += red[1][q];
            c += red[2][q];
        }
        const float sum_sin = (float)a, sum_cos = (float)b, sum_w = (float)c;
        float delta = 0.0f;
        if (!(sum_w <= 1e-6f)) {
            const float r = sqrtf(sum_sin * sum_sin + sum_cos * sum_cos) / sum_w;
            if (!(r < 0.05f)) delta = atan2f(sum_sin, sum_cos) / (2.0f * 3.14159265358979323846f);
        }
        const float lim = fabsf(cfg.tune_max_abs);
        T.key_tuning = clamp_rs(delta, -lim, lim);
    }
}

// ---- per-track chroma-folding lists for a non-zero tuning offset: 12 threads per track (one per pitch class) -------------------
// Same layout as the per-sample-rate lists of SrTables: for pitch class pc, the (bin, weight) entries in ascending bin order.
__global__ void __launch_bounds__(32) fold_table_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                        float* fa, int32_t* ia, DevCfg cfg) {
    const int t = blockIdx.x;
    const TrackDev& T = tr[t];
    if (T.status != 0 || T.Fk == 0) return;
    if (!T.beat_sync && (cfg.key_hpcp || cfg.key_log_freq)) return;  // this track never folds
    // beat-synchronous chroma passes the offset as is; the plain path only when it exceeds 1e-6 (lib.rs:1178)
    const float tuning = T.beat_sync ? T.key_tuning : (fabsf(T.key_tuning) > 1e-6f ? T.key_tuning : 0.0f);
    if (tuning == 0.0f) return;  // `semitone - 0.0`: the per-sample-rate lists are exact
    const int pc = threadIdx.x;
    const SrTables& st = srtab[sr_index[t]];
    if (pc == 12) ia[T.kfold_bin + 12 * FOLD_MAX + 12] = 1;
    if (pc >= 12) return;
    const float res = (float)T.sr / (float)cfg.key_frame;
    int32_t* bins = ia + T.kfold_bin + pc * FOLD_MAX;
    float* ws = fa + T.kfold_w + pc * FOLD_MAX;
    const float sigma = fmaxf(cfg.hpcp_sigma, 1e-6f);
    int n = 0;
    if (st.fold_lo <= st.fold_hi)
        for (uint32_t b = st.fold_lo; b <= st.fold_hi && n < FOLD_MAX; ++b) {
            const float semitone = 12.0f * log2f(((float)b * res) / 440.0f) + 57.0f - tuning;
            if (cfg.key_soft_mapping) {
                const float spc = rem_euclid12(semitone);
                const int primary = as_i32(rem_euclid12(roundf(spc)));
                for (int off = -1; off <= 1; ++off) {
                    const int tc = ((primary + off) % 12 + 12) % 12;
                    if (tc != pc) continue;
                    float dist = fabsf(spc - (float)tc);
                    dist = fminf(dist, 12.0f - dist);
                    bins[n] = (int32_t)b;
                    ws[n] = expf(-dist * dist / (2.0f * sigma * sigma));
                    ++n;
                }
            } else {
                int cls = as_i32(roundf(semitone)) % 12;
                if (cls < 0) cls += 12;
                if (cls == pc) {
                    bins[n] = (int32_t)b;
                    ws[n] = 1.0f;
                    ++n;
                }
            }
        }
    ia[T.kfold_bin + 12 * FOLD_MAX + pc] = n;
}

// ---- spectral whitening: one lane per frame -----------------------------------------------------------------------------------------
// whitened[i] = min(max(x_i, 0) / (mean_i + 1e-12), 20), mean_i from a sequential f32 prefix over the bins (extractor.rs:565-579).
// The whitened values decide which bins are peaks and which peaks are kept, so the prefix is reproduced term by term: a lane
// walks its frame's bins in order while the warp moves 32 x 32 tiles through shared memory to keep global accesses coalesced.
// Only bins [0, white_n) are produced; they never reach the right-clipped part of the window (white_n + half <= 4097 is
// checked by the launcher).
__global__ void __launch_bounds__(32) whiten_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                    float* fa, DevCfg cfg) {
    __shared__ float tin[1][32][33];
    __shared__ float tout[1][32][33];
    __shared__ float ringP[1][WHITE_RING][32];
    __shared__ float ringX[1][WHITE_RING][32];
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const int w = 0, lane = threadIdx.x & 31;
    const uint32_t f0 = blockIdx.x * 32;
    if (T.status != 0 || f0 >= T.Fk) return;
    const SrTables& st = srtab[sr_index[t]];
    const uint32_t wn = st.white_n, half = cfg.whiten_half, WS = T.kwhite_stride;
    const uint32_t nf = T.Fk, KBINS = cfg.key_bins;
    const float* K = fa + T.keyspec;
    float* W = fa + T.kwhite;
    float P = 0.0f;
    ringP[w][0][lane] = 0.0f;
    const uint32_t last = min(wn + half, KBINS);  // bins to walk
    for (uint32_t jb = 0; jb < last; jb += 32) {
        for (int r = 0; r < 32; ++r) {
            const uint32_t f = f0 + r, b = jb + lane;
            tin[w][r][lane] = (f < nf && b < KBINS) ? K[(uint64_t)f * KBINS + b] : 0.0f;
        }
        __syncwarp();
        for (int j = 0; j < 32; ++j) {
            const uint32_t i = jb + j;  // bin whose magnitude joins the prefix
            const float x = fmaxf(tin[w][lane][j], 0.0f);
            P = P + x;
            ringP[w][(i + 1) & (WHITE_RING - 1)][lane] = P;
            ringX[w][i & (WHITE_RING - 1)][lane] = x;
            float o = 0.0f;
            if (i >= half) {  // output bin c = i - half: r = c + half = i (never clipped here), l = max(c - half, 0)
                const uint32_t c = i - half;
                const uint32_t l = c >= half ? c - half : 0;
                const float denom = (float)(i + 1 - l);
                const float mean = (P - ringP[w][l & (WHITE_RING - 1)][lane]) / fmaxf(denom, 1.0f);
                o = fminf(ringX[w][c & (WHITE_RING - 1)][lane] / (mean + 1e-12f), 20.0f);
            }
            tout[w][lane][j] = o;
        }
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const uint32_t f = f0 + r;
            const int64_t c = (int64_t)jb + lane - half;
            if (f < nf && c >= 0 && c < (int64_t)wn) W[(uint64_t)f * WS + c] = tout[w][r][lane];
        }
        __syncwarp();
    }
}

// ---- log-frequency chroma: one warp per frame --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) logfreq_chroma_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                             float* fa, uint32_t KBINS) {
    __shared__ float lf[4][LOG_MAX];
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * 4 + w;
    if (T.status != 0 || f >= T.Fk) return;
    const SrTables& st = srtab[sr_index[t]];
    const float* row = fa + T.keyspec + (uint64_t)f * KBINS;
    const int ns = (int)st.log_n;
    for (int s = lane; s < ns; s += 32) {  // semitone bin s: its linear bins in ascending order (extractor.rs:773-799)
        float acc = 0.0f;
        const int a = st.log_off[s], z = st.log_off[s + 1];
        for (int q = a; q < z; ++q) {
            const float m = row[st.log_bin[q]];
            if (m > 0.0f) acc = acc + m * st.log_w[q];
        }
        lf[w][s] = acc;
    }
    __syncwarp();
    float pc = 0.0f;
    if (lane < 12) {  // extractor.rs:957-972: semitone bins ascending, pitch class = (offset + bin) rem_euclid 12
        for (int s = 0; s < ns; ++s) {
            int c = (st.log_offset + s) % 12;
            if (c < 0) c += 12;
            const float m = lf[w][s];
            if (c == lane && m > 0.0f) pc = pc + m;
        }
    }
    float ss = 0.0f;
    for (int i = 0; i < 12; ++i) {
        const float v = __shfl_sync(0xffffffffu, pc, i);
        ss = ss + v * v;
    }
    const float norm = sqrtf(ss);
    if (norm > 1e-10f) pc = pc / norm;
    if (lane < 12) fa[T.chroma + (uint64_t)f * 12 + lane] = pc;
    if (lane == 0) {  // lib.rs:1127-1130: energy of the log-frequency frame, sequential over its bins
        float e = 0.0f;
        for (int s = 0; s < ns; ++s) e = e + lf[w][s] * lf[w][s];
        fa[T.kenergy + f] = e;
    }
}

// ---- beat-synchronous chroma: one warp per beat interval ------------------------------------------------------------------------------------
// Interval i = [beats[i], beats[i+1]): the frames with start time (f as f32 * hop/sr) inside it form a contiguous run because
// the frame times are non-decreasing; their chroma vectors (chroma_fold_kernel -> chroma2) are added in frame order, divided
// by the count and L2-normalised; energies (kweights) are added in frame order (extractor.rs:872-919).
__global__ void __launch_bounds__(128) beat_sync_kernel(TrackDev* tr, float* fa, const float* __restrict__ oa, uint32_t key_hop) {
    const int t = blockIdx.y;
    TrackDev& T = tr[t];
    if (T.status != 0 || !T.beat_sync) return;
    const uint32_t ni = T.n_beats - 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) T.kf = ni;  // consumers run in later launches
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * 4 + w;
    if (i >= ni) return;
    const float* beats = oa + T.beats;
    const float b0 = beats[i], b1 = beats[i + 1];
    const float dur = (float)key_hop / (float)T.sr;  // hop_size as f32 / sample_rate as f32 (extractor.rs:859)
    const uint32_t nf = T.Fk;
    auto first_ge = [&](float x) {  // first frame with (float)f * dur >= x
        uint32_t lo = 0, hi = nf;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((float)mid * dur >= x) hi = mid;
            else lo = mid + 1;
        }
        return lo;
    };
    const uint32_t fa0 = first_ge(b0), fa1 = first_ge(b1);
    const float* ch = fa + T.chroma2;
    const float* en = fa + T.kweights;
    float acc = 0.0f;
    uint32_t cnt = 0;
    if (fa1 > fa0) {
        cnt = fa1 - fa0;
        if (lane < 12)
            for (uint32_t f = fa0; f < fa1; ++f) acc = acc + ch[(uint64_t)f * 12 + lane];
        else if (lane == 12)
            for (uint32_t f = fa0; f < fa1; ++f) acc = acc + en[f];
    }
    float pc = 0.0f;
    if (cnt > 0) {
        pc = lane < 12 ? acc / (float)cnt : 0.0f;
        float ss = 0.0f;
        for (int q = 0; q < 12; ++q) {
            const float v = __shfl_sync(0xffffffffu, pc, q);
            ss = ss + v * v;
        }
        const float norm = sqrtf(ss);
        if (norm > 1e-10f) pc = pc / norm;
    }
    const float e = __shfl_sync(0xffffffffu, acc, 12);
    if (lane < 12) fa[T.chroma + (uint64_t)i * 12 + lane] = pc;
    if (lane == 0) fa[T.kenergy + i] = cnt > 0 ? e : 0.0f;
}

// Runs after the key STFT and the cheap time mask: per-track setup, then whichever spectrogram-side variants are enabled.
void launch_key_variants_pre(const WaveCtx& c) {
    key_setup_kernel<<<(c.n_tracks + 63) / 64, 64, 0, c.stream>>>(c.tracks, c.ia, c.n_tracks, c.cfg);
    count_launch("key_mask");
    if (c.max_Fk == 0) return;
    if (c.cfg.key_hpss) {
        const uint32_t step = c.cfg.khpss_step > 1 ? c.cfg.khpss_step : 1;
        khpss_mask_kernel<<<dim3((c.max_Fk + step - 1) / step, (FOLD_MAX + 127) / 128, c.n_tracks), 128, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg);
        count_launch("key_mask");
        khpss_apply_kernel<<<dim3(c.max_Fk, c.n_tracks), 256, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg);
        count_launch("key_mask");
    }
    if (c.cfg.key_tuning && !c.cfg.key_log_freq) {
        tuning_kernel<<<c.n_tracks, 256, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg);
        count_launch("key_hpcp");
        fold_table_kernel<<<c.n_tracks, 32, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.ia, c.cfg);
        count_launch("key_hpcp");
    }
    if (c.cfg.key_whiten && c.cfg.key_hpcp && !c.cfg.key_log_freq) {
        whiten_kernel<<<dim3((c.max_Fk + 31) / 32, c.n_tracks), 32, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg);
        count_launch("key_hpcp");
    }
}

// Runs after the per-frame chroma kernels of k_key.cu.
void launch_key_chroma_variants(const WaveCtx& c) {
    if (c.max_Fk == 0) return;
    if (c.cfg.key_log_freq) {
        logfreq_chroma_kernel<<<dim3((c.max_Fk + 3) / 4, c.n_tracks), 128, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg.key_bins);
        count_launch("key_hpcp");
    }
    if (c.cfg.key_beat_sync && !c.cfg.key_log_freq) {
        beat_sync_kernel<<<dim3((c.max_beat_cap + 3) / 4, c.n_tracks), 128, 0, c.stream>>>(c.tracks, c.fa, c.oa, c.cfg.key_hop);
        count_launch("key_hpcp");
    }
}

}  // namespace sb
