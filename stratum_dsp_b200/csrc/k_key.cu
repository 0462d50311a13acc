// Key detection kernels — the default key path of analyze_audio (reference lib.rs:961-1559):
//   soft harmonic time mask            chroma/extractor.rs:1246-1349   (mask_kernel, in place)
//   HPCP per frame + frame energy      chroma/extractor.rs:529-680, 1097-1150 (hpcp_kernel)
//   5-tap median smoothing             chroma/smoothing.rs:37-94       (chroma_smooth_kernel)
//   tonalness / energy frame weights   lib.rs:1236-1287                (key_weights_kernel)
//   segment template scores            key/detector.rs:118-133, 984-1001 (segment_score_kernel)
//   normalise, circle-of-fifths bonus, ranking, clarity, clarity-weighted vote
//                                      key/detector.rs:135-313, key_clarity.rs:51-93, lib.rs:1332-1436 (key_vote_kernel)
//
// HBM traffic per key frame: the 4097-bin magnitude row is written once by the STFT, read + rewritten
// in place by the mask and read once by the HPCP kernel (4 x 16 KB); everything after that is 12 floats
// per frame.
#include "framed.cuh"
#include "kernels.h"

namespace sb {

constexpr int KBINS = 4097;
constexpr int RING = 32;           // mask ring: supports margin <= 15
constexpr int HPCP_MAX_PEAKS = 512;  // local maxima in the 100..5000 Hz band (<= (hi-lo+2)/2 = 456 at 44.1 kHz)
constexpr int HPCP_MAX_SEL = 32;     // key_hpcp_peaks_per_frame upper bound accepted by the ABI
constexpr int HPCP_MAX_HARM = 8;     // key_hpcp_num_harmonics upper bound accepted by the ABI

// ---- soft harmonic mask: one thread per (track, bin), strictly sequential f32 prefix over time ----
// The reference builds `prefix[t+1] = prefix[t] + x[t]` per bin and takes window sums as prefix
// differences (extractor.rs:1274-1287); the cancellation noise of that formulation is part of its
// output, so the scan is reproduced term by term.  Adjacent threads own adjacent bins, so every
// load and store of a frame row is coalesced.  Frames are consumed in groups of 16 whose loads are
// issued together (16 independent requests in flight per thread); stores only touch rows already
// consumed, so the mask is written in place.  MG > 0: margin known at compile time, the delayed sample
// x[t - MG] is then a register; MG = 0: run-time margin with a shared-memory ring for the samples.
constexpr int MASK_G = 16;

template <int MG>
__global__ void __launch_bounds__(128) mask_kernel(const TrackDev* __restrict__ tr, float* fa, DevCfg cfg) {
    __shared__ float ringP[RING][128];
    __shared__ float ringX[MG > 0 ? 1 : RING][128];
    const TrackDev& T = tr[blockIdx.y];
    const uint32_t nf = T.Fk;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (T.status != 0 || nf == 0 || b >= KBINS) return;
    const uint32_t mg = MG > 0 ? (uint32_t)MG : cfg.key_margin;
    float* K = fa + T.keyspec + b;
    const int tx = threadIdx.x;
    const float p = fmaxf(cfg.key_mask_power, 1.0f);
    const bool square = (p == 2.0f);
    float P = 0.0f;
    ringP[0][tx] = 0.0f;
    auto emit = [&](uint32_t t, uint32_t en, float Pen, float xt) {
        float h_est;
        if (mg == 0) {
            h_est = xt;
        } else {
            const uint32_t st = t >= mg ? t - mg : 0;
            const float sum = Pen - ringP[st & (RING - 1)][tx];
            const float denom = (float)max(en - st, 1u);
            h_est = sum / denom;
        }
        if (cfg.key_smooth_only) {  // smooth_spectrogram_time alone (extractor.rs:1246-1290, lib.rs:1043-1060)
            K[(uint64_t)t * KBINS] = h_est;
            return;
        }
        const float x = fmaxf(xt, 0.0f);
        const float h = fmaxf(h_est, 0.0f);
        const float r = fmaxf(x - h, 0.0f);
        const float hp = square ? h * h : powf(h, p);
        const float rp = square ? r * r : powf(r, p);
        const float m = hp / (hp + rp + 1e-12f);
        K[(uint64_t)t * KBINS] = x * m;
    };
    float xp[MASK_G];  // previous group (compile-time margin only)
#pragma unroll
    for (int q = 0; q < MASK_G; ++q) xp[q] = 0.0f;
    uint32_t i = 0;
    for (; i + MASK_G <= nf; i += MASK_G) {
        float xs[MASK_G];
#pragma unroll
        for (int q = 0; q < MASK_G; ++q) xs[q] = K[(uint64_t)(i + q) * KBINS];
#pragma unroll
        for (int q = 0; q < MASK_G; ++q) {
            const uint32_t ii = i + q;
            if (MG == 0) ringX[ii & (RING - 1)][tx] = xs[q];
            P = P + xs[q];
            if (ii >= mg) {  // prefix[t - mg] was written 2*mg+1 steps ago: still in the ring
                float xt;
                if (MG > 0) xt = (q >= MG) ? xs[q >= MG ? q - MG : 0] : xp[q + MASK_G - MG < MASK_G ? q + MASK_G - MG : 0];
                else xt = ringX[(ii - mg) & (RING - 1)][tx];
                emit(ii - mg, ii + 1, P, xt);
            }
            ringP[(ii + 1) & (RING - 1)][tx] = P;
        }
#pragma unroll
        for (int q = 0; q < MASK_G; ++q) xp[q] = xs[q];
    }
    // tail (< 16 frames) and flush: the delayed samples are re-read from rows that are still unmasked
    // (row t is only rewritten by emit(t)), which costs at most 16 + margin scalar loads per thread
    for (; i < nf; ++i) {
        const float x = K[(uint64_t)i * KBINS];
        P = P + x;
        if (i >= mg) emit(i - mg, i + 1, P, K[(uint64_t)(i - mg) * KBINS]);
        ringP[(i + 1) & (RING - 1)][tx] = P;
    }
    for (uint32_t t = nf > mg ? nf - mg : 0; t < nf; ++t) emit(t, nf, P, K[(uint64_t)t * KBINS]);
}

// ---- HPCP: one warp per frame ---------------------------------------------------------------------
struct HpcpSmem {
    float mag[HPCP_MAX_PEAKS];
    uint16_t bin[HPCP_MAX_PEAKS];
    uint16_t sel[HPCP_MAX_SEL];
    float val[HPCP_MAX_SEL * HPCP_MAX_HARM * 3];
    int8_t tc[HPCP_MAX_SEL * HPCP_MAX_HARM * 3];
};

__device__ __forceinline__ float rem_euclid_f(float a, float b) {
    float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}

__global__ void __launch_bounds__(128) hpcp_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                   float* fa, DevCfg cfg) {
    __shared__ HpcpSmem sm[4];
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * 4 + w;
    if (T.status != 0 || f >= T.Fk) return;
    HpcpSmem& S = sm[w];
    const SrTables& st = srtab[sr_index[t]];
    const float* row = fa + T.keyspec + (uint64_t)f * KBINS;
    // frame energy (extractor.rs:1132-1134).  Consumed only through (E/median)^0.5 frame weights, a
    // tolerance-level quantity, so the 4097-term sum is a warp tree instead of a serial fold.
    float e = 0.0f;
    for (uint32_t k = lane; k < KBINS; k += 32) {
        const float x = row[k];
        e = e + x * x;
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    // local maxima in the band, compacted in ascending bin order (extractor.rs:582-606)
    const uint32_t lo = st.key_bin_lo, hi = st.key_bin_hi;  // inclusive; lo > hi = empty band
    uint32_t np = 0;
    if (lo <= hi) {
        for (uint32_t base = lo; base <= hi; base += 32) {
            const uint32_t b = base + lane;
            bool pk = false;
            float m = 0.0f;
            if (b <= hi) {
                m = row[b];
                pk = !(m <= row[b - 1] || m < row[b + 1]);
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, pk);
            if (pk) {
                const uint32_t pos = np + __popc(mask & ((1u << lane) - 1u));
                if (pos < HPCP_MAX_PEAKS) {
                    S.mag[pos] = m;
                    S.bin[pos] = (uint16_t)b;
                }
            }
            np += __popc(mask);
        }
        np = min(np, (uint32_t)HPCP_MAX_PEAKS);
    }
    __syncwarp();
    float pc = 0.0f;  // lanes 0..11 own one pitch class each
    if (np > 0) {
        // top-K by (magnitude desc, bin asc); the reference's select_nth_unstable_by leaves the K
        // survivors in unspecified order — the documented rule is "accumulate in ascending bin order".
        // The K-th largest magnitude is found by a bitwise search on the (positive) float bit patterns:
        // 31 rounds of "how many peaks are >= candidate", each a register compare + one warp reduction.
        const uint32_t K = min(max(cfg.hpcp_peaks, 1u), np);
        uint32_t thr = 0, need = 0;  // keep magnitudes > thr, plus the first `need` ties in bin order
        if (np > K) {
            uint32_t e[HPCP_MAX_PEAKS / 32];
#pragma unroll
            for (int q = 0; q < HPCP_MAX_PEAKS / 32; ++q) {
                const uint32_t i = q * 32 + lane;
                e[q] = i < np ? __float_as_uint(S.mag[i]) : 0u;  // peaks are strictly positive, 0 never counts
            }
            for (int bit = 30; bit >= 0; --bit) {
                const uint32_t cand = thr | (1u << bit);
                uint32_t c = 0;
#pragma unroll
                for (int q = 0; q < HPCP_MAX_PEAKS / 32; ++q) c += e[q] >= cand;
                if (__reduce_add_sync(0xffffffffu, c) >= K) thr = cand;
            }
            uint32_t gt = 0;
#pragma unroll
            for (int q = 0; q < HPCP_MAX_PEAKS / 32; ++q) gt += e[q] > thr;
            need = K - __reduce_add_sync(0xffffffffu, gt);
        }
        uint32_t nsel = 0, ties_seen = 0;
        for (uint32_t base = 0; base < np; base += 32) {
            const uint32_t i = base + lane;
            bool keep = false, tie = false;
            if (i < np) {
                if (np <= K) {
                    keep = true;
                } else {
                    const uint32_t bits = __float_as_uint(S.mag[i]);
                    keep = bits > thr;
                    tie = bits == thr;
                }
            }
            const uint32_t tmask = __ballot_sync(0xffffffffu, tie);
            if (tie && ties_seen + __popc(tmask & ((1u << lane) - 1u)) < need) keep = true;
            ties_seen += __popc(tmask);
            const uint32_t mask = __ballot_sync(0xffffffffu, keep);
            if (keep) S.sel[nsel + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)i;
            nsel += __popc(mask);
        }
        __syncwarp();
        const float res = (float)T.sr / 8192.0f;
        const float fmin = fmaxf(100.0f, 20.0f), fmax = fminf(5000.0f, (float)T.sr / 2.0f);
        const float sigma = fmaxf(cfg.hpcp_sigma, 1e-6f);
        const uint32_t hmax = min(max(cfg.hpcp_harm, 1u), (uint32_t)HPCP_MAX_HARM);
        const float decay = clamp_rs(cfg.hpcp_decay, 0.0f, 1.0f);
        const float pw = clamp_rs(cfg.hpcp_pow, 0.05f, 1.0f);
        const uint32_t items = nsel * hmax;
        for (uint32_t it = lane; it < items; it += 32) {
            const uint32_t pi = it / hmax, h = it % hmax + 1;
            const uint32_t idx = S.sel[pi];
            const float f0 = (float)S.bin[idx] * res;
            const float mg = fmaxf(S.mag[idx], 0.0f);
            const float w0 = (pw == 0.5f) ? sqrtf(mg) : powf(mg, pw);
            const float fh = f0 * (float)h;
            int8_t* tcs = S.tc + it * 3;
            float* vals = S.val + it * 3;
            // `break` at fh > fmax and `continue` at fh < fmin both leave this (peak, h) without a contribution
            if (!(f0 > 0.0f) || !(w0 > 0.0f) || fh > fmax || fh < fmin) {
                tcs[0] = tcs[1] = tcs[2] = -1;
                continue;
            }
            const float semitone = 12.0f * log2f(fh / 440.0f) + 57.0f;
            const float spc = rem_euclid_f(semitone, 12.0f);
            const float ppc = rem_euclid_f(roundf(spc), 12.0f);
            const int primary = as_i32(ppc);
            float dp = 1.0f;
            for (uint32_t q = 1; q < h; ++q) dp = dp * decay;
            const float hw = dp / (float)h;
            const float contrib = w0 * hw;
#pragma unroll
            for (int off = -1; off <= 1; ++off) {
                const int tc = ((primary + off) % 12 + 12) % 12;
                float dist = fabsf(spc - (float)tc);
                dist = fminf(dist, 12.0f - dist);
                const float wgt = expf(-dist * dist / (2.0f * sigma * sigma));
                tcs[off + 1] = (int8_t)tc;
                vals[off + 1] = contrib * wgt;
            }
        }
        __syncwarp();
        if (lane < 12) {
            const uint32_t ne = items * 3;
            for (uint32_t q = 0; q < ne; ++q)
                if (S.tc[q] == lane) pc = pc + S.val[q];
        }
        // L2 normalise with the reference's sequential sum of squares (extractor.rs:668-677)
        float ss = 0.0f;
        for (int i = 0; i < 12; ++i) {
            const float v = __shfl_sync(0xffffffffu, pc, i);
            ss = ss + v * v;
        }
        const float norm = sqrtf(ss);
        if (norm > 1e-10f) pc = pc / norm;
    }
    if (lane < 12) fa[T.chroma + (uint64_t)f * 12 + lane] = pc;
    if (lane == 0) fa[T.kenergy + f] = e;
}

// ---- chroma folding (extractor.rs:393-487, enable_key_hpcp = false): one warp per frame ---------------------------
// chroma[pc] = sum over band bins (ascending) of max(x,0)^0.6 * w[bin][pc]; the weights (Gaussian soft mapping in
// circular pitch-class space, or the hard nearest-class assignment) depend only on the bin, so they come from a
// per-sample-rate table grouped by pitch class in the reference's accumulation order.  The 0.6-power of every band bin
// is evaluated once, in parallel, into shared memory; 12 lanes then fold their pitch class.
constexpr int FOLD_MAX = 1024;  // band bins (912 at 44.1 kHz; the ABI rejects sample rates whose band is wider)

__global__ void __launch_bounds__(128) chroma_fold_kernel(const TrackDev* __restrict__ tr, const SrTables* __restrict__ srtab, const int32_t* __restrict__ sr_index,
                                                          float* fa) {
    __shared__ float contrib[4][FOLD_MAX];
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t f = blockIdx.x * 4 + w;
    if (T.status != 0 || f >= T.Fk) return;
    const SrTables& st = srtab[sr_index[t]];
    const float* row = fa + T.keyspec + (uint64_t)f * KBINS;
    float e = 0.0f;  // frame energy: tree sum (tolerance-level consumer, see hpcp_kernel)
    for (uint32_t k = lane; k < KBINS; k += 32) {
        const float x = row[k];
        e = e + x * x;
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    const uint32_t lo = st.fold_lo, hi = st.fold_hi;
    float pc = 0.0f;
    if (lo <= hi) {
        for (uint32_t b = lo + lane; b <= hi; b += 32) contrib[w][b - lo] = powf(fmaxf(row[b], 0.0f), 0.6f);  // extractor.rs:431
        __syncwarp();
        if (lane < 12) {
            const int a = st.fold_off[lane], z = st.fold_off[lane + 1];
            for (int q = a; q < z; ++q) pc = pc + contrib[w][st.fold_bin[q] - lo] * st.fold_w[q];
        }
        float ss = 0.0f;
        for (int i = 0; i < 12; ++i) {
            const float v = __shfl_sync(0xffffffffu, pc, i);
            ss = ss + v * v;
        }
        const float norm = sqrtf(ss);
        if (norm > 1e-10f) pc = pc / norm;
    }
    if (lane < 12) fa[T.chroma + (uint64_t)f * 12 + lane] = pc;
    if (lane == 0) fa[T.kenergy + f] = e;
}

// ---- sharpen_chroma (chroma/normalization.rs:41-65): thread per frame ----------------------------------------------
__global__ void __launch_bounds__(256) chroma_sharpen_kernel(const TrackDev* __restrict__ tr, float* fa, float power) {
    const TrackDev& T = tr[blockIdx.y];
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (T.status != 0 || f >= T.Fk) return;
    float* ch = fa + T.chroma + (uint64_t)f * 12;
    float v[12], ss = 0.0f;
    for (int i = 0; i < 12; ++i) {
        v[i] = powf(ch[i], power);
        ss = ss + v[i] * v[i];
    }
    const float norm = sqrtf(ss);
    for (int i = 0; i < 12; ++i) ch[i] = norm > 1e-10f ? v[i] / norm : 1.0f / sqrtf(12.0f);
}

// ---- 5-tap median over time per pitch class (smoothing.rs:37-94); applied when Fk > 5 --------------
__global__ void __launch_bounds__(256) chroma_smooth_kernel(const TrackDev* __restrict__ tr, float* fa) {
    const TrackDev& T = tr[blockIdx.y];
    const uint32_t nf = T.Fk;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (T.status != 0 || i >= nf * 12) return;
    const float* src = fa + T.chroma;
    float* dst = fa + T.chroma2;
    if (nf <= 5) {
        dst[i] = src[i];
        return;
    }
    const int t = (int)(i / 12), s = (int)(i % 12);
    float v[5];
    int n = 0;
    for (int off = -2; off <= 2; ++off) {
        const int ft = t + off;
        if (ft >= 0 && ft < (int)nf) v[n++] = src[(uint64_t)ft * 12 + s];
    }
    for (int a = 1; a < n; ++a) {  // insertion sort; equal keys are indistinguishable
        const float x = v[a];
        int j = a;
        while (j > 0 && v[j - 1] > x) {
            v[j] = v[j - 1];
            --j;
        }
        v[j] = x;
    }
    dst[i] = v[n / 2];
}

// ---- frame weights (lib.rs:1236-1287): one CTA per track -------------------------------------------
__global__ void __launch_bounds__(256) key_weights_kernel(TrackDev* tr, float* fa, DevCfg cfg) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bc[2];
    __shared__ float sred[32];
    __shared__ uint32_t sused[32];
    TrackDev& T = tr[blockIdx.x];
    const uint32_t nf = T.Fk;
    if (threadIdx.x == 0) T.have_w = 0;
    if (T.status != 0 || nf == 0 || !cfg.key_weighting) return;
    const float* en = fa + T.kenergy;
    const float median = fmaxf(block_select_kth(en, nf, nf / 2, hist, bc), 1e-12f);
    const float* ch = fa + T.chroma2;
    float* wv = fa + T.kweights;
    const float tp = fmaxf(cfg.key_tonal_pow, 0.0f), ep = fmaxf(cfg.key_energy_pow, 0.0f);
    const float ln12 = logf(12.0f);
    float sw = 0.0f;
    uint32_t used = 0;
    for (uint32_t t = threadIdx.x; t < nf; t += blockDim.x) {
        const float* c = ch + (uint64_t)t * 12;
        float sum = 0.0f;
        for (int i = 0; i < 12; ++i) sum = sum + c[i];
        float tonal = 0.0f;
        if (sum > 1e-12f) {
            float ent = 0.0f;
            for (int i = 0; i < 12; ++i) {
                const float p = c[i] / sum;
                if (p > 1e-12f) ent = ent - p * logf(p);
            }
            tonal = clamp_rs(1.0f - (ent / ln12), 0.0f, 1.0f);
        }
        if (tonal < cfg.key_min_tonal) tonal = 0.0f;
        const float e = fmaxf(en[t] / median, 0.0f);
        const float wt = (tp == 2.0f) ? tonal * tonal : powf(tonal, tp);
        const float we = (ep == 0.5f) ? sqrtf(e) : powf(e, ep);
        const float wgt = fmaxf(wt * we, 0.0f);
        wv[t] = wgt;
        sw += wgt;
        used += wgt > 0.0f;
    }
    for (int o = 16; o > 0; o >>= 1) {
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
        used += __shfl_xor_sync(0xffffffffu, used, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sred[threadIdx.x >> 5] = sw;
        sused[threadIdx.x >> 5] = used;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        uint32_t u = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) {
            s += sred[q];
            u += sused[q];
        }
        T.have_w = !(s <= 1e-12f || u < 10) ? 1 : 0;
    }
}

// Segment geometry shared by the two kernels below (lib.rs:1332-1352).
__device__ __forceinline__ bool seg_geometry(const TrackDev& T, const DevCfg& cfg, uint32_t* seg_len, uint32_t* hop, uint32_t* nseg) {
    const uint32_t nf = T.Fk;
    const bool voting = cfg.key_voting && nf >= max(cfg.key_seg_len, 1u) && cfg.key_seg_len >= 120 && cfg.key_seg_hop >= 1;
    if (!voting) {
        *seg_len = nf;
        *hop = 1;
        *nseg = 0;
        return false;
    }
    *seg_len = min(cfg.key_seg_len, nf);
    *hop = max(min(cfg.key_seg_hop, *seg_len), 1u);
    *nseg = (nf - *seg_len) / *hop + 1;
    return true;
}

// ---- template scores: one warp per (segment, key) ------------------------------------------------------
// score = sum over frames (in frame order) of w_t * <c_t, T_k> (detector.rs:984-1001).  The 32 lanes
// evaluate the 12-term dot products of 32 consecutive frames in parallel (each in the reference's term
// order); the running sum then absorbs the 32 products strictly in frame order through shuffles, so the
// result equals the serial fold bit for bit while the chain per frame is one add instead of ~25 ops.
// blockIdx.x = segment index; the extra index `nseg` is the whole-track score used when no segment
// passes the clarity gate (lib.rs:1385-1411) or voting is off.  blockDim = 24 warps, warp = key.
// The whole-track row of a track that votes by segments is only needed when no segment passes the clarity gate, so the
// first launch skips it and a second launch (fallback_only) computes it for the tracks key_vote_kernel flagged.
__global__ void __launch_bounds__(768) segment_score_kernel(const TrackDev* __restrict__ tr, float* fa, Tables tab, DevCfg cfg, int fallback_only) {
    const TrackDev& T = tr[blockIdx.y];
    if (T.status != 0 || T.Fk == 0) return;
    uint32_t seg_len, hop, nseg;
    const bool voting = seg_geometry(T, cfg, &seg_len, &hop, &nseg);
    const uint32_t s = fallback_only ? nseg : blockIdx.x;
    if (s > nseg || s >= T.seg_cap) return;
    if (fallback_only ? !T.key_fallback : (voting && s == nseg)) return;
    const uint32_t start = s < nseg ? s * hop : 0;
    const uint32_t len = s < nseg ? seg_len : T.Fk;
    const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* tpl = (k < 12 ? tab.key_major + k * 12 : tab.key_minor + (k - 12) * 12);
    float tp[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) tp[i] = tpl[i];
    const float* ch = fa + T.chroma2 + (uint64_t)start * 12;
    const float* wv = T.have_w ? fa + T.kweights + start : nullptr;
    float acc = 0.0f;
    for (uint32_t t0 = 0; t0 < len; t0 += 32) {
        const uint32_t t = t0 + lane;
        float prod = 0.0f;
        bool use = false;
        if (t < len) {
            const float* c = ch + (uint64_t)t * 12;
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < 12; ++i) dot = dot + c[i] * tp[i];
            if (wv) {
                const float wt = wv[t];
                use = wt > 0.0f;
                prod = wt * dot;
            } else {
                use = true;
                prod = dot;
            }
        }
        const uint32_t um = __ballot_sync(0xffffffffu, use);
#pragma unroll
        for (int l = 0; l < 32; ++l) {
            const float p = __shfl_sync(0xffffffffu, prod, l);
            if ((um >> l) & 1u) acc = acc + p;
        }
    }
    if (lane == 0) fa[T.seg_scores + (uint64_t)s * 24 + k] = acc;
}

// detect_key_weighted steps 1.5 - 2 on 24 raw scores: ranked keys/scores out (detector.rs:135-250) and the
// same refined scores indexed by key id (by_key).  Kept out of line: each call gets its own stack
// frame, so the caller's accumulators never share local-memory slots with this function's temporaries.
__device__ __noinline__ void rank_keys(const float* raw, int* keys, float* scores, float* by_key) {
    float sc[24];
    for (int k = 0; k < 24; ++k) sc[k] = raw[k];
    float mxM = 0.0f, mxm = 0.0f;
    for (int k = 0; k < 12; ++k) mxM = fmaxf(mxM, sc[k]);
    for (int k = 12; k < 24; ++k) mxm = fmaxf(mxm, sc[k]);
    if (mxM > 1e-9f && mxm > 1e-9f) {
        for (int k = 0; k < 12; ++k) sc[k] = sc[k] / mxM;
        for (int k = 12; k < 24; ++k) sc[k] = sc[k] / mxm;
    }
    const int pos_of[12] = {0, 7, 2, 9, 4, 11, 6, 1, 8, 3, 10, 5};  // position of tonic on the circle of fifths (self-inverse table)
    int topM = 0, topm = 12;
    for (int k = 0; k < 12; ++k)
        if (sc[k] >= sc[topM]) topM = k;  // max_by: last maximal
    for (int k = 12; k < 24; ++k)
        if (sc[k] >= sc[topm]) topm = k;
    float refined[24];
    for (int k = 0; k < 24; ++k) {
        refined[k] = sc[k];
        const int ref = k < 12 ? topM : topm;
        const float ref_score = sc[ref];
        if (ref_score > 1e-9f) {
            const int d = abs(pos_of[k % 12] - pos_of[ref % 12]);
            const int dist = min(d, 12 - d);
            if (dist <= 2) {
                const float bonus = 0.20f * (1.0f - (float)dist * 0.5f);
                refined[k] = refined[k] + ref_score * bonus;
            }
        }
    }
    for (int k = 0; k < 24; ++k) by_key[k] = refined[k];
    for (int k = 0; k < 24; ++k) {  // stable insertion sort, descending
        const float x = refined[k];
        int j = k;
        while (j > 0 && scores[j - 1] < x) {
            scores[j] = scores[j - 1];
            keys[j] = keys[j - 1];
            --j;
        }
        scores[j] = x;
        keys[j] = k;
    }
}

__device__ __noinline__ float key_clarity(const float* sc, int n) {  // key_clarity.rs:51-93
    if (n < 2) return 0.0f;
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) sum = sum + sc[i];
    const float avg = sum / (float)n;
    float mn = sc[0], mx = sc[0];
    for (int i = 1; i < n; ++i) {
        if (sc[i] < mn) mn = sc[i];
        if (sc[i] >= mx) mx = sc[i];
    }
    const float range = mx - mn;
    if (range > 1e-10f) return clamp_rs((sc[0] - avg) / range, 0.0f, 1.0f);
    return 0.0f;
}

// ---- per-track vote (lib.rs:1353-1436, 1461): one thread per track -----------------------------------
__global__ void key_vote_kernel(TrackDev* tr, const float* fa, int n_tracks, DevCfg cfg, int fallback_only) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    TrackDev& T = tr[t];
    if (fallback_only) {
        if (!T.key_fallback) return;
    } else {
        T.key = 0;
        T.key_confidence = 0.0f;
        T.key_clarity = 0.0f;
        T.key_fallback = 0;
    }
    if (T.status != 0 || T.Fk == 0 || T.m < 2048) return;
    uint32_t seg_len, hop, nseg;
    const bool voting = seg_geometry(T, cfg, &seg_len, &hop, &nseg);
    nseg = min(nseg, T.seg_cap > 0 ? T.seg_cap - 1 : 0u);
    const float* ss = fa + T.seg_scores;
    int keys[24];
    float scores[24], by_key[24];
    bool voted = false;
    int key = 0;
    float confidence = 0.0f;
    float fin[24];
    if (voting && !fallback_only) {
        const float min_cl = clamp_rs(cfg.key_seg_min_clarity, 0.0f, 1.0f);
        float acc[24];
        for (int k = 0; k < 24; ++k) acc[k] = 0.0f;
        uint32_t used = 0;
        for (uint32_t s = 0; s < nseg; ++s) {
            rank_keys(ss + (uint64_t)s * 24, keys, scores, by_key);
            const float cl = key_clarity(scores, 24);
            if (cl >= min_cl) {
                ++used;
                for (int k = 0; k < 24; ++k) acc[k] = acc[k] + by_key[k] * cl;  // one add per key and segment, as lib.rs:1375-1381
            }
        }
        if (used > 0) {
            for (int k = 0; k < 24; ++k) {
                const float x = acc[k];
                int j = k;
                while (j > 0 && fin[j - 1] < x) {
                    fin[j] = fin[j - 1];
                    keys[j] = keys[j - 1];
                    --j;
                }
                fin[j] = x;
                keys[j] = k;
            }
            key = keys[0];
            confidence = fin[0] > 0.0f ? clamp_rs((fin[0] - fin[1]) / fin[0], 0.0f, 1.0f) : 0.0f;
            voted = true;
        }
    }
    if (!voted && voting && !fallback_only) {
        T.key_fallback = 1;  // no segment passed the clarity gate (lib.rs:1385-1411): the whole-track row is computed on demand
        return;
    }
    if (!voted) {
        rank_keys(ss + (uint64_t)nseg * 24, keys, fin, by_key);
        // weighted top-3 vote (detector.rs:254-275): three distinct keys, so the first-ranked key wins
        key = keys[0];
        confidence = fin[0] > 0.0f ? clamp_rs((fin[0] - fin[1]) / fin[0], 0.0f, 1.0f) : 0.0f;
    }
    T.key = key;
    T.key_confidence = confidence;
    T.key_clarity = key_clarity(fin, 24);
}

void launch_key_mask(const WaveCtx& c) {
    if (c.max_Fk > 0 && (c.cfg.key_mask || c.cfg.key_smooth_only)) {
        const dim3 g((KBINS + 127) / 128, c.n_tracks);
        if (c.cfg.key_margin == 12) mask_kernel<12><<<g, 128, 0, c.stream>>>(c.tracks, c.fa, c.cfg);  // default margin (config.rs:669)
        else mask_kernel<0><<<g, 128, 0, c.stream>>>(c.tracks, c.fa, c.cfg);
        count_launch("key_mask");
    }
}

void launch_key_hpcp(const WaveCtx& c) {
    if (c.max_Fk > 0) {
        const dim3 g((c.max_Fk + 3) / 4, c.n_tracks);
        if (c.cfg.key_hpcp) hpcp_kernel<<<g, 128, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa, c.cfg);
        else chroma_fold_kernel<<<g, 128, 0, c.stream>>>(c.tracks, c.srtab, c.sr_index, c.fa);
        count_launch("key_hpcp");
        if (c.cfg.chroma_sharpen > 1.0f) {  // lib.rs:1200-1208
            chroma_sharpen_kernel<<<dim3((c.max_Fk + 255) / 256, c.n_tracks), 256, 0, c.stream>>>(c.tracks, c.fa, c.cfg.chroma_sharpen);
            count_launch("key_hpcp");
        }
    }
}

void launch_key_vote(const WaveCtx& c) {
    if (c.max_Fk > 0) {
        chroma_smooth_kernel<<<dim3((c.max_Fk * 12 + 255) / 256, c.n_tracks), 256, 0, c.stream>>>(c.tracks, c.fa);
        count_launch("key_vote");
        key_weights_kernel<<<c.n_tracks, 256, 0, c.stream>>>(c.tracks, c.fa, c.cfg);
        count_launch("key_vote");
        segment_score_kernel<<<dim3(c.max_seg_cap, c.n_tracks), 768, 0, c.stream>>>(c.tracks, c.fa, c.tab, c.cfg, 0);
        count_launch("key_vote");
    }
    key_vote_kernel<<<(c.n_tracks + 63) / 64, 64, 0, c.stream>>>(c.tracks, c.fa, c.n_tracks, c.cfg, 0);
    count_launch("key_vote");
    if (c.max_Fk > 0 && c.cfg.key_voting) {  // rare path: tracks whose segments were all rejected
        segment_score_kernel<<<dim3(1, c.n_tracks), 768, 0, c.stream>>>(c.tracks, c.fa, c.tab, c.cfg, 1);
        count_launch("key_vote");
        key_vote_kernel<<<(c.n_tracks + 63) / 64, 64, 0, c.stream>>>(c.tracks, c.fa, c.n_tracks, c.cfg, 1);
        count_launch("key_vote");
    }
}

}  // namespace sb
