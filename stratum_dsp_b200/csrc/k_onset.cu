// Onset detection kernels: energy flux (reference onset/energy_flux.rs:67-243), per-frame and
// per-frame-pair spectrogram reductions shared by the spectral-flux / HFC detectors
// (onset/spectral_flux.rs:120-157, onset/hfc.rs:132-154) and by the tempogram novelty curves
// (period/novelty.rs:336-836), exact percentile thresholds + peak picking, and the consensus vote
// (onset/consensus.rs:111-287 with the caller-side policy of lib.rs:258-286).
//
// Every sum that feeds a comparison is accumulated in the reference's order (bin 0 upward, one
// rounded multiply and one rounded add per term), one thread per frame / frame pair.
#include "framed.cuh"
#include "kernels.h"

namespace sb {

// frame-region layout (q * fmax + t): 0 rowmax, 1 E, 2 H, 3..5 E_low/mid/high, 6..8 H_low/mid/high, 9.. mel[40]
constexpr int FQ_ROWMAX = 0, FQ_E = 1, FQ_H = 2, FQ_EB = 3, FQ_HB = 6, FQ_MEL = 9;
// pair-region layout: 0 onset spectral flux, 1 SF_full, 2..4 SF_low/mid/high, 5 SF_mel
constexpr int PQ_SFLUX = 0, PQ_SF = 1, PQ_SFB = 2, PQ_MEL = 5;
constexpr int MEL_MAX = 40;

// ---- energy flux onsets ------------------------------------------------------------------------
__global__ void __launch_bounds__(128) energy_rms_kernel(const float* __restrict__ x, const TrackDev* tr, float* fa) {
    __shared__ float tiles[4][32][33];
    const TrackDev& T = tr[blockIdx.y];
    const uint32_t nf = T.F[0];
    const uint32_t f0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 32;
    if (f0 >= nf || T.status != 0) return;
    framed_rms_warp<2048>(x + T.off + T.trim_start, T.m, T.gain, 512, f0, nf, tiles[threadIdx.x >> 5], fa + T.erms);
}

// Shared peak rule of the three detectors (energy_flux.rs:176-217, spectral_flux.rs:183-210, hfc.rs:180-204).
__device__ __forceinline__ bool is_peak(const float* f, uint32_t i, uint32_t L, float thr) {
    if (L < 2) return false;
    float v = f[i];
    if (!(v > thr)) return false;
    if (i == 0) return v >= f[1];
    if (i == L - 1) return v > f[L - 2];
    return v > f[i - 1] && v >= f[i + 1];
}

// Ordered compaction of peak indices of f[0..L) into out (as (i+1)*hop sample positions < m).
__device__ inline uint32_t compact_peaks(const float* f, uint32_t L, float thr, uint32_t hop, uint64_t m, int32_t* out, uint32_t* sc) {
    const uint32_t per = (L + blockDim.x - 1) / blockDim.x;
    const uint32_t a = threadIdx.x * per, b = min(a + per, L);
    uint32_t cnt = 0;
    for (uint32_t i = a; i < b; ++i)
        if (is_peak(f, i, L, thr) && (uint64_t)(i + 1) * hop < m) ++cnt;
    uint32_t total;
    uint32_t pos = block_exclusive_scan(cnt, &total, sc);
    for (uint32_t i = a; i < b; ++i)
        if (is_peak(f, i, L, thr) && (uint64_t)(i + 1) * hop < m) out[pos++] = (int32_t)((i + 1) * hop);
    return total;
}

__global__ void __launch_bounds__(256) energy_onset_kernel(TrackDev* tr, float* fa, int32_t* ia, DevCfg cfg) {
    __shared__ uint32_t sc[34];
    __shared__ float smax[32];
    TrackDev& T = tr[blockIdx.x];
    if (T.status != 0) return;
    const uint32_t nf = T.F[0];
    if (threadIdx.x == 0) {
        T.n_on_energy = 0;
        T.onset_method_consensus = 0.0f;
    }
    if (nf < 2) return;
    const uint32_t L = nf - 1;
    const float* rms = fa + T.erms;
    float* flux = fa + T.scratch;
    float mx = 0.0f;
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
        float v = fmaxf(__fsub_rn(rms[i + 1], rms[i]), 0.0f);
        flux[i] = v;
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, smax);
    __syncthreads();
    if (mx <= 1e-10f) return;
    const float thr = __fmul_rn(mx, cfg.energy_thr_mul);
    int32_t* out = ia + T.on_energy;
    uint32_t total = compact_peaks(flux, L, thr, 512, T.m, out, sc);
    __syncthreads();
    if (threadIdx.x == 0) {  // dedupe within hop/2 (energy_flux.rs:224-238)
        uint32_t w = 0;
        for (uint32_t i = 0; i < total; ++i)
            if (w == 0 || out[i] >= out[w - 1] + 256) out[w++] = out[i];
        T.n_on_energy = w;
        T.onset_method_consensus = w > 0 ? 1.0f : 0.0f;
    }
}

// ---- per-frame reductions over the hop-h spectrogram ------------------------------------------
__global__ void __launch_bounds__(128) frame_feat_kernel(const TrackDev* tr, const int32_t* __restrict__ list, const SrTables* srtab,
                                                         const int32_t* sr_index, int h, float* fa) {
    __shared__ float mel[MEL_MAX][128];
    const int t = list ? list[blockIdx.y] : blockIdx.y;
    const TrackDev& T = tr[t];
    const uint32_t F = T.F[h];
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= F || T.status != 0) return;
    const SrTables& st = srtab[sr_index[t]];
    const HopLayout& HL = T.hop[h];
    if (f >= F) return;
    const float* row = fa + HL.spec + (uint64_t)f * 1025;
    for (int m = 0; m < MEL_MAX; ++m) mel[m][threadIdx.x] = 0.0f;
    float rowmax = 0.0f, E = 0.0f, H = 0.0f;
    float Eb[3] = {0.0f, 0.0f, 0.0f}, Hb[3] = {0.0f, 0.0f, 0.0f};
    const uint32_t e0 = st.b0, e1 = st.b_low, e2 = st.b_mid, e3 = st.b_hi;
    for (uint32_t k = 0; k < 1025; ++k) {
        const float x = row[k];
        rowmax = fmaxf(rowmax, x);
        const float xx = __fmul_rn(x, x);
        const float kx = __fmul_rn(__fmul_rn((float)k, x), x);
        E = __fadd_rn(E, xx);
        H = __fadd_rn(H, kx);
        if (k >= e0 && k < e1) { Eb[0] = __fadd_rn(Eb[0], xx); Hb[0] = __fadd_rn(Hb[0], kx); }
        else if (k >= e1 && k < e2) { Eb[1] = __fadd_rn(Eb[1], xx); Hb[1] = __fadd_rn(Hb[1], kx); }
        else if (k >= e2 && k < e3) { Eb[2] = __fadd_rn(Eb[2], xx); Hb[2] = __fadd_rn(Hb[2], kx); }
        const float v = logf(__fadd_rn(1.0f, fmaxf(x, 0.0f)));  // novelty.rs:180
        if (v > 0.0f) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int m = st.mel_m[k * 2 + c];
                if (m >= 0) mel[m][threadIdx.x] = __fadd_rn(mel[m][threadIdx.x], __fmul_rn(v, st.mel_w[k * 2 + c]));
            }
        }
    }
    float* fr = fa + HL.frame;
    const uint64_t fm = HL.fmax;
    fr[FQ_ROWMAX * fm + f] = rowmax;
    fr[FQ_E * fm + f] = E;
    fr[FQ_H * fm + f] = H;
    for (int b = 0; b < 3; ++b) {
        fr[(FQ_EB + b) * fm + f] = Eb[b];
        fr[(FQ_HB + b) * fm + f] = Hb[b];
    }
    for (uint32_t m = 0; m < st.n_mels; ++m) fr[(FQ_MEL + m) * fm + f] = mel[m][threadIdx.x];
}

// ---- per-pair reductions (frame i -> i+1) -------------------------------------------------------
__global__ void __launch_bounds__(128) pair_feat_kernel(const TrackDev* tr, const int32_t* __restrict__ list, const SrTables* srtab,
                                                        const int32_t* sr_index, int h, float* fa, DevCfg cfg) {
    const int t = list ? list[blockIdx.y] : blockIdx.y;
    const TrackDev& T = tr[t];
    const uint32_t F = T.F[h];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (F < 2 || i >= F - 1 || T.status != 0) return;
    const SrTables& st = srtab[sr_index[t]];
    const HopLayout& HL = T.hop[h];
    const float* prev = fa + HL.spec + (uint64_t)i * 1025;
    const float* cur = prev + 1025;
    const float* fr = fa + HL.frame;
    const uint64_t fm = HL.fmax;
    const float maxp = fr[FQ_ROWMAX * fm + i], maxc = fr[FQ_ROWMAX * fm + i + 1];
    const bool np = maxp > 1e-10f, nc = maxc > 1e-10f;
    const int K = (int)min(max(cfg.sf_k, 1u), 8u);  // superflux radius (novelty.rs:349); the ABI rejects K > 8
    const uint32_t e0 = st.b0, e1 = st.b_low, e2 = st.b_mid, e3 = st.b_hi;
    float sflux = 0.0f, sf = 0.0f, sfb[3] = {0.0f, 0.0f, 0.0f};
    // sliding window of log-compressed prev bins k-8..k+8 (radius K <= 8 selected by masking)
    float win[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) win[j] = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) win[9 + j] = logf(__fadd_rn(1.0f, fmaxf(prev[j], 0.0f)));  // shifted left once before first use
    for (uint32_t k = 0; k < 1025; ++k) {
        const float xc_raw = cur[k], xp_raw = prev[k];
        // spectral_flux.rs:120-157
        const float xc = nc ? __fdiv_rn(xc_raw, maxc) : 0.0f;
        const float xp = np ? __fdiv_rn(xp_raw, maxp) : 0.0f;
        const float d0 = fmaxf(__fsub_rn(xc, xp), 0.0f);
        sflux = __fadd_rn(sflux, __fmul_rn(d0, d0));
#pragma unroll
        for (int j = 0; j < 16; ++j) win[j] = win[j + 1];
        win[16] = (k + 8 < 1025) ? logf(__fadd_rn(1.0f, fmaxf(prev[k + 8], 0.0f))) : 0.0f;
        const float Lc = logf(__fadd_rn(1.0f, fmaxf(xc_raw, 0.0f)));
        // full band: window clipped to [0, 1025) — out-of-range slots hold 0 and logs are >= 0
        float pm = 0.0f;
#pragma unroll
        for (int j = 0; j < 17; ++j)
            if (j >= 8 - K && j <= 8 + K) pm = fmaxf(pm, win[j]);
        float d = fmaxf(__fsub_rn(Lc, pm), 0.0f);
        sf = __fadd_rn(sf, __fmul_rn(d, d));
        // band variants: window additionally clipped to the band (novelty.rs:432-441)
        int b = -1;
        uint32_t lo = 0, hi = 0;
        if (k >= e0 && k < e1) { b = 0; lo = e0; hi = e1; }
        else if (k >= e1 && k < e2) { b = 1; lo = e1; hi = e2; }
        else if (k >= e2 && k < e3) { b = 2; lo = e2; hi = e3; }
        if (b >= 0) {
            float pmb = 0.0f;
#pragma unroll
            for (int j = 0; j < 17; ++j) {
                const int kb = (int)k - 8 + j;
                if (j >= 8 - K && j <= 8 + K && kb >= (int)lo && kb < (int)hi) pmb = fmaxf(pmb, win[j]);
            }
            float db = fmaxf(__fsub_rn(Lc, pmb), 0.0f);
            if (b == 0) sfb[0] = __fadd_rn(sfb[0], __fmul_rn(db, db));
            else if (b == 1) sfb[1] = __fadd_rn(sfb[1], __fmul_rn(db, db));
            else sfb[2] = __fadd_rn(sfb[2], __fmul_rn(db, db));
        }
    }
    float* pr = fa + HL.pair;
    pr[PQ_SFLUX * fm + i] = __fadd_rn(sqrtf(sflux), 0.0f);
    pr[PQ_SF * fm + i] = sqrtf(sf);
    for (int b = 0; b < 3; ++b) pr[(PQ_SFB + b) * fm + i] = sqrtf(sfb[b]);
    // mel superflux (novelty.rs:576-597)
    const int nm = (int)st.n_mels, MK = (int)max(cfg.mel_k, 1u);
    float ms = 0.0f;
    for (int m = 0; m < nm; ++m) {
        float pmx = 0.0f;
        for (int j = max(m - MK, 0); j < min(m + MK + 1, nm); ++j) pmx = fmaxf(pmx, fr[(FQ_MEL + j) * fm + i]);
        float d = fmaxf(__fsub_rn(fr[(FQ_MEL + m) * fm + i + 1], pmx), 0.0f);
        ms = __fadd_rn(ms, __fmul_rn(d, d));
    }
    pr[PQ_MEL * fm + i] = sqrtf(ms);
}

// ---- spectral-flux / HFC onsets: exact percentile threshold + peaks -----------------------------
__global__ void __launch_bounds__(256) spectral_onset_kernel(TrackDev* tr, float* fa, int32_t* ia, DevCfg cfg) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bc[2];
    __shared__ uint32_t sc[34];
    TrackDev& T = tr[blockIdx.x];
    const int which = blockIdx.y;  // 0 spectral flux, 1 HFC
    if (T.status != 0) return;
    if (threadIdx.x == 0) {
        if (which == 0) T.n_on_spectral = 0; else T.n_on_hfc = 0;
    }
    const uint32_t F = T.F[0];
    if (F < 2) return;
    const uint32_t L = F - 1;
    const HopLayout& HL = T.hop[0];
    const uint64_t fm = HL.fmax;
    const float* flux;
    if (which == 0) {
        flux = fa + HL.pair + PQ_SFLUX * fm;
    } else {
        float* hf = fa + T.scratch + fm;  // HFC flux (hfc.rs:151-154)
        const float* H = fa + HL.frame + FQ_H * fm;
        for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) hf[i] = __fadd_rn(fmaxf(__fsub_rn(H[i + 1], H[i]), 0.0f), 0.0f);
        __syncthreads();
        flux = hf;
    }
    uint32_t idx = as_u32(__fmul_rn((float)L, cfg.onset_pct));  // spectral_flux.rs:168
    idx = min(idx, L - 1);
    const float thr = block_select_kth(flux, L, idx, hist, bc);
    int32_t* out = ia + (which == 0 ? T.on_spectral : T.on_hfc);
    uint32_t total = compact_peaks(flux, L, thr, 512, T.m, out, sc);  // frame i+1 -> sample (i+1)*hop, kept if < m (lib.rs:181-190)
    if (threadIdx.x == 0) {
        if (which == 0) T.n_on_spectral = total; else T.n_on_hfc = total;
    }
}

// ---- consensus vote + caller policy ------------------------------------------------------------
__global__ void consensus_kernel(TrackDev* tr, int32_t* ia, int n_tracks, DevCfg cfg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    TrackDev& T = tr[t];
    if (T.status != 0) return;
    const int32_t* L0 = ia + T.on_energy;
    int32_t* fin = ia + T.on_final;
    const uint32_t n0 = T.n_on_energy;
    auto copy_energy = [&]() {
        for (uint32_t i = 0; i < n0; ++i) fin[i] = L0[i];
        T.n_on_final = n0;
    };
    if (!cfg.enable_consensus || T.F[0] == 0) { copy_energy(); return; }
    const int32_t* L1 = ia + T.on_spectral;
    const int32_t* L2 = ia + T.on_hfc;
    const uint32_t n1 = T.n_on_spectral, n2 = T.n_on_hfc;
    if (n0 + n1 + n2 == 0) { T.n_on_final = 0; return; }
    const uint32_t tol = as_u32(__fmul_rn(__fdiv_rn((float)cfg.consensus_tol_ms, 1000.0f), (float)T.sr));  // consensus.rs:149
    // 3-way stable merge (method order 0,1,2 on equal samples) + gap clustering; two passes: first
    // counts clusters voted by >= 2 methods, second writes the chosen set.
    uint32_t strong = 0, clusters = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const bool want_strong = strong > 0;
        uint32_t i0 = 0, i1 = 0, i2 = 0, w = 0;
        bool open = false;
        int64_t last = 0;
        uint64_t sum = 0;
        uint32_t cnt = 0, voted = 0;
        auto close = [&]() {
            const int32_t centre = (int32_t)(sum / cnt);
            const uint32_t vb = __popc(voted);
            if (pass == 0) {
                ++clusters;
                if (vb >= 2) ++strong;
            } else if (!want_strong || vb >= 2) {
                if (w == 0 || fin[w - 1] != centre) fin[w++] = centre;  // sort + dedup (lib.rs:266-271): centres ascend
            }
        };
        while (i0 < n0 || i1 < n1 || i2 < n2) {
            int method = -1;
            int32_t s = 0x7fffffff;
            if (i0 < n0 && L0[i0] < s) { s = L0[i0]; method = 0; }
            if (i1 < n1 && L1[i1] < s) { s = L1[i1]; method = 1; }
            if (i2 < n2 && L2[i2] < s) { s = L2[i2]; method = 2; }
            if (method == 0) ++i0; else if (method == 1) ++i1; else ++i2;
            if (open && (int64_t)s - last > (int64_t)tol) { close(); open = false; }
            if (!open) { open = true; sum = 0; cnt = 0; voted = 0; }
            sum += (uint64_t)s;
            ++cnt;
            voted |= 1u << method;
            last = s;
        }
        if (open) close();
        if (pass == 1) T.n_on_final = w;
    }
    if (T.n_on_final == 0) copy_energy();  // "Onset consensus produced no candidates" (lib.rs:283-285)
}

void launch_energy_onsets(const WaveCtx& c) {
    if (c.max_F[0] > 0) {
        energy_rms_kernel<<<dim3((c.max_F[0] + 127) / 128, c.n_tracks), 128, 0, c.stream>>>(c.samples, c.tracks, c.fa);
        count_launch("onsets");
    }
    energy_onset_kernel<<<c.n_tracks, 256, 0, c.stream>>>(c.tracks, c.fa, c.ia, c.cfg);
    count_launch("onsets");
}

void launch_spec_features(const WaveCtx& c, int h, const int32_t* d_list, int n_list) {
    if (c.max_F[h] == 0 || n_list == 0) return;
    dim3 grid((c.max_F[h] + 127) / 128, n_list);
    frame_feat_kernel<<<grid, 128, 0, c.stream>>>(c.tracks, d_list, c.srtab, c.sr_index, h, c.fa);
    count_launch("spec_features");
    pair_feat_kernel<<<grid, 128, 0, c.stream>>>(c.tracks, d_list, c.srtab, c.sr_index, h, c.fa, c.cfg);
    count_launch("spec_features");
}

void launch_spectral_onsets_consensus(const WaveCtx& c) {
    if (c.cfg.enable_consensus && c.max_F[0] > 0) {
        spectral_onset_kernel<<<dim3(c.n_tracks, 2), 256, 0, c.stream>>>(c.tracks, c.fa, c.ia, c.cfg);
        count_launch("onsets");
    }
    consensus_kernel<<<(c.n_tracks + 63) / 64, 64, 0, c.stream>>>(c.tracks, c.ia, c.n_tracks, c.cfg);
    count_launch("onsets");
}

}  // namespace sb
