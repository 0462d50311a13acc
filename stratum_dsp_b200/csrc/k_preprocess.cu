// Preprocessing kernels: peak normalisation gain (reference preprocessing/normalization.rs:262-322)
// and silence detection + trim (preprocessing/silence.rs:102-279).
//
// Data layout: tracks are concatenated in one f32 buffer; the gain is never written back — every
// consumer multiplies on load with one rounding (`x*g`), which is bit-identical to the reference's
// in-place `*sample *= gain` followed by a read.
#include "framed.cuh"
#include "kernels.h"

namespace sb {

// ---- peak: order-free exact max|x| per track ------------------------------------------------
__global__ void __launch_bounds__(256) peak_kernel(const float* __restrict__ x, TrackDev* tr) {
    const int t = blockIdx.y;
    const uint64_t n = tr[t].n;
    const float* p = x + tr[t].off;
    float m = 0.0f;
    // float4 body when the track start is 16-byte aligned, scalar otherwise
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((((uintptr_t)p) & 15) == 0) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const uint64_t n4 = n >> 2;
        for (uint64_t j = i; j < n4; j += stride) {
            float4 v = __ldg(p4 + j);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        for (uint64_t j = (n4 << 2) + i; j < n; j += stride) m = fmaxf(m, fabsf(p[j]));
    } else {
        for (uint64_t j = i; j < n; j += stride) m = fmaxf(m, fabsf(p[j]));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0f;
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(&tr[t].peak), __float_as_uint(m));  // m >= 0: bit order == value order
    }
}

// normalize_peak — normalization.rs:275-295
__global__ void gain_kernel(TrackDev* tr, int n_tracks, DevCfg cfg) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    float g = 1.0f;
    if (cfg.enable_normalization) {
        float peak = tr[t].peak;
        if (peak > 1e-10f) {
            g = cfg.target_peak / peak;
            g = fminf(g, 1.0f / peak);
        }
    }
    tr[t].gain = g;
}

__global__ void __launch_bounds__(128) silence_rms_kernel(const float* __restrict__ x, TrackDev* tr, float* fa) {
    __shared__ float tiles[4][32][33];
    const int t = blockIdx.y;
    const TrackDev& T = tr[t];
    const uint32_t nf = T.Fsil;
    const uint32_t f0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 32;
    if (f0 >= nf) return;
    framed_rms_warp<2048>(x + T.off, T.n, T.gain, 1024, f0, nf, tiles[threadIdx.x >> 5], fa + T.sil_rms);
}

// detect_and_trim region logic — silence.rs:171-256.  One thread per track.
__global__ void trim_kernel(TrackDev* tr, const float* fa, int n_tracks, DevCfg cfg) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    TrackDev& T = tr[t];
    const uint64_t n = T.n;
    uint64_t ts = 0, te = n;
    if (cfg.enable_trim) {
        const float* rms = fa + T.sil_rms;
        const uint32_t nf = T.Fsil;
        const uint32_t hop = 1024;
        uint32_t min_samples = as_u32(__fmul_rn(__fdiv_rn((float)cfg.silence_min_ms, 1000.0f), (float)T.sr));
        uint32_t min_frames = (min_samples + hop - 1) / hop;
        bool in_sil = false, have_first = false;
        uint32_t sil_start = 0;
        uint64_t first_a = 0, first_b = 0, last_a = 0, last_b = 0;
        for (uint32_t f = 0; f < nf; ++f) {
            bool s = rms[f] <= cfg.silence_thr_linear;
            if (s && !in_sil) {
                in_sil = true;
                sil_start = f;
            } else if (!s && in_sil) {
                in_sil = false;
                if (f - sil_start >= min_frames || sil_start == 0) {
                    uint64_t a = (uint64_t)sil_start * hop, b = (uint64_t)f * hop;
                    if (!have_first) { first_a = a; first_b = b; have_first = true; }
                    last_a = a; last_b = b;
                }
            }
        }
        if (in_sil && (nf - sil_start >= min_frames || sil_start == 0)) {
            uint64_t a = (uint64_t)sil_start * hop, b = n;
            if (!have_first) { first_a = a; first_b = b; have_first = true; }
            last_a = a; last_b = b;
        }
        if (have_first && first_a == 0) ts = first_b;
        if (have_first && last_b == n) te = last_a;
        if (ts > te) ts = te;  // trim_start.min(trim_end); trim_end.max(trim_start)
        if (!(ts < te && te <= n)) { ts = 0; te = 0; }
    }
    T.trim_start = ts;
    T.trim_end = te;
    const uint64_t m = te - ts;
    T.m = m;
    const uint32_t hops[N_HOPS] = {512, 256, 1024};
    for (int h = 0; h < N_HOPS; ++h) T.F[h] = m >= 2048 ? (uint32_t)((m - 2048) / hops[h] + 1) : 0;
    T.Fk = m >= 8192 ? (uint32_t)((m - 8192) / 512 + 1) : 0;
    if (m == 0 && T.status == 0) {
        T.status = STRATUM_PROCESSING_ERROR;
        T.err_code = 3;  // "Audio is entirely silent after trimming" (lib.rs:143-147)
    }
}

void launch_peak_gain(const WaveCtx& c) {
    if (c.cfg.enable_normalization) {
        uint64_t per = (c.max_n + 255) / 256;
        unsigned gx = (unsigned)((per + 63) / 64);  // ~64 float4-less iterations per thread
        if (gx < 1) gx = 1;
        if (gx > 2048) gx = 2048;
        peak_kernel<<<dim3(gx, c.n_tracks), 256, 0, c.stream>>>(c.samples, c.tracks);
        count_launch("preprocess");
    }
    gain_kernel<<<(c.n_tracks + 127) / 128, 128, 0, c.stream>>>(c.tracks, c.n_tracks, c.cfg);
    count_launch("preprocess");
}

void launch_silence_trim(const WaveCtx& c) {
    if (c.cfg.enable_trim && c.max_Fsil > 0) {
        silence_rms_kernel<<<dim3((c.max_Fsil + 127) / 128, c.n_tracks), 128, 0, c.stream>>>(c.samples, c.tracks, c.fa);
        count_launch("preprocess");
    }
    trim_kernel<<<(c.n_tracks + 127) / 128, 128, 0, c.stream>>>(c.tracks, c.fa, c.n_tracks, c.cfg);
    count_launch("preprocess");
}

}  // namespace sb
