// Row geometry of the key scoring stage, shared by the host planner (engine.cu) and the kernels (k_key.cu).
//
// Every detect_key_* call of the reference's key path works on a contiguous run of chroma frames ("row"):
//   segment voting   lib.rs:1332-1436          rows 0..n-1 = windows of key_segment_len_frames, row n = whole track
//   multi-scale      key/detector.rs:546-700   rows enumerate (scale, window) in the reference's loop order, then whole
//   ensemble         key/detector.rs:881-976   row 0 = whole track with Krumhansl-Kessler, row 1 = whole track with Temperley
// The whole-track row of the first two is the fallback when every window fails the clarity gate.
// All rows are relative to the (optionally edge-trimmed, lib.rs:1216-1233) chroma slice [f0, f0 + nf).
#pragma once
#include "kernels.h"

namespace sb {

enum { KEY_ROWS_VOTE = 0, KEY_ROWS_MULTI_SCALE = 1, KEY_ROWS_ENSEMBLE = 2 };

struct KeyRows {
    int mode;         // KEY_ROWS_*
    uint32_t nseg;    // window rows; the whole-track row has index nseg (ensemble: nseg = 0, rows 0 and 1)
    uint32_t nrows;   // rows in total
};

// lib.rs:1216-1233: slice of the smoothed chroma the scoring works on
__host__ __device__ inline void key_slice(uint32_t n, const DevCfg& cfg, uint32_t* f0, uint32_t* nf) {
    *f0 = 0;
    *nf = n;
    if (cfg.key_edge_trim && n >= 200) {
        float frac = cfg.key_edge_frac;
        frac = frac < 0.0f ? 0.0f : (frac > 0.49f ? 0.49f : frac);
        const float a = roundf((float)n * frac), b = roundf((float)n * (1.0f - frac));
        const uint32_t start = a > 0.0f ? (uint32_t)a : 0u, end = b > 0.0f ? (uint32_t)b : 0u;
        if (end > start + 50 && end <= n) {
            *f0 = start;
            *nf = end - start;
        }
    }
}

__host__ __device__ inline bool ms_scale_live(const DevCfg& cfg, uint32_t si, uint32_t nf, float* weight) {
    const uint32_t len = cfg.ms_len[si];
    if (len == 0 || len > nf) return false;
    const float w = (cfg.ms_nw > 0 && si < cfg.ms_nw) ? cfg.ms_w[si] : 1.0f;  // detector.rs:594-600
    *weight = w;
    return w > 0.0f;
}

__host__ __device__ inline KeyRows key_rows(uint32_t nf, const DevCfg& cfg) {
    KeyRows r;
    if (cfg.key_mode == KEY_ROWS_ENSEMBLE) {  // lib.rs:1290-1298
        r.mode = KEY_ROWS_ENSEMBLE;
        r.nseg = 0;
        r.nrows = 2;
        return r;
    }
    if (cfg.key_mode == KEY_ROWS_MULTI_SCALE && cfg.ms_n > 0) {  // lib.rs:1304-1308
        uint32_t mn = cfg.ms_len[0];
        for (uint32_t i = 1; i < cfg.ms_n; ++i) mn = cfg.ms_len[i] < mn ? cfg.ms_len[i] : mn;
        if (nf >= mn) {
            const uint32_t hop = cfg.ms_hop > 1 ? cfg.ms_hop : 1;
            uint32_t rows = 0;
            for (uint32_t i = 0; i < cfg.ms_n; ++i) {
                float w;
                if (ms_scale_live(cfg, i, nf, &w)) rows += (nf - cfg.ms_len[i]) / hop + 1;
            }
            r.mode = KEY_ROWS_MULTI_SCALE;
            r.nseg = rows;
            r.nrows = rows + 1;
            return r;
        }
    }
    r.mode = KEY_ROWS_VOTE;
    r.nseg = 0;
    const uint32_t sl = cfg.key_seg_len > 1 ? cfg.key_seg_len : 1;
    if (cfg.key_voting && nf >= sl && cfg.key_seg_len >= 120 && cfg.key_seg_hop >= 1) {  // lib.rs:1332-1336
        const uint32_t seg_len = cfg.key_seg_len < nf ? cfg.key_seg_len : nf;
        uint32_t hop = cfg.key_seg_hop < seg_len ? cfg.key_seg_hop : seg_len;
        hop = hop > 1 ? hop : 1;
        r.nseg = (nf - seg_len) / hop + 1;
    }
    r.nrows = r.nseg + 1;
    return r;
}

// Frame range (relative to the slice), scale weight and template set of a row.
__host__ __device__ inline void key_row_range(const KeyRows& R, uint32_t row, uint32_t nf, const DevCfg& cfg, uint32_t* start, uint32_t* len, float* scale_w,
                                              int* tset) {
    *start = 0;
    *len = nf;
    *scale_w = 1.0f;
    *tset = cfg.key_template_set;
    if (R.mode == KEY_ROWS_ENSEMBLE) {
        *tset = row == 0 ? 0 : 1;
        return;
    }
    if (row >= R.nseg) return;  // whole-track row
    if (R.mode == KEY_ROWS_MULTI_SCALE) {
        const uint32_t hop = cfg.ms_hop > 1 ? cfg.ms_hop : 1;
        uint32_t r = row;
        for (uint32_t i = 0; i < cfg.ms_n; ++i) {
            float w;
            if (!ms_scale_live(cfg, i, nf, &w)) continue;
            const uint32_t cnt = (nf - cfg.ms_len[i]) / hop + 1;
            if (r < cnt) {
                *start = r * hop;
                *len = cfg.ms_len[i];
                *scale_w = w;
                return;
            }
            r -= cnt;
        }
        return;
    }
    const uint32_t seg_len = cfg.key_seg_len < nf ? cfg.key_seg_len : nf;
    uint32_t hop = cfg.key_seg_hop < seg_len ? cfg.key_seg_hop : seg_len;
    hop = hop > 1 ? hop : 1;
    *start = row * hop;
    *len = seg_len;
}

}  // namespace sb
