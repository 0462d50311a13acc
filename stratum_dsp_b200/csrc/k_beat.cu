// Beat tracking kernels — reference features/beat_tracking/{hmm,tempo_variation,bayesian,time_signature,mod}.rs
// as driven by generate_beat_grid (beat_tracking/mod.rs:108-247).
//
// One warp per track.  The HMM emission pass is lane-parallel over beat frames (binary search in the
// sorted onset list instead of the reference's O(T x onsets) scan, hmm.rs:273-279 — the minimum
// distance is the same number).  The Viterbi recursion (hmm.rs:308-375) keeps the five tempo states
// in lanes 0..4: probability-domain max-times with strict `>` (ties -> lowest state), argmax by
// shuffle, back-pointers packed 3 bits per state into one word per frame, and a device-side
// backtrace.  The emission is state-independent in the reference (hmm.rs:265-270), so the decoded
// path never changes the beat list; it is still computed and exposed for parity checks.
// Tempo-variation segmentation, the Bayesian per-segment refinement, time signature, downbeats and
// grid stability are short scalar recurrences executed by lane 0 in the reference's order.
#include <algorithm>

#include "kernels.h"

namespace sb {

constexpr float TIMING_TOL_S = 0.05f;
constexpr float EMISSION_SIGMA = TIMING_TOL_S / 2.0f;

// distance to the nearest onset: min over the two neighbours found by binary search
__device__ __forceinline__ float nearest_dist(const float* __restrict__ on, int n, float t) {
    int lo = 0, hi = n;  // first index with on[idx] >= t
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (on[mid] < t) lo = mid + 1; else hi = mid;
    }
    float md = INFINITY;
    if (lo < n) md = fabsf(on[lo] - t);
    if (lo > 0) md = fminf(md, fabsf(on[lo - 1] - t));
    return md;
}

// HmmBeatTracker::track_beats on onsets on[0..n): emissions -> em[0..T), Viterbi -> bp/path (optional),
// kept beats appended to times[*count..] (and frames[] when non-null).  Warp-collective.
// Returns false when the call would have failed in the reference (hmm.rs:121-160).
__device__ inline bool warp_hmm(float bpm, const float* __restrict__ on, int n, float* em, uint32_t T_cap, uint32_t* bp, int32_t* path, float* times,
                                int32_t* frames, uint32_t cap, uint32_t* count, uint32_t* T_out) {
    const int lane = threadIdx.x & 31;
    if (bpm <= 1e-10f || bpm > 300.0f || n <= 0) return false;
    const float start = on[0], end = on[n - 1];
    const float interval = 60.0f / bpm;
    uint32_t T = as_u32(ceilf((end - start) / interval)) + 1;  // hmm.rs:248
    if (T == 0) return false;
    if (T > T_cap) T = T_cap;  // cannot happen for bpm <= 300 with the engine's capacity rule
    if (T_out) *T_out = T;
    const float sigma_sq = EMISSION_SIGMA * EMISSION_SIGMA;
    for (uint32_t t = lane; t < T; t += 32) {
        const float ft = start + ((float)t * interval);
        const float md = nearest_dist(on, n, ft);
        em[t] = expf(-(md * md) / (2.0f * sigma_sq));
    }
    __syncwarp();
    // Viterbi forward pass: lane s < 5 owns state s
    {
        // transition matrix rows normalised as in hmm.rs:184-219: A[p][s]
        float Acol[5];  // A[p][lane]
        for (int p = 0; p < 5; ++p) {
            float row[5], sum = 0.0f;
            for (int j = 0; j < 5; ++j) {
                const int d = abs(p - j);
                row[j] = d == 0 ? 0.7f : (d == 1 ? 0.15f : 0.0f);
                sum = sum + row[j];
            }
            const int s = lane < 5 ? lane : 0;
            Acol[p] = sum > 1e-10f ? row[s] / sum : row[s];
        }
        float V = (1.0f / 5.0f) * em[0];
        for (uint32_t t = 1; t < T; ++t) {
            float bestp = 0.0f;
            int bprev = 0;
#pragma unroll
            for (int p = 0; p < 5; ++p) {
                const float vp = __shfl_sync(0xffffffffu, V, p);
                const float pr = vp * Acol[p];
                if (pr > bestp) {
                    bestp = pr;
                    bprev = p;
                }
            }
            V = bestp * em[t];
            if (bp) {
                uint32_t word = lane < 5 ? ((uint32_t)bprev << (3 * lane)) : 0u;
                for (int o = 4; o > 0; o >>= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
                if (lane == 0) bp[t] = word;
            }
        }
        if (path) {
            // final argmax with strict `>` from state 0 upward (hmm.rs:345-356), then backtrace
            float bf = 0.0f;
            int bs = 0;
            for (int s = 0; s < 5; ++s) {
                const float vs = __shfl_sync(0xffffffffu, V, s);
                if (vs > bf) {
                    bf = vs;
                    bs = s;
                }
            }
            __syncwarp();
            if (lane == 0) {
                path[T - 1] = bs;
                for (uint32_t t = T - 1; t-- > 0;) {
                    bs = (int)((bp[t + 1] >> (3 * bs)) & 7u);
                    path[t] = bs;
                }
            }
        }
    }
    __syncwarp();
    // extract_beats_from_path (hmm.rs:383-441): keep frame t iff emission > 0.1; ordered compaction
    uint32_t base = *count;
    for (uint32_t t0 = 0; t0 < T; t0 += 32) {
        const uint32_t t = t0 + lane;
        const bool keep = t < T && em[t] > 0.1f;
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const uint32_t pos = base + __popc(mask & ((1u << lane) - 1u));
            if (pos < cap) {
                times[pos] = start + ((float)t * interval);
                if (frames) frames[pos] = (int32_t)t;
            }
        }
        base += __popc(mask);
    }
    __syncwarp();
    *count = min(base, cap);
    return true;
}

// BayesianBeatTracker::update_with_onsets (bayesian.rs:104-255): first strict maximum over cur-5..cur+5 step 0.5.
// Warp-collective: lane i evaluates candidate i (at most 21) over the onsets in order; the winner is then
// picked by scanning the lanes in candidate order, exactly like the reference's loop.
__device__ inline float bayes_update(float current_bpm, const float* on, int n) {
    const int lane = threadIdx.x & 31;
    const float lo = fmaxf(current_bpm - 5.0f, 60.0f), hi = fminf(current_bpm + 5.0f, 180.0f);
    const float sigma_sq = 0.05f * 0.05f;
    float b = lo;
    for (int q = 0; q < lane; ++q) b = b + 0.5f;  // the reference accumulates `bpm += 0.5` in f32
    const bool valid = b <= hi;
    float l = 0.0f;
    if (valid && n > 0) {
        const float interval = 60.0f / b, start = on[0];
        float ll = 0.0f;
        for (int i = 0; i < n; ++i) {
            const float o = on[i];
            const int bi = as_i32(roundf((o - start) / interval));
            const float eb = start + ((float)bi * interval);
            const float d = fabsf(o - eb);
            ll = ll + (-(d * d) / (2.0f * sigma_sq));
        }
        l = expf(ll / (float)n);
    }
    float best_bpm = current_bpm, best_l = 0.0f;
    for (int q = 0; q < 32; ++q) {
        const float lq = __shfl_sync(0xffffffffu, l, q);
        const float bq = __shfl_sync(0xffffffffu, b, q);
        const bool vq = __shfl_sync(0xffffffffu, (int)valid, q) != 0;
        if (vq && lq > best_l) {
            best_l = lq;
            best_bpm = bq;
        }
    }
    return best_bpm;
}

__device__ inline void insertion_sort_f(float* a, uint32_t n) {  // stable; input is nearly sorted
    for (uint32_t i = 1; i < n; ++i) {
        const float x = a[i];
        uint32_t j = i;
        while (j > 0 && a[j - 1] > x) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = x;
    }
}

// time_signature.rs:90-199 on a beat-time list; iv = scratch for the positive intervals
__device__ inline int detect_time_signature(const float* beats, uint32_t n, float* iv) {
    if (n < 8) return 4;
    uint32_t cnt = 0;
    for (uint32_t i = 1; i < n; ++i) {
        const float d = beats[i] - beats[i - 1];
        if (d > 0.0f) iv[cnt++] = d;
    }
    if (cnt == 0) return 4;
    float sum = 0.0f;
    for (uint32_t i = 0; i < cnt; ++i) sum = sum + iv[i];
    const float mean = sum / (float)cnt;
    float vs = 0.0f;
    for (uint32_t i = 0; i < cnt; ++i) {
        const float df = iv[i] - mean;
        vs = vs + df * df;
    }
    const float var = vs / (float)cnt;
    const float cv = mean > 1e-10f ? sqrtf(var) / mean : 1.0f;
    float sc[3];
    const uint32_t lags[3] = {4, 3, 6};
    for (int q = 0; q < 3; ++q) {
        const uint32_t lag = lags[q];
        sc[q] = 0.0f;
        if (cnt < lag) continue;
        float acc = 0.0f;
        uint32_t c2 = 0;
        for (uint32_t i = 0; i + lag < cnt; ++i) {
            const float diff = fabsf(iv[i] - iv[i + lag]);
            acc = acc + 1.0f / (1.0f + diff / mean);
            ++c2;
        }
        if (c2 == 0) continue;
        const float ac = acc / (float)c2;
        sc[q] = fminf(ac * 0.7f + (1.0f / (1.0f + cv)) * 0.3f, 1.0f);
    }
    int best = 4;
    float bsc = sc[0];
    if (sc[1] >= bsc) {  // max_by keeps the last maximal element (order 4/4, 3/4, 6/8)
        bsc = sc[1];
        best = 3;
    }
    if (sc[2] >= bsc) {
        bsc = sc[2];
        best = 6;
    }
    return best;
}

// One warp (= one CTA) per track.  The working lists (emissions, first-pass beats, refined beats, interval
// scratch) live in dynamic shared memory when they fit (always for tracks up to ~10 minutes), otherwise
// in the track's arena areas; the scalar post-processing on lane 0 is dominated by access latency.
constexpr uint32_t BEAT_SMEM_FLOATS = 24 * 1024;  // 96 KB

__global__ void __launch_bounds__(32) beat_kernel(TrackDev* tr, float* fa, float* oa, int32_t* ia, int n_tracks, uint32_t smem_floats) {
    extern __shared__ float bsm[];
    const int t = blockIdx.x;
    const int lane = threadIdx.x & 31;
    if (t >= n_tracks) return;
    TrackDev& T = tr[t];
    if (lane == 0) {
        T.n_beats = 0;
        T.n_downbeats = 0;
        T.n_hmm_frames = 0;
        T.grid_stability = 0.0f;
        T.time_sig = 4;
        T.beats_refined = 0;
        T.hmm_T = 0;
    }
    __syncwarp();
    if (T.status != 0) return;
    const float bpm = T.bpm;
    const uint32_t n_on = T.n_on_final;
    if (!(bpm > 0.0f) || n_on < 2) return;  // lib.rs:913
    if (bpm > 300.0f) return;                // generate_beat_grid rejects; analyze_audio degrades to an empty grid
    // onsets in seconds (lib.rs:915-919); already ascending
    const int32_t* on_i = ia + T.on_final;
    float* on = fa + T.onsets_s;
    for (uint32_t i = lane; i < n_on; i += 32) on[i] = (float)on_i[i] / (float)T.sr;
    __syncwarp();
    const bool in_smem = (uint64_t)3 * T.beat_cap + T.hmm_cap <= smem_floats;
    float* em = in_smem ? bsm + 3 * T.beat_cap : fa + T.hmm_em;
    float* pos = in_smem ? bsm : fa + T.beats_tmp;  // beat_positions of the first pass
    float* ref = pos + T.beat_cap;   // refined list; a third beat_cap-sized block holds the interval scratch
    float* beats = oa + T.beats;
    float* down = oa + T.downbeats;
    uint32_t n_pos = 0, hmm_T = 0;
    if (!warp_hmm(bpm, on, (int)n_on, em, T.hmm_cap, reinterpret_cast<uint32_t*>(ia + T.hmm_bp), ia + T.hmm_path, pos, ia + T.hmm_frames, T.beat_cap, &n_pos,
                  &hmm_T))
        return;
    if (lane == 0) {
        T.n_hmm_frames = n_pos;
        T.hmm_T = hmm_T;
    }
    if (n_pos == 0) return;  // "HMM beat tracking produced no beats" -> empty grid
    // ---- detect_tempo_variations (tempo_variation.rs:95-220) + refinement (mod.rs:160-219) ----
    // Segment boundaries are a scalar recurrence; every lane evaluates it redundantly so that the
    // warp-collective HMM of a variable segment can be called uniformly.
    uint32_t n_ref = 0;
    bool any_var = false, refined_ok = false;
    const float first = pos[0], last = pos[n_pos - 1];
    const float total = last - first;
    if (n_pos >= 4 && total >= 2.0f) {
        const float seg_dur = clamp_rs(total / 4.0f, 4.0f, 8.0f);
        const float overlap = seg_dur * 0.5f;
        // pass 1: does any segment vary?  pass 2: build the refined list
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && !any_var) break;
            float cur = first;
            float bayes_bpm = bpm;
            uint32_t nseg = 0;
            uint32_t lo_idx = 0;  // first beat index with time >= cur (monotone)
            while (cur < last) {
                const float seg_end = fminf(cur + seg_dur, last);
                while (lo_idx < n_pos && pos[lo_idx] < cur) ++lo_idx;
                uint32_t hi_idx = lo_idx;
                while (hi_idx < n_pos && pos[hi_idx] <= seg_end) ++hi_idx;
                const uint32_t nb = hi_idx - lo_idx;
                if (nb >= 3) {
                    float sum = 0.0f;
                    uint32_t cnt = 0;
                    for (uint32_t i = lo_idx + 1; i < hi_idx; ++i) {
                        const float d = pos[i] - pos[i - 1];
                        if (d > 0.0f) {
                            sum = sum + d;
                            ++cnt;
                        }
                    }
                    if (cnt > 0) {
                        const float mean = sum / (float)cnt;
                        float vs = 0.0f;
                        for (uint32_t i = lo_idx + 1; i < hi_idx; ++i) {
                            const float d = pos[i] - pos[i - 1];
                            if (d > 0.0f) {
                                const float df = d - mean;
                                vs = vs + df * df;
                            }
                        }
                        const float var = vs / (float)cnt;
                        const float sd = sqrtf(var);
                        const float cv = mean > 1e-10f ? sd / mean : 0.0f;
                        const bool variable = cv > 0.15f;
                        ++nseg;
                        if (pass == 0) {
                            any_var |= variable;
                        } else if (variable) {
                            // onsets inside [cur, seg_end] are a contiguous range of the sorted list
                            int a = 0, b = (int)n_on;
                            {
                                int l = 0, h = (int)n_on;
                                while (l < h) { const int mid = (l + h) >> 1; if (on[mid] < cur) l = mid + 1; else h = mid; }
                                a = l;
                                l = a; h = (int)n_on;
                                while (l < h) { const int mid = (l + h) >> 1; if (on[mid] <= seg_end) l = mid + 1; else h = mid; }
                                b = l;
                            }
                            if (b > a) {
                                bayes_bpm = bayes_update(bayes_bpm, on + a, b - a);
                                uint32_t cnt2 = n_ref;
                                warp_hmm(bayes_bpm, on + a, b - a, em, T.hmm_cap, nullptr, nullptr, ref, nullptr, T.beat_cap, &cnt2, nullptr);
                                n_ref = cnt2;
                            }
                        } else {
                            for (uint32_t i = lo_idx + lane; i < hi_idx; i += 32)
                                if (n_ref + (i - lo_idx) < T.beat_cap) ref[n_ref + (i - lo_idx)] = pos[i];
                            n_ref = min(n_ref + nb, T.beat_cap);
                            __syncwarp();
                        }
                    }
                }
                cur = cur + (seg_dur - overlap);
            }
            (void)nseg;
        }
        refined_ok = any_var && n_ref > 0;
    }
    __syncwarp();
    if (lane != 0) return;
    float* final_pos = pos;
    uint32_t n_final = n_pos;
    if (refined_ok) {
        insertion_sort_f(ref, n_ref);  // stable sort by time (mod.rs:211-216)
        final_pos = ref;
        n_final = n_ref;
        T.beats_refined = 1;
    }
    const int bpb = detect_time_signature(final_pos, n_final, ref + T.beat_cap);
    T.time_sig = bpb;
    // grid (mod.rs:293-321, 363-404): beats sorted (already), downbeats by bar interval
    for (uint32_t i = 0; i < n_final; ++i) beats[i] = final_pos[i];
    T.n_beats = n_final;
    {
        const float beat_iv = 60.0f / bpm;
        const float bar_iv = beat_iv * (float)bpb;
        const float tol = bar_iv * 0.1f;
        uint32_t nd = 0;
        down[nd++] = beats[0];
        for (uint32_t i = 1; i < n_final; ++i) {
            const float expected = down[nd - 1] + bar_iv;
            if (fabsf(beats[i] - expected) <= tol) down[nd++] = beats[i];
        }
        T.n_downbeats = nd;
    }
    // calculate_grid_stability (mod.rs:425-485)
    float stab = 0.0f;
    if (n_final >= 2) {
        float sum = 0.0f;
        uint32_t cnt = 0;
        for (uint32_t i = 1; i < n_final; ++i) {
            const float d = final_pos[i] - final_pos[i - 1];
            if (d > 0.0f) {
                sum = sum + d;
                ++cnt;
            }
        }
        if (cnt > 0) {
            const float mean = sum / (float)cnt;
            if (mean > 1e-10f) {
                float vs = 0.0f;
                for (uint32_t i = 1; i < n_final; ++i) {
                    const float d = final_pos[i] - final_pos[i - 1];
                    if (d > 0.0f) {
                        const float df = d - mean;
                        vs = vs + df * df;
                    }
                }
                const float var = vs / (float)cnt;
                const float cv = sqrtf(var) / mean;
                stab = 1.0f / (1.0f + cv);
            }
        }
    }
    T.grid_stability = stab;
}

void launch_beat_tracking(const WaveCtx& c) {
    static bool attr_dev[64] = {};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_dev[dev & 63]) {
        cudaFuncSetAttribute(beat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BEAT_SMEM_FLOATS * sizeof(float));
        attr_dev[dev & 63] = true;
    }
    // shared memory sized for the largest track of the wave (capped): short tracks leave room for more CTAs per SM
    const uint64_t want = (uint64_t)3 * c.max_beat_cap + (c.max_beat_cap - 64) / 3 + 16;
    const size_t smem = (size_t)std::min<uint64_t>(want, BEAT_SMEM_FLOATS) * sizeof(float);
    beat_kernel<<<c.n_tracks, 32, smem, c.stream>>>(c.tracks, c.fa, c.oa, c.ia, c.n_tracks, (uint32_t)(smem / sizeof(float)));
    count_launch("beats");
}

}  // namespace sb
