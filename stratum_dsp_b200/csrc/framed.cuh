// Framed RMS with strict left-to-right f32 accumulation, and small block-level helpers.
#pragma once
#include "common.cuh"

namespace sb {

// ---- framed RMS with the reference's strict left-to-right f32 sum -----------------------------
// One lane per frame; a warp transposes 32x32 sample tiles through shared memory so that global
// loads stay coalesced while every lane adds its own frame's squares in sample order
// (silence.rs:159, onset/energy_flux.rs:127).
template <int FRAME>
__device__ __forceinline__ void framed_rms_warp(const float* __restrict__ x, uint64_t limit, float g, uint32_t hop, uint32_t f0, uint32_t nf,
                                                float (*tile)[33], float* out) {
    const int lane = threadIdx.x & 31;
    const uint32_t f = f0 + lane;
    float sum = 0.0f;
    for (uint32_t jb = 0; jb < FRAME; jb += 32) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            uint64_t idx = (uint64_t)(f0 + r) * hop + jb + lane;
            tile[r][lane] = (f0 + r < nf && idx < limit) ? __fmul_rn(x[idx], g) : 0.0f;
        }
        __syncwarp();
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            float v = tile[lane][j];
            sum = __fadd_rn(sum, __fmul_rn(v, v));
        }
        __syncwarp();
    }
    if (f < nf) {
        uint64_t st = (uint64_t)f * hop;
        uint64_t en = st + FRAME < limit ? st + FRAME : limit;
        // frames are always full except the single short frame of a track shorter than FRAME, where
        // the zero padding above adds exact zeros (x + 0*0 == x) and the divisor is the true length.
        float len = (float)(en - st);
        out[f] = en > st ? sqrtf(__fdiv_rn(sum, len)) : 0.0f;
    }
}


// Block-wide exclusive scan of one uint per thread (blockDim.x <= 1024, multiple of 32).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* warp_sums /* >= 33 */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        uint32_t s = lane < nw ? warp_sums[lane] : 0;
        uint32_t si = s;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += y;
        }
        if (lane < nw) warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    uint32_t res = warp_sums[w] + inc - v;
    *total = warp_sums[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ float block_max(float v, float* sm /* >= 32 */) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float r = (threadIdx.x & 31) < nw ? sm[threadIdx.x & 31] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    __syncthreads();
    return r;
}

// Exact k-th smallest (0-based) of n non-negative floats: 4-pass MSB radix select on the bit
// patterns (order-preserving for x >= 0).  All threads of the block call it; result broadcast.
__device__ inline float block_select_kth(const float* __restrict__ v, uint32_t n, uint32_t k, uint32_t* hist /* 256 */, uint32_t* bcast /* 2 */) {
    uint32_t prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t b = __float_as_uint(v[i]);
            if ((b & mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0, d = 0;
            for (; d < 256; ++d) {
                if (acc + hist[d] > k) break;
                acc += hist[d];
            }
            bcast[0] = d;
            bcast[1] = k - acc;
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= 255u << shift;
        k = bcast[1];
        __syncthreads();
    }
    return __uint_as_float(prefix);
}

}  // namespace sb
