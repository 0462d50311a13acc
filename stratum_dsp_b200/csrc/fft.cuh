// SFFT arithmetic — the fixed FFT DAG shared with the CPU oracle (oracle/so_fft.cpp documents it).
//   complex multiply  w*f:  re = fma(w.re, f.re, -(w.im*f.im)),  im = fma(w.re, f.im, w.im*f.re)
//   radix-4 DIT butterfly after twiddling: see r4()
// The reference delegates this arithmetic to rustfft 6.2 (chroma/extractor.rs:326-346,
// period/tempogram_fft.rs:149-151); the DAG below is our pinned restatement of the same DFT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

__device__ __forceinline__ float2 cmul(float2 w, float2 f) {
    float2 r;
    float p = __fmul_rn(w.y, f.y);
    r.x = __fmaf_rn(w.x, f.x, -p);
    float q = __fmul_rn(w.y, f.x);
    r.y = __fmaf_rn(w.x, f.y, q);
    return r;
}

// (A,B,C,D) already twiddled -> (X0,X1,X2,X3) written back in place.
__device__ __forceinline__ void r4(float2& A, float2& B, float2& C, float2& D) {
    float2 t0 = make_float2(__fadd_rn(A.x, C.x), __fadd_rn(A.y, C.y));
    float2 t1 = make_float2(__fsub_rn(A.x, C.x), __fsub_rn(A.y, C.y));
    float2 t2 = make_float2(__fadd_rn(B.x, D.x), __fadd_rn(B.y, D.y));
    float2 t3 = make_float2(__fsub_rn(B.x, D.x), __fsub_rn(B.y, D.y));
    A = make_float2(__fadd_rn(t0.x, t2.x), __fadd_rn(t0.y, t2.y));
    B = make_float2(__fadd_rn(t1.x, t3.y), __fsub_rn(t1.y, t3.x));
    C = make_float2(__fsub_rn(t0.x, t2.x), __fsub_rn(t0.y, t2.y));
    D = make_float2(__fsub_rn(t1.x, t3.y), __fadd_rn(t1.y, t3.x));
}

// real-input split for bin k (0..M): a = Z[k mod M], b = Z[(M-k) mod M], w = RW[k]
__device__ __forceinline__ float2 rsplit(float2 a, float2 b, float2 w) {
    float2 E = make_float2(__fadd_rn(a.x, b.x), __fsub_rn(a.y, b.y));
    float2 O = make_float2(__fsub_rn(a.x, b.x), __fadd_rn(a.y, b.y));
    float2 T = cmul(w, O);
    return make_float2(__fmul_rn(0.5f, __fadd_rn(E.x, T.y)), __fmul_rn(0.5f, __fsub_rn(E.y, T.x)));
}

// One CTA-wide Stockham pass schedule over buffers that every thread of the block can see
// (shared or global memory): radix-2 first when log2(M) is odd, then radix-4 passes.  `in` holds
// the input; returns the buffer that holds the result.  tw = TW_M table (M entries).
// tws = table stride: tw may be the TW table of a larger power-of-two size (TW_M[t] == TW_{M*s}[t*s] bit for bit).
__device__ inline float2* cta_cfft(float2* in, float2* out, const float2* __restrict__ tw, uint32_t M, uint32_t tws = 1) {
    uint32_t m = 31 - __clz(M);
    uint32_t Ns = 1;
    if (m & 1) {
        const uint32_t half = M >> 1;
        for (uint32_t j = threadIdx.x; j < half; j += blockDim.x) {
            float2 a = cmul(tw[0], in[j]);
            float2 b = cmul(tw[0], in[j + half]);
            out[2 * j] = make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
            out[2 * j + 1] = make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
        }
        __syncthreads();
        float2* t = in; in = out; out = t;
        Ns = 2;
    }
    while (Ns < M) {
        const uint32_t q = M >> 2;
        const uint32_t tstep = (M / (Ns * 4)) * tws;
        for (uint32_t j = threadIdx.x; j < q; j += blockDim.x) {
            const uint32_t k = j & (Ns - 1);
            float2 A = cmul(tw[0], in[j]);
            float2 B = cmul(tw[k * tstep], in[j + q]);
            float2 C = cmul(tw[2 * k * tstep], in[j + 2 * q]);
            float2 D = cmul(tw[3 * k * tstep], in[j + 3 * q]);
            r4(A, B, C, D);
            const uint32_t o = (j - k) * 4 + k;
            out[o] = A;
            out[o + Ns] = B;
            out[o + 2 * Ns] = C;
            out[o + 3 * Ns] = D;
        }
        __syncthreads();
        float2* t = in; in = out; out = t;
        Ns *= 4;
    }
    return in;
}

}  // namespace sb
