// ORACLE — TEST INFRASTRUCTURE ONLY (see so_common.hpp).
//
// FFT used wherever the reference calls rustfft 6.2 (third-party, absent from
// /root/reference: Cargo.toml:18; call sites chroma/extractor.rs:326-346,
// period/tempogram_fft.rs:149-151, period/autocorrelation.rs:240-251).  rustfft
// computes an unnormalised forward DFT  X[k] = sum_n x[n] e^{-2 pi i k n / N}
// in f32 with f64-derived twiddles; its butterfly order (and therefore its
// rounding) depends on the planner and the host SIMD level and is not
// reproducible.  FFT-level parity is therefore UNPINNED; this file restates the
// published DFT with a fixed, documented arithmetic DAG ("SFFT") so that the
// CUDA kernels can reproduce it bit for bit:
//
//   * complex size M = 2^m: Stockham autosort, decimation in time; one radix-2
//     pass first when m is odd, then radix-4 passes (sub-transform size Ns
//     grows 1/2 -> 4Ns each pass).  In the pass with sub-size Ns, butterfly j
//     (0 <= j < M/R) reads in[j + r*M/R], multiplies input r by the twiddle
//     TW[(k*r) * M/(Ns*R)] (k = j mod Ns; TW[t] = (float)cos(2 pi t/M),
//     (float)(-sin(2 pi t/M)) evaluated in double), applies the R-point
//     butterfly and writes out[(j-k)*R + k + r*Ns].
//   * complex multiply w*f:  re = fma(w.re, f.re, -(w.im*f.im)),
//                            im = fma(w.re, f.im,   w.im*f.re)
//     (one rounded product + one fused multiply-add per component).
//   * radix-4 butterfly on (A,B,C,D) after twiddling:
//       t0=A+C t1=A-C t2=B+D t3=B-D;  X0=t0+t2  X2=t0-t2
//       X1=(t1.re+t3.im, t1.im-t3.re)  X3=(t1.re-t3.im, t1.im+t3.re)
//   * real input of even length N: pack z[n] = x[2n] + i x[2n+1], Z = cfft(z)
//     (M = N/2) and for k = 0..M with a = Z[k mod M], b = Z[(M-k) mod M]:
//       E=(a.re+b.re, a.im-b.im)  O=(a.re-b.re, a.im+b.im)
//       T = RW[k]*O   (RW[k] = (float)cos(2 pi k/N), (float)(-sin(2 pi k/N)))
//       X[k] = (0.5*(E.re+T.im), 0.5*(E.im-T.re))
// tests/test_oracle_fft.py checks it against a float64 DFT.
//
// ALTERNATIVE ARITHMETICS (robustness only, never the parity arithmetic): set_fft_variant(v) swaps the DAG above for
//   1 = the same DFT evaluated entirely in float64 (radix-2, float64 twiddles) and rounded ONCE to f32 — the correctly
//       rounded spectrum up to double precision, i.e. what any faithful f32 FFT (rustfft included) scatters around;
//   2 = a textbook f32 radix-2 decimation-in-time FFT with bit reversal, twiddles rounded from float64, plain complex
//       multiply (4 products, 2 sums, no fma), on the FULL complex frame (imaginary part 0) as the reference's call
//       sites do (extractor.rs:326-346) — a different butterfly order, operand packing and rounding pattern.
// tests/test_fft_robustness.py and tools/fft_robustness.py run the whole analysis under each and compare every discrete
// output with the SFFT run: that measures how much of "bit-exact vs the reference" survives an FFT with other roundings.
#include <atomic>
#include <cmath>
#include <complex>

#include "so_common.hpp"

namespace so {

static std::atomic<int> g_fft_variant{0};
void set_fft_variant(int v) { g_fft_variant.store(v); }
int fft_variant() { return g_fft_variant.load(); }

static inline size_t bitrev(size_t i, unsigned bits) {
    size_t r = 0;
    for (unsigned b = 0; b < bits; ++b) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

// exp(-2 pi i t / n) in float64, t < n/2, cached per size and thread (the stage of length len uses entries k * n/len:
// k/len and (k*n/len)/n are the same rational, so the table entry equals the directly evaluated twiddle bit for bit)
static const std::vector<std::complex<double>>& twiddle_f64(size_t n) {
    static thread_local std::map<size_t, std::vector<std::complex<double>>> cache;
    std::vector<std::complex<double>>& tw = cache[n];
    if (tw.size() != n / 2) {
        tw.resize(n / 2);
        for (size_t t = 0; t < n / 2; ++t) {
            const double ang = -2.0 * M_PI * (double)t / (double)n;
            tw[t] = std::complex<double>(cos(ang), sin(ang));
        }
    }
    return tw;
}

// variant 1: float64 radix-2 DIT, in place
static void cfft_f64(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    const std::vector<std::complex<double>>& tw = twiddle_f64(n);
    unsigned bits = 0;
    while (((size_t)1 << bits) < n) ++bits;
    for (size_t i = 0; i < n; ++i) {
        const size_t j = bitrev(i, bits);
        if (j > i) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len / 2;
        for (size_t k = 0; k < half; ++k) {
            const std::complex<double> w = tw[k * (n / len)];
            for (size_t i = k; i < n; i += len) {
                const std::complex<double> u = a[i];
                const std::complex<double> x = a[i + half];
                const std::complex<double> v(x.real() * w.real() - x.imag() * w.imag(), x.real() * w.imag() + x.imag() * w.real());
                a[i] = u + v;
                a[i + half] = u - v;
            }
        }
    }
}

// variant 2: f32 radix-2 DIT, plain complex multiply, twiddles rounded from float64 per stage
static void cfft_r2_f32(std::vector<cpx>& a) {
    const size_t n = a.size();
    const std::vector<std::complex<double>>& tw = twiddle_f64(n);
    unsigned bits = 0;
    while (((size_t)1 << bits) < n) ++bits;
    for (size_t i = 0; i < n; ++i) {
        const size_t j = bitrev(i, bits);
        if (j > i) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len / 2;
        for (size_t k = 0; k < half; ++k) {
            const float wr = (float)tw[k * (n / len)].real(), wi = (float)tw[k * (n / len)].imag();
            for (size_t i = k; i < n; i += len) {
                const cpx u = a[i], x = a[i + half];
                const float vr = wr * x.re - wi * x.im;
                const float vi = wr * x.im + wi * x.re;
                a[i] = cpx{u.re + vr, u.im + vi};
                a[i + half] = cpx{u.re - vr, u.im - vi};
            }
        }
    }
}

static inline cpx cmul(cpx w, cpx f) {
    cpx r;
    float p = w.im * f.im;
    r.re = fmaf(w.re, f.re, -p);
    float q = w.im * f.re;
    r.im = fmaf(w.re, f.im, q);
    return r;
}

// TW tables are cached per size and per thread (pure function of M).
static const std::vector<cpx>& twiddle_table(size_t M) {
    static thread_local std::map<size_t, std::vector<cpx>> cache;
    std::vector<cpx>& tw = cache[M];
    if (tw.size() != M) {
        tw.resize(M);
        for (size_t t = 0; t < M; ++t) {
            double a = 2.0 * M_PI * (double)t / (double)M;
            tw[t].re = (float)cos(a);
            tw[t].im = (float)(-sin(a));
        }
    }
    return tw;
}

void cfft_forward(std::vector<cpx>& x) {
    const size_t M = x.size();
    if (M <= 1) return;
    const int variant = g_fft_variant.load(std::memory_order_relaxed);
    if (variant == 1) {
        std::vector<std::complex<double>> a(M);
        for (size_t i = 0; i < M; ++i) a[i] = std::complex<double>((double)x[i].re, (double)x[i].im);
        cfft_f64(a);
        for (size_t i = 0; i < M; ++i) x[i] = cpx{(float)a[i].real(), (float)a[i].imag()};
        return;
    }
    if (variant == 2) {
        cfft_r2_f32(x);
        return;
    }
    const std::vector<cpx>& tw = twiddle_table(M);
    static thread_local std::vector<cpx> y;
    y.resize(M);
    cpx* in = x.data();
    cpx* out = y.data();
    size_t m = 0;
    while (((size_t)1 << m) < M) ++m;
    size_t Ns = 1;
    if (m & 1) {  // radix-2 pass, Ns = 1: twiddle index 0 for both inputs
        const size_t half = M / 2;
        for (size_t j = 0; j < half; ++j) {
            cpx a = cmul(tw[0], in[j]);
            cpx b = cmul(tw[0], in[j + half]);
            out[2 * j].re = a.re + b.re;
            out[2 * j].im = a.im + b.im;
            out[2 * j + 1].re = a.re - b.re;
            out[2 * j + 1].im = a.im - b.im;
        }
        std::swap(in, out);
        Ns = 2;
    }
    while (Ns < M) {
        const size_t q = M / 4;
        const size_t tstep = M / (Ns * 4);
        for (size_t j = 0; j < q; ++j) {
            const size_t k = j % Ns;
            cpx A = cmul(tw[0], in[j]);
            cpx B = cmul(tw[(k * 1) * tstep], in[j + q]);
            cpx C = cmul(tw[(k * 2) * tstep], in[j + 2 * q]);
            cpx D = cmul(tw[(k * 3) * tstep], in[j + 3 * q]);
            cpx t0{A.re + C.re, A.im + C.im}, t1{A.re - C.re, A.im - C.im};
            cpx t2{B.re + D.re, B.im + D.im}, t3{B.re - D.re, B.im - D.im};
            const size_t o = (j - k) * 4 + k;
            out[o] = cpx{t0.re + t2.re, t0.im + t2.im};
            out[o + Ns] = cpx{t1.re + t3.im, t1.im - t3.re};
            out[o + 2 * Ns] = cpx{t0.re - t2.re, t0.im - t2.im};
            out[o + 3 * Ns] = cpx{t1.re - t3.im, t1.im + t3.re};
        }
        std::swap(in, out);
        Ns *= 4;
    }
    if (in != x.data()) x.assign(in, in + M);
}

void rfft_forward(const float* x, size_t n, std::vector<cpx>& X) {
    const size_t M = n / 2;
    const int variant = g_fft_variant.load(std::memory_order_relaxed);
    if (variant == 1) {  // full complex frame in float64, rounded once
        std::vector<std::complex<double>> a(n);
        for (size_t i = 0; i < n; ++i) a[i] = std::complex<double>((double)x[i], 0.0);
        cfft_f64(a);
        X.resize(M + 1);
        for (size_t k = 0; k <= M; ++k) X[k] = cpx{(float)a[k].real(), (float)a[k].imag()};
        return;
    }
    if (variant == 2) {  // full complex frame (imaginary part 0), as the reference feeds rustfft
        std::vector<cpx> a(n);
        for (size_t i = 0; i < n; ++i) a[i] = cpx{x[i], 0.0f};
        cfft_r2_f32(a);
        X.assign(a.begin(), a.begin() + (M + 1));
        return;
    }
    static thread_local std::vector<cpx> z;
    z.resize(M);
    for (size_t i = 0; i < M; ++i) z[i] = cpx{x[2 * i], x[2 * i + 1]};
    cfft_forward(z);
    X.resize(M + 1);
    const std::vector<cpx>& rw = twiddle_table(n);  // RW[k] = TW_n[k], k <= n/2
    for (size_t k = 0; k <= M; ++k) {
        cpx a = z[k % M], b = z[(M - k) % M];
        cpx E{a.re + b.re, a.im - b.im}, O{a.re - b.re, a.im + b.im};
        cpx T = cmul(rw[k], O);
        X[k].re = 0.5f * (E.re + T.im);
        X[k].im = 0.5f * (E.im - T.re);
    }
}

}  // namespace so
