// fft_rows.h -- shared-memory mixed-radix Stockham FFT over one latitude row, used by the polar filter
// (reference low_pass.py:41-78: rfft_i -> x smmz[j][n] -> irfft_i).
//
// Two REAL rows of the same latitude j (two layers k, k+1) are packed as one complex row z = a + i b.
// The filter multiplier is real and symmetric in the wavenumber (s[n] = s[W-n]), so filtering z with one
// complex FFT pair filters both rows at once: Re -> row a, Im -> row b.  No Hermitian split is needed.
//
// Stockham autosort: pass p has radix R, ns = product of earlier radices; butterfly b (0 <= b < n/R) reads
// in[b + t n/R], twiddles by exp(-/+ 2 pi i t k / (ns R)) with k = b mod ns, and writes
// out[(b - k) R + k + t ns].  The data ping-pongs between two shared-memory rows; no bit reversal.
#pragma once
#include "gcm_common.h"

#include "fft_consts.h"

__device__ __forceinline__ double2 gcm_cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 gcm_csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// a * w for DIR = -1 (forward), a * conj(w) for DIR = +1; the table holds w = exp(-i theta)
template <int DIR>
__device__ __forceinline__ double2 gcm_cmul_tw(double2 a, double2 w) {
  const double wy = DIR < 0 ? w.y : -w.y;
  return make_double2(a.x * w.x - a.y * wy, a.x * wy + a.y * w.x);
}
// a * (i * DIR)
template <int DIR>
__device__ __forceinline__ double2 gcm_mul_i(double2 a) {
  return DIR > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <int R, int DIR>
struct GcmButterfly;

template <int DIR>
struct GcmButterfly<2, DIR> {
  __device__ static __forceinline__ void run(double2* x) {
    const double2 a = x[0], b = x[1];
    x[0] = gcm_cadd(a, b);
    x[1] = gcm_csub(a, b);
  }
};

template <int DIR>
struct GcmButterfly<3, DIR> {
  __device__ static __forceinline__ void run(double2* x) {
    const double s = 0.86602540378443864676;  // sin(2 pi / 3)
    const double2 t1 = gcm_cadd(x[1], x[2]);
    const double2 d = gcm_csub(x[1], x[2]);
    const double2 m = make_double2(x[0].x - 0.5 * t1.x, x[0].y - 0.5 * t1.y);
    const double2 id = gcm_mul_i<DIR>(make_double2(s * d.x, s * d.y));
    x[0] = gcm_cadd(x[0], t1);
    x[1] = gcm_cadd(m, id);
    x[2] = gcm_csub(m, id);
  }
};

template <int DIR>
struct GcmButterfly<4, DIR> {
  __device__ static __forceinline__ void run(double2* x) {
    const double2 a = gcm_cadd(x[0], x[2]), b = gcm_csub(x[0], x[2]);
    const double2 c = gcm_cadd(x[1], x[3]), d = gcm_mul_i<DIR>(gcm_csub(x[1], x[3]));
    x[0] = gcm_cadd(a, c);
    x[1] = gcm_cadd(b, d);
    x[2] = gcm_csub(a, c);
    x[3] = gcm_csub(b, d);
  }
};

template <int DIR>
struct GcmButterfly<5, DIR> {
  __device__ static __forceinline__ void run(double2* x) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;  // cos(2pi/5), cos(4pi/5)
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;   // sin(2pi/5), sin(4pi/5)
    const double2 a1 = gcm_cadd(x[1], x[4]), a2 = gcm_cadd(x[2], x[3]);
    const double2 b1 = gcm_csub(x[1], x[4]), b2 = gcm_csub(x[2], x[3]);
    const double2 r1 = make_double2(x[0].x + c1 * a1.x + c2 * a2.x, x[0].y + c1 * a1.y + c2 * a2.y);
    const double2 r2 = make_double2(x[0].x + c2 * a1.x + c1 * a2.x, x[0].y + c2 * a1.y + c1 * a2.y);
    const double2 i1 = gcm_mul_i<DIR>(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const double2 i2 = gcm_mul_i<DIR>(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    x[0] = make_double2(x[0].x + a1.x + a2.x, x[0].y + a1.y + a2.y);
    x[1] = gcm_cadd(r1, i1);
    x[4] = gcm_csub(r1, i1);
    x[2] = gcm_cadd(r2, i2);
    x[3] = gcm_csub(r2, i2);
  }
};

// v * w_R^a for DIR = -1 (w_R = exp(-2 pi i / R)), v * conj(w_R^a) for DIR = +1.  R and a are compile-time
// constants once the butterflies are unrolled, so the branches fold and the table values become immediates.
template <int DIR>
__device__ __forceinline__ double2 gcm_mul_const_tw(double2 v, int R, int a) {
  a %= R;
  if (a == 0) return v;
  if (2 * a == R) return make_double2(-v.x, -v.y);
  if (4 * a == R) return gcm_mul_i<DIR>(v);
  if (4 * a == 3 * R) return gcm_mul_i<-DIR>(v);
  const double c = gcm_tw_cos(R, a);
  const double wy = DIR < 0 ? -gcm_tw_sin(R, a) : gcm_tw_sin(R, a);  // imaginary part of the factor
  if ((8 * a) % R == 0) {  // odd eighth turn: |c| == |wy|
    if ((c > 0) == (wy > 0)) return make_double2(c * (v.x - v.y), c * (v.x + v.y));
    return make_double2(c * (v.x + v.y), c * (v.y - v.x));
  }
  return make_double2(v.x * c - v.y * wy, v.x * wy + v.y * c);
}

// Composite radix R = R1 R2 inside one thread (Cooley-Tukey): with t = R2 t1 + t2 and m = m1 + R1 m2,
//   X[m1 + R1 m2] = sum_t2 w_R2^(t2 m2) * w_R^(t2 m1) * sum_t1 x[R2 t1 + t2] w_R1^(t1 m1)
template <int R1, int R2, int DIR>
struct GcmButterflyCT {
  __device__ static __forceinline__ void run(double2* x) {
    constexpr int R = R1 * R2;
    double2 a[R];
#pragma unroll
    for (int t2 = 0; t2 < R2; ++t2) {
      double2 y[R1];
#pragma unroll
      for (int t1 = 0; t1 < R1; ++t1) y[t1] = x[R2 * t1 + t2];
      GcmButterfly<R1, DIR>::run(y);
#pragma unroll
      for (int m1 = 0; m1 < R1; ++m1) a[t2 * R1 + m1] = gcm_mul_const_tw<DIR>(y[m1], R, t2 * m1);
    }
#pragma unroll
    for (int m1 = 0; m1 < R1; ++m1) {
      double2 y[R2];
#pragma unroll
      for (int t2 = 0; t2 < R2; ++t2) y[t2] = a[t2 * R1 + m1];
      GcmButterfly<R2, DIR>::run(y);
#pragma unroll
      for (int m2 = 0; m2 < R2; ++m2) x[m1 + R1 * m2] = y[m2];
    }
  }
};
template <int DIR> struct GcmButterfly<6, DIR> : GcmButterflyCT<3, 2, DIR> {};
template <int DIR> struct GcmButterfly<8, DIR> : GcmButterflyCT<4, 2, DIR> {};
template <int DIR> struct GcmButterfly<9, DIR> : GcmButterflyCT<3, 3, DIR> {};
template <int DIR> struct GcmButterfly<10, DIR> : GcmButterflyCT<5, 2, DIR> {};
template <int DIR> struct GcmButterfly<12, DIR> : GcmButterflyCT<4, 3, DIR> {};
template <int DIR> struct GcmButterfly<15, DIR> : GcmButterflyCT<5, 3, DIR> {};
template <int DIR> struct GcmButterfly<16, DIR> : GcmButterflyCT<4, 4, DIR> {};

// radices with an unrolled butterfly
__host__ __device__ inline bool gcm_radix_unrolled(int r) {
  return r == 2 || r == 3 || r == 4 || r == 5 || r == 6 || r == 8 || r == 9 || r == 10 || r == 12 || r == 15 || r == 16;
}

template <int R, int DIR>
__device__ __forceinline__ void gcm_fft_pass(const double2* in, double2* out, int n, int ns, const double2* tw, int tid,
                                             int nthr) {
  const int nb = n / R;
  const int tstep = n / (ns * R);
  for (int b = tid; b < nb; b += nthr) {
    const int k = b % ns;
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      double2 v = in[b + t * nb];
      if (t > 0 && k > 0) v = gcm_cmul_tw<DIR>(v, __ldg(&tw[t * k * tstep]));  // t k tstep < n
      x[t] = v;
    }
    GcmButterfly<R, DIR>::run(x);
    const int o = (b - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) out[o + t * ns] = x[t];
  }
}

// any other (prime) radix: O(R^2) DFT reading the inputs straight from shared memory
template <int DIR>
__device__ __forceinline__ void gcm_fft_pass_generic(const double2* in, double2* out, int n, int R, int ns,
                                                     const double2* tw, int tid, int nthr) {
  const int nb = n / R;
  const int tstep = n / (ns * R);
  const int rstep = n / R;
  for (int w = tid; w < n; w += nthr) {  // one output per work item
    const int b = w % nb, m = w / nb;
    const int k = b % ns;
    double2 acc = make_double2(0.0, 0.0);
    for (int t = 0; t < R; ++t) {
      double2 v = in[b + t * nb];
      if (t > 0 && k > 0) v = gcm_cmul_tw<DIR>(v, __ldg(&tw[t * k * tstep]));
      v = gcm_cmul_tw<DIR>(v, __ldg(&tw[((t * m) % R) * rstep]));
      acc = gcm_cadd(acc, v);
    }
    out[(b - k) * R + k + m * ns] = acc;
  }
}

// Transform the row held in `a` (visible to all threads of the block on entry); returns the buffer
// (a or b) that holds the result, already synchronised.
template <int DIR>
__device__ __forceinline__ double2* gcm_fft_run(double2* a, double2* b, const GcmFftPlan& plan, const double2* tw,
                                                int tid, int nthr) {
  const int n = plan.n;
  int ns = 1;
  for (int p = 0; p < plan.npass; ++p) {
    const int r = plan.radix[p];
    switch (r) {
      case 2: gcm_fft_pass<2, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 3: gcm_fft_pass<3, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 4: gcm_fft_pass<4, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 5: gcm_fft_pass<5, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 6: gcm_fft_pass<6, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 8: gcm_fft_pass<8, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 9: gcm_fft_pass<9, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 10: gcm_fft_pass<10, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 12: gcm_fft_pass<12, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 15: gcm_fft_pass<15, DIR>(a, b, n, ns, tw, tid, nthr); break;
      case 16: gcm_fft_pass<16, DIR>(a, b, n, ns, tw, tid, nthr); break;
      default: gcm_fft_pass_generic<DIR>(a, b, n, r, ns, tw, tid, nthr); break;
    }
    __syncthreads();
    double2* t = a;
    a = b;
    b = t;
    ns *= r;
  }
  return a;
}

// Polar filter of a packed row pair: forward FFT, multiply wavenumber n (and W - n) by table[n], inverse
// FFT.  Returns the buffer holding W * (filtered rows); the caller scales by 1/W on the way out, like
// the 1/n normalisation numpy's irfft applies to its output.
__device__ __forceinline__ double2* gcm_filter_pair(double2* a, double2* b, const GcmFftPlan& plan, const double2* tw,
                                                    const double* table_row, int tid, int nthr) {
  const int W = plan.n;
  if (W == 1) return a;  // low_pass.py:58-59: a single column is returned unfiltered
  double2* f = gcm_fft_run<-1>(a, b, plan, tw, tid, nthr);
  double2* o = (f == a) ? b : a;
  for (int i = tid; i < W; i += nthr) {
    const int m = (i <= W - i) ? i : W - i;
    const double s = __ldg(&table_row[m]);
    double2 v = f[i];
    v.x *= s;
    v.y *= s;
    f[i] = v;
  }
  __syncthreads();
  return gcm_fft_run<+1>(f, o, plan, tw, tid, nthr);
}
