// fft_inplace.h -- in-place mixed-radix FFT filter over a batch of rows held in shared memory
// (reference low_pass.py:41-78: rfft_i -> x smmz[j][n] -> irfft_i).
//
// Two real rows (two layers of one latitude) are packed as one complex row z = a + i b; the multiplier is
// real and symmetric in the wavenumber, so one complex transform pair filters both (see fft_rows.h).
//
// No ping-pong buffer and no reordering pass:
//   forward  = decimation in frequency: natural order in, digit-reversed order out.  Stage s (radix r,
//              block length n = N / (r_0 .. r_{s-1})) combines the r elements q + t n/r of each block,
//              multiplies output m by w_n^{q m} and stores it at q + m n/r.  After the last stage position
//              p = m_0 N/r_0 + m_1 N/(r_0 r_1) + ... holds wavenumber k = m_0 + r_0 m_1 + r_0 r_1 m_2 + ...
//   multiply = position p is scaled by table[min(k, N - k)], k = kperm[p] (folded into the last stage);
//   inverse  = the exact mirror (decimation in time, stages in reverse order, conjugate twiddles): digit-
//              reversed in, natural order out, scaled by N.
// All batch rows go through a stage together, so a stage costs one __syncthreads for the whole batch.
#pragma once
#include "fft_rows.h"

// (row, butterfly) of flat work item w without integer division (operands < 2^16, see gcm_fastdiv)
struct GcmFftStage {
  int N, n, stride, nbf, tstep;
  unsigned magic_stride, magic_nbf;
};
__device__ __forceinline__ GcmFftStage gcm_fft_stage(const GcmFftPlan& plan, int s, int n) {
  GcmFftStage st;
  st.N = plan.n;
  st.n = n;
  st.stride = plan.stride[s];
  st.nbf = plan.n / plan.radix[s];
  st.tstep = plan.n / n;
  st.magic_stride = plan.magic_stride[s];
  st.magic_nbf = plan.magic_nbf[s];
  return st;
}

// one forward DIF stage over `nrows` rows of length N (row r starts at z + r * N)
// `table` is the multiplier row of the first latitude; packed row r of the batch uses
// table + ((pr0 + r) / NPJ) * table_stride (NPJ packed rows per latitude, pr0 = packed rows before this batch;
// NPJ = 0: one table for the whole batch)
template <int R, int NPJ>
__device__ __forceinline__ void gcm_dif_stage(double2* z, const GcmFftStage st, int nrows,
                                              const double2* __restrict__ tw, const int* __restrict__ kperm,
                                              const double* __restrict__ table, int table_stride, int pr0, bool last,
                                              int tid, int nthr) {
  const int stride = st.stride, N = st.N;
  const int total = st.nbf * nrows;
  for (int w = tid; w < total; w += nthr) {
    const int row = gcm_fastdiv(w, st.magic_nbf), b = w - row * st.nbf;
    const int blk = gcm_fastdiv(b, st.magic_stride), q = b - blk * stride;
    double2* base = z + row * N + blk * st.n + q;
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = base[t * stride];
    GcmButterfly<R, -1>::run(x);
    if (q > 0) {
      const int qt = q * st.tstep;
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<-1>(x[m], __ldg(&tw[qt * m]));
    }
    if (last) {  // stride == 1: position p = blk * n + m holds wavenumber kperm[p]
#pragma unroll
      for (int m = 0; m < R; ++m) {
        const int k = __ldg(&kperm[blk * st.n + m]);
        const double* trow = NPJ > 0 ? table + ((pr0 + row) / (NPJ > 0 ? NPJ : 1)) * table_stride : table;
        const double s = __ldg(&trow[k <= N - k ? k : N - k]);
        x[m].x *= s;
        x[m].y *= s;
      }
    }
#pragma unroll
    for (int m = 0; m < R; ++m) base[m * stride] = x[m];
  }
}

// one inverse DIT stage (mirror of gcm_dif_stage)
template <int R>
__device__ __forceinline__ void gcm_dit_stage(double2* z, const GcmFftStage st, int nrows,
                                              const double2* __restrict__ tw, int tid, int nthr) {
  const int stride = st.stride, N = st.N;
  const int total = st.nbf * nrows;
  for (int w = tid; w < total; w += nthr) {
    const int row = gcm_fastdiv(w, st.magic_nbf), b = w - row * st.nbf;
    const int blk = gcm_fastdiv(b, st.magic_stride), q = b - blk * stride;
    double2* base = z + row * N + blk * st.n + q;
    double2 x[R];
#pragma unroll
    for (int m = 0; m < R; ++m) x[m] = base[m * stride];
    if (q > 0) {
      const int qt = q * st.tstep;
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<+1>(x[m], __ldg(&tw[qt * m]));
    }
    GcmButterfly<R, +1>::run(x);
#pragma unroll
    for (int t = 0; t < R; ++t) base[t * stride] = x[t];
  }
}

// true when every radix of the plan has an unrolled in-place butterfly
__host__ __device__ inline bool gcm_plan_inplace_ok(const GcmFftPlan& plan) {
  for (int p = 0; p < plan.npass; ++p)
    if (!gcm_radix_unrolled(plan.radix[p])) return false;
  return true;
}

// Filter `nrows` packed rows in place.  On entry the rows are visible to the whole block; on exit they hold
// N x (filtered rows) and are synchronised.
template <int NPJ>
__device__ __forceinline__ void gcm_filter_rows_inplace(double2* z, int nrows, const GcmFftPlan& plan,
                                                        const double2* __restrict__ tw, const int* __restrict__ kperm,
                                                        const double* __restrict__ table_row, int table_stride, int pr0,
                                                        int tid, int nthr) {
  const int N = plan.n;
  if (N == 1) return;  // low_pass.py:58-59
  int n = N;
  for (int p = 0; p < plan.npass; ++p) {
    const int r = plan.radix[p];
    const bool last = p == plan.npass - 1;
    const GcmFftStage st = gcm_fft_stage(plan, p, n);
    switch (r) {
      case 2: gcm_dif_stage<2, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 3: gcm_dif_stage<3, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 4: gcm_dif_stage<4, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 5: gcm_dif_stage<5, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 6: gcm_dif_stage<6, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 8: gcm_dif_stage<8, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 9: gcm_dif_stage<9, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 10: gcm_dif_stage<10, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 12: gcm_dif_stage<12, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      case 15: gcm_dif_stage<15, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
      default: gcm_dif_stage<16, NPJ>(z, st, nrows, tw, kperm, table_row, table_stride, pr0, last, tid, nthr); break;
    }
    n /= r;
    __syncthreads();
  }
  // n == 1 here; walk the stages back up
  for (int p = plan.npass - 1; p >= 0; --p) {
    const int r = plan.radix[p];
    n *= r;
    const GcmFftStage st = gcm_fft_stage(plan, p, n);
    switch (r) {
      case 2: gcm_dit_stage<2>(z, st, nrows, tw, tid, nthr); break;
      case 3: gcm_dit_stage<3>(z, st, nrows, tw, tid, nthr); break;
      case 4: gcm_dit_stage<4>(z, st, nrows, tw, tid, nthr); break;
      case 5: gcm_dit_stage<5>(z, st, nrows, tw, tid, nthr); break;
      case 6: gcm_dit_stage<6>(z, st, nrows, tw, tid, nthr); break;
      case 8: gcm_dit_stage<8>(z, st, nrows, tw, tid, nthr); break;
      case 9: gcm_dit_stage<9>(z, st, nrows, tw, tid, nthr); break;
      case 10: gcm_dit_stage<10>(z, st, nrows, tw, tid, nthr); break;
      case 12: gcm_dit_stage<12>(z, st, nrows, tw, tid, nthr); break;
      case 15: gcm_dit_stage<15>(z, st, nrows, tw, tid, nthr); break;
      default: gcm_dit_stage<16>(z, st, nrows, tw, tid, nthr); break;
    }
    __syncthreads();
  }
}
