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
//   multiply = position p is scaled by the table entry of wavenumber k = kperm[p]; the table is stored in
//              transform order (GcmGeomDev::smmzp) and the multiply is folded between the last forward and the
//              first inverse stage, which are one visit (gcm_mid_stage);
//   inverse  = the exact mirror (decimation in time, stages in reverse order, conjugate twiddles): digit-
//              reversed in, natural order out; numpy's 1/N is folded into the table.
// The first forward stage is fed from global memory and the last inverse stage drains to it (gcm_filter_rows_io).
// All batch rows go through a stage together, so a stage costs one __syncthreads for the whole batch.
#pragma once
#include "fft_rows.h"

// (row, butterfly) of flat work item w without integer division (operands < 2^16, see gcm_fastdiv)
struct GcmFftStage {
  int N, n, stride, nbf, twoff;
  unsigned magic_stride, magic_nbf;
};
__device__ __forceinline__ GcmFftStage gcm_fft_stage(const GcmFftPlan& plan, int s) {
  GcmFftStage st;
  st.N = plan.n;
  st.n = plan.blk[s];
  st.stride = plan.stride[s];
  st.nbf = plan.nbf[s];
  st.twoff = plan.twoff[s];
  st.magic_stride = plan.magic_stride[s];
  st.magic_nbf = plan.magic_nbf[s];
  return st;
}

// one forward DIF stage over `nrows` rows of length N (row r starts at z + r * N)
template <int R>
__device__ __forceinline__ void gcm_dif_stage(double2* z, const GcmFftStage st, int nrows,
                                              const double2* __restrict__ tw, int tid, int nthr) {
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
      const double2* t = tw + st.twoff + q;  // lanes hold consecutive q: coalesced
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<-1>(x[m], __ldg(&t[(m - 1) * stride]));
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
      const double2* t = tw + st.twoff + q;
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<+1>(x[m], __ldg(&t[(m - 1) * stride]));
    }
    GcmButterfly<R, +1>::run(x);
#pragma unroll
    for (int t = 0; t < R; ++t) base[t * stride] = x[t];
  }
}

// last forward stage + multiply + first inverse stage in one visit (stride 1: the R elements of a butterfly are
// contiguous and q = 0, so no twiddles).  `table` is the multiplier in transform order (GcmGeomDev::smmzp, or one row
// of it when NPJ = 0); packed row r of the batch belongs to latitude gcm_seg_row(seg, (pr0 + r) / NPJ) (NPJ packed
// rows per latitude, pr0 = packed rows of the launch before this batch).
template <int R, int NPJ>
__device__ __forceinline__ void gcm_mid_stage(double2* z, const GcmFftStage st, int nrows,
                                              const double* __restrict__ table, const GcmRowSeg seg, int pr0, int tid,
                                              int nthr) {
  const int N = st.N;
  const int total = st.nbf * nrows;
  for (int w = tid; w < total; w += nthr) {
    const int row = gcm_fastdiv(w, st.magic_nbf), blk = w - row * st.nbf;
    double2* base = z + row * N + blk * R;
    const double* trow =
        (NPJ > 0 ? table + (size_t)gcm_seg_row(seg, (pr0 + row) / (NPJ > 0 ? NPJ : 1)) * N : table) + blk * R;
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = base[t];
    GcmButterfly<R, -1>::run(x);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const double s = __ldg(&trow[m]);
      x[m].x *= s;
      x[m].y *= s;
    }
    GcmButterfly<R, +1>::run(x);
#pragma unroll
    for (int t = 0; t < R; ++t) base[t] = x[t];
  }
}

// ---- transform fed from / drained to global memory ------------------------------------------------------------
// IO supplies the packed rows: ctx = io.begin(row) once per butterfly (row-invariant pointers), io.load(ctx, pos) the
// complex element at position pos of that row, io.store(ctx, pos, v) the filtered element.  The first forward stage
// reads its R inputs straight from global memory (lanes hold consecutive positions: coalesced) and the last inverse
// stage writes straight back, so the data crosses shared memory 2 npass - 2 times instead of 2 npass, and the two
// copy loops with their barriers are gone.
template <int R, class IO>
__device__ __forceinline__ void gcm_dif_stage_first(double2* z, const GcmFftStage st, int nrows,
                                                    const double2* __restrict__ tw, IO& io, int tid, int nthr) {
  const int stride = st.stride, N = st.N;  // stage 0: one block per row, stride = N / R = butterflies per row
  const int total = stride * nrows;
  for (int w = tid; w < total; w += nthr) {
    const int row = gcm_fastdiv(w, st.magic_nbf), q = w - row * stride;
    const auto ctx = io.begin(row);
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = io.load(ctx, q + t * stride);
    GcmButterfly<R, -1>::run(x);
    if (q > 0) {
      const double2* t = tw + st.twoff + q;
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<-1>(x[m], __ldg(&t[(m - 1) * stride]));
    }
    double2* base = z + row * N + q;
#pragma unroll
    for (int m = 0; m < R; ++m) base[m * stride] = x[m];
  }
}

template <int R, class IO>
__device__ __forceinline__ void gcm_dit_stage_last(const double2* z, const GcmFftStage st, int nrows,
                                                   const double2* __restrict__ tw, IO& io, int tid, int nthr) {
  const int stride = st.stride, N = st.N;
  const int total = stride * nrows;
  for (int w = tid; w < total; w += nthr) {
    const int row = gcm_fastdiv(w, st.magic_nbf), q = w - row * stride;
    const double2* base = z + row * N + q;
    double2 x[R];
#pragma unroll
    for (int m = 0; m < R; ++m) x[m] = base[m * stride];
    if (q > 0) {
      const double2* t = tw + st.twoff + q;
#pragma unroll
      for (int m = 1; m < R; ++m) x[m] = gcm_cmul_tw<+1>(x[m], __ldg(&t[(m - 1) * stride]));
    }
    GcmButterfly<R, +1>::run(x);
    const auto ctx = io.begin(row);
#pragma unroll
    for (int t = 0; t < R; ++t) io.store(ctx, q + t * stride, x[t]);
  }
}

// single-stage plans (N = R): global -> butterfly -> multiply -> inverse butterfly -> global, no shared memory
template <int R, int NPJ, class IO>
__device__ __forceinline__ void gcm_mid_stage_io(const GcmFftStage st, int nrows, const double* __restrict__ table,
                                                 const GcmRowSeg seg, int pr0, IO& io, int tid, int nthr) {
  for (int row = tid; row < nrows; row += nthr) {
    const double* trow = NPJ > 0 ? table + (size_t)gcm_seg_row(seg, (pr0 + row) / (NPJ > 0 ? NPJ : 1)) * R : table;
    const auto ctx = io.begin(row);
    double2 x[R];
#pragma unroll
    for (int t = 0; t < R; ++t) x[t] = io.load(ctx, t);
    GcmButterfly<R, -1>::run(x);
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const double s = __ldg(&trow[m]);
      x[m].x *= s;
      x[m].y *= s;
    }
    GcmButterfly<R, +1>::run(x);
#pragma unroll
    for (int t = 0; t < R; ++t) io.store(ctx, t, x[t]);
  }
  (void)st;
}

// true when every radix of the plan has an unrolled in-place butterfly
__host__ __device__ inline bool gcm_plan_inplace_ok(const GcmFftPlan& plan) {
  for (int p = 0; p < plan.npass; ++p)
    if (!gcm_radix_unrolled(plan.radix[p])) return false;
  return true;
}

// Filter `nrows` packed rows in place.  On entry the rows are visible to the whole block; on exit they hold
// N x (filtered rows) and are synchronised.
#define GCM_RADIX_SWITCH(r, CALL)   \
  switch (r) {                     \
    case 2: CALL(2); break;        \
    case 3: CALL(3); break;        \
    case 4: CALL(4); break;        \
    case 5: CALL(5); break;        \
    case 6: CALL(6); break;        \
    case 8: CALL(8); break;        \
    case 9: CALL(9); break;        \
    case 10: CALL(10); break;      \
    case 12: CALL(12); break;      \
    case 15: CALL(15); break;      \
    default: CALL(16); break;      \
  }

// Filter `nrows` packed rows read from and written to global memory through `io` (see gcm_dif_stage_first).
// z is only the inter-stage buffer (nrows * N complex); on exit it may be reused at once.
template <int NPJ, class IO>
__device__ __forceinline__ void gcm_filter_rows_io(double2* z, int nrows, const GcmFftPlan& plan,
                                                   const double2* __restrict__ tw, const double* __restrict__ table,
                                                   const GcmRowSeg seg, int pr0, IO& io, int tid, int nthr) {
  const int last = plan.npass - 1;
  if (last == 0) {
    const GcmFftStage st = gcm_fft_stage(plan, 0);
#define GCM_CALL(R) gcm_mid_stage_io<R, NPJ>(st, nrows, table, seg, pr0, io, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[0], GCM_CALL)
#undef GCM_CALL
    return;
  }
  {
    const GcmFftStage st = gcm_fft_stage(plan, 0);
#define GCM_CALL(R) gcm_dif_stage_first<R>(z, st, nrows, tw, io, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[0], GCM_CALL)
#undef GCM_CALL
    __syncthreads();
  }
  for (int p = 1; p < last; ++p) {
    const GcmFftStage st = gcm_fft_stage(plan, p);
#define GCM_CALL(R) gcm_dif_stage<R>(z, st, nrows, tw, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[p], GCM_CALL)
#undef GCM_CALL
    __syncthreads();
  }
  {
    const GcmFftStage st = gcm_fft_stage(plan, last);
#define GCM_CALL(R) gcm_mid_stage<R, NPJ>(z, st, nrows, table, seg, pr0, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[last], GCM_CALL)
#undef GCM_CALL
    __syncthreads();
  }
  for (int p = last - 1; p >= 1; --p) {
    const GcmFftStage st = gcm_fft_stage(plan, p);
#define GCM_CALL(R) gcm_dit_stage<R>(z, st, nrows, tw, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[p], GCM_CALL)
#undef GCM_CALL
    __syncthreads();
  }
  {
    const GcmFftStage st = gcm_fft_stage(plan, 0);
#define GCM_CALL(R) gcm_dit_stage_last<R>(z, st, nrows, tw, io, tid, nthr)
    GCM_RADIX_SWITCH(plan.radix[0], GCM_CALL)
#undef GCM_CALL
  }
  __syncthreads();
}

// ---- plans known at compile time ---------------------------------------------------------------------------------
// gcm_filter_rows_io picks every stage's butterfly with a runtime switch over eleven radices: the kernel image holds
// 5 x 11 unrolled stage bodies (~20 000 SASS instructions, 330 KB) of which one plan executes ~2 200, scattered over
// the image -- ncu shows `no_instruction` (instruction fetch) as the first stall reason of the filter kernels (r02m).
// For the plans of the grids BASELINE.json names the radices are template constants, so the image of a filter kernel
// is the five stage bodies it runs and nothing else (fits the 32 KB L1.5 instruction cache).
template <int PLAN>
struct GcmFixedPlan;
template <> struct GcmFixedPlan<1> { static constexpr int n = 3, r0 = 12, r1 = 8, rl = 15; };  // W = 1440
template <> struct GcmFixedPlan<2> { static constexpr int n = 3, r0 = 12, r1 = 8, rl = 3; };   // W = 288
template <> struct GcmFixedPlan<3> { static constexpr int n = 2, r0 = 8, r1 = 0, rl = 9; };    // W = 72
template <> struct GcmFixedPlan<4> { static constexpr int n = 2, r0 = 12, r1 = 0, rl = 3; };   // W = 36
#define GCM_FIXED_PLANS 4

// 0 = no compile-time twin of this plan (take the runtime switch)
__host__ inline int gcm_fixed_plan_id(const GcmFftPlan& p) {
  auto is = [&](int n, int a, int b, int c) {
    return p.npass == n && p.radix[0] == a && p.radix[1] == b && (n == 2 || p.radix[2] == c);
  };
  if (is(3, 12, 8, 15)) return 1;
  if (is(3, 12, 8, 3)) return 2;
  if (is(2, 8, 9, 0)) return 3;
  if (is(2, 12, 3, 0)) return 4;
  return 0;
}

struct GcmNoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `after_first()` runs once the first stage has consumed its inputs and the block has synchronised (the pipelined
// filter kernel starts the bulk copy of its NEXT rows there: they land under the remaining stages)
template <int NPJ, int PLAN, class IO, class Hook = GcmNoHook>
__device__ __forceinline__ void gcm_filter_rows_io_fixed(double2* z, int nrows, const GcmFftPlan& plan,
                                                         const double2* __restrict__ tw,
                                                         const double* __restrict__ table, const GcmRowSeg seg, int pr0,
                                                         IO& io, int tid, int nthr, Hook after_first = Hook()) {
  using P = GcmFixedPlan<PLAN>;
  constexpr int last = P::n - 1;
  gcm_dif_stage_first<P::r0>(z, gcm_fft_stage(plan, 0), nrows, tw, io, tid, nthr);
  __syncthreads();
  after_first();
  if constexpr (P::n == 3) {
    gcm_dif_stage<(P::r1 > 0 ? P::r1 : 2)>(z, gcm_fft_stage(plan, 1), nrows, tw, tid, nthr);
    __syncthreads();
  }
  gcm_mid_stage<P::rl, NPJ>(z, gcm_fft_stage(plan, last), nrows, table, seg, pr0, tid, nthr);
  __syncthreads();
  if constexpr (P::n == 3) {
    gcm_dit_stage<(P::r1 > 0 ? P::r1 : 2)>(z, gcm_fft_stage(plan, 1), nrows, tw, tid, nthr);
    __syncthreads();
  }
  gcm_dit_stage_last<P::r0>(z, gcm_fft_stage(plan, 0), nrows, tw, io, tid, nthr);
  __syncthreads();
}
