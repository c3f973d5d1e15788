// pe25_fast_impl.h -- fused kernels of the 2.5-D half step (reference dynamics.py:183-227).
//
// Five launches per half step, split at the only data dependences that span a whole latitude row (the zonal FFT
// filter, low_pass.py:41-78) or a whole column (sums over k) -- everything else is recomputed where it is needed
// instead of stored.  Two independent chains run side by side (caller's stream + the geometry's side stream):
//
//   F1  pe25f_filter_kernel<1>   spu = arakawa_1977(su * iph(sp))                          (dynamics.py:187-189)
//   A   pe25f_aflux_kernel       pit = sum_k conv, p_n = p - pit dt  (+ sd on the direct-load update path)  (:35-46, :194)
//   H   pe25f_hydro_tile_kernel (W >= 256) / pe25f_hydro_kernel / pe25f_hydro_narrow_kernel (W < 62)
//                                hydrostatic phi, rho by column in registers -> pgfu + phiu, fv = phiv + pgv (:111-171)
//   F0  pe25f_filter_kernel<0>   pgf = arakawa_1977(pgfu + phiu), in place                 (:202)
//   U   pe25f_update_tiled_kernel (32- or 36-wide tiles, LDGSTS) / pe25f_update_kernel / pe25f_update_cell_kernel
//                                momentum and tracer update                                (:197-222)
// Measured alternatives kept behind tuning knobs: pe25f_update_tma_kernel (TMA box loads on
// mbarriers, warp-specialised), pe25f_filter_pipe_kernel (persistent, bulk-copy prefetch), filter MODE 2 (aflux fused).
// Instantiated per layer count in pe25_fast_l{3,9,17,18}.cu.
//
// Work fields in HBM between the launches: spu, pgf, fv (3-D), pit, p_n (2-D).  phi and rho never leave the SM; sd is
// rebuilt inside U.
//   * divides by metric terms are multiplications by resident reciprocals (1/dx_j, 1/dx_h, 1/dy, 1/dsig), the
//     divides by p_n averages are done once per column, the 1/W of the inverse transform is folded into the
//     filter table, and when ptop = 0 (the reference's setting, geometry.py:147) the Exner factor
//     ((sig p + ptop)/P0)^kappa factorises into sig^kappa (resident) x (p/P0)^kappa: one exp/log per column;
//   * per-layer tables ride in the kernel parameters (constant bank), indices are 32-bit;
//   * every kernel takes row segments (GcmRowSeg), so a latitude band can compute its interior rows while the halo
//     exchange is in flight (comm.cu).
// FMA contraction is on for this file.  Results agree with the reference within the stated fp64 tolerance
// (tests/test_parity.py); the bit-exact operator kernels stay in pe25.cu.
#pragma once
#include "fft_inplace.h"
#include "gcm_common.h"
#include "gcm_tma.h"
#include "prof.h"

struct PfConst {
  const double *p, *u, *v, *t, *q;
};
struct PfMut {
  double *p, *u, *v, *t, *q;
};
struct PfWork {
  double *spu, *sd, *pgf, *fv, *pn, *pit;
};

// hydrostatic geopotential at layer centres and density of one column (dynamics.py:111-142, :150-152);
// t points at layer 0 of the column, ks = layer stride
template <int L, bool PTOP0>
__device__ __forceinline__ void pf_column(const GcmGeomDev& g, double sp_c, double hm, const double* __restrict__ t,
                                          int ks, double* phi, double* rho) {
  const double ptop = g.ptop;
  double pk_col = 0.0;
  if (PTOP0) pk_col = exp(GCM_KAPPA * log(sp_c * (1.0 / GCM_P0)));  // (p / P0)^kappa
  double tk = t[0];
  double pk = PTOP0 ? g.c_sigkap[0] * pk_col : pow((g.c_sig[0] * sp_c + ptop) / GCM_P0, GCM_KAPPA);
  const double t0 = tk, pk0 = pk;
  double sum = 0.0;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    double t_n = t0, pk_n = pk0;  // k + 1 wraps to layer 0 (coordinates_3d.py:55); sigt[L-1] = 0 kills it
    if (k + 1 < L) {
      t_n = t[(k + 1) * ks];
      pk_n = PTOP0 ? g.c_sigkap[k + 1] * pk_col : pow((g.c_sig[k + 1] * sp_c + ptop) / GCM_P0, GCM_KAPPA);
    }
    const double tp = sp_c * g.c_sig[k] + ptop;
    const double rtt = GCM_RD * (tk * pk);                      // Rd * T
    const double r = tp * gcm_rcp(rtt);                         // rho (dynamics.py:152)
    const double spa = PTOP0 ? rtt : (g.c_sig[k] * sp_c) / r;   // sig p / rho
    const double stp = GCM_CP * ((tk + t_n) * 0.5) * (pk - pk_n);
    sum += spa * g.c_dsig[k] - g.c_sigt[k] * stp;
    rho[k] = r;
    if (k + 1 < L) phi[k + 1] = stp;
    tk = t_n;
    pk = pk_n;
  }
  double run = sum + hm * GCM_G;
  phi[0] = run;
#pragma unroll
  for (int k = 1; k < L; ++k) {
    run += phi[k];
    phi[k] = run;
  }
}

// ---------------------------------------------------------------------------------------------------
// hydrostatic columns
// ---------------------------------------------------------------------------------------------------
// Column phase of one row for one lane: phi and rho of column (j, i) into (phi, rho); if emit_pre, pgfu + phiu of
// row j (dynamics.py:159, :162-165; east neighbour from lane + 1) -> pgf, unfiltered; if emit_fv, fv = phiv + pgv of
// the row to the north (:160, :167-169) from (phi_n, rho_n, sp_n) -> fv at cn.  All lanes of the warp must call.
template <int L, bool PTOP0>
__device__ __forceinline__ double pf_row_step(const GcmGeomDev& g, const double* __restrict__ sp,
                                              const double* __restrict__ st, double* pgf, double* __restrict__ fv,
                                              int plane, int c2, int cn, int j, bool own, bool emit_pre, bool emit_fv,
                                              double* phi, double* rho, const double* phi_n, const double* rho_n,
                                              double sp_n, int c2_next = -1) {
  if (c2_next >= 0) {  // the column of the next row of the march: ask L1 for it now
    gcm_prefetch_l1(sp + c2_next);
#pragma unroll
    for (int k = 0; k < L; ++k) gcm_prefetch_l1(st + k * plane + c2_next);
  }
  const double sp_c = sp[c2];
  pf_column<L, PTOP0>(g, sp_c, g.hmap[c2], st + c2, plane, phi, rho);
  if (emit_pre) {
    const double sp_e = __shfl_down_sync(0xffffffffu, sp_c, 1);
    const double rdxj = g.rdx_j[j];
    const double psum = sp_c + sp_e, gradp = (sp_e - sp_c) * rdxj;
    const double a_u = psum * gradp;       // (p_c + p_e) dp/dx
    const double b_u = psum * 0.5 * rdxj;  // iph(p) / dx
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const double phi_e = __shfl_down_sync(0xffffffffu, phi[k], 1);
      const double rho_e = __shfl_down_sync(0xffffffffu, rho[k], 1);
      const double x = g.c_sig[k] * a_u * gcm_rcp(rho[k] + rho_e) + b_u * (phi_e - phi[k]);
      if (own) pgf[k * plane + c2] = x;
    }
  }
  if (emit_fv && own) {
    const double rdy = g.rdy;
    const double psum = sp_n + sp_c;
    const double a_v = psum * ((sp_c - sp_n) * rdy);  // (p_c + p_jp) dp/dy
    const double b_v = psum * 0.5 * rdy;              // jph(p) / dy
#pragma unroll
    for (int k = 0; k < L; ++k)
      fv[k * plane + cn] = g.c_sig[k] * a_v * gcm_rcp(rho_n[k] + rho[k]) + b_v * (phi[k] - phi_n[k]);
  }
  return sp_c;
}

// ---------------------------------------------------------------------------------------------------
// The row phase: every (row, layer pair) and every (row group, column chunk) is its own unit of parallelism, so a
// latitude band of a few dozen rows (strong scaling over GPUs) still fills the chip, and the column march can span
// RG rows (RG + 1 column evaluations per RG rows).
// ---------------------------------------------------------------------------------------------------
// MODE 1: spu = arakawa_1977(su * iph(sp)) (dynamics.py:187-189);  MODE 0: x = arakawa_1977(x) in place (:202).
// One CTA per NBAT packed rows of the flattened (row, layer pair) list of the rows of `seg`.
// rows of the filter kernel as seen by the transform (fft_inplace.h: gcm_filter_rows_io)
template <int L, int MODE>
struct PfFilterIO {
  const double* __restrict__ sp;
  const double* in;
  double* out;
  GcmRowSeg seg;
  int pr0, W, plane;
  double2* zkeep;  // MODE 2: the filtered packed rows stay in the transform's work rows (read back by the aflux pass)
  struct Ctx {
    const double* s0;
    const double* spr;
    double* o;
    double2* zr;
    bool two;
  };
  __device__ __forceinline__ Ctx begin(int row) const {
    constexpr int NP = (L + 1) / 2;
    const int pr = pr0 + row, r = pr / NP, k0 = 2 * (pr - r * NP);
    const int j = gcm_seg_row(seg, r);
    Ctx c;
    c.s0 = in + k0 * plane + j * W;
    c.spr = sp + j * W;
    c.o = out + k0 * plane + j * W;
    c.zr = MODE == 2 ? zkeep + (size_t)row * W : nullptr;
    c.two = k0 + 1 < L;
    return c;
  }
  __device__ __forceinline__ double2 load(const Ctx& c, int i) const {
    double x0 = c.s0[i], x1 = c.two ? c.s0[plane + i] : 0.0;
    if (MODE >= 1) {  // su * iph(sp)  (dynamics.py:187)
      const double ph = (c.spr[i] + c.spr[gcm_ip(i, W)]) * 0.5;
      x0 *= ph;
      x1 *= ph;
    }
    return make_double2(x0, x1);
  }
  __device__ __forceinline__ void store(const Ctx& c, int i, double2 v) const {
    c.o[i] = v.x;
    if (c.two) c.o[plane + i] = v.y;
    // in place: the last inverse stage has read every input of this butterfly before it stores, and no other thread
    // touches these positions
    if (MODE == 2) c.zr[i] = v;
  }
};

// PLAN > 0: the radices of the plan are compile-time constants (GcmFixedPlan); 0: runtime switch, any plan
// MB: resident CTAs per SM the register allocation aims at (fixed plans launch 128 threads): 5 = 96 registers, which
// the filter fits without a spill (r2g: 5 CTAs per SM -4 % per launch on the 1440-wide rows; 6 and 8 spill and lose)
template <int L, int MODE, int PLAN, int MB = 5>
__global__ void __launch_bounds__(PLAN > 0 ? 128 : 256, PLAN > 0 ? MB : 2)
pe25f_filter_kernel(GcmGeomDev g, const double* __restrict__ sp, const double* in, double* out, GcmRowSeg seg, int NBAT,
                    size_t bstride2, size_t bstride3, const double* __restrict__ aux_sv, double* __restrict__ aux_part) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  GCM_DYN_SMEM(double2, z);
  constexpr int NP = (L + 1) / 2;
  const int W = g.W;
  const int npr_total = (seg.n1 + seg.n2) * NP;
  const int pr0 = blockIdx.x * NBAT;
  const int nb = npr_total - pr0 < NBAT ? npr_total - pr0 : NBAT;
  PfFilterIO<L, MODE> io{sp + blockIdx.y * bstride2, in + blockIdx.y * bstride3, out + blockIdx.y * bstride3, seg, pr0, W,
                         g.H * W, z};
  if constexpr (PLAN > 0)
    gcm_filter_rows_io_fixed<NP, PLAN>(z, nb, g.plan, g.tws, g.smmzp, seg, pr0, io, threadIdx.x, blockDim.x);
  else
    gcm_filter_rows_io<NP>(z, nb, g.plan, g.tws, g.smmzp, seg, pr0, io, threadIdx.x, blockDim.x);
  if constexpr (MODE == 2) {
    // aflux (dynamics.py:35-46) fused: the two layers of a packed row contribute conv[k0] + conv[k0+1] to pit; the sum
    // over the NP layer pairs of a latitude is taken by the update kernel in a fixed order (pair 0 first), so the band
    // decomposition stays bit-invariant.  part[pair][j][i] lives in the (otherwise unused) sd work field.
    const int H = g.H, plane = H * W;
    const double* __restrict__ spb = sp + blockIdx.y * bstride2;
    const double* __restrict__ svb = aux_sv + blockIdx.y * bstride3;
    double* __restrict__ part = aux_part + blockIdx.y * bstride3;
    const double rdy = g.rdy;
    for (int row = 0; row < nb; ++row) {
      const int pr = pr0 + row, r = pr / NP, pair = pr - r * NP, k0 = 2 * pair;
      const int j = gcm_seg_row(seg, r);
      const int jm = gcm_row(j, -1, H, g.wrap_j), jp = gcm_row(j, 1, H, g.wrap_j);
      const double rdxj = g.rdx_j[j];
      const bool two = k0 + 1 < L;
      const double ds0 = g.c_dsig[k0], ds1 = two ? g.c_dsig[k0 + 1] : 0.0;
      const double2* zr = z + (size_t)row * W;
      for (int i = threadIdx.x; i < W; i += blockDim.x) {
        const double2 pu_c = zr[i], pu_im = zr[gcm_im(i, W)];
        const double sp_c = spb[j * W + i];
        const double pjh = (sp_c + spb[jp * W + i]) * 0.5, pjh_m = (spb[jm * W + i] + sp_c) * 0.5;
        const double pv_c0 = svb[k0 * plane + j * W + i] * pjh, pv_jm0 = svb[k0 * plane + jm * W + i] * pjh_m;
        double conv = ((pu_c.x - pu_im.x) * rdxj + (pv_c0 - pv_jm0) * rdy) * ds0;
        if (two) {
          const double pv_c1 = svb[(k0 + 1) * plane + j * W + i] * pjh, pv_jm1 = svb[(k0 + 1) * plane + jm * W + i] * pjh_m;
          conv += ((pu_c.y - pu_im.y) * rdxj + (pv_c1 - pv_jm1) * rdy) * ds1;
        }
        part[pair * plane + j * W + i] = conv;
      }
    }
  }
}

// The same filter as a PERSISTENT, software-pipelined kernel (compile-time plans only): a CTA walks over units of NBAT
// packed rows; while it transforms unit u, the raw rows of its next unit travel into a staging buffer as 1-D bulk copies
// of whole rows (cp.async.bulk, the TMA unit's contiguous mode, SASS UBLKCP) completing on an mbarrier -- issued as soon as
// the first stage has consumed the staging buffer, so they land under the remaining four stages.  The first stage then
// reads shared memory instead of waiting for HBM: the non-pipelined kernel above exposes that latency once per CTA with
// nothing to overlap it (DRAM 20 % busy, issue 44 %, ncu r03j).
template <int L, int MODE>
struct PfFilterStagedIO {
  const double* __restrict__ sp;
  const double* stage;  // [NBAT][2][W] raw rows of the unit
  double* out;
  GcmRowSeg seg;
  int pr0, W, plane;
  struct Ctx {
    const double* s0;
    const double* spr;
    double* o;
    bool two;
  };
  __device__ __forceinline__ Ctx begin(int row) const {
    constexpr int NP = (L + 1) / 2;
    const int pr = pr0 + row, r = pr / NP, k0 = 2 * (pr - r * NP);
    const int j = gcm_seg_row(seg, r);
    Ctx c;
    c.s0 = stage + (size_t)row * 2 * W;
    c.spr = sp + j * W;
    c.o = out + k0 * plane + j * W;
    c.two = k0 + 1 < L;
    return c;
  }
  __device__ __forceinline__ double2 load(const Ctx& c, int i) const {
    double x0 = c.s0[i], x1 = c.two ? c.s0[W + i] : 0.0;
    if (MODE == 1) {  // su * iph(sp)  (dynamics.py:187)
      const double ph = (c.spr[i] + c.spr[gcm_ip(i, W)]) * 0.5;
      x0 *= ph;
      x1 *= ph;
    }
    return make_double2(x0, x1);
  }
  __device__ __forceinline__ void store(const Ctx& c, int i, double2 v) const {
    c.o[i] = v.x;
    if (c.two) c.o[plane + i] = v.y;
  }
};

template <int L, int MODE, int PLAN>
__global__ void __launch_bounds__(256, 2)
pe25f_filter_pipe_kernel(GcmGeomDev g, const double* __restrict__ sp, const double* in, double* out, GcmRowSeg seg, int NBAT,
                         size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  GCM_DYN_SMEM(unsigned char, smraw);
  constexpr int NP = (L + 1) / 2;
  const int W = g.W, plane = g.H * W;
  GcmMbar* full = reinterpret_cast<GcmMbar*>(smraw);
  double2* z = reinterpret_cast<double2*>(smraw + 128);               // [NBAT][W] complex work rows
  double* stage = reinterpret_cast<double*>(z + (size_t)NBAT * W);   // [NBAT][2][W] raw rows of the next unit
  const int npr_total = (seg.n1 + seg.n2) * NP;
  const int nunits = (npr_total + NBAT - 1) / NBAT;
  const double* inb = in + blockIdx.y * bstride3;
  const int tid = threadIdx.x;
  if (tid == 0) {
    gcm_mbar_init(full, 1);
    gcm_mbar_fence_init();
  }
  __syncthreads();
  auto prefetch = [&](int u) {  // one thread: the raw rows of unit u -> staging
    const int pr0 = u * NBAT;
    const int nb = npr_total - pr0 < NBAT ? npr_total - pr0 : NBAT;
    unsigned bytes = 0;
    for (int row = 0; row < nb; ++row) {
      const int pr = pr0 + row, r = pr / NP, k0 = 2 * (pr - r * NP);
      bytes += (k0 + 1 < L ? 2u : 1u) * (unsigned)W * 8u;
    }
    gcm_fence_proxy_async();
    gcm_mbar_expect_tx(full, bytes);
    for (int row = 0; row < nb; ++row) {
      const int pr = pr0 + row, r = pr / NP, k0 = 2 * (pr - r * NP);
      const int j = gcm_seg_row(seg, r);
      const double* src = inb + (size_t)k0 * plane + (size_t)j * W;
      gcm_bulk_load(stage + (size_t)row * 2 * W, src, (unsigned)W * 8u, full);
      if (k0 + 1 < L) gcm_bulk_load(stage + (size_t)row * 2 * W + W, src + plane, (unsigned)W * 8u, full);
    }
  };
  int u = blockIdx.x;
  if (tid == 0 && u < nunits) prefetch(u);
  unsigned phase = 0;
  for (; u < nunits; u += gridDim.x) {
    const int pr0 = u * NBAT;
    const int nb = npr_total - pr0 < NBAT ? npr_total - pr0 : NBAT;
    PfFilterStagedIO<L, MODE> io{sp + blockIdx.y * bstride2, stage, out + blockIdx.y * bstride3, seg, pr0, W, plane};
    gcm_mbar_wait(full, phase);
    phase ^= 1;
    const int unext = u + gridDim.x;
    auto hook = [&]() {
      if (tid == 0 && unext < nunits) prefetch(unext);
    };
    gcm_filter_rows_io_fixed<NP, PLAN>(z, nb, g.plan, g.tws, g.smmzp, seg, pr0, io, tid, blockDim.x, hook);
  }
}

// launch of the filter kernel whose image matches the plan
template <int L, int MODE, int PLAN>
static int pf_filter_launch_plan(const GcmGeomDev& d, dim3 grid, int threads, size_t smem, void* stream, const double* sp,
                                 const double* in, double* out, GcmRowSeg seg, int nbf, size_t b2, size_t b3,
                                 const double* aux_sv = nullptr, double* aux_part = nullptr) {
  if (PLAN > 0 && threads > 128) threads = 128;  // the fixed-plan kernels are compiled for at most 128 threads
  if constexpr (PLAN > 0) {
    // persistent pipelined kernel (knob 14 = 2; measured slower than the one-unit-per-CTA kernel on every grid, r2f:
    // the filter is bound by dependent-instruction latency at 16 warps per SM, not by its loads)
    const size_t smp = 128 + 2 * smem;
    const int per_sm = (int)((227 * 1024) / (smp + 1024)) < 4 ? (int)((227 * 1024) / (smp + 1024)) : 4;
    if (MODE != 2 && g_gcm_knob[14] == 2 && per_sm >= 1 && (d.W % 2) == 0) {
      int ncta = 148 * per_sm / (int)grid.y;
      ncta = ncta < 1 ? 1 : ncta;
      const dim3 gridp((unsigned)((int)grid.x < ncta ? (int)grid.x : ncta), grid.y);
#ifndef GCM_EMU
      if constexpr (MODE != 2) {
        if (smp > 48 * 1024)
          GCM_CUDA(cudaFuncSetAttribute(pe25f_filter_pipe_kernel<L, MODE, PLAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smp));
      }
#endif
      if constexpr (MODE != 2) {
        GCM_LAUNCH_DEP((pe25f_filter_pipe_kernel<L, MODE, PLAN>), gridp, dim3(threads), smp, stream, d, sp, in, out, seg,
                       nbf, b2, b3);
        GCM_CHECK_LAUNCH();
        return GCM_OK;
      }
    }
  }
#ifndef GCM_EMU
  if constexpr (PLAN == 1 && L == 9) {  // register-budget variants of the 1440-wide plan (knob 15, units digit)
    const int mb = g_gcm_knob[15] % 10;
    if (mb == 4 || mb == 6 || mb == 8) {
      if (threads > 128) return gcm_set_status(GCM_EUNSUP);
#define PF_FILTER_MB(MB_)                                                                                             \
  do {                                                                                                                \
    if (smem > 48 * 1024)                                                                                             \
      GCM_CUDA(cudaFuncSetAttribute(pe25f_filter_kernel<L, MODE, PLAN, MB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)smem));                                                                      \
    GCM_LAUNCH_DEP((pe25f_filter_kernel<L, MODE, PLAN, MB_>), grid, dim3(threads), smem, stream, d, sp, in, out, seg, nbf, \
                   b2, b3, aux_sv, aux_part);                                                                         \
  } while (0)
      if (mb == 4) PF_FILTER_MB(4); else if (mb == 6) PF_FILTER_MB(6); else PF_FILTER_MB(8);
      GCM_CHECK_LAUNCH();
      return GCM_OK;
    }
  }
  if (smem > 48 * 1024)
    GCM_CUDA(cudaFuncSetAttribute(pe25f_filter_kernel<L, MODE, PLAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
#endif
  GCM_LAUNCH_DEP((pe25f_filter_kernel<L, MODE, PLAN>), grid, dim3(threads), smem, stream, d, sp, in, out, seg, nbf, b2, b3,
                 aux_sv, aux_part);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}
template <int L, int MODE>
static int pf_filter_launch(int plan_id, const GcmGeomDev& d, dim3 grid, int threads, size_t smem, void* stream,
                            const double* sp, const double* in, double* out, GcmRowSeg seg, int nbf, size_t b2,
                            size_t b3, const double* aux_sv = nullptr, double* aux_part = nullptr) {
  switch (plan_id) {
    case 1: return pf_filter_launch_plan<L, MODE, 1>(d, grid, threads, smem, stream, sp, in, out, seg, nbf, b2, b3, aux_sv, aux_part);
    case 2: return pf_filter_launch_plan<L, MODE, 2>(d, grid, threads, smem, stream, sp, in, out, seg, nbf, b2, b3, aux_sv, aux_part);
    case 3: return pf_filter_launch_plan<L, MODE, 3>(d, grid, threads, smem, stream, sp, in, out, seg, nbf, b2, b3, aux_sv, aux_part);
    case 4: return pf_filter_launch_plan<L, MODE, 4>(d, grid, threads, smem, stream, sp, in, out, seg, nbf, b2, b3, aux_sv, aux_part);
    default: return pf_filter_launch_plan<L, MODE, 0>(d, grid, threads, smem, stream, sp, in, out, seg, nbf, b2, b3, aux_sv, aux_part);
  }
}

// aflux (dynamics.py:35-46) and p_n (:193-194): one thread per column of the rows of `seg`; needs the filtered spu.
template <int L, bool WRITE_SD>
__global__ void __launch_bounds__(128)
pe25f_aflux_kernel(GcmGeomDev g, const double* __restrict__ p, const double* __restrict__ sp_,
                   const double* __restrict__ sv_, PfWork w, double dt, GcmRowSeg seg, unsigned magicW, size_t bstride2,
                   size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  const int H = g.H, W = g.W, plane = H * W;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gcm_fastdiv(t, magicW), i = t - r * W;
  if (r >= seg.n1 + seg.n2) return;
  const size_t o2 = blockIdx.y * bstride2, o3 = blockIdx.y * bstride3;
  const double* __restrict__ sp = sp_ + o2;
  const double* __restrict__ sv = sv_ + o3;
  const double* __restrict__ spu = w.spu + o3;
  double* __restrict__ sd = w.sd + o3;
  const int j = gcm_seg_row(seg, r);
  const int jm = gcm_row(j, -1, H, g.wrap_j), jp = gcm_row(j, 1, H, g.wrap_j);
  const int c2 = j * W + i, cim = j * W + gcm_im(i, W), cjm = jm * W + i;
  const double sp_c = sp[c2];
  const double pjh = (sp_c + sp[jp * W + i]) * 0.5, pjh_m = (sp[cjm] + sp_c) * 0.5;
  const double rdxj = g.rdx_j[j], rdy = g.rdy;
  double conv[L];
  double pit = 0.0;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const double pu_c = spu[k * plane + c2], pu_im = spu[k * plane + cim];
    const double pv_c = sv[k * plane + c2] * pjh, pv_jm = sv[k * plane + cjm] * pjh_m;
    conv[k] = ((pu_c - pu_im) * rdxj + (pv_c - pv_jm) * rdy) * g.c_dsig[k];
    pit += conv[k];
  }
  if (WRITE_SD) {  // the tiled update kernel rebuilds sd from pit and the fluxes it already holds
    double acc = 0.0;
#pragma unroll
    for (int k = L - 1; k >= 0; --k) {
      acc += conv[k];
      sd[k * plane + c2] = k == 0 ? 0.0 : acc - pit * g.c_sigb[k];  // dynamics.py:42-44
    }
  }
  w.pit[o2 + c2] = pit;
  w.pn[o2 + c2] = p[o2 + c2] - pit * dt;
}

// Hydrostatic columns: one warp per (group of RG rows, chunk of 31 columns) marches south over the group's rows and
// their south neighbour (pf_row_step): pgfu + phiu (unfiltered) -> pgf, fv = phiv + pgv.  Independent of spu.
template <int L, bool PTOP0, int MB = 4>
__global__ void __launch_bounds__(128, MB)
pe25f_hydro_kernel(GcmGeomDev g, PfConst star, PfWork w, GcmRowSeg seg, int RG, size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  const int H = g.H, W = g.W, plane = H * W;
  const int lane = threadIdx.x & 31;
  const int nchunk = (W + 30) / 31, nrows = seg.n1 + seg.n2, ngrp = (nrows + RG - 1) / RG;
  const int task = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= nchunk * ngrp) return;  // whole warps leave together
  const size_t o2 = blockIdx.y * bstride2, o3 = blockIdx.y * bstride3;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ st = star.t + o3;
  double* pgf = w.pgf + o3;
  double* __restrict__ fv = w.fv + o3;
  const int grp = task / nchunk, c = task - grp * nchunk;
  int i = c * 31 + lane;
  const bool own = lane < 31 && i < W;
  i = i % W;
  // rows [j0, j0 + rg), row j0 + rg only as the south neighbour (a two-segment launch has RG = 1)
  const int j0 = gcm_seg_row(seg, grp * RG);
  const int rg = nrows - grp * RG < RG ? nrows - grp * RG : RG;
  double phiA[L], rhoA[L], phiB[L], rhoB[L];
  int jn = j0;
  int j = gcm_row(jn, 1, H, g.wrap_j);
  double spA = pf_row_step<L, PTOP0>(g, sp, st, pgf, fv, plane, jn * W + i, 0, jn, own, true, false, phiA, rhoA, phiA,
                                     rhoA, 0.0, j * W + i);
  double spB = 0.0;
#pragma unroll 1
  for (int r = 1; r <= rg; r += 2) {
    int jnext = gcm_row(j, 1, H, g.wrap_j);
    spB = pf_row_step<L, PTOP0>(g, sp, st, pgf, fv, plane, j * W + i, jn * W + i, j, own, r < rg, true, phiB, rhoB, phiA,
                                rhoA, spA, r + 1 <= rg ? jnext * W + i : -1);
    jn = j;
    j = jnext;
    if (r + 1 <= rg) {
      jnext = gcm_row(j, 1, H, g.wrap_j);
      spA = pf_row_step<L, PTOP0>(g, sp, st, pgf, fv, plane, j * W + i, jn * W + i, j, own, r + 1 < rg, true, phiA, rhoA,
                                  phiB, rhoB, spB, r + 2 <= rg ? jnext * W + i : -1);
      jn = j;
      j = jnext;
    }
  }
}

// Hydrostatic columns, tile form (the default on wide grids; knob 7 = 3 keeps the marching kernel above): a CTA owns RT
// rows x 31 columns (lane 31 of every warp is the east halo
// column, warp RT the south halo row).  Every thread evaluates ONE column -- nine times the threads of the marching
// kernel above and a ninth of its dependent chain (RT + 1 evaluations per RT rows, the same redundancy) --, the east
// neighbour comes by shuffle, the south neighbour through shared memory.  Same expressions, operand for operand, as
// pf_row_step.  One launch per contiguous row segment.
#define PFH_RT 8
template <int L, bool PTOP0, int MB = 2>
__global__ void __launch_bounds__(32 * (PFH_RT + 1), MB)
pe25f_hydro_tile_kernel(GcmGeomDev g, PfConst star, PfWork w, GcmRowSeg seg, size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  GCM_DYN_SMEM(double, xs);  // [PFH_RT + 1 rows][2 L + 1 values][32 lanes]
  const int H = g.H, W = g.W, plane = H * W;
  const int lane = threadIdx.x, ty = threadIdx.y;
  const int nrows = seg.n1;
  const int r0 = blockIdx.y * PFH_RT;                              // first row of the tile within the segment
  const int rt = nrows - r0 < PFH_RT ? nrows - r0 : PFH_RT;        // rows of this tile; row rt is only the south neighbour
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ st = star.t + o3;
  double* pgf = w.pgf + o3;
  double* __restrict__ fv = w.fv + o3;
  int i = blockIdx.x * 31 + lane;
  const bool own = lane < 31 && i < W;
  i = i % W;
  // stored row of tile row ty: consecutive rows from seg.a (periodic on a whole grid, halo rows on a band)
  int j = seg.a + r0 + ty;
  if (g.wrap_j) j = j >= H ? j - H : j;
  const bool row_ok = ty <= rt && j < H;
  const int c2 = j * W + i;
  double phi[L], rho[L];
  double sp_c = 1.0;
  double* xr = xs + (size_t)ty * (2 * L + 1) * 32;
  if (row_ok) {
    sp_c = sp[c2];
    pf_column<L, PTOP0>(g, sp_c, g.hmap[c2], st + c2, plane, phi, rho);
    xr[lane] = sp_c;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      xr[(1 + k) * 32 + lane] = phi[k];
      xr[(1 + L + k) * 32 + lane] = rho[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < L; ++k) { phi[k] = 0.0; rho[k] = 1.0; }
  }
  // pgfu + phiu of row j (dynamics.py:159, :162-165): every lane of the warp takes part in the shuffles
  {
    const double sp_e = __shfl_down_sync(0xffffffffu, sp_c, 1);
    const double rdxj = row_ok ? g.rdx_j[j] : 0.0;
    const double psum = sp_c + sp_e, gradp = (sp_e - sp_c) * rdxj;
    const double a_u = psum * gradp;       // (p_c + p_e) dp/dx
    const double b_u = psum * 0.5 * rdxj;  // iph(p) / dx
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const double phi_e = __shfl_down_sync(0xffffffffu, phi[k], 1);
      const double rho_e = __shfl_down_sync(0xffffffffu, rho[k], 1);
      const double x = g.c_sig[k] * a_u * gcm_rcp(rho[k] + rho_e) + b_u * (phi_e - phi[k]);
      if (own && row_ok && ty < rt) pgf[k * plane + c2] = x;
    }
  }
  __syncthreads();
  if (own && row_ok && ty < rt) {  // fv = phiv + pgv of this row against its south neighbour (dynamics.py:160, :167-169)
    const double* xsn = xs + (size_t)(ty + 1) * (2 * L + 1) * 32;
    const double sp_s = xsn[lane];
    const double rdy = g.rdy;
    const double psum = sp_c + sp_s;
    const double a_v = psum * ((sp_s - sp_c) * rdy);  // (p_c + p_jp) dp/dy
    const double b_v = psum * 0.5 * rdy;              // jph(p) / dy
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const double phi_s = xsn[(1 + k) * 32 + lane], rho_s = xsn[(1 + L + k) * 32 + lane];
      fv[k * plane + c2] = g.c_sig[k] * a_v * gcm_rcp(rho[k] + rho_s) + b_v * (phi_s - phi[k]);
    }
  }
}

// Hydrostatic columns on narrow grids (W < 62, the ensemble members): a 31-column warp chunk would leave most lanes
// idle, so a CTA takes G whole row groups (G * W threads, one column each) and the east neighbour comes through
// shared memory instead of a shuffle.  Same arithmetic as pe25f_hydro_kernel.
template <int L, bool PTOP0>
__global__ void __launch_bounds__(256, 2)
pe25f_hydro_narrow_kernel(GcmGeomDev g, PfConst star, PfWork w, GcmRowSeg seg, int RG, int G, size_t bstride2,
                          size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  GCM_DYN_SMEM(double, xs);  // [2 parities][2 L + 1 values][blockDim.x]
  const int H = g.H, W = g.W, plane = H * W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int nrows = seg.n1 + seg.n2, ngrp = (nrows + RG - 1) / RG;
  const int gl = tid / W, i = tid - gl * W;
  const int grp = blockIdx.x * G + gl;
  const bool valid = gl < G && grp < ngrp;
  const size_t o2 = blockIdx.y * bstride2, o3 = blockIdx.y * bstride3;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ st = star.t + o3;
  double* pgf = w.pgf + o3;
  double* __restrict__ fv = w.fv + o3;
  const int te = i + 1 < W ? tid + 1 : tid - (W - 1);  // thread of the east neighbour (periodic within the row)
  const int j0 = valid ? gcm_seg_row(seg, grp * RG) : 0;
  const int rg = valid ? (nrows - grp * RG < RG ? nrows - grp * RG : RG) : 0;
  double phi_n[L], rho_n[L];
  double sp_n = 0.0;
  int j = j0, jn = j0;
  for (int r = 0; r <= RG; ++r) {  // the same trip count for every thread of the block: barriers inside
    double* xr = xs + (r & 1) * (2 * L + 1) * nthr;
    const bool row_ok = valid && r <= rg;
    double phi[L], rho[L];
    double sp_c = 1.0;
    const int c2 = j * W + i;
    if (row_ok) {
      sp_c = sp[c2];
      pf_column<L, PTOP0>(g, sp_c, g.hmap[c2], st + c2, plane, phi, rho);
      xr[tid] = sp_c;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        xr[(1 + k) * nthr + tid] = phi[k];
        xr[(1 + L + k) * nthr + tid] = rho[k];
      }
    }
    __syncthreads();
    if (row_ok && r < rg) {  // pgfu + phiu of row j (dynamics.py:159, :162-165)
      const double sp_e = xr[te];
      const double rdxj = g.rdx_j[j];
      const double psum = sp_c + sp_e, gradp = (sp_e - sp_c) * rdxj;
      const double a_u = psum * gradp, b_u = psum * 0.5 * rdxj;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const double phi_e = xr[(1 + k) * nthr + te], rho_e = xr[(1 + L + k) * nthr + te];
        pgf[k * plane + c2] = g.c_sig[k] * a_u * gcm_rcp(rho[k] + rho_e) + b_u * (phi_e - phi[k]);
      }
    }
    if (row_ok && r >= 1) {  // fv = phiv + pgv of the row to the north (dynamics.py:160, :167-169)
      const int cn = jn * W + i;
      const double rdy = g.rdy;
      const double psum = sp_n + sp_c;
      const double a_v = psum * ((sp_c - sp_n) * rdy), b_v = psum * 0.5 * rdy;
#pragma unroll
      for (int k = 0; k < L; ++k)
        fv[k * plane + cn] = g.c_sig[k] * a_v * gcm_rcp(rho_n[k] + rho[k]) + b_v * (phi[k] - phi_n[k]);
    }
    if (row_ok) {
#pragma unroll
      for (int k = 0; k < L; ++k) {
        phi_n[k] = phi[k];
        rho_n[k] = rho[k];
      }
      sp_n = sp_c;
      jn = j;
      j = gcm_row(j, 1, H, g.wrap_j);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// U, direct loads: one thread per column, k loop (widths that are not a multiple of 32; short rows run flat)
// ---------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(128, 4)
pe25f_update_kernel(GcmGeomDev g, PfConst base, PfConst star, PfMut out, PfWork w, double dt, GcmRowSeg seg, int pfd,
                    unsigned flatW, size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  const int H = g.H, W = g.W, plane = H * W;
  int i, r;
  if (flatW) {  // short rows: threads run over the (row, column) pairs of the launch in row-major order
    const int t = blockIdx.x * (blockDim.x * blockDim.y) + threadIdx.y * blockDim.x + threadIdx.x;
    r = gcm_fastdiv(t, flatW);
    i = t - r * W;
  } else {
    i = blockIdx.x * blockDim.x + threadIdx.x;
    r = blockIdx.y * blockDim.y + threadIdx.y;
  }
  if (i >= W || r >= seg.n1 + seg.n2) return;
  const int j = gcm_seg_row(seg, r);
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* __restrict__ p = base.p + o2;
  const double* __restrict__ u = base.u + o3;
  const double* __restrict__ v = base.v + o3;
  const double* __restrict__ t = base.t + o3;
  const double* __restrict__ q = base.q + o3;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ su = star.u + o3;
  const double* __restrict__ sv = star.v + o3;
  const double* __restrict__ st = star.t + o3;
  const double* __restrict__ sq = star.q + o3;
  const double* __restrict__ spu = w.spu + o3;
  const double* __restrict__ sd = w.sd + o3;
  const double* __restrict__ pgf = w.pgf + o3;
  const double* __restrict__ fv = w.fv + o3;
  const double* __restrict__ pn = w.pn + o2;
  double* __restrict__ ou = out.u + o3;
  double* __restrict__ ov = out.v + o3;
  double* __restrict__ ot = out.t + o3;
  double* __restrict__ oq = out.q + o3;

  const int wrap = g.wrap_j;
  const int jm = gcm_row(j, -1, H, wrap), jp = gcm_row(j, 1, H, wrap), jpp = gcm_row(jp, 1, H, wrap);
  const int im = gcm_im(i, W), ip = gcm_ip(i, W);
  const double rdxj = g.rdx_j[j], rdxh = g.rdx_h[j], rdy = g.rdy;
  // element offsets of the seven stencil columns within a layer; they advance by one plane per layer
  int e_c = j * W + i, e_im = j * W + im, e_ip = j * W + ip, e_jp = jp * W + i, e_jm = jm * W + i,
      e_jp_im = jp * W + im, e_jm_ip = jm * W + ip;

  // per-column (2-D) factors
  const double p_c = p[e_c], p_ip = p[e_ip], p_jp = p[e_jp];
  const double pn_c = pn[e_c], pn_ip = pn[e_ip], pn_jp = pn[e_jp];
  const double pu_fac = (p_c + p_ip) * 0.5, pv_fac = (p_c + p_jp) * 0.5;                    // calc_pu / calc_pv
  const double r_pnu = 1.0 / ((pn_c + pn_ip) * 0.5), r_pnv = 1.0 / ((pn_c + pn_jp) * 0.5);  // un_pu / un_pv
  const double r_pn = 1.0 / pn_c;
  const double sp_c = sp[e_c], sp_ip = sp[e_ip], sp_jp = sp[e_jp], sp_jm = sp[e_jm];
  const double a_c = (sp_c + sp_jp) * 0.5;                   // jph(sp) at (j, i)
  const double a_ip = (sp_ip + sp[jp * W + ip]) * 0.5;       // (j, i+1)
  const double a_jm = (sp_jm + sp_c) * 0.5;                  // (j-1, i)
  const double a_jm_ip = (sp[e_jm_ip] + sp_ip) * 0.5;        // (j-1, i+1)
  const double a_jp = (sp_jp + sp[jpp * W + i]) * 0.5;       // (j+1, i)
  const bool zero_v = j == g.zero_v_row || j == g.zero_v_row2;

  // vertical neighbours and interface fluxes carried in registers (advec_sig, dynamics.py:49-52)
  double u_k = su[e_c], v_k = sv[e_c], t_k = st[e_c], q_k = sq[e_c];
  double sd_c = sd[e_c], sd_ip = sd[e_ip], sd_jp = sd[e_jp];
  // flux through the bottom of layer 0 pairs layer 0 with layer L-1 (np.roll) times sd[0] = 0
  double fu, fv_, ft, fq;
  {
    const int top = (L - 1) * plane + e_c;
    fu = (u_k + su[top]) * 0.5 * ((sd_c + sd_ip) * 0.5);
    fv_ = (v_k + sv[top]) * 0.5 * ((sd_c + sd_jp) * 0.5);
    ft = (t_k + st[top]) * 0.5 * sd_c;
    fq = (q_k + sq[top]) * 0.5 * sd_c;
  }
  const double fu0 = fu, fv0 = fv_, ft0 = ft, fq0 = fq;

  // ask L1 for the lines of layer k + pfd while layer k is computed (no registers held, no scoreboard)
  auto prefetch_layer = [&](int ec, int ejp, int ejm) {
    gcm_prefetch_l1(su + ec); gcm_prefetch_l1(su + ejp); gcm_prefetch_l1(su + ejm);
    gcm_prefetch_l1(sv + ec); gcm_prefetch_l1(sv + ejp); gcm_prefetch_l1(sv + ejm);
    gcm_prefetch_l1(st + ec); gcm_prefetch_l1(st + ejp); gcm_prefetch_l1(st + ejm);
    gcm_prefetch_l1(sq + ec); gcm_prefetch_l1(sq + ejp); gcm_prefetch_l1(sq + ejm);
    gcm_prefetch_l1(spu + ec); gcm_prefetch_l1(spu + ejp);
    gcm_prefetch_l1(sd + ec); gcm_prefetch_l1(sd + ejp);
    gcm_prefetch_l1(pgf + ec); gcm_prefetch_l1(fv + ec);
    gcm_prefetch_l1(u + ec); gcm_prefetch_l1(v + ec); gcm_prefetch_l1(t + ec); gcm_prefetch_l1(q + ec);
  };
  for (int k = 0; k < pfd && k < L; ++k) prefetch_layer(e_c + k * plane, e_jp + k * plane, e_jm + k * plane);
  bool bad = gcm_not_finite(pn_c);

#pragma unroll
  for (int k = 0; k < L; ++k) {
    if (pfd > 0 && k + pfd < L) prefetch_layer(e_c + pfd * plane, e_jp + pfd * plane, e_jm + pfd * plane);
    // fluxes through the top of layer k
    double fu_n = fu0, fv_n = fv0, ft_n = ft0, fq_n = fq0;
    double u_kp = 0.0, v_kp = 0.0, t_kp = 0.0, q_kp = 0.0;
    if (k + 1 < L) {
      u_kp = su[e_c + plane]; v_kp = sv[e_c + plane]; t_kp = st[e_c + plane]; q_kp = sq[e_c + plane];
      sd_c = sd[e_c + plane]; sd_ip = sd[e_ip + plane]; sd_jp = sd[e_jp + plane];
      fu_n = (u_kp + u_k) * 0.5 * ((sd_c + sd_ip) * 0.5);
      fv_n = (v_kp + v_k) * 0.5 * ((sd_c + sd_jp) * 0.5);
      ft_n = (t_kp + t_k) * 0.5 * sd_c;
      fq_n = (q_kp + q_k) * 0.5 * sd_c;
    }
    const double rds = g.c_rdsig[k];
    const double dus = (fu_n - fu) * rds, dvs = (fv_n - fv_) * rds;      // -(F_k - F_k+1) / dsig
    const double ads_t = (ft_n - ft) * rds, ads_q = (fq_n - fq) * rds;

    // horizontal neighbours
    const double u_im = su[e_im], u_ip = su[e_ip], u_jp = su[e_jp], u_jm = su[e_jm];
    const double v_im = sv[e_im], v_ip = sv[e_ip], v_jp = sv[e_jp], v_jm = sv[e_jm], v_jm_ip = sv[e_jm_ip];
    const double pu_c = spu[e_c], pu_im = spu[e_im], pu_ip = spu[e_ip], pu_jp = spu[e_jp], pu_jp_im = spu[e_jp_im];
    const double pv_c = v_k * a_c, pv_ip = v_ip * a_ip, pv_jm = v_jm * a_jm, pv_jm_ip = v_jm_ip * a_jm_ip,
                 pv_jp = v_jp * a_jp;

    // advec_m_pu (dynamics.py:55-108); (a/2)(b/2) = ab/4 exactly
    const double puum = (u_k + u_im) * (pu_c + pu_im), puup = (u_ip + u_k) * (pu_ip + pu_c);
    const double puvp = (pv_c + pv_ip) * (u_k + u_jp), puvm = (pv_jm + pv_jm_ip) * (u_jm + u_k);
    const double dut = ((puum - puup) * rdxj + (puvm - puvp) * rdy) * 0.25;
    const double pvvm = (v_k + v_jm) * (pv_c + pv_jm), pvvp = (v_jp + v_k) * (pv_jp + pv_c);
    const double pvup = (v_k + v_ip) * (pu_c + pu_jp), pvum = (v_im + v_k) * (pu_im + pu_jp_im);
    const double dvt = ((pvvm - pvvp) * rdy + (pvum - pvup) * rdxh) * 0.25;

    const double pu_n = u[e_c] * pu_fac - (dut + dus + pgf[e_c]) * dt;   // dynamics.py:206
    const double pv_n = v[e_c] * pv_fac - (dvt + dvs + fv[e_c]) * dt;    // dynamics.py:207
    const double u_n = pu_n * r_pnu;
    ou[e_c] = u_n;
    double v_n = pv_n * r_pnv;
    if (zero_v) v_n *= 0.0;  // dynamics.py:222
    ov[e_c] = v_n;

    // tracers: advec_t (dynamics.py:174-181) + advec_sig, flux form (dynamics.py:214, :219)
    double t_n, q_n;
    {
      const double x_ip = st[e_ip], x_im = st[e_im], x_jp = st[e_jp], x_jm = st[e_jm];
      const double adv = ((pu_c * (t_k + x_ip) - pu_im * (x_im + t_k)) * rdxj +
                          (pv_c * (t_k + x_jp) - pv_jm * (x_jm + t_k)) * rdy) * 0.5;
      t_n = (t[e_c] * p_c - (adv + ads_t) * dt) * r_pn;
      ot[e_c] = t_n;
    }
    {
      const double x_ip = sq[e_ip], x_im = sq[e_im], x_jp = sq[e_jp], x_jm = sq[e_jm];
      const double adv = ((pu_c * (q_k + x_ip) - pu_im * (x_im + q_k)) * rdxj +
                          (pv_c * (q_k + x_jp) - pv_jm * (x_jm + q_k)) * rdy) * 0.5;
      q_n = (q[e_c] * p_c - (adv + ads_q) * dt) * r_pn;
      oq[e_c] = q_n;
    }
    bad |= gcm_not_finite((u_n + v_n) + (t_n + q_n));
    fu = fu_n; fv_ = fv_n; ft = ft_n; fq = fq_n;
    u_k = u_kp; v_k = v_kp; t_k = t_kp; q_k = q_kp;
    e_c += plane; e_im += plane; e_ip += plane; e_jp += plane; e_jm += plane; e_jp_im += plane; e_jm_ip += plane;
  }
  out.p[o2 + j * W + i] = pn_c;
  gcm_flag_nonfinite(g.nonfinite, bad);
}

// ---------------------------------------------------------------------------------------------------
// U, one thread per CELL: the direct-load update for launches too small to fill the chip with one thread per column
// (the 72 x 46 grid is 26 CTAs of columns marching nine layers one after the other: a chain of nine dependent load
// round trips on 26 of 148 SMs).  Every (k, j, i) is its own thread and evaluates the two sigma-interface fluxes of its
// layer itself instead of carrying them up the column: the same expressions, operand for operand, as
// pe25f_update_kernel, 9 x the threads, a ninth of the chain.  Flat over (row, column); blockIdx.y = layer.
// ---------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(128)
pe25f_update_cell_kernel(GcmGeomDev g, PfConst base, PfConst star, PfMut out, PfWork w, double dt, GcmRowSeg seg,
                         unsigned flatW, size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  const int H = g.H, W = g.W, plane = H * W;
  const int tt = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gcm_fastdiv(tt, flatW), i = tt - r * W;
  if (r >= seg.n1 + seg.n2) return;
  const int k = blockIdx.y;
  const int j = gcm_seg_row(seg, r);
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* __restrict__ p = base.p + o2;
  const double* __restrict__ u = base.u + o3;
  const double* __restrict__ v = base.v + o3;
  const double* __restrict__ t = base.t + o3;
  const double* __restrict__ q = base.q + o3;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ su = star.u + o3;
  const double* __restrict__ sv = star.v + o3;
  const double* __restrict__ st = star.t + o3;
  const double* __restrict__ sq = star.q + o3;
  const double* __restrict__ spu = w.spu + o3;
  const double* __restrict__ sd = w.sd + o3;
  const double* __restrict__ pgf = w.pgf + o3;
  const double* __restrict__ fv = w.fv + o3;
  const double* __restrict__ pn = w.pn + o2;

  const int wrap = g.wrap_j;
  const int jm = gcm_row(j, -1, H, wrap), jp = gcm_row(j, 1, H, wrap), jpp = gcm_row(jp, 1, H, wrap);
  const int im = gcm_im(i, W), ip = gcm_ip(i, W);
  const double rdxj = g.rdx_j[j], rdxh = g.rdx_h[j], rdy = g.rdy;
  const int c2 = j * W + i, c2_ip = j * W + ip, c2_jp = jp * W + i, c2_jm = jm * W + i;
  const int kd = k * plane;
  const int e_c = kd + c2, e_im = kd + j * W + im, e_ip = kd + c2_ip, e_jp = kd + c2_jp, e_jm = kd + c2_jm,
            e_jp_im = kd + jp * W + im, e_jm_ip = kd + jm * W + ip;

  // per-column (2-D) factors
  const double p_c = p[c2], p_ip = p[c2_ip], p_jp = p[c2_jp];
  const double pn_c = pn[c2], pn_ip = pn[c2_ip], pn_jp = pn[c2_jp];
  const double pu_fac = (p_c + p_ip) * 0.5, pv_fac = (p_c + p_jp) * 0.5;                    // calc_pu / calc_pv
  const double r_pnu = 1.0 / ((pn_c + pn_ip) * 0.5), r_pnv = 1.0 / ((pn_c + pn_jp) * 0.5);  // un_pu / un_pv
  const double r_pn = 1.0 / pn_c;
  const double sp_c = sp[c2], sp_ip = sp[c2_ip], sp_jp = sp[c2_jp], sp_jm = sp[c2_jm];
  const double a_c = (sp_c + sp_jp) * 0.5;                   // jph(sp) at (j, i)
  const double a_ip = (sp_ip + sp[jp * W + ip]) * 0.5;       // (j, i+1)
  const double a_jm = (sp_jm + sp_c) * 0.5;                  // (j-1, i)
  const double a_jm_ip = (sp[jm * W + ip] + sp_ip) * 0.5;    // (j-1, i+1)
  const double a_jp = (sp_jp + sp[jpp * W + i]) * 0.5;       // (j+1, i)
  const bool zero_v = j == g.zero_v_row || j == g.zero_v_row2;

  // advec_sig (dynamics.py:49-52): flux through the bottom of layer k pairs it with layer k - 1 (layer 0 with L - 1,
  // np.roll, times sd[0] = 0), flux through its top pairs layer k + 1 with it (the top of layer L - 1 is the bottom of
  // layer 0 again)
  const int kb = (k == 0 ? L - 1 : k - 1) * plane, kn = (k + 1 < L ? k + 1 : 0) * plane;
  const double u_k = su[e_c], v_k = sv[e_c], t_k = st[e_c], q_k = sq[e_c];
  double fu, fv_, ft, fq, fu_n, fv_n, ft_n, fq_n;
  {
    const double sd_c = sd[e_c], sd_ip = sd[e_ip], sd_jp = sd[e_jp];
    fu = (u_k + su[kb + c2]) * 0.5 * ((sd_c + sd_ip) * 0.5);
    fv_ = (v_k + sv[kb + c2]) * 0.5 * ((sd_c + sd_jp) * 0.5);
    ft = (t_k + st[kb + c2]) * 0.5 * sd_c;
    fq = (q_k + sq[kb + c2]) * 0.5 * sd_c;
  }
  {
    const double sd_c = sd[kn + c2], sd_ip = sd[kn + c2_ip], sd_jp = sd[kn + c2_jp];
    fu_n = (su[kn + c2] + u_k) * 0.5 * ((sd_c + sd_ip) * 0.5);
    fv_n = (sv[kn + c2] + v_k) * 0.5 * ((sd_c + sd_jp) * 0.5);
    ft_n = (st[kn + c2] + t_k) * 0.5 * sd_c;
    fq_n = (sq[kn + c2] + q_k) * 0.5 * sd_c;
  }
  const double rds = g.rdsig[k];
  const double dus = (fu_n - fu) * rds, dvs = (fv_n - fv_) * rds;      // -(F_k - F_k+1) / dsig
  const double ads_t = (ft_n - ft) * rds, ads_q = (fq_n - fq) * rds;

  // horizontal neighbours
  const double u_im = su[e_im], u_ip = su[e_ip], u_jp = su[e_jp], u_jm = su[e_jm];
  const double v_im = sv[e_im], v_ip = sv[e_ip], v_jp = sv[e_jp], v_jm = sv[e_jm], v_jm_ip = sv[e_jm_ip];
  const double pu_c = spu[e_c], pu_im = spu[e_im], pu_ip = spu[e_ip], pu_jp = spu[e_jp], pu_jp_im = spu[e_jp_im];
  const double pv_c = v_k * a_c, pv_ip = v_ip * a_ip, pv_jm = v_jm * a_jm, pv_jm_ip = v_jm_ip * a_jm_ip,
               pv_jp = v_jp * a_jp;

  // advec_m_pu (dynamics.py:55-108); (a/2)(b/2) = ab/4 exactly
  const double puum = (u_k + u_im) * (pu_c + pu_im), puup = (u_ip + u_k) * (pu_ip + pu_c);
  const double puvp = (pv_c + pv_ip) * (u_k + u_jp), puvm = (pv_jm + pv_jm_ip) * (u_jm + u_k);
  const double dut = ((puum - puup) * rdxj + (puvm - puvp) * rdy) * 0.25;
  const double pvvm = (v_k + v_jm) * (pv_c + pv_jm), pvvp = (v_jp + v_k) * (pv_jp + pv_c);
  const double pvup = (v_k + v_ip) * (pu_c + pu_jp), pvum = (v_im + v_k) * (pu_im + pu_jp_im);
  const double dvt = ((pvvm - pvvp) * rdy + (pvum - pvup) * rdxh) * 0.25;

  const double pu_n = u[e_c] * pu_fac - (dut + dus + pgf[e_c]) * dt;   // dynamics.py:206
  const double pv_n = v[e_c] * pv_fac - (dvt + dvs + fv[e_c]) * dt;    // dynamics.py:207
  const double u_n = pu_n * r_pnu;
  out.u[o3 + e_c] = u_n;
  double v_n = pv_n * r_pnv;
  if (zero_v) v_n *= 0.0;  // dynamics.py:222
  out.v[o3 + e_c] = v_n;

  // tracers: advec_t (dynamics.py:174-181) + advec_sig, flux form (dynamics.py:214, :219)
  double t_n, q_n;
  {
    const double x_ip = st[e_ip], x_im = st[e_im], x_jp = st[e_jp], x_jm = st[e_jm];
    const double adv = ((pu_c * (t_k + x_ip) - pu_im * (x_im + t_k)) * rdxj +
                        (pv_c * (t_k + x_jp) - pv_jm * (x_jm + t_k)) * rdy) * 0.5;
    t_n = (t[e_c] * p_c - (adv + ads_t) * dt) * r_pn;
    out.t[o3 + e_c] = t_n;
  }
  {
    const double x_ip = sq[e_ip], x_im = sq[e_im], x_jp = sq[e_jp], x_jm = sq[e_jm];
    const double adv = ((pu_c * (q_k + x_ip) - pu_im * (x_im + q_k)) * rdxj +
                        (pv_c * (q_k + x_jp) - pv_jm * (x_jm + q_k)) * rdy) * 0.5;
    q_n = (q[e_c] * p_c - (adv + ads_q) * dt) * r_pn;
    out.q[o3 + e_c] = q_n;
  }
  if (k == 0) out.p[o2 + c2] = pn_c;
  gcm_flag_nonfinite(g.nonfinite, gcm_not_finite((u_n + v_n) + (t_n + q_n) + pn_c));
}

// ---------------------------------------------------------------------------------------------------
// U, tiled: the same update with every operand staged in shared memory by 8-byte asynchronous copies (LDGSTS), three
// layers in flight.  A CTA owns a 32 x 4 (i x j) tile; per layer it stages the 34 x 6 halo tile of su, sv, st, sq, spu,
// sd (each thread its own column, 76 threads one halo-ring element each, periodic wrap resolved per element).  The
// stencil reads of the k loop hit shared memory only: the HBM latency is carried by the copy queue instead of by
// registers and resident warps; the six values a cell reads once (pgf, fv, u, v, t, q) come straight from global
// memory.  Needs W % 32 == 0; one launch per row segment.
// ---------------------------------------------------------------------------------------------------
// value of knob 16 that fuses aflux into the filter: 1 = opt-in (the separate aflux kernel is the default), 0 = default on
#define GCM_FUSE_AFLUX_ON 1
#define PFT_TI 32
#define PFT_NS 3  // layers in flight
#define PFT_ROW (PFT_TI + 4)  // [pad, west halo, 32 columns, east halo, pad]: the interior starts 16-byte aligned
#define PFT_NF 5  // staged fields: su, sv, st, sq, spu

// TI: tile width = 32, or 36 (the 36-wide ensemble members: one tile spans the row, both seams in the same tile)
// PARTS: pit comes as NP partial sums of conv, one per layer pair (the aflux pass fused into the filter kernel, MODE 2),
// in the sd work field; pit = their sum in pair order, p_n = p - pit dt (dynamics.py:39-40, :193-194) are formed here
template <int L, int PFT_TJ, int MB = 512 / (PFT_TI * PFT_TJ), int TI = PFT_TI, bool PARTS = false>
__global__ void __launch_bounds__(TI * PFT_TJ, MB)
pe25f_update_tiled_kernel(GcmGeomDev g, PfConst base, PfConst star, PfMut out, PfWork w, double dt, GcmRowSeg seg,
                          size_t bstride2, size_t bstride3) {
  constexpr int TROW = TI + 4;  // [pad, west halo, TI columns, east halo, pad]: the interior starts 16-byte aligned
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  constexpr int pfd = 2;  // L1 prefetch distance (layers) of the once-read fields: 1..3 measured alike, off costs 14 %
  GCM_DYN_SMEM(double, sm);
  constexpr int TILE = (PFT_TJ + 2) * TROW, PFT_STAGE = PFT_NF * TILE;  // doubles per field tile / stage
  const int H = g.H, W = g.W, plane = H * W, wrap = g.wrap_j;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TI + tx;
  const int i = blockIdx.x * TI + tx;
  const int r = blockIdx.y * PFT_TJ + ty;
  const bool active = r < seg.n1;
  const int j0 = seg.a + blockIdx.y * PFT_TJ;  // first row of the tile
  // Stored row of a tile row: periodic on a whole grid; on a band clamped into the stored rows.  Rows an active
  // thread reads are real rows by construction (the caller's halo contract); the clamp only keeps the unused corners
  // of a partial tile inside the arrays.
  auto rowc = [&](int x) { return wrap ? ((x % H) + H) % H : (x < 0 ? 0 : (x >= H ? H - 1 : x)); };
  const int j = rowc(j0 + ty);
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* __restrict__ p = base.p + o2;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ pn = w.pn + o2;
  const double* __restrict__ pit = w.pit + o2;
  // fields 0..4 staged with halo (su, sv, st, sq, spu); 6..11 read once per cell (pgf, fv, u, v, t, q)
  const double* fld[12] = {star.u + o3, star.v + o3, star.t + o3, star.q + o3, w.spu + o3, nullptr,
                           w.pgf + o3,  w.fv + o3,   base.u + o3, base.v + o3, base.t + o3, base.q + o3};
  double* __restrict__ ou = out.u + o3;
  double* __restrict__ ov = out.v + o3;
  double* __restrict__ ot = out.t + o3;
  double* __restrict__ oq = out.q + o3;

  const int jm = rowc(j0 + ty - 1), jp = rowc(j0 + ty + 1), jpp = rowc(j0 + ty + 2);
  const int im = gcm_im(i, W), ip = gcm_ip(i, W);
  const int e_c = j * W + i;
  // halo ring of the tile: north row, south row, west column, east column -> (tile row, tile column, global offset)
  int hr = 0, hc = 0, e_h = 0;
  constexpr int RING_ROW = TI + 2;  // a ring row spans columns -1 .. TI
  const bool has_halo = tid < 2 * RING_ROW + 2 * PFT_TJ;
  if (has_halo) {
    int rr, cc;  // tile coordinates, -1 .. TJ and -1 .. TI
    if (tid < RING_ROW) { rr = -1; cc = tid - 1; }
    else if (tid < 2 * RING_ROW) { rr = PFT_TJ; cc = tid - RING_ROW - 1; }
    else if (tid < 2 * RING_ROW + PFT_TJ) { rr = tid - 2 * RING_ROW; cc = -1; }
    else { rr = tid - 2 * RING_ROW - PFT_TJ; cc = TI; }
    const int gj = rowc(j0 + rr);
    int gi = blockIdx.x * TI + cc;
    gi = gi < 0 ? gi + W : (gi >= W ? gi - W : gi);
    hr = rr + 1;
    hc = cc + 2;
    e_h = gj * W + gi;
  }
  const int t_c = (ty + 1) * TROW + (tx + 2);  // own position in a tile
  const int t_h = hr * TROW + hc;

  auto issue = [&](int k, int s) {  // stage layer k into stage s
    double* st = sm + s * PFT_STAGE;
    const int off = k * plane;
    if ((tx & 1) == 0) {  // two columns per copy: even columns are 16-byte aligned in the tile and in the field
#pragma unroll
      for (int f = 0; f < PFT_NF; ++f) gcm_cp_async16(st + f * TILE + t_c, fld[f] + off + e_c);
    }
    if (has_halo) {
#pragma unroll
      for (int f = 0; f < PFT_NF; ++f) gcm_cp_async8(st + f * TILE + t_h, fld[f] + off + e_h);
    }
    gcm_cp_async_commit();
  };
  for (int k = 0; k < pfd && k < L; ++k) {
#pragma unroll
    for (int f = 6; f < 12; ++f) gcm_prefetch_l1(fld[f] + k * plane + e_c);
  }
  // layers 0 .. NS-2 are issued up front; iteration k then issues layer k + NS - 1 into the stage layer k - 1 left
#pragma unroll
  for (int k = 0; k < PFT_NS - 1; ++k)
    if (k < L) issue(k, k);

  // per-column (2-D) factors while the first layers are on their way
  const double rdxj = g.rdx_j[j], rdxh = g.rdx_h[j], rdy = g.rdy;
  const double p_c = p[e_c], p_ip = p[j * W + ip], p_jp = p[jp * W + i];
  double pit_c, pit_ip, pit_jp, pn_c, pn_ip, pn_jp;
  if constexpr (PARTS) {
    constexpr int NP = (L + 1) / 2;
    const double* __restrict__ part = w.sd + o3;
    pit_c = part[e_c]; pit_ip = part[j * W + ip]; pit_jp = part[jp * W + i];
#pragma unroll
    for (int pr = 1; pr < NP; ++pr) {
      pit_c += part[pr * plane + e_c];
      pit_ip += part[pr * plane + j * W + ip];
      pit_jp += part[pr * plane + jp * W + i];
    }
    pn_c = p_c - pit_c * dt; pn_ip = p_ip - pit_ip * dt; pn_jp = p_jp - pit_jp * dt;
  } else {
    pit_c = pit[e_c]; pit_ip = pit[j * W + ip]; pit_jp = pit[jp * W + i];
    pn_c = pn[e_c]; pn_ip = pn[j * W + ip]; pn_jp = pn[jp * W + i];
  }
  const double pu_fac = (p_c + p_ip) * 0.5, pv_fac = (p_c + p_jp) * 0.5;                    // calc_pu / calc_pv
  const double r_pnu = 1.0 / ((pn_c + pn_ip) * 0.5), r_pnv = 1.0 / ((pn_c + pn_jp) * 0.5);  // un_pu / un_pv
  const double r_pn = 1.0 / pn_c;
  const double sp_c = sp[e_c], sp_ip = sp[j * W + ip], sp_jp = sp[jp * W + i], sp_jm = sp[jm * W + i];
  const double a_c = (sp_c + sp_jp) * 0.5;                     // jph(sp) at (j, i)
  const double a_ip = (sp_ip + sp[jp * W + ip]) * 0.5;         // (j, i+1)
  const double a_jm = (sp_jm + sp_c) * 0.5;                    // (j-1, i)
  const double a_jm_ip = (sp[jm * W + ip] + sp_ip) * 0.5;      // (j-1, i+1)
  const double a_jp = (sp_jp + sp[jpp * W + i]) * 0.5;         // (j+1, i)
  const bool zero_v = j == g.zero_v_row || j == g.zero_v_row2;
  // sd (dynamics.py:42-44) of the three columns the vertical fluxes need -- (j, i), (j, i+1), (j+1, i) -- is rebuilt
  // from pit and the running sums of conv, whose operands the horizontal advection loads anyway:
  //   sd[k] = sum_{l >= k} conv[l] - pit sigb[k] = (pit - sum_{l < k} conv[l]) - pit sigb[k],   sd[0] = 0

  const double rdxj_jp = g.rdx_j[jp];
  double pre_c = 0.0, pre_ip = 0.0, pre_jp = 0.0;
  // layer L-1 pairs with layer 0 through the bottom of layer 0 (np.roll), times sd[0] = 0
  const double u_top = fld[0][(L - 1) * plane + e_c], v_top = fld[1][(L - 1) * plane + e_c],
               t_top = fld[2][(L - 1) * plane + e_c], q_top = fld[3][(L - 1) * plane + e_c];
  (void)im;

  gcm_cp_async_wait<PFT_NS - 3>();  // layers 0 and 1 have landed
  __syncthreads();
  if (PFT_NS - 1 < L) issue(PFT_NS - 1, PFT_NS - 1);
  const double* s0 = sm;
  double u_k = s0[0 * TILE + t_c], v_k = s0[1 * TILE + t_c], t_k = s0[2 * TILE + t_c],
         q_k = s0[3 * TILE + t_c];
  double sd_c = 0.0, sd_ip = 0.0, sd_jp = 0.0;  // level 0
  double fu = (u_k + u_top) * 0.5 * ((sd_c + sd_ip) * 0.5);
  double fv_ = (v_k + v_top) * 0.5 * ((sd_c + sd_jp) * 0.5);
  double ft = (t_k + t_top) * 0.5 * sd_c;
  double fq = (q_k + q_top) * 0.5 * sd_c;
  const double fu0 = fu, fv0 = fv_, ft0 = ft, fq0 = fq;
  bool bad = active && gcm_not_finite(pn_c);

#pragma unroll
  for (int k = 0; k < L; ++k) {
    if (k > 0) {  // layers k and k + 1 have landed; the stage of layer k - 1 is free for layer k + NS - 1
      gcm_cp_async_wait<PFT_NS - 3>();
      __syncthreads();
      if (k + PFT_NS - 1 < L) issue(k + PFT_NS - 1, (k + PFT_NS - 1) % PFT_NS);
      else gcm_cp_async_commit();  // keep one group per iteration so that the wait count stays exact
    }
    const double* sk = sm + (k % PFT_NS) * PFT_STAGE;
    const double* sn = sm + ((k + 1) % PFT_NS) * PFT_STAGE;
    // this cell's pgf, fv, u, v, t, q: read once, straight from global memory, consumed at the end of the layer;
    // their lines are asked into L1 two layers ahead (no register, no scoreboard)
    const int e = k * plane + e_c;
    if (k + pfd < L) {
#pragma unroll
      for (int f = 6; f < 12; ++f) gcm_prefetch_l1(fld[f] + e + pfd * plane);
    }
    const double own_pgf = fld[6][e], own_fv = fld[7][e], own_u = fld[8][e], own_v = fld[9][e], own_t = fld[10][e],
                 own_q = fld[11][e];
    // horizontal neighbours from the tile
    const double* su_ = sk + 0 * TILE + t_c;
    const double* sv_ = sk + 1 * TILE + t_c;
    const double* st_ = sk + 2 * TILE + t_c;
    const double* sq_ = sk + 3 * TILE + t_c;
    const double* pu_ = sk + 4 * TILE + t_c;
    const double u_im = su_[-1], u_ip = su_[1], u_jp = su_[TROW], u_jm = su_[-TROW];
    const double v_im = sv_[-1], v_ip = sv_[1], v_jp = sv_[TROW], v_jm = sv_[-TROW], v_jm_ip = sv_[1 - TROW];
    const double pu_c = pu_[0], pu_im = pu_[-1], pu_ip = pu_[1], pu_jp = pu_[TROW], pu_jp_im = pu_[TROW - 1];
    const double pv_c = v_k * a_c, pv_ip = v_ip * a_ip, pv_jm = v_jm * a_jm, pv_jm_ip = v_jm_ip * a_jm_ip,
                 pv_jp = v_jp * a_jp;

    // fluxes through the top of layer k (advec_sig, dynamics.py:49-52) with sd at level k + 1
    double fu_n = fu0, fv_n = fv0, ft_n = ft0, fq_n = fq0;
    double u_kp = 0.0, v_kp = 0.0, t_kp = 0.0, q_kp = 0.0;
    if (k + 1 < L) {
      const double ds = g.c_dsig[k], sb = g.c_sigb[k + 1];
      pre_c += ((pu_c - pu_im) * rdxj + (pv_c - pv_jm) * rdy) * ds;           // conv of (j, i)      dynamics.py:39
      pre_ip += ((pu_ip - pu_c) * rdxj + (pv_ip - pv_jm_ip) * rdy) * ds;      //         (j, i+1)
      pre_jp += ((pu_jp - pu_jp_im) * rdxj_jp + (pv_jp - pv_c) * rdy) * ds;   //         (j+1, i)
      sd_c = (pit_c - pre_c) - pit_c * sb;
      sd_ip = (pit_ip - pre_ip) - pit_ip * sb;
      sd_jp = (pit_jp - pre_jp) - pit_jp * sb;
      u_kp = sn[0 * TILE + t_c]; v_kp = sn[1 * TILE + t_c]; t_kp = sn[2 * TILE + t_c];
      q_kp = sn[3 * TILE + t_c];
      fu_n = (u_kp + u_k) * 0.5 * ((sd_c + sd_ip) * 0.5);
      fv_n = (v_kp + v_k) * 0.5 * ((sd_c + sd_jp) * 0.5);
      ft_n = (t_kp + t_k) * 0.5 * sd_c;
      fq_n = (q_kp + q_k) * 0.5 * sd_c;
    }
    const double rds = g.c_rdsig[k];
    const double dus = (fu_n - fu) * rds, dvs = (fv_n - fv_) * rds;  // -(F_k - F_k+1) / dsig
    const double ads_t = (ft_n - ft) * rds, ads_q = (fq_n - fq) * rds;

    // advec_m_pu (dynamics.py:55-108); (a/2)(b/2) = ab/4 exactly
    const double puum = (u_k + u_im) * (pu_c + pu_im), puup = (u_ip + u_k) * (pu_ip + pu_c);
    const double puvp = (pv_c + pv_ip) * (u_k + u_jp), puvm = (pv_jm + pv_jm_ip) * (u_jm + u_k);
    const double dut = ((puum - puup) * rdxj + (puvm - puvp) * rdy) * 0.25;
    const double pvvm = (v_k + v_jm) * (pv_c + pv_jm), pvvp = (v_jp + v_k) * (pv_jp + pv_c);
    const double pvup = (v_k + v_ip) * (pu_c + pu_jp), pvum = (v_im + v_k) * (pu_im + pu_jp_im);
    const double dvt = ((pvvm - pvvp) * rdy + (pvum - pvup) * rdxh) * 0.25;

    const double pu_n = own_u * pu_fac - (dut + dus + own_pgf) * dt;  // dynamics.py:206
    const double pv_n = own_v * pv_fac - (dvt + dvs + own_fv) * dt;   // dynamics.py:207
    double v_n = pv_n * r_pnv;
    if (zero_v) v_n *= 0.0;  // dynamics.py:222
    // tracers: advec_t (dynamics.py:174-181) + advec_sig, flux form (dynamics.py:214, :219)
    const double adv_t = ((pu_c * (t_k + st_[1]) - pu_im * (st_[-1] + t_k)) * rdxj +
                          (pv_c * (t_k + st_[TROW]) - pv_jm * (st_[-TROW] + t_k)) * rdy) * 0.5;
    const double adv_q = ((pu_c * (q_k + sq_[1]) - pu_im * (sq_[-1] + q_k)) * rdxj +
                          (pv_c * (q_k + sq_[TROW]) - pv_jm * (sq_[-TROW] + q_k)) * rdy) * 0.5;
    if (active) {
      const double u_n = pu_n * r_pnu;
      const double t_n = (own_t * p_c - (adv_t + ads_t) * dt) * r_pn;
      const double q_n = (own_q * p_c - (adv_q + ads_q) * dt) * r_pn;
      ou[e] = u_n;
      ov[e] = v_n;
      ot[e] = t_n;
      oq[e] = q_n;
      bad |= gcm_not_finite((u_n + v_n) + (t_n + q_n));
    }
    fu = fu_n; fv_ = fv_n; ft = ft_n; fq = fq_n;
    u_k = u_kp; v_k = v_kp; t_k = t_kp; q_k = q_kp;
  }
  if (active) out.p[o2 + e_c] = pn_c;
  gcm_flag_nonfinite(g.nonfinite, bad);
}


// ---------------------------------------------------------------------------------------------------
// U, TMA: the tiled update with every operand delivered by the Tensor Memory Accelerator.  Per layer ONE thread issues
// bulk tensor loads (cp.async.bulk.tensor.3d, SASS UTMALDG): the (TJ + 2) x 36 halo box of su, sv, st, sq, spu and the
// TJ x 32 box of the six once-read fields (pgf, fv, and in the corrector the base u, v, t, q; in the predictor
// base == star and the cell's own staged value is used), three layers in flight, each stage completing on its own
// mbarrier.  No thread computes a load address in the layer loop: the LDGSTS kernel above spends a quarter of its
// instructions on them and keeps the LSU pipe 65 % busy (ncu r03j).  A box that leaves the grid is zero-filled by the
// hardware; the periodic wrap in i (every band and grid) and in j (whole grids) is patched by the CTAs on the seam --
// 2 of W / 32 tile columns, 2 of H / TJ tile rows -- with plain loads after the box has landed.
// ---------------------------------------------------------------------------------------------------
struct PfSeam {
  const double *f0, *f1, *f2, *f3, *f4;
  int j0, x0, H, W, wrap, edge_w, edge_e, edge_j, tid;
};
// halo elements of a landed stage that the box left zero-filled <- their periodic images (seam CTAs only)
template <int PFT_TJ>
__device__ __noinline__ void pf_seam_patch(const PfSeam sd, double* st, int HT, int koff) {
  const double* fld[PFT_NF] = {sd.f0, sd.f1, sd.f2, sd.f3, sd.f4};
  auto patch = [&](int dr, int c) {  // tile element (dr, c)
    int gj = sd.j0 - 1 + dr, gi = sd.x0 + c;
    if (gj < 0 || gj >= sd.H) {
      if (!sd.wrap) return;  // a band stores its halo rows; rows outside it are never used by an active thread
      gj = ((gj % sd.H) + sd.H) % sd.H;
    }
    gi = gi < 0 ? gi + sd.W : (gi >= sd.W ? gi - sd.W : gi);
    const int src = koff + gj * sd.W + gi, rr = dr * PFT_ROW + c;
#pragma unroll
    for (int f = 0; f < PFT_NF; ++f) st[f * HT + rr] = fld[f][src];
  };
  if (sd.edge_w || sd.edge_e) {  // the halo column beyond the seam: tile column 1 (i = -1) / 34 (i = W)
    if (sd.tid < 2 * (PFT_TJ + 2)) {
      const int side = sd.tid / (PFT_TJ + 2), dr = sd.tid - side * (PFT_TJ + 2);
      if (side == 0 ? sd.edge_w : sd.edge_e) patch(dr, side == 0 ? 1 : PFT_TI + 2);
    }
  }
  if (sd.edge_j) {  // whole rows beyond the first / last row of a periodic grid
    for (int dr = 0; dr < PFT_TJ + 2; ++dr) {
      const int gj = sd.j0 - 1 + dr;
      if (gj >= 0 && gj < sd.H) continue;
      if (sd.tid >= 1 && sd.tid <= PFT_TI + 2) patch(dr, sd.tid);
    }
  }
}

struct PfTmaMaps {
  GcmTmap halo[PFT_NF];  // su, sv, st, sq, spu: box 36 x (TJ + 2) x 1
  GcmTmap cen[6];        // pgf, fv, u, v, t, q: box 32 x TJ x 1
};

// Warp-specialised: warp PFT_TJ of the CTA is the PRODUCER (one lane issues the box loads of layer k into stage k % NS as
// soon as the consumers have released it), warps 0 .. PFT_TJ-1 are the CONSUMERS (one tile row each).  Stages are handed
// over through mbarriers only -- full[s]: the bytes of a layer have landed; empty[s]: every consumer warp is done with
// it -- so the layer loop has no block-wide barrier and no thread of a compute warp ever issues a load.
template <int L, int PFT_TJ, bool SAME, int NS, int MINB>
__global__ void __launch_bounds__(PFT_TI * (PFT_TJ + 1), MINB)
pe25f_update_tma_kernel(GcmGeomDev g, const GCM_GRID_CONSTANT PfTmaMaps maps, PfConst base, PfConst star, PfMut out,
                        PfWork w, double dt, GcmRowSeg seg, size_t bstride2, size_t bstride3) {
  if (g.pdl_early) gcm_pdl_trigger();
  gcm_pdl_wait();
  GCM_DYN_SMEM(unsigned char, smraw);
  GcmMbar* full = reinterpret_cast<GcmMbar*>(smraw);  // barriers in the first 128 bytes: full[NS], empty[NS], patch
  GcmMbar* empty = full + NS;
  GcmMbar* patchbar = empty + NS;
  double* sm = reinterpret_cast<double*>(smraw + 128);
  constexpr int HBOX = (PFT_TJ + 2) * PFT_ROW;        // doubles a halo box delivers
  constexpr int HT = (HBOX + 15) / 16 * 16;           // halo tile pitch: box destinations are 128-byte aligned
  constexpr int CT = PFT_TJ * PFT_TI;                 // centre tile
  constexpr int NCEN = SAME ? 2 : 6;
  constexpr int STAGE = PFT_NF * HT + NCEN * CT;
  constexpr unsigned STAGE_BYTES = (PFT_NF * HBOX + NCEN * CT) * sizeof(double);
  constexpr int NCONS = PFT_TI * PFT_TJ;              // consumer threads
  const int H = g.H, W = g.W, plane = H * W, wrap = g.wrap_j;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * PFT_TI + tx;
  const int j0 = seg.a + blockIdx.y * PFT_TJ;  // first row of the tile
  const int x0 = blockIdx.x * PFT_TI - 2;      // first column of the halo box: the interior starts 16-byte aligned
  const int zb = blockIdx.z * L;  // first layer of this member in the [members * L][H][W] view of the tensor maps

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      gcm_mbar_init(&full[s], 1);
      gcm_mbar_init(&empty[s], PFT_TJ);
    }
    gcm_mbar_init(patchbar, NCONS);
    gcm_mbar_fence_init();
  }
  __syncthreads();
  if (ty == PFT_TJ) {  // ---- producer warp ----
    if (tx == 0) {
#pragma unroll 1
      for (int k = 0; k < L; ++k) {
        const int s = k % NS;
        if (k >= NS) gcm_mbar_wait_backoff(&empty[s], ((k / NS) - 1) & 1);  // the consumers are done with layer k - NS
        gcm_fence_proxy_async();
        double* st = sm + s * STAGE;
        gcm_mbar_expect_tx(&full[s], STAGE_BYTES);
#pragma unroll
        for (int f = 0; f < PFT_NF; ++f) gcm_tma_load3(st + f * HT, &maps.halo[f], x0, j0 - 1, zb + k, &full[s]);
#pragma unroll
        for (int f = 0; f < NCEN; ++f)
          gcm_tma_load3(st + PFT_NF * HT + f * CT, &maps.cen[f], blockIdx.x * PFT_TI, j0, zb + k, &full[s]);
      }
    }
    return;
  }
  // ---- consumer warps ----
  const int i = blockIdx.x * PFT_TI + tx;
  const int r = blockIdx.y * PFT_TJ + ty;
  const bool active = r < seg.n1;
  auto rowc = [&](int x) { return wrap ? ((x % H) + H) % H : (x < 0 ? 0 : (x >= H ? H - 1 : x)); };
  const int j = rowc(j0 + ty);
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* __restrict__ p = base.p + o2;
  const double* __restrict__ sp = star.p + o2;
  const double* __restrict__ pn = w.pn + o2;
  const double* __restrict__ pit = w.pit + o2;
  const double* fld[PFT_NF] = {star.u + o3, star.v + o3, star.t + o3, star.q + o3, w.spu + o3};
  double* __restrict__ ou = out.u + o3;
  double* __restrict__ ov = out.v + o3;
  double* __restrict__ ot = out.t + o3;
  double* __restrict__ oq = out.q + o3;

  // CTAs whose halo box leaves the grid patch the zero-filled part with the periodic neighbour
  const bool edge_w = blockIdx.x == 0, edge_e = blockIdx.x + 1 == gridDim.x;
  const bool edge_j = wrap && (j0 - 1 < 0 || j0 + PFT_TJ + 1 > H);
  const bool edge = edge_w || edge_e || edge_j;
  // patch stage s of layer k (its bytes have landed), then meet the other consumers: the patched elements are halo
  // elements, which only the layer's own iteration reads.  Out of line: only the seam CTAs run it, and inlined nine
  // times it tripled the kernel image (instruction-cache misses were the third stall reason, ncu r2d).
  auto fixup = [&](int k, int s) {
    PfSeam sd_;
    sd_.f0 = fld[0]; sd_.f1 = fld[1]; sd_.f2 = fld[2]; sd_.f3 = fld[3]; sd_.f4 = fld[4];
    sd_.j0 = j0; sd_.x0 = x0; sd_.H = H; sd_.W = W; sd_.wrap = wrap; sd_.edge_w = edge_w; sd_.edge_e = edge_e;
    sd_.edge_j = edge_j; sd_.tid = tid;
    pf_seam_patch<PFT_TJ>(sd_, sm + s * STAGE, HT, k * plane);
    gcm_mbar_arrive(patchbar);
    gcm_mbar_wait(patchbar, k & 1);
  };

  const int jm = rowc(j0 + ty - 1), jp = rowc(j0 + ty + 1), jpp = rowc(j0 + ty + 2);
  const int ip = gcm_ip(i, W);
  const int e_c = j * W + i;
  const int t_c = (ty + 1) * PFT_ROW + (tx + 2);  // own position in a halo tile
  const int c_c = ty * PFT_TI + tx;               // own position in a centre tile

  // per-column (2-D) factors while the first layers are on their way
  const double rdxj = g.rdx_j[j], rdxh = g.rdx_h[j], rdy = g.rdy;
  const double p_c = p[e_c], p_ip = p[j * W + ip], p_jp = p[jp * W + i];
  const double pn_c = pn[e_c], pn_ip = pn[j * W + ip], pn_jp = pn[jp * W + i];
  const double pu_fac = (p_c + p_ip) * 0.5, pv_fac = (p_c + p_jp) * 0.5;                    // calc_pu / calc_pv
  const double r_pnu = 1.0 / ((pn_c + pn_ip) * 0.5), r_pnv = 1.0 / ((pn_c + pn_jp) * 0.5);  // un_pu / un_pv
  const double r_pn = 1.0 / pn_c;
  const double sp_c = sp[e_c], sp_ip = sp[j * W + ip], sp_jp = sp[jp * W + i], sp_jm = sp[jm * W + i];
  const double a_c = (sp_c + sp_jp) * 0.5;                     // jph(sp) at (j, i)
  const double a_ip = (sp_ip + sp[jp * W + ip]) * 0.5;         // (j, i+1)
  const double a_jm = (sp_jm + sp_c) * 0.5;                    // (j-1, i)
  const double a_jm_ip = (sp[jm * W + ip] + sp_ip) * 0.5;      // (j-1, i+1)
  const double a_jp = (sp_jp + sp[jpp * W + i]) * 0.5;         // (j+1, i)
  const bool zero_v = j == g.zero_v_row || j == g.zero_v_row2;
  // sd of the three columns the vertical fluxes need is rebuilt from pit and the running sums of conv (see the
  // LDGSTS kernel above)
  const double pit_c = pit[e_c], pit_ip = pit[j * W + ip], pit_jp = pit[jp * W + i];
  const double rdxj_jp = g.rdx_j[jp];
  double pre_c = 0.0, pre_ip = 0.0, pre_jp = 0.0;
  // layer L-1 pairs with layer 0 through the bottom of layer 0 (np.roll), times sd[0] = 0
  const double u_top = fld[0][(L - 1) * plane + e_c], v_top = fld[1][(L - 1) * plane + e_c],
               t_top = fld[2][(L - 1) * plane + e_c], q_top = fld[3][(L - 1) * plane + e_c];

  gcm_mbar_wait(&full[0], 0);
  double u_k = 0.0, v_k = 0.0, t_k = 0.0, q_k = 0.0;
  double fu = 0.0, fv_ = 0.0, ft = 0.0, fq = 0.0, fu0 = 0.0, fv0 = 0.0, ft0 = 0.0, fq0 = 0.0;
  double sd_c = 0.0, sd_ip = 0.0, sd_jp = 0.0;  // level 0
  bool bad = active && gcm_not_finite(pn_c);

#pragma unroll
  for (int k = 0; k < L; ++k) {
    if (edge) fixup(k, k % NS);
    if (k + 1 < L)  // the next layer's centre values feed the fluxes through the top of layer k
      gcm_mbar_wait(&full[(k + 1) % NS], ((k + 1) / NS) & 1);
    const double* sk = sm + (k % NS) * STAGE;
    const double* sn = sm + ((k + 1) % NS) * STAGE;
    if (k == 0) {
      u_k = sk[0 * HT + t_c]; v_k = sk[1 * HT + t_c]; t_k = sk[2 * HT + t_c]; q_k = sk[3 * HT + t_c];
      fu = (u_k + u_top) * 0.5 * ((sd_c + sd_ip) * 0.5);
      fv_ = (v_k + v_top) * 0.5 * ((sd_c + sd_jp) * 0.5);
      ft = (t_k + t_top) * 0.5 * sd_c;
      fq = (q_k + q_top) * 0.5 * sd_c;
      fu0 = fu; fv0 = fv_; ft0 = ft; fq0 = fq;
    }
    const int e = k * plane + e_c;
    const double* ck = sk + PFT_NF * HT + c_c;
    const double own_pgf = ck[0 * CT], own_fv = ck[1 * CT];
    const double own_u = SAME ? u_k : ck[2 * CT], own_v = SAME ? v_k : ck[3 * CT];
    const double own_t = SAME ? t_k : ck[4 * CT], own_q = SAME ? q_k : ck[5 * CT];
    // horizontal neighbours from the tile
    const double* su_ = sk + 0 * HT + t_c;
    const double* sv_ = sk + 1 * HT + t_c;
    const double* st_ = sk + 2 * HT + t_c;
    const double* sq_ = sk + 3 * HT + t_c;
    const double* pu_ = sk + 4 * HT + t_c;
    const double u_im = su_[-1], u_ip = su_[1], u_jp = su_[PFT_ROW], u_jm = su_[-PFT_ROW];
    const double v_im = sv_[-1], v_ip = sv_[1], v_jp = sv_[PFT_ROW], v_jm = sv_[-PFT_ROW], v_jm_ip = sv_[1 - PFT_ROW];
    const double pu_c = pu_[0], pu_im = pu_[-1], pu_ip = pu_[1], pu_jp = pu_[PFT_ROW], pu_jp_im = pu_[PFT_ROW - 1];
    const double pv_c = v_k * a_c, pv_ip = v_ip * a_ip, pv_jm = v_jm * a_jm, pv_jm_ip = v_jm_ip * a_jm_ip,
                 pv_jp = v_jp * a_jp;

    // fluxes through the top of layer k (advec_sig, dynamics.py:49-52) with sd at level k + 1
    double fu_n = fu0, fv_n = fv0, ft_n = ft0, fq_n = fq0;
    double u_kp = 0.0, v_kp = 0.0, t_kp = 0.0, q_kp = 0.0;
    if (k + 1 < L) {
      const double ds = g.c_dsig[k], sb = g.c_sigb[k + 1];
      pre_c += ((pu_c - pu_im) * rdxj + (pv_c - pv_jm) * rdy) * ds;           // conv of (j, i)      dynamics.py:39
      pre_ip += ((pu_ip - pu_c) * rdxj + (pv_ip - pv_jm_ip) * rdy) * ds;      //         (j, i+1)
      pre_jp += ((pu_jp - pu_jp_im) * rdxj_jp + (pv_jp - pv_c) * rdy) * ds;   //         (j+1, i)
      sd_c = (pit_c - pre_c) - pit_c * sb;
      sd_ip = (pit_ip - pre_ip) - pit_ip * sb;
      sd_jp = (pit_jp - pre_jp) - pit_jp * sb;
      u_kp = sn[0 * HT + t_c]; v_kp = sn[1 * HT + t_c]; t_kp = sn[2 * HT + t_c]; q_kp = sn[3 * HT + t_c];
      fu_n = (u_kp + u_k) * 0.5 * ((sd_c + sd_ip) * 0.5);
      fv_n = (v_kp + v_k) * 0.5 * ((sd_c + sd_jp) * 0.5);
      ft_n = (t_kp + t_k) * 0.5 * sd_c;
      fq_n = (q_kp + q_k) * 0.5 * sd_c;
    }
    const double rds = g.c_rdsig[k];
    const double dus = (fu_n - fu) * rds, dvs = (fv_n - fv_) * rds;  // -(F_k - F_k+1) / dsig
    const double ads_t = (ft_n - ft) * rds, ads_q = (fq_n - fq) * rds;

    // advec_m_pu (dynamics.py:55-108); (a/2)(b/2) = ab/4 exactly
    const double puum = (u_k + u_im) * (pu_c + pu_im), puup = (u_ip + u_k) * (pu_ip + pu_c);
    const double puvp = (pv_c + pv_ip) * (u_k + u_jp), puvm = (pv_jm + pv_jm_ip) * (u_jm + u_k);
    const double dut = ((puum - puup) * rdxj + (puvm - puvp) * rdy) * 0.25;
    const double pvvm = (v_k + v_jm) * (pv_c + pv_jm), pvvp = (v_jp + v_k) * (pv_jp + pv_c);
    const double pvup = (v_k + v_ip) * (pu_c + pu_jp), pvum = (v_im + v_k) * (pu_im + pu_jp_im);
    const double dvt = ((pvvm - pvvp) * rdy + (pvum - pvup) * rdxh) * 0.25;

    const double pu_n = own_u * pu_fac - (dut + dus + own_pgf) * dt;  // dynamics.py:206
    const double pv_n = own_v * pv_fac - (dvt + dvs + own_fv) * dt;   // dynamics.py:207
    double v_n = pv_n * r_pnv;
    if (zero_v) v_n *= 0.0;  // dynamics.py:222
    // tracers: advec_t (dynamics.py:174-181) + advec_sig, flux form (dynamics.py:214, :219)
    const double adv_t = ((pu_c * (t_k + st_[1]) - pu_im * (st_[-1] + t_k)) * rdxj +
                          (pv_c * (t_k + st_[PFT_ROW]) - pv_jm * (st_[-PFT_ROW] + t_k)) * rdy) * 0.5;
    const double adv_q = ((pu_c * (q_k + sq_[1]) - pu_im * (sq_[-1] + q_k)) * rdxj +
                          (pv_c * (q_k + sq_[PFT_ROW]) - pv_jm * (sq_[-PFT_ROW] + q_k)) * rdy) * 0.5;
    if (active) {
      const double u_n = pu_n * r_pnu;
      const double t_n = (own_t * p_c - (adv_t + ads_t) * dt) * r_pn;
      const double q_n = (own_q * p_c - (adv_q + ads_q) * dt) * r_pn;
      ou[e] = u_n;
      ov[e] = v_n;
      ot[e] = t_n;
      oq[e] = q_n;
      bad |= gcm_not_finite((u_n + v_n) + (t_n + q_n));
    }
    fu = fu_n; fv_ = fv_n; ft = ft_n; fq = fq_n;
    u_k = u_kp; v_k = v_kp; t_k = t_kp; q_k = q_kp;
    if (k + NS < L) {  // this warp is done with the stage of layer k: hand it back to the producer
      __syncwarp();
      if (tx == 0) gcm_mbar_arrive(&empty[k % NS]);
    }
  }
  if (active) out.p[o2 + e_c] = pn_c;
  gcm_flag_nonfinite(g.nonfinite, bad);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// One half step on the rows of `segR` (row phase: spu, pit, p_n, pgf, fv) and `segU` (update).  A whole grid or band
// is one segment each (segR = owned rows + the first halo row to the south in band mode); gcm_pe25_half_step_rows
// passes two-segment launches for the rows next to the halos.
template <int L>
static int pf_half_step(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out, double dt,
                        int nbatch, const PfWork& w, GcmRowSeg segR, GcmRowSeg segU, void* stream) {
  GcmGeomDev d = g->d;
  d.pdl_early = g_gcm_knob[9] == 2;
  const int H = d.H, W = d.W;
  const size_t b2 = (size_t)H * W, b3 = (size_t)L * H * W;
  const int nrowsR = segR.n1 + segR.n2, nrowsU = segU.n1 + segU.n2;
  const size_t prsmem = (size_t)W * sizeof(double2);  // one packed row (two layers of one latitude)
  constexpr int NP = (L + 1) / 2;
  const PfConst cb{base->p, base->u, base->v, base->t, base->q};
  const PfConst cs{star->p, star->u, star->v, star->t, star->q};
  const PfMut mo{out->p, out->u, out->v, out->t, out->q};
  const bool ptop0 = d.ptop == 0.0;
  const unsigned magicW = gcm_magic((unsigned)W);
  // Update kernel.  Narrow single grids (W x members <= 128: the 72 x 46 and 36 x 24 grids, not their ensembles)
  // take one thread per cell: too few columns to fill the chip with a thread per column or a CTA per tile (r03g:
  // 72 x 46 0.055 -> 0.029 ms/step).  Already at 288 x 180 the tiled kernel wins again (r03i: 0.071 vs 0.076).
  // The choice depends on the width and the member count only, never on the rows of the launch, so a latitude band
  // takes the same kernel as the whole grid (bit-identical decomposition).
  const bool cells = ((size_t)W * nbatch <= 128 || g_gcm_knob[4] == 3) && (size_t)(segU.n1 + segU.n2) * W < (1u << 22) &&
                     g_gcm_knob[4] != 2;
  // update on staged shared-memory tiles: 32-wide tiles, or 36-wide ones when only those divide the row (the 36 x 24
  // ensemble members: knob 4 = 6 or 2 keeps them on the direct-load kernel)
  const bool tile36 = W % PFT_TI != 0 && W % 36 == 0 && g_gcm_knob[4] != 6 && g_gcm_knob[4] != 2;
  const bool tiled = !cells && (W % PFT_TI == 0 || tile36) && g_gcm_knob[4] != 1;
  // aflux fused into the filter of the mass flux (filter MODE 2 writes the per-pair partial sums of conv, the tiled
  // update forms pit and p_n from them): one launch and one pass over spu less per half step.  Needs the tiled update
  // (the other update kernels read sd), a compile-time FFT plan, no opt-in terms (pe25_extras reads p_n).  Knob 16
  // selects it (GCM_FUSE_AFLUX_ON below)
  const bool fuse_aflux = tiled && g_gcm_knob[4] != 5 && g_gcm_knob[16] == GCM_FUSE_AFLUX_ON && !gcm_extras_on(g) &&
                          g_gcm_knob[8] != 1 && gcm_fixed_plan_id(d.plan) > 0;
  if (nrowsR > 0) {
    // Two independent chains:  F(su iph(sp)) -> aflux   and   hydro -> F(pgfu + phiu).  On a whole grid / band they run
    // side by side (caller's stream + the geometry's side stream).
    cudaStream_t qa = (cudaStream_t)stream, qb = (cudaStream_t)stream;
#ifndef GCM_EMU
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    const bool side = segR.n2 == 0 && g_gcm_knob[3] != 1;
    if (side) {
      void *q2, *e1, *e2;
      int st2 = gcm_geom_aux(g, &q2, &e1, &e2);
      if (st2) return st2;
      qb = (cudaStream_t)q2;
      ev_fork = (cudaEvent_t)e1;
      ev_join = (cudaEvent_t)e2;
      GCM_CUDA(cudaEventRecord(ev_fork, qa));
      GCM_CUDA(cudaStreamWaitEvent(qb, ev_fork, 0));
    }
#endif
    // filter launches: NBAT packed rows per CTA, about 1440 elements, but at least four CTAs per SM when possible
    const int npr_total = nrowsR * NP;
    int nbf = 1440 / W < 1 ? 1 : 1440 / W;
    while (nbf > 1 && (size_t)((npr_total + nbf - 1) / nbf) * nbatch < 592) --nbf;
    if (g_gcm_knob[1] > 0) nbf = g_gcm_knob[1];
    int tf = (nbf * W / 12 + 31) / 32 * 32;
    tf = tf < 32 ? 32 : (tf > 256 ? 256 : tf);
    if (g_gcm_knob[0] > 0) tf = g_gcm_knob[0] > 256 ? 256 : g_gcm_knob[0];
    const size_t smf = nbf * prsmem;
    const dim3 gridf((npr_total + nbf - 1) / nbf, nbatch);
    const int plan_id = g_gcm_knob[8] == 1 ? 0 : gcm_fixed_plan_id(d.plan);  // compile-time radices when known
    // hydro launch: warp tasks of RG rows x 31 columns; shrink RG until there are about 16 warps per SM
    const int nchunk = (W + 30) / 31;
    int rg = 8;
    while (rg > 1 && (size_t)nchunk * ((nrowsR + rg - 1) / rg) * nbatch < 2368) rg /= 2;
    if (g_gcm_knob[2] > 0) rg = g_gcm_knob[2];
    if (segR.n2 > 0) rg = 1;  // a group of rows must be contiguous
    const int ntasks = nchunk * ((nrowsR + rg - 1) / rg);
    {
      GcmProfScope ps(GCM_K_FILTER_A, qa);
      int stf = fuse_aflux ? pf_filter_launch<L, 2>(plan_id, d, gridf, tf, smf, qa, star->p, star->u, w.spu, segR, nbf, b2,
                                                    b3, star->v, w.sd)
                           : pf_filter_launch<L, 1>(plan_id, d, gridf, tf, smf, qa, star->p, star->u, w.spu, segR, nbf, b2,
                                                    b3);
      if (stf) return stf;
    }
    if (W < 62 && g_gcm_knob[7] != 1) {  // narrow rows: whole row groups per CTA, east neighbour through shared memory
      GcmProfScope ps(GCM_K_COLUMN_F, qb);
      const int ngrp = (nrowsR + rg - 1) / rg;
      int G = 128 / W < 1 ? 1 : 128 / W;
      if (G > ngrp) G = ngrp;
      const int th = (G * W + 31) / 32 * 32;
      const size_t smh = (size_t)2 * (2 * L + 1) * th * sizeof(double);
      const dim3 gridn((ngrp + G - 1) / G, nbatch);
#ifndef GCM_EMU
      if (smh > 48 * 1024) {
        GCM_CUDA(cudaFuncSetAttribute(pe25f_hydro_narrow_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smh));
        GCM_CUDA(cudaFuncSetAttribute(pe25f_hydro_narrow_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smh));
      }
#endif
      if (ptop0)
        GCM_LAUNCH_DEP((pe25f_hydro_narrow_kernel<L, true>), gridn, dim3(th), smh, qb, d, cs, w, segR, rg, G, b2, b3);
      else
        GCM_LAUNCH_DEP((pe25f_hydro_narrow_kernel<L, false>), gridn, dim3(th), smh, qb, d, cs, w, segR, rg, G, b2, b3);
    } else {
      GcmProfScope ps(GCM_K_COLUMN_F, qb);
      const dim3 gridc((ntasks + 3) / 4, nbatch);
      const int mbh = (g_gcm_knob[15] / 10) % 10;  // register-budget variants (knob 15, tens digit)
      // tile form: the default from 256 columns up (a 72-wide grid is 18 CTAs of it and loses to the marching kernel,
      // r2u); the choice depends on the width only, never on the rows of a launch.  One launch per contiguous segment.
      if (g_gcm_knob[7] == 0 && W >= 256) {
        const GcmRowSeg hp[2] = {{segR.a, segR.n1, 0, 0}, {segR.c, segR.n2, 0, 0}};
        const size_t smh = (size_t)(PFH_RT + 1) * (2 * L + 1) * 32 * sizeof(double);
        for (int s2 = 0; s2 < 2; ++s2) {
          if (hp[s2].n1 <= 0) continue;
          const dim3 gridh(nchunk, (hp[s2].n1 + PFH_RT - 1) / PFH_RT, nbatch), blockh(32, PFH_RT + 1);
#ifndef GCM_EMU
          if (smh > 48 * 1024) {
            GCM_CUDA(cudaFuncSetAttribute(pe25f_hydro_tile_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smh));
            GCM_CUDA(cudaFuncSetAttribute(pe25f_hydro_tile_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smh));
          }
#endif
          if (ptop0 && mbh == 3 && smh <= 48 * 1024)  // 3 CTAs per SM at 72 registers
            GCM_LAUNCH_DEP((pe25f_hydro_tile_kernel<L, true, 3>), gridh, blockh, smh, qb, d, cs, w, hp[s2], b2, b3);
          else if (ptop0) GCM_LAUNCH_DEP((pe25f_hydro_tile_kernel<L, true>), gridh, blockh, smh, qb, d, cs, w, hp[s2], b2, b3);
          else GCM_LAUNCH_DEP((pe25f_hydro_tile_kernel<L, false>), gridh, blockh, smh, qb, d, cs, w, hp[s2], b2, b3);
          GCM_CHECK_LAUNCH();
        }
      } else
      if (L == 9 && ptop0 && (mbh == 5 || mbh == 6 || mbh == 8)) {
        if constexpr (L == 9) {
          if (mbh == 5) GCM_LAUNCH_DEP((pe25f_hydro_kernel<L, true, 5>), gridc, dim3(128), 0, qb, d, cs, w, segR, rg, b2, b3);
          else if (mbh == 6) GCM_LAUNCH_DEP((pe25f_hydro_kernel<L, true, 6>), gridc, dim3(128), 0, qb, d, cs, w, segR, rg, b2, b3);
          else GCM_LAUNCH_DEP((pe25f_hydro_kernel<L, true, 8>), gridc, dim3(128), 0, qb, d, cs, w, segR, rg, b2, b3);
        }
      } else if (ptop0)
        GCM_LAUNCH_DEP((pe25f_hydro_kernel<L, true>), gridc, dim3(128), 0, qb, d, cs, w, segR, rg, b2, b3);
      else
        GCM_LAUNCH_DEP((pe25f_hydro_kernel<L, false>), gridc, dim3(128), 0, qb, d, cs, w, segR, rg, b2, b3);
    }
    GCM_CHECK_LAUNCH();
    if (!fuse_aflux) {
      GcmProfScope ps(GCM_K_AFLUX_F, qa);
      const dim3 grida((nrowsR * W + 127) / 128, nbatch);
      if (tiled)  // the tiled update rebuilds sd: pit and p_n only
        GCM_LAUNCH_DEP((pe25f_aflux_kernel<L, false>), grida, dim3(128), 0, qa, d, base->p, star->p, star->v, w, dt, segR,
                   magicW, b2, b3);
      else
        GCM_LAUNCH_DEP((pe25f_aflux_kernel<L, true>), grida, dim3(128), 0, qa, d, base->p, star->p, star->v, w, dt, segR,
                   magicW, b2, b3);
    }
    GCM_CHECK_LAUNCH();
    {
      GcmProfScope ps(GCM_K_FILTER_B, qb);
      int stf = pf_filter_launch<L, 0>(plan_id, d, gridf, tf, smf, qb, star->p, w.pgf, w.pgf, segR, nbf, b2, b3);
      if (stf) return stf;
    }
#ifndef GCM_EMU
    if (side) {
      GCM_CUDA(cudaEventRecord(ev_join, qb));
      GCM_CUDA(cudaStreamWaitEvent(qa, ev_join, 0));
    }
#endif
  }
  // TMA update (knob 4 = 5; the LDGSTS kernel stays the default while it measures faster, profiles/round2): tensor
  // maps of the 11 fields (cached per pointer); a driver without cuTensorMapEncodeTiled falls back to LDGSTS
  bool tma = tiled && !tile36 && g_gcm_knob[4] == 5 && (size_t)nbatch * L < 2147483647u;
  PfTmaMaps maps;
  const int tjt = g_gcm_knob[11] == 8 ? 8 : 4;                          // tile rows
  const int nst = (g_gcm_knob[10] != 3 && tjt == 4) ? 4 : 3;  // layers in flight (knob 10 = 3: three)
  if (tma && nrowsU > 0) {
    const double* hf[PFT_NF] = {star->u, star->v, star->t, star->q, w.spu};
    const double* cf[6] = {w.pgf, w.fv, base->u, base->v, base->t, base->q};
    for (int f = 0; f < PFT_NF && tma; ++f)
      tma = gcm_tmap_get(&maps.halo[f], hf[f], W, H, nbatch * L, PFT_ROW, tjt + 2) == GCM_OK;
    for (int f = 0; f < 6 && tma; ++f) tma = gcm_tmap_get(&maps.cen[f], cf[f], W, H, nbatch * L, PFT_TI, tjt) == GCM_OK;
  }
  if (nrowsU > 0 && tma) {
    GcmProfScope ps(GCM_K_UPDATE_TMA, stream);
    const bool same = base->u == star->u && base->v == star->v && base->t == star->t && base->q == star->q;
    const GcmRowSeg parts[2] = {{segU.a, segU.n1, 0, 0}, {segU.c, segU.n2, 0, 0}};
    const int HT = ((tjt + 2) * PFT_ROW + 15) / 16 * 16, CT = tjt * PFT_TI;
    const size_t smt = 128 + (size_t)nst * (PFT_NF * HT + (same ? 2 : 6) * CT) * sizeof(double);
    for (int s2 = 0; s2 < 2; ++s2) {
      if (parts[s2].n1 <= 0) continue;
      const dim3 gridt(W / PFT_TI, (parts[s2].n1 + tjt - 1) / tjt, nbatch), blockt(PFT_TI, tjt + 1);  // + producer warp
#ifndef GCM_EMU
#define PF_TMA_ATTR(K) \
  if (smt > 48 * 1024) GCM_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smt))
#else
#define PF_TMA_ATTR(K)
#endif
#define PF_TMA_GO(TJ_, SAME_, NS_, MB_)                                                                                   \
  do {                                                                                                                    \
    PF_TMA_ATTR((pe25f_update_tma_kernel<L, TJ_, SAME_, NS_, MB_>));                                                      \
    GCM_LAUNCH_DEP((pe25f_update_tma_kernel<L, TJ_, SAME_, NS_, MB_>), gridt, blockt, smt, stream, d, maps, cb, cs, mo, w, \
                   dt, parts[s2], b2, b3);                                                                                \
  } while (0)
#define PF_TMA_T4(SAME_)                                  \
  do {                                                    \
    if (nst == 4 && minb4) PF_TMA_GO(4, SAME_, 4, 4);     \
    else if (nst == 4) PF_TMA_GO(4, SAME_, 4, 3);         \
    else if (minb4) PF_TMA_GO(4, SAME_, 3, 4);            \
    else PF_TMA_GO(4, SAME_, 3, 3);                       \
  } while (0)
      const bool minb4 = g_gcm_knob[13] == 1;  // 4 CTAs per SM at 96 registers (spills) instead of 3 at 128
      if (tjt == 8) {
        if (same) PF_TMA_GO(8, true, 3, 2); else PF_TMA_GO(8, false, 3, 2);
      } else {
        if (same) PF_TMA_T4(true); else PF_TMA_T4(false);
      }
      GCM_CHECK_LAUNCH();
    }
  } else if (nrowsU > 0 && tiled) {
    GcmProfScope ps(GCM_K_UPDATE_TILED, stream);
    // a tile needs contiguous rows: a two-segment launch becomes one launch per segment (same kernel for every
    // row, so a band stays bit-identical to the whole grid)
    const GcmRowSeg parts[2] = {{segU.a, segU.n1, 0, 0}, {segU.c, segU.n2, 0, 0}};
    constexpr int tj = 4;  // tile height (8 measured slower on B200: r02a)
    const int ti = tile36 ? 36 : PFT_TI;
    const size_t smt = (size_t)PFT_NS * PFT_NF * (tj + 2) * (ti + 4) * sizeof(double);
    for (int s2 = 0; s2 < 2; ++s2) {
      if (parts[s2].n1 <= 0) continue;
      const dim3 gridt(W / ti, (parts[s2].n1 + tj - 1) / tj, nbatch), blockt(ti, tj);
      const int mbu = (g_gcm_knob[15] / 100) % 10;  // register-budget variants (knob 15, hundreds digit)
      if (tile36) {
        if (fuse_aflux)
          GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 3, 36, true>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
        else if (mbu == 4)  // 4 CTAs of 144 threads per SM at 112 registers instead of 3 at 128
          GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 4, 36>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
        else
          GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 3, 36>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
      } else if (fuse_aflux) {
        GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 4, PFT_TI, true>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
      } else if (L == 9 && (mbu == 5 || mbu == 6)) {
        if constexpr (L == 9) {
          if (mbu == 5) GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 5>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
          else GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj, 6>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
        }
      } else
        GCM_LAUNCH_DEP((pe25f_update_tiled_kernel<L, tj>), gridt, blockt, smt, stream, d, cb, cs, mo, w, dt, parts[s2], b2, b3);
      GCM_CHECK_LAUNCH();
    }
  } else if (nrowsU > 0) {
    // direct loads: 32 x 4 (i x j) tiles; rows shorter than 128 that do not fill 32-wide tiles run flat over
    // (row, column)
    GcmProfScope ps(GCM_K_UPDATE_FAST, stream);
    if (cells) {
      const dim3 gridc((nrowsU * W + 127) / 128, L, nbatch);
      GCM_LAUNCH_DEP((pe25f_update_cell_kernel<L>), gridc, dim3(128), 0, stream, d, cb, cs, mo, w, dt, segU,
                     gcm_magic((unsigned)W), b2, b3);
      GCM_CHECK_LAUNCH();
      return GCM_OK;
    }
    const bool flat = W < 128 && W % 32 != 0 && (size_t)nrowsU * W < (1u << 22);
    const unsigned flatW = flat ? gcm_magic((unsigned)W) : 0u;
    const dim3 block(32, 4);
    const dim3 grid = flat ? dim3((nrowsU * W + 127) / 128, 1, nbatch) : dim3((W + 31) / 32, (nrowsU + 3) / 4, nbatch);
    const int pfd = g_gcm_knob[5] > 0 ? g_gcm_knob[5] - 1 : 1;
    GCM_LAUNCH_DEP((pe25f_update_kernel<L>), grid, block, 0, stream, d, cb, cs, mo, w, dt, segU, pfd, flatW, b2, b3);
    GCM_CHECK_LAUNCH();
  }
  return GCM_OK;
}


// one translation unit per layer count (pe25_fast_l*.cu): the kernels are unrolled over the layers, and the units
// compile side by side
#define GCM_PF_INSTANTIATE(LAYERS)                                                                                       \
  int gcm_pf_half_step_l##LAYERS(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out, \
                                 double dt, int nbatch, const PfWork& w, GcmRowSeg sr, GcmRowSeg su, void* stream) {    \
    return pf_half_step<LAYERS>(g, base, star, out, dt, nbatch, w, sr, su, stream);                                     \
  }
