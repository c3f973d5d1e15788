// sw2d.cu -- 2-D schemes on a uniform doubly periodic C-grid:
//   matsuno_c_grid.matsumo_scheme + its 5 operators   (reference matsuno_c_grid.py:15-142)
//   matsumo_temp.matsumo_scheme                        (reference matsumo_temp.py:13-99)
//   viscosity.finite_laplacian_2d / incompressible_viscosity_2d (reference viscosity.py:12-25)
//
// Only + - * / appear in the shallow-water scheme, so this translation unit is compiled with
// -fmad=false and keeps the reference's operation order: the results are bit-identical to numpy.
#include "gcm_common.h"
#include "stencil2d.h"

// ---------------------------------------------------------------------------------------------------
// operators (one cell), written once for global arrays and shared-memory tiles
// ---------------------------------------------------------------------------------------------------
// matsuno_c_grid.py:15-51
__device__ __forceinline__ double sw_adv_u(const double* u, const double* v, const GcmNb& n, double dx) {
  const double u_c = n(u, 0, 0), u_ip = n(u, 0, 1), u_im = n(u, 0, -1), u_jp = n(u, 1, 0), u_jm = n(u, -1, 0);
  const double a_ipj = (u_ip + u_c) / 2;
  const double a_imj = (u_im + u_c) / 2;
  const double v_ijm = (n(v, 0, -1) + n(v, 0, 0)) / 2;
  const double v_ijp = (n(v, 1, -1) + n(v, 1, 0)) / 2;
  const double du_ipj = u_ip - u_c, du_imj = u_c - u_im, du_ijp = u_jp - u_c, du_ijm = u_c - u_jm;
  return (a_ipj * du_ipj + a_imj * du_imj + v_ijp * du_ijp + v_ijm * du_ijm) / dx;
}
// matsuno_c_grid.py:54-80
__device__ __forceinline__ double sw_adv_v(const double* u, const double* v, const GcmNb& n, double dx) {
  const double v_c = n(v, 0, 0), v_ip = n(v, 0, 1), v_im = n(v, 0, -1), v_jp = n(v, 1, 0), v_jm = n(v, -1, 0);
  const double a_ijp = (v_jp + v_c) / 2;
  const double a_ijm = (v_jm + v_c) / 2;
  const double u_ipj = (n(u, 0, 0) + n(u, -1, 0)) / 2;
  const double u_imj = (n(u, 0, -1) + n(u, 1, -1)) / 2;
  const double dv_ipj = v_ip - v_c, dv_imj = v_c - v_im, dv_ijp = v_jp - v_c, dv_ijm = v_c - v_jm;
  return (u_ipj * dv_ipj + u_imj * dv_imj + a_ijp * dv_ijp + a_ijm * dv_ijm) / dx;
}
// matsuno_c_grid.py:97, :103
__device__ __forceinline__ double sw_grad_u(const double* p, const GcmNb& n, double dx) {
  return (n(p, 0, 1) - n(p, 0, 0)) / dx * GCM_G;
}
__device__ __forceinline__ double sw_grad_v(const double* p, const GcmNb& n, double dx) {
  return (n(p, 1, 0) - n(p, 0, 0)) / dx * GCM_G;
}
// matsuno_c_grid.py:109-118
__device__ __forceinline__ double sw_adv_p(const double* u, const double* v, const double* p, const GcmNb& n,
                                           double dx) {
  const double p_c = n(p, 0, 0);
  const double up_imj = (n(p, 0, -1) + p_c) / 2 * n(u, 0, -1);
  const double up_ipj = (n(p, 0, 1) + p_c) / 2 * n(u, 0, 0);
  const double vp_ijm = (n(p, -1, 0) + p_c) / 2 * n(v, -1, 0);
  const double vp_ijp = (n(p, 1, 0) + p_c) / 2 * n(v, 0, 0);
  return (up_ipj - up_imj) / dx + (vp_ijp - vp_ijm) / dx;
}
// viscosity.py:12-19
__device__ __forceinline__ double sw_laplacian(const double* q, const GcmNb& n, double dx) {
  const double top = n(q, 1, 0) + n(q, -1, 0) + n(q, 0, 1) + n(q, 0, -1) - 4 * n(q, 0, 0);
  return top / (dx * dx);
}

// ---------------------------------------------------------------------------------------------------
// stand-alone operators
// ---------------------------------------------------------------------------------------------------
__global__ void sw2d_operator_kernel(int op, const double* __restrict__ u, const double* __restrict__ v,
                                     const double* __restrict__ p, double* __restrict__ out, int H, int W, double dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  const GcmNb n = gcm_nb_global(j, i, H, W);
  double r;
  switch (op) {
    case 0: r = sw_adv_u(u, v, n, dx); break;
    case 1: r = sw_adv_v(u, v, n, dx); break;
    case 2: r = sw_grad_u(p, n, dx); break;
    case 3: r = sw_grad_v(p, n, dx); break;
    default: r = sw_adv_p(u, v, p, n, dx); break;
  }
  out[(size_t)j * W + i] = r;
}

extern "C" int gcm_sw2d_operator(int op, const double* u, const double* v, const double* p, double* out, int H, int W,
                                 double dx, void* stream) {
  GCM_REQUIRE(out, GCM_ENULL);
  GCM_REQUIRE(op >= 0 && op <= 4, GCM_EUNSUP);
  GCM_REQUIRE(H > 0 && W > 0, GCM_ESHAPE);
  if (op <= 1 || op == 4) GCM_REQUIRE(u && v, GCM_ENULL);
  if (op >= 2) GCM_REQUIRE(p, GCM_ENULL);
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  GCM_LAUNCH(sw2d_operator_kernel, dim3((W + tc - 1) / tc, H, 1), dim3(tc), 0, stream, op, u, v, p, out, H, W, dx);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

__global__ void laplacian5_kernel(const double* __restrict__ q, double* __restrict__ out, int H, int W, double dx,
                                  double mu, int apply_mu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  const GcmNb n = gcm_nb_global(j, i, H, W);
  const double lap = sw_laplacian(q, n, dx);
  out[(size_t)j * W + i] = apply_mu ? mu * lap : lap;
}

extern "C" int gcm_laplacian5(const double* q, double* out, int H, int W, double dx, double mu, int apply_mu,
                              void* stream) {
  GCM_REQUIRE(q && out, GCM_ENULL);
  GCM_REQUIRE(H > 0 && W > 0, GCM_ESHAPE);
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  GCM_LAUNCH(laplacian5_kernel, dim3((W + tc - 1) / tc, H, 1), dim3(tc), 0, stream, q, out, H, W, dx, mu, apply_mu);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// Matsuno step, tiled: one launch = predictor + corrector (matsuno_c_grid.py:125-142).
// A CTA owns a TH x TW tile of outputs, stages the (TH+4) x (TW+4) base tile in shared memory, forms the
// star state on the (TH+2) x (TW+2) ring-1 tile, then the corrected state on its own tile.
// ---------------------------------------------------------------------------------------------------
#define SW_TW 32
#define SW_TH 16
#define SW_BP (SW_TW + 4)  // base tile pitch
#define SW_SP (SW_TW + 2)  // star tile pitch

__global__ void __launch_bounds__(256) sw2d_matsuno_tile_kernel(const double* __restrict__ u, const double* __restrict__ v,
                                                                const double* __restrict__ p, double* __restrict__ uo,
                                                                double* __restrict__ vo, double* __restrict__ po, int H,
                                                                int W, double dx, double dt, unsigned int* nonfinite) {
  __shared__ double bu[(SW_TH + 4) * SW_BP], bv[(SW_TH + 4) * SW_BP], bp[(SW_TH + 4) * SW_BP];
  __shared__ double su[(SW_TH + 2) * SW_SP], sv[(SW_TH + 2) * SW_SP], sp[(SW_TH + 2) * SW_SP];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int j0 = blockIdx.y * SW_TH, i0 = blockIdx.x * SW_TW;
  for (int e = tid; e < (SW_TH + 4) * SW_BP; e += nthr) {
    const int tj = e / SW_BP, ti = e % SW_BP;
    const size_t g = (size_t)gcm_wrap(j0 + tj - 2, H) * W + gcm_wrap(i0 + ti - 2, W);
    bu[e] = u[g]; bv[e] = v[g]; bp[e] = p[g];
  }
  __syncthreads();
  for (int e = tid; e < (SW_TH + 2) * SW_SP; e += nthr) {  // star tile cell (sj, si) = base tile cell (sj+1, si+1)
    const int sj = e / SW_SP, si = e % SW_SP;
    const GcmNb n = gcm_nb_tile(sj + 1, si + 1, SW_BP);
    const int c = (sj + 1) * SW_BP + si + 1;
    su[e] = bu[c] - dt * (sw_adv_u(bu, bv, n, dx) + sw_grad_u(bp, n, dx));
    sv[e] = bv[c] - dt * (sw_adv_v(bu, bv, n, dx) + sw_grad_v(bp, n, dx));
    sp[e] = bp[c] - dt * sw_adv_p(bu, bv, bp, n, dx);
  }
  __syncthreads();
  bool bad = false;
  for (int e = tid; e < SW_TH * SW_TW; e += nthr) {
    const int tj = e / SW_TW, ti = e % SW_TW;
    const int j = j0 + tj, i = i0 + ti;
    if (j >= H || i >= W) continue;
    const GcmNb n = gcm_nb_tile(tj + 1, ti + 1, SW_SP);
    const int c = (tj + 2) * SW_BP + ti + 2;
    const size_t g = (size_t)j * W + i;
    const double u_n = bu[c] - dt * (sw_adv_u(su, sv, n, dx) + sw_grad_u(sp, n, dx));
    const double v_n = bv[c] - dt * (sw_adv_v(su, sv, n, dx) + sw_grad_v(sp, n, dx));
    const double p_n = bp[c] - dt * sw_adv_p(su, sv, sp, n, dx);
    uo[g] = u_n;
    vo[g] = v_n;
    po[g] = p_n;
    bad |= gcm_not_finite(u_n + v_n + p_n);
  }
  gcm_flag_nonfinite(nonfinite, bad);
}

// Matsuno steps, resident: the whole grid lives in one CTA's shared memory for all `nsteps` steps (the
// reference's own 64 x 64 case is 96 KB of state).  base is updated in place: the corrector at a cell reads
// base only at that cell, star at its neighbours.
__global__ void __launch_bounds__(1024) sw2d_matsuno_resident_kernel(const double* __restrict__ u,
                                                                     const double* __restrict__ v,
                                                                     const double* __restrict__ p, double* __restrict__ uo,
                                                                     double* __restrict__ vo, double* __restrict__ po,
                                                                     int H, int W, double dx, double dt, int nsteps,
                                                                     unsigned int* nonfinite) {
  GCM_DYN_SMEM(double, smem);
  const int n2 = H * W;
  double *bu = smem, *bv = bu + n2, *bp = bv + n2, *su = bp + n2, *sv = su + n2, *sp = sv + n2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int e = tid; e < n2; e += nthr) { bu[e] = u[e]; bv[e] = v[e]; bp[e] = p[e]; }
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    for (int e = tid; e < n2; e += nthr) {
      const GcmNb n = gcm_nb_global(e / W, e % W, H, W);
      su[e] = bu[e] - dt * (sw_adv_u(bu, bv, n, dx) + sw_grad_u(bp, n, dx));
      sv[e] = bv[e] - dt * (sw_adv_v(bu, bv, n, dx) + sw_grad_v(bp, n, dx));
      sp[e] = bp[e] - dt * sw_adv_p(bu, bv, bp, n, dx);
    }
    __syncthreads();
    for (int e = tid; e < n2; e += nthr) {
      const GcmNb n = gcm_nb_global(e / W, e % W, H, W);
      bu[e] = bu[e] - dt * (sw_adv_u(su, sv, n, dx) + sw_grad_u(sp, n, dx));
      bv[e] = bv[e] - dt * (sw_adv_v(su, sv, n, dx) + sw_grad_v(sp, n, dx));
      bp[e] = bp[e] - dt * sw_adv_p(su, sv, sp, n, dx);
    }
    __syncthreads();
  }
  bool bad = false;
  for (int e = tid; e < n2; e += nthr) {
    uo[e] = bu[e]; vo[e] = bv[e]; po[e] = bp[e];
    bad |= gcm_not_finite(bu[e] + bv[e] + bp[e]);
  }
  gcm_flag_nonfinite(nonfinite, bad);
}

#define SW_RESIDENT_MAX_BYTES (220 * 1024)

extern "C" size_t gcm_sw2d_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  return 3 * n2 * sizeof(double);  // one ping-pong state for the tiled path
}

extern "C" int gcm_sw2d_matsuno_step(const double* u, const double* v, const double* p, double* uo, double* vo,
                                     double* po, int H, int W, double dx, double dt, int nsteps, void* ws,
                                     size_t ws_bytes, void* stream) {
  GCM_REQUIRE(u && v && p && uo && vo && po, GCM_ENULL);
  GCM_REQUIRE(H > 0 && W > 0 && nsteps > 0, GCM_ESHAPE);
  const size_t resident = 6 * (size_t)H * W * sizeof(double);
  if (resident <= SW_RESIDENT_MAX_BYTES) {
#ifndef GCM_EMU
    GCM_CUDA(cudaFuncSetAttribute(sw2d_matsuno_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)resident));
#endif
    const int n2 = H * W;
    const int thr = n2 >= 1024 ? 1024 : (n2 + 31) / 32 * 32;
    GCM_LAUNCH(sw2d_matsuno_resident_kernel, dim3(1), dim3(thr), resident, stream, u, v, p, uo, vo, po, H, W, dx, dt,
               nsteps, gcm_nonfinite_word());
    GCM_CHECK_LAUNCH();
    return GCM_OK;
  }
  GCM_REQUIRE(nsteps == 1 || ws, GCM_ENULL);
  GCM_REQUIRE(nsteps == 1 || ws_bytes >= gcm_sw2d_workspace_bytes(H, W), GCM_EWORK);
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  double* t[3] = {(double*)ws, (double*)ws + n2, (double*)ws + 2 * n2};
  const double *cu = u, *cv = v, *cp = p;
  const dim3 grid((W + SW_TW - 1) / SW_TW, (H + SW_TH - 1) / SW_TH, 1);
  for (int s = 0; s < nsteps; ++s) {
    const bool to_out = (nsteps - 1 - s) % 2 == 0;
    double* du = to_out ? uo : t[0];
    double* dv = to_out ? vo : t[1];
    double* dp = to_out ? po : t[2];
    GCM_LAUNCH(sw2d_matsuno_tile_kernel, grid, dim3(256), 0, stream, cu, cv, cp, du, dv, dp, H, W, dx, dt,
               gcm_nonfinite_word());
    GCM_CHECK_LAUNCH();
    cu = du; cv = dv; cp = dp;
  }
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// matsumo_temp.matsumo_scheme (matsumo_temp.py:66-99): shallow water + temperature + viscosity.
// Two launches per step (star state in the workspace).  Keeps the reference quirk that the v equation
// is damped with the Laplacian of u (:75, :91).
// ---------------------------------------------------------------------------------------------------
// matsumo_temp.py:13-16
__device__ __forceinline__ double swt_density(double p, double t) {
  const double temp = t / pow(100000.0 / p, GCM_RD / GCM_CP);
  return p / (GCM_RD * temp);
}

__global__ void swt2d_half_kernel(const double* __restrict__ u, const double* __restrict__ v,
                                  const double* __restrict__ p, const double* __restrict__ t,
                                  const double* __restrict__ su, const double* __restrict__ sv,
                                  const double* __restrict__ sp, const double* __restrict__ st, double* __restrict__ uo,
                                  double* __restrict__ vo, double* __restrict__ po, double* __restrict__ to, int H, int W,
                                  double dx, double dt, double mu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  const GcmNb n = gcm_nb_global(j, i, H, W);
  const size_t c = (size_t)j * W + i;
  // geopotential = p / (G rho) at the cell and its i+1, j+1 neighbours (:69-70, :85-86)
  const double rho_c = swt_density(n(sp, 0, 0), n(st, 0, 0));
  const double geo_c = n(sp, 0, 0) / (GCM_G * rho_c);
  const double geo_ip = n(sp, 0, 1) / (GCM_G * swt_density(n(sp, 0, 1), n(st, 0, 1)));
  const double geo_jp = n(sp, 1, 0) / (GCM_G * swt_density(n(sp, 1, 0), n(st, 1, 0)));
  const double visc = mu * sw_laplacian(su, n, dx) / rho_c;
  const double fu = sw_adv_u(su, sv, n, dx) + (geo_ip - geo_c) / dx * GCM_G - visc;
  const double fv = sw_adv_v(su, sv, n, dx) + (geo_jp - geo_c) / dx * GCM_G - visc;
  uo[c] = u[c] - dt * fu;
  vo[c] = v[c] - dt * fv;
  const double p_n = p[c] - dt * sw_adv_p(su, sv, sp, n, dx);
  po[c] = p_n;
  // flux-form advection of the scaled temperature p t dx dx (:67, :79-82, :84, :95-97)
  const double sc_c = sp[c] * st[c] * dx * dx;
  const double sc_im = n(sp, 0, -1) * n(st, 0, -1) * dx * dx, sc_ip = n(sp, 0, 1) * n(st, 0, 1) * dx * dx;
  const double sc_jm = n(sp, -1, 0) * n(st, -1, 0) * dx * dx, sc_jp = n(sp, 1, 0) * n(st, 1, 0) * dx * dx;
  const double up_imj = (sc_im + sc_c) / 2 * n(su, 0, -1);
  const double up_ipj = (sc_ip + sc_c) / 2 * n(su, 0, 0);
  const double vp_ijm = (sc_jm + sc_c) / 2 * n(sv, -1, 0);
  const double vp_ijp = (sc_jp + sc_c) / 2 * n(sv, 0, 0);
  const double adv = (up_ipj - up_imj) / dx + (vp_ijp - vp_ijm) / dx;
  const double scaled_base = p[c] * t[c] * dx * dx;
  to[c] = (scaled_base - dt * adv) / (p_n * dx * dx);
}

extern "C" size_t gcm_swt2d_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  return 8 * n2 * sizeof(double);  // star state + one ping-pong state
}

extern "C" int gcm_swt2d_matsuno_step(const double* u, const double* v, const double* p, const double* t, double* uo,
                                      double* vo, double* po, double* to, int H, int W, double dx, double dt, double mu,
                                      int nsteps, void* ws, size_t ws_bytes, void* stream) {
  GCM_REQUIRE(u && v && p && t && uo && vo && po && to && ws, GCM_ENULL);
  GCM_REQUIRE(H > 0 && W > 0 && nsteps > 0, GCM_ESHAPE);
  GCM_REQUIRE(ws_bytes >= gcm_swt2d_workspace_bytes(H, W), GCM_EWORK);
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  double* w = (double*)ws;
  double *su = w, *sv = w + n2, *sp = w + 2 * n2, *st = w + 3 * n2;
  double* tmp[4] = {w + 4 * n2, w + 5 * n2, w + 6 * n2, w + 7 * n2};
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  const dim3 grid((W + tc - 1) / tc, H, 1);
  const double *cu = u, *cv = v, *cp = p, *ct = t;
  for (int s = 0; s < nsteps; ++s) {
    const bool to_out = (nsteps - 1 - s) % 2 == 0;
    double* du = to_out ? uo : tmp[0];
    double* dv = to_out ? vo : tmp[1];
    double* dp = to_out ? po : tmp[2];
    double* dtt = to_out ? to : tmp[3];
    GCM_LAUNCH(swt2d_half_kernel, grid, dim3(tc), 0, stream, cu, cv, cp, ct, cu, cv, cp, ct, su, sv, sp, st, H, W, dx, dt,
               mu);
    GCM_CHECK_LAUNCH();
    GCM_LAUNCH(swt2d_half_kernel, grid, dim3(tc), 0, stream, cu, cv, cp, ct, (const double*)su, (const double*)sv,
               (const double*)sp, (const double*)st, du, dv, dp, dtt, H, W, dx, dt, mu);
    GCM_CHECK_LAUNCH();
    cu = du; cv = dv; cp = dp; ct = dtt;
  }
  return GCM_OK;
}
