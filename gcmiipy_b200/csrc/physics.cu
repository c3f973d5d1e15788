// physics.cu -- grey-radiation column physics, the step after the dynamics in no_limits_2_5d.full_timestep
// (reference grey_solar.py:40-68 zenith_angle, :323-333 basic_grey_transmittances, :358-563 basic_grey_radiation,
// no_limits_2_5d.py:66-75 solar_timestep).  Columns are independent: one thread per column,
// every layer quantity in registers, the layer recursions (down- and up-welling long wave) sequential in k exactly as
// the reference loops.  Compiled with -fmad=false and the reference's operation order: everything but cos() of the
// hour angle and pow() of the Exner factor is bit-identical to numpy.
//
// Per-layer scalars (transmittances and their cumulative products, identical for every column because the reference's
// t ** dsig depends on the layer only) are evaluated once on the host, in the reference's order, and ride in the kernel
// parameters.
#include <math.h>

#include "gcm_common.h"

#define GCM_SB 5.67e-8              /* W m-2 K-4, constants.py:71 */
#define GCM_SOLAR (1.3608 * 1000.0) /* W m-2,     constants.py:59 */
#define GCM_CG 1.13e6               /* J K-1 m-3, constants.py:25 */

struct GreyTabs {
  int L;
  double lw[GCM_MAXLC], sw[GCM_MAXLC];  // transmittance of layer k               (grey_solar.py:323-333)
  double clw_b_div[GCM_MAXLC];          // cumprod(lw)[k] / lw[k]                  (:373-377)
  double cum_sw_top[GCM_MAXLC];         // cumprod(sw[::-1])[::-1][k]              (:371)
  double dsig[GCM_MAXLC];
};

// cos of the solar zenith angle, declination 0 (grey_solar.py:40-46, :49-68), clipped at 0
__device__ __forceinline__ double grey_sza(double sinlat, double coslat, double lon, double hour_angle) {
  const double pa = lon + hour_angle;
  const double v = sinlat * 0.0 + coslat * 1.0 * cos(pa);
  return v > 0.0 ? v : (v == v ? 0.0 : v);  // np.maximum(x, 0): NaN propagates
}

// one column: tt[k] true temperature -> dTdt[k], returns d(ground temperature)/dt
template <class TT>
__device__ __forceinline__ double grey_column(const GreyTabs& tb, TT tt, double p_s, double gt, double sza, double albedo,
                                              double* dTdt) {
  const int L = tb.L;
  double emission[GCM_MAXLC], lwa_a[GCM_MAXLC], lwa_b[GCM_MAXLC];
  double B = 0.0;
  for (int k = 0; k < L; ++k) {
    const double t1 = tt(k), t2 = t1 * t1;
    emission[k] = (1 - tb.lw[k]) * GCM_SB * (t2 * t2);  // 2.25
    const double term = emission[k] * tb.clw_b_div[k];
    B = k == 0 ? term : B + term;
  }
  const double Sc = GCM_SOLAR * sza;
  const double S = (1 - albedo) * Sc * tb.cum_sw_top[0];  // 2.26
  const double g2 = gt * gt;
  const double U_s = 1 * GCM_SB * (g2 * g2);              // 2.27
  const double dt_ground = (B + S - U_s) / GCM_CG / 0.1;
  double dw = 0.0;
  for (int k = L - 1; k >= 0; --k) {  // long wave from above (:467-475)
    lwa_a[k] = dw * (1 - tb.lw[k]);
    dw = dw * tb.lw[k] + emission[k];
  }
  double uw = 0.0;
  for (int k = 0; k < L; ++k) {       // long wave from the layers below (:500-503)
    lwa_b[k] = uw * (1 - tb.lw[k]);
    uw = uw * tb.lw[k] + emission[k];
  }
  for (int k = 0; k < L; ++k) {
    const double U_n = tb.clw_b_div[k] * U_s * (1 - tb.lw[k]);                    // 2.30
    const double S_n = (1 - tb.sw[k]) * tb.cum_sw_top[k] / tb.sw[k] * Sc;         // 2.31
    const double B_n = emission[k];                                              // 2.32
    dTdt[k] = (U_n + S_n - 2 * B_n + lwa_a[k] + lwa_b[k]) * (GCM_G / (GCM_CP * p_s * tb.dsig[k]));  // 2.34
  }
  return dt_ground;
}

__global__ void __launch_bounds__(128)
grey_radiation_kernel(GreyTabs tb, int H, int W, const double* __restrict__ p, const double* __restrict__ tt,
                      const double* __restrict__ gt, const double* __restrict__ sinlat, const double* __restrict__ coslat,
                      const double* __restrict__ lon, double hour_angle, double albedo, double* __restrict__ dTdt,
                      double* __restrict__ dt_ground) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H * W) return;
  const int j = c / W, i = c - j * W;
  const size_t plane = (size_t)H * W;
  const double sza = grey_sza(sinlat[j], coslat[j], lon[i], hour_angle);
  double out[GCM_MAXLC];
  const double dtg = grey_column(tb, [&](int k) { return tt[k * plane + c]; }, p[c], gt[c], sza, albedo, out);
  for (int k = 0; k < tb.L; ++k) dTdt[k * plane + c] = out[k];
  dt_ground[c] = dtg;
}

// no_limits_2_5d.solar_timestep (:66-75): theta -> true temperature, radiation, explicit update, back to theta
__global__ void __launch_bounds__(128)
solar_timestep_kernel(GreyTabs tb, GcmGeomDev g, const double* __restrict__ p, const double* __restrict__ t,
                      const double* __restrict__ gt, const double* __restrict__ sinlat, const double* __restrict__ coslat,
                      const double* __restrict__ lon, double hour_angle, double albedo, double dt, double* __restrict__ t_n,
                      double* __restrict__ gt_n) {
  const int H = g.H, W = g.W;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H * W) return;
  const int j = c / W, i = c - j * W;
  const size_t plane = (size_t)H * W;
  const double p_s = p[c];
  double tt[GCM_MAXLC], ex[GCM_MAXLC], out[GCM_MAXLC];
  for (int k = 0; k < tb.L; ++k) {
    const double tp = p_s * g.sig[k] + g.ptop;
    ex[k] = pow(GCM_P0 / tp, GCM_KAPPA);
    tt[k] = t[k * plane + c] / ex[k];  // temperature.to_true_temp
  }
  const double sza = grey_sza(sinlat[j], coslat[j], lon[i], hour_angle);
  const double gt_c = gt[c];
  const double dtg = grey_column(tb, [&](int k) { return tt[k]; }, p_s, gt_c, sza, albedo, out);
  for (int k = 0; k < tb.L; ++k) t_n[k * plane + c] = (tt[k] + out[k] * dt) * ex[k];  // temperature.to_potential_temp
  gt_n[c] = gt_c + dtg * dt;
  bool bad = gcm_not_finite(gt_c + dtg * dt);
  gcm_flag_nonfinite(g.nonfinite, bad);
}

static int grey_tabs(const gcm_geom* g, const double* h_lw, const double* h_sw, GreyTabs* tb) {
  const int L = g->d.L;
  GCM_REQUIRE(L >= 1 && L <= GCM_MAXLC, GCM_EUNSUP);
  tb->L = L;
  double cum = 1.0;
  for (int k = 0; k < L; ++k) {
    tb->lw[k] = h_lw[k];
    tb->sw[k] = h_sw[k];
    tb->dsig[k] = g->d.c_dsig[k];
    cum = k == 0 ? h_lw[0] : cum * h_lw[k];  // np.cumprod(lw)
    tb->clw_b_div[k] = cum / h_lw[k];
  }
  double cs = 1.0;
  for (int k = L - 1; k >= 0; --k) {  // np.cumprod(sw[::-1])[::-1]
    cs = k == L - 1 ? h_sw[k] : cs * h_sw[k];
    tb->cum_sw_top[k] = cs;
  }
  return GCM_OK;
}

// h_lw / h_sw: HOST arrays [L], the per-layer transmittances t_lw ** dsig, t_sw ** dsig as the caller's numpy evaluates
// them (grey_solar.py:323-333); sinlat / coslat [H], lon [W]: DEVICE tables (radians); hour_angle in radians (:51)
extern "C" int gcm_grey_radiation(const gcm_geom* g, const double* p, const double* tt, const double* gt,
                                  const double* h_lw, const double* h_sw, double albedo, const double* sinlat,
                                  const double* coslat, const double* lon, double hour_angle, double* dTdt,
                                  double* dt_ground, void* stream) {
  GCM_REQUIRE(g && p && tt && gt && h_lw && h_sw && sinlat && coslat && lon && dTdt && dt_ground, GCM_ENULL);
  GreyTabs tb;
  int st = grey_tabs(g, h_lw, h_sw, &tb);
  if (st) return st;
  const int n = g->d.H * g->d.W;
  GCM_LAUNCH(grey_radiation_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, tb, g->d.H, g->d.W, p, tt, gt, sinlat, coslat,
             lon, hour_angle, albedo, dTdt, dt_ground);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

extern "C" int gcm_solar_timestep(const gcm_geom* g, const double* p, const double* t, const double* gt, const double* h_lw,
                                  const double* h_sw, double albedo, const double* sinlat, const double* coslat,
                                  const double* lon, double hour_angle, double dt, double* t_n, double* gt_n, void* stream) {
  GCM_REQUIRE(g && p && t && gt && h_lw && h_sw && sinlat && coslat && lon && t_n && gt_n, GCM_ENULL);
  GreyTabs tb;
  int st = grey_tabs(g, h_lw, h_sw, &tb);
  if (st) return st;
  const int n = g->d.H * g->d.W;
  GCM_LAUNCH(solar_timestep_kernel, dim3((n + 127) / 128), dim3(128), 0, stream, tb, g->d, p, t, gt, sinlat, coslat, lon,
             hour_angle, albedo, dt, t_n, gt_n);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}
