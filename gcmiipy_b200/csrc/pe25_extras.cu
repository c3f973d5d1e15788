// pe25_extras.cu -- opt-in terms of the 2.5-D half step, default OFF.
//
//   coriolis        dynamics.py:82-95 with the reference's dead `if False:` branch switched on
//   nu              horizontal viscosity pi * nu * lap(u), pi * nu * lap(v): viscosity.py:12-25 on the lat-lon metric
//                   (the mass-weighted form of matsumo_temp.py:55-59's mu * lap(u) / rho)
//   limit_q/limit_t van Leer flux-limited horizontal advection of q / theta: flux_limiter.py:10-32 (donor cell,
//                   calc_r, van_leer) composed into advec_t (dynamics.py:174-181; the TODO at dynamics.py:217-218)
//
// Every term is linear in the tendencies of the step, so it is applied to the state a half step has just
// written (`out`), by ONE extra launch that reads the star state and the filtered mass flux the step leaves in its
// workspace:
//   u_n += -dt (cor_u - visc_u) / iph(p_n)        v_n += -dt (cor_v - visc_v) / jph(p_n)
//   q_n += -dt div(G) / p_n,   G = mass flux * (limited edge value - centred edge value)
// The default path (no option set) launches nothing from this file and stays bit-identical to the reference step.
// Whole-grid geometries only (rows periodic in j): the limiter reads j - 2 ... j + 2.
#include <math.h>

#include "gcm_common.h"
#include "prof.h"

#define IDX3(k, j, i) (((size_t)(k) * H + (size_t)(j)) * W + (size_t)(i))
#define IDX2(j, i) ((size_t)(j) * W + (size_t)(i))

// flux_limiter.py:10
__device__ __forceinline__ double px_van_leer(double r) { return (r + fabs(r)) / (1 + fabs(r)); }

// limited edge value minus the centred one at the edge between q0 and q1 (qm, q0 | q1, q2), upwind by the sign
// of the mass flux (flux_limiter.py:23-27); slope ratio as calc_r (:14-20): 0 where the denominator is 0
__device__ __forceinline__ double px_edge_excess(double qm, double q0, double q1, double q2, double flux) {
  const double b = q1 - q0;
  double e;
  if (flux > 0) {
    const double a = q0 - qm;
    const double r = b != 0 ? a / b : 0.0;
    e = q0 + 0.5 * px_van_leer(r) * b;
  } else {
    const double c = q2 - q1;
    const double r = b != 0 ? c / b : 0.0;
    e = q1 - 0.5 * px_van_leer(r) * b;
  }
  return e - (q0 + q1) / 2;
}

// divergence of the correction flux G of one tracer at (k, j, i)
__device__ __forceinline__ double px_limiter_div(const double* __restrict__ f, const double* __restrict__ spu,
                                                 double spv_c, double spv_n, int k, int j, int jm1, int jm2, int jp1,
                                                 int jp2, int i, int im1, int im2, int ip1, int ip2, int H, int W,
                                                 double dx, double dy) {
  const double f0 = f[IDX3(k, j, i)];
  const double fw1 = f[IDX3(k, j, im1)], fw2 = f[IDX3(k, j, im2)];
  const double fe1 = f[IDX3(k, j, ip1)], fe2 = f[IDX3(k, j, ip2)];
  const double fn1 = f[IDX3(k, jm1, i)], fn2 = f[IDX3(k, jm2, i)];
  const double fs1 = f[IDX3(k, jp1, i)], fs2 = f[IDX3(k, jp2, i)];
  const double pu_c = spu[IDX3(k, j, i)], pu_w = spu[IDX3(k, j, im1)];
  const double gi_c = pu_c * px_edge_excess(fw1, f0, fe1, fe2, pu_c);    // edge i + 1/2
  const double gi_w = pu_w * px_edge_excess(fw2, fw1, f0, fe1, pu_w);    // edge i - 1/2
  const double gj_c = spv_c * px_edge_excess(fn1, f0, fs1, fs2, spv_c);  // edge j + 1/2
  const double gj_n = spv_n * px_edge_excess(fn2, fn1, f0, fs1, spv_n);  // edge j - 1/2
  return (gi_c - gi_w) / dx + (gj_c - gj_n) / dy;
}

__global__ void __launch_bounds__(128)
pe25x_extras_kernel(GcmGeomDev g, GcmExtras x, const double* __restrict__ sp, const double* __restrict__ su,
                    const double* __restrict__ sv, const double* __restrict__ st, const double* __restrict__ sq,
                    const double* __restrict__ spu, const double* __restrict__ pn, double* __restrict__ u,
                    double* __restrict__ v, double* __restrict__ t, double* __restrict__ q, double dt, size_t b2,
                    size_t b3) {
  const int H = g.H, W = g.W, L = g.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const int j = g.row_lo + blockIdx.y;
  const int k = blockIdx.z % L, b = blockIdx.z / L;
  sp += b * b2; pn += b * b2;
  su += b * b3; sv += b * b3; st += b * b3; sq += b * b3; spu += b * b3;
  u += b * b3; v += b * b3; t += b * b3; q += b * b3;

  const int ip1 = gcm_ip(i, W), im1 = gcm_im(i, W), ip2 = gcm_ip(ip1, W), im2 = gcm_im(im1, W);
  const int jp1 = gcm_row(j, 1, H, 1), jm1 = gcm_row(j, -1, H, 1);
  const int jp2 = gcm_row(jp1, 1, H, 1), jm2 = gcm_row(jm1, -1, H, 1);
  const size_t c = IDX3(k, j, i);
  const double dy = g.dy, dxj = g.dx_j[j], dxh = g.dx_h[j];

  const double sp_c = sp[IDX2(j, i)], sp_e = sp[IDX2(j, ip1)], sp_s = sp[IDX2(jp1, i)], sp_n = sp[IDX2(jm1, i)];
  const double pn_c = pn[IDX2(j, i)];
  // star mass flux in j (dynamics.py:191): spv = sv * jph(sp)
  const double spv_c = sv[c] * ((sp_c + sp_s) / 2);
  const double spv_n = sv[IDX3(k, jm1, i)] * ((sp_n + sp_c) / 2);

  if (x.coriolis || x.nu != 0.0) {
    double fu = 0.0, fv = 0.0;  // what is added to dut + dus + pgfu and to dvt + dvs + phiv + pgv
    if (x.coriolis) {
      const double sp_se = sp[IDX2(jp1, ip1)], sp_ne = sp[IDX2(jm1, ip1)];
      const double spv_e = sv[IDX3(k, j, ip1)] * ((sp_e + sp_se) / 2);
      const double spv_ne = sv[IDX3(k, jm1, ip1)] * ((sp_ne + sp_e) / 2);
      const double pv_at_pu = ((spv_c + spv_n) / 2 + (spv_e + spv_ne) / 2) / 2;  // iph(jmh(pv)), dynamics.py:87
      const double pu_at_pv = ((spu[c] + spu[IDX3(k, jp1, i)]) / 2 +
                               (spu[IDX3(k, j, im1)] + spu[IDX3(k, jp1, im1)]) / 2) / 2;  // imh(jph(pu)), :86
      fu += x.cor_u[j] * -pv_at_pu;  // :94
      fv += x.cor_v[j] * pu_at_pv;   // :95
    }
    const bool wall = (j == g.zero_v_row || j == g.zero_v_row2);  // v_n[:, -1, :] stays 0 (dynamics.py:222)
    if (x.nu != 0.0) {
      const double u0 = su[c];
      const double lap_u = (su[IDX3(k, j, ip1)] + su[IDX3(k, j, im1)] - 2 * u0) / (dxj * dxj) +
                           (su[IDX3(k, jp1, i)] + su[IDX3(k, jm1, i)] - 2 * u0) / (dy * dy);
      fu -= (sp_c + sp_e) / 2 * (x.nu * lap_u);
      if (!wall) {
        const double v0 = sv[c];
        const double lap_v = (sv[IDX3(k, j, ip1)] + sv[IDX3(k, j, im1)] - 2 * v0) / (dxh * dxh) +
                             (sv[IDX3(k, jp1, i)] + sv[IDX3(k, jm1, i)] - 2 * v0) / (dy * dy);
        fv -= (sp_c + sp_s) / 2 * (x.nu * lap_v);
      }
    }
    u[c] += -(fu * dt) / ((pn_c + pn[IDX2(j, ip1)]) / 2);
    if (!wall) v[c] += -(fv * dt) / ((pn_c + pn[IDX2(jp1, i)]) / 2);
  }
  if (x.limit_q)
    q[c] += -(px_limiter_div(sq, spu, spv_c, spv_n, k, j, jm1, jm2, jp1, jp2, i, im1, im2, ip1, ip2, H, W, dxj, dy) *
              dt) / pn_c;
  if (x.limit_t)
    t[c] += -(px_limiter_div(st, spu, spv_c, spv_n, k, j, jm1, jm2, jp1, jp2, i, im1, im2, ip1, ip2, H, W, dxj, dy) *
              dt) / pn_c;
}

int gcm_pe25_extras_apply(const gcm_geom* g, const gcm_state* star, const gcm_state* out, const double* spu, double dt,
                          int nbatch, void* stream) {
  const GcmGeomDev& d = g->d;
  GCM_REQUIRE(d.wrap_j, GCM_EUNSUP);
  const int H = d.H, W = d.W, L = d.L;
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  const unsigned gx = (unsigned)((W + tc - 1) / tc);
  {
    GcmProfScope ps(GCM_K_EXTRAS, stream);
    GCM_LAUNCH(pe25x_extras_kernel, dim3(gx, d.row_hi - d.row_lo, L * nbatch), dim3(tc), 0, stream, d, g->x, star->p,
               star->u, star->v, star->t, star->q, spu, out->p, out->u, out->v, out->t, out->q, dt, (size_t)H * W,
               (size_t)L * H * W);
  }
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

extern "C" int gcm_pe25_set_options(gcm_geom* g, const gcm_pe25_options* opt) {
  GCM_REQUIRE(g, GCM_ENULL);
  GcmExtras x = {0, 0, 0, 0.0, nullptr, nullptr};
  if (opt) {
    GCM_REQUIRE(opt->nu >= 0.0 && opt->nu == opt->nu, GCM_ESHAPE);
    x.coriolis = opt->coriolis ? 1 : 0;
    x.limit_q = opt->limit_q ? 1 : 0;
    x.limit_t = opt->limit_t ? 1 : 0;
    x.nu = opt->nu;
  }
  const bool any = x.coriolis || x.limit_q || x.limit_t || x.nu != 0.0;
  GCM_REQUIRE(!any || g->d.wrap_j, GCM_EUNSUP);  // latitude bands carry halos for the reference's stencil only
  if (x.coriolis) {
    GCM_REQUIRE(opt->h_cor_u && opt->h_cor_v, GCM_ENULL);
    const size_t n = (size_t)g->d.H;
    if (!g->d_cor) GCM_CUDA(cudaMalloc(&g->d_cor, 2 * n * sizeof(double)));
    GCM_CUDA(cudaMemcpy(g->d_cor, opt->h_cor_u, n * sizeof(double), cudaMemcpyHostToDevice));
    GCM_CUDA(cudaMemcpy((double*)g->d_cor + n, opt->h_cor_v, n * sizeof(double), cudaMemcpyHostToDevice));
    x.cor_u = (const double*)g->d_cor;
    x.cor_v = (const double*)g->d_cor + n;
  }
  g->x = x;
  ++g_gcm_tuning_epoch;  // cached step graphs were captured without / with other extras
  return GCM_OK;
}
