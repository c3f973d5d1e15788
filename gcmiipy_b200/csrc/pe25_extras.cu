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
// One thread per column (i, k) marches PX_RJ rows south: the tracer column j-2 ... j+2 and the flux through the north
// edge ride in registers, so a row costs one new load per tracer in j and three limiter evaluations instead of four.
// Divides are reciprocals (metric tables, gcm_rcp): fp64 division is a ~30-instruction software sequence.  The march is
// bound by load latency (ncu r03c: long-scoreboard 12.9 cycles per issue), so each row asks the next row's lines into L1.
// The default path (no option set) launches nothing from this file and stays bit-identical to the reference step.
// The limiter reads j - 2 ... j + 2: whole-grid geometries (rows periodic in j), or latitude bands that store at least
// two halo rows on either side (gcm_pe25_half_step only; the row-segment schedules keep the reference's halo widths).
#include <math.h>

#include "gcm_common.h"
#include "prof.h"


// Limited edge value minus the centred one at the edge between q0 and q1 (qm, q0 | q1, q2), upwind by the sign of the
// mass flux (flux_limiter.py:23-27).  With r = a / b the slope ratio of calc_r (:14-20; 0 where b == 0) seen from the
// upwind cell, van_leer(r) * b (:10) = 2 a b / (a + b) where a and b have the same sign and 0 elsewhere -- the
// harmonic-mean form of the same limiter: one reciprocal instead of two divides, and no r to overflow.
__device__ __forceinline__ double px_edge_excess(double qm, double q0, double q1, double q2, double flux) {
  const double b = q1 - q0;
  const double a = flux > 0 ? q0 - qm : q2 - q1;
  const double ab = a * b;
  const double lim = ab > 0 ? 2 * ab * gcm_rcp(a + b) : 0.0;
  return flux > 0 ? 0.5 * (lim - b) : 0.5 * (b - lim);   // (q0 + lim/2) - (q0+q1)/2   |   (q1 - lim/2) - (q0+q1)/2
}

// A correction flux is one rounded product: the flux through an edge is evaluated by the cell north of it (carried in
// a register) or at the start of a march, and both must give the same bits whatever the march length and the band
// decomposition -- no FMA contraction into the divergence that follows.
__device__ __forceinline__ double px_mul(double a, double b) { return __dmul_rn(a, b); }

// column i of one tracer while the thread marches south: the five rows j-2 ... j+2 and the correction flux through
// the north edge of row j (the south edge of the row before)
struct PxColumn {
  double m2, m1, c, p1, p2, gj_n;
};

// rows a thread marches over: 8 on large grids, fewer when the launch would not fill the chip
#define PX_RJ_MAX 8

__global__ void __launch_bounds__(128)
pe25x_extras_kernel(GcmGeomDev g, GcmExtras x, const double* __restrict__ sp, const double* __restrict__ su,
                    const double* __restrict__ sv, const double* __restrict__ st, const double* __restrict__ sq,
                    const double* __restrict__ spu, const double* __restrict__ pn, double* __restrict__ u,
                    double* __restrict__ v, double* __restrict__ t, double* __restrict__ q, double dt, int rj,
                    size_t b2, size_t b3) {
  gcm_pdl_wait();  // launched with the programmatic-stream-serialization attribute right behind the update kernel
  const int H = g.H, W = g.W, L = g.L, wrap = g.wrap_j;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const int j0 = g.row_lo + blockIdx.y * rj;
  const int j1 = j0 + rj < g.row_hi ? j0 + rj : g.row_hi;
  const int k = blockIdx.z % L, b = blockIdx.z / L;
  // member and layer folded into the pointers: every access below is pointer + 32-bit (row * W + column)
  const size_t o3 = b * b3 + (size_t)k * H * W;
  sp += b * b2; pn += b * b2;
  su += o3; sv += o3; st += o3; sq += o3; spu += o3;
  u += o3; v += o3; t += o3; q += o3;

  const int ip1 = gcm_ip(i, W), im1 = gcm_im(i, W), ip2 = gcm_ip(ip1, W), im2 = gcm_im(im1, W);
  const double rdy = g.rdy;
  const bool momentum = x.coriolis || x.nu != 0.0;
  const double* const trc_in[2] = {sq, st};
  double* const trc_out[2] = {q, t};
  const bool trc_on[2] = {x.limit_q != 0, x.limit_t != 0};

  // state carried from row to row (row offsets r* = row * W)
  int jm1 = gcm_row(j0, -1, H, wrap), jp1 = gcm_row(j0, 1, H, wrap);
  int rn = jm1 * W, rc = j0 * W, rs = jp1 * W;
  double sp_n = sp[rn + i], sp_c = sp[rc + i], sp_s = sp[rs + i];
  double spv_n = sv[rn + i] * ((sp_n + sp_c) / 2);  // star mass flux in j (dynamics.py:191) through the north edge
  PxColumn col[2];
#pragma unroll
  for (int f = 0; f < 2; ++f)
    if (trc_on[f]) {
      const double* __restrict__ a = trc_in[f];
      const int rn2 = gcm_row(jm1, -1, H, wrap) * W, rs2 = gcm_row(jp1, 1, H, wrap) * W;
      col[f].m2 = a[rn2 + i]; col[f].m1 = a[rn + i]; col[f].c = a[rc + i];
      col[f].p1 = a[rs + i]; col[f].p2 = a[rs2 + i];
      col[f].gj_n = px_mul(spv_n, px_edge_excess(col[f].m2, col[f].m1, col[f].c, col[f].p1, spv_n));
    }

  for (int j = j0; j < j1; ++j) {
    const int jp2 = gcm_row(jp1, 1, H, wrap);
    const int rs2 = jp2 * W;
    const int c = rc + i;
    if (j + 1 < j1) {  // the next row's lines into L1 while this row is computed (the march is load-latency bound)
      const int cn = rs + i;
      gcm_prefetch_l1(pn + cn);
      gcm_prefetch_l1(spu + cn);
      if (momentum) {
        gcm_prefetch_l1(u + cn);
        gcm_prefetch_l1(v + cn);
        gcm_prefetch_l1(su + rs2 + i);
        gcm_prefetch_l1(sv + rs2 + i);
      }
#pragma unroll
      for (int f = 0; f < 2; ++f)
        if (trc_on[f]) {
          gcm_prefetch_l1(trc_out[f] + cn);
          gcm_prefetch_l1(trc_in[f] + gcm_row(jp2, 1, H, wrap) * W + i);
        }
    }
    const double rdxj = g.rdx_j[j];
    const double pn_c = pn[c];
    const double sv_c = sv[c];
    const double spv_c = sv_c * ((sp_c + sp_s) / 2);

    if (momentum) {
      double fu = 0.0, fv = 0.0;  // what is added to dut + dus + pgfu and to dvt + dvs + phiv + pgv
      const double sp_e = sp[rc + ip1];
      if (x.coriolis) {
        const double sp_se = sp[rs + ip1], sp_ne = sp[rn + ip1];
        const double spv_e = sv[rc + ip1] * ((sp_e + sp_se) / 2);
        const double spv_ne = sv[rn + ip1] * ((sp_ne + sp_e) / 2);
        const double pv_at_pu = ((spv_c + spv_n) / 2 + (spv_e + spv_ne) / 2) / 2;  // iph(jmh(pv)), dynamics.py:87
        const double pu_at_pv = ((spu[c] + spu[rs + i]) / 2 + (spu[rc + im1] + spu[rs + im1]) / 2) / 2;  // imh(jph(pu)), :86
        fu += x.cor_u[j] * -pv_at_pu;  // :94
        fv += x.cor_v[j] * pu_at_pv;   // :95
      }
      const bool wall = (j == g.zero_v_row || j == g.zero_v_row2);  // v_n[:, -1, :] stays 0 (dynamics.py:222)
      if (x.nu != 0.0) {
        const double u0 = su[c];
        const double lap_u = (su[rc + ip1] + su[rc + im1] - 2 * u0) * (rdxj * rdxj) +
                             (su[rs + i] + su[rn + i] - 2 * u0) * (rdy * rdy);
        fu -= (sp_c + sp_e) / 2 * (x.nu * lap_u);
        if (!wall) {
          const double rdxh = g.rdx_h[j];
          const double lap_v = (sv[rc + ip1] + sv[rc + im1] - 2 * sv_c) * (rdxh * rdxh) +
                               (sv[rs + i] + sv[rn + i] - 2 * sv_c) * (rdy * rdy);
          fv -= (sp_c + sp_s) / 2 * (x.nu * lap_v);
        }
      }
      u[c] -= (fu * dt) * gcm_rcp((pn_c + pn[rc + ip1]) / 2);
      if (!wall) v[c] -= (fv * dt) * gcm_rcp((pn_c + pn[rs + i]) / 2);
    }

    if (trc_on[0] || trc_on[1]) {
      const double pu_c = spu[c], pu_w = spu[rc + im1];
      const double rpn = gcm_rcp(pn_c);
      const int rs3 = gcm_row(jp2, 1, H, wrap) * W;
#pragma unroll
      for (int f = 0; f < 2; ++f)
        if (trc_on[f]) {
          const double* __restrict__ a = trc_in[f];
          PxColumn& cl = col[f];
          const double fw1 = a[rc + im1], fw2 = a[rc + im2];
          const double fe1 = a[rc + ip1], fe2 = a[rc + ip2];
          const double gi_c = px_mul(pu_c, px_edge_excess(fw1, cl.c, fe1, fe2, pu_c));          // edge i + 1/2
          const double gi_w = px_mul(pu_w, px_edge_excess(fw2, fw1, cl.c, fe1, pu_w));          // edge i - 1/2
          const double gj_c = px_mul(spv_c, px_edge_excess(cl.m1, cl.c, cl.p1, cl.p2, spv_c));  // edge j + 1/2
          const double div = (gi_c - gi_w) * rdxj + (gj_c - cl.gj_n) * rdy;
          trc_out[f][c] -= (div * dt) * rpn;
          cl.m2 = cl.m1; cl.m1 = cl.c; cl.c = cl.p1; cl.p1 = cl.p2;
          if (j + 1 < j1) cl.p2 = a[rs3 + i];
          cl.gj_n = gj_c;
        }
    }
    spv_n = spv_c;
    sp_n = sp_c; sp_c = sp_s;
    if (j + 1 < j1) sp_s = sp[rs2 + i];
    rn = rc; rc = rs; rs = rs2;
    jp1 = jp2;
  }
}

// pn: p_n of the half step on the owned rows AND the first row to their south (the step's work field: `out->p` of a
// band does not hold the halo row that jph(p_n) reaches)
int gcm_pe25_extras_apply(const gcm_geom* g, const gcm_state* star, const gcm_state* out, const double* spu,
                          const double* pn, double dt, int nbatch, void* stream) {
  const GcmGeomDev& d = g->d;
  GCM_REQUIRE(gcm_extras_rows_ok(g), GCM_EUNSUP);
  GCM_REQUIRE((double)d.H * d.W < 2147483648.0, GCM_EUNSUP);  // 32-bit offsets within a layer
  const int H = d.H, W = d.W, L = d.L, nrows = d.row_hi - d.row_lo;
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  const unsigned gx = (unsigned)((W + tc - 1) / tc);
  int rj = PX_RJ_MAX;  // shorter marches until the launch has about four CTAs per SM
  while (rj > 1 && (size_t)gx * ((nrows + rj - 1) / rj) * L * nbatch < 592) rj /= 2;
  const unsigned gy = (unsigned)((nrows + rj - 1) / rj);
  {
    GcmProfScope ps(GCM_K_EXTRAS, stream);
    GCM_LAUNCH_DEP(pe25x_extras_kernel, dim3(gx, gy, L * nbatch), dim3(tc), 0, stream, d, g->x, star->p, star->u, star->v,
               star->t, star->q, spu, pn, out->u, out->v, out->t, out->q, dt, rj, (size_t)H * W,
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
  GCM_REQUIRE(!any || gcm_extras_rows_ok(g), GCM_EUNSUP);  // a band needs two halo rows on either side
  if (x.coriolis) {
    GCM_REQUIRE(opt->h_cor_u && opt->h_cor_v, GCM_ENULL);
    const size_t n = (size_t)g->d.H;
    if (!g->d_cor) GCM_CUDA(cudaMalloc(&g->d_cor, 2 * n * sizeof(double)));
#ifndef GCM_EMU
    // steps of this geometry may still be reading the table on non-blocking side streams, which a cudaMemcpy on the
    // legacy stream does not wait for: drain the device first (configure() is a rare, host-synchronous call)
    GCM_CUDA(cudaDeviceSynchronize());
#endif
    GCM_CUDA(cudaMemcpy(g->d_cor, opt->h_cor_u, n * sizeof(double), cudaMemcpyHostToDevice));
    GCM_CUDA(cudaMemcpy((double*)g->d_cor + n, opt->h_cor_v, n * sizeof(double), cudaMemcpyHostToDevice));
    x.cor_u = (const double*)g->d_cor;
    x.cor_v = (const double*)g->d_cor + n;
  }
  g->x = x;
  ++g_gcm_tuning_epoch;  // cached step graphs were captured without / with other extras
  return GCM_OK;
}
