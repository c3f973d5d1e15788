// pe2d.cu -- 2-D primitive equations (p, u, v, theta; q passive) Matsuno step on a uniform doubly
// periodic C-grid (reference no_limits_2d.py:21-131).  Compiled with -fmad=false, reference operation
// order: everything except the pow() inside the pressure-gradient term is bit-identical to numpy.
#include "gcm_common.h"
#include "stencil2d.h"

// The momentum advection reaches column i+2 (puvp = ipj(puvm) reads vph and p_mid at i+1, hence v, p at
// i+2), so the half-step kernel addresses columns explicitly instead of through GcmNb.
struct Pe2dIdx {
  int r[4];  // rows j-1 .. j+2 (offsets, times W)
  int c[4];  // columns i-1 .. i+2
};
#define A2(a, dj, di) (a[ix.r[(dj) + 1] + ix.c[(di) + 1]])

__device__ __forceinline__ Pe2dIdx pe2d_idx(int j, int i, int H, int W) {
  Pe2dIdx ix;
  for (int d = 0; d < 4; ++d) {
    ix.r[d] = gcm_wrap(j - 1 + d, H) * W;
    ix.c[d] = gcm_wrap(i - 1 + d, W);
  }
  return ix;
}

// puvm at the cell (0, di): jmh(u) * ijm(iph(v)) * ijm(iph(jph(p)))   (no_limits_2d.py:53-58)
__device__ __forceinline__ double pe2d_puvm(const double* p, const double* u, const double* v, const Pe2dIdx& ix, int di) {
  const double jmh_u = (A2(u, 0, di) + A2(u, -1, di)) / 2;
  const double vph = (A2(v, -1, di) + A2(v, -1, di + 1)) / 2;
  const double jph_a = (A2(p, -1, di) + A2(p, 0, di)) / 2;
  const double jph_b = (A2(p, -1, di + 1) + A2(p, 0, di + 1)) / 2;
  return jmh_u * vph * ((jph_a + jph_b) / 2);
}
// pvum at the cell (0, di): imj(p_mid) * imh(v) * imj(jph(u))   (no_limits_2d.py:66-68)
__device__ __forceinline__ double pe2d_pvum(const double* p, const double* u, const double* v, const Pe2dIdx& ix, int di) {
  const double jph_a = (A2(p, 0, di - 1) + A2(p, 1, di - 1)) / 2;
  const double jph_b = (A2(p, 0, di) + A2(p, 1, di)) / 2;
  const double pmid_im = (jph_a + jph_b) / 2;
  const double imh_v = (A2(v, 0, di) + A2(v, 0, di - 1)) / 2;
  const double jph_u_im = (A2(u, 0, di - 1) + A2(u, 1, di - 1)) / 2;
  return pmid_im * imh_v * jph_u_im;
}

__device__ __forceinline__ void pe2d_advec_m_full(const double* p, const double* u, const double* v, const Pe2dIdx& ix,
                                                  double dx, double* dut, double* dvt) {
  const double a = (A2(u, 0, 0) + A2(u, 0, -1)) / 2;
  const double puum = a * a * A2(p, 0, 0);
  const double b = (A2(u, 0, 1) + A2(u, 0, 0)) / 2;
  const double puup = b * b * A2(p, 0, 1);
  const double puvm = pe2d_puvm(p, u, v, ix, 0);
  const double puvp = pe2d_puvm(p, u, v, ix, 1);
  *dut = (puum - puup) / dx + (puvm - puvp) / dx;
  const double c = (A2(v, 0, 0) + A2(v, -1, 0)) / 2;
  const double pvvm = c * c * A2(p, 0, 0);
  const double d = (A2(v, 1, 0) + A2(v, 0, 0)) / 2;
  const double pvvp = d * d * A2(p, 1, 0);
  const double pvum = pe2d_pvum(p, u, v, ix, 0);
  const double pvup = pe2d_pvum(p, u, v, ix, 1);
  *dvt = (pvvm - pvvp) / dx + (pvum - pvup) / dx;
}

// pgf (no_limits_2d.py:76-89)
__device__ __forceinline__ void pe2d_pgf(const double* p, const double* t, const Pe2dIdx& ix, double dx, double* pgfu,
                                         double* pgfv) {
  const double p_c = A2(p, 0, 0), p_ip = A2(p, 0, 1), p_jp = A2(p, 1, 0);
  const double t_c = A2(t, 0, 0);
  const double ppih = (p_c + p_ip) / 2;
  const double ttu = ((t_c + A2(t, 0, 1)) / 2) / pow(GCM_P0 / ppih, GCM_KAPPA);
  const double rhou = ppih / (GCM_RD * ttu);
  *pgfu = ppih / rhou * ((p_ip - p_c) / dx);
  const double ppjh = (p_c + p_jp) / 2;
  const double ttv = ((t_c + A2(t, 1, 0)) / 2) / pow(GCM_P0 / ppjh, GCM_KAPPA);
  const double rhov = ppjh / (GCM_RD * ttv);
  *pgfv = ppjh / rhov * ((p_jp - p_c) / dx);
}

// p_n at the cell (dj, di) in {(0,0), (0,1), (1,0)}: p - advec_p(spu, spv) dt   (no_limits_2d.py:41-44, :112)
__device__ __forceinline__ double pe2d_pn(const double* p, const double* sp, const double* su, const double* sv,
                                          const Pe2dIdx& ix, int dj, int di, double dx, double dt) {
  const double sp_c = A2(sp, dj, di);
  const double spu = A2(su, dj, di) * ((sp_c + A2(sp, dj, di + 1)) / 2);
  const double spu_im = A2(su, dj, di - 1) * ((A2(sp, dj, di - 1) + sp_c) / 2);
  const double spv = A2(sv, dj, di) * ((sp_c + A2(sp, dj + 1, di)) / 2);
  const double spv_jm = A2(sv, dj - 1, di) * ((A2(sp, dj - 1, di) + sp_c) / 2);
  return A2(p, dj, di) - ((spu - spu_im) / dx + (spv - spv_jm) / dx) * dt;
}

__global__ void pe2d_half_kernel(const double* __restrict__ p, const double* __restrict__ u, const double* __restrict__ v,
                                 const double* __restrict__ t, const double* __restrict__ q,
                                 const double* __restrict__ sp, const double* __restrict__ su,
                                 const double* __restrict__ sv, const double* __restrict__ st, double* __restrict__ po,
                                 double* __restrict__ uo, double* __restrict__ vo, double* __restrict__ to,
                                 double* __restrict__ qo, int H, int W, double dt, double dx,
                                 unsigned int* nonfinite) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  const Pe2dIdx ix = pe2d_idx(j, i, H, W);
  const size_t c = (size_t)j * W + i;
  const double pn_c = pe2d_pn(p, sp, su, sv, ix, 0, 0, dx, dt);
  const double pn_ip = pe2d_pn(p, sp, su, sv, ix, 0, 1, dx, dt);
  const double pn_jp = pe2d_pn(p, sp, su, sv, ix, 1, 0, dx, dt);
  double dut, dvt, pgu, pgv;
  pe2d_advec_m_full(sp, su, sv, ix, dx, &dut, &dvt);
  pe2d_pgf(sp, st, ix, dx, &pgu, &pgv);
  const double p_c = A2(p, 0, 0);
  const double pu = u[c] * ((p_c + A2(p, 0, 1)) / 2);
  const double pv = v[c] * ((p_c + A2(p, 1, 0)) / 2);
  const double pu_n = pu - (dut + pgu) * dt;
  const double pv_n = pv - (dvt + pgv) * dt;
  const double u_n = pu_n / ((pn_c + pn_ip) / 2), v_n = pv_n / ((pn_c + pn_jp) / 2);
  uo[c] = u_n;
  vo[c] = v_n;
  // advec_t (no_limits_2d.py:92-101)
  const double sp_c = A2(sp, 0, 0), st_c = A2(st, 0, 0);
  const double spu = A2(su, 0, 0) * ((sp_c + A2(sp, 0, 1)) / 2);
  const double spu_im = A2(su, 0, -1) * ((A2(sp, 0, -1) + sp_c) / 2);
  const double spv = A2(sv, 0, 0) * ((sp_c + A2(sp, 1, 0)) / 2);
  const double spv_jm = A2(sv, -1, 0) * ((A2(sp, -1, 0) + sp_c) / 2);
  const double tpu = spu * ((st_c + A2(st, 0, 1)) / 2);
  const double tpu_im = spu_im * ((A2(st, 0, -1) + st_c) / 2);
  const double tpv = spv * ((st_c + A2(st, 1, 0)) / 2);
  const double tpv_jm = spv_jm * ((A2(st, -1, 0) + st_c) / 2);
  const double adv = (tpu - tpu_im) / dx + (tpv - tpv_jm) / dx;
  const double t_n = t[c] - (adv / pn_c) * dt;
  to[c] = t_n;
  po[c] = pn_c;
  qo[c] = q[c];  // q is passed through (no_limits_2d.py:126)
  gcm_flag_nonfinite(nonfinite, gcm_not_finite((u_n + v_n) + (t_n + pn_c)));
}

static int pe2d_check(const gcm_state* s) {
  GCM_REQUIRE(s && s->p && s->u && s->v && s->t && s->q, GCM_ENULL);
  return GCM_OK;
}

static int pe2d_half_impl(const gcm_state* b, const gcm_state* s, const gcm_state* o, int H, int W, double dt, double dx,
                          void* stream) {
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  GCM_LAUNCH(pe2d_half_kernel, dim3((W + tc - 1) / tc, H, 1), dim3(tc), 0, stream, (const double*)b->p,
             (const double*)b->u, (const double*)b->v, (const double*)b->t, (const double*)b->q, (const double*)s->p,
             (const double*)s->u, (const double*)s->v, (const double*)s->t, o->p, o->u, o->v, o->t, o->q, H, W, dt, dx,
             gcm_nonfinite_word());
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

extern "C" int gcm_pe2d_half_step(const gcm_state* base, const gcm_state* star, const gcm_state* out, int H, int W,
                                  double dt, double dx, void* stream) {
  int st;
  if ((st = pe2d_check(base)) || (st = pe2d_check(star)) || (st = pe2d_check(out))) return st;
  GCM_REQUIRE(H > 0 && W > 0, GCM_ESHAPE);
  return pe2d_half_impl(base, star, out, H, W, dt, dx, stream);
}

extern "C" size_t gcm_pe2d_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  return 10 * n2 * sizeof(double);  // star state + one ping-pong state
}

extern "C" int gcm_pe2d_matsuno_step(const gcm_state* in, const gcm_state* out, int H, int W, double dt, double dx,
                                     int nsteps, void* ws, size_t ws_bytes, void* stream) {
  int st;
  if ((st = pe2d_check(in)) || (st = pe2d_check(out))) return st;
  GCM_REQUIRE(ws, GCM_ENULL);
  GCM_REQUIRE(H > 0 && W > 0 && nsteps > 0, GCM_ESHAPE);
  GCM_REQUIRE(ws_bytes >= gcm_pe2d_workspace_bytes(H, W), GCM_EWORK);
  const size_t n2 = ((size_t)H * W + 1) / 2 * 2;
  double* w = (double*)ws;
  gcm_state star = {w, w + n2, w + 2 * n2, w + 3 * n2, w + 4 * n2};
  gcm_state tmp = {w + 5 * n2, w + 6 * n2, w + 7 * n2, w + 8 * n2, w + 9 * n2};
  const gcm_state* cur = in;
  for (int s = 0; s < nsteps; ++s) {
    const gcm_state* dst = ((nsteps - 1 - s) % 2 == 0) ? out : &tmp;
    if ((st = pe2d_half_impl(cur, cur, &star, H, W, dt, dx, stream))) return st;   // no_limits_2d.py:130
    if ((st = pe2d_half_impl(cur, &star, dst, H, W, dt, dx, stream))) return st;   // no_limits_2d.py:131
    cur = dst;
  }
  return GCM_OK;
}

// op 0: advec_m -> (dut, dvt)   op 1: pgf -> (pgfu, pgfv)
__global__ void pe2d_operator_kernel(int op, const double* __restrict__ p, const double* __restrict__ u,
                                     const double* __restrict__ v, const double* __restrict__ t, double* __restrict__ o0,
                                     double* __restrict__ o1, int H, int W, double dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  const Pe2dIdx ix = pe2d_idx(j, i, H, W);
  double a, b;
  if (op == 0) pe2d_advec_m_full(p, u, v, ix, dx, &a, &b);
  else pe2d_pgf(p, t, ix, dx, &a, &b);
  o0[(size_t)j * W + i] = a;
  o1[(size_t)j * W + i] = b;
}

extern "C" int gcm_pe2d_operator(int op, const double* p, const double* u, const double* v, const double* t, double* o0,
                                 double* o1, int H, int W, double dx, void* stream) {
  GCM_REQUIRE(p && o0 && o1, GCM_ENULL);
  GCM_REQUIRE(op == 0 || op == 1, GCM_EUNSUP);
  if (op == 0) GCM_REQUIRE(u && v, GCM_ENULL);
  if (op == 1) GCM_REQUIRE(t, GCM_ENULL);
  GCM_REQUIRE(H > 0 && W > 0, GCM_ESHAPE);
  const int tc = W >= 128 ? 128 : (W + 31) / 32 * 32;
  GCM_LAUNCH(pe2d_operator_kernel, dim3((W + tc - 1) / tc, H, 1), dim3(tc), 0, stream, op, p, u, v, t, o0, o1, H, W, dx);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}
