// ops.cu -- stand-alone operators, diagnostics and halo-row movers behind include/gcm_b200.h.
// Compiled with -fmad=false and the reference's operation order (bit-identical to numpy except pow()).
#include <math.h>

#include "gcm_common.h"

#define IDX3(k, j, i) (((size_t)(k) * H + (size_t)(j)) * W + (size_t)(i))
#define IDX2(j, i) ((size_t)(j) * W + (size_t)(i))

static inline unsigned gcm_blocks(size_t n, int threads, unsigned cap) {
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  return (unsigned)(b > cap ? cap : b);
}

// ---------------------------------------------------------------------------------------------------
// phi_port.PGF (phi_port.py:5-113): Fortran-faithful hydrostatic column integration.  Only column i = 0 of
// every row is computed (IMAX = 1, :50-54); theta-bar is the arithmetic mean (:78); FDATA is the heightmap
// taken as it is (:101).  grid = (ceil(W/T), H)
// ---------------------------------------------------------------------------------------------------
__global__ void phi_port_kernel(GcmGeomDev g, const double* __restrict__ p, const double* __restrict__ t,
                                double* __restrict__ phi) {
  const int H = g.H, W = g.W, L = g.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= W) return;
  if (i != 0) {
    for (int k = 0; k < L; ++k) phi[IDX3(k, j, i)] = 0.0;
    return;
  }
  const double sha = GCM_RD / GCM_KAPPA;
  const double sp = p[IDX2(j, 0)];
  double sum1 = 0.0, sum2 = 0.0;
  double pdn = g.sig[0] * sp + g.ptop;
  double pkdn = pow(pdn, GCM_KAPPA);
  for (int k = 0; k < L - 1; ++k) {
    const double tk = t[IDX3(k, j, 0)];
    const double spa = g.sig[k] * sp * GCM_RD * tk * pkdn / pdn;
    sum1 = sum1 + spa * g.dsig[k];
    const double pup = g.sig[k + 1] * sp + g.ptop;
    const double pkup = pow(pup, GCM_KAPPA);
    const double theta = (t[IDX3(k + 1, j, 0)] + tk) / 2;
    const double ph = sha * theta * (pkdn - pkup);
    phi[IDX3(k + 1, j, 0)] = ph;
    sum2 = sum2 + g.sigt[k] * ph;  // SIGE[L+1] = sigt[L]
    pdn = pup;
    pkdn = pkup;
  }
  const double spa = g.sig[L - 1] * sp * GCM_RD * t[IDX3(L - 1, j, 0)] * pkdn / pdn;
  sum1 = sum1 + spa * g.dsig[L - 1];
  double run = g.hmap[IDX2(j, 0)] + sum1 - sum2;
  phi[IDX3(0, j, 0)] = run;
  for (int k = 1; k < L; ++k) {
    run = phi[IDX3(k, j, 0)] + run;
    phi[IDX3(k, j, 0)] = run;
  }
}

extern "C" int gcm_phi_port_pgf(const gcm_geom* g, const double* p, const double* t, double* phi, void* stream) {
  GCM_REQUIRE(g && p && t && phi, GCM_ENULL);
  const GcmGeomDev& d = g->d;
  const int tc = d.W >= 128 ? 128 : (d.W + 31) / 32 * 32;
  GCM_LAUNCH(phi_port_kernel, dim3((d.W + tc - 1) / tc, d.H, 1), dim3(tc), 0, stream, d, p, t, phi);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// flux_limiter.py:10-32 on nrows independent periodic rows of length n
// ---------------------------------------------------------------------------------------------------
__global__ void fl_van_leer_kernel(const double* __restrict__ r, double* __restrict__ out, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const double x = r[e], a = fabs(x);
    out[e] = (x + a) / (1 + a);
  }
}

// op 0 calc_r   1 donor_cell_flux   2 donor_cell_advection
__global__ void fl_row_kernel(int op, const double* __restrict__ q, const double* __restrict__ u, double* __restrict__ out,
                              int n, double dx, double dt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t r0 = (size_t)blockIdx.y * n;
  const int im = gcm_im(i, n), ip = gcm_ip(i, n);
  const double q_c = q[r0 + i], q_ip = q[r0 + ip], q_im = q[r0 + im];
  if (op == 0) {
    const double a = q_c - q_im, b = q_ip - q_c;
    out[r0 + i] = b != 0 ? a / b : 0.0;
  } else {
    const double u_c = u[r0 + i];
    const double flux = (u_c > 0 ? q_c : q_ip) * u_c;
    if (op == 1) {
      out[r0 + i] = flux;
    } else {
      const double u_im = u[r0 + im];
      const double flux_im = (u_im > 0 ? q_im : q_c) * u_im;
      out[r0 + i] = q_c + (flux_im - flux) * dt / dx;
    }
  }
}

extern "C" int gcm_fl_van_leer(const double* r, double* out, size_t n, void* stream) {
  GCM_REQUIRE(r && out, GCM_ENULL);
  if (n == 0) return GCM_OK;
  GCM_LAUNCH(fl_van_leer_kernel, dim3(gcm_blocks(n, 256, 148 * 8)), dim3(256), 0, stream, r, out, n);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

static int fl_rows(int op, const double* q, const double* u, double* out, int nrows, int n, double dx, double dt,
                   void* stream) {
  GCM_REQUIRE(nrows > 0 && n > 0, GCM_ESHAPE);
  const int tc = n >= 128 ? 128 : (n + 31) / 32 * 32;
  GCM_LAUNCH(fl_row_kernel, dim3((n + tc - 1) / tc, nrows, 1), dim3(tc), 0, stream, op, q, u, out, n, dx, dt);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}
extern "C" int gcm_fl_calc_r(const double* q, double* out, int nrows, int n, void* stream) {
  GCM_REQUIRE(q && out, GCM_ENULL);
  return fl_rows(0, q, nullptr, out, nrows, n, 1.0, 1.0, stream);
}
extern "C" int gcm_fl_donor_cell_flux(const double* q, const double* u, double* out, int nrows, int n, void* stream) {
  GCM_REQUIRE(q && u && out, GCM_ENULL);
  return fl_rows(1, q, u, out, nrows, n, 1.0, 1.0, stream);
}
extern "C" int gcm_fl_donor_cell_advection(const double* q, const double* u, double* out, int nrows, int n, double dx,
                                           double dt, int nsteps, double* tmp, void* stream) {
  GCM_REQUIRE(q && u && out, GCM_ENULL);
  GCM_REQUIRE(nsteps > 0, GCM_ESHAPE);
  GCM_REQUIRE(nsteps == 1 || tmp, GCM_ENULL);
  const double* cur = q;
  for (int s = 0; s < nsteps; ++s) {
    double* dst = ((nsteps - 1 - s) % 2 == 0) ? out : tmp;
    int st = fl_rows(2, cur, u, dst, nrows, n, dx, dt, stream);
    if (st) return st;
    cur = dst;
  }
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// coordinates*.py shift / half-average / gradient helpers (constants.py:85 unit_roll = np.roll)
// ---------------------------------------------------------------------------------------------------
__global__ void shift_op_kernel(int op, const double* __restrict__ q, double* __restrict__ out, int n2, int n1, int n0,
                                int axis, int shift, double d) {
  const size_t n = (size_t)n2 * n1 * n0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const int i0 = (int)(e % n0), i1 = (int)((e / n0) % n1), i2 = (int)(e / ((size_t)n0 * n1));
    const int s = op == 2 ? -1 : shift;  // np.roll(q, s)[x] = q[(x - s) mod n]
    int j0 = i0, j1 = i1, j2 = i2;
    if (axis == 0) { j0 = (i0 - s) % n0; if (j0 < 0) j0 += n0; }
    else if (axis == 1) { j1 = (i1 - s) % n1; if (j1 < 0) j1 += n1; }
    else { j2 = (i2 - s) % n2; if (j2 < 0) j2 += n2; }
    const double r = q[((size_t)j2 * n1 + j1) * n0 + j0], c = q[e];
    out[e] = op == 0 ? r : (op == 1 ? (c + r) / 2 : (r - c) / d);
  }
}

extern "C" int gcm_shift_op(int op, const double* q, double* out, int n2, int n1, int n0, int axis, int shift, double d,
                            void* stream) {
  GCM_REQUIRE(q && out, GCM_ENULL);
  GCM_REQUIRE(op >= 0 && op <= 2 && axis >= 0 && axis <= 2, GCM_EUNSUP);
  GCM_REQUIRE(n2 > 0 && n1 > 0 && n0 > 0, GCM_ESHAPE);
  const size_t n = (size_t)n2 * n1 * n0;
  GCM_LAUNCH(shift_op_kernel, dim3(gcm_blocks(n, 256, 148 * 8)), dim3(256), 0, stream, op, q, out, n2, n1, n0, axis,
             shift, d);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// temperature.py:7-19
// ---------------------------------------------------------------------------------------------------
__global__ void temperature_kernel(int dir, const double* __restrict__ t, const double* __restrict__ p,
                                   double* __restrict__ out, size_t n) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const double f = pow(GCM_P0 / p[e], GCM_KAPPA);
    out[e] = dir == 0 ? t[e] / f : t[e] * f;
  }
}

extern "C" int gcm_temperature_convert(int dir, const double* t, const double* p, double* out, size_t n, void* stream) {
  GCM_REQUIRE(t && p && out, GCM_ENULL);
  GCM_REQUIRE(dir == 0 || dir == 1, GCM_EUNSUP);
  if (n == 0) return GCM_OK;
  GCM_LAUNCH(temperature_kernel, dim3(gcm_blocks(n, 256, 148 * 8)), dim3(256), 0, stream, dir, t, p, out, n);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// diagnostics of no_limits_2_5d.full_timestep (no_limits_2_5d.py:85-91)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gcm_atomic_min(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}
__device__ __forceinline__ void gcm_atomic_max(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (seen == old) break;
    old = seen;
  }
}

__global__ void minmax_init_kernel(double* out) {
  out[0] = INFINITY;
  out[1] = -INFINITY;
  out[2] = 0.0;
}

// min / max propagate NaN like np.min / np.max; out[2] counts the non-finite values
__global__ void minmax_kernel(const double* __restrict__ x, size_t n, double* out) {
  __shared__ double smin[32], smax[32], scnt[32], snan[32];
  double mn = INFINITY, mx = -INFINITY, cnt = 0.0, isn = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
    const double v = x[e];
    if (v != v) { isn = 1.0; cnt += 1.0; continue; }
    if (isinf(v)) cnt += 1.0;
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    isn += __shfl_xor_sync(0xffffffffu, isn, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { smin[warp] = mn; smax[warp] = mx; scnt[warp] = cnt; snan[warp] = isn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) / 32;
    for (int w = 1; w < nw; ++w) {
      mn = smin[w] < mn ? smin[w] : mn;
      mx = smax[w] > mx ? smax[w] : mx;
      cnt += scnt[w];
      isn += snan[w];
    }
    if (isn > 0) {
      const unsigned long long qnan = 0x7ff8000000000000ull;
      atomicExch((unsigned long long*)&out[0], qnan);
      atomicExch((unsigned long long*)&out[1], qnan);
    } else {
      gcm_atomic_min(&out[0], mn);
      gcm_atomic_max(&out[1], mx);
    }
    if (cnt > 0) atomicAdd(&out[2], cnt);
  }
}

extern "C" int gcm_diag_minmax(const double* x, size_t n, double* out3, void* stream) {
  GCM_REQUIRE(x && out3, GCM_ENULL);
  GCM_REQUIRE(n > 0, GCM_ESHAPE);
  GCM_LAUNCH(minmax_init_kernel, dim3(1), dim3(1), 0, stream, out3);
  GCM_CHECK_LAUNCH();
  GCM_LAUNCH(minmax_kernel, dim3(gcm_blocks(n, 256, 148 * 4)), dim3(256), 0, stream, x, n, out3);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// no_limits_2_5d.calc_energy (no_limits_2_5d.py:35-60), one thread per column; block sums are combined
// with atomicAdd, so the three totals agree with numpy's pairwise np.sum to round-off, not bit for bit.
__global__ void energy_kernel(GcmGeomDev g, const double* __restrict__ p, const double* __restrict__ u,
                              const double* __restrict__ v, const double* __restrict__ t,
                              const double* __restrict__ area_by_i, double* out) {
  __shared__ double s[3][32];
  const int H = g.H, W = g.W, L = g.L;
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  double ke = 0.0, ate = 0.0, geo = 0.0;
  if (i < W) {
    const int im = gcm_im(i, W), jm = gcm_row(j, -1, H, 1);
    const double pc = p[IDX2(j, i)], area = area_by_i[i];
    double total_depth = 0.0;
    for (int k = 0; k < L; ++k) {
      const double a = (u[IDX3(k, j, i)] + u[IDX3(k, j, im)]) / 2, b = (v[IDX3(k, j, i)] + v[IDX3(k, jm, i)]) / 2;
      const double mag = sqrt(a * a + b * b);
      const double tp = pc * g.sig[k] + g.ptop;
      const double tt = t[IDX3(k, j, i)] / pow(GCM_P0 / tp, GCM_KAPPA);
      const double rho = tp / (GCM_RD * tt);
      const double depth = (pc * g.dsig[k]) / (rho * GCM_G);
      const double airmass = rho * depth * area;
      total_depth += depth;
      geo += total_depth * airmass * GCM_G;
      ke += mag * mag * .5 * airmass;
      ate += tt * GCM_CP * airmass;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ke += __shfl_xor_sync(0xffffffffu, ke, o);
    ate += __shfl_xor_sync(0xffffffffu, ate, o);
    geo += __shfl_xor_sync(0xffffffffu, geo, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s[0][warp] = ke; s[1][warp] = ate; s[2][warp] = geo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) / 32;
    for (int w = 1; w < nw; ++w) { ke += s[0][w]; ate += s[1][w]; geo += s[2][w]; }
    atomicAdd(&out[0], ke);
    atomicAdd(&out[1], ate);
    atomicAdd(&out[2], geo);
  }
}

extern "C" int gcm_pe25_energy(const gcm_geom* g, const gcm_state* s, const double* area_by_i, double* out3,
                               void* stream) {
  GCM_REQUIRE(g && s && s->p && s->u && s->v && s->t && area_by_i && out3, GCM_ENULL);
  GCM_REQUIRE(g->d.wrap_j, GCM_EUNSUP);
  const GcmGeomDev& d = g->d;
  GCM_CUDA(cudaMemsetAsync(out3, 0, 3 * sizeof(double), (cudaStream_t)stream));
  const int tc = d.W >= 128 ? 128 : (d.W + 31) / 32 * 32;
  GCM_LAUNCH(energy_kernel, dim3((d.W + tc - 1) / tc, d.H, 1), dim3(tc), 0, stream, d, (const double*)s->p,
             (const double*)s->u, (const double*)s->v, (const double*)s->t, area_by_i, out3);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

// ---------------------------------------------------------------------------------------------------
// latitude-band halo rows (np.roll over j across ranks, coordinates_3d.py:43-48)
// buffer layout: [p rows][u rows][v rows][t rows][q rows], each [layer][row][i]
// ---------------------------------------------------------------------------------------------------
struct GcmRows5 {
  const double* src[5];
  double* dst[5];
};

// mode 0: state rows -> buffer   1: buffer -> state rows   2: state rows -> state rows
__global__ void halo_rows_kernel(GcmRows5 a, int mode, int H, int W, int L, int src_row0, int dst_row0, int nrows) {
  const size_t per_p = (size_t)nrows * W, per_3 = (size_t)L * nrows * W;
  const size_t total = per_p + 4 * per_3;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int f;
    size_t r;
    if (e < per_p) { f = 0; r = e; } else { f = 1 + (int)((e - per_p) / per_3); r = (e - per_p) % per_3; }
    const int i = (int)(r % W), row = (int)((r / W) % nrows), k = (int)(r / ((size_t)W * nrows));
    const size_t in_state_src = ((size_t)k * H + src_row0 + row) * W + i;
    const size_t in_state_dst = ((size_t)k * H + dst_row0 + row) * W + i;
    const size_t in_buf = r;
    const size_t so = mode == 1 ? in_buf : in_state_src;
    const size_t dofs = mode == 0 ? in_buf : in_state_dst;
    a.dst[f][dofs] = a.src[f][so];
  }
}

extern "C" size_t gcm_halo_buffer_doubles(const gcm_geom* g, int nrows) {
  if (!g || nrows <= 0) return 0;
  return (size_t)nrows * g->d.W * (1 + 4 * (size_t)g->d.L);
}

static int halo_launch(const gcm_geom* g, GcmRows5 a, int mode, int src_row0, int dst_row0, int nrows, void* stream) {
  const GcmGeomDev& d = g->d;
  GCM_REQUIRE(nrows > 0 && src_row0 >= 0 && dst_row0 >= 0 && src_row0 + nrows <= d.H && dst_row0 + nrows <= d.H,
              GCM_ESHAPE);
  const size_t total = gcm_halo_buffer_doubles(g, nrows);
  GCM_LAUNCH(halo_rows_kernel, dim3(gcm_blocks(total, 256, 148 * 4)), dim3(256), 0, stream, a, mode, d.H, d.W, d.L,
             src_row0, dst_row0, nrows);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

static void halo_buf_ptrs(const gcm_geom* g, int nrows, double* buf, double* out[5]) {
  const size_t per_p = (size_t)nrows * g->d.W, per_3 = per_p * g->d.L;
  out[0] = buf;
  for (int f = 1; f < 5; ++f) out[f] = buf + per_p + (size_t)(f - 1) * per_3;
}

extern "C" int gcm_halo_pack(const gcm_geom* g, const gcm_state* s, int row0, int nrows, double* buf, void* stream) {
  GCM_REQUIRE(g && s && s->p && s->u && s->v && s->t && s->q && buf, GCM_ENULL);
  GcmRows5 a;
  double* b[5];
  halo_buf_ptrs(g, nrows > 0 ? nrows : 1, buf, b);
  const double* src[5] = {s->p, s->u, s->v, s->t, s->q};
  for (int f = 0; f < 5; ++f) { a.src[f] = src[f]; a.dst[f] = b[f]; }
  return halo_launch(g, a, 0, row0, 0, nrows, stream);
}

extern "C" int gcm_halo_unpack(const gcm_geom* g, const gcm_state* s, int row0, int nrows, const double* buf,
                               void* stream) {
  GCM_REQUIRE(g && s && s->p && s->u && s->v && s->t && s->q && buf, GCM_ENULL);
  GcmRows5 a;
  double* b[5];
  halo_buf_ptrs(g, nrows > 0 ? nrows : 1, (double*)buf, b);
  double* dst[5] = {s->p, s->u, s->v, s->t, s->q};
  for (int f = 0; f < 5; ++f) { a.src[f] = b[f]; a.dst[f] = dst[f]; }
  return halo_launch(g, a, 1, 0, row0, nrows, stream);
}

extern "C" int gcm_halo_copy_rows(const gcm_geom* g, const gcm_state* src, int src_row0, const gcm_state* dst,
                                  int dst_row0, int nrows, void* stream) {
  GCM_REQUIRE(g && src && dst && src->p && src->u && src->v && src->t && src->q && dst->p && dst->u && dst->v &&
                  dst->t && dst->q, GCM_ENULL);
  GcmRows5 a;
  const double* s[5] = {src->p, src->u, src->v, src->t, src->q};
  double* d[5] = {dst->p, dst->u, dst->v, dst->t, dst->q};
  for (int f = 0; f < 5; ++f) { a.src[f] = s[f]; a.dst[f] = d[f]; }
  return halo_launch(g, a, 2, src_row0, dst_row0, nrows, stream);
}
