// pe25_fast.cu -- ALU-lean, fused kernels of the 2.5-D half step (reference dynamics.py:183-227).
//
// The step is bound by the FP64 pipe and by HBM about equally, so this path cuts both:
//   * 2 launches per half step instead of 4:
//       R  pe25f_row_kernel     one CTA per latitude row, all layers: spu = filter(su iph(sp)) with an in-place
//                               shared-memory FFT (fft_inplace.h), then the column work of that row straight
//                               from shared memory (conv, pit, sd, p_n; hydrostatic phi, rho), then
//                               pgf = filter(pgfu + phiu) through the same buffer      (dynamics.py:186-202)
//       U  pe25f_update_kernel  one thread per column, k loop with the vertical neighbours and interface
//                               fluxes carried in registers: momentum, tracer update     (dynamics.py:197-222)
//   * divides by metric terms become multiplications by resident reciprocals (1/dx_j, 1/dx_h, 1/dy, 1/dsig),
//     the three divides by p_n averages are done once per column instead of once per cell, and when ptop = 0
//     (the reference's setting, geometry.py:147) the Exner factor ((sig p + ptop)/P0)^kappa factorises into
//     sig^kappa (resident table) x (p/P0)^kappa: one pow per column instead of one per cell.
// FMA contraction is on for this file.  Results agree with the reference within the stated fp64 tolerance
// (tests/test_parity.py); the bit-exact operator kernels stay in pe25.cu.
#include "fft_inplace.h"
#include "gcm_common.h"
#include "prof.h"

#define IDX3(k, j, i) (((size_t)(k) * H + (size_t)(j)) * W + (size_t)(i))
#define IDX2(j, i) ((size_t)(j) * W + (size_t)(i))

struct PfConst {
  const double *p, *u, *v, *t, *q;
};
struct PfMut {
  double *p, *u, *v, *t, *q;
};
struct PfWork {
  double *spu, *sd, *phi, *rho, *pgf, *pn;
};

// ---------------------------------------------------------------------------------------------------
// R: one CTA per (row, member)
// ---------------------------------------------------------------------------------------------------
template <int L, bool PTOP0>
__global__ void __launch_bounds__(512, 1)
pe25f_row_kernel(GcmGeomDev g, const double* __restrict__ p, PfConst star, PfWork w, double dt, int ja, size_t bstride2,
                 size_t bstride3) {
  GCM_DYN_SMEM(double2, z);
  constexpr int NP = (L + 1) / 2;
  const int H = g.H, W = g.W;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int j = ja + blockIdx.x;
  const size_t o2 = blockIdx.y * bstride2, o3 = blockIdx.y * bstride3;
  const double* sp = star.p + o2;
  const double* su = star.u + o3;
  const double* sv = star.v + o3;
  const double* st = star.t + o3;
  p += o2;
  double* spu = w.spu + o3;
  double* sd = w.sd + o3;
  double* phi = w.phi + o3;
  double* rho = w.rho + o3;
  double* pgf = w.pgf + o3;
  double* pn = w.pn + o2;
  const int jm = gcm_row(j, -1, H, g.wrap_j), jp = gcm_row(j, 1, H, g.wrap_j);
  const double* sprow = sp + IDX2(j, 0);
  const double* table = g.smmz + (size_t)j * (W / 2 + 1);
  const double inv = W == 1 ? 1.0 : 1.0 / W;
  const double rdxj = g.rdx_j[j], rdy = g.rdy;

  // 1. spu_orig = su * iph(sp), two layers per complex row  (dynamics.py:187)
  for (int e = tid; e < NP * W; e += nthr) {
    const int pr = e / W, i = e - pr * W;
    const int k0 = 2 * pr, k1 = k0 + 1;
    const double ph = (sprow[i] + sprow[gcm_ip(i, W)]) * 0.5;
    const double x0 = su[IDX3(k0, j, i)] * ph;
    const double x1 = k1 < L ? su[IDX3(k1, j, i)] * ph : 0.0;
    z[e] = make_double2(x0, x1);
  }
  __syncthreads();
  gcm_filter_rows_inplace(z, NP, g.plan, g.tw, g.kperm, table, tid, nthr);  // dynamics.py:189
  for (int e = tid; e < NP * W; e += nthr) {
    const int pr = e / W, i = e - pr * W;
    const int k0 = 2 * pr, k1 = k0 + 1;
    double2 v = z[e];
    v.x *= inv;
    v.y *= inv;
    z[e] = v;
    spu[IDX3(k0, j, i)] = v.x;
    if (k1 < L) spu[IDX3(k1, j, i)] = v.y;
  }
  __syncthreads();

  // 2. column work of this row: aflux (dynamics.py:35-46), p_n (:194), geopotential and rho (:111-142, :150-152)
  for (int i = tid; i < W; i += nthr) {
    const int im = gcm_im(i, W);
    const double sp_c = sprow[i];
    const double pjh = (sp_c + sp[IDX2(jp, i)]) * 0.5, pjh_m = (sp[IDX2(jm, i)] + sp_c) * 0.5;
    double conv[L];
    double pit = 0.0;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      const double2 a = z[(k >> 1) * W + i], b = z[(k >> 1) * W + im];
      const double pu_c = (k & 1) ? a.y : a.x, pu_im = (k & 1) ? b.y : b.x;
      const double pv_c = sv[IDX3(k, j, i)] * pjh, pv_jm = sv[IDX3(k, jm, i)] * pjh_m;
      conv[k] = ((pu_c - pu_im) * rdxj + (pv_c - pv_jm) * rdy) * g.dsig[k];
      pit += conv[k];
    }
    double acc = 0.0;
#pragma unroll
    for (int k = L - 1; k >= 0; --k) {
      acc += conv[k];
      sd[IDX3(k, j, i)] = k == 0 ? 0.0 : acc - pit * g.sigb[k];  // dynamics.py:42-44
    }
    pn[IDX2(j, i)] = p[IDX2(j, i)] - pit * dt;

    // hydrostatic geopotential at layer centres
    const double ptop = g.ptop;
    double pk_col = 0.0;
    if (PTOP0) pk_col = pow(sp_c / GCM_P0, GCM_KAPPA);
    double tk = st[IDX3(0, j, i)];
    double pk = PTOP0 ? g.sigkap[0] * pk_col : pow((g.sig[0] * sp_c + ptop) / GCM_P0, GCM_KAPPA);
    const double t0 = tk, pk0 = pk;
    double stp[L];
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < L; ++k) {
      double t_n = t0, pk_n = pk0;  // k + 1 wraps to layer 0 (coordinates_3d.py:55); sigt[L-1] = 0 kills it
      if (k + 1 < L) {
        t_n = st[IDX3(k + 1, j, i)];
        pk_n = PTOP0 ? g.sigkap[k + 1] * pk_col : pow((g.sig[k + 1] * sp_c + ptop) / GCM_P0, GCM_KAPPA);
      }
      const double tp = sp_c * g.sig[k] + ptop;
      const double rtt = GCM_RD * (tk * pk);  // Rd * T
      const double r = tp / rtt;              // rho (dynamics.py:152)
      const double spa = PTOP0 ? rtt : (g.sig[k] * sp_c) / r;  // sig p / rho
      stp[k] = GCM_CP * ((tk + t_n) * 0.5) * (pk - pk_n);
      sum += spa * g.dsig[k] - g.sigt[k] * stp[k];
      rho[IDX3(k, j, i)] = r;
      tk = t_n;
      pk = pk_n;
    }
    double run = sum + g.hmap[IDX2(j, i)] * GCM_G;
    phi[IDX3(0, j, i)] = run;
#pragma unroll
    for (int k = 1; k < L; ++k) {
      run += stp[k - 1];
      phi[IDX3(k, j, i)] = run;
    }
  }
  __syncthreads();  // z is free again; phi and rho of this row are visible to the block

  // 3. pgfu + phiu (dynamics.py:159, :162-165) into the FFT buffer, filter (:202), write
  for (int e = tid; e < NP * W; e += nthr) {
    const int pr = e / W, i = e - pr * W;
    const int ip = gcm_ip(i, W);
    const int k0 = 2 * pr, k1 = k0 + 1;
    const double p_c = sprow[i], p_ip = sprow[ip];
    const double psum = p_c + p_ip, gradp = (p_ip - p_c) * rdxj;
    const double a_u = psum * gradp;       // (p_c + p_ip) dp/dx
    const double b_u = psum * 0.5 * rdxj;  // iph(p) / dx
    double x[2] = {0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int k = s ? k1 : k0;
      if (k < L) {
        const size_t c = IDX3(k, j, i), cp = IDX3(k, j, ip);
        const double pgu = g.sig[k] * a_u / (rho[c] + rho[cp]);
        x[s] = pgu + b_u * (phi[cp] - phi[c]);
      }
    }
    z[e] = make_double2(x[0], x[1]);
  }
  __syncthreads();
  gcm_filter_rows_inplace(z, NP, g.plan, g.tw, g.kperm, table, tid, nthr);
  for (int e = tid; e < NP * W; e += nthr) {
    const int pr = e / W, i = e - pr * W;
    const int k0 = 2 * pr, k1 = k0 + 1;
    const double2 v = z[e];
    pgf[IDX3(k0, j, i)] = v.x * inv;
    if (k1 < L) pgf[IDX3(k1, j, i)] = v.y * inv;
  }
}

// ---------------------------------------------------------------------------------------------------
// U: one thread per column, k loop
// ---------------------------------------------------------------------------------------------------
template <int L, int MINB>
__global__ void __launch_bounds__(160, MINB)
pe25f_update_kernel(GcmGeomDev g, PfConst base, PfConst star, PfMut out, PfWork w, double dt, int ja, size_t bstride2,
                    size_t bstride3) {
  const int H = g.H, W = g.W;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const int j = ja + blockIdx.y;
  const size_t o2 = blockIdx.z * bstride2, o3 = blockIdx.z * bstride3;
  const double* p = base.p + o2;
  const double* u = base.u + o3;
  const double* v = base.v + o3;
  const double* t = base.t + o3;
  const double* q = base.q + o3;
  const double* sp = star.p + o2;
  const double* su = star.u + o3;
  const double* sv = star.v + o3;
  const double* st = star.t + o3;
  const double* sq = star.q + o3;
  const double* spu = w.spu + o3;
  const double* sd = w.sd + o3;
  const double* phi = w.phi + o3;
  const double* rho = w.rho + o3;
  const double* pgf = w.pgf + o3;
  const double* pn = w.pn + o2;

  const int wrap = g.wrap_j;
  const int jm = gcm_row(j, -1, H, wrap), jp = gcm_row(j, 1, H, wrap), jpp = gcm_row(jp, 1, H, wrap);
  const int im = gcm_im(i, W), ip = gcm_ip(i, W);
  const double rdxj = g.rdx_j[j], rdxh = g.rdx_h[j], rdy = g.rdy;

  // per-column (2-D) factors
  const double p_c = p[IDX2(j, i)], p_ip = p[IDX2(j, ip)], p_jp = p[IDX2(jp, i)];
  const double pn_c = pn[IDX2(j, i)], pn_ip = pn[IDX2(j, ip)], pn_jp = pn[IDX2(jp, i)];
  const double pu_fac = (p_c + p_ip) * 0.5, pv_fac = (p_c + p_jp) * 0.5;                 // calc_pu / calc_pv
  const double r_pnu = 1.0 / ((pn_c + pn_ip) * 0.5), r_pnv = 1.0 / ((pn_c + pn_jp) * 0.5);  // un_pu / un_pv
  const double r_pn = 1.0 / pn_c;
  const double sp_c = sp[IDX2(j, i)], sp_ip = sp[IDX2(j, ip)], sp_jp = sp[IDX2(jp, i)], sp_jm = sp[IDX2(jm, i)];
  const double a_c = (sp_c + sp_jp) * 0.5;                                  // jph(sp) at (j, i)
  const double a_ip = (sp_ip + sp[IDX2(jp, ip)]) * 0.5;                    // (j, i+1)
  const double a_jm = (sp_jm + sp_c) * 0.5;                                // (j-1, i)
  const double a_jm_ip = (sp[IDX2(jm, ip)] + sp_ip) * 0.5;                 // (j-1, i+1)
  const double a_jp = (sp_jp + sp[IDX2(jpp, i)]) * 0.5;                    // (j+1, i)
  const double a_v = (sp_c + sp_jp) * ((sp_jp - sp_c) * rdy);              // (p_c + p_jp) dp/dy
  const double b_v = a_c * rdy;                                            // jph(p) / dy
  const bool zero_v = j == g.zero_v_row;

  // vertical neighbours and interface fluxes carried in registers (advec_sig, dynamics.py:49-52)
  double u_k = su[IDX3(0, j, i)], v_k = sv[IDX3(0, j, i)], t_k = st[IDX3(0, j, i)], q_k = sq[IDX3(0, j, i)];
  double sd_c = sd[IDX3(0, j, i)], sd_ip = sd[IDX3(0, j, ip)], sd_jp = sd[IDX3(0, jp, i)];
  // flux through the bottom of layer 0 pairs layer 0 with layer L-1 (np.roll) times sd[0] = 0
  double fu, fv, ft, fq;
  {
    const size_t c = IDX3(L - 1, j, i);
    fu = (u_k + su[c]) * 0.5 * ((sd_c + sd_ip) * 0.5);
    fv = (v_k + sv[c]) * 0.5 * ((sd_c + sd_jp) * 0.5);
    ft = (t_k + st[c]) * 0.5 * sd_c;
    fq = (q_k + sq[c]) * 0.5 * sd_c;
  }
  const double fu0 = fu, fv0 = fv, ft0 = ft, fq0 = fq;

#pragma unroll
  for (int k = 0; k < L; ++k) {
    const size_t c = IDX3(k, j, i), c_im = IDX3(k, j, im), c_ip = IDX3(k, j, ip), c_jp = IDX3(k, jp, i),
                 c_jm = IDX3(k, jm, i);
    // fluxes through the top of layer k
    double fu_n = fu0, fv_n = fv0, ft_n = ft0, fq_n = fq0;
    double u_kp = 0.0, v_kp = 0.0, t_kp = 0.0, q_kp = 0.0;
    if (k + 1 < L) {
      const size_t cn = IDX3(k + 1, j, i);
      u_kp = su[cn]; v_kp = sv[cn]; t_kp = st[cn]; q_kp = sq[cn];
      sd_c = sd[cn]; sd_ip = sd[IDX3(k + 1, j, ip)]; sd_jp = sd[IDX3(k + 1, jp, i)];
      fu_n = (u_kp + u_k) * 0.5 * ((sd_c + sd_ip) * 0.5);
      fv_n = (v_kp + v_k) * 0.5 * ((sd_c + sd_jp) * 0.5);
      ft_n = (t_kp + t_k) * 0.5 * sd_c;
      fq_n = (q_kp + q_k) * 0.5 * sd_c;
    }
    const double rds = g.rdsig[k];
    const double dus = -((fu - fu_n) * rds), dvs = -((fv - fv_n) * rds);
    const double ads_t = -((ft - ft_n) * rds), ads_q = -((fq - fq_n) * rds);

    // horizontal neighbours
    const double u_im = su[c_im], u_ip = su[c_ip], u_jp = su[c_jp], u_jm = su[c_jm];
    const double v_im = sv[c_im], v_ip = sv[c_ip], v_jp = sv[c_jp], v_jm = sv[c_jm], v_jm_ip = sv[IDX3(k, jm, ip)];
    const double pu_c = spu[c], pu_im = spu[c_im], pu_ip = spu[c_ip], pu_jp = spu[c_jp], pu_jp_im = spu[IDX3(k, jp, im)];
    const double pv_c = v_k * a_c, pv_ip = v_ip * a_ip, pv_jm = v_jm * a_jm, pv_jm_ip = v_jm_ip * a_jm_ip,
                 pv_jp = v_jp * a_jp;

    // advec_m_pu (dynamics.py:55-108); (a/2)(b/2) = ab/4 exactly
    const double puum = (u_k + u_im) * (pu_c + pu_im), puup = (u_ip + u_k) * (pu_ip + pu_c);
    const double puvp = (pv_c + pv_ip) * (u_k + u_jp), puvm = (pv_jm + pv_jm_ip) * (u_jm + u_k);
    const double dut = ((puum - puup) * rdxj + (puvm - puvp) * rdy) * 0.25;
    const double pvvm = (v_k + v_jm) * (pv_c + pv_jm), pvvp = (v_jp + v_k) * (pv_jp + pv_c);
    const double pvup = (v_k + v_ip) * (pu_c + pu_jp), pvum = (v_im + v_k) * (pu_im + pu_jp_im);
    const double dvt = ((pvvm - pvvp) * rdy + (pvum - pvup) * rdxh) * 0.25;

    // pressure-gradient force, v direction (dynamics.py:160, :167-169)
    const double phiv = b_v * (phi[c_jp] - phi[c]);
    const double pgv = g.sig[k] * a_v / (rho[c] + rho[c_jp]);

    const double pu_n = u[c] * pu_fac - (dut + dus + pgf[c]) * dt;           // dynamics.py:206
    const double pv_n = v[c] * pv_fac - (dvt + dvs + phiv + pgv) * dt;       // dynamics.py:207
    out.u[o3 + c] = pu_n * r_pnu;
    double v_n = pv_n * r_pnv;
    if (zero_v) v_n *= 0.0;  // dynamics.py:222
    out.v[o3 + c] = v_n;

    // tracers: advec_t (dynamics.py:174-181) + advec_sig, flux form (dynamics.py:214, :219)
    {
      const double x_ip = st[c_ip], x_im = st[c_im], x_jp = st[c_jp], x_jm = st[c_jm];
      const double adv = ((pu_c * (t_k + x_ip) - pu_im * (x_im + t_k)) * rdxj +
                          (pv_c * (t_k + x_jp) - pv_jm * (x_jm + t_k)) * rdy) * 0.5;
      out.t[o3 + c] = (t[c] * p_c - (adv + ads_t) * dt) * r_pn;
    }
    {
      const double x_ip = sq[c_ip], x_im = sq[c_im], x_jp = sq[c_jp], x_jm = sq[c_jm];
      const double adv = ((pu_c * (q_k + x_ip) - pu_im * (x_im + q_k)) * rdxj +
                          (pv_c * (q_k + x_jp) - pv_jm * (x_jm + q_k)) * rdy) * 0.5;
      out.q[o3 + c] = (q[c] * p_c - (adv + ads_q) * dt) * r_pn;
    }
    fu = fu_n; fv = fv_n; ft = ft_n; fq = fq_n;
    u_k = u_kp; v_k = v_kp; t_k = t_kp; q_k = q_kp;
  }
  out.p[o2 + IDX2(j, i)] = pn_c;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int g_gcm_knob[8] = {0, 0, 0, 0, 0, 0, 0, 0};

// tuning knobs of the fast path (bench.py --knob i=v): 0 = min resident blocks of the update kernel (1..4),
// 1 = threads per block of the update kernel (32..160, 0 = automatic), 2 = threads of the row kernel (0 = auto)
extern "C" int gcm_tuning_knob(int idx, int value) {
  GCM_REQUIRE(idx >= 0 && idx < 8, GCM_ESHAPE);
  g_gcm_knob[idx] = value;
  return GCM_OK;
}

static size_t pf_row_smem(int W, int L) { return (size_t)((L + 1) / 2) * W * sizeof(double2); }

bool gcm_pe25_fast_supported(const gcm_geom* g) {
  const GcmGeomDev& d = g->d;
  if (d.L != 9 && d.L != 3) return false;
  if (!gcm_plan_inplace_ok(d.plan)) return false;
  return pf_row_smem(d.W, d.L) <= 200 * 1024;
}

template <int L>
static int pf_half_step(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out, double dt,
                        int nbatch, const PfWork& w, void* stream) {
  const GcmGeomDev& d = g->d;
  const int H = d.H, W = d.W;
  const size_t b2 = (size_t)H * W, b3 = (size_t)L * H * W;
  const int ja = d.row_lo, nrows = d.row_hi - d.row_lo;
  const int nrows_ext = d.wrap_j ? nrows : nrows + 1;  // band: also the first halo row to the south
  const size_t smem = pf_row_smem(W, L);
  const int work = ((L + 1) / 2) * W;
  int tr = (work / 4 + 31) / 32 * 32;  // about four elements of the FFT buffer per thread
  tr = tr < 64 ? 64 : (tr > 512 ? 512 : tr);
  if (g_gcm_knob[2] > 0) tr = g_gcm_knob[2];
  const PfConst cb{base->p, base->u, base->v, base->t, base->q};
  const PfConst cs{star->p, star->u, star->v, star->t, star->q};
  const PfMut mo{out->p, out->u, out->v, out->t, out->q};
  const bool ptop0 = d.ptop == 0.0;
#ifndef GCM_EMU
  if (smem > 48 * 1024) {
    GCM_CUDA(cudaFuncSetAttribute(pe25f_row_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GCM_CUDA(cudaFuncSetAttribute(pe25f_row_kernel<L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
#endif
  {
    GcmProfScope ps(GCM_K_ROW, stream);
    if (ptop0)
      GCM_LAUNCH((pe25f_row_kernel<L, true>), dim3(nrows_ext, nbatch), dim3(tr), smem, stream, d, base->p, cs, w, dt, ja,
                 b2, b3);
    else
      GCM_LAUNCH((pe25f_row_kernel<L, false>), dim3(nrows_ext, nbatch), dim3(tr), smem, stream, d, base->p, cs, w, dt, ja,
                 b2, b3);
  }
  GCM_CHECK_LAUNCH();
  int tc = g_gcm_knob[1] > 0 ? g_gcm_knob[1] : 160;
  if (g_gcm_knob[1] > 0) {
  } else if (W < 160) tc = (W + 31) / 32 * 32;
  else if (W % 128 == 0) tc = 128;
  else if (W % 96 == 0 && W % 160 != 0) tc = 96;
  {
    GcmProfScope ps(GCM_K_UPDATE_FAST, stream);
    const dim3 grid((W + tc - 1) / tc, nrows, nbatch);
    switch (g_gcm_knob[0]) {  // registers per thread vs resident warps (tuning knob 0)
      case 1: GCM_LAUNCH((pe25f_update_kernel<L, 1>), grid, dim3(tc), 0, stream, d, cb, cs, mo, w, dt, ja, b2, b3); break;
      case 3: GCM_LAUNCH((pe25f_update_kernel<L, 3>), grid, dim3(tc), 0, stream, d, cb, cs, mo, w, dt, ja, b2, b3); break;
      case 4: GCM_LAUNCH((pe25f_update_kernel<L, 4>), grid, dim3(tc), 0, stream, d, cb, cs, mo, w, dt, ja, b2, b3); break;
      default: GCM_LAUNCH((pe25f_update_kernel<L, 2>), grid, dim3(tc), 0, stream, d, cb, cs, mo, w, dt, ja, b2, b3); break;
    }
  }
  GCM_CHECK_LAUNCH();
  return GCM_OK;
}

int gcm_pe25_fast_half_step(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out,
                            double dt, int nbatch, double* spu, double* sd, double* phi, double* rho, double* pgf,
                            double* pn, void* stream) {
  const PfWork w{spu, sd, phi, rho, pgf, pn};
  if (g->d.L == 9) return pf_half_step<9>(g, base, star, out, dt, nbatch, w, stream);
  return pf_half_step<3>(g, base, star, out, dt, nbatch, w, stream);
}
