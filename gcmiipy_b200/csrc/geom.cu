// geom.cu -- device-resident metric / sigma / filter tables (reference geometry.py:9-182, low_pass.py:61-72)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "gcm_common.h"
#include "fft_rows.h"

extern "C" int gcm_version(void) { return 100; }
int g_gcm_tuning_epoch = 0;

static thread_local int tl_last_status = 0;
int gcm_set_status(int st) {
  if (st != 0) tl_last_status = st;
  return st;
}
extern "C" int gcm_last_status(int clear) {
  const int st = tl_last_status;
  if (clear) tl_last_status = 0;
  return st;
}

// one counter per device, allocated on first use and never freed (4 bytes)
unsigned int* gcm_nonfinite_word() {
#ifdef GCM_EMU
  static unsigned int word = 0;
  return &word;
#else
  static unsigned int* words[64] = {nullptr};
  static std::mutex mu;  // two threads may create their first geometry at the same time
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!words[dev]) {
    unsigned int* w = nullptr;
    if (cudaMalloc((void**)&w, 256) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    cudaMemset(w, 0, 256);
    words[dev] = w;
  }
  return words[dev];
#endif
}

extern "C" int gcm_nonfinite_read(unsigned int* host_out, int reset, void* stream) {
  unsigned int* w = gcm_nonfinite_word();
  GCM_REQUIRE(w, (int)cudaErrorMemoryAllocation);
  if (host_out) GCM_CUDA(cudaMemcpyAsync(host_out, w, sizeof(unsigned int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (reset) GCM_CUDA(cudaMemsetAsync(w, 0, sizeof(unsigned int), (cudaStream_t)stream));
  return GCM_OK;
}

extern "C" const char* gcm_status_string(int s) {
  switch (s) {
    case GCM_OK: return "ok";
    case GCM_ENULL: return "required pointer is NULL";
    case GCM_ESHAPE: return "bad extent or row range";
    case GCM_EALIGN: return "device pointer not 16-byte aligned";
    case GCM_EUNSUP: return "unsupported combination";
    case GCM_EWORK: return "workspace too small";
    default:
      if (s >= 1000000) return "NCCL error (status - 1000000 = ncclResult_t)";
      return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown status";
  }
}

// Radix plan of one row transform.  The 2-3-5-smooth part of n is split into as few in-register butterflies
// as possible (radix <= 16, fft_rows.h), cheapest combination first by an estimate of flops per point plus a
// fixed charge per pass (one shared-memory round trip, one barrier, one twiddle multiply); an odd radix goes
// last so that the unit-stride stage is free of shared-memory bank conflicts.  Any other prime factor takes
// the generic O(R^2) pass.
static const int kRadix[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};
static const double kRadixCost[] = {10.5, 16.5, 11.33, 12.4, 13.33, 7.0, 9.33, 8.0, 4.0, 5.33, 2.0};
static const double kPassCost = 10.0;

static void plan_search(int m, int first, int* cur, int ncur, double cost, int* best, int* nbest, double* bestcost) {
  if (m == 1) {
    if (cost < *bestcost) {
      *bestcost = cost;
      *nbest = ncur;
      memcpy(best, cur, ncur * sizeof(int));
    }
    return;
  }
  if (ncur >= 12 || cost >= *bestcost) return;
  for (int f = first; f < (int)(sizeof(kRadix) / sizeof(int)); ++f)
    if (m % kRadix[f] == 0) {
      cur[ncur] = kRadix[f];
      plan_search(m / kRadix[f], f, cur, ncur + 1, cost + kRadixCost[f] + kPassCost, best, nbest, bestcost);
    }
}

int gcm_fft_make_plan(int n, GcmFftPlan* plan) {
  plan->n = n;
  plan->npass = 0;
  int m = n, smooth = 1;
  const int small[3] = {2, 3, 5};
  for (int f = 0; f < 3; ++f)
    while (m % small[f] == 0) {
      m /= small[f];
      smooth *= small[f];
    }
  int cur[16], best[16], nbest = 0;
  double bestcost = 1e30;
  plan_search(smooth, 0, cur, 0, 0.0, best, &nbest, &bestcost);
  // descending, one odd radix (if any) moved to the end
  for (int a = 0; a < nbest; ++a)
    for (int b = a + 1; b < nbest; ++b)
      if (best[b] > best[a]) { int t = best[a]; best[a] = best[b]; best[b] = t; }
  for (int a = 0; a < nbest; ++a)
    if (best[a] % 2) {
      const int odd = best[a];
      for (int b = a; b + 1 < nbest; ++b) best[b] = best[b + 1];
      best[nbest - 1] = odd;
      break;
    }
  // tuning hook: GCM_FFT_PLAN="16,10,9" overrides the radix sequence of the smooth part (product must match)
  if (const char* env = getenv("GCM_FFT_PLAN")) {
    int over[16], no = 0, prod = 1;
    for (const char* c = env; *c && no < 16;) {
      const int r = atoi(c);
      if (r < 2) break;
      over[no++] = r;
      prod *= r;
      while (*c && *c != ',') ++c;
      if (*c == ',') ++c;
    }
    bool ok = prod == smooth;
    for (int a = 0; a < no; ++a) ok = ok && gcm_radix_unrolled(over[a]);
    if (ok) {
      nbest = no;
      memcpy(best, over, no * sizeof(int));
    }
  }
  for (int a = 0; a < nbest; ++a) plan->radix[plan->npass++] = best[a];
  for (int p = 7; m > 1; p += 2)
    while (m % p == 0) {
      if (plan->npass >= GCM_MAX_RADIX_PASSES) return GCM_EUNSUP;
      plan->radix[plan->npass++] = p;
      m /= p;
    }
  int len = n;
  for (int s = 0; s < plan->npass; ++s) {
    const int r = plan->radix[s];
    plan->stride[s] = len / r;
    plan->magic_stride[s] = gcm_magic((unsigned)(len / r));
    plan->magic_nbf[s] = gcm_magic((unsigned)(n / r));
    plan->blk[s] = len;
    plan->nbf[s] = n / r;
    plan->tstep[s] = n / len;
    plan->twoff[s] = s == 0 ? 0 : plan->twoff[s - 1] + (plan->radix[s - 1] - 1) * plan->stride[s - 1];
    len /= r;
  }
  return GCM_OK;
}

static size_t up256(size_t b) { return (b + 255) / 256 * 256; }

extern "C" int gcm_geom_create(const gcm_geom_desc* d, gcm_geom** out) {
  GCM_REQUIRE(d && out, GCM_ENULL);
  GCM_REQUIRE(d->h_sig && d->h_dsig && d->h_sigb && d->h_sigt && d->h_dx_j && d->h_dx_h, GCM_ENULL);
  const int H = d->H, W = d->W, L = d->L;
  GCM_REQUIRE(H > 0 && W > 0 && L > 0, GCM_ESHAPE);
  GCM_REQUIRE(W == 1 || W % 2 == 0, GCM_ESHAPE);  // the reference filter breaks on odd W (low_pass.py:57)
  GCM_REQUIRE(W == 1 || d->h_smmz, GCM_ENULL);
  GCM_REQUIRE(d->row_lo >= 0 && d->row_lo < d->row_hi && d->row_hi <= H, GCM_ESHAPE);
  if (d->wrap_j) GCM_REQUIRE(d->row_lo == 0 && d->row_hi == H, GCM_ESHAPE);
  // a band needs 1 halo row to the north and 2 to the south of its owned rows
  if (!d->wrap_j) GCM_REQUIRE(d->row_lo >= 1 && d->row_hi + 2 <= H, GCM_ESHAPE);
  GCM_REQUIRE(d->zero_v_row >= -1 && d->zero_v_row < H, GCM_ESHAPE);
  GCM_REQUIRE(d->zero_v_row2 >= -1 && d->zero_v_row2 < H, GCM_ESHAPE);

  gcm_geom* g = (gcm_geom*)calloc(1, sizeof(gcm_geom));
  GCM_REQUIRE(g, (int)cudaErrorMemoryAllocation);
  int st = gcm_fft_make_plan(W, &g->d.plan);
  if (st != GCM_OK) { free(g); return st; }

  const size_t nw = (size_t)W / 2 + 1;
  const size_t o_sig = 0, o_dsig = o_sig + up256(L * 8), o_sigb = o_dsig + up256(L * 8), o_sigt = o_sigb + up256(L * 8);
  const size_t o_dxj = o_sigt + up256(L * 8), o_dxh = o_dxj + up256(H * 8), o_hmap = o_dxh + up256(H * 8);
  const size_t o_smmz = o_hmap + up256((size_t)H * W * 8), o_tw = o_smmz + up256((size_t)H * nw * 8);
  const size_t o_rdxj = o_tw + up256((size_t)W * 16), o_rdxh = o_rdxj + up256(H * 8);
  const size_t o_rdsig = o_rdxh + up256(H * 8), o_sigkap = o_rdsig + up256(L * 8);
  const size_t o_kperm = o_sigkap + up256(L * 8);
  const size_t o_smmzp = o_kperm + up256((size_t)W * 4);
  const GcmFftPlan& pl0 = g->d.plan;
  size_t ntws = 0;
  for (int s2 = 0; s2 < pl0.npass; ++s2) ntws += (size_t)(pl0.radix[s2] - 1) * pl0.stride[s2];
  const size_t o_tws = o_smmzp + up256((size_t)H * W * 8);
  const size_t total = o_tws + up256(ntws * 16);

  std::vector<unsigned char> host(total, 0);
  memcpy(&host[o_sig], d->h_sig, L * 8);
  memcpy(&host[o_dsig], d->h_dsig, L * 8);
  memcpy(&host[o_sigb], d->h_sigb, L * 8);
  memcpy(&host[o_sigt], d->h_sigt, L * 8);
  memcpy(&host[o_dxj], d->h_dx_j, H * 8);
  memcpy(&host[o_dxh], d->h_dx_h, H * 8);
  if (d->h_heightmap) memcpy(&host[o_hmap], d->h_heightmap, (size_t)H * W * 8);
  if (d->h_smmz) memcpy(&host[o_smmz], d->h_smmz, (size_t)H * nw * 8);
  double* tw = reinterpret_cast<double*>(&host[o_tw]);
  for (int m = 0; m < W; ++m) {  // exp(-2 pi i m / W), evaluated in long double
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)W;
    tw[2 * m] = (double)cosl(a);
    tw[2 * m + 1] = (double)(-sinl(a));
  }

  {
    double* rdxj = reinterpret_cast<double*>(&host[o_rdxj]);
    double* rdxh = reinterpret_cast<double*>(&host[o_rdxh]);
    double* rdsig = reinterpret_cast<double*>(&host[o_rdsig]);
    double* sigkap = reinterpret_cast<double*>(&host[o_sigkap]);
    int* kperm = reinterpret_cast<int*>(&host[o_kperm]);
    for (int j = 0; j < H; ++j) { rdxj[j] = 1.0 / d->h_dx_j[j]; rdxh[j] = 1.0 / d->h_dx_h[j]; }
    for (int k = 0; k < L; ++k) { rdsig[k] = 1.0 / d->h_dsig[k]; sigkap[k] = pow(d->h_sig[k], GCM_KAPPA); }
    const GcmFftPlan& pl = g->d.plan;
    for (int p = 0; p < W; ++p) {  // position after the forward DIF stages -> wavenumber (fft_inplace.h)
      int k = 0, mult = 1, n = W, rem = p;
      for (int s = 0; s < pl.npass; ++s) {
        n /= pl.radix[s];
        k += (rem / n) * mult;
        rem %= n;
        mult *= pl.radix[s];
      }
      kperm[p] = k;
    }
    if (d->h_smmz) {  // multiplier in transform order, 1/W folded in
      double* sp = reinterpret_cast<double*>(&host[o_smmzp]);
      const double inv = 1.0 / W;
      for (int j = 0; j < H; ++j)
        for (int p = 0; p < W; ++p) {
          const int k = kperm[p];
          sp[(size_t)j * W + p] = d->h_smmz[(size_t)j * nw + (k <= W - k ? k : W - k)] * inv;
        }
    }
    double* tws = reinterpret_cast<double*>(&host[o_tws]);
    const long double twopi = 2.0L * 3.14159265358979323846264338327950288L;
    for (int s2 = 0; s2 < pl.npass; ++s2) {
      const int R = pl.radix[s2], stride = pl.stride[s2], nblk = pl.blk[s2];
      for (int m = 1; m < R; ++m)
        for (int q = 0; q < stride; ++q) {
          const long double a = twopi * (long double)((long long)q * m % nblk) / (long double)nblk;
          const size_t e = (size_t)pl.twoff[s2] + (size_t)(m - 1) * stride + q;
          tws[2 * e] = (double)cosl(a);
          tws[2 * e + 1] = (double)(-sinl(a));
        }
    }
  }

  void* blk = nullptr;
  cudaError_t e = cudaMalloc(&blk, total);
  if (e != cudaSuccess) { free(g); return (int)e; }
  e = cudaMemcpy(blk, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(blk); free(g); return (int)e; }

  unsigned char* b = (unsigned char*)blk;
  g->d_block = blk;
  g->d.H = H; g->d.W = W; g->d.L = L;
  g->d.wrap_j = d->wrap_j ? 1 : 0;
  g->d.row_lo = d->row_lo; g->d.row_hi = d->row_hi;
  g->d.zero_v_row = d->zero_v_row;
  g->d.zero_v_row2 = d->zero_v_row2;
  g->d.dy = d->dy; g->d.ptop = d->ptop;
  g->d.sig = (const double*)(b + o_sig);
  g->d.dsig = (const double*)(b + o_dsig);
  g->d.sigb = (const double*)(b + o_sigb);
  g->d.sigt = (const double*)(b + o_sigt);
  g->d.dx_j = (const double*)(b + o_dxj);
  g->d.dx_h = (const double*)(b + o_dxh);
  g->d.hmap = (const double*)(b + o_hmap);
  g->d.smmz = (const double*)(b + o_smmz);
  g->d.tw = (const double2*)(b + o_tw);
  g->d.rdx_j = (const double*)(b + o_rdxj);
  g->d.rdx_h = (const double*)(b + o_rdxh);
  g->d.rdsig = (const double*)(b + o_rdsig);
  g->d.sigkap = (const double*)(b + o_sigkap);
  g->d.kperm = (const int*)(b + o_kperm);
  g->d.smmzp = (const double*)(b + o_smmzp);
  g->d.tws = (const double2*)(b + o_tws);
  g->d.rdy = 1.0 / d->dy;
  g->d.nonfinite = gcm_nonfinite_word();
  for (int k = 0; k < L && k < GCM_MAXLC; ++k) {
    g->d.c_sig[k] = d->h_sig[k];
    g->d.c_dsig[k] = d->h_dsig[k];
    g->d.c_sigb[k] = d->h_sigb[k];
    g->d.c_sigt[k] = d->h_sigt[k];
    g->d.c_rdsig[k] = 1.0 / d->h_dsig[k];
    g->d.c_sigkap[k] = pow(d->h_sig[k], GCM_KAPPA);
  }
  *out = g;
  return GCM_OK;
}

int gcm_geom_aux(const gcm_geom* cg, void** stream, void** ev_fork, void** ev_join) {
  gcm_geom* g = const_cast<gcm_geom*>(cg);
#ifndef GCM_EMU
  if (!g->aux_stream) {
    cudaStream_t q;
    cudaEvent_t a, b;
    if (g_gcm_knob[3] == 2) {  // the side stream carries the longer chain (hydro -> filter): schedule its CTAs first
      int lo = 0, hi = 0;
      GCM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      GCM_CUDA(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, hi));
    } else {
      GCM_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    }
    GCM_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    GCM_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    g->aux_stream = q;
    g->ev_fork = a;
    g->ev_join = b;
  }
#endif
  *stream = g->aux_stream;
  *ev_fork = g->ev_fork;
  *ev_join = g->ev_join;
  return GCM_OK;
}

extern "C" int gcm_geom_destroy(gcm_geom* g) {
  if (!g) return GCM_OK;
#ifndef GCM_EMU
  for (int s = 0; s < 2; ++s)
    if (g->graph[s].exec) cudaGraphExecDestroy((cudaGraphExec_t)g->graph[s].exec);
  if (g->cap_stream) cudaStreamDestroy((cudaStream_t)g->cap_stream);
  if (g->aux_stream) cudaStreamDestroy((cudaStream_t)g->aux_stream);
  if (g->ev_fork) cudaEventDestroy((cudaEvent_t)g->ev_fork);
  if (g->ev_join) cudaEventDestroy((cudaEvent_t)g->ev_join);
#endif
  cudaFree(g->d_block);
  if (g->d_cor) cudaFree(g->d_cor);
  free(g);
  return GCM_OK;
}
