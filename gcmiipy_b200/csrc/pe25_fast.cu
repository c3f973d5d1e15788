// pe25_fast.cu -- host entry of the fused 2.5-D half step (kernels: pe25_fast_impl.h, one translation unit per layer
// count: pe25_fast_l3.cu, _l9.cu, _l17.cu, _l18.cu), tuning knobs.
#include "fft_inplace.h"
#include "gcm_common.h"

struct PfWork {
  double *spu, *sd, *pgf, *fv, *pn, *pit;
};
#define GCM_PF_DECLARE(LAYERS)                                                                                           \
  int gcm_pf_half_step_l##LAYERS(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out, \
                                 double dt, int nbatch, const PfWork& w, GcmRowSeg sr, GcmRowSeg su, void* stream);
GCM_PF_DECLARE(3)
GCM_PF_DECLARE(9)
GCM_PF_DECLARE(17)
GCM_PF_DECLARE(18)

int g_gcm_knob[GCM_NKNOBS] = {0};

// tuning knobs (bench.py --knob i=v; 0 = automatic):
//   0  threads of the filter kernel                 1  packed rows (layer pairs) per CTA of the filter kernel
//   2  rows per warp task of the hydro kernel (RG)  3  1 = the two chains one after the other on the caller's stream; 2 = side stream at the highest priority
//   4  1 = update kernel with direct global loads even when W % 32 == 0; 2 = never the one-thread-per-cell update (and
//      no 36-wide tiles: the column march); 6 = no 36-wide tiles;
//      3 = the one-thread-per-cell update for every member count; 5 = the warp-specialised TMA update (pe25f_update_tma_kernel)
//      instead of the LDGSTS tiled update
//   5  direct-load update kernel: L1 prefetch distance in layers + 1 (1 = off)
//   6  latitude blocks of the host-resident step (host_step.cu)
//   7  hydro kernel: 0 = W < 62 pe25f_hydro_narrow_kernel, else the tile form pe25f_hydro_tile_kernel (one column per
//      thread); 1 = the marching warp-chunk kernel everywhere; 3 = the marching kernel on wide grids only
//   8  1 = filter kernel with the runtime radix switch even when the plan has a compile-time twin (GcmFixedPlan)
//   9  programmatic dependent launch of the half-step kernels: 0 / 1 = on (a kernel's launch overlaps the drain of its
//      predecessor in the stream: -3 ... -5 % per step, r03e), 2 = on + every kernel triggers its dependents at entry
//      (waiting CTAs then hold SM slots the running kernel could use: slower on most grids), 3 = off
//  10  TMA update kernel: layers in flight (3; 0 = 4)     11  TMA update kernel: tile rows (8; 0 = 4)
//  13  TMA update kernel with 4-row tiles: 1 = 4 CTAs per SM (96 registers, spills) instead of 3 (128 registers)
//  12  L2 promotion of the tensor maps: 0 = 128 B, 1 = none, 2 = 256 B
//  14  2 = the persistent pipelined filter kernel (bulk-copy prefetch of the next rows) instead of the one-unit-per-CTA one
//  15  register budgets (L = 9): units digit = CTAs per SM the filter aims at (4, 6, 8; default 5), tens = hydro (5, 6, 8;
//      default 4), hundreds = tiled update (5, 6; default 4)
//  16  1 = aflux fused into the filter of the mass flux (per-pair partial sums of conv) instead of its own kernel
extern "C" int gcm_tuning_knob(int idx, int value) {
  GCM_REQUIRE(idx >= 0 && idx < GCM_NKNOBS, GCM_ESHAPE);
  g_gcm_knob[idx] = value;
  ++g_gcm_tuning_epoch;
  return GCM_OK;
}

bool gcm_pe25_fast_supported(const gcm_geom* g) {
  const GcmGeomDev& d = g->d;
  if (d.L != 9 && d.L != 3 && d.L != 17 && d.L != 18) return false;  // layer counts with an instantiation (others: pe25.cu)
  if (d.W < 2 || !gcm_plan_inplace_ok(d.plan)) return false;
  if ((double)d.L * d.H * d.W >= 2147483648.0) return false;  // 32-bit element offsets within a member
  return (size_t)d.W * sizeof(double2) <= 200 * 1024;
}

int gcm_pe25_fast_half_step(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out,
                            double dt, int nbatch, double* spu, double* sd, double* fv, double* pgf, double* pn,
                            double* pit, const int* seg_r, const int* seg_u, void* stream) {
  const PfWork w{spu, sd, pgf, fv, pn, pit};
  const GcmGeomDev& d = g->d;
  const int nrows = d.row_hi - d.row_lo;
  GcmRowSeg sr{d.row_lo, d.wrap_j ? nrows : nrows + 1, 0, 0};  // band: also the first halo row to the south
  GcmRowSeg su{d.row_lo, nrows, 0, 0};
  if (seg_r) sr = GcmRowSeg{seg_r[0], seg_r[1], seg_r[2], seg_r[3]};
  if (seg_u) su = GcmRowSeg{seg_u[0], seg_u[1], seg_u[2], seg_u[3]};
  switch (d.L) {
    case 9: return gcm_pf_half_step_l9(g, base, star, out, dt, nbatch, w, sr, su, stream);
    case 3: return gcm_pf_half_step_l3(g, base, star, out, dt, nbatch, w, sr, su, stream);
    case 17: return gcm_pf_half_step_l17(g, base, star, out, dt, nbatch, w, sr, su, stream);
    case 18: return gcm_pf_half_step_l18(g, base, star, out, dt, nbatch, w, sr, su, stream);
    default: return gcm_set_status(GCM_EUNSUP);
  }
}
