// prof.h -- optional per-kernel timing with CUDA events on the launching stream (used by bench.py for the
// live roofline of the dominant kernel; off by default, no cost when off).
#pragma once
#include "gcm_common.h"

enum GcmProfKind {
  GCM_K_FILTER_SPU = 0,   // pe25_filter_kernel<1>
  GCM_K_COLUMN = 1,       // pe25_column_kernel
  GCM_K_FILTER_PGF = 2,   // pe25_pgf_filter_kernel
  GCM_K_UPDATE = 3,       // pe25_update_kernel
  GCM_K_ROW = 4,          // (unused: the single-launch row kernel of rounds r01d-r01o)
  GCM_K_UPDATE_FAST = 5,  // pe25f_update_kernel (pe25_fast.cu, direct loads)
  GCM_K_FILTER_A = 6,     // pe25f_filter_kernel<1>
  GCM_K_COLUMN_F = 7,     // pe25f_hydro_kernel
  GCM_K_FILTER_B = 8,     // pe25f_filter_kernel<0>
  GCM_K_AFLUX_F = 9,      // pe25f_aflux_kernel
  GCM_K_UPDATE_TILED = 10,  // pe25f_update_tiled_kernel
  GCM_K_EXTRAS = 11,        // pe25x_extras_kernel (pe25_extras.cu, opt-in terms)
  GCM_K_HALO = 12,          // one halo exchange of a latitude band: pack + NCCL / peer copies + unpack (comm.cu)
  GCM_K_UPDATE_TMA = 13,    // pe25f_update_tma_kernel (TMA box loads + mbarrier)
  GCM_K_COUNT = 14
};

#ifdef GCM_EMU
struct GcmProfScope {
  GcmProfScope(int, void*) {}
};
#else
void gcm_prof_begin(int kind, void* stream);
void gcm_prof_end(int kind, void* stream);
extern int g_gcm_prof_on;
struct GcmProfScope {
  int kind;
  void* stream;
  GcmProfScope(int k, void* s) : kind(k), stream(s) {
    if (g_gcm_prof_on) gcm_prof_begin(kind, stream);
  }
  ~GcmProfScope() {
    if (g_gcm_prof_on) gcm_prof_end(kind, stream);
  }
};
#endif
