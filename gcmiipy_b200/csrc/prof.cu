// prof.cu -- see prof.h
#include "prof.h"

#include <vector>

#ifndef GCM_EMU
int g_gcm_prof_on = 0;
static std::vector<cudaEvent_t> g_ev[GCM_K_COUNT][2];
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_pool.empty()) { e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
void gcm_prof_begin(int kind, void* stream) {
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, (cudaStream_t)stream);
  g_ev[kind][0].push_back(e);
}
void gcm_prof_end(int kind, void* stream) {
  cudaEvent_t e = prof_event();
  cudaEventRecord(e, (cudaStream_t)stream);
  g_ev[kind][1].push_back(e);
}
#endif

extern "C" int gcm_prof_enable(int on) {
#ifndef GCM_EMU
  g_gcm_prof_on = on ? 1 : 0;
#else
  (void)on;
#endif
  return GCM_OK;
}

extern "C" int gcm_prof_kinds(void) { return GCM_K_COUNT; }

extern "C" const char* gcm_prof_kind_name(int kind) {
  switch (kind) {
    case GCM_K_FILTER_SPU: return "pe25_filter_kernel<1>";
    case GCM_K_COLUMN: return "pe25_column_kernel";
    case GCM_K_FILTER_PGF: return "pe25_pgf_filter_kernel";
    case GCM_K_UPDATE: return "pe25_update_kernel";
    case GCM_K_ROW: return "pe25f_row_kernel";
    case GCM_K_UPDATE_FAST: return "pe25f_update_kernel";
    case GCM_K_FILTER_A: return "pe25f_filter_kernel<1>";
    case GCM_K_COLUMN_F: return "pe25f_hydro_kernel";
    case GCM_K_AFLUX_F: return "pe25f_aflux_kernel";
    case GCM_K_UPDATE_TILED: return "pe25f_update_tiled_kernel";
    case GCM_K_FILTER_B: return "pe25f_filter_kernel<0>";
    case GCM_K_EXTRAS: return "pe25x_extras_kernel";
    case GCM_K_HALO: return "halo_exchange";
    case GCM_K_UPDATE_TMA: return "pe25f_update_tma_kernel";
    default: return "?";
  }
}

// Synchronises the device, then writes per kind the summed milliseconds and the launch count recorded
// since the last collect; the records are cleared.
extern "C" int gcm_prof_collect(double* ms, long long* launches) {
  GCM_REQUIRE(ms && launches, GCM_ENULL);
#ifndef GCM_EMU
  GCM_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < GCM_K_COUNT; ++k) {
    double tot = 0.0;
    const size_t n = g_ev[k][0].size() < g_ev[k][1].size() ? g_ev[k][0].size() : g_ev[k][1].size();
    for (size_t i = 0; i < n; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, g_ev[k][0][i], g_ev[k][1][i]);
      tot += t;
    }
    for (int s = 0; s < 2; ++s) {
      for (cudaEvent_t e : g_ev[k][s]) g_pool.push_back(e);
      g_ev[k][s].clear();
    }
    ms[k] = tot;
    launches[k] = (long long)n;
  }
#else
  for (int k = 0; k < GCM_K_COUNT; ++k) { ms[k] = 0.0; launches[k] = 0; }
#endif
  return GCM_OK;
}
