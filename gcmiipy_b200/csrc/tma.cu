// tma.cu -- host side of gcm_tma.h: tiled tensor maps of the fp64 fields, encoded through the driver entry point
// cuTensorMapEncodeTiled (looked up at run time: the library does not link libcuda) and cached, because a ping-pong
// run asks for the same handful of (pointer, extents, box) over and over and an encode costs microseconds of host time.
#include "gcm_tma.h"

#include <string.h>

#include <mutex>

#ifdef GCM_EMU
int gcm_tmap_get(GcmTmap* out, const double* base, int W, int H, int NZ, int bw, int bh) {
  GCM_REQUIRE(out && base, GCM_ENULL);
  out->base = base;
  out->W = W; out->H = H; out->NZ = NZ; out->bw = bw; out->bh = bh;
  return GCM_OK;
}
#else
typedef CUresult (*GcmEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapEntry {
  const double* base;
  int W, H, NZ, bw, bh;
  unsigned long long stamp;
  int epoch;  // tuning epoch the map was encoded under (knob 12 = L2 promotion)
  CUtensorMap map;
};
#define GCM_TMAP_CACHE 64
static TmapEntry g_cache[GCM_TMAP_CACHE];
static unsigned long long g_stamp = 0;
static std::mutex g_mu;
static GcmEncodeTiled g_encode = nullptr;
static int g_encode_state = 0;  // 0 = not looked up, 1 = available, -1 = missing

int gcm_tmap_get(GcmTmap* out, const double* base, int W, int H, int NZ, int bw, int bh) {
  GCM_REQUIRE(out && base, GCM_ENULL);
  GCM_REQUIRE(W > 0 && H > 0 && NZ > 0 && bw > 0 && bh > 0 && bw <= 256 && bh <= 256, GCM_ESHAPE);
  GCM_REQUIRE((W % 2) == 0 && (bw % 2) == 0 && gcm_aligned16(base), GCM_EALIGN);
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_encode_state == 0) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
        qres == cudaDriverEntryPointSuccess) {
      g_encode = (GcmEncodeTiled)fn;
      g_encode_state = 1;
    } else {
      cudaGetLastError();
      g_encode_state = -1;
    }
  }
  GCM_REQUIRE(g_encode_state == 1, GCM_EUNSUP);
  int victim = 0;
  for (int e = 0; e < GCM_TMAP_CACHE; ++e) {
    TmapEntry& t = g_cache[e];
    if (t.base == base && t.W == W && t.H == H && t.NZ == NZ && t.bw == bw && t.bh == bh && t.epoch == g_gcm_tuning_epoch) {
      t.stamp = ++g_stamp;
      memcpy(out, &t.map, sizeof(CUtensorMap));
      return GCM_OK;
    }
    if (t.stamp < g_cache[victim].stamp) victim = e;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NZ};
  const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(double), (cuuint64_t)W * H * sizeof(double)};  // bytes, dims 1..2
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUtensorMap m;
  const CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              g_gcm_knob[12] == 1 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                                  : (g_gcm_knob[12] == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                                         : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GCM_REQUIRE(r == CUDA_SUCCESS, GCM_EUNSUP);
  TmapEntry& t = g_cache[victim];
  t.base = base; t.W = W; t.H = H; t.NZ = NZ; t.bw = bw; t.bh = bh;
  t.stamp = ++g_stamp;
  t.epoch = g_gcm_tuning_epoch;
  t.map = m;
  memcpy(out, &m, sizeof(CUtensorMap));
  return GCM_OK;
}
#endif
