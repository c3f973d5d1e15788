// gcm_tma.h -- Tensor Memory Accelerator helpers (sm_100a): tiled tensor maps of the fp64 fields, bulk tensor loads
// (cp.async.bulk.tensor -> SASS UTMALDG), 1-D bulk copies (cp.async.bulk -> UBLKCP) and the mbarrier they complete on.
//
// A field is described to the TMA unit as a 3-D tensor [NZ][H][W] (W fastest) of doubles with a fixed box of
// bw x bh x 1 elements.  A box may start at negative coordinates or run past the extents: the out-of-range part is
// zero-filled by the hardware (the callers patch the periodic wrap afterwards, pe25_fast.cu).
//
// On the CPU emulator build (tests only) a "tensor map" is the plain description and a load is a synchronous copy with
// the same zero fill; the mbarrier is emulated with its real phase semantics (a waiting fiber yields until the phase
// flips), so warp-specialised producer / consumer kernels run on the emulator too.
#pragma once
#include "gcm_common.h"

#ifdef GCM_EMU
struct GcmTmap {
  const double* base;
  int W, H, NZ, bw, bh;
};
#define GCM_GRID_CONSTANT
#else
#include <cuda.h>
typedef CUtensorMap GcmTmap;
#define GCM_GRID_CONSTANT __grid_constant__
#endif

// Host: the tensor map of `base` viewed as [NZ][H][W] doubles with box bw x bh x 1 (cached per (pointer, extents, box):
// a ping-pong run uses a handful).  bw * 8 must be a multiple of 16 bytes, W even, base 16-byte aligned.
// Returns GCM_OK, GCM_EUNSUP when the driver has no cuTensorMapEncodeTiled, or a CUDA error.
int gcm_tmap_get(GcmTmap* out, const double* base, int W, int H, int NZ, int bw, int bh);

// ---- device side ------------------------------------------------------------------------------------------------
#ifdef GCM_EMU
// mbarrier on the emulator: bit 63 = phase, bits 32..62 = arrivals per phase, bits 0..31 = arrivals still pending.  A
// block's fibers run on one OS thread and switch only when they yield, so plain read-modify-write is atomic; the bytes
// of a bulk copy "land" at issue, so expect_tx is an arrival.  A waiter yields (stays runnable) until the phase flips.
typedef unsigned long long GcmMbar;
__device__ __forceinline__ void gcm_mbar_init(GcmMbar* b, int count) {
  *b = ((unsigned long long)count << 32) | (unsigned long long)count;
}
__device__ __forceinline__ void gcm_mbar_fence_init() {}
__device__ __forceinline__ void gcm_fence_proxy_async() {}
__device__ __forceinline__ void gcm_mbar_arrive(GcmMbar* b) {
  unsigned long long v = *b;
  const unsigned long long init = (v >> 32) & 0x7fffffffull;
  unsigned long long pending = (v & 0xffffffffull) - 1;
  unsigned long long phase = v >> 63;
  if (pending == 0) { phase ^= 1; pending = init; }
  *b = (phase << 63) | (init << 32) | pending;
}
__device__ __forceinline__ void gcm_mbar_expect_tx(GcmMbar* b, unsigned) { gcm_mbar_arrive(b); }
__device__ __forceinline__ void gcm_mbar_wait(GcmMbar* b, unsigned parity) {
  while ((unsigned)(*b >> 63) == (parity & 1u)) gcm_emu::yield_state(0);
}
__device__ __forceinline__ void gcm_mbar_wait_backoff(GcmMbar* b, unsigned parity) { gcm_mbar_wait(b, parity); }
__device__ __forceinline__ void gcm_tma_load3(double* dst, const GcmTmap* m, int x, int y, int z, GcmMbar*) {
  for (int r = 0; r < m->bh; ++r)
    for (int c = 0; c < m->bw; ++c) {
      const int gi = x + c, gj = y + r;
      const bool in = gi >= 0 && gi < m->W && gj >= 0 && gj < m->H && z >= 0 && z < m->NZ;
      dst[r * m->bw + c] = in ? m->base[((size_t)z * m->H + gj) * m->W + gi] : 0.0;
    }
}
__device__ __forceinline__ void gcm_bulk_load(double* dst, const double* src, unsigned bytes, GcmMbar*) {
  memcpy(dst, src, bytes);
}
#else
typedef unsigned long long GcmMbar;
__device__ __forceinline__ unsigned gcm_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gcm_mbar_init(GcmMbar* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gcm_smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (TMA) -- once, by the initialising thread, before the
// block-wide barrier that publishes them
__device__ __forceinline__ void gcm_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void gcm_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void gcm_mbar_expect_tx(GcmMbar* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gcm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gcm_mbar_arrive(GcmMbar* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gcm_smem_u32(bar)) : "memory");
}
// blocks until the phase with the given parity has completed (every expected byte has landed)
__device__ __forceinline__ void gcm_mbar_wait(GcmMbar* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(gcm_smem_u32(bar)), "r"(parity)
      : "memory");
}
// the same for a thread that has nothing else to do (a producer waiting for its consumers): sleeps between probes
// instead of competing with the compute warps for issue slots
__device__ __forceinline__ void gcm_mbar_wait_backoff(GcmMbar* bar, unsigned parity) {
  for (;;) {
    unsigned done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(gcm_smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(100);
  }
}
// box of the tensor map at element coordinates (x, y, z) -> dst (128-byte aligned shared memory), completing on bar
__device__ __forceinline__ void gcm_tma_load3(double* dst, const GcmTmap* m, int x, int y, int z, GcmMbar* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          gcm_smem_u32(dst)),
      "l"(m), "r"(gcm_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
// contiguous bytes (multiple of 16, both addresses 16-byte aligned) global -> shared, completing on bar
__device__ __forceinline__ void gcm_bulk_load(double* dst, const double* src, unsigned bytes, GcmMbar* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   gcm_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(gcm_smem_u32(bar))
               : "memory");
}
#endif
