// gcm_common.h -- shared declarations of the sm_100a kernels behind include/gcm_b200.h
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "gcm_b200.h"

#define GCM_NKNOBS 24  // tuning knobs (gcm_tuning_knob)

#ifdef GCM_EMU
#include "cuda_emu.h"  // tests/emu: CPU execution of these sources for the no-GPU test-suite only
#define GCM_LAUNCH(kern, grid, block, smem, stream, ...) \
  gcm_emu::launch((grid), (block), (smem), [=]() { kern(__VA_ARGS__); })
#define GCM_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(gcm_emu::dyn_smem())
#else
#include <cuda_runtime.h>
#define GCM_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#define GCM_DYN_SMEM(type, name)                                 \
  extern __shared__ __align__(128) unsigned char gcm_dyn_smem_[]; \
  type* name = reinterpret_cast<type*>(gcm_dyn_smem_)
#endif

extern int g_gcm_knob[GCM_NKNOBS];  // tuning knobs (pe25_fast.cu)
#ifdef GCM_EMU
#define GCM_LAUNCH_DEP GCM_LAUNCH
#else
// launch with the programmatic-stream-serialization attribute when `pdl` is set (see gcm_pdl_wait), else like <<<>>>
template <class... KArgs, class... Args>
static inline void gcm_launch_dep(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, void* stream,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#define GCM_LAUNCH_DEP(kern, grid, block, smem, stream, ...) \
  gcm_launch_dep(g_gcm_knob[9] != 3, kern, (grid), (block), (smem), (stream), __VA_ARGS__)
#endif

// geom.cu: remembers the newest non-zero status of the calling thread (gcm_last_status) and returns it unchanged
int gcm_set_status(int st);
#define GCM_CHECK_LAUNCH()                                        \
  do {                                                            \
    cudaError_t e_ = cudaGetLastError();                          \
    if (e_ != cudaSuccess) return gcm_set_status((int)e_);        \
  } while (0)
#define GCM_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e_ = (call);                                      \
    if (e_ != cudaSuccess) return gcm_set_status((int)e_);        \
  } while (0)
#define GCM_REQUIRE(cond, code)                       \
  do {                                                \
    if (!(cond)) return gcm_set_status((code));       \
  } while (0)

// physical constants, SI (reference constants.py:16-48)
#define GCM_RD 287.0
#define GCM_CP 1004.0
#define GCM_KAPPA (287.0 / 1004.0)
#define GCM_P0 100000.0
#define GCM_G 9.8

#define GCM_MAX_RADIX_PASSES 16
#define GCM_MAXLC 24

// Stockham mixed-radix plan for one row of length n (see fft_rows.h)
struct GcmFftPlan {
  int n;
  int npass;
  int radix[GCM_MAX_RADIX_PASSES];
  // in-place transform (fft_inplace.h): per stage the butterfly stride n_s / r_s and ceil(2^32 / x) multipliers
  // that turn the two index divisions of a stage into one multiply-high each (operands are below 2^16)
  int stride[GCM_MAX_RADIX_PASSES];
  unsigned magic_stride[GCM_MAX_RADIX_PASSES];
  unsigned magic_nbf[GCM_MAX_RADIX_PASSES];
  int blk[GCM_MAX_RADIX_PASSES];    // block length n_s entering stage s (n_0 = n)
  int nbf[GCM_MAX_RADIX_PASSES];    // butterflies per row: n / r_s
  int tstep[GCM_MAX_RADIX_PASSES];  // twiddle table step n / n_s
  int twoff[GCM_MAX_RADIX_PASSES];  // start of stage s in the per-stage twiddle table (GcmGeomDev::tws)
};

// floor(x / d) for x < 2^16, d < 2^16, with m = ceil(2^32 / d); m == 0 encodes d == 1
__host__ __device__ __forceinline__ unsigned gcm_magic(unsigned d) {
  return d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d);
}
__device__ __forceinline__ int gcm_fastdiv(int x, unsigned m) {
#ifdef GCM_EMU
  return m ? (int)(((unsigned long long)(unsigned)x * m) >> 32) : x;
#else
  return m ? (int)__umulhi((unsigned)x, m) : x;
#endif
}

// 1 / b to about 1 ulp without the special-case branches of the IEEE division sequence: hardware seed
// (rcp.approx.ftz.f64, 2^-23) + two Newton steps.  Only for the fast kernels (pe25_fast.cu).
__device__ __forceinline__ double gcm_rcp(double b) {
#ifdef GCM_EMU
  return 1.0 / b;
#else
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
#endif
}

// ask L2 for the 128-byte line at p (no register, no scoreboard): used to start a later phase's rows early
__device__ __forceinline__ void gcm_prefetch_l2(const void* p) {
#ifndef GCM_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

__device__ __forceinline__ void gcm_prefetch_l1(const void* p) {
#ifndef GCM_EMU
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// 8-byte asynchronous copy global -> shared (LDGSTS): no register, no scoreboard; completion by group
__device__ __forceinline__ void gcm_cp_async8(double* smem_dst, const double* gmem_src) {
#ifdef GCM_EMU
  *smem_dst = *gmem_src;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
// 16-byte form (two doubles; both addresses 16-byte aligned)
__device__ __forceinline__ void gcm_cp_async16(double* smem_dst, const double* gmem_src) {
#ifdef GCM_EMU
  smem_dst[0] = gmem_src[0];
  smem_dst[1] = gmem_src[1];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void gcm_cp_async_commit() {
#ifndef GCM_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
// wait until at most N of this thread's committed groups are still in flight
template <int N>
__device__ __forceinline__ void gcm_cp_async_wait() {
#ifndef GCM_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// be scheduled before its predecessor in the stream has drained; gcm_pdl_wait() blocks until every prerequisite grid
// has completed and its writes are visible, so it must precede the first global access of a kernel; both are no-ops in
// a kernel launched the ordinary way.  gcm_pdl_trigger() lets the dependents of THIS grid start being scheduled.
__device__ __forceinline__ void gcm_pdl_wait() {
#ifndef GCM_EMU
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void gcm_pdl_trigger() {
#ifndef GCM_EMU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// Asynchronous "non-finite seen" watch (replaces the caller-side np.isnan(u).any() poll, matsuno_c_grid.py:184-187):
// one 32-bit counter per device; every step kernel adds 1 per thread that wrote a non-finite value.  Read it with
// gcm_nonfinite_read (a stream-ordered 4-byte copy: no synchronisation of the caller).
unsigned int* gcm_nonfinite_word();  // geom.cu: the current device's counter (NULL if it cannot be allocated)
__device__ __forceinline__ bool gcm_not_finite(double s) { return !(fabs(s) <= 1.7976931348623157e308); }
__device__ __forceinline__ void gcm_flag_nonfinite(unsigned int* flag, bool bad) {
  if (bad && flag) atomicAdd(flag, 1u);
}

// device-resident geometry tables, passed to kernels by value
struct GcmGeomDev {
  int H, W, L;
  int wrap_j;
  int row_lo, row_hi;
  int zero_v_row, zero_v_row2;
  double dy, ptop;
  const double* sig;
  const double* dsig;
  const double* sigb;
  const double* sigt;
  const double* dx_j;
  const double* dx_h;
  const double* hmap;
  const double* smmz;   // [H][W/2+1]
  const double2* tw;    // [W]  exp(-2 pi i m / W)
  // tables of the ALU-lean step kernels (pe25_fast.cu): reciprocals instead of divides, sig^kappa, and the
  // digit-reversal map of the in-place FFT (fft_inplace.h)
  const double* rdx_j;  // [H]  1 / dx_j
  const double* rdx_h;  // [H]  1 / dx_h
  const double* rdsig;  // [L]  1 / dsig
  const double* sigkap; // [L]  sig^kappa
  const int* kperm;     // [W]  wavenumber held at position p after the forward DIF transform
  const double* smmzp;  // [H][W]  smmz[j][min(k, W-k)] / W at the position p that holds wavenumber k = kperm[p]
                        //         after the forward in-place transform (the 1/n of numpy's irfft folded in)
  const double2* tws;   // per-stage twiddles of the in-place transform, contiguous in the butterfly index:
                        //         tws[twoff[s] + (m-1) stride_s + q] = exp(-2 pi i q m / n_s)
  double rdy;           // 1 / dy
  unsigned int* nonfinite;  // the device's non-finite counter (gcm_nonfinite_word)
  int pdl_early;        // 1 = kernels trigger their dependents at entry (set per launch from tuning knob 9)
  // per-layer tables by value (kernel parameters live in the constant bank: no load instruction) when L <= 16
  double c_sig[GCM_MAXLC], c_dsig[GCM_MAXLC], c_sigb[GCM_MAXLC], c_sigt[GCM_MAXLC], c_rdsig[GCM_MAXLC],
      c_sigkap[GCM_MAXLC];
  GcmFftPlan plan;
};

// opt-in terms of the 2.5-D half step (pe25_extras.cu; gcm_pe25_set_options): all off by default
struct GcmExtras {
  int coriolis, limit_q, limit_t;
  double nu;
  const double* cor_u;  // [H] 2 sin(lat) w at the u rows (dynamics.py:91)
  const double* cor_v;  // [H] at the v rows (dynamics.py:92)
};

struct gcm_geom {
  GcmGeomDev d;
  void* d_block;  // one device allocation holding every table
  GcmExtras x;    // opt-in terms; x.cor_u / x.cor_v point into d_cor
  void* d_cor;
  // side stream + fork/join events: the two independent kernel chains of a half step's row phase run side by side
  // (pe25_fast.cu); created on first use, one user thread per geometry
  void* aux_stream;
  void* ev_fork;
  void* ev_join;
  // two consecutive Matsuno steps (ping-pong A -> B -> A) captured as a CUDA graph for multi-step calls (pe25.cu):
  // a small cache keyed by the buffers, dt, member count and the tuning epoch
  void* cap_stream;
  struct {
    unsigned long long key[20];
    void* exec;
  } graph[2];
  int graph_next;
};
extern int g_gcm_tuning_epoch;  // bumped by gcm_tuning_knob / gcm_pe25_select_path: cached graphs go stale
// lazily creates the side stream and events; returns a cudaError_t / GCM_OK
int gcm_geom_aux(const gcm_geom* g, void** stream, void** ev_fork, void** ev_join);

// rows of one launch: n1 rows from a, then n2 rows from c.  One segment covers a whole band; the rows next to the
// halos (first owned row + last owned rows) form a two-segment launch once the halo exchange has landed.
struct GcmRowSeg {
  int a, n1, c, n2;
};
__host__ __device__ __forceinline__ int gcm_seg_row(const GcmRowSeg& s, int r) {
  return r < s.n1 ? s.a + r : s.c + (r - s.n1);
}

static inline bool gcm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// neighbour row in j: periodic like np.roll when the geometry stores the whole grid, plain offset
// when it stores a band with halo rows
__host__ __device__ __forceinline__ int gcm_row(int j, int d, int H, int wrap) {
  int r = j + d;
  if (wrap) {
    if (r < 0) r += H;
    if (r >= H) r -= H;
  }
  return r;
}
__host__ __device__ __forceinline__ int gcm_ip(int i, int W) { return i + 1 == W ? 0 : i + 1; }
__host__ __device__ __forceinline__ int gcm_im(int i, int W) { return i == 0 ? W - 1 : i - 1; }

int gcm_fft_make_plan(int n, GcmFftPlan* plan);

static inline bool gcm_extras_on(const gcm_geom* g) {
  return g->x.coriolis || g->x.limit_q || g->x.limit_t || g->x.nu != 0.0;
}
// rows j - 2 ... j + 2 of every owned row are stored: whole grid, or a band with two halo rows on either side
static inline bool gcm_extras_rows_ok(const gcm_geom* g) {
  return g->d.wrap_j || (g->d.row_lo >= 2 && g->d.H - g->d.row_hi >= 2);
}
// pe25_extras.cu: adds the opt-in terms to the freshly written `out` of a half step (whole-grid geometry)
int gcm_pe25_extras_apply(const gcm_geom* g, const gcm_state* star, const gcm_state* out, const double* spu,
                          const double* pn, double dt, int nbatch, void* stream);
