// comm.cu -- latitude-band Matsuno loop with the halo exchange over NCCL (NVLink 5 / NVSwitch).
//
// The reference is one process: every j-shift is np.roll over the whole array (coordinates_3d.py:43-48).  Here rank r
// owns a band of rows (all i, all k) with one halo row to the north and two to the south; ranks form a ring (the
// roll is periodic across the pole).  Per half step the band
//   side stream : packs its first two / last owned rows, ncclSend/ncclRecv with both neighbours in one group,
//                 unpacks into the halo rows, then computes the three rows next to the halos
//                 (gcm_pe25_half_step_rows, two-segment launches);
//   main stream : meanwhile computes the interior rows (they read owned rows only) and joins the side stream.
// With 2 + 4 halo rows the band exchanges ONCE per step and recomputes the three predictor rows its corrector reads
// across the edges.  The whole loop over steps runs here, so the host cost per step is a dozen launches and one or
// two NCCL groups.
//
// NCCL is bound at run time (dlopen of the libnccl the process already uses -- torch's -- else the system one): the
// library has no link-time dependency on it, so single-GPU users and the CPU-side ABI check never need NCCL.
#include <stdlib.h>
#include <string.h>

#include "gcm_common.h"
#include "prof.h"

#ifndef GCM_EMU
#include <dlfcn.h>
#endif

extern "C" int gcm_pe25_half_step(const gcm_geom*, const gcm_state*, const gcm_state*, const gcm_state*, double, int,
                                  void*, size_t, void*);
extern "C" int gcm_pe25_half_step_rows(const gcm_geom*, const gcm_state*, const gcm_state*, const gcm_state*, double, int,
                                       void*, size_t, const int*, const int*, void*);

#define GCM_HALO_N 1
#define GCM_HALO_S 2
#define GCM_ENCCL_BASE 1000000  // status = GCM_ENCCL_BASE + ncclResult_t

typedef struct {
  char internal[128];
} gcm_nccl_id;
typedef void* gcm_nccl_comm;

struct GcmNccl {
  void* handle;
  int (*GetUniqueId)(gcm_nccl_id*);
  int (*CommInitRank)(gcm_nccl_comm*, int, gcm_nccl_id, int);
  int (*CommDestroy)(gcm_nccl_comm);
  int (*Send)(const void*, size_t, int, int, gcm_nccl_comm, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, gcm_nccl_comm, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
};
static GcmNccl g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

static int gcm_nccl_load() {
#ifdef GCM_EMU
  return GCM_EUNSUP;
#else
  if (g_nccl.handle) return GCM_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int pass = 0; pass < 2 && !h; ++pass)  // first the copy already in the process (torch's), then a fresh load
    for (int n = 0; n < 2 && !h; ++n) h = dlopen(names[n], RTLD_NOW | (pass == 0 ? RTLD_NOLOAD : 0));
  GCM_REQUIRE(h, GCM_EUNSUP);
  GcmNccl t;
  t.handle = h;
  t.GetUniqueId = (int (*)(gcm_nccl_id*))dlsym(h, "ncclGetUniqueId");
  t.CommInitRank = (int (*)(gcm_nccl_comm*, int, gcm_nccl_id, int))dlsym(h, "ncclCommInitRank");
  t.CommDestroy = (int (*)(gcm_nccl_comm))dlsym(h, "ncclCommDestroy");
  t.Send = (int (*)(const void*, size_t, int, int, gcm_nccl_comm, cudaStream_t))dlsym(h, "ncclSend");
  t.Recv = (int (*)(void*, size_t, int, int, gcm_nccl_comm, cudaStream_t))dlsym(h, "ncclRecv");
  t.GroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
  t.GroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
  GCM_REQUIRE(t.GetUniqueId && t.CommInitRank && t.CommDestroy && t.Send && t.Recv && t.GroupStart && t.GroupEnd,
              GCM_EUNSUP);
  g_nccl = t;
  return GCM_OK;
#endif
}

#define GCM_NCCL(call)                                             \
  do {                                                             \
    int r_ = (call);                                               \
    if (r_ != 0) return gcm_set_status(GCM_ENCCL_BASE + r_);       \
  } while (0)

struct gcm_comm {
  int nranks, rank;
  gcm_nccl_comm comm;
  cudaStream_t stream;          // halo traffic runs here, beside the caller's stream
  cudaEvent_t ev_ready, ev_halo, ev_rint;
  double* buf;                  // send_n | send_s | recv_s | recv_n
  size_t buf_doubles;
};

extern "C" int gcm_comm_unique_id(unsigned char* out128) {
  GCM_REQUIRE(out128, GCM_ENULL);
  int st = gcm_nccl_load();
  if (st) return st;
  gcm_nccl_id id;
  GCM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return GCM_OK;
}

extern "C" int gcm_comm_destroy(gcm_comm* c);
extern "C" int gcm_comm_create(int nranks, int rank, const unsigned char* id128, gcm_comm** out) {
  GCM_REQUIRE(out, GCM_ENULL);
  GCM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, GCM_ESHAPE);
  gcm_comm* c = (gcm_comm*)calloc(1, sizeof(gcm_comm));
  GCM_REQUIRE(c, (int)cudaErrorMemoryAllocation);
  c->nranks = nranks;
  c->rank = rank;
#ifndef GCM_EMU
  if (nranks > 1) {
    int st = id128 ? gcm_nccl_load() : GCM_ENULL;
    if (st) { free(c); return st; }
    gcm_nccl_id id;
    memcpy(id.internal, id128, 128);
    int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (r != 0) { c->comm = nullptr; gcm_comm_destroy(c); return gcm_set_status(GCM_ENCCL_BASE + r); }
  }
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_rint, cudaEventDisableTiming);
  if (e != cudaSuccess) { gcm_comm_destroy(c); return gcm_set_status((int)e); }  // frees whatever exists so far
#else
  (void)id128;
  GCM_REQUIRE(nranks == 1, GCM_EUNSUP);
#endif
  *out = c;
  return GCM_OK;
}

extern "C" int gcm_comm_destroy(gcm_comm* c) {
  if (!c) return GCM_OK;
#ifndef GCM_EMU
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_halo) cudaEventDestroy(c->ev_halo);
  if (c->ev_rint) cudaEventDestroy(c->ev_rint);
#endif
  if (c->buf) cudaFree(c->buf);
  free(c);
  return GCM_OK;
}

extern "C" size_t gcm_halo_buffer_doubles(const gcm_geom* g, int nrows);
extern "C" int gcm_halo_pack(const gcm_geom*, const gcm_state*, int, int, double*, void*);
extern "C" int gcm_halo_unpack(const gcm_geom*, const gcm_state*, int, int, const double*, void*);
extern "C" int gcm_halo_copy_rows(const gcm_geom*, const gcm_state*, int, const gcm_state*, int, int, void*);

// both halo messages of a band in ONE launch: rows [a.row0, +a.nrows) <-> a.buf and rows [b.row0, +b.nrows) <-> b.buf
// (unpack = 0: state -> buffers, 1: buffers -> state).  Buffer layout per job: [p rows][u][v][t][q], each [layer][row][i].
struct GcmHaloJob {
  int row0, nrows;
  double* buf;
};
__global__ void band_halo2_kernel(gcm_state s, int H, int W, int L, GcmHaloJob a, GcmHaloJob b, int unpack) {
  double* fld[5] = {s.p, s.u, s.v, s.t, s.q};
  const size_t na = (size_t)a.nrows * W * (1 + 4 * (size_t)L), nb = (size_t)b.nrows * W * (1 + 4 * (size_t)L);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < na + nb; e += (size_t)gridDim.x * blockDim.x) {
    const GcmHaloJob& jb = e < na ? a : b;
    const size_t r0 = e < na ? e : e - na;
    const size_t per_p = (size_t)jb.nrows * W, per_3 = (size_t)L * per_p;
    int f;
    size_t r;
    if (r0 < per_p) { f = 0; r = r0; } else { f = 1 + (int)((r0 - per_p) / per_3); r = (r0 - per_p) % per_3; }
    const int i = (int)(r % W), row = (int)((r / W) % jb.nrows), k = (int)(r / per_p);
    const size_t in_state = ((size_t)k * H + jb.row0 + row) * W + i;
    if (unpack) fld[f][in_state] = jb.buf[r0];
    else jb.buf[r0] = fld[f][in_state];
  }
}

// fill the halo rows of `s` (hn north, hs south) from the ring neighbours, on stream `q`
static int band_exchange(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, cudaStream_t q) {
  GcmProfScope ps(GCM_K_HALO, q);
  const int lo = g->d.row_lo, hi = g->d.row_hi;
  int st;
  if (c->nranks == 1) {  // the ring closes on the band itself
    if ((st = gcm_halo_copy_rows(g, s, lo, s, hi, hs, q))) return st;
    return gcm_halo_copy_rows(g, s, hi - hn, s, lo - hn, hn, q);
  }
#ifdef GCM_EMU
  return GCM_EUNSUP;
#else
  const size_t ns = gcm_halo_buffer_doubles(g, hs), nn = gcm_halo_buffer_doubles(g, hn);
  if (c->buf_doubles < 2 * (ns + nn)) {
    if (c->buf) cudaFree(c->buf);
    c->buf = nullptr;
    GCM_CUDA(cudaMalloc((void**)&c->buf, 2 * (ns + nn) * sizeof(double)));
    c->buf_doubles = 2 * (ns + nn);
  }
  double *send_n = c->buf, *send_s = send_n + ns, *recv_s = send_s + nn, *recv_n = recv_s + ns;
  const int north = (c->rank + c->nranks - 1) % c->nranks, south = (c->rank + 1) % c->nranks;
  const int H = g->d.H, W = g->d.W, L = g->d.L;
  const unsigned nblk = (unsigned)((ns + nn + 255) / 256 < 148 * 4 ? (ns + nn + 255) / 256 : 148 * 4);
  // first owned rows -> north's south halo; last owned rows -> south's north halo
  band_halo2_kernel<<<nblk, 256, 0, q>>>(*s, H, W, L, GcmHaloJob{lo, hs, send_n}, GcmHaloJob{hi - hn, hn, send_s}, 0);
  GCM_CHECK_LAUNCH();
  GCM_NCCL(g_nccl.GroupStart());
  GCM_NCCL(g_nccl.Send(send_n, ns, 8 /* ncclFloat64 */, north, c->comm, q));
  GCM_NCCL(g_nccl.Send(send_s, nn, 8, south, c->comm, q));
  GCM_NCCL(g_nccl.Recv(recv_s, ns, 8, south, c->comm, q));
  GCM_NCCL(g_nccl.Recv(recv_n, nn, 8, north, c->comm, q));
  GCM_NCCL(g_nccl.GroupEnd());
  band_halo2_kernel<<<nblk, 256, 0, q>>>(*s, H, W, L, GcmHaloJob{hi, hs, recv_s}, GcmHaloJob{lo - hn, hn, recv_n}, 1);
  GCM_CHECK_LAUNCH();
  return GCM_OK;
#endif
}

// out = base + dt F(star) on the band, halos of `star` exchanged first; interior rows overlap the exchange
static int band_half_step(const gcm_geom* g, gcm_comm* c, const gcm_state* base, const gcm_state* star,
                          const gcm_state* out, double dt, int overlap, void* ws, size_t ws_bytes, cudaStream_t main) {
  const int lo = g->d.row_lo, hi = g->d.row_hi, n = hi - lo;
  int st;
#ifndef GCM_EMU
  if (overlap && n >= 4) {
    // main stream: row phase and update of the interior rows (they read owned rows only);
    // side stream: halo exchange, then the rows next to the halos (two-segment launches), beside the interior.
    const int none[4] = {0, 0, 0, 0};
    const int ri[4] = {lo + 1, n - 2, 0, 0}, ui[4] = {lo + 1, n - 3, 0, 0};
    const int rb[4] = {lo, 1, hi - 1, 2}, ub[4] = {lo, 1, hi - 2, 2};
    GCM_CUDA(cudaEventRecord(c->ev_ready, main));  // `star` is complete on the main stream
    GCM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_ready, 0));
    if ((st = band_exchange(g, c, star, GCM_HALO_N, GCM_HALO_S, c->stream))) return st;
    st = gcm_pe25_half_step_rows(g, base, star, out, dt, 1, ws, ws_bytes, ri, none, main);
    if (st == GCM_OK) {
      GCM_CUDA(cudaEventRecord(c->ev_rint, main));  // row phase of the interior is done
      if ((st = gcm_pe25_half_step_rows(g, base, star, out, dt, 1, ws, ws_bytes, none, ui, main))) return st;
      if ((st = gcm_pe25_half_step_rows(g, base, star, out, dt, 1, ws, ws_bytes, rb, none, c->stream))) return st;
      GCM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_rint, 0));  // update of rows lo and hi-2 reads rows lo+1, hi-2
      if ((st = gcm_pe25_half_step_rows(g, base, star, out, dt, 1, ws, ws_bytes, none, ub, c->stream))) return st;
      GCM_CUDA(cudaEventRecord(c->ev_halo, c->stream));
      GCM_CUDA(cudaStreamWaitEvent(main, c->ev_halo, 0));
      return GCM_OK;
    }
    if (st != GCM_EUNSUP) return st;
    GCM_CUDA(cudaEventRecord(c->ev_halo, c->stream));
    GCM_CUDA(cudaStreamWaitEvent(main, c->ev_halo, 0));  // no segmented kernels for this geometry: whole band now
    return gcm_pe25_half_step(g, base, star, out, dt, 1, ws, ws_bytes, main);
  }
#else
  (void)overlap;
  (void)n;
#endif
  if ((st = band_exchange(g, c, star, GCM_HALO_N, GCM_HALO_S, main))) return st;
  return gcm_pe25_half_step(g, base, star, out, dt, 1, ws, ws_bytes, main);
}

// dynamics.matsuno_timestep (dynamics.py:230-237) `nsteps` times on one latitude band.  `cur` holds the band with
// halo rows (any content); after the call the newest state is in `nxt` if nsteps is odd, else in `cur` (the two
// alternate); `star` is scratch.  Halos of the result are NOT filled.
extern "C" int gcm_band_matsuno_step(const gcm_geom* g, gcm_comm* c, const gcm_state* cur, const gcm_state* star,
                                     const gcm_state* nxt, double dt, int nsteps, int overlap, void* ws, size_t ws_bytes,
                                     void* stream) {
  GCM_REQUIRE(g && c && cur && star && nxt && ws, GCM_ENULL);
  GCM_REQUIRE(nsteps > 0, GCM_ESHAPE);
  GCM_REQUIRE(!g->d.wrap_j, GCM_EUNSUP);
  const int hn = g->d.row_lo, hs = g->d.H - g->d.row_hi, lo = g->d.row_lo, n = g->d.row_hi - g->d.row_lo;
  // opt-in terms (gcm_pe25_set_options): they reach j - 2 ... j + 2, so the band carries two halo rows on either side
  // and both states are exchanged, two rows each way, before a whole-band half step (no row-segment schedule)
  const bool ext = gcm_extras_on(g);
  const bool wide = !ext && hn == 2 * GCM_HALO_N && hs == 2 * GCM_HALO_S;
  GCM_REQUIRE(ext ? (hn == 2 && hs == 2) : (wide || (hn == GCM_HALO_N && hs == GCM_HALO_S)), GCM_ESHAPE);
  GCM_REQUIRE(n >= hs, GCM_ESHAPE);
  cudaStream_t main = (cudaStream_t)stream;
  const gcm_state *a = cur, *b = nxt;
  int st;
  for (int s = 0; s < nsteps; ++s) {
    if (ext) {
      if ((st = band_exchange(g, c, a, hn, hs, main))) return st;
      if ((st = gcm_pe25_half_step(g, a, a, star, dt, 1, ws, ws_bytes, main))) return st;           // dynamics.py:231
      if ((st = band_exchange(g, c, star, hn, hs, main))) return st;
      if ((st = gcm_pe25_half_step(g, a, star, b, dt, 1, ws, ws_bytes, main))) return st;           // dynamics.py:234
    } else if (wide) {
      // ONE exchange per step: with 2 + 4 halo rows of the base state the band computes the predictor also on the
      // three rows its corrector reads across the band edges (the neighbours compute the same values from the same
      // inputs with the same kernels), so the star state needs no exchange.  Halves the latency-bound messages.
      if ((st = band_exchange(g, c, a, hn, hs, main))) return st;
      const int rp[4] = {lo - 1, n + 4, 0, 0}, up[4] = {lo - 1, n + 3, 0, 0};
      const int rc[4] = {lo, n + 1, 0, 0}, uc[4] = {lo, n, 0, 0};
      if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, rp, up, main))) return st;  // dynamics.py:231
      if ((st = gcm_pe25_half_step_rows(g, a, star, b, dt, 1, ws, ws_bytes, rc, uc, main))) return st;  // dynamics.py:234
    } else {
      if ((st = band_half_step(g, c, a, a, star, dt, overlap, ws, ws_bytes, main))) return st;     // dynamics.py:231
      if ((st = band_half_step(g, c, a, star, b, dt, overlap, ws, ws_bytes, main))) return st;     // dynamics.py:234
    }
    const gcm_state* t = a;
    a = b;
    b = t;
  }
  return GCM_OK;
}
