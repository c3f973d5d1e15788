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

extern "C" size_t gcm_halo_buffer_doubles(const gcm_geom* g, int nrows);

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
  // Peer-memory halo exchange (NVLink peer stores instead of NCCL): every rank owns a MAILBOX (cudaMalloc, exported
  // with cudaIpcGetMemHandle); the neighbours map it and write my halo rows straight into it, then raise a flag in it.
  unsigned char* mbox;          // my mailbox: GcmMboxHeader | slot 0 | slot 1, slot = rows from north | rows from south
  size_t mbox_bytes;
  size_t part_n, part_s;        // doubles per slot part: halo rows from the north / from the south neighbour
  int mb_hn, mb_hs, mb_W, mb_L; // halo layout the mailbox was sized for
  unsigned char* peer_n;        // the north / south neighbour's mailbox as mapped into this process
  unsigned char* peer_s;
  int peer_mode;                // 0 = NCCL, 1 = mapped through CUDA IPC, 2 = plain pointers of the same process (tests)
  int push_blocks;              // grid of the push kernel (the ticket of its last block closes the message)
};

// first 256 bytes of a mailbox
struct GcmMboxHeader {
  unsigned long long flag_from_n;  // sequence number of the newest complete message from the north neighbour
  unsigned long long flag_from_s;  //                                              ... from the south neighbour
  unsigned long long seq;          // messages this rank has sent (local)
  unsigned int ticket;             // finished blocks of the running push kernel (local)
  unsigned int timeouts;           // pull kernels that gave up waiting (local; reported by gcm_comm_peer_status)
};
#define GCM_MBOX_HDR 256

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
  if (nranks > 1 && id128) {  // id128 == NULL: no NCCL communicator, the ring runs on peer mailboxes only
    int st = gcm_nccl_load();
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
  (void)id128;  // no NCCL on the emulator build: ranks > 1 exist only for the same-process peer-mailbox tests
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
#ifndef GCM_EMU
  if (c->peer_mode == 1) {
    if (c->peer_n) cudaIpcCloseMemHandle(c->peer_n);
    if (c->peer_s && c->peer_s != c->peer_n) cudaIpcCloseMemHandle(c->peer_s);
  }
#endif
  if (c->mbox) cudaFree(c->mbox);
  free(c);
  return GCM_OK;
}

// ---- peer-memory mailbox ----------------------------------------------------------------------------------------
// Allocates this rank's mailbox for the halo layout of band geometry `g` and writes its 64-byte CUDA IPC handle to
// handle64 (zeros on the emulator build).  The caller ships the handles over its own control plane
// (bands.BandStepper: one all_gather) and calls gcm_comm_peer_connect.
extern "C" int gcm_comm_peer_setup(gcm_comm* c, const gcm_geom* g, unsigned char* handle64) {
  GCM_REQUIRE(c && g && handle64, GCM_ENULL);
  GCM_REQUIRE(!g->d.wrap_j && !c->mbox, GCM_EUNSUP);
  const int hn = g->d.row_lo, hs = g->d.H - g->d.row_hi;
  c->mb_hn = hn; c->mb_hs = hs; c->mb_W = g->d.W; c->mb_L = g->d.L;
  c->part_n = gcm_halo_buffer_doubles(g, hn);
  c->part_s = gcm_halo_buffer_doubles(g, hs);
  c->mbox_bytes = GCM_MBOX_HDR + 2 * (c->part_n + c->part_s) * sizeof(double);
  GCM_CUDA(cudaMalloc((void**)&c->mbox, c->mbox_bytes));
  GCM_CUDA(cudaMemset(c->mbox, 0, c->mbox_bytes));
  memset(handle64, 0, 64);
#ifndef GCM_EMU
  cudaIpcMemHandle_t h;
  GCM_CUDA(cudaIpcGetMemHandle(&h, c->mbox));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle64, &h, 64);
#endif
  return GCM_OK;
}

// Maps the neighbours' mailboxes (IPC handles from their gcm_comm_peer_setup).  same_process != 0: the two arguments
// are gcm_comm* of this process instead (single-process tests of the protocol: ranks stepped one after the other).
extern "C" int gcm_comm_peer_connect(gcm_comm* c, const void* north, const void* south, int same_process) {
  GCM_REQUIRE(c && north && south && c->mbox, GCM_ENULL);
  GCM_REQUIRE(c->peer_mode == 0, GCM_EUNSUP);
  if (same_process) {
    const gcm_comm* cn = (const gcm_comm*)north;
    const gcm_comm* cs = (const gcm_comm*)south;
    GCM_REQUIRE(cn->mbox && cs->mbox && cn->mbox_bytes == c->mbox_bytes && cs->mbox_bytes == c->mbox_bytes, GCM_ESHAPE);
    c->peer_n = cn->mbox;
    c->peer_s = cs->mbox;
    c->peer_mode = 2;
    return GCM_OK;
  }
#ifdef GCM_EMU
  return GCM_EUNSUP;
#else
  cudaIpcMemHandle_t hn_, hs_;
  memcpy(&hn_, north, 64);
  memcpy(&hs_, south, 64);
  void *pn = nullptr, *ps = nullptr;
  GCM_CUDA(cudaIpcOpenMemHandle(&pn, hn_, cudaIpcMemLazyEnablePeerAccess));
  if (memcmp(north, south, 64) == 0) {
    ps = pn;  // two ranks: both neighbours are the same rank; a handle is mapped once per process
  } else {
    cudaError_t e = cudaIpcOpenMemHandle(&ps, hs_, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaIpcCloseMemHandle(pn); return gcm_set_status((int)e); }
  }
  c->peer_n = (unsigned char*)pn;
  c->peer_s = (unsigned char*)ps;
  c->peer_mode = 1;
  return GCM_OK;
#endif
}

// 0 = NCCL, 1 = CUDA IPC peers, 2 = same-process peers; *timeouts (optional) = pull kernels that gave up waiting for a
// neighbour (synchronises the device to read the counter)
extern "C" int gcm_comm_peer_status(gcm_comm* c, unsigned int* timeouts) {
  GCM_REQUIRE(c, GCM_ENULL);
  if (timeouts) {
    *timeouts = 0;
    if (c->mbox) {
      GcmMboxHeader h;
      GCM_CUDA(cudaMemcpy(&h, c->mbox, sizeof(h), cudaMemcpyDeviceToHost));
      *timeouts = h.timeouts;
    }
  }
  return c->peer_mode;
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


// ---- peer-memory exchange kernels ---------------------------------------------------------------------------------
__device__ __forceinline__ void band_halo_copy(double* const* fld, int H, int W, int L, const GcmHaloJob& a,
                                               const GcmHaloJob& b, int unpack) {
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

__device__ __forceinline__ void gcm_fence_system() {
#ifndef GCM_EMU
  __threadfence_system();
#else
  __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}
__device__ __forceinline__ unsigned long long gcm_load_sys(const unsigned long long* p) {
#ifndef GCM_EMU
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#else
  return __atomic_load_n(p, __ATOMIC_SEQ_CST);
#endif
}
__device__ __forceinline__ void gcm_store_sys(unsigned long long* p, unsigned long long v) {
#ifndef GCM_EMU
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
  __atomic_store_n(p, v, __ATOMIC_SEQ_CST);
#endif
}

// PUSH: my first owned rows -> the north neighbour's mailbox (its halo rows to the south), my last owned rows -> the
// south neighbour's (its halo rows to the north), slot = parity of the message number.  The block that finishes last
// publishes the message: system-wide fence, then the message number into both neighbours' flags.
__global__ void band_push_kernel(gcm_state s, int H, int W, int L, GcmHaloJob to_n, GcmHaloJob to_s, size_t slot_doubles,
                                 GcmMboxHeader* mine, GcmMboxHeader* north, GcmMboxHeader* south) {
  gcm_pdl_wait();  // launched as a programmatic dependent of the update kernel: the rows it sends are complete here
  __shared__ unsigned long long seq_s;
  __shared__ int last_s;
  if (threadIdx.x == 0) seq_s = mine->seq + 1;  // written only by the closing block of the previous push
  __syncthreads();
  const unsigned long long seq = seq_s;
  to_n.buf += (seq & 1) * slot_doubles;
  to_s.buf += (seq & 1) * slot_doubles;
  double* fld[5] = {s.p, s.u, s.v, s.t, s.q};
  band_halo_copy(fld, H, W, L, to_n, to_s, 0);
  gcm_fence_system();  // my peer stores are performed before the ticket below
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&mine->ticket, 1u);
    last_s = t + 1 == gridDim.x;
  }
  __syncthreads();
  if (last_s && threadIdx.x == 0) {
    gcm_fence_system();  // every block's stores (ordered before its ticket) before the flags
    mine->ticket = 0;
    mine->seq = seq;
    gcm_store_sys(&north->flag_from_s, seq);
    gcm_store_sys(&south->flag_from_n, seq);
  }
}

// PULL: wait until both neighbours have published message `seq` (my own count: every rank sends the same number of
// messages), then copy my mailbox slot into my halo rows.  A bounded wait: a neighbour that never answers raises
// `timeouts` instead of hanging the GPU.
__global__ void band_pull_kernel(gcm_state s, int H, int W, int L, GcmHaloJob from_s, GcmHaloJob from_n,
                                 size_t slot_doubles, GcmMboxHeader* mine) {
  gcm_pdl_wait();  // the push of this rank (same stream) has bumped mine->seq
  __shared__ unsigned long long seq_s;
  if (threadIdx.x == 0) {
    const unsigned long long seq = mine->seq;
    bool ok = false;
#ifndef GCM_EMU
    const long long t0 = clock64();
    for (;;) {
      ok = gcm_load_sys(&mine->flag_from_n) >= seq && gcm_load_sys(&mine->flag_from_s) >= seq;
      if (ok || clock64() - t0 > 4000000000ll) break;  // ~2 s
      __nanosleep(64);
    }
#else
    ok = gcm_load_sys(&mine->flag_from_n) >= seq && gcm_load_sys(&mine->flag_from_s) >= seq;
#endif
    if (!ok && blockIdx.x == 0) atomicAdd(&mine->timeouts, 1u);
    gcm_fence_system();  // the neighbours' rows were performed before their flags
    seq_s = seq;
  }
  __syncthreads();
  from_s.buf += (seq_s & 1) * slot_doubles;
  from_n.buf += (seq_s & 1) * slot_doubles;
  double* fld[5] = {s.p, s.u, s.v, s.t, s.q};
  band_halo_copy(fld, H, W, L, from_s, from_n, 1);
}

// halo rows of `s` through the peer mailboxes (hn north, hs south rows; at most the layout the mailbox was sized for)
// phase: 1 = push only, 2 = pull only, 3 = both
static int band_exchange_peer(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, cudaStream_t q,
                              int phase = 3) {
  GCM_REQUIRE(c->peer_mode && c->mbox && c->peer_n && c->peer_s, GCM_EUNSUP);
  GCM_REQUIRE(hn <= c->mb_hn && hs <= c->mb_hs && g->d.W == c->mb_W && g->d.L == c->mb_L, GCM_ESHAPE);
  const int lo = g->d.row_lo, hi = g->d.row_hi, H = g->d.H, W = g->d.W, L = g->d.L;
  const size_t slot = c->part_n + c->part_s;
  GcmMboxHeader* mine = (GcmMboxHeader*)c->mbox;
  GcmMboxHeader* north = (GcmMboxHeader*)c->peer_n;
  GcmMboxHeader* south = (GcmMboxHeader*)c->peer_s;
  // slot layout: [rows from the north neighbour (part_n)][rows from the south neighbour (part_s)]
  double* n_from_s = (double*)(c->peer_n + GCM_MBOX_HDR) + c->part_n;  // north's "from south" part: my first rows
  double* s_from_n = (double*)(c->peer_s + GCM_MBOX_HDR);              // south's "from north" part: my last rows
  double* my_from_n = (double*)(c->mbox + GCM_MBOX_HDR);
  double* my_from_s = my_from_n + c->part_n;
  const size_t total = gcm_halo_buffer_doubles(g, hs) + gcm_halo_buffer_doubles(g, hn);
  unsigned nblk = (unsigned)((total + 255) / 256);
  nblk = nblk < 1 ? 1 : (nblk > 148 * 2 ? 148 * 2 : nblk);
  if (phase & 1) {
    GCM_LAUNCH_DEP(band_push_kernel, dim3(nblk), dim3(256), 0, q, *s, H, W, L, GcmHaloJob{lo, hs, n_from_s},
                   GcmHaloJob{hi - hn, hn, s_from_n}, slot, mine, north, south);
    GCM_CHECK_LAUNCH();
  }
  if (phase & 2) {
    GCM_LAUNCH_DEP(band_pull_kernel, dim3(nblk), dim3(256), 0, q, *s, H, W, L, GcmHaloJob{hi, hs, my_from_s},
                   GcmHaloJob{lo - hn, hn, my_from_n}, slot, mine);
    GCM_CHECK_LAUNCH();
  }
  return GCM_OK;
}

// The two halves of a peer-mailbox exchange on their own (phase 1 = push my boundary rows to the neighbours, 2 = wait
// for theirs and fill my halo rows): hn / hs halo rows north / south.  A caller that steps several ranks from one
// process (tests) pushes on every rank before it pulls on any.
extern "C" int gcm_band_halo_peer(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, int phase,
                                  void* stream) {
  GCM_REQUIRE(g && c && s && s->p && s->u && s->v && s->t && s->q, GCM_ENULL);
  GCM_REQUIRE(!g->d.wrap_j && hn >= 0 && hs >= 0 && phase >= 1 && phase <= 3, GCM_ESHAPE);
  GCM_REQUIRE(hn <= g->d.row_lo && hs <= g->d.H - g->d.row_hi && hn <= g->d.row_hi - g->d.row_lo &&
                  hs <= g->d.row_hi - g->d.row_lo, GCM_ESHAPE);
  return band_exchange_peer(g, c, s, hn, hs, (cudaStream_t)stream, phase);
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
  if (c->peer_mode) return band_exchange_peer(g, c, s, hn, hs, q);
#ifdef GCM_EMU
  return GCM_EUNSUP;
#else
  GCM_REQUIRE(c->comm, GCM_ENULL);  // neither peer mailboxes nor an NCCL communicator
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

// SURVEY 8(b) names the two halves of an exchange gcm_halo_exchange_begin / _end: begin = push my boundary rows to the
// neighbours' mailboxes (returns at once: the caller may queue interior work behind it on another stream), end = wait
// for theirs and fill my halo rows.  Peer mailboxes only (an NCCL exchange is one grouped call: gcm_band_matsuno_step).
extern "C" int gcm_halo_exchange_begin(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, void* stream) {
  return gcm_band_halo_peer(g, c, s, hn, hs, 1, stream);
}
extern "C" int gcm_halo_exchange_end(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, void* stream) {
  return gcm_band_halo_peer(g, c, s, hn, hs, 2, stream);
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
  const int hn = g->d.row_lo, hs = g->d.H - g->d.row_hi, lo = g->d.row_lo, hi = g->d.row_hi, n = hi - lo;
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
      const int rp[4] = {lo - 1, n + 4, 0, 0}, up[4] = {lo - 1, n + 3, 0, 0};
      const int rc[4] = {lo, n + 1, 0, 0}, uc[4] = {lo, n, 0, 0};
      const int none[4] = {0, 0, 0, 0};
      bool done = false;
#ifndef GCM_EMU
      cudaStream_t side = c->stream;
#else
      cudaStream_t side = main;  // streams are program order on the emulator: the same calls, one after the other
#endif
      // overlap = 1: the exchange of the NEW state runs under the interior rows of the corrector's update.  The comm
      // stream updates the rows the neighbours are waiting for (the first 4 and the last 2 owned rows: one two-segment
      // launch) and pushes them / pulls the neighbours' while the caller's stream updates the other n - 6 rows -- the
      // longest kernel of the step.  One extra small launch per step; the next predictor waits for the halos.  overlap = 2: the exchange runs under the
      // interior rows of the PREDICTOR instead (ten extra small launches per step: measured slower on 8 GPUs, r2j).
      const bool tail = overlap == 1 && n >= 12;
      const bool split = overlap == 2 && n >= 16;
      if (tail && s > 0) {
#ifndef GCM_EMU
        GCM_CUDA(cudaStreamWaitEvent(main, c->ev_halo, 0));  // halos of `a`: exchanged under the previous corrector
#endif
        if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, rp, up, main))) return st;
        done = true;
      } else if (split) {
        // Rows whose stencil stays inside the owned rows need no halo: row phase of rows [lo+1, hi-2], update of rows
        // [lo+1, hi-3] on the caller's stream, while the comm stream exchanges and then computes the rows next to the
        // halos (two-segment launches); the update of rows lo and hi-2 reads the row phase of rows lo+1 and hi-2.
        const int ri[4] = {lo + 1, n - 2, 0, 0}, ui[4] = {lo + 1, n - 3, 0, 0};
        const int rb[4] = {lo - 1, 2, hi - 1, 4}, ub[4] = {lo - 1, 2, hi - 2, 4};
#ifndef GCM_EMU
        GCM_CUDA(cudaEventRecord(c->ev_ready, main));  // `a` is complete on the caller's stream
        GCM_CUDA(cudaStreamWaitEvent(side, c->ev_ready, 0));
#endif
        if ((st = band_exchange(g, c, a, hn, hs, side))) return st;
        st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, ri, none, main);
        if (st == GCM_OK) {
#ifndef GCM_EMU
          GCM_CUDA(cudaEventRecord(c->ev_rint, main));  // row phase of the interior is done
#endif
          if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, none, ui, main))) return st;
          if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, rb, none, side))) return st;
#ifndef GCM_EMU
          GCM_CUDA(cudaStreamWaitEvent(side, c->ev_rint, 0));
#endif
          if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, none, ub, side))) return st;
#ifndef GCM_EMU
          GCM_CUDA(cudaEventRecord(c->ev_halo, side));
          GCM_CUDA(cudaStreamWaitEvent(main, c->ev_halo, 0));
#endif
          done = true;
        } else {
          if (st != GCM_EUNSUP) return st;
#ifndef GCM_EMU
          GCM_CUDA(cudaEventRecord(c->ev_halo, side));
          GCM_CUDA(cudaStreamWaitEvent(main, c->ev_halo, 0));  // no segmented kernels: the whole predictor after the exchange
#endif
          if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, rp, up, main))) return st;
          done = true;
        }
      }
      if (!done) {
        if ((st = band_exchange(g, c, a, hn, hs, main))) return st;
        if ((st = gcm_pe25_half_step_rows(g, a, a, star, dt, 1, ws, ws_bytes, rp, up, main))) return st;  // dynamics.py:231
      }
      if (tail && s + 1 < nsteps) {
        // corrector (dynamics.py:234): row phase of every row on the caller's stream; then, side by side, the update of
        // the n - 6 interior rows there and, on the comm stream, the update of the rows the neighbours need followed by
        // the exchange of the new state
        const int ub2[4] = {lo, hs, hi - hn, hn}, ui2[4] = {lo + hs, n - hs - hn, 0, 0};
        if ((st = gcm_pe25_half_step_rows(g, a, star, b, dt, 1, ws, ws_bytes, rc, none, main))) return st;
#ifndef GCM_EMU
        GCM_CUDA(cudaEventRecord(c->ev_ready, main));  // the row phase of the corrector is complete
        GCM_CUDA(cudaStreamWaitEvent(side, c->ev_ready, 0));
#endif
        if ((st = gcm_pe25_half_step_rows(g, a, star, b, dt, 1, ws, ws_bytes, none, ub2, side))) return st;
        if ((st = band_exchange(g, c, b, hn, hs, side))) return st;
#ifndef GCM_EMU
        GCM_CUDA(cudaEventRecord(c->ev_halo, side));
#endif
        if ((st = gcm_pe25_half_step_rows(g, a, star, b, dt, 1, ws, ws_bytes, none, ui2, main))) return st;
      } else {
        if ((st = gcm_pe25_half_step_rows(g, a, star, b, dt, 1, ws, ws_bytes, rc, uc, main))) return st;  // dynamics.py:234
      }
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
