// host_step.cu -- one Matsuno step for a caller whose state lives in HOST memory (the reference's own situation:
// dynamics.matsuno_timestep takes and returns numpy arrays, dynamics.py:230-237).
//
// Copying the state in, stepping and copying it out one after the other leaves the PCIe link half idle and the
// GPU waiting.  The j-stencil of a full step is short (output rows [r0, r1) need input rows [r0-2, r1+4)), so the grid
// is cut into latitude blocks and pipelined over three streams:
//     copy-in stream  : block b+1 host -> device
//     caller's stream : block b   predictor on rows [r0-1, r1+2), corrector on rows [r0, r1)   (row-segment launches)
//     copy-out stream : block b-1 device -> host
// so both directions of the link run at once and the kernels hide under the copies.  A block recomputes the three
// predictor rows it shares with its neighbours (same kernels, same inputs: identical values), so the result is
// bit-identical to gcm_pe25_matsuno_step.
#include <stdlib.h>

#include "gcm_common.h"

extern "C" int gcm_pe25_half_step_rows(const gcm_geom*, const gcm_state*, const gcm_state*, const gcm_state*, double, int,
                                       void*, size_t, const int*, const int*, void*);
extern "C" int gcm_pe25_matsuno_step(const gcm_geom*, const gcm_state*, const gcm_state*, double, int, int, void*, size_t,
                                     void*);
bool gcm_pe25_fast_supported(const gcm_geom* g);
extern int g_gcm_knob[GCM_NKNOBS];  // pe25_fast.cu; knob 6 = latitude blocks of the host-resident step

#define GCM_HOST_MAX_BLOCKS 64

struct GcmHostPipe {
  cudaStream_t q_in, q_out;
  cudaEvent_t ev_start, ev_in[GCM_HOST_MAX_BLOCKS + 1], ev_done[GCM_HOST_MAX_BLOCKS], ev_out;
  cudaEvent_t ev_outb[GCM_HOST_MAX_BLOCKS];  // block b of the newest pipelined call has reached the host
  int outb_valid;                            // ev_outb[] have been recorded at least once
  int outb_blocks, outb_rows;                // block partition (nblocks, H) the recorded ev_outb[] belong to
  int pending;                               // a pipelined call has not been joined yet
};

#ifndef GCM_EMU
// one per device (streams and events belong to the device that was current when they were made); a step_host call is
// stream-ordered on the caller's stream
static GcmHostPipe* g_pipes[64] = {nullptr};
static int host_pipe(GcmHostPipe** out) {
  int dev = 0;
  GCM_CUDA(cudaGetDevice(&dev));
  GCM_REQUIRE(dev >= 0 && dev < 64, GCM_EUNSUP);
  GcmHostPipe*& g_pipe = g_pipes[dev];
  if (!g_pipe) {
    GcmHostPipe* p = (GcmHostPipe*)calloc(1, sizeof(GcmHostPipe));
    GCM_REQUIRE(p, (int)cudaErrorMemoryAllocation);
    GCM_CUDA(cudaStreamCreateWithFlags(&p->q_in, cudaStreamNonBlocking));
    GCM_CUDA(cudaStreamCreateWithFlags(&p->q_out, cudaStreamNonBlocking));
    GCM_CUDA(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    GCM_CUDA(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
    for (int b = 0; b <= GCM_HOST_MAX_BLOCKS; ++b) GCM_CUDA(cudaEventCreateWithFlags(&p->ev_in[b], cudaEventDisableTiming));
    for (int b = 0; b < GCM_HOST_MAX_BLOCKS; ++b) GCM_CUDA(cudaEventCreateWithFlags(&p->ev_done[b], cudaEventDisableTiming));
    for (int b = 0; b < GCM_HOST_MAX_BLOCKS; ++b) GCM_CUDA(cudaEventCreateWithFlags(&p->ev_outb[b], cudaEventDisableTiming));
    g_pipe = p;
  }
  *out = g_pipe;
  return GCM_OK;
}
#endif

// rows [r0, r1) of the five fields between host and device (dir 0: in, 1: out); 3-D fields as one strided copy
static int copy_rows(const gcm_geom* g, const gcm_state* dst, const gcm_state* src, int r0, int r1, int dir,
                     cudaStream_t q) {
  const int H = g->d.H, W = g->d.W, L = g->d.L;
  const size_t off = (size_t)r0 * W, width = (size_t)(r1 - r0) * W * sizeof(double), pitch = (size_t)H * W * sizeof(double);
  const cudaMemcpyKind kind = dir == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  GCM_CUDA(cudaMemcpyAsync(dst->p + off, src->p + off, width, kind, q));
  double* d3[4] = {dst->u, dst->v, dst->t, dst->q};
  const double* s3[4] = {src->u, src->v, src->t, src->q};
  for (int f = 0; f < 4; ++f) GCM_CUDA(cudaMemcpy2DAsync(d3[f] + off, pitch, s3[f] + off, pitch, width, L, kind, q));
  return GCM_OK;
}

// rows [a, b) of the periodic grid (a may be negative, b may exceed H) as a two-segment launch descriptor
static void wrap_seg(int a, int b, int H, int* seg) {
  if (a < 0) {
    seg[0] = a + H; seg[1] = -a; seg[2] = 0; seg[3] = b;
  } else if (b > H) {
    seg[0] = a; seg[1] = H - a; seg[2] = 0; seg[3] = b - H;
  } else {
    seg[0] = a; seg[1] = b - a; seg[2] = 0; seg[3] = 0;
  }
}

// dynamics.matsuno_timestep (dynamics.py:230-237) once, state in host memory (pinned for full speed): h_in -> h_out.
// d_cur, d_star, d_nxt are device states of the grid's shape (scratch; d_nxt ends up holding the new state too).
// nblocks <= 0: automatic.  Whole-grid geometry, one member.  Stream-ordered on `stream`: h_out is valid once the
// stream has been synchronised.
extern "C" int gcm_pe25_matsuno_step_host(const gcm_geom* g, const gcm_state* h_in, const gcm_state* h_out,
                                          const gcm_state* d_cur, const gcm_state* d_star, const gcm_state* d_nxt,
                                          double dt, int nblocks, void* ws, size_t ws_bytes, void* stream) {
  GCM_REQUIRE(g && h_in && h_out && d_cur && d_star && d_nxt && ws, GCM_ENULL);
  GCM_REQUIRE(g->d.wrap_j, GCM_EUNSUP);
  GCM_REQUIRE(!gcm_extras_on(g), GCM_EUNSUP);  // the latitude blocks carry the reference's halo widths
  const int H = g->d.H;
  cudaStream_t main = (cudaStream_t)stream;
  // automatic: pipeline only when the copies dwarf the per-block launch overhead (about 2 M cells and up)
  if (nblocks <= 0) nblocks = g_gcm_knob[6] > 0 ? g_gcm_knob[6] : ((double)H * g->d.W * g->d.L >= 2e6 ? 8 : 1);
  if (nblocks > GCM_HOST_MAX_BLOCKS) nblocks = GCM_HOST_MAX_BLOCKS;
  while (nblocks > 1 && H / nblocks < 8) --nblocks;
  int st;
#ifndef GCM_EMU
  if (nblocks > 1 && gcm_pe25_fast_supported(g)) {
    GcmHostPipe* pp;
    if ((st = host_pipe(&pp))) return st;
    if (pp->pending) {  // an unjoined pipelined call: its copy-outs come first
      GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_out, 0));
      pp->pending = 0;
    }
    pp->outb_valid = 0;
    GCM_CUDA(cudaEventRecord(pp->ev_start, main));  // everything the caller queued before is done first
    GCM_CUDA(cudaStreamWaitEvent(pp->q_in, pp->ev_start, 0));
    GCM_CUDA(cudaStreamWaitEvent(pp->q_out, pp->ev_start, 0));
    // the last two rows first (block 0 reads them across the periodic edge), then the blocks in order
    if ((st = copy_rows(g, d_cur, h_in, H - 2, H, 0, pp->q_in))) return st;
    const int rows = (H + nblocks - 1) / nblocks;
    for (int b = 0; b < nblocks; ++b) {
      const int r0 = b * rows, r1 = r0 + rows < H ? r0 + rows : H;
      if ((st = copy_rows(g, d_cur, h_in, r0, r1, 0, pp->q_in))) return st;
      GCM_CUDA(cudaEventRecord(pp->ev_in[b], pp->q_in));
    }
    for (int b = 0; b < nblocks; ++b) {
      const int r0 = b * rows, r1 = r0 + rows < H ? r0 + rows : H;
      if (r0 >= H) break;
      // rows up to r1 + 3 must have arrived: they lie in the next block (blocks have at least 8 rows)
      const int need = b + 1 < nblocks && (b + 1) * rows < H ? b + 1 : b;
      GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_in[need], 0));
      if (b == 0) GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_in[0], 0));
      int sr[4], su[4];
      wrap_seg(r0 - 1, r1 + 3, H, sr);  // predictor: star rows [r0-1, r1+2) need the row phase of [r0-1, r1+3)
      wrap_seg(r0 - 1, r1 + 2, H, su);
      if ((st = gcm_pe25_half_step_rows(g, d_cur, d_cur, d_star, dt, 1, ws, ws_bytes, sr, su, main))) return st;
      wrap_seg(r0, r1 + 1, H, sr);      // corrector: rows [r0, r1)
      wrap_seg(r0, r1, H, su);
      if ((st = gcm_pe25_half_step_rows(g, d_cur, d_star, d_nxt, dt, 1, ws, ws_bytes, sr, su, main))) return st;
      GCM_CUDA(cudaEventRecord(pp->ev_done[b], main));
      GCM_CUDA(cudaStreamWaitEvent(pp->q_out, pp->ev_done[b], 0));
      if ((st = copy_rows(g, h_out, d_nxt, r0, r1, 1, pp->q_out))) return st;
    }
    GCM_CUDA(cudaEventRecord(pp->ev_out, pp->q_out));
    GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_out, 0));
    return GCM_OK;
  }
#endif
  // one block: copy in, step, copy out on the caller's stream
  if ((st = copy_rows(g, d_cur, h_in, 0, H, 0, main))) return st;
  if ((st = gcm_pe25_matsuno_step(g, d_cur, d_nxt, dt, 1, 1, ws, ws_bytes, main))) return st;
  (void)d_star;
  return copy_rows(g, h_out, d_nxt, 0, H, 1, main);
}

// ---------------------------------------------------------------------------------------------------
// The same step for a caller that steps again and again through host buffers (a time loop whose state lives in host
// memory between steps): consecutive calls are PIPELINED ACROSS STEPS.  gcm_pe25_matsuno_step_host joins the copy-out
// stream into the caller's stream before it returns, so the copy-in of the next call cannot start before the last
// block of this one has reached the host, and each direction of the PCIe link idles while the other fills / drains.
// Here
//   * the call does NOT join: h_out is complete only after gcm_host_pipe_join(stream) (or a later non-pipelined call);
//   * the blocks are visited in the rotated order start_block, start_block + 1, ... (mod nblocks); a caller that passes
//     start_block = call number makes block k of call n + 1 depend only on blocks k - 1 .. k + 1 of call n (the
//     periodic grid is a ring of blocks), which call n finished two positions earlier;
//   * the copy-in of block b waits for the copy-out of block b of the previous pipelined call (the caller may be feeding
//     the previous output back in: bench.py and a host-resident time loop do), nothing else;
//   * d_in / d_out must ALTERNATE between two pairs of device states from call to call (the copy-in of call n + 1
//     overlaps the compute and copy-out of call n); d_star and the workspace are shared (compute is in stream order).
// Same kernels, same row segments per block: bit-identical to gcm_pe25_matsuno_step.  nblocks >= 3.
// ---------------------------------------------------------------------------------------------------
extern "C" int gcm_pe25_matsuno_step_host_pipelined(const gcm_geom* g, const gcm_state* h_in, const gcm_state* h_out,
                                                    const gcm_state* d_in, const gcm_state* d_star, const gcm_state* d_out,
                                                    double dt, int nblocks, int start_block, void* ws, size_t ws_bytes,
                                                    void* stream) {
  GCM_REQUIRE(g && h_in && h_out && d_in && d_star && d_out && ws, GCM_ENULL);
  GCM_REQUIRE(g->d.wrap_j && !gcm_extras_on(g) && gcm_pe25_fast_supported(g), GCM_EUNSUP);
#ifdef GCM_EMU
  (void)dt; (void)nblocks; (void)start_block; (void)ws_bytes; (void)stream;
  return GCM_EUNSUP;
#else
  const int H = g->d.H;
  cudaStream_t main = (cudaStream_t)stream;
  if (nblocks <= 0) nblocks = g_gcm_knob[6] > 0 ? g_gcm_knob[6] : 8;
  if (nblocks > GCM_HOST_MAX_BLOCKS) nblocks = GCM_HOST_MAX_BLOCKS;
  while (nblocks > 3 && H / nblocks < 8) --nblocks;
  GCM_REQUIRE(nblocks >= 3 && H / nblocks >= 8, GCM_ESHAPE);
  GcmHostPipe* pp;
  int st;
  if ((st = host_pipe(&pp))) return st;
  auto lo = [&](int b) { return (int)((long long)b * H / nblocks); };  // balanced blocks: sizes differ by at most one
  auto wrapb = [&](int b) { return ((b % nblocks) + nblocks) % nblocks; };
  const int s0 = wrapb(start_block);
  GCM_CUDA(cudaEventRecord(pp->ev_start, main));  // everything the caller queued before (incl. the previous compute)
  GCM_CUDA(cudaStreamWaitEvent(pp->q_in, pp->ev_start, 0));
  if (pp->outb_valid && (pp->outb_blocks != nblocks || pp->outb_rows != H)) {
    // the previous pipelined call cut another grid or other blocks: its per-block events do not describe these rows;
    // wait for all of its copy-outs instead
    GCM_CUDA(cudaStreamWaitEvent(pp->q_in, pp->ev_out, 0));
    pp->outb_valid = 0;
  }
  // copy-in, blocks s0 - 1, s0, s0 + 1, ... : block x waits for the previous call's copy-out of block x
  for (int k = 0; k < nblocks; ++k) {
    const int x = wrapb(s0 - 1 + k);
    if (pp->outb_valid) GCM_CUDA(cudaStreamWaitEvent(pp->q_in, pp->ev_outb[x], 0));
    if ((st = copy_rows(g, d_in, h_in, lo(x), lo(x + 1), 0, pp->q_in))) return st;
    GCM_CUDA(cudaEventRecord(pp->ev_in[x], pp->q_in));
  }
  for (int k = 0; k < nblocks; ++k) {
    const int b = wrapb(s0 + k);
    const int r0 = lo(b), r1 = lo(b + 1);
    // rows [r0 - 2, r1 + 4) of the ring: the block itself and its two neighbours
    GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_in[wrapb(b - 1)], 0));
    GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_in[b], 0));
    GCM_CUDA(cudaStreamWaitEvent(main, pp->ev_in[wrapb(b + 1)], 0));
    int sr[4], su[4];
    wrap_seg(r0 - 1, r1 + 3, H, sr);  // predictor: star rows [r0-1, r1+2) need the row phase of [r0-1, r1+3)
    wrap_seg(r0 - 1, r1 + 2, H, su);
    if ((st = gcm_pe25_half_step_rows(g, d_in, d_in, d_star, dt, 1, ws, ws_bytes, sr, su, main))) return st;
    wrap_seg(r0, r1 + 1, H, sr);      // corrector: rows [r0, r1)
    wrap_seg(r0, r1, H, su);
    if ((st = gcm_pe25_half_step_rows(g, d_in, d_star, d_out, dt, 1, ws, ws_bytes, sr, su, main))) return st;
    GCM_CUDA(cudaEventRecord(pp->ev_done[b], main));
    GCM_CUDA(cudaStreamWaitEvent(pp->q_out, pp->ev_done[b], 0));
    if ((st = copy_rows(g, h_out, d_out, r0, r1, 1, pp->q_out))) return st;
    GCM_CUDA(cudaEventRecord(pp->ev_outb[b], pp->q_out));
  }
  GCM_CUDA(cudaEventRecord(pp->ev_out, pp->q_out));
  pp->outb_valid = 1;
  pp->outb_blocks = nblocks;
  pp->outb_rows = H;
  pp->pending = 1;
  return GCM_OK;
#endif
}

// makes `stream` wait for the copy-outs of every pipelined call issued so far on this device
extern "C" int gcm_host_pipe_join(void* stream) {
#ifndef GCM_EMU
  GcmHostPipe* pp;
  int st;
  if ((st = host_pipe(&pp))) return st;
  if (pp->pending) {
    GCM_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, pp->ev_out, 0));
    pp->pending = 0;
  }
#else
  (void)stream;
#endif
  return GCM_OK;
}
