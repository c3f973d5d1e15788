/* gcm_b200.h -- C ABI of the B200-native Matsuno C-grid dynamical-core stepper.
 *
 * Drop-in boundary for ONE path of marthinwurer/gcmiipy: the Matsuno forward-backward step on the
 * Arakawa C-grid and the operators it is built from.  The reference has no FFI layer: its boundary
 * is plain Python module functions.  Each entry point below names the
 * reference function (file:line under the reference checkout) whose result it reproduces; the Python
 * package `gcmiipy_b200` binds them with ctypes under the reference's own module/function names
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer to C-contiguous float64; h_* is a host pointer;
 *   - 2-D fields are [j][i] (H rows x W columns, i fastest, row 0 = north); 3-D fields are
 *     [k][j][i] (L layers, k = 0 = surface); a leading ensemble dimension [b] is allowed where a
 *     function takes `nbatch`;
 *   - out-of-place: inputs are never written, outputs never alias inputs (the reference functions
 *     are pure);
 *   - all work is enqueued on the caller's stream (`stream` is a cudaStream_t passed as void*; NULL =
 *     default stream); no hidden synchronisation unless stated;
 *   - every function returns int: 0 = ok, < 0 = argument error (GCM_E*), > 0 = cudaError_t.
 *     Nothing throws or aborts.  There is no CPU fallback: without a CUDA device every compute entry
 *     point returns the CUDA error.
 */
#ifndef GCM_B200_H
#define GCM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCM_OK 0
#define GCM_ENULL (-1)      /* required pointer is NULL */
#define GCM_ESHAPE (-2)     /* bad extent (<= 0, odd W for the polar filter, row range outside the band ...) */
#define GCM_EALIGN (-3)     /* device pointer not 16-byte aligned */
#define GCM_EUNSUP (-4)     /* unsupported combination */
#define GCM_EWORK (-5)      /* workspace too small */

int gcm_version(void);
const char* gcm_status_string(int status);

/* The newest non-zero status any entry point returned on the calling thread (0 if none since the last clear);
 * clear != 0 resets it.  For callers that batch several calls and look once (the reference's callers have nothing to
 * check: numpy raises or propagates NaN, matsuno_c_grid.py:184-187 polls np.isnan). */
int gcm_last_status(int clear);

/* Asynchronous "non-finite seen" watch -- replaces the caller-side poll `np.isnan(u).any()` of
 * matsuno_c_grid.py:184-187 / no_limits_2_5d.py:85-88.  Every step kernel (2.5-D update, 2-D shallow water, 2-D
 * primitive equations) adds, per thread that has just written an inf or NaN, one to a 32-bit counter that lives on the
 * device.  gcm_nonfinite_read enqueues on `stream` a 4-byte copy of the current device's counter into *host_out
 * (pinned memory keeps it asynchronous; NULL = no read) and, if reset != 0, zeroes it afterwards.  The value is valid
 * once the stream has reached that point (event / stream query): nothing here synchronises the caller. */
int gcm_nonfinite_read(unsigned int* host_out, int reset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry: metric + sigma tables kept resident on the device (geometry.py:9-182, Geom).
 * A geometry describes the rows STORED by one process: either the whole grid (wrap_j = 1, rows are
 * periodic in j exactly like np.roll in coordinates_3d.py:43-48) or one latitude band with halo rows
 * (wrap_j = 0; stored row r holds global row (band_j0 - halo_n + r) mod H_global; the caller fills
 * halo rows, see gcm_halo_*).  A step writes only the owned rows [row_lo, row_hi).
 * ---------------------------------------------------------------------------------------------- */
typedef struct gcm_geom gcm_geom;

typedef struct {
  int H;            /* stored rows */
  int W;            /* columns; must be even or 1 (low_pass.py:57-59) */
  int L;            /* layers */
  int wrap_j;       /* 1 = rows periodic in j; 0 = band with halo rows */
  int row_lo;       /* owned rows [row_lo, row_hi) */
  int row_hi;
  int zero_v_row;   /* stored row of the global last row: v_n[:, -1, :] *= 0 (dynamics.py:222); -1 if absent */
  double dy;        /* geometry.py:138 */
  double ptop;      /* geometry.py:147 (Pa) */
  const double* h_sig;      /* [L]  geometry.py:84 */
  const double* h_dsig;     /* [L]  geometry.py:83 */
  const double* h_sigb;     /* [L]  geometry.py:81 */
  const double* h_sigt;     /* [L]  geometry.py:80 */
  const double* h_dx_j;     /* [H]  geometry.py:136 */
  const double* h_dx_h;     /* [H]  geometry.py:137 */
  const double* h_heightmap;/* [H*W] geometry.py:149 (m); NULL = zeros */
  const double* h_smmz;     /* [H*(W/2+1)] polar-filter multipliers, low_pass.py:61-72; NULL iff W == 1 */
  int zero_v_row2;  /* a second stored row holding the global last row (a band whose halo rows are recomputed by the
                       one-exchange schedule of gcm_band_matsuno_step can hold it twice); -1 if absent */
} gcm_geom_desc;

int gcm_geom_create(const gcm_geom_desc* desc, gcm_geom** out);
int gcm_geom_destroy(gcm_geom* g);

/* ------------------------------------------------------------------------------------------------
 * 2.5-D sigma-layer primitive equations (dynamics.py).  State = surface pressure p[H][W] and
 * u, v, t, q [L][H][W], with an optional leading ensemble dimension of `nbatch` members.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double* p;
  double* u;
  double* v;
  double* t;
  double* q;
} gcm_state;

/* bytes of device scratch a half step / Matsuno step needs for `nbatch` members */
size_t gcm_pe25_workspace_bytes(const gcm_geom* g, int nbatch);

/* element offset (doubles, member 0) of a work field a half step leaves in the workspace: 0 = spu, the filtered mass
 * flux arakawa_1977(su * iph(sp)) (dynamics.py:187-189); 1 = the filtered pgfu + phiu (:202); 2 = pit (:39-40);
 * 3 = p_n (:193-194; 2 and 3 are produced only when aflux runs as its own kernel: wide grids with a compile-time FFT
 * plan fuse it into the filter and form pit / p_n inside the update kernel).  (size_t)-1 for an unknown field.
 * Diagnostic / test access only. */
size_t gcm_pe25_workspace_field(const gcm_geom* g, int nbatch, int which);

/* dynamics.half_timestep (dynamics.py:183-227): out = base + dt * F(star).  Rows of `star` within
 * [-1, +2] of an owned row (and row +1 of base.p) must be valid (halo rows in band mode). */
int gcm_pe25_half_step(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out,
                       double dt, int nbatch, void* d_workspace, size_t workspace_bytes, void* stream);

/* The same half step restricted to explicit stored-row segments (fused kernels only, else GCM_EUNSUP):
 * seg = {a, n1, c, n2} = n1 rows from a, then n2 rows from c.  The row phase (filtered mass flux, sigma-dot, p_n,
 * filtered pressure-gradient force) runs on seg_r, the update on seg_u; row j of the update needs the row phase of
 * rows j and j + 1.  A latitude band computes its interior rows while the halo exchange is in flight and the rows
 * next to the halos afterwards.  No counterpart in the reference (single process, np.roll). */
int gcm_pe25_half_step_rows(const gcm_geom* g, const gcm_state* base, const gcm_state* star, const gcm_state* out,
                            double dt, int nbatch, void* d_workspace, size_t workspace_bytes, const int* seg_r,
                            const int* seg_u, void* stream);

/* dynamics.matsuno_timestep (dynamics.py:230-237), `nsteps` times, whole-grid geometry (wrap_j = 1).
 * `in` is not modified; the result of the last step lands in `out`. */
int gcm_pe25_matsuno_step(const gcm_geom* g, const gcm_state* in, const gcm_state* out, double dt, int nsteps,
                          int nbatch, void* d_workspace, size_t workspace_bytes, void* stream);

/* dynamics.matsuno_timestep once for a caller whose state lives in HOST memory (what the reference's callers have:
 * numpy arrays in, numpy arrays out).  The grid is cut into latitude blocks that are copied in, stepped and copied
 * out on three streams, so both directions of the PCIe link run at once and the kernels hide under the copies
 * (csrc/host_step.cu); bit-identical to gcm_pe25_matsuno_step.  h_in / h_out: host states (pinned for full speed);
 * d_cur, d_star, d_nxt: device scratch states of the grid's shape (d_nxt also ends up holding the new state).
 * nblocks <= 0: automatic.  Whole-grid geometry, one member.  Stream-ordered: h_out is valid after the stream syncs. */
int gcm_pe25_matsuno_step_host(const gcm_geom* g, const gcm_state* h_in, const gcm_state* h_out, const gcm_state* d_cur,
                               const gcm_state* d_star, const gcm_state* d_nxt, double dt, int nblocks,
                               void* d_workspace, size_t workspace_bytes, void* stream);

/* The same step pipelined ACROSS consecutive calls, for a time loop whose state lives in host memory between steps
 * (the reference's own situation: numpy arrays in, numpy arrays out, dynamics.py:230-237).  The call does not join its
 * copy-out stream into `stream`: h_out is complete after gcm_host_pipe_join(stream).  Blocks are visited in the
 * rotated order start_block, start_block + 1, ... (pass the call number); the copy-in of a block waits only for the
 * previous call's copy-out of the same block, so both directions of the PCIe link stay busy from step to step.
 * d_in / d_out must alternate between two pairs of device states from call to call.  Bit-identical to
 * gcm_pe25_matsuno_step.  GCM_EUNSUP without the fused row-segment kernels. */
int gcm_pe25_matsuno_step_host_pipelined(const gcm_geom* g, const gcm_state* h_in, const gcm_state* h_out,
                                         const gcm_state* d_in, const gcm_state* d_star, const gcm_state* d_out,
                                         double dt, int nblocks, int start_block, void* d_workspace,
                                         size_t workspace_bytes, void* stream);
int gcm_host_pipe_join(void* stream);

/* Kernel path of the half step: 0 (default) = the fused kernels of pe25_fast_impl.h whenever the geometry allows
 * (L in {3, 9, 17, 18}, W a product of 2, 3, 5), else the general 4-kernel path; 1 = always the general path (A/B
 * comparisons, widths with other prime factors). */
int gcm_pe25_select_path(int path);
/* launch-shape / kernel-choice tuning knobs of the fused kernels (idx 0..23, listed in pe25_fast.cu); 0 = automatic.
 * They select between measured alternatives (e.g. 4 = 5: the TMA update kernel, 14 = 2: the pipelined filter, 7 = 3: the
 * marching hydro kernel) and never change results beyond the last bits; GCM_ESHAPE for an unknown index. */
int gcm_tuning_knob(int idx, int value);

/* Opt-in terms of the 2.5-D half step.  All OFF by default: the step is then the
 * reference's dynamics.half_timestep and nothing else.  The reference only sketches these terms -- the Coriolis
 * block dynamics.py:82-95 is complete but sits behind `if False:`, tracer flux limiting is a TODO (dynamics.py:217-218)
 * and viscosity.py is wired into the 2-D matsumo_temp.py only -- so they are composed from the reference's own
 * building blocks:
 *   coriolis  dynamics.py:86-95 as written (h_cor_u/h_cor_v = 2 sin(lat) w at the u / v rows, :91-92, [H] each);
 *   nu        kinematic horizontal viscosity (m2/s): pi * nu * lap(u), lap = viscosity.py:12-19 with dx_j (dx_h for v) in i
 *             and dy in j;
 *   limit_q / limit_t  horizontal advection of q / theta (dynamics.py:174-181) with the edge value
 *             upwind + 1/2 van_leer(r) * (downwind - upwind), r = flux_limiter.calc_r seen from the upwind cell
 *             (flux_limiter.py:10-27); phi = 1 is the reference's centred flux, phi = 0 its donor cell.
 * Applied by one extra launch per half step (csrc/pe25_extras.cu) inside gcm_pe25_half_step / gcm_pe25_matsuno_step.
 * The limiter reads rows j - 2 ... j + 2: whole-grid geometries, or band geometries that store at least two halo rows
 * on either side (stepped with gcm_pe25_half_step around halo exchanges; h_cor_u / h_cor_v then hold the values of the
 * STORED rows; gcm_band_matsuno_step then exchanges two rows each way before each whole-band half step).  With any
 * option on, gcm_pe25_half_step_rows and gcm_pe25_matsuno_step_host (schedules built on the reference's halo widths)
 * return GCM_EUNSUP.  opt = NULL switches everything off. */
typedef struct {
  int coriolis;
  int limit_q;
  int limit_t;
  double nu;
  const double* h_cor_u;   /* [H], required when coriolis != 0 */
  const double* h_cor_v;   /* [H] */
} gcm_pe25_options;
int gcm_pe25_set_options(gcm_geom* g, const gcm_pe25_options* opt);

/* the operators half_timestep is built from, each on the owned rows of one member */
int gcm_pe25_calc_pu(const gcm_geom* g, const double* d_p, const double* d_u, double* d_pu, void* stream);   /* dynamics.py:15 */
int gcm_pe25_calc_pv(const gcm_geom* g, const double* d_p, const double* d_v, double* d_pv, void* stream);   /* dynamics.py:20 */
int gcm_pe25_un_pu(const gcm_geom* g, const double* d_pu, const double* d_p, double* d_u, void* stream);     /* dynamics.py:25 */
int gcm_pe25_un_pv(const gcm_geom* g, const double* d_pv, const double* d_p, double* d_v, void* stream);     /* dynamics.py:30 */
int gcm_pe25_aflux(const gcm_geom* g, const double* d_pu, const double* d_pv, double* d_pit, double* d_sd,
                   void* stream);                                                                          /* dynamics.py:35 */
int gcm_pe25_advec_sig(const gcm_geom* g, const double* d_sd, const double* d_q, double* d_out, void* stream); /* dynamics.py:49 */
int gcm_pe25_advec_m_pu(const gcm_geom* g, const double* d_p, const double* d_u, const double* d_v,
                        const double* d_pu, const double* d_pv, double* d_dut, double* d_dvt, void* stream);  /* dynamics.py:55 */
int gcm_pe25_geopotential(const gcm_geom* g, const double* d_p, const double* d_t, double* d_phi,
                          void* stream);                                                                   /* dynamics.py:111 */
int gcm_pe25_pgf(const gcm_geom* g, const double* d_p, const double* d_t, double* d_pgfu, double* d_pgfv,
                 double* d_phiu, double* d_phiv, void* d_workspace, size_t workspace_bytes, void* stream);   /* dynamics.py:147 */
int gcm_pe25_advec_t(const gcm_geom* g, const double* d_pu, const double* d_pv, const double* d_t,
                     double* d_out, void* stream);                                                         /* dynamics.py:174 */

/* low_pass.arakawa_1977 (low_pass.py:41-78): per-row rFFT x smmz[j][n] x irFFT on `nlayers` layers of
 * the geometry's owned rows.  low_pass.avrx (low_pass.py:14-38) is the same kernel with a 0/1 table:
 * pass its table through `d_table` ([H][W/2+1], device) or NULL for the geometry's smmz. */
int gcm_polar_filter(const gcm_geom* g, const double* d_in, double* d_out, int nlayers, const double* d_table,
                     void* stream);

/* diagnostics of no_limits_2_5d.full_timestep (no_limits_2_5d.py:85-91): min and max of a field
 * -> d_out[0], d_out[1]; number of non-finite values -> d_out[2] (as a double). */
int gcm_diag_minmax(const double* d_x, size_t n, double* d_out3, void* stream);
/* no_limits_2_5d.calc_energy (no_limits_2_5d.py:35-60): ke, cpT, geopotential sums -> d_out[0..2];
 * h_area_by_i is the reference's broadcast of geom.area along i (valid when H == W or H == 1). */
int gcm_pe25_energy(const gcm_geom* g, const gcm_state* s, const double* d_area_by_i, double* d_out3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latitude-band halo rows (emulates np.roll over j across ranks, coordinates_3d.py:43-48).
 * pack: copy `nrows` stored rows starting at `row0` of each of the 5 state fields (p: 1 layer,
 * u,v,t,q: L layers) into one contiguous buffer; unpack is the inverse.  Buffer layout:
 * [p rows][u rows][v rows][t rows][q rows], each [layer][row][i].
 * ---------------------------------------------------------------------------------------------- */
size_t gcm_halo_buffer_doubles(const gcm_geom* g, int nrows);
int gcm_halo_pack(const gcm_geom* g, const gcm_state* s, int row0, int nrows, double* d_buf, void* stream);
int gcm_halo_unpack(const gcm_geom* g, const gcm_state* s, int row0, int nrows, const double* d_buf, void* stream);
/* same, but copies straight between two states (used for the periodic self-wrap and for peer-mapped
 * neighbour memory over NVLink): dst rows [dst_row0, +nrows) <- src rows [src_row0, +nrows) */
int gcm_halo_copy_rows(const gcm_geom* g, const gcm_state* src, int src_row0, const gcm_state* dst, int dst_row0,
                       int nrows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latitude bands across GPUs: one process per GPU, ring of bands (np.roll over j is periodic), halo rows moved with
 * ncclSend / ncclRecv over NVLink on a side stream while the interior rows are computed.  NCCL is bound at run time
 * (dlopen of the libnccl already in the process, e.g. torch's); status >= 1000000 = 1000000 + ncclResult_t.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gcm_comm gcm_comm;
/* rank 0 makes the 128-byte NCCL unique id; the caller ships it to the other ranks (e.g. torch.distributed) */
int gcm_comm_unique_id(unsigned char* h_out128);
/* collective over the ranks of the ring; nranks == 1 needs no NCCL and no id (the ring closes on the band itself);
 * h_id128 == NULL with nranks > 1: no NCCL communicator, the ring then needs the peer mailboxes below */
int gcm_comm_create(int nranks, int rank, const unsigned char* h_id128, gcm_comm** out);
int gcm_comm_destroy(gcm_comm* c);

/* Peer-memory halo exchange over NVLink / NVSwitch (replaces pack -> ncclSend/ncclRecv -> unpack: emulates the same
 * np.roll over j, coordinates_3d.py:43-48).  Every rank owns a MAILBOX in device memory; a neighbour maps it (CUDA
 * IPC) and its push kernel stores my halo rows straight into it over NVLink, then publishes a message number in it
 * (system-scope fence + flag); my pull kernel waits for the number (bounded spin) and copies the rows into my halo
 * rows.  Two slots alternate, so a rank may run one step ahead of its neighbours.  No host involvement per step.
 *   gcm_comm_peer_setup   allocates the mailbox for the halo layout of band geometry g; h_handle64 = its IPC handle
 *   gcm_comm_peer_connect maps the ring neighbours' mailboxes from their handles (same_process != 0: `north` /
 *                         `south` are gcm_comm* of this process instead -- single-process tests of the protocol)
 *   gcm_comm_peer_status  returns 0 = NCCL ring, 1 = IPC peers, 2 = same-process peers; *timeouts = pull kernels
 *                         that gave up waiting for a neighbour (synchronises the device)
 *   gcm_band_halo_peer    one exchange of hn / hs halo rows of `s`, or one half of it: phase 1 = push, 2 = pull
 * Once connected, gcm_band_matsuno_step exchanges through the mailboxes. */
int gcm_comm_peer_setup(gcm_comm* c, const gcm_geom* g, unsigned char* h_handle64);
int gcm_comm_peer_connect(gcm_comm* c, const void* north, const void* south, int same_process);
int gcm_comm_peer_status(gcm_comm* c, unsigned int* h_timeouts);
int gcm_band_halo_peer(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, int phase, void* stream);
/* the two halves under the names SURVEY 8(b) gives them: begin = push (phase 1), end = wait + fill my halo rows (2) */
int gcm_halo_exchange_begin(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, void* stream);
int gcm_halo_exchange_end(const gcm_geom* g, gcm_comm* c, const gcm_state* s, int hn, int hs, void* stream);
/* dynamics.matsuno_timestep (dynamics.py:230-237) `nsteps` times on this rank's band (geometry with wrap_j = 0,
 * 1 halo row north and 2 south, or 2 + 4 for the one-exchange schedule, or 2 + 2 with opt-in terms on).  `cur` =
 * the band incl. halo rows (halo content is overwritten), `star` = scratch
 * state of the same shape; the newest state ends in `nxt` if nsteps is odd, else in `cur`.  overlap = 1: interior
 * rows run while the halos are in flight.  Bit-identical to the whole-grid step for any number of ranks. */
int gcm_band_matsuno_step(const gcm_geom* g, gcm_comm* c, const gcm_state* cur, const gcm_state* star,
                          const gcm_state* nxt, double dt, int nsteps, int overlap, void* d_workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2-D schemes on a uniform doubly periodic grid.
 * ---------------------------------------------------------------------------------------------- */
/* matsuno_c_grid.matsumo_scheme (matsuno_c_grid.py:125-142), nsteps times; both sub-steps fused in
 * one launch (radius-2 halo tile).  Bit-identical to the reference arithmetic (no FMA contraction). */
int gcm_sw2d_matsuno_step(const double* d_u, const double* d_v, const double* d_p, double* d_u_out,
                          double* d_v_out, double* d_p_out, int H, int W, double dx, double dt, int nsteps,
                          void* d_workspace, size_t workspace_bytes, void* stream);
size_t gcm_sw2d_workspace_bytes(int H, int W);
/* op: 0 advection_of_velocity_u (:15)  1 advection_of_velocity_v (:54)  2 geopotential_gradient_u (:97)
 *     3 geopotential_gradient_v (:103) 4 advection_of_geopotential (:109); unused inputs may be NULL */
int gcm_sw2d_operator(int op, const double* d_u, const double* d_v, const double* d_p, double* d_out, int H, int W,
                      double dx, void* stream);

/* no_limits_2d.half_timestep / matsuno_timestep (no_limits_2d.py:104-131); q passes through */
int gcm_pe2d_half_step(const gcm_state* base, const gcm_state* star, const gcm_state* out, int H, int W, double dt,
                       double dx, void* stream);
int gcm_pe2d_matsuno_step(const gcm_state* in, const gcm_state* out, int H, int W, double dt, double dx, int nsteps,
                          void* d_workspace, size_t workspace_bytes, void* stream);
size_t gcm_pe2d_workspace_bytes(int H, int W);
/* op: 0 advec_m -> (dut, dvt) (:47)   1 pgf -> (pgfu, pgfv) (:76) */
int gcm_pe2d_operator(int op, const double* d_p, const double* d_u, const double* d_v, const double* d_t,
                      double* d_out0, double* d_out1, int H, int W, double dx, void* stream);

/* matsumo_temp.matsumo_scheme (matsumo_temp.py:66-99): shallow water + temperature + viscosity */
int gcm_swt2d_matsuno_step(const double* d_u, const double* d_v, const double* d_p, const double* d_t,
                           double* d_u_out, double* d_v_out, double* d_p_out, double* d_t_out, int H, int W,
                           double dx, double dt, double mu, int nsteps, void* d_workspace, size_t workspace_bytes,
                           void* stream);
size_t gcm_swt2d_workspace_bytes(int H, int W);

/* viscosity.finite_laplacian_2d (viscosity.py:12): out = (q[j+1]+q[j-1]+q[i+1]+q[i-1]-4q)/(dx*dx) * scale;
 * scale = 1 for the Laplacian, mu for incompressible_viscosity_2d (viscosity.py:22: mu * lap) */
int gcm_laplacian5(const double* d_q, double* d_out, int H, int W, double dx, double mu, int apply_mu, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone operators
 * ---------------------------------------------------------------------------------------------- */
/* phi_port.PGF (phi_port.py:5-113) on [k][j][i] inputs (the transposes of the reference's T[W][H][L],
 * P[W][H]); only column i = 0 of every row is computed (IMAX = 1, :50-54), the rest of phi is zero */
int gcm_phi_port_pgf(const gcm_geom* g, const double* d_p, const double* d_t, double* d_phi, void* stream);

/* flux_limiter.py:10-32 on `nrows` independent periodic rows of length n */
int gcm_fl_van_leer(const double* d_r, double* d_out, size_t n, void* stream);                               /* :10 */
int gcm_fl_calc_r(const double* d_q, double* d_out, int nrows, int n, void* stream);                         /* :14 */
int gcm_fl_donor_cell_flux(const double* d_q, const double* d_u, double* d_out, int nrows, int n, void* stream); /* :23 */
int gcm_fl_donor_cell_advection(const double* d_q, const double* d_u, double* d_out, int nrows, int n, double dx,
                                double dt, int nsteps, double* d_tmp, void* stream);                         /* :30 */

/* coordinates*.py shift / half-average / gradient helpers on an [n2][n1][n0] array (n0 fastest).
 * axis: 0 = i (fastest), 1 = j, 2 = k.  op: 0 roll by `shift` (np.roll semantics, constants.py:85)
 * 1 (q + roll(q, shift))/2   2 (roll(q,-1) - q)/d */
int gcm_shift_op(int op, const double* d_q, double* d_out, int n2, int n1, int n0, int axis, int shift, double d,
                 void* stream);

/* temperature.py:7-19: dir 0: T = theta / (P0/p)^kappa ; dir 1: theta = T * (P0/p)^kappa */
int gcm_temperature_convert(int dir, const double* d_t, const double* d_p, double* d_out, size_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optional per-kernel timing (CUDA events on the launching stream), used by bench.py for the live
 * roofline of the dominant kernel.  Off by default.  gcm_prof_collect synchronises the device, writes the
 * summed milliseconds and launch counts per kernel kind ([gcm_prof_kinds()] entries) and clears them.
 * The reference has no counterpart (its only instrumentation is tqdm, no_limits_2_5d.py:230).
 * ---------------------------------------------------------------------------------------------- */
int gcm_prof_enable(int on);
int gcm_prof_kinds(void);
const char* gcm_prof_kind_name(int kind);
int gcm_prof_collect(double* h_ms, long long* h_launches);

/* ------------------------------------------------------------------------------------------------
 * Grey-radiation column physics: the step after the dynamics in no_limits_2_5d.full_timestep.
 * One thread per column, layer recursions in the reference's order.
 *   gcm_grey_radiation  grey_solar.basic_grey_radiation (grey_solar.py:358-563) with zenith_angle (:49-68, declination
 *                       0): p surface pressure [H][W], tt true temperature [L][H][W], gt ground temperature [H][W] ->
 *                       dTdt [L][H][W] (K/s), dt_ground [H][W] (K/s).
 *   gcm_solar_timestep  no_limits_2_5d.solar_timestep (:66-75): potential temperature t -> t_n, gt -> gt_n over dt.
 * h_lw / h_sw: HOST arrays [L] = t_lw ** dsig, t_sw ** dsig (basic_grey_transmittances, :323-333); sinlat / coslat
 * [H], lon [W]: device tables in radians; hour_angle = utc / (-24 h) * 360 deg in radians (:51).  Needs L <= 24.
 * ---------------------------------------------------------------------------------------------- */
int gcm_grey_radiation(const gcm_geom* g, const double* p, const double* tt, const double* gt, const double* h_lw,
                       const double* h_sw, double albedo, const double* sinlat, const double* coslat, const double* lon,
                       double hour_angle, double* dTdt, double* dt_ground, void* stream);
int gcm_solar_timestep(const gcm_geom* g, const double* p, const double* t, const double* gt, const double* h_lw,
                       const double* h_sw, double albedo, const double* sinlat, const double* coslat, const double* lon,
                       double hour_angle, double dt, double* t_n, double* gt_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCM_B200_H */
