/* sphb200.h -- C ABI of the B200-native JAX-SPH per-step engine (libsphb200.so).
 *
 * Everything a host binding (the jax.ffi shim in csrc/ffi_shim.cc, the ctypes
 * loader in jax_sph_b200/_lib.py) needs.  Plain C: POD structs, raw pointers,
 * sizes, a cudaStream_t passed as void*.  No torch / jax types.
 *
 * The reference (tumaer/jax-sph) has NO native interface for this path -- it is
 * pure Python/JAX -- so each entry point names the Python callable it replaces
 * (citations relative to the reference tree):
 *
 *   sphb200_neighbor_list   <- jax_sph/partition.py:492-571  neighbor_list(...).allocate/.update
 *                              (jax_sph/jax_md/partition.py:914-983, Sparse format, .idx)
 *   sphb200_forward         <- jax_sph/solver.py:702-951     WCSPH.forward_wrapper()(state, neighbors)
 *   sphb200_advance         <- jax_sph/integrator.py:22-56   si_euler(...).advance(dt, state, neighbors)
 *                              (+ the case bc_fn / g_ext_fn in table form, cases/*.py)
 *   sphb200_engine_*        <- jax_sph/simulate.py:110-134   the step loop, state resident in HBM
 *
 * Conventions
 *   - All functions return 0 on success or a negative SPHB200_E* code; they never
 *     abort, throw or print.  sphb200_strerror() gives the text.
 *   - Device work is enqueued on the caller's stream; nothing synchronises the
 *     device unless documented ("sync").  Run-time conditions (neighbour-list
 *     overflow, staging overflow) are reported through a device-side error word
 *     (bits below), mirroring PartitionErrorCode (jax_md/partition.py:434-457).
 *   - Arrays use the reference layouts: vector fields are (N, dim) row-major
 *     float32, scalar fields (N,) float32, tag (N,) int32.
 *   - float32 only: the reference's float64 mode is not offered (documented drift
 *     bound instead); a float64 request fails with SPHB200_EDTYPE.
 *   - Re-entrant: engines share nothing; one engine per stream/device.  The only global state
 *     is the mutex-protected workspace table of the stateless entry points (see below).
 */
#ifndef SPHB200_H_
#define SPHB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------- */
#define SPHB200_OK 0
#define SPHB200_EINVAL (-1)   /* bad argument / inconsistent config            */
#define SPHB200_ENOMEM (-2)   /* cudaMalloc failed or workspace too small      */
#define SPHB200_ECUDA (-3)    /* a CUDA runtime call failed                    */
#define SPHB200_EDTYPE (-4)   /* unsupported dtype (float64)                   */
#define SPHB200_EUNSUP (-5)   /* unsupported variant (e.g. DELTA + density evolution) */
#define SPHB200_ENODEV (-6)   /* no CUDA device / not an sm_100 device         */

/* ---- device-side error word (sphb200_engine_error) ---------------------- */
#define SPHB200_ERR_NEIGHBOR_OVERFLOW (1u << 0) /* == PartitionErrorCode.NEIGHBOR_LIST_OVERFLOW */
#define SPHB200_ERR_CELL_OVERFLOW (1u << 1)     /* == PartitionErrorCode.CELL_LIST_OVERFLOW     */
#define SPHB200_ERR_STAGE_OVERFLOW (1u << 2)    /* one stencil row exceeds the staging buffer   */
#define SPHB200_ERR_NONFINITE (1u << 3)         /* non-finite position met while hashing        */
#define SPHB200_ERR_OUTSIDE_BOX (1u << 4)       /* a position outside [0, box]: the periodic    *
                                                 * fold assumes shift_fn-wrapped positions       */
#define SPHB200_ERR_SLAB_OVERFLOW (1u << 5)     /* slab engine: own / halo / migration capacity  */
#define SPHB200_ERR_SLAB_TIMEOUT (1u << 8)       /* slab engine, direct transport: a ring neighbour's  *
                                                 * signal did not arrive within 20 s                   */
#define SPHB200_ERR_HINT (1u << 7)           /* a SPHB200_HINT_* promise of the config does not hold  */
#define SPHB200_ERR_SLAB_MIGRATION (1u << 6)    /* slab engine: a particle crossed more than the *
                                                 * halo width in one step                        */

/* ---- enums --------------------------------------------------------------- */
/* solver.py:666-686.  DELTA (Marrone et al. 2011): velocity diffusion of acceleration_delta_fn
 * (solver.py:259-313) and, with SPHB200_F_RHO_EVOL, the renormalised density diffusion of
 * rho_evol_fn_delta (solver.py:33-105; on a slab engine its three sweeps are separated by two
 * more halo refreshes). */
enum { SPHB200_SOLVER_SPH = 0, SPHB200_SOLVER_RIE = 1, SPHB200_SOLVER_DELTA = 2 };
/* kernel.py: Quintic :51-72, Wendland C2 :75-103 (compile-time specialised sweeps), and
 * Cubic :26-48, Wendland C4 :106-134, C6 :137-165, Gaussian :168-182, SuperGaussian :185-201
 * (one run-time-switched instance of every sweep). */
enum {
  SPHB200_KERNEL_QSK = 0, SPHB200_KERNEL_WC2K = 1, SPHB200_KERNEL_CSK = 2, SPHB200_KERNEL_WC4K = 3,
  SPHB200_KERNEL_WC6K = 4, SPHB200_KERNEL_GK = 5, SPHB200_KERNEL_SGK = 6
};
enum { SPHB200_EOS_TAIT = 0, SPHB200_EOS_RIEMANN = 1 };   /* eos.py:20-57                            */

/* solver flags (WCSPH ctor booleans, solver.py:616-637) */
#define SPHB200_F_BC_TRICK (1u << 0)        /* is_bc_trick        */
#define SPHB200_F_RHO_EVOL (1u << 1)        /* is_rho_evol        */
#define SPHB200_F_RHO_RENORM (1u << 2)      /* is_rho_renorm      */
#define SPHB200_F_FREE_SLIP (1u << 3)       /* is_free_slip       */
#define SPHB200_F_HEAT (1u << 4)            /* is_heat_conduction */

/* tags, utils.py:22-32 */
enum {
  SPHB200_TAG_PAD = -1,
  SPHB200_TAG_FLUID = 0,
  SPHB200_TAG_SOLID_WALL = 1,
  SPHB200_TAG_MOVING_WALL = 2,
  SPHB200_TAG_DIRICHLET_WALL = 3
};

/* external force g_ext_fn(r) in table form (cases/tgv.py:53, db.py:129-132, pf.py/ht.py band) */
enum { SPHB200_G_NONE = 0, SPHB200_G_CONST = 1, SPHB200_G_BAND = 2, SPHB200_G_ARRAY = 3 };

/* case bc_fn in table form: per tag, which fields are overwritten after forward
 * (cases/db.py:134-142, cf.py:143-160, pf.py, ht.py:154-187) */
#define SPHB200_BC_SET_U (1u << 0)
#define SPHB200_BC_SET_V (1u << 1)
#define SPHB200_BC_ZERO_DUDT (1u << 2)
#define SPHB200_BC_ZERO_DVDT (1u << 3)
#define SPHB200_BC_SET_P (1u << 4)
#define SPHB200_BC_SET_T (1u << 5)
#define SPHB200_BC_ZERO_DTDT (1u << 6)

typedef struct sphb200_bc_rule {
  uint32_t flags; /* SPHB200_BC_* */
  float u[3];
  float v[3];
  float p;
  float T;
} sphb200_bc_rule;

/* ---- configuration -------------------------------------------------------
 * Mirrors the WCSPH constructor (solver.py:616-637), the EoS objects
 * (eos.py), si_euler's tvf (integrator.py:8-10), space.periodic(side=box)
 * (case_setup.py:133) and the table form of the case callables.  Physical
 * parameters are doubles and are rounded to float32 exactly where the
 * reference's weakly-typed Python scalars are. */
typedef struct sphb200_config {
  uint32_t struct_size; /* = sizeof(sphb200_config), ABI check */
  int32_t dim;          /* 2 or 3 */
  int32_t solver;       /* SPHB200_SOLVER_* */
  int32_t kernel;       /* SPHB200_KERNEL_* */
  int32_t eos;          /* SPHB200_EOS_*    */
  uint32_t flags;       /* SPHB200_F_*      */
  double box[3];        /* periodic box sides (box[2] ignored in 2D) */
  double dx;            /* particle spacing                          */
  double h;             /* smoothing length = h_fac * dx             */
  double dt;            /* WCSPH.dt: rho / T integration inside forward */
  double tvf;           /* si_euler tvf                               */
  double c_ref;         /* RIE limiter                                */
  double eta_limiter;   /* RIE, -1 = off                              */
  double artificial_alpha;
  double p_ref, rho_ref, p_bg, gamma; /* TaitEoS                      */
  double u_ref;                       /* RIEMANNEoS                   */
  double r_cutoff; /* neighbour cutoff; 0 = kernel support (3h QSK, 2h WC2K, kernel.py:64,88).
                    * Set explicitly only by the neighbor_list drop-in (partition.py:492-507). */
  /* g_ext table */
  int32_t g_mode;   /* SPHB200_G_* */
  int32_t g_axis;   /* BAND: coordinate axis tested */
  double g[3];      /* CONST / BAND value */
  double g_lo, g_hi; /* BAND: g applied where lo < r[axis] < hi */
  /* bc table, indexed by tag 0..3 */
  sphb200_bc_rule bc[4];
  int32_t bc_inflow_on;  /* ht.py:159-161: fluid with r.x < x  -> T = T_in, dTdt = 0 */
  float bc_inflow_x, bc_inflow_T;
  int32_t bc_outflow_on; /* ht.py:182-185: fluid with r.x > x  -> dTdt = 0           */
  float bc_outflow_x;
  /* tuning (0 = automatic) */
  int32_t cell_sub[3]; /* cells per cutoff along each axis (1 or 2)  */
  int32_t tile[3];     /* tile size in cells                         */
  int32_t threads;     /* sweep block size                           */
  int32_t list_cap;    /* per-thread pair-list capacity              */
  int32_t stage_cap;   /* staged particles per block (0 = auto)      */
  int32_t nl_cap;      /* row length of the per-step neighbour lists shared by the sweeps of one
                        * forward(): 0 = auto (1.3 x the uniform-fluid neighbour count), -1 = off */
  float diff_delta;    /* DELTA: solver.py:623, defaults.py:97 */
  float diff_alpha;    /* DELTA: solver.py:624, defaults.py:99 */
  float skin;          /* neighbour-list skin as a fraction of the cutoff: the search (cell sort +
                        * candidate walk) runs only when a particle has travelled more than half the
                        * skin since the last one; the exact d^2 < cutoff^2 membership test of
                        * jax_md/partition.py:897 still runs for every pair on every step, so the
                        * neighbour sets are the reference's.  0 = automatic, < 0 = search every step */
  uint32_t hints;      /* SPHB200_HINT_*: promises about the state the engine may exploit; each is
                        * verified on the device every step, a broken one raises
                        * SPHB200_ERR_HINT in the device error word */
  int32_t reserved[3];
} sphb200_config;

/* State dict of the reference (solver.py:930-947), device or host pointers.
 * Optional members may be NULL: nw (zeros), T (ones), dTdt, kappa, Cp, drhodt (zeros). */
typedef struct sphb200_state {
  float *r, *u, *v, *dudt, *dvdt, *nw;                          /* (N, dim) */
  float *rho, *p, *drhodt, *mass, *eta, *T, *dTdt, *kappa, *Cp; /* (N,)     */
  int32_t *tag;                                                 /* (N,)     */
  float *g_ext; /* (N, dim), only read when g_mode == SPHB200_G_ARRAY */
} sphb200_state;

typedef struct sphb200_engine sphb200_engine;

/* config hints */
#define SPHB200_HINT_UNIFORM_ETA (1u << 0) /* state.eta is the same for every particle (every case of
                                            * the reference sets eta = viscosity, case_setup.py:152-181):
                                            * eta_ij of solver.py:243 is a constant of the run and the
                                            * force sweep does not stage eta */

/* step flags */
#define SPHB200_STEP_INTEGRATE (1u << 0) /* si_euler kick+drift before forward (advance); unset = forward only */
#define SPHB200_STEP_BC (1u << 1)        /* apply the bc table after forward                                  */

/* ---- library ------------------------------------------------------------- */
int sphb200_abi_version(void);
const char *sphb200_strerror(int code);
void sphb200_config_default(sphb200_config *cfg);

/* ---- resident engine (state stays cell-sorted in HBM between steps) ------ */
/* Bytes of device memory an engine for n particles needs. */
int sphb200_engine_bytes(const sphb200_config *cfg, int64_t n, size_t *bytes);
/* Allocates its own arena with cudaMalloc on the current device. */
int sphb200_engine_create(const sphb200_config *cfg, int64_t n, sphb200_engine **out);
/* Places the engine in caller-owned device memory (>= sphb200_engine_bytes). */
int sphb200_engine_create_in(const sphb200_config *cfg, int64_t n, void *workspace,
                             size_t workspace_bytes, sphb200_engine **out);
int sphb200_engine_destroy(sphb200_engine *e);
/* Copy a state in reference layout into the engine (on_host: pointers are host memory). */
int sphb200_engine_upload(sphb200_engine *e, const sphb200_state *s, int on_host, void *stream);
/* Copy a state of the SAME particles (row i is still particle i) into the slots they occupy in
 * the cell-sorted frame.  Unlike sphb200_engine_upload this keeps the cell table and the
 * neighbour lists: the next step tests every position against the one the particle had when the
 * cells were made and sorts again only if the skin lists may have become incomplete (a particle
 * further than half the list skin from there AND a block of neighbouring cells whose particles
 * drifted apart by more than the skin, csrc/cells.cuh k_drift / k_drift_box) -- the result is
 * the same either way.  Before the first step it is an ordinary upload. */
int sphb200_engine_refresh(sphb200_engine *e, const sphb200_state *s, int on_host, void *stream);
/* nsteps x advance(dt) (integrator.py:22-56) or forward only, per `flags`. */
int sphb200_engine_step(sphb200_engine *e, double dt, int nsteps, uint32_t flags, void *stream);
/* One advance(dt, state, neighbors) (integrator.py:22-56) on HOST buffers, the call a reference
 * user makes with a host-resident state: copies the non-NULL members of *in host -> device,
 * runs one step (`flags` as in sphb200_engine_step) and copies the non-NULL members of *out
 * device -> host in the original particle order.  The entries that are final after the
 * integrate / reorder pass (r; u and v too unless a wall sweep or the bc table rewrites them)
 * travel back on an internal high-priority stream WHILE the interaction sweeps run; `stream`
 * waits for that copy before the call's work on it ends, so synchronising `stream` is enough.
 * Pinned host memory is needed for the copies to overlap. */
int sphb200_engine_advance_host(sphb200_engine *e, double dt, const sphb200_state *in,
                                sphb200_state *out, uint32_t flags, void *stream);
/* Write the state back in the ORIGINAL particle order (index-stable API). NULL members are skipped. */
int sphb200_engine_download(sphb200_engine *e, sphb200_state *out, int on_host, void *stream);
/* Vector-Jacobian product of the step the engine ran LAST -- one advance(dt) with `flags` as
 * in sphb200_engine_step (SPHB200_STEP_INTEGRATE) or one forward() (flags 0) -- at the state it
 * holds: what jax.grad / jax.vjp give for `advance` of jax_sph/integrator.py:22-56 in
 * notebooks/iclr24_grads.ipynb (cell 5), by hand-written adjoint sweeps over the same cell-sorted
 * particles (csrc/adjoint.cuh).
 *   cot_out  cotangents of the step's outputs r, u, v, dudt, rho, p (DEVICE pointers, the
 *            caller's particle order, NULL = zero)
 *   cot_in   receives the cotangents of the step's inputs r, u, v, dudt, dvdt, rho, p (non-NULL
 *            members; rho and p are overwritten by the step: zero)
 * Supported: solver SPH, density by summation, tvf = 0, no wall / bc sweep, no artificial
 * viscosity, constant external force, Quintic or Wendland C2 kernel; otherwise SPHB200_EUNSUP.
 * A gradient through K steps is K calls in reverse order, each after re-running the step from
 * its saved input state (jax_sph_b200.engine.grad_through_steps). */
int sphb200_engine_vjp(sphb200_engine *e, double dt, uint32_t flags, const sphb200_state *cot_out,
                       sphb200_state *cot_in, void *stream);
/* sync: read and clear the device error word. */
int sphb200_engine_error(sphb200_engine *e, uint32_t *code, void *stream);
/* Sparse neighbour list of the CURRENT resident positions in original indices:
 * idx is (2, capacity) int32 on the device, row 0 receiver, row 1 sender, sorted
 * by (sender, receiver), padded with N (jax_md/partition.py:885-909).  count
 * (device int64, may be NULL) receives the number of edges found; overflow sets
 * SPHB200_ERR_NEIGHBOR_OVERFLOW.  idx == NULL with capacity 0 only counts. */
int sphb200_engine_neighbor_list(sphb200_engine *e, int32_t *idx, int64_t capacity,
                                 int mask_self, int64_t *count, void *stream);
/* Wall-normal recomputation, the integrator's `nw_fn` (jax_sph/integrator.py:33-34;
 * compute_nws_jax_wrapper, jax_sph/utils.py:197-277; enabled by case_setup.py:209-218 when a
 * MOVING_WALL exists and the solver reads normals).  layer: HOST array (n_layer, dim) float32,
 * the one-layer discretisation `wall_part_fn(dx / 5, 1) - offset / n_walls / 5`; offset: HOST
 * (dim) float32 subtracted from the wall positions; cutoff = dx * n_walls * sqrt(2) * 1.01.
 * From then on every integrating step rewrites nw after the drift: each wall particle gets
 * disp(closest layer particle, particle) / (dist + EPS), all other particles zero.
 * n_layer = 0 switches it off (normals stay what upload() provided).  sync (copies). */
int sphb200_engine_set_wall_layer(sphb200_engine *e, const float *layer, int n_layer,
                                  const float *offset, double cutoff);
/* sync: kinetic energy 0.5*sum(m u.u) (utils.py:128-133) and max |u| (utils.py:136-166). */
int sphb200_engine_stats(sphb200_engine *e, double *ekin, double *u_max, void *stream);
/* sync: get_stats (jax_sph/utils.py:156-166, what Logger.print_stats prints, :288-296) on the
 * resident state, one pass:
 *   out[0]         Ekin = 0.5 * dx^dim * sum over FLUID particles of |v|^2 (get_ekin, :128-133:
 *                  the TRANSPORT velocity, fluid only -- unlike sphb200_engine_stats)
 *   out[1 + 3k..]  min, max, SUM of get_array_stats (:136-153; Euclidean norm for vectors) for
 *                  k = 0..4: u, v, rho, p, T, over all particles (mean = sum / out[16])
 *   out[16]        particles counted (a slab engine counts its own particles: the caller
 *                  reduces min / max / sum / count over the ranks) */
#define SPHB200_NSTATS 20
int sphb200_engine_get_stats(sphb200_engine *e, double out[SPHB200_NSTATS], void *stream);
/* Kernel launches issued by this engine so far (bench.py's gpu_launches). */
int64_t sphb200_engine_launches(const sphb200_engine *e);
/* Sweep timing: CUDA-event ms of the last step's passes: [0] integrate+hash, [1] sort/reorder,
 * [2] density sweep, [3] wall sweep, [4] force sweep, [5] total.  Enabled by sphb200_engine_profile(e, 1). sync.
 * While enabled a slab engine does not overlap its exchanges with interior tiles (the passes are
 * timed whole, on one stream). */
int sphb200_engine_profile(sphb200_engine *e, int enable);
int sphb200_engine_last_times(sphb200_engine *e, float ms[8]);
/* cell-grid facts for reports: ncells[3], sub[3], tile[3], threads, list_cap, stage_cap(A,B,C). */
int sphb200_engine_plan(const sphb200_engine *e, int32_t out[16]);
/* sync: [0] steps run since creation, [1] searches (cell sort + candidate walk) among them,
 * [2] row length of the neighbour lists, [3] skin in 1e-6 of the cutoff, [4] tiles,
 * [5] tiles whose lists do not exist (swept by their own search), [6] 1 when the duo sweeps serve
 * this variant, [7] directed pairs (self pairs included) in the exact neighbour lists of the last
 * step -- duo engines whose every tile has lists, else -1. */
int sphb200_engine_counters(sphb200_engine *e, int64_t out[8], void *stream);
/* sync: measured FP32 throughput of this device in TFLOP/s (2 flop per lane-FMA): a grid of
 * independent FMA chains, `packed` = 0: FFMA, 1: FFMA2 (fma.rn.f32x2).  The roofline denominator
 * of the interaction sweeps (bench.py); ms = kernel time of the measurement. */
int sphb200_fp32_peak(int packed, double *tflops, double *ms, void *stream);

/* ---- slab decomposition: one engine per GPU, the caller moves the messages ----------------
 * The reference drives ONE device (jax_sph/simulate.py:110-134); north_star asks for the
 * periodic box to be cut into slabs with a halo exchange and particle migration each step.
 * The box is cut along the slowest-varying cell axis (z in 3D, y in 2D) into `nranks` slabs
 * of whole cell layers; rank r owns global layers [z0, z1) (sphb200_slab_info) and keeps
 * S = cell_sub halo layers (one cutoff) of its ring neighbours on each side.  One step is
 *
 *     for (phase = 0;; ++phase) {
 *       sphb200_slab_run(e, phase, dt, flags, send_lo, send_hi, recv_lo, recv_hi, stream, &nbytes);
 *       if (nbytes == 0) break;                       // step complete
 *       if (nbytes < 0)                               // -4: agree on the re-sort decision
 *         max-reduce the int32 at send_lo[0:4] over ALL ranks, in place, stream-ordered
 *       else
 *         send send_lo[0:nbytes] to rank-1, send_hi[0:nbytes] to rank+1 (periodic ring),
 *         receive recv_lo[0:nbytes] from rank-1, recv_hi[0:nbytes] from rank+1, stream-ordered
 *     }
 *
 * phase 0 integrates in place and publishes whether a particle of this rank has travelled more
 * than half the skin of the neighbour lists since the last sort; phase 1 (on the steps where
 * some rank said so) hashes and emits the emigrants, phase 2 takes the immigrants, sorts, and
 * emits the boundary layers (every step: positions and state of the halo particles change, the
 * particle sets only when the ranks sort); every later phase takes a halo message and runs the
 * sweeps up to the next one whose results the neighbours need (density -> rho, p; wall BC -> u,
 * v, rho, p).
 * Message sizes depend only on the capacities, counts travel in the message headers and stay
 * on the device: nothing synchronises with the host.  The four buffers are device memory of
 * at least outi[9] bytes each (sphb200_slab_info), owned by the caller (the transport:
 * NCCL send/recv in jax_sph_b200/slab.py).  Capacities <= 0 are chosen from the config. */
int sphb200_slab_create(const sphb200_config *cfg, int rank, int nranks, int64_t own_cap,
                        int64_t halo_cap, int64_t mig_cap, sphb200_engine **out);
/* outi: rank, nranks, axis, z0, z1, global layers, own_cap, halo_cap, mig_cap, message buffer
 * bytes, S, slots, arena bytes.  outd: float32 1/cell and box side along the slab axis (a
 * particle at r lies in layer min(int(f32(r) * f32(outd[0])), layers - 1)). */
int sphb200_slab_info(const sphb200_engine *e, int64_t outi[16], double outd[4]);
/* This rank's own particles (rows <= own_cap) and their global indices. */
int sphb200_slab_upload(sphb200_engine *e, const sphb200_state *s, const int32_t *ids, int64_t rows,
                        int on_host, void *stream);
/* Own particles in local (cell-sorted) order; ids receives their global indices; rows = array
 * capacity (>= current own count, see sphb200_slab_counts). */
int sphb200_slab_download(sphb200_engine *e, sphb200_state *out, int32_t *ids, int64_t rows,
                          int on_host, void *stream);
/* sync: device counters [own, immigrants, emigrants lo, hi, halo lo, hi, sent lo, hi]. */
int sphb200_slab_counts(sphb200_engine *e, int32_t out[8], void *stream);
int sphb200_slab_run(sphb200_engine *e, int phase, double dt, uint32_t flags, void *send_lo,
                     void *send_hi, const void *recv_lo, const void *recv_hi, void *stream,
                     int64_t *xbytes);
/* Transport without collectives (ranks that map each other's device memory: CUDA IPC / symmetric
 * memory over NVLink).  The engine only ever WRITES the send buffers, and only the live part of
 * a message (counts travel in the header), so send_lo / send_hi of sphb200_slab_run may be the
 * ring neighbours' receive buffers themselves: the pack kernels then store straight into the
 * neighbour's memory and an exchange is a pair of flags (jax_sph_b200/slab.py, DirectRing).
 * sphb200_slab_set_agree hands the engine every rank's CONTROL BLOCK -- flag_arrays[r] is rank
 * r's zero-initialised int32[3 * nranks + 2] as mapped into THIS process (flag_arrays[rank] the
 * local one): agreement words [step parity][source rank], agreement arrivals [source rank],
 * and the two exchange signals (from above, from below); all signals are sequence numbers that
 * only grow.  Phase 0 then stores this rank's re-sort word and the step number into every block
 * and phase 1 waits for all ranks' step numbers and takes the maximum of the local words: the
 * transport answers nbytes == -4 with nothing at all.  After every phase that returns nbytes > 0
 * the transport calls sphb200_slab_signal(e, stream) instead of moving bytes: one kernel on the
 * step's stream that tells both ring neighbours "my messages are complete" (system-scope release
 * behind the pack kernels' stores) and waits for theirs.  The caller alternates two sets of
 * receive buffers from exchange to exchange (a rank that signals exchange x has consumed x - 1).
 * A signal that does not arrive within 20 s raises SPHB200_ERR_SLAB_TIMEOUT instead of hanging.
 * nranks <= 16; flag_arrays == NULL switches back to the message transport. */
int sphb200_slab_set_agree(sphb200_engine *e, int32_t *const *flag_arrays, int nranks);
int sphb200_slab_signal(sphb200_engine *e, void *stream);

/* ---- on-device case initialisation ----------
 * Replaces the host-side lattice generators pos_init_cartesian_2d / _3d
 * (jax_sph/utils.py:35-54), the velocity initialisation of cases/tgv.py:37-51 and the uniform
 * field setup of SimulationSetup.initialize() (jax_sph/case_setup.py:152-181) for the
 * regular-lattice cases: at 16-64 M particles the NumPy meshgrid / vstack and the host-to-device
 * copy of the full state take seconds, the kernel a fraction of a millisecond.
 *   row order   = np.meshgrid(range(n0), range(n1)[, range(n2)], indexing="xy") ravelled
 *                 (utils.py:41,52): row = (iy * n0 + ix) * n2 + iz in 3D, iy * n0 + ix in 2D;
 *   position    = (i + 0.5) * dx per axis, float32 (utils.py:42,53);
 *   planes      = [k_lo, k_hi) of the LAST axis only (one rank's slab); rows are then counted
 *                 within the slab and `ids` (optional) receives the full-lattice row of each;
 *   walls       = the n_walls outer planes on both sides of wall_axis get SPHB200_TAG_SOLID_WALL;
 *                 lower-wall particles with hot_lo < x < hot_hi get SPHB200_TAG_DIRICHLET_WALL and
 *                 T_hot (cases/ht.py:90-97, :154-187); wall_axis = -1: no walls;
 *   velocity    = u = v = the selected field evaluated at the position (fluid particles only). */
enum { SPHB200_VEL_REST = 0, SPHB200_VEL_TGV2D = 1, SPHB200_VEL_TGV3D = 2 };

typedef struct sphb200_lattice {
  uint32_t struct_size; /* = sizeof(sphb200_lattice) */
  int32_t dim;
  int32_t n[3];       /* planes per axis = round(box / dx), utils.py:40,51 */
  int32_t k_lo, k_hi; /* planes of the last axis; 0, n[dim - 1] = the whole lattice */
  int32_t velocity;   /* SPHB200_VEL_* */
  int32_t wall_axis, n_walls;
  float hot_lo, hot_hi, T_hot;
  float dx, rho, p, mass, eta, T, kappa, Cp; /* uniform fields, case_setup.py:152-181 */
} sphb200_lattice;

/* Rows sphb200_init_lattice writes (n0 * n1 * (k_hi - k_lo) in 3D), or a negative code. */
int64_t sphb200_lattice_rows(const sphb200_lattice *l);
/* Fill a state in the reference layout, DEVICE pointers; NULL members are skipped; dudt, dvdt,
 * drhodt, dTdt, nw are zeroed.  `ids` (int32, device) may be NULL. */
int sphb200_init_lattice(const sphb200_lattice *l, sphb200_state *out, int32_t *ids, void *stream);
/* u (and v, either may be NULL) = the case's velocity field at the given positions (device
 * pointers, (n, dim) float32): vmap(self._init_velocity2D / 3D)(r), case_setup.py:146-150, for
 * starts whose positions do not come from the lattice kernel (noise, relaxed states). */
int sphb200_eval_velocity(int32_t dim, int64_t n, int32_t velocity, const float *r, float *u,
                          float *v, void *stream);
/* Position noise of SimulationSetup.initialize() (case_setup.py:138-144, utils.py:120-125):
 * r += std * N(0, 1) on the FLUID particles (tag NULL: all), wrapped into the periodic box as
 * shift_fn does (jnp.mod(r + dr, side), space.py:207-209).  The deviates come from a
 * counter-based Philox-4x32-10 generator keyed by `seed` and the particle's row (`ids`, or the
 * array index when NULL), Box-Muller: the same start for any slab decomposition.  NOT
 * jax.random's threefry stream -- same distribution, other numbers. */
int sphb200_add_noise(int32_t dim, int64_t n, float *r, const int32_t *tag, const int32_t *ids,
                      double std, uint64_t seed, const double box[3], void *stream);

/* ---- stateless entry points (device pointers, caller-owned workspace) -----
 * Pure functions of their arguments: out = f(cfg, in); the workspace is scratch (>=
 * sphb200_workspace_bytes), nothing is kept in it after the call. */
int sphb200_workspace_bytes(const sphb200_config *cfg, int64_t n, size_t *bytes);
int sphb200_neighbor_list(const sphb200_config *cfg, int64_t n, const float *r, int32_t *idx,
                          int64_t capacity, int mask_self, int64_t *count, uint32_t *err,
                          void *workspace, size_t workspace_bytes, void *stream);
int sphb200_forward(const sphb200_config *cfg, int64_t n, const sphb200_state *in,
                    sphb200_state *out, uint32_t *err, void *workspace, size_t workspace_bytes,
                    void *stream);
int sphb200_advance(const sphb200_config *cfg, int64_t n, double dt, const sphb200_state *in,
                    sphb200_state *out, uint32_t *err, void *workspace, size_t workspace_bytes,
                    void *stream);
/* sphb200_advance for a caller that OWNS the workspace from call to call (memory nobody else
 * writes in between -- NOT an XLA scratch buffer; the jax.ffi shim keeps one allocation per
 * (cfg, n)).  Still out = f(cfg, in): the result never depends on what the
 * workspace held.  But when it is handed in again with the same cfg and n, the particles are
 * found cell-sorted in it with their neighbour lists: the new state goes into the slots its
 * particles occupy (sphb200_engine_refresh), every position is tested against the sorted one, and
 * the cell sort + neighbour search run only if the lists may have become incomplete (see
 * sphb200_engine_refresh) -- a state that has nothing to do with the previous call simply sorts
 * again.  The call then
 * costs a resident step plus the two state copies.  The library keeps a table workspace pointer
 * -> engine for this (mutex-protected, at most 8 workspaces); call
 * sphb200_workspace_release(ws) before the memory is freed or reused (NULL: all). */
int sphb200_advance_persistent(const sphb200_config *cfg, int64_t n, double dt,
                               const sphb200_state *in, sphb200_state *out, uint32_t *err,
                               void *workspace, size_t workspace_bytes, void *stream);
/* sphb200_advance_persistent for a caller that lets the ENGINE choose the row order of its state
 * arrays: row p of `in` is taken as the particle in slot p, row p of `out` is the particle the
 * step left in slot p, and in_order / out_order (int32[n]) carry the particles' labels through
 * the cell sorts (in_order == NULL: rows are labelled 0 .. n-1 as given; out_order is required).
 * The reference keeps the caller's order for ever (jax_sph/simulate.py:110-134 carries `state`
 * from advance() to advance()); a caller that feeds `out` / `out_order` of one call to the next
 * and un-permutes only where it looks at particles by index (io_state.write_state every
 * write_every steps: state[k][argsort(order)]) saves the two permuted copies of every step: the
 * copies in and out are then plain streaming passes and the call costs about what a resident
 * step costs.  Any row order is accepted -- out = f(in) row for row whatever the workspace held;
 * rows that do not lie where the workspace's cells expect them simply trigger a sort.  All
 * sixteen entries of `out` the caller wants must be non-NULL: a step that sorted has moved
 * every row, including the entries advance() does not change. */
int sphb200_advance_ordered(const sphb200_config *cfg, int64_t n, double dt, const sphb200_state *in,
                            const int32_t *in_order, sphb200_state *out, int32_t *out_order,
                            uint32_t *err, void *workspace, size_t workspace_bytes, void *stream);
void sphb200_workspace_release(void *workspace);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H_ */
