// engine.cu -- host side of libsphb200.so: planning, arena, launches, C ABI.
//
// The per-step pipeline (one call of si_euler.advance, integrator.py:22-56) of a resident engine:
//   k_drift       kick + drift + wrap in place; raises the step's re-sort flag when a particle is
//                 further than half the skin of the neighbour lists from where it was sorted
//   [flag set]    k_hash -> k_scan_* -> k_scatter_src -> k_reorder -> k_copyback   (cell sort)
//                 k_sweep<PhysNone, LIST_BUILD>                                     (the search)
//   k_sweep<PhysDensity, LIST_FILTER>  exact membership test + density + exact list of the step
//   [k_sweep<PhysRenorm>]  [k_sweep<PhysWall>]  k_sweep<PhysForce>  (LIST_CONSUME)  [k_bc]
// The flag lives on the device: every kernel of the bracketed group is launched every step and
// returns at once unless it is set, so a step never synchronises with the host.  Slab engines
// (one per GPU) sort and search every step: particles migrate between ranks (slab.cuh).
// The state never leaves HBM between steps; the original particle order is
// restored only by k_unpack (download).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>

#include "cells.cuh"
#include "common.cuh"
#include "init.cuh"
#include "phys.cuh"
#include "slab.cuh"
#include "sweep.cuh"
#include "sweep2.cuh"
#include "adjoint.cuh"

using namespace sphb200;

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) {                                                            \
      fprintf(stderr, "sphb200: %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SPHB200_ECUDA;                                                             \
    }                                                                                   \
  } while (0)

namespace {

constexpr size_t ALIGN = 256;
constexpr int DEFAULT_TPB = 512;  // tuned on B200, profiles/ (tune logs)
inline size_t up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

struct SweepPlan {
  int sb, cap, lcap;  // bytes staged per particle, staged particles, list entries per thread
  size_t smem;
};

}  // namespace

struct sphb200_engine {
  sphb200_config cfg;
  int n, dim;
  Grid grid;
  Consts consts;
  char* arena;
  size_t arena_bytes;
  bool own_arena;
  Frame fr[2];
  int cur;
  bool cells_valid;
  int *key, *rnk, *src, *count, *start, *bsum, *maxocc, *wallcount, *nl_counts;
  // neighbour lists (sweep.cuh, NList): pl_* the exact list of the step, sl_* the skin list
  unsigned short* pl_list;
  int* pl_cnt;
  unsigned char* pl_ok;
  int pl_lmax;
  unsigned short* sl_list;
  int* sl_cnt;
  // frozen sort (resident single-GPU engines): the particles keep their slots and the cell table
  // stays as it is until a particle is further than path_limit from where it was at the last sort
  bool inplace;
  double skin_frac;         // skin / cutoff
  float path_limit;         // < 0: sort + search every step
  bool rel_drift;           // re-sort criterion: relative drift of neighbours (cells.cuh, k_drift_box)
  float rel_limit;          // ... its limit (the skin); path_limit (half the skin, absolute) stays
                            // sufficient on its own, rel_guard is the absolute guard of the seam
  float rel_guard;
  DriftBlocks dblocks;
  float4* rb;               // [n] positions at the last sort
  int* ctl;                 // [0], [1] re-sort flag of even / odd steps, [2] searches so far
  unsigned long long step_no;
  AgreePtrs agree;  // sphb200_slab_set_agree: every rank's control block (peer-mapped), n = 0: off
  int ring_seq;     // exchanges signalled so far (sphb200_slab_signal)
  const int* gate_cur;      // flag word of the step being enqueued (nullptr: ungated)
  bool force_rebuild;       // the next step must sort + search (new state, lists stale)
  bool maybe_drifted;       // particles may have left the cells of the frozen table
  bool positions_replaced;  // a state was uploaded into the sorted slots: test it against rb
  int num_sms;
  unsigned* err;
  double* stats;  // [ekin, umax] of sphb200_engine_stats / the SPHB200_NSTATS words of _get_stats
  int nscan_blocks;
  int tpb, lcap;
  SweepPlan planA, planR, planW, planC, planN, planB;
  // duo sweeps (sweep2.cuh): two slot neighbours per thread on one union list; the headline SPH
  // variants (summation density, compact force record, no wall / renormalisation sweep)
  bool duo;
  int duo_tpb, duo_lmax, duo_rows, duo_desc_stride;
  float4 *rec0, *rec1, *rec2;  // compact force records of every slot
  float* rec_e;
  int* duo_desc;               // tile descriptors
  SweepPlan planDB, planDA;  // search, density filter (the force plan depends on tvf: forward_stage)
  bool has_kc, has_nw, has_ut, has_ge;
  int64_t launches;
  bool profile;
  cudaEvent_t ev[8];
  float times[8];
  // host staging for on_host upload / download
  char* hstage;
  size_t hstage_bytes;
  int max_smem;
  bool needs_zero;
  // slab decomposition (slab.cuh); slab_on == false: one GPU owns the whole periodic box
  bool slab_on;
  int slab_rank, slab_nranks, slab_axis;
  int slab_z0, slab_z1;    // global cell layers [z0, z1) owned along slab_axis
  Slab slab;               // base slot, device counters, migration buffers (set per call)
  SlabGeom sgeom;
  int* dn;
  int slab_stage;          // next sweep of the current step (0 density, 1 renorm, 2 wall, 3 force)
  int slab_pending_mask;   // arrays carried by the exchange in flight
  uint32_t slab_flags;
  bool slab_v_is_u;
  Kick slab_kick;
  // Slab overlap: while a halo message is in flight the INTERIOR tile layers of the next big
  // sweep (density or force; tiles whose stencil touches no halo layer) run on a side stream.
  int part;            // 0: all tiles, 1: interior tile layers only, 2: boundary tile layers only
  int int_lo, int_hi;  // interior tile layers [int_lo, int_hi) along the slab axis
  cudaStream_t side;
  cudaEvent_t ev_main, ev_side;
  bool overlap;       // side stream and events exist (slab engines)
  // advance_host: the entries that are final after the integrate / reorder pass (r, and u, v
  // when no wall sweep or bc table rewrites them) go back to the host on this stream WHILE the
  // sweeps run
  cudaStream_t io_side;
  cudaEvent_t ev_io_fork, ev_io_join;
  bool io_on;
  int pre_stage;      // stage whose interior tiles are already running on the side stream, -1
  int delta_sub;      // slab engine, Delta-SPH density diffusion: next of its three sweeps
  bool stage_more;    // forward_stage ran only part of the stage (another halo refresh first)
  // wall-normal recomputation for moving walls (utils.py:197-277): the static one-layer
  // discretisation of the wall surface; wl_n == 0: normals are an input that never changes
  float4* adj_mem;  // work arrays of sphb200_engine_vjp (4 n quads), allocated on first use
  float4* wl_pts;
  int wl_n;
  float wl_off[3];
  float wl_c2;
};

namespace {

double kernel_cutoff(const sphb200_config& c) {
  if (c.r_cutoff > 0.0) return c.r_cutoff;
  const bool wide = c.kernel == SPHB200_KERNEL_QSK || c.kernel == SPHB200_KERNEL_GK ||
                    c.kernel == SPHB200_KERNEL_SGK;  // kernel.py:56, :172, :190: 3 h, else 2 h
  return (wide ? 3.0 : 2.0) * c.h;
}

int validate(const sphb200_config* c, int64_t n) {
  if (!c || c->struct_size != sizeof(sphb200_config)) return SPHB200_EINVAL;
  if (c->dim != 2 && c->dim != 3) return SPHB200_EINVAL;
  if (n <= 0 || n > 2000000000LL) return SPHB200_EINVAL;
  if (c->solver != SPHB200_SOLVER_SPH && c->solver != SPHB200_SOLVER_RIE &&
      c->solver != SPHB200_SOLVER_DELTA)
    return SPHB200_EUNSUP;

  if (c->kernel < SPHB200_KERNEL_QSK || c->kernel > SPHB200_KERNEL_SGK) return SPHB200_EUNSUP;
  if (c->eos != SPHB200_EOS_TAIT && c->eos != SPHB200_EOS_RIEMANN) return SPHB200_EINVAL;
  if (!(c->h > 0) || !(c->dx > 0)) return SPHB200_EINVAL;
  for (int a = 0; a < c->dim; ++a)
    if (!(c->box[a] > 0)) return SPHB200_EINVAL;
  if (c->g_mode < 0 || c->g_mode > 3) return SPHB200_EINVAL;
  if (c->g_mode == SPHB200_G_BAND && (c->g_axis < 0 || c->g_axis >= c->dim)) return SPHB200_EINVAL;
  return SPHB200_OK;
}

// Global layer range [z0, z1) of rank r of p along an axis of ng cell layers.
void slab_range(int ng, int r, int p, int& z0, int& z1) {
  z0 = (int)((long long)ng * r / p);
  z1 = (int)((long long)ng * (r + 1) / p);
}

// Largest record (bytes per staged particle) any sweep of this solver variant stages: the
// records of phys.cuh (PhysDensity / PhysDelta / PhysRenorm / PhysWall / PhysForce).
bool duo_variant(const sphb200_config& c);
int max_stage_bytes(const sphb200_config& c) {
  const bool rie = c.solver == SPHB200_SOLVER_RIE, delta = c.solver == SPHB200_SOLVER_DELTA;
  const bool bc_trick = c.flags & SPHB200_F_BC_TRICK, evol = c.flags & SPHB200_F_RHO_EVOL,
             renorm = c.flags & SPHB200_F_RHO_RENORM, free_slip = c.flags & SPHB200_F_FREE_SLIP,
             heat = c.flags & SPHB200_F_HEAT;
  const bool has_ut = rie && bc_trick && !free_slip;
  int m = !evol ? ((has_ut || (rie && bc_trick && heat)) ? 48 : 16)
                : (delta ? (c.dim == 3 ? 64 : 48) : (!rie ? 32 : 64));
  const int force_nq = 3 + ((!rie && c.tvf != 0.0) ? 1 : 0) + (heat ? 1 : 0) + (rie ? 1 : 0) +
                       (has_ut ? 1 : 0);
  const bool generic = rie || heat || c.artificial_alpha != 0.0 || delta;
  // (the duo force sweep under SPHB200_HINT_UNIFORM_ETA stages three quads, no eta column)
  const bool eta_u = (c.hints & SPHB200_HINT_UNIFORM_ETA) && duo_variant(c);
  const int force = generic ? 16 * force_nq : (c.tvf != 0.0 ? (eta_u ? 48 : 52) : (duo_variant(c) ? 48 : 40));
  if (force > m) m = force;
  if (bc_trick && !rie && 64 > m) m = 64;  // PhysWall
  if (evol && renorm && 32 > m) m = 32;    // PhysRenorm
  return m;
}

// Cell grid, stencil and tiling (host).  nranks > 1: the local view of rank `rank`.
// skin / cutoff of an engine for this config
bool duo_variant(const sphb200_config& c);
double plan_skin(const sphb200_config& c) {
  if (c.nl_cap < 0 || c.skin < 0.f) return 0.0;
  if (c.skin > 0.f) return c.skin > 1.f ? 1.0 : (double)c.skin;
  // tuned on B200 (profiles/r02_plan_tuning_*): the cheaper search + filter of the 3D duo sweeps
  // move the optimum out a little
  return (c.dim == 3 && duo_variant(c)) ? 0.12 : 0.10;
}

void plan_grid(const sphb200_config& c, Grid& g, int tpb, int rank = 0, int nranks = 1,
               double skin_frac = 0.0, int duo_tpb = 0) {
  const double cutoff = kernel_cutoff(c) * (1.0 + skin_frac);  // what the cells and the search cover
  double pop = 1.0;
  g.exact_all = 0;
  g.imargin = 0;
  g.ncells = 1;
  for (int a = 0; a < 3; ++a) {
    if (a < c.dim) {
      int sub = c.cell_sub[a] > 0 ? c.cell_sub[a] : 2;  // cells of half a cutoff by default
      if (sub > 4) sub = 4;
      int n = (int)floor(c.box[a] * sub / (cutoff * 1.001));
      if (n < 1) n = 1;
      g.n[a] = n;
      g.S[a] = sub;
      g.W[a] = (n >= 2 * sub + 1) ? 2 * sub + 1 : n;
      if (n < 2 * sub + 1) g.exact_all = 1;
      g.inv_cell[a] = (float)(n / c.box[a]);
      g.box[a] = (float)c.box[a];
      g.half[a] = g.box[a] * 0.5f;
      pop *= (c.box[a] / n) / c.dx;
    } else {
      g.n[a] = 1; g.S[a] = 0; g.W[a] = 1;
      g.inv_cell[a] = 0.f; g.box[a] = 1.f; g.half[a] = 0.5f;
    }
    g.ncells *= g.n[a];
  }
  // tile: T1 = T2 = sub (so a tile is about one cutoff wide across), T0 fills the block
  for (int a = 1; a < 3; ++a) {
    int t = c.tile[a] > 0 ? c.tile[a] : (a < c.dim ? 2 * g.S[a] : 1);
    if (t > g.n[a]) t = g.n[a];
    g.T[a] = t;
  }
  while (g.T[1] * g.T[2] > MAX_RUNS) (g.T[2] > 1 ? g.T[2] : g.T[1])--;
  // particles a block wants: one per thread, or two per thread of a duo sweep (with room for the
  // half-filled duo that ends every odd run)
  const double want_pop = duo_tpb > 0 ? 0.90 * 2.0 * duo_tpb : 0.93 * tpb;
  int t0 = c.tile[0] > 0 ? c.tile[0] : (int)floor(want_pop / (pop * g.T[1] * g.T[2]) + 0.5);
  if (t0 < 1) t0 = 1;
  if (t0 > g.n[0]) t0 = g.n[0];
  // MAX_SOFF bound on staged (row, cell) entries
  auto entries = [&](int t) {
    int rows = 1;
    for (int a = 1; a < 3; ++a)
      rows *= (g.n[a] >= 2 * g.S[a] + 1) ? g.T[a] + 2 * g.S[a] : g.n[a];
    int nxs = (g.n[0] >= 2 * g.S[0] + 1) ? t + 2 * g.S[0] : g.n[0];
    return rows * nxs;
  };
  while (t0 > 1 && entries(t0) > MAX_SOFF) --t0;
  if (c.tile[0] <= 0) {
    // The step's shared neighbour lists need the tile's stencil to fit ONE staging group of the
    // sweep with the largest record (sweep.cuh, NList): shorten the tile until the expected
    // stencil population (+8 % for disorder) fits what 227 KB of shared memory hold next to the
    // minimum per-thread lists.
    const double fit = ((227.0 * 1024 - 1024) - (duo_tpb > 0 ? (double)duo_smem_bytes(0, 0, 0, 0)
                                                             : (double)sweep_smem_bytes(0, 0, 24, tpb))) /
                       max_stage_bytes(c);
    // (duo sweeps have no second staging group to fall back on inside the kernel: a stencil that
    // does not fit is swept by the slow per-particle search, so the bound also covers the most
    // a LATTICE start can put into the stencil, one more plane per axis than the mean)
    auto lattice_max = [&](int t) {
      double m = 1.0;
      for (int a = 0; a < c.dim; ++a) {
        const int len = a == 0 ? ((g.n[0] >= 2 * g.S[0] + 1) ? t + 2 * g.S[0] : g.n[0])
                               : ((g.n[a] >= 2 * g.S[a] + 1) ? g.T[a] + 2 * g.S[a] : g.n[a]);
        m *= floor(len * (c.box[a] / g.n[a]) / c.dx) + 1.0;
      }
      return m + 1.0;
    };
    while (t0 > 2 && (entries(t0) * pop * 1.08 > fit || (duo_tpb > 0 && lattice_max(t0) > fit))) --t0;
  }
  g.T[0] = t0;
  for (int a = 0; a < 3; ++a) {
    g.goff[a] = 0;
    g.ng[a] = g.n[a];
    g.own_lo[a] = 0;
    g.own_hi[a] = g.n[a];
  }
  if (nranks > 1) {
    const int ax = c.dim - 1;
    int z0, z1;
    slab_range(g.n[ax], rank, nranks, z0, z1);
    g.goff[ax] = z0 - g.S[ax];
    g.own_lo[ax] = g.S[ax];
    g.own_hi[ax] = g.S[ax] + (z1 - z0);
    g.n[ax] = (z1 - z0) + 2 * g.S[ax];
    g.W[ax] = 2 * g.S[ax] + 1;
    if (g.T[ax] > z1 - z0) g.T[ax] = z1 - z0;
    g.ncells = g.n[0] * g.n[1] * g.n[2];
  }
  for (int a = 0; a < 3; ++a) g.nt[a] = (g.own_hi[a] - g.own_lo[a] + g.T[a] - 1) / g.T[a];

  const float cf = (float)kernel_cutoff(c);
  g.c2 = cf * cf;  // float32 product, jax_md/partition.py:820-822
  g.block0 = 0;
  g.ntl = g.nt[0] * g.nt[1] * g.nt[2];
  if (g.exact_all) {
    g.c2_hi = INFINITY;
    g.c2_lo = -1.0f;
    g.c2_fb = INFINITY;
  } else {
    double side = 0;
    for (int a = 0; a < c.dim; ++a) side = fmax(side, c.box[a]);
    const double errd = 16.0 * 1.1920929e-7 * side * sqrt((double)c.dim);
    const double chi = cutoff + errd, clo = fmax(0.0, cutoff - errd);
    g.c2_hi = (float)(chi * chi * (1.0 + 1e-6));
    g.c2_lo = (float)(clo * clo * (1.0 - 1e-6));
    g.c2_fb = skin_frac > 0.0 ? INFINITY : g.c2_hi;
  }
}

void plan_consts(const sphb200_config& c, Consts& k) {
  memset(&k, 0, sizeof(k));
  k.dim = c.dim; k.solver = c.solver; k.kernel = c.kernel; k.eos = c.eos; k.flags = c.flags;
  k.eps = 1.1920928955078125e-07f;
  const double ooh = 1.0 / c.h;
  double sigma;
  const double oohd = c.dim == 2 ? ooh * ooh : ooh * ooh * ooh;
  switch (c.kernel) {  // kernel.py: the _sigma of each class for dim 2 / 3
    case SPHB200_KERNEL_QSK: sigma = (c.dim == 2 ? 7.0 / 478.0 : 3.0 / 359.0) / M_PI * oohd; break;
    case SPHB200_KERNEL_WC2K: sigma = (c.dim == 2 ? 7.0 / 4.0 : 21.0 / 16.0) / M_PI * oohd; break;
    case SPHB200_KERNEL_CSK: sigma = (c.dim == 2 ? 10.0 / 7.0 : 1.0) / M_PI * oohd; break;
    case SPHB200_KERNEL_WC4K: sigma = (c.dim == 2 ? 9.0 / 4.0 : 495.0 / 256.0) / M_PI * oohd; break;
    case SPHB200_KERNEL_WC6K: sigma = (c.dim == 2 ? 78.0 / 28.0 : 1365.0 / 512.0) / M_PI * oohd; break;
    default: sigma = 1.0 / pow(M_PI, c.dim / 2.0) * oohd; break;  // GK, SGK
  }
  k.ooh = (float)ooh;
  k.sigma = (float)sigma;
  k.sigma_ooh = k.sigma * k.ooh;
  k.gwk[0] = k.sigma_ooh * -5.0f;
  k.gwk[1] = k.sigma_ooh * 30.0f;
  k.gwk[2] = k.sigma_ooh * -75.0f;
  k.dt_s = (float)c.dt;
  k.p_ref = (float)c.p_ref; k.rho_ref = (float)c.rho_ref; k.p_bg = (float)c.p_bg;
  k.gamma = (float)c.gamma;
  k.inv_gamma = (float)(1.0 / c.gamma);
  k.c100 = (float)(100.0 * c.u_ref * c.u_ref);
  if (c.eos == SPHB200_EOS_TAIT) {
    // p_fn(0) = p_ref * ((0/rho_ref)^gamma - 1) + p_bg in float32
    k.p_bg_tvf = k.p_ref * (0.0f - 1.0f) + k.p_bg;
  } else {
    k.p_bg_tvf = k.c100 * (0.0f - k.rho_ref) + k.p_bg;
  }
  k.c_ref = (float)c.c_ref;
  k.use_lim = !(c.eta_limiter == -1.0);
  k.eta_lim = (float)c.eta_limiter;
  k.av_coef = (float)(c.artificial_alpha * c.dx * 10.0);  // h_ab = dx, c_ab = 10 * 1.0
  k.av_eps = (float)(0.01 * c.dx * c.dx);
  k.delta_coef = (float)c.diff_alpha * (float)c.h * (float)c.c_ref * (float)c.rho_ref;
  k.delta_rho = (float)c.c_ref * (float)c.diff_delta * (float)c.h;
  k.g_mode = c.g_mode; k.g_axis = c.g_axis;
  for (int a = 0; a < 3; ++a) k.g[a] = (float)c.g[a];
  k.g_lo = (float)c.g_lo; k.g_hi = (float)c.g_hi;
  for (int t = 0; t < 4; ++t) k.bc[t] = c.bc[t];
  k.inflow_on = c.bc_inflow_on; k.inflow_x = c.bc_inflow_x; k.inflow_T = c.bc_inflow_T;
  k.outflow_on = c.bc_outflow_on; k.outflow_x = c.bc_outflow_x;
  k.any_walls = 1;
}

bool bc_table_on(const sphb200_config& c) {
  for (int t = 0; t < 4; ++t)
    if (c.bc[t].flags) return true;
  return c.bc_inflow_on || c.bc_outflow_on;
}

// duo sweeps (sweep2.cuh): whether this engine uses them and with what
struct DuoPlan {
  bool on;
  int tpb, lmax, rows, desc_stride;
};

struct Layout {
  size_t frame[2][12];
  size_t key, rnk, src, count, start, bsum, maxocc, wallcount, err, stats, nl_counts, ut;
  size_t pl_list, pl_cnt, pl_ok, sl_list, sl_cnt, path, ctl;
  size_t dl, dg;  // Delta-SPH density diffusion (PhysDelta)
  size_t rec, desc;  // duo sweeps: compact force records (52 B per slot), tile descriptors
  size_t dbox;       // duo engines: drift boxes of the S^3-cell blocks (cells.cuh, k_drift_box)
  int pl_lmax;
  size_t dn;
  size_t total;
};

void feature_flags(const sphb200_config& c, bool& kc, bool& nw, bool& ut, bool& ge) {
  kc = (c.flags & SPHB200_F_HEAT) != 0;
  nw = (c.solver == SPHB200_SOLVER_RIE) || (c.flags & SPHB200_F_FREE_SLIP);
  ut = (c.solver == SPHB200_SOLVER_RIE) && (c.flags & SPHB200_F_BC_TRICK) &&
       !(c.flags & SPHB200_F_FREE_SLIP);
  ge = c.g_mode == SPHB200_G_ARRAY;
}

// Row length of the per-step neighbour lists: 1.3 x the expected number of
// neighbours of a particle in a uniform fluid (+8), a multiple of 8; 0 = lists off.
int plan_lmax(const sphb200_config& c, double skin_frac) {
  if (c.nl_cap < 0) return 0;
  if (c.nl_cap > 0) return (c.nl_cap + 7) / 8 * 8;
  const double q = kernel_cutoff(c) * (1.0 + skin_frac) / c.dx;
  const double expect = c.dim == 2 ? M_PI * q * q : 4.0 / 3.0 * M_PI * q * q * q;
  int lmax = ((int)(1.3 * expect) + 8 + 7) / 8 * 8;
  if (lmax > 1024) lmax = 0;  // very wide kernels: not worth the memory
  return lmax;
}

void plan_layout(const sphb200_config& c, int64_t n, const Grid& g, double skin_frac, Layout& L,
                 const DuoPlan& dp) {
  bool kc, nw, ut, ge;
  feature_flags(c, kc, nw, ut, ge);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += up(bytes);
    return o;
  };
  for (int f = 0; f < 2; ++f) {
    for (int q = 0; q < 6; ++q) L.frame[f][q] = take((size_t)n * 16);  // pt um vv st du dv
    L.frame[f][6] = take((size_t)n * 4);                              // id
    L.frame[f][7] = kc ? take((size_t)n * 8) : (size_t)-1;
    L.frame[f][8] = nw ? take((size_t)n * 16) : (size_t)-1;
    L.frame[f][9] = ge ? take((size_t)n * 16) : (size_t)-1;
  }
  L.ut = ut ? take((size_t)n * 16) : (size_t)-1;
  const bool delta_evol = c.solver == SPHB200_SOLVER_DELTA && (c.flags & SPHB200_F_RHO_EVOL);
  L.dl = delta_evol ? take((size_t)n * 48) : (size_t)-1;
  L.dg = delta_evol ? take((size_t)n * 32) : (size_t)-1;
  L.key = take((size_t)n * 4);
  L.rnk = take((size_t)n * 4);
  L.src = take((size_t)n * 4);
  L.nl_counts = take(((size_t)n + 1) * 4);
  L.count = take((size_t)g.ncells * 4);
  L.start = take(((size_t)g.ncells + 1) * 4);
  size_t nb = ((size_t)(n > g.ncells ? n : g.ncells) + SCAN_TILE) / SCAN_TILE + 1;
  L.bsum = take(nb * 4);
  L.maxocc = take(4);
  L.wallcount = take(4);
  L.err = take(4);
  L.stats = take(SPHB200_NSTATS * 8);
  L.dn = take(DN_WORDS * 4);
  L.path = take((size_t)n * 16);
  L.ctl = take(8 * 4);
  L.pl_lmax = plan_lmax(c, skin_frac);
  if (L.pl_lmax > 0) {
    L.pl_list = take((size_t)n * L.pl_lmax * 2);
    L.pl_cnt = take((size_t)n * 4);
    L.sl_list = take((size_t)n * L.pl_lmax * 2);
    L.sl_cnt = take((size_t)n * 4);
    L.pl_ok = take((size_t)g.ncells);  // one flag per tile; sized for any tiling of the cells
  } else {
    L.pl_list = L.pl_cnt = L.pl_ok = L.sl_list = L.sl_cnt = (size_t)-1;
  }
  L.rec = L.desc = L.dbox = (size_t)-1;
  if (dp.on) {
    L.rec = take((size_t)n * 52);
    L.desc = take((size_t)g.nt[0] * g.nt[1] * g.nt[2] * dp.desc_stride * 4);
    size_t blocks = 1;
    for (int a = 0; a < c.dim; ++a) blocks *= (size_t)((g.n[a] + g.S[a] - 1) / g.S[a]);
    L.dbox = take(blocks * 32 * 3);  // the blocks' boxes + two buffers of the axis passes
  }
  L.total = off;
}

// lcap: per-thread pair-list entries in shared memory.  The sweeps that search (density =
// list builder, the materialiser) want it long -- fewer, more even flushes; a list consumer
// only searches in its fall-back path, so it gets the minimum and the staging buffer the rest.
SweepPlan plan_sweep_bytes(const sphb200_engine* e, int sb, int lcap);
SweepPlan plan_sweep(const sphb200_engine* e, int nq, int lcap) { return plan_sweep_bytes(e, 16 * nq, lcap); }
// sb: bytes staged per particle (a multiple of 4; capacities are multiples of 32 particles, so
// whatever follows the staging buffer stays 16-byte aligned)
SweepPlan plan_sweep_bytes(const sphb200_engine* e, int sb, int lcap) {
  const Grid& g = e->grid;
  const sphb200_config& c = e->cfg;
  double pop = 1.0;
  for (int a = 0; a < c.dim; ++a) pop *= (c.box[a] / g.ng[a]) / c.dx;  // global cells: n[] is the slab-local view
  int rows = 1;
  for (int a = 1; a < 3; ++a) rows *= (g.n[a] >= 2 * g.S[a] + 1) ? g.T[a] + 2 * g.S[a] : g.n[a];
  int nxs = (g.n[0] >= 2 * g.S[0] + 1) ? g.T[0] + 2 * g.S[0] : g.n[0];
  long long want = (long long)(rows * (double)nxs * pop * 1.3) + 64;
  if (c.stage_cap > 0) want = c.stage_cap;
  if (want > e->n + 32) want = e->n + 32;  // never more than (a few images of) everything
  size_t fixed = sweep_smem_bytes(sb, 0, lcap, e->tpb);
  long long fit = ((long long)e->max_smem - (long long)fixed) / (long long)sb;
  while (fit < want && lcap > 24) {  // the staging buffer comes first: a stencil that does not
    lcap -= 8;                       // fit is swept in several groups and gets no shared lists
    fixed = sweep_smem_bytes(sb, 0, lcap, e->tpb);
    fit = ((long long)e->max_smem - (long long)fixed) / (long long)sb;
  }
  if (want > fit) want = fit;
  if (want > 65535) want = 65535;
  want = want / 32 * 32;
  if (want < 32) want = 32;
  SweepPlan p;
  p.sb = sb;
  p.cap = (int)want;
  p.lcap = lcap;
  p.smem = sweep_smem_bytes(sb, p.cap, lcap, e->tpb);
  return p;
}

// gate: device flag word the launch is conditional on (the search of a resident engine: a small
// persistent grid, so that a gated-off launch costs one block exit per SM); fresh: the cell table
// was rebuilt from the current positions just before (the fall-back search may use its cheap
// reject although the engine's lists carry a skin).
template <class K>
int launch_sweep(sphb200_engine* e, K kern, const SweepPlan& sp, const Frame& f, const Extra& ex,
                 cudaStream_t st,
                 const NList& nl = NList{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr},
                 const int* gate = nullptr, bool persistent = false, bool fresh = false,
                 int gate_want = 1) {
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
  // a sweep may stage fewer bytes than its plan was sized for, never more
  const int sb = ex.sb > 0 ? ex.sb : 16 * ex.nq;
  if (sb > sp.sb) return SPHB200_EINVAL;
  SweepDims sd{sp.cap, sp.lcap, sb, gate, gate_want};
  const int all = e->grid.nt[0] * e->grid.nt[1] * e->grid.nt[2];
  Grid g = e->grid;
  if (fresh) g.c2_fb = g.c2_hi;
  int resident = 0;
  if (persistent) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, e->tpb, sp.smem));
    resident = (per_sm > 0 ? per_sm : 1) * e->num_sms;
  }
  // tile layers along the slab axis (the slowest tile index) covered by this launch
  const int ax = e->dim - 1, nl_ax = g.nt[ax], per_layer = all / nl_ax;
  int ranges[2][2] = {{0, nl_ax}, {0, 0}};
  if (e->part == 1) {
    ranges[0][0] = e->int_lo; ranges[0][1] = e->int_hi;
  } else if (e->part == 2) {
    ranges[0][0] = 0; ranges[0][1] = e->int_lo;
    ranges[1][0] = e->int_hi; ranges[1][1] = nl_ax;
  }
  for (int r = 0; r < 2; ++r) {
    const int blocks = (ranges[r][1] - ranges[r][0]) * per_layer;
    if (blocks <= 0) continue;
    g.block0 = ranges[r][0] * per_layer;
    g.ntl = blocks;
    const int grid = (persistent && blocks > resident) ? resident : blocks;
    kern<<<grid, e->tpb, sp.smem, st>>>(g, e->consts, f, e->start, sd, ex, e->err, nl);
    e->launches++;
  }
  CK(cudaGetLastError());
  return SPHB200_OK;
}

// duo sweeps (sweep2.cuh): same tile ranges and gate protocol as launch_sweep
template <class K>
int launch_duo(sphb200_engine* e, K kern, const SweepPlan& sp, const Frame& f, const Extra& ex,
               cudaStream_t st, const DuoList& dl, const int* gate = nullptr, bool persistent = false,
               int split = 1) {
  const int threads = e->duo_tpb * split;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
  const int sb = ex.sb > 0 ? ex.sb : 16 * ex.nq;
  if (sb > sp.sb) return SPHB200_EINVAL;
  SweepDims sd{sp.cap, sp.lcap, sb, gate, 1};
  const int all = e->grid.nt[0] * e->grid.nt[1] * e->grid.nt[2];
  Grid g = e->grid;
  int resident = 0;
  if (persistent) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, sp.smem));
    resident = (per_sm > 0 ? per_sm : 1) * e->num_sms;
  }
  const int ax = e->dim - 1, nl_ax = g.nt[ax], per_layer = all / nl_ax;
  int ranges[2][2] = {{0, nl_ax}, {0, 0}};
  if (e->part == 1) {
    ranges[0][0] = e->int_lo; ranges[0][1] = e->int_hi;
  } else if (e->part == 2) {
    ranges[0][0] = 0; ranges[0][1] = e->int_lo;
    ranges[1][0] = e->int_hi; ranges[1][1] = nl_ax;
  }
  for (int r = 0; r < 2; ++r) {
    const int blocks = (ranges[r][1] - ranges[r][0]) * per_layer;
    if (blocks <= 0) continue;
    g.block0 = ranges[r][0] * per_layer;
    g.ntl = blocks;
    const int grid = (persistent && blocks > resident) ? resident : blocks;
    kern<<<grid, threads, sp.smem, st>>>(g, e->consts, f, e->start, sd, ex, e->err, dl);
    e->launches++;
  }
  CK(cudaGetLastError());
  return SPHB200_OK;
}

// staging capacity of a duo sweep: what the tile's stencil is expected to hold (+30 %), bounded by
// shared memory and by the 14 index bits of a list entry
SweepPlan plan_duo(const sphb200_engine* e, int sb, int lcap, int blocks_per_sm = 1) {
  const Grid& g = e->grid;
  const sphb200_config& c = e->cfg;
  double pop = 1.0;
  for (int a = 0; a < c.dim; ++a) pop *= (c.box[a] / g.ng[a]) / c.dx;
  int rows = 1;
  for (int a = 1; a < 3; ++a) rows *= (g.n[a] >= 2 * g.S[a] + 1) ? g.T[a] + 2 * g.S[a] : g.n[a];
  const int nxs = (g.n[0] >= 2 * g.S[0] + 1) ? g.T[0] + 2 * g.S[0] : g.n[0];
  long long want = (long long)(rows * (double)nxs * pop * 1.3) + 64;
  if (c.stage_cap > 0) want = c.stage_cap;
  if (want > e->n + 32) want = e->n + 32;
  // (a block's share of the SM's shared memory when several are to be resident: 1 KB each is
  // the system's)
  const long long room = ((long long)e->max_smem + 1024) / blocks_per_sm - 1024 - 64;
  const int desc_ints = lcap > 0 ? 0 : e->duo_desc_stride;  // (lcap > 0: the search, DUO_BUILD)
  const long long fit = (room - (long long)duo_smem_bytes(0, 0, lcap, e->duo_tpb, desc_ints)) / sb;
  if (want > fit) want = fit;
  if (want > DUO_IDX + 1) want = DUO_IDX + 1;
  want = want / 32 * 32;
  if (want < 32) want = 32;
  SweepPlan p;
  p.sb = sb;
  p.cap = (int)want;
  p.lcap = lcap;
  p.smem = duo_smem_bytes(sb, p.cap, lcap, e->duo_tpb, desc_ints);
  return p;
}

// the solver variants the duo sweeps cover: SPH with summation density and the compact force
// record (phys.cuh: DENS_SUM, FORCE_PLAIN / FORCE_TVF), nothing between the two sweeps
bool duo_variant(const sphb200_config& c) {
  const char* env = getenv("SPHB200_DUO");
  if (env && env[0] == '0') return false;
  if (c.solver != SPHB200_SOLVER_SPH || c.nl_cap < 0) return false;
  if (c.flags & (SPHB200_F_BC_TRICK | SPHB200_F_RHO_EVOL | SPHB200_F_RHO_RENORM | SPHB200_F_HEAT |
                 SPHB200_F_FREE_SLIP))
    return false;
  return c.artificial_alpha == 0.0;
}

Extra make_extra() {
  Extra ex;
  memset(&ex, 0, sizeof(ex));
  ex.q_v = ex.q_h = ex.q_nw = ex.q_ut = -1;
  ex.nq = 1;
  return ex;
}

#define DISPATCH_DK(e, CALL)                                                       \
  do {                                                                             \
    if ((e)->dim == 2) {                                                           \
      if ((e)->cfg.kernel == SPHB200_KERNEL_QSK) { CALL(2, SPHB200_KERNEL_QSK); }  \
      else if ((e)->cfg.kernel == SPHB200_KERNEL_WC2K) { CALL(2, SPHB200_KERNEL_WC2K); } \
      else { CALL(2, KERN_ANY); }                                                  \
    } else {                                                                       \
      if ((e)->cfg.kernel == SPHB200_KERNEL_QSK) { CALL(3, SPHB200_KERNEL_QSK); }  \
      else if ((e)->cfg.kernel == SPHB200_KERNEL_WC2K) { CALL(3, SPHB200_KERNEL_WC2K); } \
      else { CALL(3, KERN_ANY); }                                                  \
    }                                                                              \
  } while (0)

void swap_st(sphb200_engine* e) {
  float4* t = e->fr[0].st;
  e->fr[0].st = e->fr[1].st;
  e->fr[1].st = t;
}

// nw_fn of the integrator (integrator.py:33-34, compute_nws_jax_wrapper utils.py:197-277):
// every wall particle takes the direction to the closest particle of the wall-surface layer
// among those within the list cutoff; everything else gets a zero normal.  The layer is a
// few thousand points (5 per dx of wall length), scanned linearly by each wall particle --
// the case that needs this (a MOVING_WALL with free-slip or Riemann walls) is the 2D/3D
// Couette channel, O(walls x layer) is a fraction of one sweep there.  Reference quirk kept:
// layer particle 0 is never matched (`idx > len(r_walls)`, utils.py:252).
template <int DIM>
__global__ void __launch_bounds__(256)
    k_wall_normals(int n, Grid g, Slab sl, Consts c, Frame f, const float4* __restrict__ layer,
                   int nl, float ox, float oy, float oz, float c2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (sl.dn ? sl.dn[DN_OWN] : n)) return;
  const int p = sl.base + t;
  const float4 pt = f.pt[p];
  float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
  if (is_wall_tag(__float_as_int(pt.w))) {
    const float rw[3] = {__fsub_rn(pt.x, ox), __fsub_rn(pt.y, oy), __fsub_rn(pt.z, oz)};
    float best = __int_as_float(0x7f800000), bd[3] = {0.f, 0.f, 0.f};
    for (int j = 1; j < nl; ++j) {
      const float4 q = __ldg(layer + j);
      float d[3];
      d[0] = __fsub_rn(mod_side(__fadd_rn(__fsub_rn(q.x, rw[0]), g.half[0]), g.box[0]), g.half[0]);
      d[1] = __fsub_rn(mod_side(__fadd_rn(__fsub_rn(q.y, rw[1]), g.half[1]), g.box[1]), g.half[1]);
      d[2] = DIM == 3
                 ? __fsub_rn(mod_side(__fadd_rn(__fsub_rn(q.z, rw[2]), g.half[2]), g.box[2]), g.half[2])
                 : 0.f;
      const float d2 = sumsq<DIM>(d);
      if (d2 < c2 && d2 < best) {  // first minimum wins, like argmin
        best = d2;
        bd[0] = d[0]; bd[1] = d[1]; bd[2] = d[2];
      }
    }
    if (best < __int_as_float(0x7f800000)) {
      const float den = __fadd_rn(best > 0.f ? __fsqrt_rn(best) : 0.f, c.eps);
      out = make_float4(__fdiv_rn(bd[0], den), __fdiv_rn(bd[1], den),
                        DIM == 3 ? __fdiv_rn(bd[2], den) : 0.f, 0.f);
    }
  }
  f.nw[p] = out;
}

// first half of the cell pipeline: kick + drift + wrap (recomputed later, not stored), cell
// key, arrival rank, histogram; slab mode: emigrants leave through e->slab.mig_*
// blocks of a grid-stride streaming kernel over `bound` particles
int stream_blocks(const sphb200_engine* e, int bound) {
  const int nb = (bound + 255) / 256, cap = e->num_sms * 16;
  return nb < cap ? (nb > 0 ? nb : 1) : cap;
}

int hash_cells(sphb200_engine* e, const Kick& k, cudaStream_t st, const int* gate = nullptr) {
  const int bound = e->slab_on ? e->sgeom.own_cap : e->n;
  const int nb = stream_blocks(e, bound);
  Frame& A = e->fr[e->cur];
  if (e->dim == 2)
    k_hash<2><<<nb, 256, 0, st>>>(bound, e->grid, k, e->slab, A, e->key, e->rnk, e->count, e->err, gate);
  else
    k_hash<3><<<nb, 256, 0, st>>>(bound, e->grid, k, e->slab, A, e->key, e->rnk, e->count, e->err, gate);
  e->launches += 1;
  CK(cudaGetLastError());
  return SPHB200_OK;
}

// second half: cell table, slot -> source, stable reorder into the other frame
// nw_fn of the integrator on the current frame (integrator.py:33-34: after the drift)
int wall_normals(sphb200_engine* e, cudaStream_t st);

int sort_cells(sphb200_engine* e, const Kick& k, cudaStream_t st, const int* gate = nullptr,
               bool keep_acc = false) {
  // slab mode: sources are own + immigrants; the new own count is only known on the device
  const int bound = e->slab_on ? e->sgeom.own_cap + 2 * e->slab.mig_cap : e->n;
  const int nb = stream_blocks(e, bound);
  Frame& A = e->fr[e->cur];
  Frame& B = e->fr[1 - e->cur];
  const int c = e->grid.ncells;
  const int sb = (c + SCAN_TILE) / SCAN_TILE;  // covers index c itself (start[c] = total)
  k_scan_partial<<<sb, SCAN_TPB, 0, st>>>(c, e->count, e->bsum, gate);
  k_scan_bsum<<<1, 1024, 0, st>>>(sb, e->bsum, gate);
  k_scan_final<<<sb, SCAN_TPB, 0, st>>>(c, e->slab.base, e->count, e->bsum, e->start, e->maxocc, gate);
  k_scatter_src<<<nb, 256, 0, st>>>(bound, e->slab, e->key, e->rnk, e->start, e->src, gate);
  ReorderOpt o{e->has_kc ? 1 : 0, e->has_nw ? 1 : 0, e->has_ge ? 1 : 0, keep_acc ? 1 : 0};
  if (e->dim == 2)
    k_reorder<2><<<nb, 256, 0, st>>>(bound, e->grid, k, e->slab, o, A, B, e->key, e->start, e->src, gate);
  else
    k_reorder<3><<<nb, 256, 0, st>>>(bound, e->grid, k, e->slab, o, A, B, e->key, e->start, e->src, gate);
  e->launches += 5;
  CK(cudaGetLastError());
  {  // the sorted particles go back to the frame the host keeps launching on
    k_copyback<<<nb, 256, 0, st>>>(e->n, e->slab, e->grid.ncells, o, A, B, e->rb, e->start, gate,
                                   e->ctl + 2);
    e->launches++;
    CK(cudaGetLastError());
  }
  (void)k;
  e->cells_valid = true;
  return SPHB200_OK;  // the callers run nw_fn themselves, on every integrating step
}

int wall_normals(sphb200_engine* e, cudaStream_t st) {
  if (e->wl_n > 0 && e->has_nw) {
    const int wb = e->slab_on ? e->sgeom.own_cap : e->n;
    Frame& F = e->fr[e->cur];
    if (e->dim == 2)
      k_wall_normals<2><<<(wb + 255) / 256, 256, 0, st>>>(wb, e->grid, e->slab, e->consts, F, e->wl_pts,
                                                        e->wl_n, e->wl_off[0], e->wl_off[1],
                                                        e->wl_off[2], e->wl_c2);
    else
      k_wall_normals<3><<<(wb + 255) / 256, 256, 0, st>>>(wb, e->grid, e->slab, e->consts, F, e->wl_pts,
                                                        e->wl_n, e->wl_off[0], e->wl_off[1],
                                                        e->wl_off[2], e->wl_c2);
    e->launches++;
    CK(cudaGetLastError());
  }
  return SPHB200_OK;
}

// First half of one step of a resident engine: kick + drift + wrap in place and, when the
// device decides so (k_drift) or the host knows the lists are stale, the cell sort.  The
// search itself is the first launch of forward_stage(0), gated on the same flag word.
// kick + drift + wrap in place and this engine's own re-sort decision -> the step's flag word
int drift_step(sphb200_engine* e, const Kick& k, cudaStream_t st) {
  int* fc = e->ctl + (e->step_no & 1ull);
  int* fn = e->ctl + ((e->step_no + 1ull) & 1ull);
  const int force = (e->force_rebuild || !e->cells_valid) ? 1 : 0;
  Frame& F = e->fr[e->cur];
  const float lim2 = e->path_limit < 0.f ? -1.0f : e->path_limit * e->path_limit;
  // relative criterion: ctl[5], ctl[6] are the alternating `maybe` words of k_drift
  int* mc = e->rel_drift ? e->ctl + 5 + (e->step_no & 1ull) : nullptr;
  int* mn = e->rel_drift ? e->ctl + 5 + ((e->step_no + 1ull) & 1ull) : nullptr;
  const float guard2 = e->rel_guard * e->rel_guard;
  if (k.on || e->positions_replaced) {  // (a forward-only call on a re-uploaded state tests it too)
    const int nb = stream_blocks(e, e->slab_on ? e->sgeom.own_cap : e->n);
    if (e->dim == 2)
      k_drift<2><<<nb, 256, 0, st>>>(e->n, e->grid, k, e->slab, F, e->rb, fc, fn, force, lim2, e->err, mc, mn, guard2);
    else
      k_drift<3><<<nb, 256, 0, st>>>(e->n, e->grid, k, e->slab, F, e->rb, fc, fn, force, lim2, e->err, mc, mn, guard2);
    if (e->rel_drift && !force) {
      const int blocks = e->dblocks.nb[0] * e->dblocks.nb[1] * e->dblocks.nb[2];
      const int nbk = (blocks + 127) / 128;
      const float rl2 = e->rel_limit * e->rel_limit;
      // boxes -> buffer 0; joins along x (0 -> 1), y (1 -> 2, or the test in 2D), z (the test)
      if (e->dim == 2) {
        k_drift_box<2><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, e->start, F.pt, e->rb, fc, mc);
        k_drift_join<0, false><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, 0, 1, rl2, fc, mc);
        k_drift_join<1, true><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, 1, 2, rl2, fc, mc);
      } else {
        k_drift_box<3><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, e->start, F.pt, e->rb, fc, mc);
        k_drift_join<0, false><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, 0, 1, rl2, fc, mc);
        k_drift_join<1, false><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, 1, 2, rl2, fc, mc);
        k_drift_join<2, true><<<nbk, 128, 0, st>>>(e->grid, e->dblocks, 2, 0, rl2, fc, mc);
      }
      e->launches += 1 + e->dim;
    }
    e->maybe_drifted = true;
    e->positions_replaced = false;
  } else {
    k_gate<<<1, 1, 0, st>>>(fc, fn, force);
  }
  e->launches++;
  CK(cudaGetLastError());
  e->gate_cur = fc;
  e->force_rebuild = false;
  e->step_no++;
  return SPHB200_OK;
}

int begin_step(sphb200_engine* e, const Kick& k, cudaStream_t st) {
  int rc = drift_step(e, k, st);
  if (rc) return rc;
  const Kick off{0.f, 0.f, 0};
  rc = hash_cells(e, off, st, e->gate_cur);
  if (!rc) rc = sort_cells(e, off, st, e->gate_cur);
  if (rc) return rc;
  if (k.on) rc = wall_normals(e, st);
  return rc;
}

// Sort the current positions now (ungated) -- for callers that search on their own right
// after (the neighbour-list materialiser).  The engine's lists no longer describe the new
// order: the next step searches again.
int sort_now(sphb200_engine* e, cudaStream_t st) {
  const Kick off{0.f, 0.f, 0};
  int rc = hash_cells(e, off, st);
  if (!rc) rc = sort_cells(e, off, st, nullptr, true);  // the next kick needs dudt / dvdt
  e->force_rebuild = true;
  e->maybe_drifted = false;
  return rc;
}

// One stage of WCSPH.forward (solver.py:705-949): 0 density + EoS, 1 Shepard renormalisation,
// 2 generalized wall BC, 3 force (+ case bc_fn).  *wrote = HX_* mask of the per-particle
// arrays the stage changed that later stages read from NEIGHBOURS (what a slab halo must
// refresh); 0 when the stage is not part of this solver variant.
// part (slab overlap, stages 0 and 3 only): 1 = launch the interior tile layers and nothing else
// (no swap, no bookkeeping: the boundary call finishes the stage), 2 = the boundary tile layers.
int forward_stage(sphb200_engine* e, int stage, uint32_t flags, bool v_is_u, cudaStream_t st,
                  int* wrote, int part = 0) {
  const sphb200_config& c = e->cfg;
  *wrote = 0;
  struct PartGuard {  // launch_sweep reads e->part; always back to "all tiles" on return
    sphb200_engine* e;
    ~PartGuard() { e->part = 0; }
  } part_guard{e};
  e->part = part;
  const bool bc_trick = c.flags & SPHB200_F_BC_TRICK, evol = c.flags & SPHB200_F_RHO_EVOL,
             renorm = c.flags & SPHB200_F_RHO_RENORM, free_slip = c.flags & SPHB200_F_FREE_SLIP,
             heat = c.flags & SPHB200_F_HEAT;
  const bool rie = c.solver == SPHB200_SOLVER_RIE;
  const bool wall_sweep = bc_trick && !rie;
  int rc = SPHB200_OK;
  // quads staged by the force sweep (decides its staging capacity)
  int force_nq = 3;
  const int fq_v = (!rie && !v_is_u) ? force_nq++ : -1;
  const int fq_h = heat ? force_nq++ : -1;
  const int fq_nw = rie ? force_nq++ : -1;
  const int fq_ut = (rie && e->has_ut) ? force_nq++ : -1;
  // the two headline SPH variants stage a compact record (phys.cuh, PhysForce)
  const bool av_on = c.artificial_alpha != 0.0;
  const bool delta_on = c.solver == SPHB200_SOLVER_DELTA;
  const int force_feat =
      (heat || av_on || delta_on) ? FORCE_GENERIC : (v_is_u ? FORCE_PLAIN : FORCE_TVF);
  const int force_sb = (rie || force_feat == FORCE_GENERIC) ? 16 * force_nq
                                                           : (force_feat == FORCE_TVF ? 52 : 40);
  const SweepPlan planF =
      plan_sweep_bytes(e, force_sb, e->pl_lmax > 0 ? (e->lcap < 24 ? e->lcap : 24) : e->lcap);
  const bool dens_extras = e->has_ut || (rie && bc_trick && heat);
  const SweepPlan planD = !evol ? (dens_extras ? e->planW : e->planA) : (!rie ? e->planR : e->planW);
  // neighbour lists (sweep.cuh): skin list from the last search, exact list written by the
  // density sweep of this step and consumed by every later sweep
  NList nl{e->pl_list, e->pl_cnt, e->sl_list, e->sl_cnt, e->pl_ok, e->pl_lmax, 0, nullptr};
  const SweepPlan planG = plan_sweep(e, e->dim == 3 ? 4 : 2, 24);  // PhysDelta<1>
  {
    int mc = planD.cap < planF.cap ? planD.cap : planF.cap;
    if (delta_on && evol) {
      if (planG.cap < mc) mc = planG.cap;
      if (e->planC.cap < mc) mc = e->planC.cap;
    }
    if (evol && renorm && e->planR.cap < mc) mc = e->planR.cap;
    if (wall_sweep && e->planC.cap < mc) mc = e->planC.cap;
    nl.min_cap = mc;
  }
  // duo sweeps (sweep2.cuh) for the tiles whose stencil fits; the others are left to sweep.cuh
  // kernels that search on their own (nlb: no lists, skip the ok tiles), launched on a small
  // persistent grid and only when the search counted such tiles (ctl[4])
  const int ntiles = e->grid.nt[0] * e->grid.nt[1] * e->grid.nt[2];
  SweepPlan planDF = planF;
  DuoList dl{e->duo_desc, e->duo_desc_stride, e->sl_list, e->pl_list, e->sl_cnt, e->pl_cnt,
             e->pl_ok,    e->duo_lmax,        0,          e->duo_rows};
  NList nlb{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr, e->pl_ok};
  int* const nbad = e->ctl + 4;
  // bytes of the duo force record in shared memory (phys.cuh, PhysForce: FORCE_PLAIN keeps eta and
  // (m/rho)^2 in a third quad there)
  const int duo_feat = (force_feat == FORCE_TVF && (c.hints & SPHB200_HINT_UNIFORM_ETA)) ? FORCE_TVF_U
                                                                                         : force_feat;
  const int duo_force_sb = duo_feat == FORCE_TVF ? 52 : 48;
  if (e->duo) {
    planDF = plan_duo(e, duo_force_sb, 0);
    int mc = e->planDB.cap < e->planDA.cap ? e->planDB.cap : e->planDA.cap;
    if (planDF.cap < mc) mc = planDF.cap;
    dl.min_cap = mc;
  }
  // ---- the search (on the steps that re-sorted), then density ---------------
  if (stage == 0 && e->duo) {
    Extra exb = make_extra();
    exb.nq = 1;
    Frame& F = e->fr[e->cur];
    const int* gate = e->inplace ? e->gate_cur : nullptr;
    if (e->dim == 2)
      rc = launch_duo(e, k_duo<2, PhysNone, DUO_BUILD>, e->planDB, F, exb, st, dl, gate, gate != nullptr);
    else
      rc = launch_duo(e, k_duo<3, PhysNone, DUO_BUILD>, e->planDB, F, exb, st, dl, gate, gate != nullptr);
    if (rc) return rc;
    k_duo_count_bad<<<1, 1024, 0, st>>>(e->pl_ok, ntiles, nbad, gate);
    e->launches++;
    CK(cudaGetLastError());
    Extra ex = make_extra();
    ex.finalT = false;
    ex.st_out = e->fr[1 - e->cur].st;
    ex.nq = 1;
#define CALL(D, K) rc = launch_duo(e, k_duo<D, PhysDensity<D, K, DENS_SUM>, DUO_FILTER>, e->planDA, F, ex, st, dl)
    DISPATCH_DK(e, CALL);
#undef CALL
    if (rc) return rc;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDensity<D, K, DENS_SUM>, LIST_FILTER>, planD, F, ex, st, nlb, nbad, true, false, -1)
    DISPATCH_DK(e, CALL);
#undef CALL
    if (rc) return rc;
    if (part == 1) return SPHB200_OK;
    swap_st(e);
    *wrote = HX_ST;
    if (e->profile) cudaEventRecord(e->ev[3], st);
  } else if (stage == 0) {
    const bool first_sub = !(e->slab_on && delta_on && evol && e->delta_sub > 0);
    if (e->pl_lmax > 0 && first_sub) {
      Extra exb = make_extra();
      exb.nq = 1;
      Frame& FB = e->fr[e->cur];
      const int* gate = e->inplace ? e->gate_cur : nullptr;
      if (e->dim == 2)
        rc = launch_sweep(e, k_sweep<2, PhysNone, LIST_BUILD>, e->planB, FB, exb, st, nl, gate, gate != nullptr);
      else
        rc = launch_sweep(e, k_sweep<3, PhysNone, LIST_BUILD>, e->planB, FB, exb, st, nl, gate, gate != nullptr);
      if (rc) return rc;
    }
    Extra ex = make_extra();
    ex.utilde = e->has_ut;
    ex.wallT = rie && bc_trick && heat;
    ex.finalT = heat && !wall_sweep;
    ex.heat = heat;
    Frame& F = e->fr[e->cur];
    ex.st_out = e->fr[1 - e->cur].st;
    if (!evol) {
      if (dens_extras) {
        ex.nq = 3;  // 3 quads fit in the 4-quad plan
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDensity<D, K, DENS_SUM_X>, LIST_FILTER>, planD, F, ex, st, nl)
        DISPATCH_DK(e, CALL);
#undef CALL
      } else {
        ex.nq = 1;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDensity<D, K, DENS_SUM>, LIST_FILTER>, planD, F, ex, st, nl)
        DISPATCH_DK(e, CALL);
#undef CALL
      }
    } else if (delta_on) {
      // rho_evol_fn_delta (solver.py:36-103): L matrices (list builder), gradient terms, update.
      // A slab engine runs ONE of the three sweeps per call: the neighbours' L rows and H terms
      // of the halo particles have to arrive in between (wrote = what to send next).
      const int sub0 = e->slab_on ? e->delta_sub : 0, sub1 = e->slab_on ? e->delta_sub + 1 : 3;
      for (int sub = sub0; sub < sub1 && !rc; ++sub) {
        if (sub == 0) {
          ex.nq = 1;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDelta<D, K, 0>, LIST_FILTER>, e->planA, F, ex, st, nl)
          DISPATCH_DK(e, CALL);
#undef CALL
        } else if (sub == 1) {
          ex.nq = e->dim == 3 ? 4 : 2;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDelta<D, K, 1>, LIST_CONSUME>, planG, F, ex, st, nl)
          DISPATCH_DK(e, CALL);
#undef CALL
        } else {
          ex.nq = 3;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDelta<D, K, 2>, LIST_CONSUME>, e->planC, F, ex, st, nl)
          DISPATCH_DK(e, CALL);
#undef CALL
        }
      }
      if (rc) return rc;
      if (e->slab_on) {
        if (e->delta_sub < 2) {
          *wrote = e->delta_sub == 0 ? (HX_DL0 | (e->dim == 3 ? (HX_DL1 | HX_DL2) : 0)) : HX_DG1;
          e->delta_sub++;
          e->stage_more = true;
          return SPHB200_OK;
        }
        e->delta_sub = 0;
      }
    } else if (!rie) {
      ex.nq = 2;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDensity<D, K, DENS_EVOL_SPH>, LIST_FILTER>, planD, F, ex, st, nl)
      DISPATCH_DK(e, CALL);
#undef CALL
    } else {
      ex.nq = 4;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysDensity<D, K, DENS_EVOL_RIE>, LIST_FILTER>, planD, F, ex, st, nl)
      DISPATCH_DK(e, CALL);
#undef CALL
    }
    if (rc) return rc;
    if (part == 1) return SPHB200_OK;
    swap_st(e);
    *wrote = HX_ST | (e->has_ut ? HX_UT : 0);
    if (e->profile) cudaEventRecord(e->ev[3], st);
  }
  if (stage == 1 && evol && renorm) {
    Extra ex = make_extra();
    ex.nq = 2;
    Frame& F = e->fr[e->cur];
    ex.st_out = e->fr[1 - e->cur].st;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysRenorm<D, K>, LIST_CONSUME>, e->planR, F, ex, st, nl)
    DISPATCH_DK(e, CALL);
#undef CALL
    if (rc) return rc;
    swap_st(e);
    *wrote = HX_ST;
  }
  // ---- generalized wall boundary condition ---------------------------------
  if (stage == 2 && wall_sweep) {
    Extra ex = make_extra();
    ex.nq = 4;
    ex.heat = heat;
    ex.free_slip = free_slip;
    Frame& F = e->fr[e->cur];
    ex.st_out = e->fr[1 - e->cur].st;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysWall<D, K>, LIST_CONSUME>, e->planC, F, ex, st, nl)
    DISPATCH_DK(e, CALL);
#undef CALL
    if (rc) return rc;
    swap_st(e);
    *wrote = HX_UM | HX_VV | HX_ST;
  }
  // ---- force -----------------------------------------------------------------
  if (stage == 3) {
    if (e->profile && part != 1) cudaEventRecord(e->ev[4], st);
    Extra ex = make_extra();
    ex.q_v = fq_v; ex.q_h = fq_h; ex.q_nw = fq_nw; ex.q_ut = fq_ut;
    ex.nq = force_nq;
    ex.sb = force_sb;
    ex.heat = heat;
    ex.av = av_on;
    ex.delta = delta_on;
    ex.bc_on = (flags & SPHB200_STEP_BC) && bc_table_on(c);
    ex.free_slip = free_slip;
    ex.bc_trick = bc_trick;
    const SweepPlan& sp = planF;
    Frame& F = e->fr[e->cur];
    if (e->duo) {
      // the compact record of every slot the sweep may stage (own slots; on a slab engine the
      // halo slots too, once their (rho, p) has arrived: the boundary part of an overlapped step)
      Extra exd = ex;
      exd.sb = duo_force_sb;
      exd.rec0 = e->rec0; exd.rec1 = e->rec1; exd.rec2 = e->rec2; exd.rec_e = e->rec_e;
      exd.eta_ref = F.vv + e->slab.base;
      exd.err_word = e->err;
      {
        const int mode = !e->slab_on ? 0 : (part == 1 ? 1 : (part == 2 ? 2 : 0));
        const int nb = stream_blocks(e, e->n);
#define REC(D, FEAT) k_force_rec<PhysForce<D, SPHB200_KERNEL_QSK, SPHB200_SOLVER_SPH, FEAT>><<<nb, 256, 0, st>>>(e->n, e->slab, mode, F, exd)
        if (e->dim == 2) {
          if (duo_feat == FORCE_PLAIN) REC(2, FORCE_PLAIN);
          else if (duo_feat == FORCE_TVF) REC(2, FORCE_TVF);
          else REC(2, FORCE_TVF_U);
        } else {
          if (duo_feat == FORCE_PLAIN) REC(3, FORCE_PLAIN);
          else if (duo_feat == FORCE_TVF) REC(3, FORCE_TVF);
          else REC(3, FORCE_TVF_U);
        }
#undef REC
        e->launches++;
        CK(cudaGetLastError());
      }
      if (force_feat == FORCE_PLAIN) {
#define CALL(D, K) rc = launch_duo(e, k_duo<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_PLAIN>, DUO_CONSUME, 2>, planDF, F, exd, st, dl, nullptr, false, 2)
        DISPATCH_DK(e, CALL);
#undef CALL
        if (rc) return rc;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_PLAIN>, LIST_CONSUME>, sp, F, ex, st, nlb, nbad, true, false, -1)
        DISPATCH_DK(e, CALL);
#undef CALL
      } else if (duo_feat == FORCE_TVF_U) {
#define CALL(D, K) rc = launch_duo(e, k_duo<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_TVF_U>, DUO_CONSUME, 2>, planDF, F, exd, st, dl, nullptr, false, 2)
        DISPATCH_DK(e, CALL);
#undef CALL
        if (rc) return rc;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_TVF>, LIST_CONSUME>, sp, F, ex, st, nlb, nbad, true, false, -1)
        DISPATCH_DK(e, CALL);
#undef CALL
      } else {
#define CALL(D, K) rc = launch_duo(e, k_duo<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_TVF>, DUO_CONSUME, 2>, planDF, F, exd, st, dl, nullptr, false, 2)
        DISPATCH_DK(e, CALL);
#undef CALL
        if (rc) return rc;
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_TVF>, LIST_CONSUME>, sp, F, ex, st, nlb, nbad, true, false, -1)
        DISPATCH_DK(e, CALL);
#undef CALL
      }
    } else if (!rie) {
      const int feat = force_feat;
      if (feat == FORCE_PLAIN) {
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_PLAIN>, LIST_CONSUME>, sp, F, ex, st, nl)
        DISPATCH_DK(e, CALL);
#undef CALL
      } else if (feat == FORCE_TVF) {
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_TVF>, LIST_CONSUME>, sp, F, ex, st, nl)
        DISPATCH_DK(e, CALL);
#undef CALL
      } else {
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_SPH, FORCE_GENERIC>, LIST_CONSUME>, sp, F, ex, st, nl)
        DISPATCH_DK(e, CALL);
#undef CALL
      }
    } else {
#define CALL(D, K) rc = launch_sweep(e, k_sweep<D, PhysForce<D, K, SPHB200_SOLVER_RIE, FORCE_GENERIC>, LIST_CONSUME>, sp, F, ex, st, nl)
      DISPATCH_DK(e, CALL);
#undef CALL
    }
    if (rc) return rc;
    if (part == 1) return SPHB200_OK;
    if (ex.bc_on) {
      const int bound = e->slab_on ? e->sgeom.own_cap : e->n;
      const int nb = (bound + 255) / 256;
      if (e->dim == 2) k_bc<2><<<nb, 256, 0, st>>>(bound, e->slab, e->consts, F);
      else k_bc<3><<<nb, 256, 0, st>>>(bound, e->slab, e->consts, F);
      e->launches++;
      CK(cudaGetLastError());
    }
  }
  return SPHB200_OK;
}

int run_forward(sphb200_engine* e, uint32_t flags, bool v_is_u, cudaStream_t st) {
  for (int stage = 0; stage < 4; ++stage) {
    int wrote;
    int rc = forward_stage(e, stage, flags, v_is_u, st, &wrote);
    if (rc) return rc;
  }
  return SPHB200_OK;
}

__global__ void k_fill(int* a, long long n, int v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// exclusive scan of counts[0..n) (original order) using the cell-scan kernels;
// total edge count -> *count64, overflow -> err
__global__ void k_nl_total(int n, const int* offs_last, const int* counts, long long capacity,
                           long long* count64, unsigned* err) {
  long long tot = (long long)(unsigned)offs_last[n - 1] + counts[n - 1];
  if (count64) *count64 = tot;
  if (tot > capacity) atomicOr(err, SPHB200_ERR_NEIGHBOR_OVERFLOW);
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_final_keep(int c, const int* __restrict__ count,
                                                              const int* __restrict__ bsum,
                                                              int* __restrict__ start) {
  __shared__ int sh[32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < c) ? count[base + i] : 0;
    s += v[i];
  }
  int tot;
  int inc = block_incl_scan(s, sh, &tot);
  int run = bsum[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < c) start[base + i] = run;
    run += v[i];
  }
}

template <int DIM>
__global__ void __launch_bounds__(256) k_stats(int n, Slab sl, Frame f, double* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = sl.base + t;
  double ek = 0.0;
  float um = 0.f;
  if (t < (sl.dn ? sl.dn[DN_OWN] : n)) {
    float4 u = f.um[p];
    float s = u.x * u.x + u.y * u.y + (DIM == 3 ? u.z * u.z : 0.f);
    ek = 0.5 * (double)u.w * (double)s;
    um = sqrtf(s);
  }
  for (int o = 16; o > 0; o >>= 1) {
    ek += __shfl_xor_sync(FULL_MASK, ek, o);
    um = fmaxf(um, __shfl_xor_sync(FULL_MASK, um, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], ek);
    atomicMax(reinterpret_cast<unsigned long long*>(&out[1]),
              (unsigned long long)__double_as_longlong((double)um));
  }
}

// get_stats (jax_sph/utils.py:128-166) on the resident frame, any particle order:
//   out[0]            sum over FLUID particles of |v|^2           (get_ekin, utils.py:128-133)
//   out[1 + 3k + 0/1] min / max as order-preserving uint keys, out[1 + 3k + 2] sum, for
//                     k = |u|, |v|, rho, p, T                      (get_array_stats, :136-153)
//   out[16]           particles counted
__device__ __forceinline__ unsigned f2key(float x) {
  const unsigned b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_get_stats(int n, Slab sl, Frame f, double* out) {
  const int bound = sl.dn ? sl.dn[DN_OWN] : n;
  double sum[6] = {0, 0, 0, 0, 0, 0};  // ekin, u, v, rho, p, T
  float mn[5], mx[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    mn[k] = INFINITY;
    mx[k] = -INFINITY;
  }
  int cnt = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < bound; t += gridDim.x * blockDim.x) {
    const int p = sl.base + t;
    const float4 pt = f.pt[p], u = f.um[p], v = f.vv[p], st = f.st[p];
    const float uu = u.x * u.x + u.y * u.y + (DIM == 3 ? u.z * u.z : 0.f);
    const float vv = v.x * v.x + v.y * v.y + (DIM == 3 ? v.z * v.z : 0.f);
    const float val[5] = {sqrtf(uu), sqrtf(vv), st.x, st.y, st.z};
    if (__float_as_int(pt.w) == SPHB200_TAG_FLUID) sum[0] += (double)vv;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      sum[k + 1] += (double)val[k];
      mn[k] = fminf(mn[k], val[k]);
      mx[k] = fmaxf(mx[k], val[k]);
    }
    ++cnt;
  }
  __shared__ double s_sum[8][6];
  __shared__ float s_mn[8][5], s_mx[8][5];
  __shared__ int s_cnt[8];
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 6; ++k) sum[k] += __shfl_xor_sync(FULL_MASK, sum[k], o);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(FULL_MASK, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(FULL_MASK, mx[k], o));
    }
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    for (int k = 0; k < 6; ++k) s_sum[warp][k] = sum[k];
    for (int k = 0; k < 5; ++k) {
      s_mn[warp][k] = mn[k];
      s_mx[warp][k] = mx[k];
    }
    s_cnt[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      for (int k = 0; k < 6; ++k) s_sum[0][k] += s_sum[w][k];
      for (int k = 0; k < 5; ++k) {
        s_mn[0][k] = fminf(s_mn[0][k], s_mn[w][k]);
        s_mx[0][k] = fmaxf(s_mx[0][k], s_mx[w][k]);
      }
      s_cnt[0] += s_cnt[w];
    }
    if (s_cnt[0] > 0) {
      atomicAdd(&out[0], s_sum[0][0]);
      unsigned long long* key = reinterpret_cast<unsigned long long*>(out);
      for (int k = 0; k < 5; ++k) {
        atomicMin(&key[1 + 3 * k], (unsigned long long)f2key(s_mn[0][k]));
        atomicMax(&key[2 + 3 * k], (unsigned long long)f2key(s_mx[0][k]));
        atomicAdd(&out[3 + 3 * k], s_sum[0][k + 1]);
      }
      atomicAdd(&out[16], (double)s_cnt[0]);
    }
  }
}

bool duo_variant(const sphb200_config& c);

// Grid + duo decision of an engine for n particle slots (every entry point that sizes or creates
// an engine plans through here, so that they agree on the tiling).
void plan_all(const sphb200_config& c, int64_t n, int rank, int nranks, Grid& g, DuoPlan& dp) {
  const int tpb = c.threads > 0 ? (c.threads + 31) / 32 * 32 : DEFAULT_TPB;
  const double skin = plan_skin(c);
  dp.on = duo_variant(c);
  dp.tpb = dp.lmax = dp.rows = dp.desc_stride = 0;
  if (dp.on) {
    // tuned on B200 (profiles/r02_duo_*): one block per SM holds the force record of the stencil
    // (48-byte force records leave room for a ninth cell along x: 384 threads hold its duos)
    dp.tpb = c.dim == 3 ? (((c.hints & SPHB200_HINT_UNIFORM_ETA) && c.tvf != 0.0) ? 384 : 352) : 256;
    if (const char* env = getenv("SPHB200_DUO_TPB")) {
      const int v = atoi(env);
      if (v >= 32 && v <= 384) dp.tpb = (v + 31) / 32 * 32;  // (two lanes per duo in the force sweep)
    }
  }
  plan_grid(c, g, tpb > 512 ? 512 : tpb, rank, nranks, skin, dp.tpb);
  if (!dp.on) return;
  // the union window of a duo (first particle's window start to the second one's window end)
  // must not wrap onto itself, and the tile tables must fit
  bool fits = !g.exact_all && g.n[0] >= g.T[0] + 2 * g.S[0] && g.T[1] * g.T[2] <= DUO_RUNS;
  for (int a = 1; a < c.dim; ++a) fits = fits && g.n[a] >= 2 * g.S[a] + 1;
  // Duo list rows live in the buffers of the per-particle lists: about half as many rows (one
  // per duo; sweep2.cuh numbers them ceil((slot + run ordinal) / 2)), each the union of two
  // neighbour sets (1.23 x one set for slot neighbours one spacing apart, 1.75 x at worst).
  const int pl_lmax = plan_lmax(c, skin);
  const long long runs = (long long)g.n[1] * g.n[2] * g.nt[0];
  const long long rows = (n + runs) / 2 + 2;
  const double q = kernel_cutoff(c) * (1.0 + skin) / c.dx;
  const double expect = c.dim == 2 ? M_PI * q * q : 4.0 / 3.0 * M_PI * q * q * q;
  long long lmax = ((long long)(1.3 * 1.4 * expect) + 8 + 7) / 8 * 8;
  const long long room = pl_lmax > 0 ? (long long)n * pl_lmax / rows / 8 * 8 : 0;
  if (lmax > room) lmax = room;
  int srows = 1;
  for (int a = 1; a < 3; ++a) srows *= (g.n[a] >= 2 * g.S[a] + 1) ? g.T[a] + 2 * g.S[a] : g.n[a];
  const int stride = duo_desc_ints(srows);
  fits = fits && rows <= n && lmax >= 16 && stride <= DUO_SOFF;
  dp.rows = (int)rows;
  dp.lmax = (int)lmax;
  dp.desc_stride = stride;
  if (!fits) {
    dp.on = false;
    dp.tpb = dp.lmax = dp.rows = dp.desc_stride = 0;
    plan_grid(c, g, tpb > 512 ? 512 : tpb, rank, nranks, skin, 0);
  }
}

struct SlabSpec {
  int rank, nranks;
  int own_cap, halo_cap, mig_cap;
};

// capacity (particle slots) of a slab engine
int64_t slab_slots(const SlabSpec& sp) { return (int64_t)sp.halo_cap + sp.own_cap + sp.halo_cap; }

int init_engine(sphb200_engine* e, const sphb200_config* cfg, int64_t n, void* ws, size_t ws_bytes,
                bool own, const SlabSpec* sp = nullptr) {
  e->cfg = *cfg;
  e->n = (int)n;
  e->dim = cfg->dim;
  e->slab_on = sp != nullptr;
  e->slab_rank = sp ? sp->rank : 0;
  e->slab_nranks = sp ? sp->nranks : 1;
  e->slab_axis = cfg->dim - 1;
  e->slab_stage = 0;
  e->slab_pending_mask = 0;
  e->part = 0;
  e->pre_stage = -1;
  e->overlap = false;
  e->io_on = false;
  e->delta_sub = 0;
  e->stage_more = false;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  int maxs = 0;
  CK(cudaDeviceGetAttribute(&maxs, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  e->max_smem = maxs - 1024;
  e->tpb = cfg->threads > 0 ? (cfg->threads + 31) / 32 * 32 : DEFAULT_TPB;
  if (e->tpb > 512) e->tpb = 512;
  e->lcap = cfg->list_cap > 0 ? cfg->list_cap : 48;  // tuned on B200 (profiles/, tune logs)
  if (e->lcap < SWEEP_CHUNK) e->lcap = SWEEP_CHUNK;
  e->inplace = true;
  e->skin_frac = plan_skin(*cfg);
  DuoPlan dp;
  plan_all(*cfg, n, e->slab_rank, e->slab_nranks, e->grid, dp);
  e->duo = dp.on;
  e->duo_tpb = dp.tpb;
  e->duo_lmax = dp.lmax;
  e->duo_rows = dp.rows;
  e->duo_desc_stride = dp.desc_stride;
  // half the skin, minus a margin for the rounding of positions and of the accumulated path
  e->path_limit = e->skin_frac > 0.0 ? (float)(0.5 * e->skin_frac * kernel_cutoff(*cfg) * (1.0 - 1e-3)) : -1.0f;
  e->step_no = 0;
  e->agree.n = 0;
  e->ring_seq = 0;
  e->gate_cur = nullptr;
  e->force_rebuild = true;
  e->maybe_drifted = false;
  e->positions_replaced = false;
  CK(cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, dev));
  plan_consts(*cfg, e->consts);
  feature_flags(*cfg, e->has_kc, e->has_nw, e->has_ut, e->has_ge);
  Layout L;
  plan_layout(*cfg, n, e->grid, e->skin_frac, L, dp);
  if (ws_bytes < L.total) return SPHB200_ENOMEM;
  e->arena = (char*)ws;
  e->arena_bytes = L.total;
  e->own_arena = own;
  for (int f = 0; f < 2; ++f) {
    Frame& F = e->fr[f];
    memset(&F, 0, sizeof(F));
    F.pt = (float4*)(e->arena + L.frame[f][0]);
    F.um = (float4*)(e->arena + L.frame[f][1]);
    F.vv = (float4*)(e->arena + L.frame[f][2]);
    F.st = (float4*)(e->arena + L.frame[f][3]);
    F.du = (float4*)(e->arena + L.frame[f][4]);
    F.dv = (float4*)(e->arena + L.frame[f][5]);
    F.id = (int*)(e->arena + L.frame[f][6]);
    F.kc = e->has_kc ? (float2*)(e->arena + L.frame[f][7]) : nullptr;
    F.nw = e->has_nw ? (float4*)(e->arena + L.frame[f][8]) : nullptr;
    F.ge = e->has_ge ? (float4*)(e->arena + L.frame[f][9]) : nullptr;
    F.ut = e->has_ut ? (float4*)(e->arena + L.ut) : nullptr;
    F.dl0 = F.dl1 = F.dl2 = F.dg0 = F.dg1 = nullptr;
    if (L.dl != (size_t)-1) {  // planes of n quads
      F.dl0 = (float4*)(e->arena + L.dl);
      F.dl1 = F.dl0 + n;
      F.dl2 = F.dl1 + n;
      F.dg0 = (float4*)(e->arena + L.dg);
      F.dg1 = F.dg0 + n;
    }
  }
  e->key = (int*)(e->arena + L.key);
  e->rnk = (int*)(e->arena + L.rnk);
  e->src = (int*)(e->arena + L.src);
  e->nl_counts = (int*)(e->arena + L.nl_counts);
  e->count = (int*)(e->arena + L.count);
  e->start = (int*)(e->arena + L.start);
  e->bsum = (int*)(e->arena + L.bsum);
  e->maxocc = (int*)(e->arena + L.maxocc);
  e->wallcount = (int*)(e->arena + L.wallcount);
  e->err = (unsigned*)(e->arena + L.err);
  e->stats = (double*)(e->arena + L.stats);
  e->pl_lmax = L.pl_lmax;
  e->pl_list = L.pl_lmax ? (unsigned short*)(e->arena + L.pl_list) : nullptr;
  e->pl_cnt = L.pl_lmax ? (int*)(e->arena + L.pl_cnt) : nullptr;
  e->pl_ok = L.pl_lmax ? (unsigned char*)(e->arena + L.pl_ok) : nullptr;
  e->sl_list = L.pl_lmax ? (unsigned short*)(e->arena + L.sl_list) : nullptr;
  e->sl_cnt = L.pl_lmax ? (int*)(e->arena + L.sl_cnt) : nullptr;
  e->rb = (float4*)(e->arena + L.path);
  e->ctl = (int*)(e->arena + L.ctl);
  e->rec0 = e->rec1 = e->rec2 = nullptr;
  e->rec_e = nullptr;
  e->duo_desc = nullptr;
  if (dp.on) {
    e->rec0 = (float4*)(e->arena + L.rec);
    e->rec1 = e->rec0 + n;
    e->rec2 = e->rec1 + n;
    e->rec_e = (float*)(e->rec2 + n);
    e->duo_desc = (int*)(e->arena + L.desc);
  }
  // Relative-drift criterion: single-GPU duo engines on a grid with room for the wider seam
  // margin of the interior shortcut (SPHB200_REL_DRIFT=0 keeps the absolute criterion alone).
  e->rel_drift = false;
  e->rel_guard = 0.f;
  e->grid.imargin = 0;
  if (dp.on && !sp && e->skin_frac > 0.0 && !e->grid.exact_all) {
    const char* env = getenv("SPHB200_REL_DRIFT");
    bool on = !(env && env[0] == '0');
    for (int a = 0; a < e->dim; ++a) on = on && e->grid.n[a] >= 6 * e->grid.S[a];
    if (on) {
      e->rel_drift = true;
      e->grid.imargin = 1;
      size_t blocks = 1;
      for (int a = 0; a < 3; ++a) {
        e->dblocks.nb[a] = a < e->dim ? (e->grid.n[a] + e->grid.S[a] - 1) / e->grid.S[a] : 1;
        blocks *= (size_t)e->dblocks.nb[a];
      }
      e->dblocks.bmin = (float4*)(e->arena + L.dbox);
      e->dblocks.bmax = e->dblocks.bmin + 3 * blocks;
      const double rc = kernel_cutoff(*cfg);
      // pairs that share a window of 2 S cells: relative drift below the skin; all others were two
      // (cutoff + skin) apart at the sort and stay a cutoff apart while every particle's own drift
      // is below (2 (cutoff + skin) - cutoff) / 2 (cells.cuh, k_drift_join); the same guard keeps
      // a particle that crossed the periodic seam a cutoff away from the interior tiles
      e->rel_limit = (float)(e->skin_frac * rc * (1.0 - 1e-3));
      e->rel_guard = (float)(0.5 * (2.0 * rc * (1.0 + e->skin_frac) - rc) * (1.0 - 1e-3));
    }
  }
  e->dn = (int*)(e->arena + L.dn);
  memset(&e->slab, 0, sizeof(e->slab));
  memset(&e->sgeom, 0, sizeof(e->sgeom));
  if (sp) {
    const Grid& g = e->grid;
    const int ax = e->slab_axis;
    int layer = 1;  // cells per layer along the slab axis
    for (int a = 0; a < ax; ++a) layer *= g.n[a];
    e->slab.base = sp->halo_cap;
    e->slab.dn = e->dn;
    e->slab.mig_cap = sp->mig_cap;
    e->sgeom.halo_cap = sp->halo_cap;
    e->sgeom.ncl = g.S[ax] * layer;
    e->sgeom.c_own_lo = g.own_lo[ax] * layer;
    e->sgeom.c_own_hi = g.own_hi[ax] * layer;
    e->sgeom.own_cap = sp->own_cap;
    e->sgeom.cap_total = (int)n;
    slab_range(g.ng[ax], sp->rank, sp->nranks, e->slab_z0, e->slab_z1);
  }
  e->cur = 0;
  e->adj_mem = nullptr;
  e->cells_valid = false;
  e->launches = 0;
  e->profile = false;
  e->hstage = nullptr;
  e->hstage_bytes = 0;
  memset(e->times, 0, sizeof(e->times));
  // sweeps that search (the list builder, the materialiser; every sweep when the lists are
  // off) want long per-thread columns, list consumers the minimum (their fall-back path only)
  const int lc = e->pl_lmax > 0 ? (e->lcap < 24 ? e->lcap : 24) : e->lcap;
  e->planA = plan_sweep(e, 1, lc);
  e->planR = plan_sweep(e, 2, lc);
  e->planW = plan_sweep(e, 4, lc);
  e->planC = plan_sweep(e, 4, lc);
  e->planN = plan_sweep(e, 2, e->lcap);
  e->planB = plan_sweep(e, 1, e->lcap);
  if (e->duo) {
    // the search and the density filter stage positions only: two blocks per SM
    e->planDB = plan_duo(e, 16, e->lcap < 40 ? e->lcap : 40, 2);
    e->planDA = plan_duo(e, 16, 0, 2);
  }
  e->needs_zero = true;  // control words are zeroed on the first upload's stream
  return SPHB200_OK;
}

// device-side staging of a reference-layout state for host pointers
struct HostStage {
  sphb200_state dev;
};

size_t state_floats(int dim) { return (size_t)(7 * dim + 10); }

__global__ void k_count_bad(int tiles, const unsigned char* __restrict__ ok, int* out) {
  int bad = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < tiles; t += gridDim.x * blockDim.x)
    bad += ok[t] == 0;
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(FULL_MASK, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, bad);
}

// Independent FMA chains: 16 per thread (8 packed pairs), no memory traffic.
template <bool PACKED>
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  if (PACKED) {
    unsigned long long v[8], pa, pb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(pa), "l"(pb));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v[i]));
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


}  // namespace

// ===========================================================================
extern "C" {

int sphb200_abi_version(void) { return SPHB200_ABI_VERSION; }

const char* sphb200_strerror(int code) {
  switch (code) {
    case SPHB200_OK: return "ok";
    case SPHB200_EINVAL: return "invalid argument or inconsistent configuration";
    case SPHB200_ENOMEM: return "out of device memory or workspace too small";
    case SPHB200_ECUDA: return "CUDA runtime error";
    case SPHB200_EDTYPE: return "unsupported dtype (float32 only)";
    case SPHB200_EUNSUP: return "unsupported solver/kernel variant";
    case SPHB200_ENODEV: return "no usable CUDA device";
    default: return "unknown error";
  }
}

void sphb200_config_default(sphb200_config* c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = sizeof(*c);
  c->dim = 3;
  c->solver = SPHB200_SOLVER_SPH;
  c->kernel = SPHB200_KERNEL_QSK;
  c->eos = SPHB200_EOS_TAIT;
  c->box[0] = c->box[1] = c->box[2] = 1.0;
  c->dx = 0.05; c->h = 0.05; c->dt = 0.0;
  c->c_ref = 10.0; c->eta_limiter = 3.0;
  c->p_ref = 100.0; c->rho_ref = 1.0; c->gamma = 1.0; c->u_ref = 1.0;
}

int sphb200_engine_bytes(const sphb200_config* cfg, int64_t n, size_t* bytes) {
  int rc = validate(cfg, n);
  if (rc) return rc;
  if (!bytes) return SPHB200_EINVAL;
  Grid g;
  const double skin = plan_skin(*cfg);
  DuoPlan dp;
  plan_all(*cfg, n, 0, 1, g, dp);
  Layout L;
  plan_layout(*cfg, n, g, skin, L, dp);
  *bytes = L.total;
  return SPHB200_OK;
}

int sphb200_workspace_bytes(const sphb200_config* cfg, int64_t n, size_t* bytes) {
  return sphb200_engine_bytes(cfg, n, bytes);
}

int sphb200_engine_create_in(const sphb200_config* cfg, int64_t n, void* ws, size_t ws_bytes,
                             sphb200_engine** out) {
  int rc = validate(cfg, n);
  if (rc) return rc;
  if (!ws || !out) return SPHB200_EINVAL;
  sphb200_engine* e = new (std::nothrow) sphb200_engine();
  if (!e) return SPHB200_ENOMEM;
  rc = init_engine(e, cfg, n, ws, ws_bytes, false);
  if (rc) {
    delete e;
    return rc;
  }
  *out = e;
  return SPHB200_OK;
}

int sphb200_engine_create(const sphb200_config* cfg, int64_t n, sphb200_engine** out) {
  int rc = validate(cfg, n);
  if (rc) return rc;
  if (!out) return SPHB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHB200_ENODEV;
  size_t bytes = 0;
  rc = sphb200_engine_bytes(cfg, n, &bytes);
  if (rc) return rc;
  void* ws = nullptr;
  if (cudaMalloc(&ws, bytes) != cudaSuccess) return SPHB200_ENOMEM;
  sphb200_engine* e = new (std::nothrow) sphb200_engine();
  if (!e) {
    cudaFree(ws);
    return SPHB200_ENOMEM;
  }
  rc = init_engine(e, cfg, n, ws, bytes, true);
  if (rc) {
    cudaFree(ws);
    delete e;
    return rc;
  }
  *out = e;
  return SPHB200_OK;
}

int sphb200_engine_destroy(sphb200_engine* e) {
  if (!e) return SPHB200_EINVAL;
  if (e->profile)
    for (int i = 0; i < 8; ++i) cudaEventDestroy(e->ev[i]);
  if (e->hstage) cudaFree(e->hstage);
  if (e->wl_pts) cudaFree(e->wl_pts);
  if (e->adj_mem) cudaFree(e->adj_mem);
  if (e->overlap) {
    cudaStreamDestroy(e->side);
    cudaEventDestroy(e->ev_main);
    cudaEventDestroy(e->ev_side);
  }
  if (e->io_on) {
    cudaStreamDestroy(e->io_side);
    cudaEventDestroy(e->ev_io_fork);
    cudaEventDestroy(e->ev_io_join);
  }
  if (e->own_arena) cudaFree(e->arena);
  delete e;
  return SPHB200_OK;
}

int sphb200_engine_set_wall_layer(sphb200_engine* e, const float* layer, int n_layer,
                                  const float* offset, double cutoff) {
  if (!e || n_layer < 0 || (n_layer > 0 && (!layer || !offset || !(cutoff > 0.0)))) return SPHB200_EINVAL;
  if (n_layer > 0 && !e->has_nw) return SPHB200_EINVAL;  // this solver variant reads no normals
  if (e->wl_pts) cudaFree(e->wl_pts);
  e->wl_pts = nullptr;
  e->wl_n = 0;
  if (n_layer == 0) return SPHB200_OK;
  float4* h = new (std::nothrow) float4[n_layer];
  if (!h) return SPHB200_ENOMEM;
  for (int i = 0; i < n_layer; ++i)
    h[i] = make_float4(layer[(size_t)i * e->dim], layer[(size_t)i * e->dim + 1],
                       e->dim == 3 ? layer[(size_t)i * e->dim + 2] : 0.f, 0.f);
  int rc = SPHB200_OK;
  if (cudaMalloc((void**)&e->wl_pts, (size_t)n_layer * sizeof(float4)) != cudaSuccess) rc = SPHB200_ENOMEM;
  else if (cudaMemcpy(e->wl_pts, h, (size_t)n_layer * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess)
    rc = SPHB200_ECUDA;
  delete[] h;
  if (rc) return rc;
  e->wl_n = n_layer;
  for (int a = 0; a < 3; ++a) e->wl_off[a] = a < e->dim ? offset[a] : 0.f;
  const float c = (float)cutoff;  // weak-typed scalar squared in float32 (jax_md/partition.py:820-821)
  e->wl_c2 = c * c;
  return SPHB200_OK;
}

static int ensure_hstage(sphb200_engine* e) {
  size_t need = up((size_t)e->n * 4) * (state_floats(e->dim) + 1);
  if (e->hstage && e->hstage_bytes >= need) return SPHB200_OK;
  if (e->hstage) cudaFree(e->hstage);
  e->hstage = nullptr;
  if (cudaMalloc((void**)&e->hstage, need) != cudaSuccess) return SPHB200_ENOMEM;
  e->hstage_bytes = need;
  return SPHB200_OK;
}

// rows: particles in *s (slab mode: this rank's own particles, ids = their global indices)
// in_place (stateless calls in engine order): row p of *s goes into slot p and ids[p] is the
// particle's label; the cells and lists stay valid as far as the positions allow (k_drift).
static int upload_impl(sphb200_engine* e, const sphb200_state* s, int rows, const int32_t* ids,
                       int on_host, void* stream, bool keep_order = false, bool in_place = false) {
  if (!e || !s || !s->r) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = rows, d = e->dim;
  sphb200_state dv = *s;
  const int32_t* dids = ids;
  if (on_host) {
    int rc = ensure_hstage(e);
    if (rc) return rc;
    char* base = e->hstage;
    size_t off = 0;
    auto stage = [&](const void* h, size_t words) -> void* {
      if (!h) return nullptr;
      void* dptr = base + off;
      off += up(words * 4) ;
      cudaMemcpyAsync(dptr, h, words * 4, cudaMemcpyHostToDevice, st);
      return dptr;
    };
    const size_t nv = (size_t)n * d, ns = (size_t)n;
    dv.r = (float*)stage(s->r, nv); dv.u = (float*)stage(s->u, nv); dv.v = (float*)stage(s->v, nv);
    dv.dudt = (float*)stage(s->dudt, nv); dv.dvdt = (float*)stage(s->dvdt, nv);
    dv.nw = e->has_nw ? (float*)stage(s->nw, nv) : nullptr;
    dv.rho = (float*)stage(s->rho, ns); dv.p = (float*)stage(s->p, ns);
    dv.drhodt = (float*)stage(s->drhodt, ns); dv.mass = (float*)stage(s->mass, ns);
    dv.eta = (float*)stage(s->eta, ns); dv.T = (float*)stage(s->T, ns);
    dv.dTdt = (float*)stage(s->dTdt, ns);
    dv.kappa = e->has_kc ? (float*)stage(s->kappa, ns) : nullptr;
    dv.Cp = e->has_kc ? (float*)stage(s->Cp, ns) : nullptr;
    dv.tag = (int32_t*)stage(s->tag, ns);
    dv.g_ext = e->has_ge ? (float*)stage(s->g_ext, nv) : nullptr;
    if (ids) dids = (const int32_t*)stage(ids, ns);
    CK(cudaGetLastError());
  }
  StatePtrs sp{dv.r, dv.u, dv.v, dv.dudt, dv.dvdt, dv.nw, dv.rho, dv.p, dv.drhodt, dv.mass,
               dv.eta, dv.T, dv.dTdt, dv.kappa, dv.Cp, dv.g_ext, dv.tag};
  if (e->has_ge && !dv.g_ext) return SPHB200_EINVAL;
  // keep_order: the same particles as the resident ones go into the slots they occupy; the
  // cells and lists survive if the new positions are close to the sorted ones (k_drift)
  keep_order = keep_order && e->cells_valid && !e->slab_on && !ids;
  in_place = in_place && e->cells_valid && !e->slab_on && !keep_order;
  if (keep_order || in_place) {
    e->positions_replaced = true;
  } else {
    e->cur = 0;
    e->cells_valid = false;
    e->force_rebuild = true;
    e->maybe_drifted = false;
  }
  if (e->needs_zero) {
    CK(cudaMemsetAsync(e->count, 0, (size_t)e->grid.ncells * 4, st));
    CK(cudaMemsetAsync(e->maxocc, 0, 4, st));
    CK(cudaMemsetAsync(e->err, 0, 4, st));
    CK(cudaMemsetAsync(e->ctl, 0, 8 * 4, st));
    e->needs_zero = false;
  }
  CK(cudaMemsetAsync(e->wallcount, 0, 4, st));
  if (e->slab_on) {
    int h[DN_WORDS] = {0};
    h[DN_OWN] = n;
    CK(cudaMemcpyAsync(e->dn, h, sizeof(h), cudaMemcpyHostToDevice, st));  // pageable: staged at once
  }
  const int nb = (n + 255) / 256;
  if (n > 0) {
    if (d == 2) k_pack<2><<<nb, 256, 0, st>>>(n, sp, e->fr[0], e->slab.base, dids, e->wallcount, keep_order ? 1 : 0);
    else k_pack<3><<<nb, 256, 0, st>>>(n, sp, e->fr[0], e->slab.base, dids, e->wallcount, keep_order ? 1 : 0);
    e->launches++;
  }
  CK(cudaGetLastError());
  return SPHB200_OK;
}

int sphb200_engine_upload(sphb200_engine* e, const sphb200_state* s, int on_host, void* stream) {
  if (!e || e->slab_on) return SPHB200_EINVAL;
  return upload_impl(e, s, e->n, nullptr, on_host, stream);
}

int sphb200_engine_refresh(sphb200_engine* e, const sphb200_state* s, int on_host, void* stream) {
  if (!e || e->slab_on) return SPHB200_EINVAL;
  return upload_impl(e, s, e->n, nullptr, on_host, stream, true);
}

// rows: capacity of the arrays in *out; ids != NULL (slab mode): local order + global indices
// `mask` (host downloads only): which entries of *out this call moves (DL_ALL: every one).  The
// staging offsets are those of the full *out for any mask, so the parts of a split download
// (advance_host) never share staging memory.
enum : unsigned {
  DL_R = 1u, DL_U = 2u, DL_V = 4u, DL_RHO = 8u, DL_P = 16u, DL_DUDT = 32u, DL_DVDT = 64u,
  DL_REST = 128u, DL_ALL = 255u
};
static int download_impl(sphb200_engine* e, sphb200_state* out, int rows, int32_t* ids,
                         int on_host, void* stream, unsigned mask = DL_ALL) {
  if (!e || !out) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = rows, d = e->dim;
  sphb200_state dv = *out;
  int32_t* dids = ids;
  struct Back { void* h; void* dptr; size_t bytes; } back[20];
  int nback = 0;
  if (on_host) {
    int rc = ensure_hstage(e);
    if (rc) return rc;
    char* base = e->hstage;
    size_t off = 0;
    auto stage = [&](void* h, size_t words) -> void* {
      if (!h) return nullptr;
      void* dptr = base + off;
      off += up(words * 4);
      back[nback++] = Back{h, dptr, words * 4};
      return dptr;
    };
    const size_t nv = (size_t)n * d, ns = (size_t)n;
    dv.r = (float*)stage(out->r, nv); dv.u = (float*)stage(out->u, nv); dv.v = (float*)stage(out->v, nv);
    dv.dudt = (float*)stage(out->dudt, nv); dv.dvdt = (float*)stage(out->dvdt, nv);
    dv.nw = e->has_nw ? (float*)stage(out->nw, nv) : nullptr;
    dv.rho = (float*)stage(out->rho, ns); dv.p = (float*)stage(out->p, ns);
    dv.drhodt = (float*)stage(out->drhodt, ns); dv.mass = (float*)stage(out->mass, ns);
    dv.eta = (float*)stage(out->eta, ns); dv.T = (float*)stage(out->T, ns);
    dv.dTdt = (float*)stage(out->dTdt, ns);
    dv.kappa = e->has_kc ? (float*)stage(out->kappa, ns) : nullptr;
    dv.Cp = e->has_kc ? (float*)stage(out->Cp, ns) : nullptr;
    dv.tag = (int32_t*)stage(out->tag, ns);
    if (ids) dids = (int32_t*)stage(ids, ns);
  }
  if (mask != DL_ALL) {
    if (!on_host || ids) return SPHB200_EINVAL;
    auto drop = [&](void* dptr) {
      int m = 0;
      for (int i = 0; i < nback; ++i)
        if (back[i].dptr != dptr) back[m++] = back[i];
      nback = m;
    };
    struct Sel { float** f; unsigned bit; } sel[] = {
        {&dv.r, DL_R}, {&dv.u, DL_U}, {&dv.v, DL_V}, {&dv.rho, DL_RHO}, {&dv.p, DL_P},
        {&dv.dudt, DL_DUDT}, {&dv.dvdt, DL_DVDT}, {&dv.nw, DL_REST}, {&dv.drhodt, DL_REST},
        {&dv.mass, DL_REST}, {&dv.eta, DL_REST}, {&dv.T, DL_REST}, {&dv.dTdt, DL_REST},
        {&dv.kappa, DL_REST}, {&dv.Cp, DL_REST}};
    for (const Sel& q : sel)
      if (!(mask & q.bit) && *q.f) {
        drop(*q.f);
        *q.f = nullptr;
      }
    if (!(mask & DL_REST) && dv.tag) {
      drop(dv.tag);
      dv.tag = nullptr;
    }
    if (nback == 0) return SPHB200_OK;
  }
  StateOut so{dv.r, dv.u, dv.v, dv.dudt, dv.dvdt, dv.nw, dv.rho, dv.p, dv.drhodt, dv.mass,
              dv.eta, dv.T, dv.dTdt, dv.kappa, dv.Cp, dv.tag, dids};
  const int nb = (n + 255) / 256;
  if (d == 2) k_unpack<2><<<nb, 256, 0, st>>>(n, e->slab, e->fr[e->cur], so);
  else k_unpack<3><<<nb, 256, 0, st>>>(n, e->slab, e->fr[e->cur], so);
  e->launches++;
  CK(cudaGetLastError());
  for (int i = 0; i < nback; ++i)
    CK(cudaMemcpyAsync(back[i].h, back[i].dptr, back[i].bytes, cudaMemcpyDeviceToHost, st));
  return SPHB200_OK;
}

int sphb200_engine_download(sphb200_engine* e, sphb200_state* out, int on_host, void* stream) {
  if (!e || e->slab_on) return SPHB200_EINVAL;
  return download_impl(e, out, e->n, nullptr, on_host, stream);
}

int sphb200_engine_advance_host(sphb200_engine* e, double dt, const sphb200_state* in,
                                sphb200_state* out, uint32_t flags, void* stream) {
  if (!e || !in || !out || e->slab_on) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (!e->io_on) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = greatest priority
    if (cudaStreamCreateWithPriority(&e->io_side, cudaStreamNonBlocking, hi) != cudaSuccess)
      return SPHB200_ECUDA;
    if (cudaEventCreateWithFlags(&e->ev_io_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_io_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaStreamDestroy(e->io_side);
      return SPHB200_ECUDA;
    }
    e->io_on = true;
  }
  // the same particles call after call: the state goes into the slots they occupy, cells and
  // neighbour lists survive as far as the positions allow (sphb200_engine_refresh)
  int rc = upload_impl(e, in, e->n, nullptr, 1, stream, true);
  if (rc) return rc;
  Kick k;
  k.dt = (float)dt;
  k.c2 = (float)(e->cfg.tvf * 0.5) * (float)dt;
  k.on = (flags & SPHB200_STEP_INTEGRATE) ? 1 : 0;
  const bool v_is_u = k.on && e->cfg.tvf == 0.0;
  if (e->profile) cudaEventRecord(e->ev[0], st);
  rc = begin_step(e, k, st);
  if (rc) return rc;
  if (e->profile) cudaEventRecord(e->ev[2], st);
  // r is final once the particles are integrated and reordered; u and v too unless the wall
  // sweep (generalized wall BC) or the case's bc table rewrites them after forward().  With
  // v == u (no transport velocity) the frame's v is written by the force sweep: not early.
  // rho and p are final after the density / renormalisation / wall stages (the force sweep
  // reads them only) unless the bc table sets p afterwards.
  const bool bc_on = bc_table_on(e->cfg);
  const bool early_uv = !(e->cfg.flags & SPHB200_F_BC_TRICK) && !bc_on && !v_is_u;
  const unsigned m_early = DL_R | (early_uv ? (DL_U | DL_V) : 0u);
  const unsigned m_mid = bc_on ? 0u : (DL_RHO | DL_P);
  CK(cudaEventRecord(e->ev_io_fork, st));
  CK(cudaStreamWaitEvent(e->io_side, e->ev_io_fork, 0));
  rc = download_impl(e, out, e->n, nullptr, 1, (void*)e->io_side, m_early);
  if (rc) return rc;
  for (int stage = 0; stage < 4; ++stage) {  // run_forward, with the rho / p copy before the force stage
    if (stage == 3 && m_mid) {
      CK(cudaEventRecord(e->ev_io_fork, st));
      CK(cudaStreamWaitEvent(e->io_side, e->ev_io_fork, 0));
      rc = download_impl(e, out, e->n, nullptr, 1, (void*)e->io_side, m_mid);
      if (rc) return rc;
    }
    int wrote;
    rc = forward_stage(e, stage, flags, v_is_u, st, &wrote);
    if (rc) return rc;
  }
  CK(cudaEventRecord(e->ev_io_join, e->io_side));
  if (e->profile) cudaEventRecord(e->ev[5], st);
  rc = download_impl(e, out, e->n, nullptr, 1, stream, DL_ALL & ~(m_early | m_mid));
  if (rc) return rc;
  CK(cudaStreamWaitEvent(st, e->ev_io_join, 0));  // the caller's stream sees the whole result
  return SPHB200_OK;
}

int sphb200_engine_step(sphb200_engine* e, double dt, int nsteps, uint32_t flags, void* stream) {
  if (!e || nsteps < 0 || e->slab_on) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  Kick k;
  k.dt = (float)dt;
  k.c2 = (float)(e->cfg.tvf * 0.5) * (float)dt;
  k.on = (flags & SPHB200_STEP_INTEGRATE) ? 1 : 0;
  const bool v_is_u = k.on && e->cfg.tvf == 0.0;
  for (int s = 0; s < nsteps; ++s) {
    if (e->profile) cudaEventRecord(e->ev[0], st);
    int rc = begin_step(e, k, st);
    if (rc) return rc;
    if (e->profile) cudaEventRecord(e->ev[2], st);
    rc = run_forward(e, flags, v_is_u, st);
    if (rc) return rc;
    if (e->profile) cudaEventRecord(e->ev[5], st);
  }
  return SPHB200_OK;
}

int sphb200_engine_vjp(sphb200_engine* e, double dt, uint32_t flags, const sphb200_state* cot_out,
                       sphb200_state* cot_in, void* stream) {
  if (!e || !cot_out || !cot_in) return SPHB200_EINVAL;
  const sphb200_config& c = e->cfg;
  // the variant adjoint.cuh differentiates (the setting of notebooks/iclr24_grads.ipynb, cell 5)
  if (e->slab_on || c.solver != SPHB200_SOLVER_SPH || c.tvf != 0.0 || c.artificial_alpha != 0.0 ||
      (c.flags & (SPHB200_F_BC_TRICK | SPHB200_F_RHO_EVOL | SPHB200_F_HEAT | SPHB200_F_FREE_SLIP)) ||
      bc_table_on(c) || (c.g_mode != SPHB200_G_NONE && c.g_mode != SPHB200_G_CONST) ||
      (c.kernel != SPHB200_KERNEL_QSK && c.kernel != SPHB200_KERNEL_WC2K))
    return SPHB200_EUNSUP;
  if (!e->cells_valid) return SPHB200_EINVAL;  // no step has been run: nothing to linearise about
  cudaStream_t st = (cudaStream_t)stream;
  const int n = e->n;
  if (!e->adj_mem) CK(cudaMalloc((void**)&e->adj_mem, (size_t)n * 64));
  AdjBufs ab{e->adj_mem, e->adj_mem + n, e->adj_mem + 2 * (size_t)n, e->adj_mem + 3 * (size_t)n};
  CotPtrs ct{cot_out->r, cot_out->u, cot_out->v, cot_out->dudt, cot_out->rho, cot_out->p};
  CotOut out{cot_in->r, cot_in->u, cot_in->v, cot_in->dudt, cot_in->dvdt, cot_in->rho, cot_in->p};
  Frame& F = e->fr[e->cur];
  const int nb = (n + 255) / 256;
  const int integrate = (flags & SPHB200_STEP_INTEGRATE) ? 1 : 0;
  if (e->dim == 2) k_adj_begin<2><<<nb, 256, 0, st>>>(n, F, ab, ct);
  else k_adj_begin<3><<<nb, 256, 0, st>>>(n, F, ab, ct);
  CK(cudaGetLastError());
  Extra ex = make_extra();
  ex.adj = ab;
  ex.nq = 4;
  int rc = SPHB200_OK;
  const SweepPlan pf = plan_sweep(e, 4, e->lcap), pd = plan_sweep(e, 1, e->lcap);
  const bool qsk = c.kernel == SPHB200_KERNEL_QSK;
#define ADJ(P, PLAN)                                                                                 \
  do {                                                                                               \
    if (e->dim == 2) {                                                                               \
      if (qsk) rc = launch_sweep(e, k_sweep<2, P<2, SPHB200_KERNEL_QSK>, LIST_NONE>, PLAN, F, ex, st); \
      else rc = launch_sweep(e, k_sweep<2, P<2, SPHB200_KERNEL_WC2K>, LIST_NONE>, PLAN, F, ex, st);   \
    } else {                                                                                         \
      if (qsk) rc = launch_sweep(e, k_sweep<3, P<3, SPHB200_KERNEL_QSK>, LIST_NONE>, PLAN, F, ex, st); \
      else rc = launch_sweep(e, k_sweep<3, P<3, SPHB200_KERNEL_WC2K>, LIST_NONE>, PLAN, F, ex, st);   \
    }                                                                                                \
  } while (0)
  ADJ(PhysForceAdj, pf);
  if (rc) return rc;
  k_adj_eos<<<nb, 256, 0, st>>>(n, e->consts, F, ab, ct);
  CK(cudaGetLastError());
  ex.nq = 1;
  ADJ(PhysDensAdj, pd);
#undef ADJ
  if (rc) return rc;
  if (e->dim == 2) k_adj_finish<2><<<nb, 256, 0, st>>>(n, F, ab, ct, out, (float)dt, integrate);
  else k_adj_finish<3><<<nb, 256, 0, st>>>(n, F, ab, ct, out, (float)dt, integrate);
  CK(cudaGetLastError());
  e->launches += 3;
  return SPHB200_OK;
}

int sphb200_engine_error(sphb200_engine* e, uint32_t* code, void* stream) {
  if (!e || !code) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned h = 0;
  CK(cudaMemcpyAsync(&h, e->err, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(e->err, 0, 4, st));
  CK(cudaStreamSynchronize(st));
  *code = h;
  return SPHB200_OK;
}

int sphb200_engine_neighbor_list(sphb200_engine* e, int32_t* idx, int64_t capacity, int mask_self,
                                 int64_t* count, void* stream) {
  if (!e || e->slab_on || capacity < 0 || capacity >= 2147483647LL) return SPHB200_EINVAL;
  if (!idx && capacity != 0) return SPHB200_EINVAL;  // idx == NULL: count only
  cudaStream_t st = (cudaStream_t)stream;
  if (!e->cells_valid || e->maybe_drifted) {  // the materialiser searches fresh cells
    int rc = sort_now(e, st);
    if (rc) return rc;
  }
  const int n = e->n;
  int rc = 0;
  Frame& F = e->fr[e->cur];
  Extra ex = make_extra();
  ex.nq = 2;
  ex.nl_counts = e->nl_counts;
  ex.nl_mask_self = mask_self;
  ex.nl_n = n;
  ex.nl_idx = idx;
  ex.nl_capacity = capacity;
  ex.nl_fill = 0;
  const NList none{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr};
  if (e->dim == 2) rc = launch_sweep(e, k_sweep<2, PhysNeighbors<2>>, e->planN, F, ex, st, none, nullptr, false, true);
  else rc = launch_sweep(e, k_sweep<3, PhysNeighbors<3>>, e->planN, F, ex, st, none, nullptr, false, true);
  if (rc) return rc;
  // exclusive scan of the per-sender counts (original order) -> offsets (reuse rnk as scratch)
  const int sb = (n + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_partial<<<sb, SCAN_TPB, 0, st>>>(n, e->nl_counts, e->bsum);
  k_scan_bsum<<<1, 1024, 0, st>>>(sb, e->bsum);
  k_scan_final_keep<<<sb, SCAN_TPB, 0, st>>>(n, e->nl_counts, e->bsum, e->rnk);
  k_nl_total<<<1, 1, 0, st>>>(n, e->rnk, e->nl_counts, idx ? capacity : (1LL << 62),
                              (long long*)count, e->err);
  e->launches += 4;
  if (!idx) {
    CK(cudaGetLastError());
    return SPHB200_OK;
  }
  const long long tot = 2 * capacity;
  if (tot > 0) k_fill<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(idx, tot, n);
  e->launches += 1;
  ex.nl_fill = 1;
  ex.nl_offsets = e->rnk;
  if (e->dim == 2) rc = launch_sweep(e, k_sweep<2, PhysNeighbors<2>>, e->planN, F, ex, st, none, nullptr, false, true);
  else rc = launch_sweep(e, k_sweep<3, PhysNeighbors<3>>, e->planN, F, ex, st, none, nullptr, false, true);
  CK(cudaGetLastError());
  return rc;
}

int sphb200_engine_stats(sphb200_engine* e, double* ekin, double* u_max, void* stream) {
  if (!e) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(e->stats, 0, 16, st));
  const int bound = e->slab_on ? e->sgeom.own_cap : e->n;
  const int nb = (bound + 255) / 256;
  if (e->dim == 2) k_stats<2><<<nb, 256, 0, st>>>(bound, e->slab, e->fr[e->cur], e->stats);
  else k_stats<3><<<nb, 256, 0, st>>>(bound, e->slab, e->fr[e->cur], e->stats);
  e->launches++;
  double h[2];
  CK(cudaMemcpyAsync(h, e->stats, 16, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (ekin) *ekin = h[0];
  if (u_max) *u_max = h[1];
  return SPHB200_OK;
}

// ---- on-device case initialisation (row f1; kernel in init.cuh) ----------------------------
int64_t sphb200_lattice_rows(const sphb200_lattice* l) {
  if (!l || l->struct_size != sizeof(sphb200_lattice)) return SPHB200_EINVAL;
  if (l->dim != 2 && l->dim != 3) return SPHB200_EINVAL;
  for (int d = 0; d < l->dim; ++d)
    if (l->n[d] <= 0) return SPHB200_EINVAL;
  if (l->k_lo < 0 || l->k_hi < l->k_lo || l->k_hi > l->n[l->dim - 1]) return SPHB200_EINVAL;
  if (l->wall_axis >= l->dim || (l->wall_axis >= 0 && l->n_walls < 0)) return SPHB200_EINVAL;
  if (l->velocity < SPHB200_VEL_REST || l->velocity > SPHB200_VEL_TGV3D) return SPHB200_EINVAL;
  if (l->velocity == SPHB200_VEL_TGV3D && l->dim != 3) return SPHB200_EINVAL;
  if (!(l->dx > 0.f)) return SPHB200_EINVAL;
  int64_t plane = l->n[0];
  if (l->dim == 3) plane *= l->n[1];
  const int64_t full = plane * l->n[l->dim - 1];
  if (full > INT32_MAX) return SPHB200_EINVAL;  // ids are int32, like the reference's indices
  return plane * (l->k_hi - l->k_lo);
}

int sphb200_init_lattice(const sphb200_lattice* l, sphb200_state* out, int32_t* ids, void* stream) {
  const int64_t rows = sphb200_lattice_rows(l);
  if (rows < 0) return (int)rows;
  if (!out) return SPHB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHB200_ENODEV;
  if (rows == 0) return SPHB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t vb = (size_t)rows * l->dim * sizeof(float), sb = (size_t)rows * sizeof(float);
  float* zv[] = {out->dudt, out->dvdt, out->nw};
  for (float* p : zv)
    if (p) CK(cudaMemsetAsync(p, 0, vb, st));
  float* zs[] = {out->drhodt, out->dTdt};
  for (float* p : zs)
    if (p) CK(cudaMemsetAsync(p, 0, sb, st));
  LatticeArgs a;
  a.l = *l;
  a.out = *out;
  a.ids = ids;
  a.rows = rows;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (rows + 255) / 256;
  const int nb = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  if (l->dim == 2) k_init_lattice<2><<<nb, 256, 0, st>>>(a);
  else k_init_lattice<3><<<nb, 256, 0, st>>>(a);
  CK(cudaGetLastError());
  return SPHB200_OK;
}

int sphb200_engine_get_stats(sphb200_engine* e, double out[SPHB200_NSTATS], void* stream) {
  if (!e || !out) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  // min slots start at the largest key, everything else at zero
  unsigned long long init[SPHB200_NSTATS];
  memset(init, 0, sizeof(init));
  for (int k = 0; k < 5; ++k) init[1 + 3 * k] = ~0ull;
  CK(cudaMemcpyAsync(e->stats, init, sizeof(init), cudaMemcpyHostToDevice, st));
  const int bound = e->slab_on ? e->sgeom.own_cap : e->n;
  int nb = (bound + 255) / 256;
  if (nb > 148 * 8) nb = 148 * 8;
  if (e->dim == 2) k_get_stats<2><<<nb, 256, 0, st>>>(bound, e->slab, e->fr[e->cur], e->stats);
  else k_get_stats<3><<<nb, 256, 0, st>>>(bound, e->slab, e->fr[e->cur], e->stats);
  CK(cudaGetLastError());
  e->launches++;
  unsigned long long h[SPHB200_NSTATS];
  CK(cudaMemcpyAsync(h, e->stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memset(out, 0, SPHB200_NSTATS * sizeof(double));
  double d[SPHB200_NSTATS];
  memcpy(d, h, sizeof(h));
  const double count = d[16];
  out[16] = count;
  double vol = 1.0;
  for (int a = 0; a < e->dim; ++a) vol *= e->cfg.dx;
  out[0] = 0.5 * d[0] * vol;  // utils.py:133
  for (int k = 0; k < 5; ++k) {
    for (int m = 0; m < 2; ++m) {
      unsigned key = (unsigned)h[1 + 3 * k + m];
      unsigned bits = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
      float v;
      memcpy(&v, &bits, 4);
      out[1 + 3 * k + m] = count > 0 ? (double)v : 0.0;
    }
    out[3 + 3 * k] = d[3 + 3 * k];
  }
  return SPHB200_OK;
}

int sphb200_eval_velocity(int32_t dim, int64_t n, int32_t velocity, const float* r, float* u,
                          float* v, void* stream) {
  if ((dim != 2 && dim != 3) || n < 0 || !r) return SPHB200_EINVAL;
  if (velocity < SPHB200_VEL_REST || velocity > SPHB200_VEL_TGV3D) return SPHB200_EINVAL;
  if (velocity == SPHB200_VEL_TGV3D && dim != 3) return SPHB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHB200_ENODEV;
  if (n == 0) return SPHB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = (n + 255) / 256;
  const int nb = (int)(want < 148 * 8 ? want : 148 * 8);
  if (dim == 2) k_eval_velocity<2><<<nb, 256, 0, st>>>(n, velocity, r, u, v);
  else k_eval_velocity<3><<<nb, 256, 0, st>>>(n, velocity, r, u, v);
  CK(cudaGetLastError());
  return SPHB200_OK;
}

int sphb200_add_noise(int32_t dim, int64_t n, float* r, const int32_t* tag, const int32_t* ids,
                      double std, uint64_t seed, const double box[3], void* stream) {
  if ((dim != 2 && dim != 3) || n < 0 || !r || !box || !(std >= 0.0)) return SPHB200_EINVAL;
  for (int d = 0; d < dim; ++d)
    if (!(box[d] > 0.0)) return SPHB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHB200_ENODEV;
  if (n == 0 || std == 0.0) return SPHB200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = (n + 255) / 256;
  const int nb = (int)(want < 148 * 8 ? want : 148 * 8);
  const float b2 = dim == 3 ? (float)box[2] : 1.0f;
  if (dim == 2)
    k_add_noise<2><<<nb, 256, 0, st>>>(n, r, tag, ids, (float)std, seed, (float)box[0], (float)box[1], b2);
  else
    k_add_noise<3><<<nb, 256, 0, st>>>(n, r, tag, ids, (float)std, seed, (float)box[0], (float)box[1], b2);
  CK(cudaGetLastError());
  return SPHB200_OK;
}

int64_t sphb200_engine_launches(const sphb200_engine* e) { return e ? e->launches : -1; }

int sphb200_engine_profile(sphb200_engine* e, int enable) {
  if (!e) return SPHB200_EINVAL;
  if (enable && !e->profile) {
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&e->ev[i]));
    e->profile = true;
  } else if (!enable && e->profile) {
    for (int i = 0; i < 8; ++i) cudaEventDestroy(e->ev[i]);
    e->profile = false;
  }
  return SPHB200_OK;
}

int sphb200_engine_last_times(sphb200_engine* e, float ms[8]) {
  if (!e || !ms || !e->profile) return SPHB200_EINVAL;
  CK(cudaEventSynchronize(e->ev[5]));
  memset(ms, 0, 8 * sizeof(float));
  // ev0 start, ev2 cells done, ev3 density done, ev4 wall done, ev5 force done
  CK(cudaEventElapsedTime(&ms[1], e->ev[0], e->ev[2]));
  CK(cudaEventElapsedTime(&ms[2], e->ev[2], e->ev[3]));
  CK(cudaEventElapsedTime(&ms[3], e->ev[3], e->ev[4]));
  CK(cudaEventElapsedTime(&ms[4], e->ev[4], e->ev[5]));
  CK(cudaEventElapsedTime(&ms[5], e->ev[0], e->ev[5]));
  return SPHB200_OK;
}

int sphb200_engine_plan(const sphb200_engine* e, int32_t out[16]) {
  if (!e || !out) return SPHB200_EINVAL;
  for (int a = 0; a < 3; ++a) {
    out[a] = e->grid.n[a];
    out[3 + a] = e->grid.S[a];
    out[6 + a] = e->grid.T[a];
  }
  out[9] = e->duo ? e->duo_tpb : e->tpb;
  out[10] = e->lcap;
  out[11] = e->planA.cap;
  out[12] = e->planW.cap;
  out[13] = e->planC.cap;
  out[14] = e->grid.exact_all;
  out[15] = e->grid.ncells;
  return SPHB200_OK;
}

int sphb200_engine_counters(sphb200_engine* e, int64_t out[8], void* stream) {
  if (!e || !out) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = e->grid.nt[0] * e->grid.nt[1] * e->grid.nt[2];
  int h[4] = {0, 0, 0, 0};
  CK(cudaMemsetAsync(e->ctl + 3, 0, 4, st));
  if (e->pl_ok && e->cells_valid) {
    k_count_bad<<<64, 256, 0, st>>>(tiles, e->pl_ok, e->ctl + 3);
    e->launches++;
  }
  CK(cudaMemcpyAsync(h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < 8; ++i) out[i] = 0;
  out[0] = (int64_t)e->step_no;
  out[1] = h[2];
  out[2] = e->duo ? e->duo_lmax : e->pl_lmax;
  out[6] = e->duo ? 1 : 0;
  out[3] = (int64_t)(e->skin_frac * 1e6 + 0.5);
  out[4] = tiles;
  out[5] = e->pl_ok ? h[3] : tiles;
  out[7] = -1;
  if (e->duo && e->cells_valid && e->pl_ok && h[3] == 0 && e->step_no > 0) {
    // directed pairs in the exact lists of the last step (diagnostics: k_duo_pair_count)
    unsigned long long* dcount = reinterpret_cast<unsigned long long*>(e->stats);
    unsigned long long hc = 0ull;
    CK(cudaMemsetAsync(dcount, 0, 8, st));
    DuoList dl{e->duo_desc, e->duo_desc_stride, e->sl_list, e->pl_list, e->sl_cnt, e->pl_cnt,
               e->pl_ok,    e->duo_lmax,        0,          e->duo_rows};
    k_duo_pair_count<<<tiles, 256, 0, st>>>(dl, dcount);
    e->launches++;
    CK(cudaMemcpyAsync(&hc, dcount, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out[7] = (int64_t)hc;
  }
  return SPHB200_OK;
}

int sphb200_fp32_peak(int packed, double* tflops, double* ms, void* stream) {
  if (!tflops) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8, threads = 256, iters = 1 << 14;
  float* out = nullptr;
  if (cudaMalloc((void**)&out, (size_t)blocks * threads * 4) != cudaSuccess) return SPHB200_ENOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {  // first round warms up
    cudaEventRecord(e0, st);
    if (packed) k_fma_peak<true><<<blocks, threads, 0, st>>>(out, iters, 0.999f, 1e-3f);
    else k_fma_peak<false><<<blocks, threads, 0, st>>>(out, iters, 0.999f, 1e-3f);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1);
    if (rep > 0 && t < best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return SPHB200_ECUDA;
  const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;  // 64 lane-FMAs per iteration
  *tflops = flop / (best * 1e-3) * 1e-12;
  if (ms) *ms = best;
  return SPHB200_OK;
}

// ---- slab decomposition (one engine per rank; the caller moves the messages) ----------------
static int slab_spec(const sphb200_config* cfg, int rank, int nranks, int64_t own_cap,
                     int64_t halo_cap, int64_t mig_cap, SlabSpec* sp) {
  if (nranks < 2 || rank < 0 || rank >= nranks) return SPHB200_EINVAL;
  Grid g;
  plan_grid(*cfg, g, cfg->threads > 0 ? cfg->threads : DEFAULT_TPB, 0, 1, plan_skin(*cfg));
  const int ax = cfg->dim - 1;
  if (g.exact_all && g.n[ax] < 2 * g.S[ax] + 1) return SPHB200_EINVAL;
  for (int r = 0; r < nranks; ++r) {
    int z0, z1;
    slab_range(g.n[ax], r, nranks, z0, z1);
    if (z1 - z0 < 2 * g.S[ax]) return SPHB200_EINVAL;  // a slab must be at least two cutoffs thick
  }
  double per_layer = 1.0;  // nominal particles in one cell layer along the slab axis
  for (int a = 0; a < cfg->dim; ++a) per_layer *= cfg->box[a] / cfg->dx;
  per_layer /= g.n[ax];
  int z0, z1;
  slab_range(g.n[ax], rank, nranks, z0, z1);
  if (own_cap <= 0) own_cap = (int64_t)(1.25 * per_layer * (z1 - z0)) + 4096;
  if (halo_cap <= 0) halo_cap = (int64_t)(1.5 * per_layer * g.S[ax]) + 4096;
  if (mig_cap <= 0) mig_cap = (int64_t)(0.25 * per_layer) + 1024;
  if (2 * mig_cap > halo_cap) halo_cap = 2 * mig_cap;  // immigrants are parked in the upper halo slots
  if (2 * halo_cap + own_cap > 2000000000LL) return SPHB200_EINVAL;
  sp->rank = rank; sp->nranks = nranks;
  sp->own_cap = (int)own_cap; sp->halo_cap = (int)halo_cap; sp->mig_cap = (int)mig_cap;
  return SPHB200_OK;
}

int sphb200_slab_create(const sphb200_config* cfg, int rank, int nranks, int64_t own_cap,
                        int64_t halo_cap, int64_t mig_cap, sphb200_engine** out) {
  int rc = validate(cfg, 1);
  if (rc) return rc;
  if (!out) return SPHB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return SPHB200_ENODEV;
  SlabSpec sp;
  rc = slab_spec(cfg, rank, nranks, own_cap, halo_cap, mig_cap, &sp);
  if (rc) return rc;
  Grid g;
  DuoPlan dp;
  plan_all(*cfg, slab_slots(sp), rank, nranks, g, dp);
  Layout L;
  plan_layout(*cfg, slab_slots(sp), g, plan_skin(*cfg), L, dp);
  void* ws = nullptr;
  if (cudaMalloc(&ws, L.total) != cudaSuccess) return SPHB200_ENOMEM;
  sphb200_engine* e = new (std::nothrow) sphb200_engine();
  if (!e) {
    cudaFree(ws);
    return SPHB200_ENOMEM;
  }
  rc = init_engine(e, cfg, slab_slots(sp), ws, L.total, true, &sp);
  if (rc) {
    cudaFree(ws);
    delete e;
    return rc;
  }
  // overlap of the halo exchanges with the interior tile layers (prelaunch_interior): tile
  // layer t covers own cell layers [t T, (t + 1) T); its stencil stays inside the own layers
  // iff t T >= S and (t + 1) T + S <= n_own
  {
    const Grid& gg = e->grid;
    const int ax = e->dim - 1, T = gg.T[ax], S = gg.S[ax], nown = gg.own_hi[ax] - gg.own_lo[ax];
    e->int_lo = (S + T - 1) / T;
    e->int_hi = (nown - S) / T;
    if (e->int_hi > gg.nt[ax]) e->int_hi = gg.nt[ax];
    if (e->int_hi < e->int_lo) e->int_hi = e->int_lo;
    const char* off = getenv("SPHB200_SLAB_OVERLAP");
    if (!(off && off[0] == '0') && e->int_hi > e->int_lo &&
        cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) == cudaSuccess) {
      if (cudaEventCreateWithFlags(&e->ev_main, cudaEventDisableTiming) == cudaSuccess &&
          cudaEventCreateWithFlags(&e->ev_side, cudaEventDisableTiming) == cudaSuccess)
        e->overlap = true;
      else
        cudaStreamDestroy(e->side);
    }
  }
  *out = e;
  return SPHB200_OK;
}

static int slab_mask_a(const sphb200_engine* e) {
  return HX_PT | HX_UM | HX_VV | HX_ST | HX_CELLS | (e->has_kc ? HX_KC : 0) |
         (e->has_nw ? HX_NW : 0) | (e->has_ge ? HX_GE : 0);
}

int sphb200_slab_info(const sphb200_engine* e, int64_t outi[16], double outd[4]) {
  if (!e || !e->slab_on || !outi || !outd) return SPHB200_EINVAL;
  const int ax = e->slab_axis;
  const SlabGeom& sg = e->sgeom;
  size_t xb = mig_bytes(e->slab.mig_cap);
  const size_t hb = halo_bytes(slab_mask_a(e) | HX_UT, sg.halo_cap, sg.ncl);
  if (hb > xb) xb = hb;
  outi[0] = e->slab_rank; outi[1] = e->slab_nranks; outi[2] = ax;
  outi[3] = e->slab_z0; outi[4] = e->slab_z1; outi[5] = e->grid.ng[ax];
  outi[6] = sg.own_cap; outi[7] = sg.halo_cap; outi[8] = e->slab.mig_cap;
  outi[9] = (int64_t)xb;  // bytes each of the four message buffers must hold
  outi[10] = e->grid.S[ax]; outi[11] = sg.cap_total; outi[12] = (int64_t)e->arena_bytes;
  outi[13] = outi[14] = outi[15] = 0;
  outd[0] = (double)e->grid.inv_cell[ax];  // layer = min(int(f32(r) * f32(inv_cell)), ng - 1)
  outd[1] = (double)e->grid.box[ax];
  outd[2] = outd[3] = 0.0;
  return SPHB200_OK;
}

int sphb200_slab_upload(sphb200_engine* e, const sphb200_state* s, const int32_t* ids, int64_t rows,
                        int on_host, void* stream) {
  if (!e || !e->slab_on || rows < 0 || rows > e->sgeom.own_cap || !ids) return SPHB200_EINVAL;
  return upload_impl(e, s, (int)rows, ids, on_host, stream);
}

int sphb200_slab_download(sphb200_engine* e, sphb200_state* out, int32_t* ids, int64_t rows,
                          int on_host, void* stream) {
  if (!e || !e->slab_on || rows < 0 || rows > e->sgeom.own_cap || !ids) return SPHB200_EINVAL;
  return download_impl(e, out, (int)rows, ids, on_host, stream);
}

int sphb200_slab_counts(sphb200_engine* e, int32_t out[8], void* stream) {
  if (!e || !e->slab_on || !out) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(out, e->dn, DN_WORDS * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return SPHB200_OK;
}

// ---- slab overlap ---------------------------------------------------------------------------
// The stage the next sweep launch belongs to (stages 1 and 2 exist only for some variants).
static int next_effective_stage(const sphb200_engine* e, int from) {
  const sphb200_config& c = e->cfg;
  const bool evol = c.flags & SPHB200_F_RHO_EVOL, renorm = c.flags & SPHB200_F_RHO_RENORM;
  const bool wall_sweep = (c.flags & SPHB200_F_BC_TRICK) && c.solver != SPHB200_SOLVER_RIE;
  for (int s = from; s < 4; ++s) {
    if (s == 1 && !(evol && renorm)) continue;
    if (s == 2 && !wall_sweep) continue;
    return s;
  }
  return 4;
}

// While the halo message just packed is in flight (the caller enqueues the exchange on `st`
// after this returns), start the interior tile layers of stage `stage` (0 density, 3 force) on
// the side stream: their stencils touch no halo layer, so they need nothing from the message.
static int prelaunch_interior(sphb200_engine* e, int stage, cudaStream_t st) {
  e->pre_stage = -1;
  // per-pass event timing (sphb200_engine_profile) measures the sweeps whole, on one stream
  if (!e->overlap || e->profile || (stage != 0 && stage != 3) || e->int_hi <= e->int_lo)
    return SPHB200_OK;
  // the three sweeps of the Delta-SPH density diffusion need their own halo refreshes
  if (stage == 0 && e->cfg.solver == SPHB200_SOLVER_DELTA && (e->cfg.flags & SPHB200_F_RHO_EVOL))
    return SPHB200_OK;
  CK(cudaEventRecord(e->ev_main, st));
  CK(cudaStreamWaitEvent(e->side, e->ev_main, 0));
  int wrote = 0;
  int rc = forward_stage(e, stage, e->slab_flags, e->slab_v_is_u, e->side, &wrote, 1);
  if (rc) return rc;
  CK(cudaEventRecord(e->ev_side, e->side));
  e->pre_stage = stage;
  return SPHB200_OK;
}

// a neighbour that never answers costs this long, then the error word says so
static const unsigned long long SLAB_SIGNAL_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

int sphb200_slab_signal(sphb200_engine* e, void* stream) {
  if (!e || !e->slab_on || e->agree.n <= 0) return SPHB200_EINVAL;
  const int n = e->agree.n, r = e->slab_rank;
  k_ring_signal<<<1, 32, 0, (cudaStream_t)stream>>>(e->agree, (r + n - 1) % n, (r + 1) % n, r, ++e->ring_seq,
                                                    e->err, SLAB_SIGNAL_TIMEOUT_NS);
  e->launches++;
  CK(cudaGetLastError());
  return SPHB200_OK;
}

int sphb200_slab_set_agree(sphb200_engine* e, int32_t* const* flag_arrays, int nranks) {
  if (!e || !e->slab_on) return SPHB200_EINVAL;
  if (!flag_arrays || nranks <= 0) {
    e->agree.n = 0;
    return SPHB200_OK;
  }
  if (nranks != e->slab_nranks || nranks > 16) return SPHB200_EINVAL;
  for (int r = 0; r < nranks; ++r) {
    if (!flag_arrays[r]) return SPHB200_EINVAL;
    e->agree.p[r] = flag_arrays[r];
  }
  e->agree.n = nranks;
  return SPHB200_OK;
}

int sphb200_slab_run(sphb200_engine* e, int phase, double dt, uint32_t flags, void* send_lo,
                     void* send_hi, const void* recv_lo, const void* recv_hi, void* stream,
                     int64_t* xbytes) {
  if (!e || !e->slab_on || !xbytes || phase < 0) return SPHB200_EINVAL;
  if (!send_lo || !send_hi || !recv_lo || !recv_hi) return SPHB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  Slab& sl = e->slab;
  const SlabGeom& sg = e->sgeom;
  const int hblocks = 4 * 148;
  *xbytes = 0;
  const Kick off{0.f, 0.f, 0};
  if (phase == 0) {
    // kick + drift + wrap in place; this rank's re-sort decision goes into a message word that
    // the transport max-reduces over the ranks (xbytes = -4): all ranks sort, or none does
    Kick k;
    k.dt = (float)dt;
    k.c2 = (float)(e->cfg.tvf * 0.5) * (float)dt;
    k.on = (flags & SPHB200_STEP_INTEGRATE) ? 1 : 0;
    e->slab_kick = k;
    e->slab_flags = flags;
    e->slab_v_is_u = k.on && e->cfg.tvf == 0.0;
    sl.mig_lo = (char*)send_lo;
    sl.mig_hi = (char*)send_hi;
    if (e->profile) cudaEventRecord(e->ev[0], st);
    int rc = drift_step(e, k, st);
    if (rc) return rc;
    if (e->agree.n > 0)
      k_flag_bcast<<<1, 32, 0, st>>>(e->gate_cur, e->agree, e->slab_rank,
                                     (int)((e->step_no - 1ull) & 1ull) * e->agree.n, (int)e->step_no);
    else
      k_flag_out<<<1, 1, 0, st>>>(e->gate_cur, (int*)send_lo);
    e->launches++;
    CK(cudaGetLastError());
    *xbytes = -4;
    return SPHB200_OK;
  }
  int* fc = e->ctl + ((e->step_no - 1ull) & 1ull);  // the flag word of this step (drift_step)
  if (phase == 1) {
    if (e->agree.n > 0)
      k_flag_in_max<<<1, 32, 0, st>>>(e->agree.p[e->slab_rank], e->agree.n,
                                      (int)((e->step_no - 1ull) & 1ull) * e->agree.n, (int)e->step_no, fc,
                                      e->err, SLAB_SIGNAL_TIMEOUT_NS);
    else
      k_flag_in<<<1, 1, 0, st>>>((const int*)send_lo, fc);
    int rc = hash_cells(e, off, st, fc);  // emigrants -> send buffers (on the steps that sort)
    if (rc) return rc;
    k_mig_header<<<1, 1, 0, st>>>(sl);
    e->launches += 2;
    CK(cudaGetLastError());
    *xbytes = (int64_t)mig_bytes(sl.mig_cap);
    return SPHB200_OK;
  }
  if (phase == 2) {
    const dim3 gi((sl.mig_cap + 255) / 256, 2);
    Frame& A = e->fr[e->cur];
    if (e->dim == 2)
      k_immigrate<2><<<gi, 256, 0, st>>>(e->grid, sl, sg, A, (const char*)recv_lo,
                                         (const char*)recv_hi, e->key, e->rnk, e->count, e->err);
    else
      k_immigrate<3><<<gi, 256, 0, st>>>(e->grid, sl, sg, A, (const char*)recv_lo,
                                         (const char*)recv_hi, e->key, e->rnk, e->count, e->err);
    int rc = sort_cells(e, off, st, fc);
    if (rc) return rc;
    k_slab_after_sort<<<1, 1, 0, st>>>(e->grid, sl, sg, e->start, e->err, fc);
    if (e->slab_kick.on) {
      rc = wall_normals(e, st);
      if (rc) return rc;
    }
    const int mask = slab_mask_a(e);
    k_halo_pack<<<dim3(hblocks, 2), 256, 0, st>>>(sl, sg, e->fr[e->cur], mask, e->start,
                                                  (char*)send_lo, (char*)send_hi, e->err);
    e->launches += 3;
    CK(cudaGetLastError());
    if (e->profile) cudaEventRecord(e->ev[2], st);
    e->slab_pending_mask = mask;
    e->slab_stage = 0;
    e->delta_sub = 0;
    *xbytes = (int64_t)halo_bytes(mask, sg.halo_cap, sg.ncl);
    return prelaunch_interior(e, next_effective_stage(e, 0), st);
  }
  // phase >= 3: take in the halo message of the previous phase, then sweep until the next
  // stage whose results the neighbours need
  if (e->slab_pending_mask) {
    const int mask = e->slab_pending_mask;
    if (mask & HX_CELLS) {
      k_halo_cells<<<2, 1024, 0, st>>>(e->grid, sl, sg, mask, (const char*)recv_lo,
                                       (const char*)recv_hi, e->start, e->err);
      e->launches++;
    }
    k_halo_unpack<<<dim3(hblocks, 2), 256, 0, st>>>(sl, sg, e->fr[e->cur], mask,
                                                    (const char*)recv_lo, (const char*)recv_hi);
    e->launches++;
    CK(cudaGetLastError());
    e->slab_pending_mask = 0;
  }
  while (e->slab_stage < 4) {
    int wrote = 0;
    const int stage = e->slab_stage;
    e->stage_more = false;
    int part = 0;
    if (e->pre_stage == stage) {  // its interior tiles ran on the side stream during the exchange
      CK(cudaStreamWaitEvent(st, e->ev_side, 0));
      e->pre_stage = -1;
      part = 2;
    }
    int rc = forward_stage(e, stage, e->slab_flags, e->slab_v_is_u, st, &wrote, part);
    if (rc) return rc;
    if (!e->stage_more) e->slab_stage++;
    if (wrote && stage < 3) {
      k_halo_pack<<<dim3(hblocks, 2), 256, 0, st>>>(sl, sg, e->fr[e->cur], wrote, e->start,
                                                    (char*)send_lo, (char*)send_hi, e->err);
      e->launches++;
      CK(cudaGetLastError());
      e->slab_pending_mask = wrote;
      *xbytes = (int64_t)halo_bytes(wrote, sg.halo_cap, sg.ncl);
      return prelaunch_interior(e, next_effective_stage(e, e->slab_stage), st);
    }
  }
  if (e->profile) cudaEventRecord(e->ev[5], st);
  return SPHB200_OK;
}

// ---- stateless entry points -------------------------------------------------
static int with_engine(const sphb200_config* cfg, int64_t n, void* ws, size_t ws_bytes,
                       sphb200_engine** e) {
  return sphb200_engine_create_in(cfg, n, ws, ws_bytes, e);
}

int sphb200_neighbor_list(const sphb200_config* cfg, int64_t n, const float* r, int32_t* idx,
                          int64_t capacity, int mask_self, int64_t* count, uint32_t* err,
                          void* ws, size_t ws_bytes, void* stream) {
  if (!r || !cfg) return SPHB200_EINVAL;
  sphb200_engine* e = nullptr;
  sphb200_config one = *cfg;  // a one-shot search: cells of half a cutoff, no skin
  one.skin = -1.f;
  int rc = with_engine(&one, n, ws, ws_bytes, &e);
  if (rc) return rc;
  sphb200_state s;
  memset(&s, 0, sizeof(s));
  s.r = const_cast<float*>(r);
  rc = sphb200_engine_upload(e, &s, 0, stream);
  if (!rc) rc = sphb200_engine_neighbor_list(e, idx, capacity, mask_self, count, stream);
  if (!rc && err)
    rc = cudaMemcpyAsync(err, e->err, 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess
             ? SPHB200_OK : SPHB200_ECUDA;
  sphb200_engine_destroy(e);
  return rc;
}

// The stateless entry points are pure functions of their arguments.  sphb200_advance_persistent
// is for a caller that OWNS its workspace between calls (nobody else writes it): handing in the
// same workspace again gets the engine that lives in it back: the particles are already
// cell-sorted there and the neighbour lists exist.
// The new state goes into the slots its particles occupy (k_pack keep_order) and k_drift tests
// every position against the sorted one, so the result never depends on what the workspace held
// -- only the cost does.  The cache maps workspace pointers to the host-side engine objects
// (config + particle count must match, else the workspace is re-initialised).
namespace {
struct CachedEngine {
  void* ws;
  size_t ws_bytes;
  int64_t n;
  sphb200_config cfg;
  sphb200_engine* e;
  unsigned long long stamp;
};
constexpr int MAX_CACHED = 8;
CachedEngine g_cache[MAX_CACHED];
unsigned long long g_stamp = 0;
std::mutex g_cache_mutex;
}  // namespace

// DL_* mask of the state entries one step of this solver variant changes (solver.py:930-947:
// everything else passes through advance() untouched); `nw` only with per-step wall normals.
static unsigned written_mask(const sphb200_engine* e, bool* rest_some) {
  const sphb200_config& c = e->cfg;
  const bool evol = c.flags & SPHB200_F_RHO_EVOL, heat = c.flags & SPHB200_F_HEAT;
  *rest_some = evol || heat || (e->has_nw && e->wl_n > 0);
  return DL_R | DL_U | DL_V | DL_RHO | DL_P | DL_DUDT | DL_DVDT;
}

// ordered: the caller keeps its arrays in ENGINE order (sphb200_advance_ordered): row p goes
// into slot p and comes back from slot p, in_order / out_order carry the particle labels through
// the sorts -- no permuted copy on either side of the step.
static int stateless_step(const sphb200_config* cfg, int64_t n, double dt, uint32_t flags,
                          const sphb200_state* in, sphb200_state* out, uint32_t* err, void* ws,
                          size_t ws_bytes, void* stream, bool persistent = false,
                          bool ordered = false, const int32_t* in_order = nullptr,
                          int32_t* out_order = nullptr) {
  if (!in || !out || !cfg || !ws) return SPHB200_EINVAL;
  if (ordered && (!persistent || !out_order)) return SPHB200_EINVAL;
  if (!persistent) {  // scratch workspace: nothing survives the call
    sphb200_engine* e = nullptr;
    int rc = with_engine(cfg, n, ws, ws_bytes, &e);
    if (rc) return rc;
    rc = sphb200_engine_upload(e, in, 0, stream);
    if (!rc) rc = sphb200_engine_step(e, dt, 1, flags, stream);
    if (!rc) rc = sphb200_engine_download(e, out, 0, stream);
    if (!rc && err)
      rc = cudaMemcpyAsync(err, e->err, 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess
               ? SPHB200_OK : SPHB200_ECUDA;
    sphb200_engine_destroy(e);
    return rc;
  }
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  CachedEngine* slot = nullptr;
  for (CachedEngine& c : g_cache)
    if (c.e && c.ws == ws) slot = &c;
  const bool hit = slot && slot->n == n && slot->ws_bytes == ws_bytes &&
                   memcmp(&slot->cfg, cfg, sizeof(*cfg)) == 0;
  int rc = SPHB200_OK;
  if (!hit) {
    if (slot) {
      sphb200_engine_destroy(slot->e);
      slot->e = nullptr;
    } else {
      slot = &g_cache[0];
      for (CachedEngine& c : g_cache) {  // a free entry, else the least recently used one
        if (!c.e) { slot = &c; break; }
        if (c.stamp < slot->stamp) slot = &c;
      }
      if (slot->e) {
        sphb200_engine_destroy(slot->e);
        slot->e = nullptr;
      }
    }
    sphb200_engine* e = nullptr;
    rc = with_engine(cfg, n, ws, ws_bytes, &e);
    if (rc) return rc;
    slot->ws = ws; slot->ws_bytes = ws_bytes; slot->n = n; slot->cfg = *cfg; slot->e = e;
  }
  slot->stamp = ++g_stamp;
  sphb200_engine* e = slot->e;
  if (ordered) {
    rc = upload_impl(e, in, (int)n, in_order, 0, stream, false, hit);
    if (!rc) rc = sphb200_engine_step(e, dt, 1, flags, stream);
    // every entry comes back from the frame: a step that sorted has moved the rows
    if (!rc) rc = download_impl(e, out, (int)n, out_order, 0, stream);
  } else {
    rc = hit ? sphb200_engine_refresh(e, in, 0, stream) : sphb200_engine_upload(e, in, 0, stream);
    if (!rc) rc = sphb200_engine_step(e, dt, 1, flags, stream);
  }
  if (!rc && !ordered) {
    // entries the step changes come back through the permutation (k_unpack scatters by particle
    // id); the others are the caller's own values: straight copies in -> out
    bool rest_some = false;
    written_mask(e, &rest_some);
    sphb200_state o = *out;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nv = (size_t)n * e->dim * 4, ns = (size_t)n * 4;
    auto pass = [&](float*& dst, const float* src, size_t bytes) {
      if (dst && src && dst != src)
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = SPHB200_ECUDA;
      if (src) dst = nullptr;  // not part of the scatter
    };
    const bool evol = cfg->flags & SPHB200_F_RHO_EVOL, heat = cfg->flags & SPHB200_F_HEAT;
    pass(o.mass, in->mass, ns);
    pass(o.eta, in->eta, ns);
    if (!heat) {
      pass(o.T, in->T, ns);
      pass(o.dTdt, in->dTdt, ns);
    }
    pass(o.kappa, in->kappa, ns);
    pass(o.Cp, in->Cp, ns);
    if (!evol) pass(o.drhodt, in->drhodt, ns);
    if (!(e->has_nw && e->wl_n > 0)) pass(o.nw, in->nw, nv);
    if (o.tag && in->tag) {
      if (o.tag != in->tag &&
          cudaMemcpyAsync(o.tag, in->tag, ns, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        rc = SPHB200_ECUDA;
      o.tag = nullptr;
    }
    if (!rc) rc = sphb200_engine_download(e, &o, 0, stream);
  }
  if (!rc && err)
    rc = cudaMemcpyAsync(err, e->err, 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) == cudaSuccess
             ? SPHB200_OK : SPHB200_ECUDA;
  if (!rc) rc = cudaMemsetAsync(e->err, 0, 4, (cudaStream_t)stream) == cudaSuccess ? SPHB200_OK : SPHB200_ECUDA;
  if (rc) {  // never keep an engine in an unknown state
    sphb200_engine_destroy(e);
    slot->e = nullptr;
  }
  return rc;
}

void sphb200_workspace_release(void* ws) {
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  for (CachedEngine& c : g_cache)
    if (c.e && (ws == nullptr || c.ws == ws)) {
      sphb200_engine_destroy(c.e);
      c.e = nullptr;
    }
}

int sphb200_forward(const sphb200_config* cfg, int64_t n, const sphb200_state* in,
                    sphb200_state* out, uint32_t* err, void* ws, size_t ws_bytes, void* stream) {
  return stateless_step(cfg, n, 0.0, 0u, in, out, err, ws, ws_bytes, stream);
}

int sphb200_advance(const sphb200_config* cfg, int64_t n, double dt, const sphb200_state* in,
                    sphb200_state* out, uint32_t* err, void* ws, size_t ws_bytes, void* stream) {
  return stateless_step(cfg, n, dt, SPHB200_STEP_INTEGRATE | SPHB200_STEP_BC, in, out, err, ws,
                        ws_bytes, stream);
}

int sphb200_advance_persistent(const sphb200_config* cfg, int64_t n, double dt,
                               const sphb200_state* in, sphb200_state* out, uint32_t* err, void* ws,
                               size_t ws_bytes, void* stream) {
  return stateless_step(cfg, n, dt, SPHB200_STEP_INTEGRATE | SPHB200_STEP_BC, in, out, err, ws,
                        ws_bytes, stream, true);
}

int sphb200_advance_ordered(const sphb200_config* cfg, int64_t n, double dt, const sphb200_state* in,
                            const int32_t* in_order, sphb200_state* out, int32_t* out_order,
                            uint32_t* err, void* ws, size_t ws_bytes, void* stream) {
  return stateless_step(cfg, n, dt, SPHB200_STEP_INTEGRATE | SPHB200_STEP_BC, in, out, err, ws,
                        ws_bytes, stream, true, true, in_order, out_order);
}

}  // extern "C"
