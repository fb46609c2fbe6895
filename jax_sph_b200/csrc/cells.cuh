// cells.cuh -- integrate + hash, cell table, stable reorder, pack / unpack.
//
// Replaces, per step: integrator.py:26-30 (kick, drift, periodic wrap), the cell
// hash + argsort + scatter of jax_md/partition.py:367-393 and the occupancy
// bookkeeping of :401-403.  The sort is a one-digit radix (counting) sort with
// radix = number of cells, made stable by an in-cell rank fix, so particle order
// inside a cell is the order of the previous frame -- deterministic sums.
//
// All passes are HBM-bound streaming kernels: one particle per thread, float4
// accesses, grid = ceil(N / 256).
#pragma once
#include "common.cuh"

namespace sphb200 {

// Slab decomposition: where a rank's particles live and what is in flight (slab.cuh).
// Particle slots are [base - n_halo_lo, base) lower halo | [base, base + n_own) own |
// [base + n_own, ...) upper halo; the counts are device-side (dn), so a step never
// synchronises with the host.  Single GPU: base = 0, dn = nullptr, n fixed.
enum { DN_OWN = 0, DN_IN = 1, DN_EMIG_LO = 2, DN_EMIG_HI = 3, DN_HALO_LO = 4, DN_HALO_HI = 5,
       DN_SEND_LO = 6, DN_SEND_HI = 7, DN_WORDS = 8 };

struct Slab {
  int base;      // first own slot
  int* dn;       // device counters (DN_*), nullptr on a single GPU
  int mig_cap;   // emigrant records per direction
  char* mig_lo;  // emigrant send buffers (records leaving towards the lower / upper neighbour)
  char* mig_hi;
};

// Emigrant / immigrant record buffer: [count, pad x3] then SoA arrays of `cap` entries each.
struct MigView {
  int* hdr;
  float4 *pt, *um, *vv, *st, *du, *nw, *ge;
  float2* kc;
  int* id;
};
__host__ __device__ inline size_t mig_bytes(int cap) { return 16 + (size_t)cap * (7 * 16 + 8 + 4); }
__host__ __device__ inline MigView mig_view(char* b, int cap) {
  MigView v;
  v.hdr = reinterpret_cast<int*>(b);
  char* q = b + 16;
  v.pt = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.um = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.vv = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.st = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.du = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.nw = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.ge = reinterpret_cast<float4*>(q); q += (size_t)cap * 16;
  v.kc = reinterpret_cast<float2*>(q); q += (size_t)cap * 8;
  v.id = reinterpret_cast<int*>(q);
  return v;
}

// K1: integrate (optionally), hash, histogram.  Reads pt, um, du, dv (64 B),
// writes key + arrival rank (8 B); cell counters live in L2.  Slab mode: a particle
// whose new cell lies outside the rank's own layers is an emigrant: its integrated
// record goes to the send buffer of that direction and its key becomes -1.
// `gate` (all kernels of the re-sort): a device word; the launch does nothing unless it is set
// (nullptr: always runs).  The kernels are grid-stride loops so that a gated-off launch costs a
// few hundred block exits, not one per 256 particles.
template <int DIM>
__global__ void __launch_bounds__(256) k_hash(int n, Grid g, Kick k, Slab sl, Frame f,
                                              int* __restrict__ key, int* __restrict__ rnk,
                                              int* __restrict__ count, unsigned* __restrict__ err,
                                              const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int n_int = sl.dn ? sl.dn[DN_OWN] : n;           // integrated sources (own)
  const int n_src = sl.dn ? n_int + sl.dn[DN_IN] : n;    // + immigrants (already integrated)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_src; t += gridDim.x * blockDim.x) {
  const int p = sl.base + t;
  float4 a = f.pt[p];
  float r[3] = {a.x, a.y, a.z};
  float u[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
  const bool integ = k.on && t < n_int;
  if (integ) {
    float4 b = f.um[p];
    u[0] = b.x; u[1] = b.y; u[2] = b.z;
    integrate_one<DIM>(k, g, r, u, v, f.du[p], f.dv[p]);
  }
  bool finite = isfinite(r[0]) && isfinite(r[1]) && (DIM == 2 || isfinite(r[2]));
  if (!finite) {
    atomicOr(err, SPHB200_ERR_NONFINITE);
    r[0] = r[1] = r[2] = 0.0f;
  }
  bool inside = r[0] >= 0.f && r[0] <= g.box[0] && r[1] >= 0.f && r[1] <= g.box[1] &&
                (DIM == 2 || (r[2] >= 0.f && r[2] <= g.box[2]));
  if (!inside) atomicOr(err, SPHB200_ERR_OUTSIDE_BOX);
  int c[3];
  int cell = cell_of<DIM>(g, r, c);
  if (sl.dn) {
    const int ax = DIM - 1;
    if (cell < 0 || c[ax] < g.own_lo[ax] || c[ax] >= g.own_hi[ax]) {
      // left the slab: one step moves a particle by a fraction of a cell, so it can only be
      // in the halo layer next to the own range
      const bool down = cell >= 0 && c[ax] < g.own_lo[ax];
      const bool up = cell >= 0 && c[ax] >= g.own_hi[ax];
      key[p] = -1;
      if (!(down || up) || t >= n_int) {
        atomicOr(err, SPHB200_ERR_SLAB_MIGRATION);
        continue;
      }
      const int slot = atomicAdd(&sl.dn[down ? DN_EMIG_LO : DN_EMIG_HI], 1);
      if (slot >= sl.mig_cap) {
        atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
        continue;
      }
      MigView m = mig_view(down ? sl.mig_lo : sl.mig_hi, sl.mig_cap);
      const float4 um = f.um[p], vv = f.vv[p];
      m.pt[slot] = make_float4(r[0], r[1], r[2], a.w);
      m.um[slot] = integ ? make_float4(u[0], u[1], u[2], um.w) : um;
      m.vv[slot] = integ ? make_float4(v[0], v[1], v[2], vv.w) : vv;
      m.st[slot] = f.st[p];
      m.du[slot] = make_float4(0.f, 0.f, 0.f, f.du[p].w);
      m.id[slot] = f.id[p];
      if (f.kc) m.kc[slot] = f.kc[p];
      if (f.nw) m.nw[slot] = f.nw[p];
      if (f.ge) m.ge[slot] = f.ge[p];
      continue;
    }
  }
  key[p] = cell;
  rnk[p] = atomicAdd(&count[cell], 1);
  }
}

// K2: exclusive scan of the cell histogram -> cell_start[0..C].  Three small
// launches (C <= a few million): block sums, scan of block sums, final scan.
constexpr int SCAN_TPB = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive block scan of one value per thread; returns inclusive value, total in *total
__device__ __forceinline__ int block_incl_scan(int v, int* sh /* >= 32 ints */, int* total) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = warp_incl_scan(v, lane);
  if (lane == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = lane < nw ? sh[lane] : 0;
    s = warp_incl_scan(s, lane);
    sh[lane] = s;
  }
  __syncthreads();
  int off = w > 0 ? sh[w - 1] : 0;
  *total = sh[nw - 1];
  __syncthreads();
  return inc + off;
}

__global__ void __launch_bounds__(SCAN_TPB) k_scan_partial(int c, const int* __restrict__ count,
                                                           int* __restrict__ bsum,
                                                           const int* __restrict__ gate = nullptr) {
  if (gate != nullptr && *gate == 0) return;
  __shared__ int sh[32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) s += (base + i < c) ? count[base + i] : 0;
  int tot;
  block_incl_scan(s, sh, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: exclusive scan of bsum[0..nb) in place (nb arbitrary)
__global__ void __launch_bounds__(1024) k_scan_bsum(int nb, int* __restrict__ bsum,
                                                    const int* __restrict__ gate = nullptr) {
  if (gate != nullptr && *gate == 0) return;
  __shared__ int sh[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    int tot;
    int inc = block_incl_scan(v, sh, &tot);
    int carry = carry_s;
    if (i < nb) bsum[i] = carry + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

// start[i] = off + exclusive prefix for i in [0, c]; start[c] is the grand total.
__global__ void __launch_bounds__(SCAN_TPB) k_scan_final(int c, int off, int* __restrict__ count,
                                                         const int* __restrict__ bsum,
                                                         int* __restrict__ start,
                                                         int* __restrict__ maxocc,
                                                         const int* __restrict__ gate = nullptr) {
  if (gate != nullptr && *gate == 0) return;
  __shared__ int sh[32];
  int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0, mx = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < c) ? count[base + i] : 0;
    s += v[i];
    mx = max(mx, v[i]);
  }
  int tot;
  int inc = block_incl_scan(s, sh, &tot);
  int run = off + bsum[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i <= c) start[base + i] = run;
    if (base + i < c) count[base + i] = 0;  // ready for the next step's histogram
    run += v[i];
  }
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 16));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 8));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 4));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 2));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 1));
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(maxocc, mx);
}

// K3: slot (arrival order) -> source index.
__global__ void __launch_bounds__(256) k_scatter_src(int n, Slab sl, const int* __restrict__ key,
                                                     const int* __restrict__ rnk,
                                                     const int* __restrict__ start,
                                                     int* __restrict__ src,
                                                     const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int bound = sl.dn ? sl.dn[DN_OWN] + sl.dn[DN_IN] : n;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < bound; t += gridDim.x * blockDim.x) {
    const int p = sl.base + t;
    const int k = key[p];
    if (k >= 0) src[start[k] + rnk[p]] = p;  // emigrants (key -1) drop out here
  }
}

// K4: stable in-cell rank + integrate + gather the whole frame into the new
// (cell-sorted) frame.  One thread per arrival slot.
struct ReorderOpt {
  int heat, has_nw, has_ge;
  int keep_acc;  // carry dudt / dvdt along (a sort BETWEEN two steps: the next kick needs them)
};

template <int DIM>
__global__ void __launch_bounds__(256) k_reorder(int n, Grid g, Kick k, Slab sl, ReorderOpt o,
                                                 Frame a, Frame b, const int* __restrict__ key,
                                                 const int* __restrict__ start,
                                                 const int* __restrict__ src,
                                                 const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  // slab mode: the own count after migration is the scan total (halo cells are still empty)
  const int bound = sl.dn ? start[g.ncells] - sl.base : n;
  const int kick_on = k.on;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < bound; t += gridDim.x * blockDim.x) {
  const int s = sl.base + t;
  int p = src[s];
  k.on = kick_on;
  if (sl.dn && p - sl.base >= sl.dn[DN_OWN]) k.on = 0;  // immigrants arrive integrated
  int cell = key[p];
  int lo = start[cell], hi = start[cell + 1];
  int rank = 0;
  for (int q = lo; q < hi; ++q) rank += (__ldg(&src[q]) < p);
  int f = lo + rank;

  float4 pt = a.pt[p], um = a.um[p], vv = a.vv[p], du = a.du[p];
  float r[3] = {pt.x, pt.y, pt.z}, u[3] = {um.x, um.y, um.z}, v[3] = {vv.x, vv.y, vv.z};
  if (k.on) integrate_one<DIM>(k, g, r, u, v, du, a.dv[p]);
  b.pt[f] = make_float4(r[0], r[1], r[2], pt.w);
  b.um[f] = make_float4(u[0], u[1], u[2], um.w);
  b.vv[f] = make_float4(v[0], v[1], v[2], vv.w);
  b.st[f] = a.st[p];
  b.du[f] = o.keep_acc ? du : make_float4(0.f, 0.f, 0.f, du.w);  // drhodt passes through when not evolved
  if (o.keep_acc) b.dv[f] = a.dv[p];
  b.id[f] = a.id[p];
  if (o.heat) b.kc[f] = a.kc[p];
  if (o.has_nw) b.nw[f] = a.nw[p];
  if (o.has_ge) b.ge[f] = a.ge[p];
  }
}

// ---------------------------------------------------------------------------
// Resident single-GPU engines keep the particles in place between two searches (sweep.cuh):
//
// k_drift: kick + drift + wrap of integrator.py:26-30 / space.py:207-209 IN PLACE, and the
// decision whether THIS step re-sorts and searches: a particle further than `limit` (half the
// skin of the neighbour lists) from where it was at the last sort (`rb`, minimum image) raises
// *flag_cur.  Sweeps only ever see positions within the limit, so a pair that is inside the
// cutoff now was inside cutoff + skin at the last search and is in the skin list.  The criterion
// is a function of the current positions alone -- a caller may hand in ANY state of the same
// particles (the stateless entry points do): if it is not close to the sorted one, the step sorts.
// The flag words alternate between steps: this launch also clears the next step's word.
template <int DIM>
// `maybe` (engines with the relative criterion, k_drift_box below): exceeding `limit` only asks
// for the relative test (word maybe[0], maybe[1] is cleared for the next step), exceeding
// `guard2` raises the flag at once.
__global__ void __launch_bounds__(256) k_drift(int n, Grid g, Kick k, Slab sl, Frame f,
                                               const float4* __restrict__ rb, int* flag_cur,
                                               int* flag_next, int force, float limit2,
                                               unsigned* __restrict__ err, int* maybe_cur = nullptr,
                                               int* maybe_next = nullptr, float guard2 = 0.f) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *flag_next = 0;
    if (maybe_next) *maybe_next = 0;
    if (force) atomicOr(flag_cur, 1);
  }
  bool over = false, over_guard = false;
  const int n_own = sl.dn ? sl.dn[DN_OWN] : n;  // slab engines: own particles only
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_own; t += gridDim.x * blockDim.x) {
    const int p = sl.base + t;
    const float4 a = f.pt[p];
    float r[3] = {a.x, a.y, a.z};
    if (k.on) {
      const float4 b = f.um[p], w = f.vv[p];
      float u[3] = {b.x, b.y, b.z}, v[3] = {w.x, w.y, w.z};
      integrate_one<DIM>(k, g, r, u, v, f.du[p], f.dv[p]);
      f.pt[p] = make_float4(r[0], r[1], r[2], a.w);
      f.um[p] = make_float4(u[0], u[1], u[2], b.w);
      f.vv[p] = make_float4(v[0], v[1], v[2], w.w);
    }
    const bool finite = isfinite(r[0]) && isfinite(r[1]) && (DIM == 2 || isfinite(r[2]));
    if (!finite) atomicOr(err, SPHB200_ERR_NONFINITE);
    const bool inside = r[0] >= 0.f && r[0] <= g.box[0] && r[1] >= 0.f && r[1] <= g.box[1] &&
                        (DIM == 2 || (r[2] >= 0.f && r[2] <= g.box[2]));
    if (finite && !inside) atomicOr(err, SPHB200_ERR_OUTSIDE_BOX);
    if (!force) {  // (a forced sort needs no test: rb may not exist yet)
      const float4 q = rb[p];
      float d[3] = {disp1(r[0], q.x, g.half[0], g.box[0]), disp1(r[1], q.y, g.half[1], g.box[1]),
                    DIM == 3 ? disp1(r[2], q.z, g.half[2], g.box[2]) : 0.f};
      const float d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      over = over || !(d2 <= limit2);
      over_guard = over_guard || !(d2 <= guard2);
    }
  }
  if (maybe_cur == nullptr) {
    if (__syncthreads_or(over) && threadIdx.x == 0) atomicOr(flag_cur, 1);
  } else {
    const int any = __syncthreads_or(over), anyg = __syncthreads_or(over_guard);
    if (threadIdx.x == 0) {
      if (anyg) atomicOr(flag_cur, 1);
      else if (any) atomicOr(maybe_cur, 1);
    }
  }
}

// ---------------------------------------------------------------------------
// The re-sort criterion by RELATIVE drift (single-GPU duo engines).  A pair inside the cutoff now
// was inside cutoff + skin at the last search as long as the two particles have moved less than
// `skin` RELATIVE to each other since then -- and neighbours move almost together: in a smooth
// flow the relative drift is the strain rate times the pair distance, an order of magnitude below
// the particles' own drift.  The bound is taken per block of S x S x S cells of the frozen table
// (slot ranges; a block edge is at least one cutoff + skin):
//   k_drift_box   the component-wise box [min, max] of the minimum-image drift r - rb of a block's
//                 particles;
//   k_drift_join  joins, axis after axis (min / max are separable), the boxes of every block that
//                 holds a cell within 2 S cells of the block's own cells, and -- last axis --
//                 raises the step's flag when the diagonal of the joined box exceeds the skin.
// Why this is sufficient.  Take any two particles.  If their sort-time cells are within 2 S cells
// of each other's BLOCK on every axis, both lie in one window: their relative drift is below the
// skin, so a pair inside the cutoff now was inside cutoff + skin at the search and is in the skin
// list.  Otherwise their sort-time cells are 2 S cells = two (cutoff + skin) apart along some axis,
// and with every particle's own drift below `guard` = (2 (cutoff + skin) - cutoff) / 2 (k_drift
// raises the flag at once beyond it) they are still further apart than the cutoff.  Either bound
// -- this one, or "no particle further than half the skin from where it was sorted" -- suffices,
// so the boxes are made only on the steps where the second one fails (k_drift's `maybe` word).
// The same guard keeps the "interior" shortcut of the sweeps valid: their tiles keep S more cells
// away from the periodic seam in these engines (Grid::imargin), so a particle that crossed the
// seam stays more than a cutoff away from every own particle of an interior tile.
struct DriftBlocks {
  int nb[3];       // blocks per axis
  float4* bmin;    // [3][blocks]: the blocks' own boxes, then two buffers of the axis passes
  float4* bmax;
};

template <int DIM>
__global__ void __launch_bounds__(128) k_drift_box(Grid g, DriftBlocks db, const int* __restrict__ start,
                                                   const float4* __restrict__ pt,
                                                   const float4* __restrict__ rb,
                                                   const int* __restrict__ flag_cur,
                                                   const int* __restrict__ maybe_cur) {
  // the step sorts anyway, or no particle is further than half the skin from where it was sorted
  // (the absolute criterion already proves the lists complete)
  if (*flag_cur != 0 || *maybe_cur == 0) return;
  const int nblocks = db.nb[0] * db.nb[1] * db.nb[2];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const int bx = b % db.nb[0], by = (b / db.nb[0]) % db.nb[1], bz = b / (db.nb[0] * db.nb[1]);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  const int x0 = bx * g.S[0], x1 = min(x0 + g.S[0], g.n[0]);
  const int z0 = DIM == 3 ? bz * g.S[2] : 0, z1 = DIM == 3 ? min((bz + 1) * g.S[2], g.n[2]) : 1;
  for (int cz = z0; cz < z1; ++cz) {
    for (int cy = by * g.S[1]; cy < min((by + 1) * g.S[1], g.n[1]); ++cy) {
      const int row = (cz * g.n[1] + cy) * g.n[0];
      const int s0 = start[row + x0], s1 = start[row + x1];
      for (int p = s0; p < s1; ++p) {
        const float4 a = pt[p], q = rb[p];
        const float d[3] = {disp1(a.x, q.x, g.half[0], g.box[0]), disp1(a.y, q.y, g.half[1], g.box[1]),
                            DIM == 3 ? disp1(a.z, q.z, g.half[2], g.box[2]) : 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          lo[k] = fminf(lo[k], d[k]);
          hi[k] = fmaxf(hi[k], d[k]);
        }
      }
    }
  }
  db.bmin[b] = make_float4(lo[0], lo[1], lo[2], 0.f);
  db.bmax[b] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

// One axis of the window join: dst[b] = join of src over the blocks that hold a cell of
// [S b - 2 S, S b + 3 S) along AXIS (periodic in CELLS: the last block of an axis may be a partial
// one).  LAST: nothing is stored, the diagonal of the result is tested against the limit.
template <int AXIS, bool LAST>
__global__ void __launch_bounds__(128) k_drift_join(Grid g, DriftBlocks db, int src, int dst, float limit2,
                                                    int* flag_cur, const int* __restrict__ maybe_cur) {
  if (*flag_cur != 0 || *maybe_cur == 0) return;
  const int nblocks = db.nb[0] * db.nb[1] * db.nb[2];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  bool over = false;
  if (b < nblocks) {
    int bc[3] = {b % db.nb[0], (b / db.nb[0]) % db.nb[1], b / (db.nb[0] * db.nb[1])};
    const int S = g.S[AXIS], n = g.n[AXIS], own = bc[AXIS];
    const float4* smin = db.bmin + (size_t)src * nblocks;
    const float4* smax = db.bmax + (size_t)src * nblocks;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int last = -1;
    for (int off = -2 * S; off < 3 * S; ++off) {
      int cc = S * own + off;
      cc = cc < 0 ? cc + n : (cc >= n ? cc - n : cc);
      const int nbk = cc / S;
      if (nbk == last) continue;  // (joining a block twice after a wrap is harmless)
      last = nbk;
      bc[AXIS] = nbk;
      const int q = (bc[2] * db.nb[1] + bc[1]) * db.nb[0] + bc[0];
      const float4 mn = smin[q], mx = smax[q];
      lo[0] = fminf(lo[0], mn.x); lo[1] = fminf(lo[1], mn.y); lo[2] = fminf(lo[2], mn.z);
      hi[0] = fmaxf(hi[0], mx.x); hi[1] = fmaxf(hi[1], mx.y); hi[2] = fmaxf(hi[2], mx.z);
    }
    if (LAST) {
      // (an empty window leaves lo = +inf, hi = -inf: spread -inf, clamped to 0)
      float d2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float w = fmaxf(hi[k] - lo[k], 0.f);
        d2 += w * w;
      }
      over = !(d2 <= limit2);
    } else {
      db.bmin[(size_t)dst * nblocks + b] = make_float4(lo[0], lo[1], lo[2], 0.f);
      db.bmax[(size_t)dst * nblocks + b] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
  }
  if (LAST && __syncthreads_or(over) && threadIdx.x == 0) atomicOr(flag_cur, 1);
}

// forward-only steps (no integration): just the flag protocol of k_drift
__global__ void k_gate(int* flag_cur, int* flag_next, int force) {
  *flag_next = 0;
  if (force) *flag_cur = 1;
}

// Slab engines agree on the re-sort decision: every rank publishes its flag in a message word,
// the transport max-reduces the word over the ranks, every rank takes the result.
__global__ void k_flag_out(const int* flag, int* word) { *word = *flag; }
__global__ void k_flag_in(const int* word, int* flag) { *flag = *word != 0 ? 1 : 0; }
// The same agreement, and the ring exchange, without a collective (slab engines whose ranks map
// each other's memory, include/sphb200.h: sphb200_slab_set_agree).  Every rank owns a control
// block of int32 words, mapped into all ranks:
//   [0, 2n)    agreement words [step parity][source rank]
//   [2n, 3n)   agreement arrivals [source rank]: the step number of the last word stored
//   [3n]       exchange signal from the rank above (its downward message is complete)
//   [3n + 1]   exchange signal from the rank below
// Signals are sequence numbers that only grow: a waiter spins until the word reaches the number it
// expects (system-scope acquire), a producer stores it after a system-scope fence behind its
// data.  A wait that lasts longer than `timeout_ns` gives up and raises SPHB200_ERR_SLAB_TIMEOUT.
struct AgreePtrs {
  int* p[16];
  int n;
};
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool wait_reach(const int* p, int want, unsigned long long timeout_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(p) - want < 0) {  // (difference: wrap-safe)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) return false;
    __nanosleep(200);
  }
  return true;
}
__global__ void k_flag_bcast(const int* flag, AgreePtrs a, int rank, int off, int seq) {
  if ((int)threadIdx.x < a.n) {
    int* blk = a.p[threadIdx.x];
    blk[off + rank] = *flag;
    __threadfence_system();
    st_release_sys(blk + 2 * a.n + rank, seq);
  }
}
__global__ void k_flag_in_max(const int* blk, int n, int off, int seq, int* flag,
                              unsigned* __restrict__ err, unsigned long long timeout_ns) {
  int m = 0;
  if ((int)threadIdx.x < n) {
    if (!wait_reach(blk + 2 * n + threadIdx.x, seq, timeout_ns)) atomicOr(err, SPHB200_ERR_SLAB_TIMEOUT);
    m = blk[off + threadIdx.x] != 0 ? 1 : 0;
  }
  m = __any_sync(0xffffffffu, m != 0) ? 1 : 0;
  if (threadIdx.x == 0) *flag = m;
}
// one exchange of the direct ring: my two messages are complete (the pack kernels ran earlier on
// this stream) -> tell both neighbours, then wait for theirs
__global__ void k_ring_signal(AgreePtrs a, int lo, int hi, int rank, int seq,
                              unsigned* __restrict__ err, unsigned long long timeout_ns) {
  const int n = a.n;
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(a.p[lo] + 3 * n, seq);      // the rank below: "from above"
  } else if (threadIdx.x == 1) {
    __threadfence_system();
    st_release_sys(a.p[hi] + 3 * n + 1, seq);  // the rank above: "from below"
  } else if (threadIdx.x == 2 || threadIdx.x == 3) {
    const int* w = a.p[rank] + 3 * n + (threadIdx.x - 2);
    if (!wait_reach(w, seq, timeout_ns)) atomicOr(err, SPHB200_ERR_SLAB_TIMEOUT);
  }
}

// After a re-sort (k_reorder: frame a -> frame b) the sorted particles go back to frame a, so
// that the host always launches on the same frame whether or not the device decided to re-sort.
__global__ void __launch_bounds__(256) k_copyback(int n, Slab sl, int ncells, ReorderOpt o,
                                                  Frame a, Frame b, float4* __restrict__ rb,
                                                  const int* __restrict__ start,
                                                  const int* __restrict__ gate, int* nsorts) {
  if (gate != nullptr && *gate == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0 && nsorts) atomicAdd(nsorts, 1);
  const int bound = sl.dn ? start[ncells] - sl.base : n;  // as k_reorder: the new own count
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < bound; t += gridDim.x * blockDim.x) {
    const int p = sl.base + t;
    a.pt[p] = b.pt[p];
    a.um[p] = b.um[p];
    a.vv[p] = b.vv[p];
    a.st[p] = b.st[p];
    a.du[p] = b.du[p];
    if (o.keep_acc) a.dv[p] = b.dv[p];
    a.id[p] = b.id[p];
    if (o.heat) a.kc[p] = b.kc[p];
    if (o.has_nw) a.nw[p] = b.nw[p];
    if (o.has_ge) a.ge[p] = b.ge[p];
    rb[p] = b.pt[p];  // where the particle was when the cells and lists were made
  }
}

// ---------------------------------------------------------------------------
// pack: reference layout -> frame (identity order); unpack: frame -> reference
// layout in ORIGINAL order (scatter through id).
struct StatePtrs {
  const float *r, *u, *v, *dudt, *dvdt, *nw, *rho, *p, *drhodt, *mass, *eta, *T, *dTdt, *kappa,
      *Cp, *g_ext;
  const int* tag;
};
struct StateOut {
  float *r, *u, *v, *dudt, *dvdt, *nw, *rho, *p, *drhodt, *mass, *eta, *T, *dTdt, *kappa, *Cp;
  int* tag;
  int* ids;  // slab mode: rows are written in local order and ids[row] = global particle index
};

template <int DIM>
__device__ __forceinline__ float4 load_vec(const float* a, int p, float w) {
  if (a == nullptr) return make_float4(0.f, 0.f, 0.f, w);
  if (DIM == 2) {
    float2 t = reinterpret_cast<const float2*>(a)[p];
    return make_float4(t.x, t.y, 0.f, w);
  }
  return make_float4(a[3 * p], a[3 * p + 1], a[3 * p + 2], w);
}

template <int DIM>
__device__ __forceinline__ void store_vec(float* a, int p, float4 q) {
  if (a == nullptr) return;
  if (DIM == 2) {
    reinterpret_cast<float2*>(a)[p] = make_float2(q.x, q.y);
  } else {
    a[3 * p] = q.x; a[3 * p + 1] = q.y; a[3 * p + 2] = q.z;
  }
}

// `ids` (slab mode): global particle indices of the uploaded rows; frame slots start at `base`.
// keep_order: the frame already holds THESE particles, cell-sorted (f.id[slot] = row of the
// particle): slot p takes row f.id[p] of the new state instead of row p -- the cell table and
// the neighbour lists stay valid as far as the positions allow (k_drift decides).
template <int DIM>
__global__ void __launch_bounds__(256) k_pack(int n, StatePtrs s, Frame fin, int base,
                                              const int* __restrict__ ids,
                                              int* __restrict__ wallcount, int keep_order) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  Frame f = fin;
  f.pt += base; f.um += base; f.vv += base; f.st += base; f.du += base; f.dv += base; f.id += base;
  if (f.kc) f.kc += base;
  if (f.nw) f.nw += base;
  if (f.ge) f.ge += base;
  const int q = keep_order ? f.id[p] : p;  // row of the state this slot takes
  int tag = s.tag ? s.tag[q] : 0;
  f.pt[p] = load_vec<DIM>(s.r, q, __int_as_float(tag));
  f.um[p] = load_vec<DIM>(s.u, q, s.mass ? s.mass[q] : 1.0f);
  f.vv[p] = load_vec<DIM>(s.v, q, s.eta ? s.eta[q] : 0.0f);
  f.st[p] = make_float4(s.rho ? s.rho[q] : 1.0f, s.p ? s.p[q] : 0.0f, s.T ? s.T[q] : 1.0f,
                        s.dTdt ? s.dTdt[q] : 0.0f);
  f.du[p] = load_vec<DIM>(s.dudt, q, s.drhodt ? s.drhodt[q] : 0.0f);
  f.dv[p] = load_vec<DIM>(s.dvdt, q, 0.0f);
  if (f.kc) f.kc[p] = make_float2(s.kappa ? s.kappa[q] : 0.0f, s.Cp ? s.Cp[q] : 0.0f);
  if (f.nw) f.nw[p] = load_vec<DIM>(s.nw, q, 0.0f);
  if (f.ge) f.ge[p] = load_vec<DIM>(s.g_ext, q, 0.0f);
  if (!keep_order) f.id[p] = ids ? ids[p] : p;
  if (is_wall_tag(tag)) atomicAdd(wallcount, 1);
}

template <int DIM>
__global__ void __launch_bounds__(256) k_unpack(int n, Slab sl, Frame f, StateOut o) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (sl.dn ? sl.dn[DN_OWN] : n)) return;
  const int s = sl.base + t;
  int p = f.id[s];
  if (o.ids) {
    o.ids[t] = p;
    p = t;
  }
  // a split download (advance_host) asks for a few entries only: load just their quads
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 pt = (o.r || o.tag) ? f.pt[s] : z4, um = (o.u || o.mass) ? f.um[s] : z4;
  const float4 vv = (o.v || o.eta) ? f.vv[s] : z4;
  const float4 st = (o.rho || o.p || o.T || o.dTdt) ? f.st[s] : z4;
  const float4 du = (o.dudt || o.drhodt) ? f.du[s] : z4, dv = o.dvdt ? f.dv[s] : z4;
  store_vec<DIM>(o.r, p, pt);
  store_vec<DIM>(o.u, p, um);
  store_vec<DIM>(o.v, p, vv);
  store_vec<DIM>(o.dudt, p, du);
  store_vec<DIM>(o.dvdt, p, dv);
  if (o.tag) o.tag[p] = __float_as_int(pt.w);
  if (o.mass) o.mass[p] = um.w;
  if (o.eta) o.eta[p] = vv.w;
  if (o.rho) o.rho[p] = st.x;
  if (o.p) o.p[p] = st.y;
  if (o.T) o.T[p] = st.z;
  if (o.dTdt) o.dTdt[p] = st.w;
  if (o.drhodt) o.drhodt[p] = du.w;
  if (f.kc) {
    float2 kc = f.kc[s];
    if (o.kappa) o.kappa[p] = kc.x;
    if (o.Cp) o.Cp[p] = kc.y;
  }
  if (o.nw && f.nw) store_vec<DIM>(o.nw, p, f.nw[s]);
}

}  // namespace sphb200
