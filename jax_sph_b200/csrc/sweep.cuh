// sweep.cuh -- the pairwise sweeps over cell-sorted particles.
//
// Replaces the reference's candidate gather + prune + E-sized gathers +
// segment_sum chains (jax_md/partition.py:832-909, solver.py:722-731 and every
// ops.segment_sum in solver.py:750-928) by one kernel template in four roles.
//
// One thread block owns a tile of T[0] x T[1] x T[2] cells and stages the particles of the
// tile's stencil (tile + S cells each side) into shared memory with coalesced float4 loads.
//
//   LIST_BUILD    (PhysNone, on the steps that re-sort the particles) the SEARCH: every thread
//                 walks its own (2S+1)^d window of the staged cells with a cheap squared-distance
//                 test against (cutoff + skin)^2 and writes the survivors' staged indices to its
//                 row of the SKIN LIST in HBM.  No physics.
//   LIST_FILTER   (the first sweep of forward(): density) every thread steps through its
//                 skin-list row, re-derives the displacement with the reference's exact float32
//                 arithmetic (space.py:170-181), keeps the pairs with d^2 < cutoff^2
//                 (jax_md/partition.py:897 -- membership is decided every step, so the neighbour
//                 set is the reference's although the search is amortised over several steps),
//                 accumulates the physics and writes the survivors to the EXACT LIST of the step.
//   LIST_CONSUME  (renormalisation, wall, force) steps through the exact list: every lane on a
//                 real pair, no test.
//   LIST_NONE     search + exact test + physics in one kernel (the neighbour-list materialiser;
//                 also the fall-back path inside FILTER / CONSUME kernels for a tile whose lists
//                 do not exist: stencil larger than one staging group, or a row overflow).
// Between two searches the particles keep their slots and the cell table is frozen: a staged
// index means the same particle in every sweep of every step until the next search (engine.cu).
// The physics (what is accumulated per pair, what is written per particle) is a policy class P,
// see phys.cuh.
#pragma once
#include "common.cuh"

namespace sphb200 {

constexpr int MAX_RUNS = 16;     // T[1]*T[2] upper bound
constexpr int MAX_SOFF = 1023;   // staged (row, cell) entries upper bound
constexpr int SWEEP_CHUNK = 16;  // phase-1 candidates between list-room checks
constexpr int SWEEP_MAXT = 512;  // largest block size the kernels are compiled for
constexpr int LS = SWEEP_MAXT;   // row stride of the per-thread lists: a compile-time constant, so
                                 // that the append in the phase-1 loop is one immediate add

struct SweepDims {
  int cap;   // staged particles per block
  int lcap;  // list entries per thread (>= SWEEP_CHUNK)
  int sb;    // bytes staged per particle (16 per quad + scalar columns, see phys.cuh)
  const int* gate;  // device word; the launch does nothing unless *gate == gate_want (nullptr: always
  int gate_want;    // runs); gate_want < 0: unless *gate != 0
};

// Neighbour lists in HBM.  Entries are STAGED indices (uint16): every sweep stages a tile's
// stencil in the same order, so the index means the same particle everywhere.  A particle's row
// is lmax entries, read and written in 16-byte chunks of 8.
//   sl / scnt  skin list: pairs within cutoff + skin when the particles were last sorted
//              (LIST_BUILD writes, LIST_FILTER reads)
//   xl / xcnt  exact list of the current step: the subset with d^2 < cutoff^2 now
//              (LIST_FILTER writes, LIST_CONSUME reads)
// A tile whose stencil does not fit one staging group of every sweep, or that holds a particle
// with more than lmax skin neighbours, is marked not-ok and every sweep searches it on its own
// (the lists are an accelerator, never a correctness dependency).
struct NList {
  unsigned short* xl;    // [n][lmax], nullptr = lists off
  int* xcnt;             // [n]
  unsigned short* sl;    // [n][lmax]
  int* scnt;             // [n]
  unsigned char* ok;     // [tiles]
  int lmax;              // multiple of 8
  int min_cap;           // smallest staging capacity among the step's sweeps
  int* nbuilds;          // optional device counter, one per LIST_BUILD launch
  const unsigned char* skip_ok;  // [tiles] or nullptr: tiles marked ok here belong to the duo sweeps
                                 // (sweep2.cuh); this launch sweeps the others, searching on its own
};

enum { LIST_NONE = 0, LIST_BUILD = 1, LIST_CONSUME = 2, LIST_FILTER = 3 };

__host__ __device__ inline size_t sweep_smem_bytes(int sb, int cap, int lcap, int tpb) {
  (void)tpb;
  return (size_t)sb * cap + (size_t)lcap * LS * 2 + (MAX_SOFF + 1 + 2 * MAX_RUNS + 2) * 4;
}

// unwrapped (global) cell index u in [-n, 2n)  ->  periodic image count / wrapped index
__device__ __forceinline__ int wrap_count(int u, int n) { return u < 0 ? -1 : (u >= n ? 1 : 0); }
__device__ __forceinline__ int wrap_cell(int u, int n) { return u < 0 ? u + n : (u >= n ? u - n : u); }

// Reference displacement r_i - r_j (space.py:170-181) of one staged neighbour; INTERIOR: no
// periodic image inside the tile's stencil, the fold is two adds.  (A particle that drifted
// across the periodic seam since the last sort is more than a cutoff away from every own
// particle of an interior tile: the unfolded distance rejects it, as it must.)
template <int DIM, bool INTERIOR>
__device__ __forceinline__ void pair_disp(const Grid& g, const float (&ri)[3], const float4 pj,
                                          float (&dr)[3]) {
  if (INTERIOR) {
    dr[0] = disp1_nowrap(ri[0], pj.x, g.half[0]);
    dr[1] = disp1_nowrap(ri[1], pj.y, g.half[1]);
    dr[2] = (DIM == 3) ? disp1_nowrap(ri[2], pj.z, g.half[2]) : 0.0f;
  } else {
    dr[0] = disp1(ri[0], pj.x, g.half[0], g.box[0]);
    dr[1] = disp1(ri[1], pj.y, g.half[1], g.box[1]);
    dr[2] = (DIM == 3) ? disp1(ri[2], pj.z, g.half[2], g.box[2]) : 0.0f;
  }
}

// ---------------------------------------------------------------------------
// Physics policies that provide a packed pair body (pair2 on Acc2: two neighbours in the two
// halves of packed float32 registers, common.cuh) opt in with HAS_PAIR2; the others get two
// predicated scalar pair() calls.
template <class P, bool = P::HAS_PAIR2>
struct Pair2 {
  using Acc2 = typename P::Acc;
  __device__ static __forceinline__ void init(Acc2& a) { P::init(a); }
  __device__ static __forceinline__ void pair2(const Consts& c, const Extra& ex,
                                               const typename P::Own& o, Acc2& a, const float4* sq,
                                               int cap, int j0, int j1, float4 p0, float4 p1,
                                               const F2 (&dr)[3], F2 d2, bool v0, bool v1) {
    if (v0) {
      const float d[3] = {lo(dr[0]), lo(dr[1]), lo(dr[2])};
      P::pair(c, ex, o, a, sq, cap, j0, p0, d, lo(d2));
    }
    if (v1) {
      const float d[3] = {hi(dr[0]), hi(dr[1]), hi(dr[2])};
      P::pair(c, ex, o, a, sq, cap, j1, p1, d, hi(d2));
    }
  }
  __device__ static __forceinline__ void fold(const Acc2& a2, typename P::Acc& a) { a = a2; }
};
template <class P>
struct Pair2<P, true> {
  using Acc2 = typename P::Acc2;
  __device__ static __forceinline__ void init(Acc2& a) { P::init2(a); }
  __device__ static __forceinline__ void pair2(const Consts& c, const Extra& ex,
                                               const typename P::Own& o, Acc2& a, const float4* sq,
                                               int cap, int j0, int j1, float4 p0, float4 p1,
                                               const F2 (&dr)[3], F2 d2, bool v0, bool v1) {
    P::pair2(c, ex, o, a, sq, cap, j0, j1, p0, p1, dr, d2, v0, v1);
  }
  __device__ static __forceinline__ void fold(const Acc2& a2, typename P::Acc& a) { P::fold(a2, a); }
};

// List consumer: every thread steps through the row of its own particle, one 32-bit word = two
// entries per trip, evaluated in the two halves of packed float32 registers (FADD2 / FMUL2 /
// FFMA2): half the issue slots of the displacement, distance and kernel arithmetic, and two
// independent LDS -> displacement -> rsqrt -> kernel -> pair-term chains in flight per thread.
// FILTER: the row is the skin list; the exact membership test of the reference (d^2 < cutoff^2
// on the float32 displacement of space.py:170-181) runs on every entry and the survivors are
// appended to the exact list of the step through a 64-bit shift register (one STG.64 per four
// survivors, branch-free).
// (Eight lanes per particle with ballot compaction were tried instead: 38 % fewer bank
// conflicts, but the per-particle prologue / reduction / epilogue and the exposed latency of the
// short rows cost more than they save: profiles/r02_sweeps_qw_tgv3d_128_raw.txt.)
template <int DIM, class P, bool INTERIOR, bool FILTER>
__device__ __forceinline__ void consume_list(const Grid& g, const Consts& c, const Extra& ex,
                                             const NList& nl, const float4* sq, int cap, int p,
                                             int nn, bool have, const float (&ri)[3],
                                             const typename P::Own& own, typename P::Acc& acc) {
  using P2 = Pair2<P>;
  const size_t roff = (size_t)p * nl.lmax;
  const uint4* lrow = reinterpret_cast<const uint4*>((FILTER ? nl.sl : nl.xl) + roff);
  unsigned long long* xp = reinterpret_cast<unsigned long long*>(nl.xl + roff);
  unsigned long long xw = 0ull;
  int m = 0;
  typename P2::Acc2 acc2;
  P2::init(acc2);
  uint4 cur = make_uint4(0u, 0u, 0u, 0u), nxt = cur;
  if (nn > 0) nxt = __ldg(lrow);
#pragma unroll 1
  for (int k = 0; k < nn; k += 2) {
    if ((k & 7) == 0) {
      cur = nxt;
      ++lrow;
      if (k + 8 < nn) nxt = __ldg(lrow);
    }
    const unsigned w = cur.x;
    cur.x = cur.y; cur.y = cur.z; cur.z = cur.w;
    bool v0 = true, v1 = k + 1 < nn;
    const int j0 = (int)(w & 0xffffu), j1 = v1 ? (int)(w >> 16) : 0;
    const float4 p0 = sq[j0], p1 = sq[j1];
    F2 dr[3];
    if (INTERIOR) {
      dr[0] = disp2_nowrap(ri[0], f2(p0.x, p1.x), g.half[0]);
      dr[1] = disp2_nowrap(ri[1], f2(p0.y, p1.y), g.half[1]);
      dr[2] = (DIM == 3) ? disp2_nowrap(ri[2], f2(p0.z, p1.z), g.half[2]) : f2(0.0f);
    } else {
      dr[0] = f2(disp1(ri[0], p0.x, g.half[0], g.box[0]), disp1(ri[0], p1.x, g.half[0], g.box[0]));
      dr[1] = f2(disp1(ri[1], p0.y, g.half[1], g.box[1]), disp1(ri[1], p1.y, g.half[1], g.box[1]));
      dr[2] = (DIM == 3) ? f2(disp1(ri[2], p0.z, g.half[2], g.box[2]),
                              disp1(ri[2], p1.z, g.half[2], g.box[2]))
                         : f2(0.0f);
    }
    const F2 d2 = sumsq2<DIM>(dr);
    if (FILTER) {
      v0 = lo(d2) < g.c2;
      v1 = v1 && hi(d2) < g.c2;
      // survivors -> the exact list: shift in, count, store every fourth (no divergent blocks)
      const unsigned long long x0 = (xw >> 16) | ((unsigned long long)j0 << 48);
      xw = v0 ? x0 : xw;
      m += v0 ? 1 : 0;
      if (v0 && (m & 3) == 0) *xp++ = xw;
      const unsigned long long x1 = (xw >> 16) | ((unsigned long long)j1 << 48);
      xw = v1 ? x1 : xw;
      m += v1 ? 1 : 0;
      if (v1 && (m & 3) == 0) *xp++ = xw;
    }
    P2::pair2(c, ex, own, acc2, sq, cap, j0, j1, p0, p1, dr, d2, v0, v1);
  }
  P2::fold(acc2, acc);
  if (FILTER && have) {
    if (m & 3) *xp = xw >> (16 * (4 - (m & 3)));  // the last, partial group of four
    nl.xcnt[p] = m;
  }
}

// The search: one thread's walk through its (2S+1)^d window of the staged cells, as
// 2 * W1 * W2 (row, x-segment) pieces (the x range of a row is split where the periodic image
// changes).  All lanes of a warp advance through the pieces together.
template <int DIM>
struct Walk {
  int w0[3], ka, kb, ks, it_step, nit;
  float xs0, xs1;
  int it, wy, wz, j, jb, row;
  bool inwin;
  float xs, ys, zs;

  // ci: the particle's cell at the time of the last sort, relative to the local grid
  __device__ __forceinline__ void init(const Grid& g, const int (&sa0)[3], const int (&ci)[3],
                                       const float (&ri)[3], bool have) {
#pragma unroll
    for (int a = 0; a < 3; ++a) w0[a] = (g.n[a] >= 2 * g.S[a] + 1) ? (ci[a] - g.S[a] - sa0[a]) : 0;
    const int kz = -sa0[0], kn = g.n[0] - sa0[0];  // staged x index of unwrapped cells 0 and n
    ka = w0[0];
    kb = w0[0] + g.W[0];
    ks = kb;
    if (ka < kz && kz < kb) ks = kz;
    else if (ka < kn && kn < kb) ks = kn;
    xs0 = ri[0] - (float)wrap_count(sa0[0] + ka, g.n[0]) * g.box[0];  // x is never the slab axis
    xs1 = ri[0] - (float)wrap_count(sa0[0] + ks, g.n[0]) * g.box[0];
    // the second x segment exists only where the window crosses the periodic seam: elsewhere
    // (all lanes of the warp agree) the row loop visits first segments only
    it_step = __any_sync(FULL_MASK, have && ks < kb) ? 1 : 2;
    nit = 2 * g.W[1] * g.W[2];
    it = -1; wy = -1; wz = 0; j = 0; jb = 0; row = 0;
    inwin = false;
    xs = ys = zs = 0.f;
  }

  // Phase 1: cheap reject against `thr`, survivors appended to the thread's shared-memory column
  // (cnt entries so far, lcap at most).  Returns true when the window is exhausted, false when
  // some lane's column cannot take the next chunk (the caller drains the columns and calls again).
  __device__ __forceinline__ bool run(const Grid& g, const float4* sq, unsigned short* list, int tid,
                                      int lcap, int& cnt, const int* soff, int base, int e_a,
                                      int e_b, int nxs, int slen1, const int (&sa0)[3],
                                      const float (&ri)[3], bool act, float thr) {
    for (;;) {
      const int rem = jb - j;
      if (!__any_sync(FULL_MASK, rem > 0)) {
        it = it < 0 ? 0 : it + it_step;
        if (it >= nit) return true;
        const int sgm = it & 1;
        if (sgm == 0) {
          if (++wy == g.W[1]) {
            wy = 0;
            ++wz;
          }
          const int ry = w0[1] + wy, rz = w0[2] + wz;
          row = rz * slen1 + ry;
          inwin = act;
          ys = ri[1] - (float)wrap_count(sa0[1] + ry + g.goff[1], g.ng[1]) * g.box[1];
          zs = ri[2] - (float)wrap_count(sa0[2] + rz + g.goff[2], g.ng[2]) * g.box[2];
        }
        const int kk0 = sgm == 0 ? ka : ks, kk1 = sgm == 0 ? ks : kb;
        j = jb = 0;
        if (inwin && kk0 < kk1) {
          const int ea = max(row * nxs + kk0, e_a), eb = min(row * nxs + kk1, e_b);
          if (ea < eb) {
            j = soff[ea] - base;
            jb = soff[eb] - base;
          }
        }
        xs = sgm == 0 ? xs0 : xs1;
        continue;
      }
      const int want = rem > 0 ? min(rem, SWEEP_CHUNK) : 0;
      if (__any_sync(FULL_MASK, want > lcap - cnt)) return false;  // drain first
      const int e = j + want;
      int lo = cnt * LS + tid;  // tid < LS: lo / LS is the entry count at any time
      // four candidates per trip, loaded before the first append: the appends go to the
      // same shared-memory window the candidates come from, so the compiler would
      // otherwise keep every load behind the previous store (one LDS latency per candidate)
      for (; j + 4 <= e; j += 4) {
        const float4 p0 = sq[j], p1 = sq[j + 1], p2 = sq[j + 2], p3 = sq[j + 3];
        float t0 = xs - p0.x, t1 = xs - p1.x, t2 = xs - p2.x, t3 = xs - p3.x;
        float d0 = t0 * t0, d1 = t1 * t1, d2 = t2 * t2, d3 = t3 * t3;
        t0 = ys - p0.y; t1 = ys - p1.y; t2 = ys - p2.y; t3 = ys - p3.y;
        d0 += t0 * t0; d1 += t1 * t1; d2 += t2 * t2; d3 += t3 * t3;
        if (DIM == 3) {
          t0 = zs - p0.z; t1 = zs - p1.z; t2 = zs - p2.z; t3 = zs - p3.z;
          d0 += t0 * t0; d1 += t1 * t1; d2 += t2 * t2; d3 += t3 * t3;
        }
        if (d0 < thr) { list[lo] = (unsigned short)j; lo += LS; }
        if (d1 < thr) { list[lo] = (unsigned short)(j + 1); lo += LS; }
        if (d2 < thr) { list[lo] = (unsigned short)(j + 2); lo += LS; }
        if (d3 < thr) { list[lo] = (unsigned short)(j + 3); lo += LS; }
      }
#pragma unroll 1
      for (; j < e; ++j) {
        const float4 pj = sq[j];
        const float dx = xs - pj.x, dy = ys - pj.y;
        float d2 = dx * dx + dy * dy;
        if (DIM == 3) {
          const float dz = zs - pj.z;
          d2 += dz * dz;
        }
        if (d2 < thr) {
          list[lo] = (unsigned short)j;
          lo += LS;
        }
      }
      cnt = lo / LS;
    }
  }
};

template <int DIM, class P, int LM = LIST_NONE>
__global__ void __launch_bounds__(SWEEP_MAXT, P::MINB)
    k_sweep(const Grid g, const Consts c, const Frame f, const int* __restrict__ cs,
            const SweepDims sd, const Extra ex, unsigned* __restrict__ err, const NList nl) {
  if (sd.gate != nullptr && (sd.gate_want >= 0 ? *sd.gate != sd.gate_want : *sd.gate == 0)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TPB = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = TPB >> 5;
  float4* sq = reinterpret_cast<float4*>(smem_raw);
  unsigned short* list = reinterpret_cast<unsigned short*>(smem_raw + (size_t)sd.sb * sd.cap);
  int* soff = reinterpret_cast<int*>(list + (size_t)sd.lcap * LS);
  int* own_off = soff + (MAX_SOFF + 1);
  int* own_start = own_off + (MAX_RUNS + 1);
  __shared__ int s_bad;
  if (LM == LIST_BUILD && blockIdx.x == 0 && tid == 0 && nl.nbuilds) atomicAdd(nl.nbuilds, 1);

  // a launch covers tiles [block0, block0 + ntl); gated launches use a small persistent grid
  for (int tq = blockIdx.x; tq < g.ntl; tq += gridDim.x) {
    if (nl.skip_ok != nullptr && nl.skip_ok[tq + g.block0] != 0) continue;
    if (tq != (int)blockIdx.x) __syncthreads();  // readers of the previous tile's tables are done

    // ---- tile geometry (uniform) -------------------------------------------
    int b = tq + g.block0;
    const int tile_id = b;
    const int tx = b % g.nt[0];
    b /= g.nt[0];
    const int ty = b % g.nt[1];
    const int tz = b / g.nt[1];
    int c0[3] = {g.own_lo[0] + tx * g.T[0], g.own_lo[1] + ty * g.T[1], g.own_lo[2] + tz * g.T[2]};
    int no[3], sa0[3], slen[3];
    bool interior = true;  // no periodic image inside this tile's stencil
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      no[a] = min(g.T[a], g.own_hi[a] - c0[a]);
      if (g.n[a] >= 2 * g.S[a] + 1) {
        sa0[a] = c0[a] - g.S[a];
        slen[a] = no[a] + 2 * g.S[a];
      } else {
        sa0[a] = 0;
        slen[a] = g.n[a];
      }
      const int im = g.imargin ? g.S[a] : 0;
      interior = interior && (a >= DIM || (sa0[a] + g.goff[a] - im >= 0 &&
                                           sa0[a] + g.goff[a] + slen[a] + im <= g.ng[a] &&
                                           g.n[a] >= 2 * g.S[a] + 2));
    }
    if (g.exact_all) interior = false;
    const int nxs = slen[0];
    const int nrows = slen[1] * slen[2];
    const int E = nrows * nxs;
    const int nruns = no[1] * no[2];

    if (tid < nruns) {
      int ry = tid % no[1], rz = tid / no[1];
      int cell = ((c0[2] + rz) * g.n[1] + (c0[1] + ry)) * g.n[0] + c0[0];
      int s = cs[cell];
      own_start[tid] = s;
      own_off[tid + 1] = cs[cell + no[0]] - s;
    }
    for (int e = tid; e < E; e += TPB) {
      int k = e % nxs, row = e / nxs;
      int ry = row % slen[1], rz = row / slen[1];
      int cell = (wrap_cell(sa0[2] + rz, g.n[2]) * g.n[1] + wrap_cell(sa0[1] + ry, g.n[1])) * g.n[0] +
                 wrap_cell(sa0[0] + k, g.n[0]);
      soff[e] = cs[cell + 1] - cs[cell];
    }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan of soff[0..E) in place, soff[E] = total
      int carry = 0;
      for (int base = 0; base < E; base += 32) {
        int i = base + lane;
        int v = i < E ? soff[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(FULL_MASK, inc, o);
          if (lane >= o) inc += t;
        }
        if (i < E) soff[i] = carry + inc - v;
        carry += __shfl_sync(FULL_MASK, inc, 31);
      }
      if (lane == 0) {
        soff[E] = carry;
        int acc = 0;
        own_off[0] = 0;
        for (int r = 0; r < nruns; ++r) {
          acc += own_off[r + 1];
          own_off[r + 1] = acc;
        }
      }
    }
    __syncthreads();
    const int tile_n = own_off[nruns];
    if (tile_n == 0) continue;
    const int total_staged = soff[E];
    const bool single_group = total_staged <= sd.cap;

    // list mode of this tile (uniform over the block, see NList)
    bool nl_use = false;
    if (LM == LIST_BUILD) {
      if (nl.sl == nullptr || total_staged > nl.min_cap) {  // no lists for this tile
        if (tid == 0 && nl.ok) nl.ok[tile_id] = 0;
        continue;
      }
      if (tid == 0) s_bad = 0;  // ordered before any other write by the staging barrier
    }
    if (LM == LIST_CONSUME || LM == LIST_FILTER) nl_use = nl.xl != nullptr && nl.ok[tile_id] != 0;

    // stage the (row, cell) entries [e_a, e_b) of the stencil; `base` = staged index of e_a
    auto stage_group = [&](int e_a, int e_b, int base) {
      const int r_first = e_a / nxs, r_last = (e_b - 1) / nxs;
      const int njobs = (r_last - r_first + 1) * 3;
      for (int job = warp; job < njobs; job += nwarps) {
        const int row = r_first + job / 3;
        const int seg = job % 3 - 1;
        const int klo = row == r_first ? e_a - row * nxs : 0;
        const int khi = row == r_last ? e_b - row * nxs : nxs;
        int k0 = max(klo, seg * g.n[0] - sa0[0]);
        int k1 = min(khi, (seg + 1) * g.n[0] - sa0[0]);
        if (k0 >= k1) continue;
        const int ry = row % slen[1], rz = row / slen[1];
        const int rowcell =
            (wrap_cell(sa0[2] + rz, g.n[2]) * g.n[1] + wrap_cell(sa0[1] + ry, g.n[1])) * g.n[0];
        const int gstart = cs[rowcell + (sa0[0] + k0 - seg * g.n[0])];
        const int dst = soff[row * nxs + k0] - base;
        const int len = soff[row * nxs + k1] - soff[row * nxs + k0];
        for (int m = lane; m < len; m += 32) P::stage(c, f, ex, gstart + m, sq, sd.cap, dst + m);
      }
    };

    bool staged = false;  // single-group tiles stage once for all rounds of own particles
    for (int ib = 0; ib < tile_n; ib += TPB) {
      // ---- own particle ------------------------------------------------------
      const int t = ib + tid;
      const bool have = t < tile_n;
      int p = 0, run = 0;
      if (have) {
        while (t >= own_off[run + 1]) ++run;
        p = own_start[run] + (t - own_off[run]);
      }
      typename P::Own own;
      typename P::Acc acc;
      float ri[3] = {0.f, 0.f, 0.f};
      bool act = false;
      if (have) {
        float4 q = f.pt[p];
        ri[0] = q.x; ri[1] = q.y; ri[2] = q.z;
        P::load_own(c, f, ex, p, q, own);
        act = P::active(c, own);
      }
      P::init(acc);
      // sparse policies (the wall sweep: wall particles only) skip tiles without work
      const bool any_act = P::SPARSE ? (__syncthreads_or(have && act) != 0) : true;

      // The particle's cell at the time of the last sort (the window of the search is laid
      // around it): its run gives (y, z), the row's cell table the x cell.  Only search paths
      // need it.
      int ci[3] = {0, 0, 0};
      if ((LM == LIST_BUILD || !nl_use) && have) {
        const int ry = run % no[1], rz = run / no[1];
        const int rowcell0 = ((c0[2] + rz) * g.n[1] + (c0[1] + ry)) * g.n[0] + c0[0];
        int kx = 0;
        while (kx + 1 < no[0] && __ldg(cs + rowcell0 + kx + 1) <= p) ++kx;
        ci[0] = c0[0] + kx; ci[1] = c0[1] + ry; ci[2] = c0[2] + rz;
      }

      if constexpr (LM == LIST_CONSUME || LM == LIST_FILTER) {
        if (nl_use) {
          // ---------------- list consumer (FILTER: exact test + exact list of the step) ------
          if (any_act) {
            if (!staged) {
              if (ib > 0) __syncthreads();
              stage_group(0, E, 0);
              __syncthreads();
              staged = true;
            }
            const int nn = (have && act) ? (LM == LIST_FILTER ? nl.scnt[p] : nl.xcnt[p]) : 0;
            if (interior)
              consume_list<DIM, P, true, LM == LIST_FILTER>(g, c, ex, nl, sq, sd.cap, p, nn,
                                                            have && act, ri, own, acc);
            else
              consume_list<DIM, P, false, LM == LIST_FILTER>(g, c, ex, nl, sq, sd.cap, p, nn,
                                                             have && act, ri, own, acc);
          }
          if (have) P::finish(c, f, ex, p, own, acc);
          continue;
        }
      }

      if (LM == LIST_BUILD) {
        // ---------------- the search: skin list of every own particle -> HBM ----------------
        if (!staged) {
          stage_group(0, E, 0);
          __syncthreads();
          staged = true;
        }
        Walk<DIM> wk;
        wk.init(g, sa0, ci, ri, have);
        int cnt = 0, gk = 0;  // entries waiting in the column / already in HBM
        unsigned short* col = list + tid;
        unsigned short* grow = nl.sl + (size_t)p * nl.lmax;
        for (;;) {
          const bool fin = wk.run(g, sq, list, tid, sd.lcap, cnt, soff, 0, 0, E, nxs, slen[1], sa0,
                                  ri, have, g.c2_hi);
          // full chunks of 8 -> one STG.128 each; the remainder waits at the head of the column
          int w = 0;
          for (; cnt - w >= 8; w += 8) {
            uint4 v;
            v.x = (unsigned)col[(w + 0) * LS] | ((unsigned)col[(w + 1) * LS] << 16);
            v.y = (unsigned)col[(w + 2) * LS] | ((unsigned)col[(w + 3) * LS] << 16);
            v.z = (unsigned)col[(w + 4) * LS] | ((unsigned)col[(w + 5) * LS] << 16);
            v.w = (unsigned)col[(w + 6) * LS] | ((unsigned)col[(w + 7) * LS] << 16);
            if (gk + 8 <= nl.lmax) *reinterpret_cast<uint4*>(grow + gk) = v;
            else s_bad = 1;
            gk += 8;
          }
          const int r = cnt - w;
          if (w > 0)
            for (int i = 0; i < r; ++i) col[i * LS] = col[(w + i) * LS];
          cnt = r;
          if (fin) break;
        }
        if (have) {
          if (cnt > 0) {  // tail chunk (its unused slots are never read: scnt says how many are real)
            unsigned e8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e8[i] = i < cnt ? (unsigned)col[i * LS] : 0u;
            if (gk + 8 <= nl.lmax)
              *reinterpret_cast<uint4*>(grow + gk) =
                  make_uint4(e8[0] | (e8[1] << 16), e8[2] | (e8[3] << 16), e8[4] | (e8[5] << 16),
                             e8[6] | (e8[7] << 16));
            else
              s_bad = 1;
          }
          nl.scnt[p] = gk + cnt;
        }
        continue;
      }

      // ---------------- search + exact test + physics (no lists for this tile) ----------------
      // Staging groups: maximal runs [e_a, e_b) of consecutive (row, cell) entries that fit the
      // staging buffer (one group unless the stencil is unusually crowded).
      Walk<DIM> wk;
      int e_a = 0;
      while (any_act && e_a < E) {
        const int base = soff[e_a];
        int e_b = E;
        if (!single_group) {
          int lo = e_a, hi = E;  // largest e_b with soff[e_b] - base <= cap (soff is monotone)
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (soff[mid] - base <= sd.cap) lo = mid;
            else hi = mid - 1;
          }
          e_b = lo;
        }
        bool skip = false;
        if (e_b == e_a) {  // one cell alone exceeds the staging buffer
          if (tid == 0) atomicOr(err, SPHB200_ERR_STAGE_OVERFLOW);
          e_b = e_a + 1;
          skip = true;
        }
        if (!skip && !(single_group && staged)) {
          if (ib > 0 || e_a > 0) __syncthreads();  // previous readers done
          stage_group(e_a, e_b, base);
          __syncthreads();
          staged = single_group;
        }
        if (!skip) {
          wk.init(g, sa0, ci, ri, have);
          int cnt = 0;
          for (;;) {
            // phase 1: cheap reject; c2_fb is +inf while the cell table may be frozen (a particle
            // that drifted across the periodic seam is not where the image shift expects it)
            const bool fin = wk.run(g, sq, list, tid, sd.lcap, cnt, soff, base, e_a, e_b, nxs, slen[1],
                                    sa0, ri, act, g.c2_fb);
            // phase 2: real neighbours, exact arithmetic
#pragma unroll 1
            for (int k = 0; k < cnt; ++k) {
              const int jn = list[k * LS + tid];
              const float4 pj = sq[jn];
              float dr[3];
              if (interior) pair_disp<DIM, true>(g, ri, pj, dr);
              else pair_disp<DIM, false>(g, ri, pj, dr);
              const float d2 = sumsq<DIM>(dr);
              // Membership.  The reference decides it on d(r_sender, r_receiver)^2 < cutoff^2
              // (jax_md/partition.py:897).  For the materialiser the thread's particle IS the
              // sender, so d2 is that very number and the list is the reference's bit for bit.
              // In the physics sweeps d2 is the receiver view, which can differ in the last bit
              // for a pair on the rounding edge of the cutoff -- where every kernel is
              // max(0, .)-clamped to exactly zero (and ~(1e-7)^4 of its peak one ulp inside), so
              // such a pair contributes nothing either way.
              if (!(d2 < g.c2)) continue;
              P::pair(c, ex, own, acc, sq, sd.cap, jn, pj, dr, d2);
            }
            cnt = 0;
            if (fin) break;
          }
        }
        e_a = e_b;
      }
      if (have) P::finish(c, f, ex, p, own, acc);
    }
    if (LM == LIST_BUILD) {
      __syncthreads();
      if (tid == 0) nl.ok[tile_id] = s_bad ? 0 : 1;
    }
  }
}

}  // namespace sphb200
