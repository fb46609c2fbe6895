// sweep2.cuh -- "duo" sweeps: one thread owns TWO cell-sorted neighbours of a row and one
// neighbour list that is the union of theirs.
//
// Same job as sweep.cuh (the reference's candidate gather + prune + E-sized gathers +
// segment_sum chains, jax_md/partition.py:832-909, solver.py:722-731, :750-928) for the tiles
// whose stencil fits one staging group; what changes is the mapping of work to threads:
//
//   * the own particles of a tile are taken in pairs of slot neighbours inside a run (one row of
//     cells of the tile): particles 2k and 2k+1 of the run are one DUO (an odd run ends in a
//     duo with one particle).  They are about one particle spacing apart, so their neighbour
//     sets overlap by ~75 %: the union list of a duo is ~1.25-1.35 lists long instead of 2;
//   * every trip of the pair loop loads ONE staged neighbour j (the random shared-memory gather
//     that bounds these sweeps) and evaluates the two pairs (i0, j), (i1, j) in the two halves of
//     packed float32 registers (FADD2 / FMUL2 / FFMA2, common.cuh): the own values are packed once
//     per duo, the neighbour's values enter as broadcast scalar operands, so -- unlike two
//     neighbours per trip -- nothing has to be transposed into register pairs;
//   * list rows are stored chunk-interleaved over the duos of a run, so that the lanes of a warp
//     read (and write) consecutive 16- / 8-byte words.
//
// Roles (as in sweep.cuh, SweepDims / gate protocol shared):
//   DUO_BUILD    the search after a re-sort: union of the two particles' candidates within
//                cutoff + skin -> skin list (staged indices, ascending)
//   DUO_FILTER   first sweep of forward(): exact float32 membership test of the reference for
//                both particles (space.py:170-181, jax_md/partition.py:897) on every skin entry,
//                every step; physics; survivors (member of either) -> exact list of the step, the
//                two membership bits in bits 14 / 15 of the entry
//   DUO_CONSUME  later sweeps: step through the exact list, no test
// A tile whose stencil does not fit (or whose rows overflow) is marked not-ok by DUO_BUILD and
// is swept by the kernels of sweep.cuh, launched over the not-ok tiles only.
#pragma once
#include "cells.cuh"
#include "common.cuh"
#include "sweep.cuh"

namespace sphb200 {

constexpr int DUO_MAXT = 512;    // largest block the kernels are compiled for
constexpr int DUO_RUNS = 36;     // T[1] * T[2] upper bound
constexpr int DUO_SOFF = 2047;   // staged (row, cell) entries upper bound
constexpr int DUO_IDX = 0x3fff;  // staged index bits of a list entry
constexpr float DUO_FAR = -1.0e30f;  // coordinate of the far sentinel of the skin rows' padding
constexpr float DUO_FAR_OWN = 1.0e30f;  // ... and of the missing second particle of a duo (apart from the sentinel too)

struct DuoList {
  int* desc;           // [tiles][desc_stride] tile descriptors (k_duo)
  int desc_stride;
  unsigned short* sl;  // skin list
  unsigned short* xl;  // exact list of the step
  int* scnt;           // [rows] skin entries
  int* xcnt;           // [rows] exact entries
  unsigned char* ok;   // [tiles]
  int lmax;            // entries per row, a multiple of 8
  int min_cap;         // smallest staging capacity among the step's sweeps
  int rows_cap;        // rows allocated
};

enum { DUO_BUILD = 1, DUO_CONSUME = 2, DUO_FILTER = 3 };

// (the per-thread columns of DUO_BUILD are strided by the block size; FILTER / CONSUME keep the
// tile's descriptor, desc_ints > 0, where DUO_BUILD keeps its tile tables)
__host__ __device__ inline size_t duo_smem_bytes(int sb, int cap, int lcap, int tpb, int desc_ints = 0) {
  const size_t tables = desc_ints > 0 ? ((size_t)desc_ints * 4 + 15) / 16 * 16
                                      : (size_t)(DUO_SOFF + 1 + 4 * (DUO_RUNS + 1)) * 4;
  return (size_t)sb * cap + (size_t)lcap * tpb * 2 + tables;
}

// ---------------------------------------------------------------------------
// Policy adaptor: a policy with a packed duo body (HAS_DUO) keeps the two particles in the
// halves of packed registers; any other policy gets two scalar Own / Acc sets and two
// predicated pair() calls.
template <class P, bool = P::HAS_DUO>
struct Duo {
  struct OwnD {
    typename P::Own o[2];
  };
  struct AccD {
    typename P::Acc a[2];
  };
  __device__ static __forceinline__ void load(const typename P::Own& o0, const typename P::Own& o1,
                                              OwnD& d) {
    d.o[0] = o0;
    d.o[1] = o1;
  }
  __device__ static __forceinline__ void init(AccD& a) {
    P::init(a.a[0]);
    P::init(a.a[1]);
  }
  __device__ static __forceinline__ void pair(const Consts& c, const Extra& ex, const OwnD& o,
                                              AccD& a, const float4* sq, int cap, int j, float4 pj,
                                              const F2 (&dr)[3], F2 d2, bool v0, bool v1) {
    if (v0) {
      const float d[3] = {lo(dr[0]), lo(dr[1]), lo(dr[2])};
      P::pair(c, ex, o.o[0], a.a[0], sq, cap, j, pj, d, lo(d2));
    }
    if (v1) {
      const float d[3] = {hi(dr[0]), hi(dr[1]), hi(dr[2])};
      P::pair(c, ex, o.o[1], a.a[1], sq, cap, j, pj, d, hi(d2));
    }
  }
  __device__ static __forceinline__ void fold(const AccD& a, typename P::Acc& a0,
                                              typename P::Acc& a1) {
    a0 = a.a[0];
    a1 = a.a[1];
  }
  // sum over the two lanes that share a duo (SPLIT == 2)
  __device__ static __forceinline__ void xsum(AccD& a) {
    auto f = [](float& v) { v += __shfl_xor_sync(FULL_MASK, v, 1); };
    P::each_acc(a.a[0], f);
    P::each_acc(a.a[1], f);
  }
};
template <class P>
struct Duo<P, true> {
  using OwnD = typename P::OwnD;
  using AccD = typename P::AccD;
  __device__ static __forceinline__ void load(const typename P::Own& o0, const typename P::Own& o1,
                                              OwnD& d) {
    P::load_duo(o0, o1, d);
  }
  __device__ static __forceinline__ void init(AccD& a) { P::init_duo(a); }
  __device__ static __forceinline__ void pair(const Consts& c, const Extra& ex, const OwnD& o,
                                              AccD& a, const float4* sq, int cap, int j, float4 pj,
                                              const F2 (&dr)[3], F2 d2, bool v0, bool v1) {
    P::pair_duo(c, ex, o, a, sq, cap, j, pj, dr, d2, v0, v1);
  }
  __device__ static __forceinline__ void fold(const AccD& a, typename P::Acc& a0,
                                              typename P::Acc& a1) {
    P::fold_duo(a, a0, a1);
  }
  __device__ static __forceinline__ void xsum(AccD& a) {
    P::each_acc_duo(a, [](F2& v) {
      F2 o;
      o.v = __shfl_xor_sync(FULL_MASK, v.v, 1);
      v = add2(v, o);
    });
  }
};

// Reference displacement r_i - r_j (space.py:170-181) of one staged neighbour for both particles
// of the duo.  INTERIOR: no periodic image inside the tile's stencil (see sweep.cuh pair_disp);
// otherwise `wrap` says along which axes there is one (uniform per tile: a tile on a face of the
// box folds along one axis only, the others keep the three packed operations).
template <int DIM, bool INTERIOR>
__device__ __forceinline__ void duo_disp(const Grid& g, const F2 (&ri)[3], const float (&rA)[3],
                                         const float (&rB)[3], const float4 pj, F2 (&dr)[3],
                                         unsigned wrap) {
  const float pc[3] = {pj.x, pj.y, pj.z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (a >= DIM) {
      dr[a] = f2(0.0f);
    } else if (INTERIOR || !(wrap & (1u << a))) {
      dr[a] = disp2_nowrap2(ri[a], pc[a], g.half[a]);
    } else {
      dr[a] = f2(disp1(rA[a], pc[a], g.half[a], g.box[a]), disp1(rB[a], pc[a], g.half[a], g.box[a]));
    }
  }
}

// ---------------------------------------------------------------------------
// FILTER / CONSUME: the pair loop of one duo.
//   row_run0  list row of the first duo of the thread's run, nr duos of the run, k the thread's
//             duo inside the run: 16-byte word c of the skin row lives at
//             sl16[row_run0 * lmax / 8 + c * nr + k], 8-byte word c of the exact row at
//             xl8[row_run0 * lmax / 4 + c * nr + k]
// SPLIT lanes share a duo (CONSUME only): lane `part` takes the words part, part + SPLIT, ... of
// the exact row; the caller sums the accumulators over the lanes.
template <int DIM, class P, bool INTERIOR, bool FILTER, int SPLIT>
__device__ __forceinline__ void duo_consume(const Grid& g, const Consts& c, const Extra& ex,
                                            const DuoList& dl, const float4* sq, int cap,
                                            int row_run0, int nr, int k, int part, int nn, bool has1,
                                            unsigned wrap, const float (&rA)[3], const float (&rB)[3],
                                            const typename Duo<P>::OwnD& own,
                                            typename Duo<P>::AccD& acc, int& n_exact) {
  using D = Duo<P>;
  const F2 ri[3] = {f2(rA[0], rB[0]), f2(rA[1], rB[1]), f2(rA[2], rB[2])};
  unsigned long long* xp =
      reinterpret_cast<unsigned long long*>(dl.xl) + (size_t)row_run0 * (dl.lmax / 4) + k;
  if (FILTER) {
    // Skin rows are whole chunks of 8: DUO_BUILD pads the last one with the index of the far
    // sentinel the FILTER kernel stages behind the stencil, which fails the test like any other
    // non-member; a duo without a second particle carries a far second position (k_duo).
    const uint4* sp = reinterpret_cast<const uint4*>(dl.sl) + (size_t)row_run0 * (dl.lmax / 8) + k;
    unsigned xlo = 0u, xhi = 0u;
    int m = 0;
    uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
    if (nn > 0) nxt = __ldg(sp);
#pragma unroll 1
    for (int kk = 0; kk < nn; kk += 8) {
      const uint4 cur = nxt;
      sp += nr;
      if (kk + 8 < nn) nxt = __ldg(sp);
      const unsigned w[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned e = (u & 1) ? (w[u >> 1] >> 16) : (w[u >> 1] & 0xffffu);
        const int j = (int)e;
        const float4 pj = sq[j];
        F2 dr[3];
        duo_disp<DIM, INTERIOR>(g, ri, rA, rB, pj, dr, wrap);
        const F2 d2 = sumsq2<DIM>(dr);
        const bool v0 = lo(d2) < g.c2;
        const bool v1 = hi(d2) < g.c2;
        // member of either -> the exact list: shift into a 64-bit window, store every fourth
        if (v0 || v1) {
          const unsigned ent = e | (v0 ? 0x4000u : 0u) | (v1 ? 0x8000u : 0u);
          xlo = __funnelshift_r(xlo, xhi, 16);
          xhi = (xhi >> 16) | (ent << 16);
          ++m;
          if ((m & 3) == 0) {
            *xp = ((unsigned long long)xhi << 32) | xlo;
            xp += nr;
          }
        }
        D::pair(c, ex, own, acc, sq, cap, j, pj, dr, d2, v0, v1);
      }
    }
    if (m & 3) {  // last, partial group: unused slots are zero
      const unsigned long long xw = ((unsigned long long)xhi << 32) | xlo;
      *xp = xw >> (16 * (4 - (m & 3)));
    }
    n_exact = m;
  } else {
    // exact list: a zero entry (unused slot of the last word) has no membership bit set
    // (words are fetched two trips ahead: a trip is shorter than a round trip to HBM)
    const int nwords = (nn + 3) >> 2, step = SPLIT * nr;
    unsigned long long nxt = 0ull, nxt2 = 0ull;
    xp += part * nr;
    if (part < nwords) nxt = __ldg(xp);
    if (part + SPLIT < nwords) nxt2 = __ldg(xp + step);
    xp += step;
#pragma unroll 1
    for (int w = part; w < nwords; w += SPLIT) {
      const unsigned long long cur = nxt;
      nxt = nxt2;
      xp += step;
      if (w + 2 * SPLIT < nwords) nxt2 = __ldg(xp);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned e = (unsigned)(cur >> (16 * u)) & 0xffffu;
        const int j = (int)(e & DUO_IDX);
        const bool v0 = (e & 0x4000u) != 0u, v1 = (e & 0x8000u) != 0u;
        const float4 pj = sq[j];
        F2 dr[3];
        duo_disp<DIM, INTERIOR>(g, ri, rA, rB, pj, dr, wrap);
        // (membership was decided by the FILTER sweep on the unfused sum: here d^2 only feeds
        // the physics and may keep the extra bits of the fused form)
        F2 d2 = fma2(dr[1], dr[1], mul2(dr[0], dr[0]));
        if (DIM == 3) d2 = fma2(dr[2], dr[2], d2);
        D::pair(c, ex, own, acc, sq, cap, j, pj, dr, d2, v0, v1);
      }
    }
    n_exact = nn;
  }
}

// ---------------------------------------------------------------------------
// The search of one duo: the walk of sweep.cuh (Walk) through the union of the two particles'
// windows -- same rows (the duo shares its row of cells), x cells from the first particle's
// window start to the second one's window end -- with the cheap squared-distance test of both
// particles in the two halves of packed registers.
template <int DIM>
struct Walk2 {
  int w0y, w0z, ka, kb, ks, it_step, nit;
  F2 xs0, xs1;
  int it, wy, wz, j, jb, row;
  bool inwin;
  F2 xs, ys, zs;

  __device__ __forceinline__ void init(const Grid& g, const int (&sa0)[3], const int (&ci)[3],
                                       int cx1, const float (&rA)[3], const float (&rB)[3],
                                       bool have) {
    w0y = (g.n[1] >= 2 * g.S[1] + 1) ? (ci[1] - g.S[1] - sa0[1]) : 0;
    w0z = (g.n[2] >= 2 * g.S[2] + 1) ? (ci[2] - g.S[2] - sa0[2]) : 0;
    const bool wide = g.n[0] >= 2 * g.S[0] + 1;
    const int kz = -sa0[0], kn = g.n[0] - sa0[0];  // staged x index of unwrapped cells 0 and n
    ka = wide ? (ci[0] - g.S[0] - sa0[0]) : 0;
    kb = wide ? (cx1 + g.S[0] + 1 - sa0[0]) : g.W[0];
    ks = kb;
    if (ka < kz && kz < kb) ks = kz;
    else if (ka < kn && kn < kb) ks = kn;
    const float sh0 = (float)wrap_count(sa0[0] + ka, g.n[0]) * g.box[0];
    const float sh1 = (float)wrap_count(sa0[0] + ks, g.n[0]) * g.box[0];
    xs0 = f2(rA[0] - sh0, rB[0] - sh0);
    xs1 = f2(rA[0] - sh1, rB[0] - sh1);
    it_step = __any_sync(FULL_MASK, have && ks < kb) ? 1 : 2;
    nit = 2 * g.W[1] * g.W[2];
    it = -1; wy = -1; wz = 0; j = 0; jb = 0; row = 0;
    inwin = false;
    xs = ys = zs = f2(0.f);
  }

  __device__ __forceinline__ bool near(const float4 p, float thr) const {
    F2 t = sub2(xs, f2(p.x));
    F2 d = mul2(t, t);
    t = sub2(ys, f2(p.y));
    d = fma2(t, t, d);
    if (DIM == 3) {
      t = sub2(zs, f2(p.z));
      d = fma2(t, t, d);
    }
    return fminf(lo(d), hi(d)) < thr;
  }

  // appends the candidates near either particle to the thread's shared-memory column; returns
  // true when the window is exhausted, false when some lane's column cannot take the next chunk
  __device__ __forceinline__ bool run(const Grid& g, const float4* sq, unsigned short* list, int tid,
                                      int ls, int lcap, int& cnt, const int* soff, int E, int nxs,
                                      int slen1, const int (&sa0)[3], const float (&rA)[3],
                                      const float (&rB)[3], bool act, float thr) {
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
          const int ry = w0y + wy, rz = w0z + wz;
          row = rz * slen1 + ry;
          inwin = act;
          const float shy = (float)wrap_count(sa0[1] + ry + g.goff[1], g.ng[1]) * g.box[1];
          const float shz = (float)wrap_count(sa0[2] + rz + g.goff[2], g.ng[2]) * g.box[2];
          ys = f2(rA[1] - shy, rB[1] - shy);
          zs = f2(rA[2] - shz, rB[2] - shz);
        }
        const int kk0 = sgm == 0 ? ka : ks, kk1 = sgm == 0 ? ks : kb;
        j = jb = 0;
        if (inwin && kk0 < kk1) {
          const int ea = max(row * nxs + kk0, 0), eb = min(row * nxs + kk1, E);
          if (ea < eb) {
            j = soff[ea];
            jb = soff[eb];
          }
        }
        xs = sgm == 0 ? xs0 : xs1;
        continue;
      }
      const int want = rem > 0 ? min(rem, SWEEP_CHUNK) : 0;
      if (__any_sync(FULL_MASK, want > lcap - cnt)) return false;  // drain first
      const int e = j + want;
      unsigned short* col = list + tid;
      // four candidates per trip, loaded before the first append (see sweep.cuh, Walk::run)
      for (; j + 4 <= e; j += 4) {
        const float4 p0 = sq[j], p1 = sq[j + 1], p2 = sq[j + 2], p3 = sq[j + 3];
        const bool n0 = near(p0, thr), n1 = near(p1, thr), n2 = near(p2, thr), n3 = near(p3, thr);
        if (n0) { col[cnt * ls] = (unsigned short)j; ++cnt; }
        if (n1) { col[cnt * ls] = (unsigned short)(j + 1); ++cnt; }
        if (n2) { col[cnt * ls] = (unsigned short)(j + 2); ++cnt; }
        if (n3) { col[cnt * ls] = (unsigned short)(j + 3); ++cnt; }
      }
#pragma unroll 1
      for (; j < e; ++j) {
        if (near(sq[j], thr)) {
          col[cnt * ls] = (unsigned short)j;
          ++cnt;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Tile descriptor (ints), written by DUO_BUILD after every re-sort and read by the FILTER /
// CONSUME sweeps of every step until the next one (the particles keep their slots and the cell
// table is frozen in between, so the staging plan of a tile does not change):
//   [0] staged particles  [1] own particles  [2] duos  [3] copy jobs  [4] runs
//   [DD_OWN_OFF + r]  own particles before run r      (runs + 1)
//   [DD_OWN_START + r] first slot of run r
//   [DD_DUO_OFF + r]  duos before run r               (runs + 1)
//   [DD_ROW0 + r]     list row of the run's first duo
//   [DD_JOBS + 2 j]   copy job j: first source slot; [.. + 1] staged index | particles << 16
// A copy job is one contiguous slot range of one stencil row (a row has up to three: the periodic
// images of its x range); FILTER / CONSUME turn every job into one bulk copy
// (cp.async.bulk, global -> shared, completion on an mbarrier) per staged 16-byte array.
constexpr int DD_OWN_OFF = 8;
constexpr int DD_OWN_START = DD_OWN_OFF + DUO_RUNS + 1;
constexpr int DD_DUO_OFF = DD_OWN_START + DUO_RUNS;
constexpr int DD_ROW0 = DD_DUO_OFF + DUO_RUNS + 1;
constexpr int DD_JOBS = (DD_ROW0 + DUO_RUNS + 3) / 4 * 4;
__host__ __device__ inline int duo_desc_ints(int rows) { return (DD_JOBS + 2 * 3 * rows + 3) / 4 * 4; }

// (register cap per policy: sweeps with a small staged record run two blocks of up to 384
// threads per SM, P::DUO_MINB == 2; the block size itself is a run-time choice up to DUO_MAXT)
// SPLIT (DUO_CONSUME): lanes per duo, see duo_consume.
template <int DIM, class P, int ROLE, int SPLIT = 1>
__global__ void __maxnreg__(SPLIT > 1 ? 80 : (P::DUO_MINB > 1 ? 80 : 128))
    k_duo(const Grid g, const Consts c, const Frame f, const int* __restrict__ cs,
          const SweepDims sd, const Extra ex, unsigned* __restrict__ err, const DuoList dl) {
  if (sd.gate != nullptr && *sd.gate != sd.gate_want) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using D = Duo<P>;
  const int TPB = blockDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = TPB >> 5;
  float4* sq = reinterpret_cast<float4*>(smem_raw);
  unsigned short* list = reinterpret_cast<unsigned short*>(smem_raw + (size_t)sd.sb * sd.cap);
  // BUILD: tile tables built from the cell table; FILTER / CONSUME: the tile's descriptor
  int* soff = reinterpret_cast<int*>(list + (size_t)sd.lcap * TPB);
  int* own_start = soff + (DUO_SOFF + 1);
  int* own_off = own_start + (DUO_RUNS + 1);
  int* duo_off = own_off + (DUO_RUNS + 1);
  int* row0 = duo_off + (DUO_RUNS + 1);
  int* dsc = soff;
  __shared__ int s_bad;
  __shared__ __align__(8) unsigned long long s_bar;
  unsigned bar_parity = 0;
  (void)err;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (ROLE != DUO_BUILD) {
    own_off = dsc + DD_OWN_OFF;
    own_start = dsc + DD_OWN_START;
    duo_off = dsc + DD_DUO_OFF;
    row0 = dsc + DD_ROW0;
  }

  for (int tq = blockIdx.x; tq < g.ntl; tq += gridDim.x) {
    int b = tq + g.block0;
    const int tile_id = b;
    if (ROLE != DUO_BUILD && dl.ok[tile_id] == 0) continue;  // swept by sweep.cuh
    if (tq != (int)blockIdx.x) __syncthreads();  // readers of the previous tile's tables are done

    // ---- tile geometry (uniform), as in sweep.cuh ----------------------------
    const int tx = b % g.nt[0];
    b /= g.nt[0];
    const int ty = b % g.nt[1];
    const int tz = b / g.nt[1];
    int c0[3] = {g.own_lo[0] + tx * g.T[0], g.own_lo[1] + ty * g.T[1], g.own_lo[2] + tz * g.T[2]};
    int no[3], sa0[3], slen[3];
    unsigned wrap = 0u;  // axes along which the stencil holds a periodic image
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
      if (a < DIM && !(sa0[a] + g.goff[a] - im >= 0 && sa0[a] + g.goff[a] + slen[a] + im <= g.ng[a] &&
                       g.n[a] >= 2 * g.S[a] + 2))
        wrap |= 1u << a;
    }
    const bool interior = wrap == 0u;
    const int nxs = slen[0];
    const int nrows = slen[1] * slen[2];
    const int E = nrows * nxs;
    const int nruns = no[1] * no[2];
    int* gdsc = dl.desc + (size_t)tile_id * dl.desc_stride;
    int tile_n, tile_duos, total_staged;

    if constexpr (ROLE == DUO_BUILD) {
      if (tid < nruns) {
        const int ry = tid % no[1], rz = tid / no[1];
        const int cy = c0[1] + ry, cz = c0[2] + rz;
        const int cell = (cz * g.n[1] + cy) * g.n[0] + c0[0];
        const int s = cs[cell];
        own_start[tid] = s;
        own_off[tid + 1] = cs[cell + no[0]] - s;
        // list row of the run's first duo: runs in slot order are numbered o = (cz n1 + cy) nt0 + tx,
        // and ceil((s + o) / 2) leaves room for the ceil(len / 2) rows of every run before it
        const int o = (cz * g.n[1] + cy) * g.nt[0] + tx;
        row0[tid] = (s + o + 1) >> 1;
      }
      for (int e = tid; e < E; e += TPB) {
        const int k = e % nxs, row = e / nxs;
        const int ry = row % slen[1], rz = row / slen[1];
        const int cell = (wrap_cell(sa0[2] + rz, g.n[2]) * g.n[1] + wrap_cell(sa0[1] + ry, g.n[1])) * g.n[0] +
                         wrap_cell(sa0[0] + k, g.n[0]);
        soff[e] = cs[cell + 1] - cs[cell];
      }
      __syncthreads();
      if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < E; base += 32) {
          const int i = base + lane;
          const int v = i < E ? soff[i] : 0;
          int inc = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, inc, o);
            if (lane >= o) inc += t;
          }
          if (i < E) soff[i] = carry + inc - v;
          carry += __shfl_sync(FULL_MASK, inc, 31);
        }
        if (lane == 0) {
          soff[E] = carry;
          int acc = 0, dacc = 0;
          own_off[0] = 0;
          duo_off[0] = 0;
          for (int r = 0; r < nruns; ++r) {
            const int len = own_off[r + 1];
            acc += len;
            dacc += (len + 1) >> 1;
            own_off[r + 1] = acc;
            duo_off[r + 1] = dacc;
          }
        }
      }
      __syncthreads();
      tile_n = own_off[nruns];
      tile_duos = duo_off[nruns];
      total_staged = soff[E];
      // ---- the tile's descriptor for the FILTER / CONSUME sweeps of the coming steps ----
      if (tid == 0) {
        gdsc[0] = total_staged; gdsc[1] = tile_n; gdsc[2] = tile_duos; gdsc[3] = 3 * nrows; gdsc[4] = nruns;
        gdsc[DD_OWN_OFF] = 0;
        gdsc[DD_DUO_OFF] = 0;
      }
      if (tid < nruns) {
        gdsc[DD_OWN_OFF + tid + 1] = own_off[tid + 1];
        gdsc[DD_DUO_OFF + tid + 1] = duo_off[tid + 1];
        gdsc[DD_OWN_START + tid] = own_start[tid];
        gdsc[DD_ROW0 + tid] = row0[tid];
      }
      if (tile_n == 0) {
        if (tid == 0) dl.ok[tile_id] = 1;
        continue;
      }
      if (total_staged + 1 > dl.min_cap || total_staged > DUO_IDX) {  // (+1: the far sentinel)
        if (tid == 0) dl.ok[tile_id] = 0;
        continue;
      }
      if (tid == 0) {
        s_bad = 0;  // ordered before any other write by the staging barrier
        mbar_expect_tx(&s_bar, (unsigned)total_staged * 16u);
      }
      fence_proxy_async();  // the previous tile's reads of the buffer precede the copies
      for (int job = tid; job < 3 * nrows; job += TPB) {
        const int row = job / 3;
        const int seg = job % 3 - 1;
        const int k0 = max(0, seg * g.n[0] - sa0[0]);
        const int k1 = min(nxs, (seg + 1) * g.n[0] - sa0[0]);
        int src = 0, dst = 0, len = 0;
        if (k0 < k1) {
          const int ry = row % slen[1], rz = row / slen[1];
          const int rowcell =
              (wrap_cell(sa0[2] + rz, g.n[2]) * g.n[1] + wrap_cell(sa0[1] + ry, g.n[1])) * g.n[0];
          src = cs[rowcell + (sa0[0] + k0 - seg * g.n[0])];
          dst = soff[row * nxs + k0];
          len = soff[row * nxs + k1] - dst;
        }
        gdsc[DD_JOBS + 2 * job] = src;
        gdsc[DD_JOBS + 2 * job + 1] = dst | (len << 16);
        // the search stages positions only: one bulk copy per job
        if (len > 0) bulk_g2s(sq + dst, f.pt + src, (unsigned)len * 16u, &s_bar);
      }
    } else {
      // ---- descriptor -> shared memory, then one bulk copy per job and staged array ----
      const int nd = dl.desc_stride;
      for (int i = tid; i < nd; i += TPB) dsc[i] = __ldg(gdsc + i);
      __syncthreads();
      total_staged = dsc[0];
      tile_n = dsc[1];
      tile_duos = dsc[2];
      if (tile_n == 0) continue;
      const int njobs = dsc[3];
      const float4* src_arr[P::DUO_COPIES > 0 ? P::DUO_COPIES : 1];
      P::duo_sources(f, ex, src_arr);
      if (P::DUO_COPIES > 0) {
        // every thread issues the copies of its jobs (a copy is issued by one lane at a time, so
        // the issue is spread over all warps); the arrival that arms the barrier may come after
        // some copies have completed: the transaction count is signed
        if (tid == 0) mbar_expect_tx(&s_bar, (unsigned)total_staged * 16u * P::DUO_COPIES);
        fence_proxy_async();  // the previous tile's reads of the buffer precede the copies
        for (int job = tid; job < njobs; job += TPB) {
          const int src = dsc[DD_JOBS + 2 * job], dlen = dsc[DD_JOBS + 2 * job + 1];
          const int dst = dlen & 0xffff, len = dlen >> 16;
          if (len > 0) {
#pragma unroll
            for (int a = 0; a < P::DUO_COPIES; ++a)
              bulk_g2s(sq + (size_t)a * sd.cap + dst, src_arr[a] + src, (unsigned)len * 16u, &s_bar);
          }
        }
      }
      // what the bulk copies do not cover: a 4-byte column (asynchronous 4-byte copies, a warp per
      // job), or -- a policy without source arrays -- everything, through registers
      if (P::DUO_COPIES == 0 || P::DUO_REST) {
        for (int job = warp; job < njobs; job += nwarps) {
          const int src = dsc[DD_JOBS + 2 * job], dlen = dsc[DD_JOBS + 2 * job + 1];
          const int dst = dlen & 0xffff, len = dlen >> 16;
          for (int m = lane; m < len; m += 32) {
            if (P::DUO_COPIES == 0) P::stage(c, f, ex, src + m, sq, sd.cap, dst + m);
            else P::stage_rest(c, f, ex, src + m, sq, sd.cap, dst + m);
          }
        }
      }
    }
    bool staged = false;  // the wait for the staged data comes after the thread's own loads

    for (int ib = 0; ib < tile_duos; ib += TPB / SPLIT) {
      // ---- the thread's duo ----------------------------------------------------
      const int t = ib + tid / SPLIT, part = tid % SPLIT;
      const bool have = t < tile_duos;
      int run = 0, k = 0, p0 = 0, nr = 1;
      bool has1 = false;
      if (have) {
        while (t >= duo_off[run + 1]) ++run;
        k = t - duo_off[run];
        nr = duo_off[run + 1] - duo_off[run];
        p0 = own_start[run] + 2 * k;
        has1 = 2 * k + 1 < own_off[run + 1] - own_off[run];
      }
      const int p1 = has1 ? p0 + 1 : p0;
      const int row_run0 = row0[run];
      float rA[3] = {0.f, 0.f, 0.f}, rB[3] = {0.f, 0.f, 0.f};
      float4 qA = make_float4(0.f, 0.f, 0.f, 0.f), qB = qA;
      if (have) {
        qA = f.pt[p0];
        qB = f.pt[p1];
        rA[0] = qA.x; rA[1] = qA.y; rA[2] = qA.z;
        rB[0] = qB.x; rB[1] = qB.y; rB[2] = qB.z;
        // no second particle: a position no neighbour is near (every distance overflows to +inf,
        // which fails the search's and the FILTER's tests without a special case)
        if (!has1) rB[0] = rB[1] = rB[2] = DUO_FAR_OWN;
      }

      if constexpr (ROLE == DUO_BUILD) {
        // ---------------- the search: union skin list of the duo -> HBM ----------------
        int ci[3] = {0, 0, 0}, cx1 = 0;
        if (have) {
          const int ry = run % no[1], rz = run / no[1];
          const int rowcell0 = ((c0[2] + rz) * g.n[1] + (c0[1] + ry)) * g.n[0] + c0[0];
          int kx = 0;
          while (kx + 1 < no[0] && __ldg(cs + rowcell0 + kx + 1) <= p0) ++kx;
          ci[0] = c0[0] + kx; ci[1] = c0[1] + ry; ci[2] = c0[2] + rz;
          while (kx + 1 < no[0] && __ldg(cs + rowcell0 + kx + 1) <= p1) ++kx;
          cx1 = c0[0] + kx;
        }
        if (!staged) {  // (uniform) the stencil has arrived
          if (ROLE == DUO_BUILD || P::DUO_COPIES > 0) {
            mbar_wait(&s_bar, bar_parity);
            bar_parity ^= 1u;
          }
          if (ROLE != DUO_BUILD && P::DUO_REST) cp_async_wait_all();
          // the far sentinel behind the stencil: what the padding of the skin rows points at
          if (ROLE == DUO_FILTER && tid == 0)
            sq[total_staged] = make_float4(DUO_FAR, DUO_FAR, DUO_FAR, 0.f);
          __syncthreads();
          staged = true;
        }
        Walk2<DIM> wk;
        wk.init(g, sa0, ci, cx1, rA, rB, have);
        int cnt = 0, gk = 0;  // entries waiting in the column / already in HBM
        unsigned short* col = list + tid;
        uint4* gp = reinterpret_cast<uint4*>(dl.sl) + (size_t)row_run0 * (dl.lmax / 8) + k;
        const bool rows_fit = row_run0 + nr <= dl.rows_cap;  // (holds by construction of rows_cap)
        if (have && !rows_fit) s_bad = 1;
        for (;;) {
          const bool fin = wk.run(g, sq, list, tid, TPB, sd.lcap, cnt, soff, E, nxs, slen[1], sa0, rA,
                                  rB, have, g.c2_hi);
          int w = 0;
          for (; cnt - w >= 8; w += 8) {
            uint4 v;
            v.x = (unsigned)col[(w + 0) * TPB] | ((unsigned)col[(w + 1) * TPB] << 16);
            v.y = (unsigned)col[(w + 2) * TPB] | ((unsigned)col[(w + 3) * TPB] << 16);
            v.z = (unsigned)col[(w + 4) * TPB] | ((unsigned)col[(w + 5) * TPB] << 16);
            v.w = (unsigned)col[(w + 6) * TPB] | ((unsigned)col[(w + 7) * TPB] << 16);
            if (gk + 8 <= dl.lmax && rows_fit) {
              *gp = v;
              gp += nr;
            } else {
              s_bad = 1;
            }
            gk += 8;
          }
          const int r = cnt - w;
          if (w > 0)
            for (int i = 0; i < r; ++i) col[i * TPB] = col[(w + i) * TPB];
          cnt = r;
          if (fin) break;
        }
        if (have) {
          if (cnt > 0) {  // tail chunk, padded with the index of the far sentinel (see duo_consume)
            unsigned e8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e8[i] = i < cnt ? (unsigned)col[i * TPB] : (unsigned)total_staged;
            if (gk + 8 <= dl.lmax && rows_fit)
              *gp = make_uint4(e8[0] | (e8[1] << 16), e8[2] | (e8[3] << 16), e8[4] | (e8[5] << 16),
                               e8[6] | (e8[7] << 16));
            else
              s_bad = 1;
          }
          if (rows_fit) dl.scnt[row_run0 + k] = gk + cnt;
        }
      } else {
        // ---------------- list consumer (FILTER: exact test + exact list of the step) ------
        typename P::Own oA, oB;
        bool actA = false, actB = false;
        if (have) {
          P::load_own(c, f, ex, p0, qA, oA);
          P::load_own(c, f, ex, p1, qB, oB);
          actA = P::active(c, oA);
          actB = has1 && P::active(c, oB);
        } else {
          oA = typename P::Own();
          oB = oA;
        }
        typename P::Acc aA, aB;
        P::init(aA);
        P::init(aB);
        const bool any_act = P::SPARSE ? (__syncthreads_or(actA || actB) != 0) : true;
        const int row = row_run0 + k;
        int nn = 0;
        if (any_act && have && (actA || actB)) nn = ROLE == DUO_FILTER ? dl.scnt[row] : dl.xcnt[row];
        if (!staged) {  // (uniform) the stencil has arrived
          if (ROLE == DUO_BUILD || P::DUO_COPIES > 0) {
            mbar_wait(&s_bar, bar_parity);
            bar_parity ^= 1u;
          }
          if (ROLE != DUO_BUILD && P::DUO_REST) cp_async_wait_all();
          // the far sentinel behind the stencil: what the padding of the skin rows points at
          if (ROLE == DUO_FILTER && tid == 0)
            sq[total_staged] = make_float4(DUO_FAR, DUO_FAR, DUO_FAR, 0.f);
          __syncthreads();
          staged = true;
        }
        if (any_act) {
          typename D::OwnD od;
          typename D::AccD ad;
          D::load(oA, oB, od);
          D::init(ad);
          int n_exact = 0;
          if (interior)
            duo_consume<DIM, P, true, ROLE == DUO_FILTER, SPLIT>(g, c, ex, dl, sq, sd.cap, row_run0,
                                                                 nr, k, part, nn, has1, wrap, rA, rB, od, ad,
                                                                 n_exact);
          else
            duo_consume<DIM, P, false, ROLE == DUO_FILTER, SPLIT>(g, c, ex, dl, sq, sd.cap, row_run0,
                                                                  nr, k, part, nn, has1, wrap, rA, rB, od,
                                                                  ad, n_exact);
          if (SPLIT > 1) D::xsum(ad);
          D::fold(ad, aA, aB);
          if (ROLE == DUO_FILTER && have && part == 0) dl.xcnt[row] = n_exact;
        }
        if (have) {
          // the epilogue re-reads the own values (nothing of them has to stay in registers
          // across the pair loop beyond what the packed body keeps); lanes that share a duo
          // take one particle each
          if (SPLIT == 1 || part == 0) {
            P::load_own(c, f, ex, p0, qA, oA);
            P::finish(c, f, ex, p0, oA, aA);
          }
          if (has1 && (SPLIT == 1 || part == 1)) {
            P::load_own(c, f, ex, p1, qB, oB);
            P::finish(c, f, ex, p1, oB, aB);
          }
        }
      }
    }
    if (ROLE == DUO_BUILD) {
      __syncthreads();
      if (tid == 0) dl.ok[tile_id] = s_bad ? 0 : 1;
    }
  }
}


// The compact force record of every slot (phys.cuh, PhysForce::make_record) -> Extra::rec*, from
// where the duo force sweep stages it with bulk copies.  mode 0: all slots, 1: own slots only,
// 2: the two halo ranges only (slab engines, cells.cuh Slab).
template <class P>
__global__ void __launch_bounds__(256) k_force_rec(int n, Slab sl, int mode, Frame f, Extra ex) {
  int lo = 0, hi = n, lo2 = 0, hi2 = 0;
  if (sl.dn != nullptr && mode == 0) {  // slab engines: the live slots only (halo + own + halo)
    lo = sl.base - sl.dn[DN_HALO_LO];
    hi = sl.base + sl.dn[DN_OWN] + sl.dn[DN_HALO_HI];
  } else if (sl.dn != nullptr) {
    const int own = sl.dn[DN_OWN];
    if (mode == 1) {
      lo = sl.base;
      hi = sl.base + own;
    } else {
      lo = sl.base - sl.dn[DN_HALO_LO];
      hi = sl.base;
      lo2 = sl.base + own;
      hi2 = lo2 + sl.dn[DN_HALO_HI];
    }
  }
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int p = lo + t0; p < hi; p += stride) P::make_record(f, ex, p);
  for (int p = lo2 + t0; p < hi2; p += stride) P::make_record(f, ex, p);
}

// Diagnostics (sphb200_engine_counters): directed pairs in the exact lists of the last step -- the
// membership bits of every duo row of every tile that has lists.  Equal to the brute-force count
// of pairs (i, j) with d^2 < cutoff^2 at the step's positions, self pairs included, whenever the
// lists are complete (tests/test_gpu_duo.py).
__global__ void __launch_bounds__(256) k_duo_pair_count(DuoList dl, unsigned long long* out) {
  const int tile = blockIdx.x;
  unsigned long long cnt = 0ull;
  if (dl.ok[tile] != 0) {
    const int* d = dl.desc + (size_t)tile * dl.desc_stride;
    const int duos = d[1] > 0 ? d[2] : 0;
    for (int t = threadIdx.x; t < duos; t += blockDim.x) {
      int run = 0;
      while (t >= d[DD_DUO_OFF + run + 1]) ++run;
      const int k = t - d[DD_DUO_OFF + run], nr = d[DD_DUO_OFF + run + 1] - d[DD_DUO_OFF + run];
      const int row_run0 = d[DD_ROW0 + run];
      const int nwords = (dl.xcnt[row_run0 + k] + 3) >> 2;
      const unsigned long long* xp =
          reinterpret_cast<const unsigned long long*>(dl.xl) + (size_t)row_run0 * (dl.lmax / 4) + k;
      for (int w = 0; w < nwords; ++w) {
        const unsigned long long cur = xp[(size_t)w * nr];
        // bits 14 / 15 of every 16-bit entry
        cnt += __popcll(cur & 0xC000C000C000C000ull);
      }
    }
  }
  __shared__ unsigned long long s;
  if (threadIdx.x == 0) s = 0ull;
  __syncthreads();
  if (cnt) atomicAdd(&s, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && s) atomicAdd(out, s);
}

// number of tiles the duo sweeps leave to sweep.cuh -> *nbad (the gate of those launches)
__global__ void k_duo_count_bad(const unsigned char* __restrict__ ok, int ntiles, int* nbad,
                                const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  __shared__ int s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  int n = 0;
  for (int i = threadIdx.x; i < ntiles; i += blockDim.x) n += ok[i] == 0 ? 1 : 0;
  if (n) atomicAdd(&s, n);
  __syncthreads();
  if (threadIdx.x == 0) *nbad = s;
}

}  // namespace sphb200
