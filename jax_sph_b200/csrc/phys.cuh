// phys.cuh -- pair physics policies for k_sweep (see sweep.cuh).
//
// Each policy restates, for ONE receiver particle i and ONE neighbour j, the
// per-edge expressions of the reference and what its segment_sum accumulates;
// finish() is the per-particle epilogue.  Citations: jax_sph/solver.py.
//
//   PhysDensity   :750-802  density summation / continuity (SPH) / Riemann continuity,
//                           EoS; RIE wall helpers u_tilde (:547-552) and heat_bc (:556-565)
//   PhysRenorm    :167-173  Shepard density renormalisation
//   PhysDelta     :33-105   Delta-SPH density evolution (renormalised density diffusion)
//   PhysWall      :450-528  generalized wall boundary condition (SPH)
//   PhysForce     :221-256 standard acceleration, :199-213 TVF, :316-401 Riemann,
//                 :404-428 artificial viscosity, :591-610 heat conduction,
//                 integrator.py:54 case bc_fn (table form, dudt/dvdt/dTdt part)
//   PhysNeighbors jax_md/partition.py:885-909  sparse list materialiser (parity / drop-in .idx)
//
// Registers held across the pair loop are kept to what pair() reads; finish()
// re-reads the rest of the particle from global memory.
#pragma once
#include "cells.cuh"
#include "common.cuh"

namespace sphb200 {

__device__ __forceinline__ float dot3(const float (&a)[3], const float (&b)[3], int dim) {
  float s = a[0] * b[0] + a[1] * b[1];
  if (dim == 3) s += a[2] * b[2];
  return s;
}

// ---------------------------------------------------------------------------
// The search (LIST_BUILD) carries no physics: it stages positions only.
struct PhysNone {
  static constexpr int MINB = 2;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 2;
  static constexpr int DUO_COPIES = 1;  // positions, as they lie in the frame
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame& f, const Extra&, const float4* (&a)[1]) { a[0] = f.pt; }
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  struct Own {};
  struct Acc {};
  template <class F>
  __device__ static void each_acc(Acc&, F) {}
  __device__ static void stage(const Consts&, const Frame& f, const Extra&, int gp, float4* sq,
                               int, int d) {
    sq[d] = f.pt[gp];
  }
  __device__ static void load_own(const Consts&, const Frame&, const Extra&, int, float4, Own&) {}
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc&) {}
  __device__ static void pair(const Consts&, const Extra&, const Own&, Acc&, const float4*, int,
                              int, float4, const float (&)[3], float) {}
  __device__ static void finish(const Consts&, const Frame&, const Extra&, int, const Own&,
                                const Acc&) {}
};

// ---------------------------------------------------------------------------
// DENS_SUM_X: summation plus the RIE wall helpers (u_tilde, wall temperature); the helpers
// are compiled out of DENS_SUM / DENS_EVOL_SPH, which never need them.
enum { DENS_SUM = 0, DENS_EVOL_SPH = 1, DENS_EVOL_RIE = 2, DENS_SUM_X = 3 };

template <int DIM, int KERN, int MODE>
struct PhysDensity {
  static constexpr int MINB = (MODE == DENS_EVOL_RIE || MODE == DENS_SUM_X) ? 1 : 2;  // the heavy modes stage 48-64 B: one block per SM anyway
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;  // true: most particles are inactive (tile skip, sweep.cuh)
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = MODE != DENS_EVOL_RIE;
  static constexpr bool SUM = MODE == DENS_SUM || MODE == DENS_SUM_X;
  static constexpr bool XTRA = MODE == DENS_SUM_X || MODE == DENS_EVOL_RIE;
  struct Own {
    float u[3], g[3];
    float rho, p;
  };
  struct Acc {
    float s;        // sum w  |  continuity sum
    float swf, sT;  // RIE wall helpers
    float su[3];
  };
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
    f(a.s);
    if (XTRA) {
      f(a.swf); f(a.sT); f(a.su[0]); f(a.su[1]); f(a.su[2]);
    }
  }
  // plain summation: the packed pair body of the list consumers (sweep.cuh, Pair2)
  static constexpr bool HAS_PAIR2 = MODE == DENS_SUM;
  struct Acc2 {
    F2 s;
  };
  __device__ static void init2(Acc2& a) { a.s = f2(0.0f); }
  __device__ static void pair2(const Consts& c, const Extra&, const Own&, Acc2& a, const float4*,
                               int, int, int, float4, float4, const F2 (&)[3], F2 d2, bool v0,
                               bool v1) {
    add2_into(a.s, sel2(v0, v1, kernel_w2<KERN>(c, fsqrt2(d2))));
  }
  __device__ static void fold(const Acc2& a2, Acc& a) {
    init(a);
    a.s = lo(a2.s) + hi(a2.s);
  }
  // plain summation: packed body of the duo sweeps (sweep2.cuh): the two particles of a duo in
  // the two halves, one neighbour per call
  static constexpr bool HAS_DUO = MODE == DENS_SUM;
  static constexpr int DUO_MINB = MODE == DENS_SUM ? 2 : 1;
  static constexpr int DUO_COPIES = MODE == DENS_SUM ? 1 : 0;  // positions, as they lie in the frame
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame& f, const Extra&, const float4* (&a)[1]) { a[0] = f.pt; }
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  struct OwnD {};
  struct AccD {
    F2 s;
  };
  __device__ static void load_duo(const Own&, const Own&, OwnD&) {}
  __device__ static void init_duo(AccD& a) { a.s = f2(0.0f); }
  __device__ static void pair_duo(const Consts& c, const Extra&, const OwnD&, AccD& a,
                                  const float4*, int, int, float4, const F2 (&)[3], F2 d2, bool v0,
                                  bool v1) {
    add2_into(a.s, sel2(v0, v1, kernel_w2<KERN>(c, fsqrt2(d2))));
  }
  __device__ static void fold_duo(const AccD& d, Acc& a0, Acc& a1) {
    init(a0);
    init(a1);
    a0.s = lo(d.s);
    a1.s = hi(d.s);
  }
  template <class F>
  __device__ static void each_acc_duo(AccD& a, F f) { f(a.s); }
  __device__ static void stage(const Consts& c, const Frame& f, const Extra& ex, int gp,
                               float4* sq, int cap, int d) {
    sq[d] = f.pt[gp];
    if (MODE == DENS_SUM) {
    } else if (MODE == DENS_SUM_X) {
      sq[cap + d] = f.um[gp];
      sq[2 * cap + d] = f.st[gp];
    } else if (MODE == DENS_EVOL_SPH) {
      float4 um = f.um[gp], st = f.st[gp];
      sq[cap + d] = make_float4(um.x, um.y, um.z, um.w / st.x);  // (mass / rho)[j], solver.py:26
    } else {
      sq[cap + d] = f.um[gp];
      sq[2 * cap + d] = f.st[gp];
      sq[3 * cap + d] = f.nw ? f.nw[gp] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __device__ static void load_own(const Consts& c, const Frame& f, const Extra& ex, int p,
                                  float4 pt, Own& o) {
    if (!SUM) {
      float4 um = f.um[p];
      o.u[0] = um.x; o.u[1] = um.y; o.u[2] = um.z;
    }
    if (MODE == DENS_EVOL_RIE) {
      float4 st = f.st[p];
      o.rho = st.x; o.p = st.y;
      float r[3] = {pt.x, pt.y, pt.z};
      g_ext_of<DIM>(c, f, p, r, o.g);
    }
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) {
    a.s = 0.f; a.swf = 0.f; a.sT = 0.f;
    a.su[0] = a.su[1] = a.su[2] = 0.f;
  }
  __device__ static void pair(const Consts& c, const Extra& ex, const Own& o, Acc& a,
                              const float4* sq, int cap, int j, float4 pj, const float (&dr)[3],
                              float d2) {
    const float dist = fsqrt(d2);
    const int tag_j = __float_as_int(pj.w);
    if (SUM || (XTRA && (ex.utilde || ex.wallT))) {
      const float w = kernel_w<KERN>(c, dist);
      if (SUM) a.s += w;
      if (XTRA && (ex.utilde || ex.wallT)) {
        if (tag_j == SPHB200_TAG_FLUID) {
          const float4 uj = sq[cap + j];
          const float4 sj = sq[2 * cap + j];
          a.swf += w;
          a.su[0] += w * uj.x; a.su[1] += w * uj.y; a.su[2] += w * uj.z;
          a.sT += w * sj.z;
        }
      }
    }
    if (MODE == DENS_EVOL_SPH) {
      const float4 uj = sq[cap + j];
      const float gw = kernel_gw<KERN>(c, dist);
      const float id = frcp(dist + c.eps);
      float s = (o.u[0] - uj.x) * (gw * (dr[0] * id)) + (o.u[1] - uj.y) * (gw * (dr[1] * id));
      if (DIM == 3) s += (o.u[2] - uj.z) * (gw * (dr[2] * id));
      a.s += uj.w * s;
    }
    if (MODE == DENS_EVOL_RIE) {
      const float4 uj4 = sq[cap + j], sj = sq[2 * cap + j], nj4 = sq[3 * cap + j];
      const float uj[3] = {uj4.x, uj4.y, uj4.z}, nwj[3] = {nj4.x, nj4.y, nj4.z};
      const float gw = kernel_gw<KERN>(c, dist);
      const float id = frcp(dist + c.eps);
      float e[3] = {dr[0] * id, dr[1] * id, DIM == 3 ? dr[2] * id : 0.f};
      float kg[3] = {gw * e[0], gw * e[1], gw * e[2]};
      const bool is_w = is_wall_tag(tag_j);
      float ne[3] = {-e[0], -e[1], -e[2]}, nn[3] = {-nwj[0], -nwj[1], -nwj[2]};
      float ndr[3] = {-dr[0], -dr[1], -dr[2]};
      const float u_L = is_w ? dot3(o.u, nn, DIM) : dot3(o.u, ne, DIM);
      const float p_L = o.p, rho_L = o.rho;
      const float u_R = is_w ? (-u_L + 2.0f * dot3(uj, nwj, DIM)) : dot3(uj, ne, DIM);
      const float p_R = is_w ? (p_L + rho_L * dot3(o.g, ndr, DIM)) : sj.y;
      const float rho_R = is_w ? eos_rho(c, p_R) : sj.x;
      const float U_avg = (u_L + u_R) / 2.0f;
      const float rho_avg = (rho_L + rho_R) / 2.0f;
      const float U_star = U_avg + fdiv(0.5f * (p_L - p_R), rho_avg * c.c_ref);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const float v_avg = (o.u[k] + uj[k]) / 2.0f;
        const float v_star = U_star * ne[k] + (v_avg - U_avg * ne[k]);
        s += (o.u[k] - v_star) * kg[k];
      }
      a.s += fdiv(2.0f * o.rho * uj4.w, sj.x) * s;
    }
  }
  __device__ static void finish(const Consts& c, const Frame& f, const Extra& ex, int p,
                                const Own&, const Acc& a) {
    const float4 st = f.st[p];
    const int tag = __float_as_int(f.pt[p].w);
    float rho, drhodt = 0.f;
    if (SUM) {
      const float rho_ = f.um[p].w * a.s;
      rho = (tag == SPHB200_TAG_FLUID) ? rho_ : st.x;  // solver.py:795-796
    } else if (MODE == DENS_EVOL_SPH) {
      drhodt = st.x * a.s;           // :28
      rho = st.x + c.dt_s * drhodt;  // :29
    } else {
      drhodt = a.s * ((tag == SPHB200_TAG_FLUID) ? 1.0f : 0.0f);  // :789
      rho = st.x + c.dt_s * drhodt;                               // :790
    }
    const float pnew = eos_p(c, rho);  // :801
    float T = st.z;
    if (XTRA && ex.wallT && (tag == SPHB200_TAG_SOLID_WALL || tag == SPHB200_TAG_MOVING_WALL))
      T = a.sT / (a.swf + c.eps);  // :556-565
    if (ex.finalT) T = T + c.dt_s * st.w;  // :834
    ex.st_out[p] = make_float4(rho, pnew, T, st.w);
    if (!SUM) reinterpret_cast<float*>(&f.du[p])[3] = drhodt;
    if (XTRA && ex.utilde) {
      const float den = a.swf + c.eps;  // :547-552
      f.ut[p] = make_float4(a.su[0] / den, a.su[1] / den, a.su[2] / den, 0.f);
    }
  }
};

// ---------------------------------------------------------------------------
// Delta-SPH density evolution, rho_evol_fn_delta (solver.py:36-103, Marrone et al. 2011), as
// three sweeps over the same neighbour lists:
//   STEP 0 (list builder)  M_i = sum_j (-r_ij) (x) (grad W_ij V_j),  L_i = inv(M_i)      :52-55
//   STEP 1                 G_i = sum_j (rho_j - rho_i) L_i (grad W_ij V_j)               :59-62
//                          H_i = sum_j (rho_i - rho_j) L_j (grad W_ij V_i)               :63-65
//                          (the reference builds L_matj = inv(sum over senders), which is
//                          -L_j for the symmetric edge list; the two signs cancel)
//   STEP 2                 psi_ij = 2 (rho_j - rho_i)(-r_ij)/(d+EPS)^2 - G_i - H_j       :77-79
//                          drhodt = rho_i sum_j (u_i-u_j).grad W_ij V_j
//                                   + c_ref delta h sum_j psi_ij.grad W_ij V_j [j fluid]  :80-101
//                          rho += dt drhodt, p = eos(rho)                                :102, :801
// All three read the INCOMING rho; only STEP 2 writes (into the other frame's st).
template <int DIM, int KERN, int STEP>
struct PhysDelta {
  static constexpr int MINB = STEP == 0 ? 2 : 1;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;  // true: most particles are inactive (tile skip, sweep.cuh)
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = true;
  // Staged records, packed so that the stencil of a tile fits ONE staging group (a tile whose
  // stencil does not fit gets no shared neighbour lists and every sweep searches on its own):
  //   STEP 0           (x, y, z, V)                                                  16 B
  //   STEP 1  3D       (x, y, z, V) (L00 L01 L02 rho) (L10 L11 L12 L20) (L21 L22 - -) 64 B
  //           2D       (x, y, rho, V) (L00 L01 L10 L11)                               32 B
  //   STEP 2           (x, y, z, tag) (u, V) (H, rho)                                 48 B
  struct Own {
    float rho, V;
    float u[3], G[3];
    float L[9];
  };
  struct Acc {
    float m[9];  // STEP 0: M;  STEP 1: G (0..2), H (3..5);  STEP 2: diff (0), cont (1)
  };
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;  // arrays the duo sweeps stage by bulk copy (sweep2.cuh)
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
#pragma unroll
    for (int k = 0; k < (STEP == 0 ? 9 : (STEP == 1 ? 6 : 2)); ++k) f(a.m[k]);
  }
  __host__ __device__ static int nq() { return STEP == 0 ? 1 : (STEP == 1 ? (DIM == 3 ? 4 : 2) : 3); }
  __device__ static void stage(const Consts&, const Frame& f, const Extra&, int gp, float4* sq,
                               int cap, int d) {
    const float4 pt = f.pt[gp], um = f.um[gp], st = f.st[gp];
    const float V = um.w / st.x;  // V_j = m_j / rho_j (:44-45)
    if (STEP == 0) sq[d] = make_float4(pt.x, pt.y, pt.z, V);
    if (STEP == 1) {
      const float4 l0 = f.dl0[gp];
      if (DIM == 3) {
        const float4 l1 = f.dl1[gp], l2 = f.dl2[gp];
        sq[d] = make_float4(pt.x, pt.y, pt.z, V);
        sq[cap + d] = make_float4(l0.x, l0.y, l0.z, st.x);
        sq[2 * cap + d] = make_float4(l1.x, l1.y, l1.z, l2.x);
        sq[3 * cap + d] = make_float4(l2.y, l2.z, 0.f, 0.f);
      } else {
        sq[d] = make_float4(pt.x, pt.y, st.x, V);  // the 2D sweeps never read the z slot
        sq[cap + d] = l0;
      }
    }
    if (STEP == 2) {
      const float4 h = f.dg1[gp];
      sq[d] = pt;
      sq[cap + d] = make_float4(um.x, um.y, um.z, V);
      sq[2 * cap + d] = make_float4(h.x, h.y, h.z, st.x);
    }
  }
  __device__ static void load_own(const Consts&, const Frame& f, const Extra&, int p, float4,
                                  Own& o) {
    const float4 um = f.um[p], st = f.st[p];
    o.rho = st.x;
    o.V = um.w / st.x;
    o.u[0] = um.x; o.u[1] = um.y; o.u[2] = um.z;
    if (STEP == 1) {
      const float4 a = f.dl0[p];
      if (DIM == 3) {
        const float4 b = f.dl1[p], cc = f.dl2[p];
        o.L[0] = a.x; o.L[1] = a.y; o.L[2] = a.z;
        o.L[3] = b.x; o.L[4] = b.y; o.L[5] = b.z;
        o.L[6] = cc.x; o.L[7] = cc.y; o.L[8] = cc.z;
      } else {
        o.L[0] = a.x; o.L[1] = a.y; o.L[2] = a.z; o.L[3] = a.w;
      }
    }
    if (STEP == 2) {
      const float4 g = f.dg0[p];
      o.G[0] = g.x; o.G[1] = g.y; o.G[2] = g.z;
    }
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) {
#pragma unroll
    for (int k = 0; k < 9; ++k) a.m[k] = 0.f;
  }
  // y = L x for a row-major DIM x DIM matrix
  __device__ static void matvec(const float* L, const float (&x)[3], float (&y)[3]) {
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      float s = L[a * DIM] * x[0] + L[a * DIM + 1] * x[1];
      if (DIM == 3) s += L[a * DIM + 2] * x[2];
      y[a] = s;
    }
    if (DIM == 2) y[2] = 0.f;
  }
  __device__ static void pair(const Consts& c, const Extra&, const Own& o, Acc& a,
                              const float4* sq, int cap, int j, float4 pj, const float (&dr)[3],
                              float d2) {
    const float dist = fsqrt(d2);
    const float gw = kernel_gw<KERN>(c, dist);
    const float id = frcp(dist + c.eps);
    float kg[3] = {gw * (dr[0] * id), gw * (dr[1] * id), DIM == 3 ? gw * (dr[2] * id) : 0.f};  // :42-43
    if (STEP == 0) {
      const float V_j = pj.w;
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int s = 0; s < DIM; ++s) a.m[r * DIM + s] += (-dr[r]) * (kg[s] * V_j);  // :49-50
    }
    if (STEP == 1) {
      const float V_j = pj.w;
      float Lj[9], rho_j;
      const float4 l0 = sq[cap + j];
      if (DIM == 3) {
        const float4 l1 = sq[2 * cap + j], l2 = sq[3 * cap + j];
        Lj[0] = l0.x; Lj[1] = l0.y; Lj[2] = l0.z;
        rho_j = l0.w;
        Lj[3] = l1.x; Lj[4] = l1.y; Lj[5] = l1.z;
        Lj[6] = l1.w; Lj[7] = l2.x; Lj[8] = l2.y;
      } else {
        Lj[0] = l0.x; Lj[1] = l0.y; Lj[2] = l0.z; Lj[3] = l0.w;
        rho_j = pj.z;
      }
      const float xj[3] = {kg[0] * V_j, kg[1] * V_j, kg[2] * V_j};
      const float xi[3] = {kg[0] * o.V, kg[1] * o.V, kg[2] * o.V};
      float gi[3], hj[3];
      matvec(o.L, xj, gi);
      matvec(Lj, xi, hj);
      const float dji = rho_j - o.rho, dij = o.rho - rho_j;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        a.m[k] += dji * gi[k];      // :59-62
        a.m[3 + k] += dij * hj[k];  // :63-65
      }
    }
    if (STEP == 2) {
      const float4 uj = sq[cap + j], hj = sq[2 * cap + j];
      const float V_j = uj.w, rho_j = hj.w;
      const float idd = frcp((dist + c.eps) * (dist + c.eps));
      const float two_d = 2.0f * (rho_j - o.rho);
      const float H[3] = {hj.x, hj.y, hj.z};
      const float du[3] = {o.u[0] - uj.x, o.u[1] - uj.y, o.u[2] - uj.z};
      float psi_kg = 0.f, du_kg = 0.f;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const float psi = (two_d * (-dr[k])) * idd - o.G[k] - H[k];  // :77-78
        psi_kg += psi * kg[k];
        du_kg += du[k] * kg[k];
      }
      const float fm = (__float_as_int(pj.w) == SPHB200_TAG_FLUID) ? 1.0f : 0.0f;
      a.m[0] += psi_kg * V_j * fm;  // :79
      a.m[1] += du_kg * V_j;        // :95-98
    }
  }
  __device__ static void finish(const Consts& c, const Frame& f, const Extra& ex, int p,
                                const Own& o, const Acc& a) {
    if (STEP == 0) {
      float L[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* m = a.m;
      if (DIM == 2) {
        const float idet = 1.0f / (m[0] * m[3] - m[1] * m[2]);
        L[0] = m[3] * idet; L[1] = -m[1] * idet; L[2] = -m[2] * idet; L[3] = m[0] * idet;
        f.dl0[p] = make_float4(L[0], L[1], L[2], L[3]);
      } else {
        const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8],
                    c02 = m[3] * m[7] - m[4] * m[6];
        const float idet = 1.0f / (m[0] * c00 + m[1] * c01 + m[2] * c02);
        f.dl0[p] = make_float4(c00 * idet, (m[2] * m[7] - m[1] * m[8]) * idet,
                                  (m[1] * m[5] - m[2] * m[4]) * idet, 0.f);
        f.dl1[p] = make_float4(c01 * idet, (m[0] * m[8] - m[2] * m[6]) * idet,
                                      (m[2] * m[3] - m[0] * m[5]) * idet, 0.f);
        f.dl2[p] = make_float4(c02 * idet, (m[1] * m[6] - m[0] * m[7]) * idet,
                                      (m[0] * m[4] - m[1] * m[3]) * idet, 0.f);
      }
    }
    if (STEP == 1) {
      f.dg0[p] = make_float4(a.m[0], a.m[1], a.m[2], 0.f);
      f.dg1[p] = make_float4(a.m[3], a.m[4], a.m[5], 0.f);
    }
    if (STEP == 2) {
      const float4 st = f.st[p];
      const float drhodt = st.x * a.m[1] + c.delta_rho * a.m[0];  // :99-101
      const float rho = st.x + c.dt_s * drhodt;                   // :102
      float T = st.z;
      if (ex.finalT) T = T + c.dt_s * st.w;  // :834
      ex.st_out[p] = make_float4(rho, eos_p(c, rho), T, st.w);
      reinterpret_cast<float*>(&f.du[p])[3] = drhodt;
    }
  }
};

// ---------------------------------------------------------------------------
template <int DIM, int KERN>
struct PhysRenorm {
  static constexpr int MINB = 2;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;  // true: most particles are inactive (tile skip, sweep.cuh)
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = true;
  struct Own {};
  struct Acc {
    float num, den;
  };
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;  // arrays the duo sweeps stage by bulk copy (sweep2.cuh)
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
    f(a.num); f(a.den);
  }
  __device__ static void stage(const Consts&, const Frame& f, const Extra&, int gp, float4* sq,
                               int cap, int d) {
    sq[d] = f.pt[gp];
    const float m = f.um[gp].w, rho = f.st[gp].x;
    sq[cap + d] = make_float4(m, m / rho, 0.f, 0.f);
  }
  __device__ static void load_own(const Consts&, const Frame&, const Extra&, int, float4, Own&) {}
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) { a.num = a.den = 0.f; }
  __device__ static void pair(const Consts& c, const Extra&, const Own&, Acc& a, const float4* sq,
                              int cap, int j, float4, const float (&)[3], float d2) {
    const float w = kernel_w<KERN>(c, fsqrt(d2));
    const float4 mj = sq[cap + j];
    a.num += mj.x * w;
    a.den += mj.y * w;
  }
  __device__ static void finish(const Consts& c, const Frame& f, const Extra& ex, int p,
                                const Own&, const Acc& a) {
    const float4 st = f.st[p];
    const float den = a.den > 1.0f ? 1.0f : a.den;
    const float rho = a.num / den;
    ex.st_out[p] = make_float4(rho, eos_p(c, rho), st.z, st.w);
  }
};

// ---------------------------------------------------------------------------
template <int DIM, int KERN>
struct PhysWall {
  static constexpr int MINB = 1;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = true;   // true: most particles are inactive (tile skip, sweep.cuh)
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = true;
  struct Own {
    int tag;
  };
  struct Acc {
    float sw, sp, sT;
    float su[3], sv[3], srr[3];
  };
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;  // arrays the duo sweeps stage by bulk copy (sweep2.cuh)
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
    f(a.sw); f(a.sp); f(a.sT);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f(a.su[k]); f(a.sv[k]); f(a.srr[k]);
    }
  }
  __device__ static void stage(const Consts&, const Frame& f, const Extra&, int gp, float4* sq,
                               int cap, int d) {
    sq[d] = f.pt[gp];
    sq[cap + d] = f.um[gp];
    sq[2 * cap + d] = f.vv[gp];
    sq[3 * cap + d] = f.st[gp];
  }
  __device__ static void load_own(const Consts&, const Frame&, const Extra&, int, float4 pt,
                                  Own& o) {
    o.tag = __float_as_int(pt.w);
  }
  // only wall particles need the Shepard sums (everything else is discarded by the
  // jnp.where(mask_bc, ...) of solver.py:461,485,514)
  __device__ static bool active(const Consts&, const Own& o) { return is_wall_tag(o.tag); }
  __device__ static void init(Acc& a) {
    a.sw = a.sp = a.sT = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) a.su[k] = a.sv[k] = a.srr[k] = 0.f;
  }
  __device__ static void pair(const Consts& c, const Extra&, const Own&, Acc& a, const float4* sq,
                              int cap, int j, float4 pj, const float (&dr)[3], float d2) {
    if (__float_as_int(pj.w) != SPHB200_TAG_FLUID) return;  // w * mask_j_s_fluid == 0
    const float w = kernel_w<KERN>(c, fsqrt(d2));
    const float4 uj = sq[cap + j], vj = sq[2 * cap + j], sj = sq[3 * cap + j];
    a.sw += w;
    a.su[0] += w * uj.x; a.su[1] += w * uj.y; a.su[2] += w * uj.z;
    a.sv[0] += w * vj.x; a.sv[1] += w * vj.y; a.sv[2] += w * vj.z;
    a.sp += w * sj.y;
    const float rw = sj.x * w;
    a.srr[0] += rw * dr[0]; a.srr[1] += rw * dr[1]; a.srr[2] += rw * dr[2];
    a.sT += w * sj.z;
  }
  __device__ static void finish(const Consts& c, const Frame& f, const Extra& ex, int p,
                                const Own& o, const Acc& a) {
    const bool wall = is_wall_tag(o.tag);
    const float4 st = f.st[p];
    float p_new = st.y, T = st.z;
    if (wall) {
      const float4 pt = f.pt[p], um = f.um[p], vv = f.vv[p];
      const float u[3] = {um.x, um.y, um.z}, v[3] = {vv.x, vv.y, vv.z};
      float r[3] = {pt.x, pt.y, pt.z}, g[3];
      g_ext_of<DIM>(c, f, p, r, g);
      const float den = a.sw + c.eps;
      float uw[3], vw[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uw[k] = a.su[k] / den;
        vw[k] = a.sv[k] / den;
      }
      if (ex.free_slip) {  // :464-486, wall_inner_normals = -nw
        float4 n = f.nw ? f.nw[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        float win[3] = {-n.x, -n.y, -n.z};
        float su = dot3(uw, win, DIM), sv = dot3(vw, win, DIM);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          uw[k] = win[k] * su;
          vw[k] = win[k] * sv;
        }
      }
      f.um[p] = make_float4(2.0f * u[0] - uw[0], 2.0f * u[1] - uw[1],
                            DIM == 3 ? 2.0f * u[2] - uw[2] : 0.f, um.w);
      f.vv[p] = make_float4(2.0f * v[0] - vw[0], 2.0f * v[1] - vw[1],
                            DIM == 3 ? 2.0f * v[2] - vw[2] : 0.f, vv.w);
      const float p_ext = dot3(g, a.srr, DIM);  // :509-511
      p_new = (a.sp + p_ext) / den;             // :513
      if (ex.heat && (o.tag == SPHB200_TAG_SOLID_WALL || o.tag == SPHB200_TAG_MOVING_WALL))
        T = a.sT / den;  // :518-526
    }
    const float rho = eos_rho(c, p_new);  // :516, every particle
    if (ex.heat) T = T + c.dt_s * st.w;   // :834
    ex.st_out[p] = make_float4(rho, p_new, T, st.w);
  }
};

// ---------------------------------------------------------------------------
// Force sweep.  Staged per particle (GENERIC / RIE): quads 0 (x,y,z,tag) 1 (u, rho)
// 2 (p, eta, mass, (m/rho)^2) [q_v] (rho/2 (v - u), 0)  [q_h] (T, kappa, 0, 0)  [q_nw] (nw, 0)
// [q_ut] (u_tilde, 0).  The two headline variants stage a COMPACT record -- the sweep is bound
// by shared-memory wavefronts of these random 16-byte gathers, so every byte counts:
//   FORCE_TVF   quads 0 (x,y,z,rho) 1 (u, p) 2 (rho/2 (v - u), (m/rho)^2) + float column eta  (52 B)
//   FORCE_PLAIN quads 0 (x,y,z,rho) 1 (u, p)     + float2 column (eta, (m/rho)^2)            (40 B)
// FEAT: compile-time specialisation of the two headline variants; GENERIC reads
// the switches from Extra at run time.
//
// Standard acceleration (solver.py:221-256) per pair, with c = ((m_i/rho_i)^2 + (m_j/rho_j)^2)/m_i
// * grad_w / (d + EPS):
//     a_i += c (-p_ij r + 1/2 (A_i + A_j) r + eta_ij u_ij),   A = rho u (x) (v - u)
// is accumulated as  a_i += -(c p_ij) r + (c eta_ij) u_ij + (c (rho_j/2 (v_j-u_j)).r) u_j  and the
// A_i term is factored out of the sum: 1/2 rho_i u_i ((v_i-u_i) . sum_j c r); sum_j c r is the
// transport-velocity sum the sweep keeps anyway (dvdt = p_bg sum_j c r, solver.py:199-213).
// FORCE_TVF_U: FORCE_TVF under SPHB200_HINT_UNIFORM_ETA (duo sweeps only): eta is not staged, eta_ij
// is a constant of the duo
enum { FORCE_PLAIN = 0, FORCE_TVF = 1, FORCE_GENERIC = 2, FORCE_TVF_U = 3 };

template <int DIM, int KERN, int SOLVER, int FEAT>
struct PhysForce {
  static constexpr int MINB = 1;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;  // true: most particles are inactive (tile skip, sweep.cuh)
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = true;
  struct Own {
    float u[3], dvu[3], g[3];
    float rho, p, eta, eta2, inv_m, V2, T, kappa, Cp;
    int tag;
  };
  struct Acc {
    float a[3], tv[3], av[3];
    float dT;
  };
  static constexpr bool COMPACT = SOLVER == SPHB200_SOLVER_SPH && FEAT != FORCE_GENERIC;
  static constexpr bool TVF = FEAT == FORCE_TVF || FEAT == FORCE_TVF_U;
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      f(a.a[k]); f(a.tv[k]);
      if (FEAT == FORCE_GENERIC) f(a.av[k]);
    }
    if (FEAT == FORCE_GENERIC) f(a.dT);
  }
  // The two headline variants: packed pair body of the list consumers (sweep.cuh, Pair2) -- the
  // arithmetic of pair() below for two neighbours in the two halves of packed registers.  A lane
  // without a pair (v false) computes on staged slot 0 with its coefficient c forced to zero.
  static constexpr bool HAS_PAIR2 = COMPACT;
  struct Acc2 {
    F2 a[3], tv[3];
  };
  __device__ static void init2(Acc2& a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a.a[k] = a.tv[k] = f2(0.0f);
  }
  __device__ static void pair2(const Consts& c, const Extra&, const Own& o, Acc2& a,
                               const float4* sq, int cap, int j0, int j1, float4 p0, float4 p1,
                               const F2 (&dr)[3], F2 d2, bool v0, bool v1) {
    const float4 qa = sq[cap + j0], qb = sq[cap + j1];
    const F2 uj[3] = {f2(qa.x, qb.x), f2(qa.y, qb.y), f2(qa.z, qb.z)};
    const F2 rho_j = f2(p0.w, p1.w), p_j = f2(qa.w, qb.w);
    F2 eta_j, V2_j, hdv[3];
    if (FEAT == FORCE_TVF) {
      const float4 ra = sq[2 * cap + j0], rb = sq[2 * cap + j1];
      hdv[0] = f2(ra.x, rb.x); hdv[1] = f2(ra.y, rb.y); hdv[2] = f2(ra.z, rb.z);
      V2_j = f2(ra.w, rb.w);
      const float* ec = reinterpret_cast<const float*>(sq + 3 * cap);
      eta_j = f2(ec[j0], ec[j1]);
    } else {
      const float2* ev = reinterpret_cast<const float2*>(sq + 2 * cap);
      const float2 ea = ev[j0], eb = ev[j1];
      eta_j = f2(ea.x, eb.x);
      V2_j = f2(ea.y, eb.y);
    }
    const F2 dist = fsqrt2(d2);
    const F2 gw = kernel_gw2<KERN>(c, dist);
    const F2 id = frcp2(add2(dist, f2(c.eps)));
    const F2 wv = mul2(add2(f2(o.V2), V2_j), f2(o.inv_m));                  // :205 / :247
    const F2 cc = sel2(v0, v1, mul2(mul2(wv, gw), id));                     // :206 / :248
    const F2 eta_ij = mul2(mul2(f2(o.eta2), eta_j),
                           frcp2(add2(add2(f2(o.eta), eta_j), f2(c.eps))));  // :243
    const F2 p_ij = mul2(fma2(rho_j, f2(o.p), mul2(f2(o.rho), p_j)),
                         frcp2(add2(f2(o.rho), rho_j)));                     // :244
    const F2 ncp = mul2(cc, neg2(p_ij)), ce = mul2(cc, eta_ij);
    F2 cj = f2(0.0f);
    if (FEAT == FORCE_TVF) {  // c (A_j r)_k = cj u_j[k]   (:250-251)
      F2 d = mul2(hdv[0], dr[0]);
      d = fma2(hdv[1], dr[1], d);
      if (DIM == 3) d = fma2(hdv[2], dr[2], d);
      cj = mul2(cc, d);
    }
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      fma2_into(a.tv[k], cc, dr[k]);                      // sum_j c r  (:199-213, :912-921)
      fma2_into(a.a[k], ncp, dr[k]);                      // -c p_ij r
      fma2_into(a.a[k], ce, sub2(f2(o.u[k]), uj[k]));     // c eta_ij u_ij
      if (FEAT == FORCE_TVF) fma2_into(a.a[k], cj, uj[k]);
    }
  }
  __device__ static void fold(const Acc2& a2, Acc& a) {
    init(a);
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      a.a[k] = lo(a2.a[k]) + hi(a2.a[k]);
      a.tv[k] = lo(a2.tv[k]) + hi(a2.tv[k]);
    }
  }
  // The same arithmetic for the two particles of a duo against ONE neighbour (sweep2.cuh): own
  // values packed once per duo, the neighbour's values broadcast scalars.
  static constexpr bool HAS_DUO = COMPACT;
  static constexpr int DUO_MINB = 1;
  // The duo sweep stages the compact record from per-slot copies of it in HBM (k_force_rec,
  // Extra::rec*): three quads by bulk copy, and for FORCE_TVF the eta column through registers.
  //   quad 2 of FORCE_PLAIN there: (eta, (m/rho)^2, 0, 0)
  static constexpr int DUO_COPIES = COMPACT ? 3 : 0;
  static constexpr bool DUO_REST = COMPACT && FEAT == FORCE_TVF;  // (FORCE_TVF_U: no eta column)
  __device__ static void duo_sources(const Frame&, const Extra& ex, const float4* (&a)[COMPACT ? 3 : 1]) {
    if (COMPACT) {
      a[0] = ex.rec0; a[1] = ex.rec1; a[2] = ex.rec2;
    }
  }
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra& ex, int gp, float4* sq,
                                    int cap, int d) {
    cp_async4(reinterpret_cast<float*>(sq + 3 * cap) + d, ex.rec_e + gp);
  }
  // the record of one slot, as stage() builds it
  __device__ static void make_record(const Frame& f, const Extra& ex, int gp) {
    const float4 pt = f.pt[gp], um = f.um[gp], st = f.st[gp], vv = f.vv[gp];
    const float vol = um.w / st.x;
    ex.rec0[gp] = make_float4(pt.x, pt.y, pt.z, st.x);
    ex.rec1[gp] = make_float4(um.x, um.y, um.z, st.y);
    if (TVF) {
      const float hr = 0.5f * st.x;
      ex.rec2[gp] = make_float4(hr * (vv.x - um.x), hr * (vv.y - um.y), hr * (vv.z - um.z), vol * vol);
      if (FEAT == FORCE_TVF) ex.rec_e[gp] = vv.w;
      else if (vv.w != ex.eta_ref[0].w) atomicOr(ex.err_word, SPHB200_ERR_HINT);  // the promise
    } else {
      ex.rec2[gp] = make_float4(vv.w, vol * vol, 0.f, 0.f);
    }
  }
  struct OwnD {
    F2 u[3];
    // eta_e = eta + EPS, V2m = (m/rho)^2 / m; FORCE_TVF_U: eta2 holds eta_ij itself
    F2 rho, p, eta_e, eta2, inv_m, V2m;
  };
  struct AccD {
    F2 a[3], tv[3];
  };
  __device__ static void load_duo(const Own& o0, const Own& o1, OwnD& d) {
#pragma unroll
    for (int k = 0; k < 3; ++k) d.u[k] = f2(o0.u[k], o1.u[k]);
    d.rho = f2(o0.rho, o1.rho);
    d.p = f2(o0.p, o1.p);
    d.eta_e = f2(o0.eta + 1.1920928955078125e-07f, o1.eta + 1.1920928955078125e-07f);
    d.eta2 = f2(o0.eta2, o1.eta2);
    if (FEAT == FORCE_TVF_U)  // every eta_j equals the own eta: the pair value, same operations
      d.eta2 = mul2(mul2(d.eta2, f2(o0.eta, o1.eta)), frcp2(add2(d.eta_e, f2(o0.eta, o1.eta))));
    d.inv_m = f2(o0.inv_m, o1.inv_m);
    d.V2m = f2(o0.V2 * o0.inv_m, o1.V2 * o1.inv_m);
  }
  __device__ static void init_duo(AccD& a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a.a[k] = a.tv[k] = f2(0.0f);
  }
  __device__ static void pair_duo(const Consts& c, const Extra&, const OwnD& o, AccD& a,
                                  const float4* sq, int cap, int j, float4 pj, const F2 (&dr)[3],
                                  F2 d2, bool v0, bool v1) {
    const float4 q = sq[cap + j];  // (u_j, p_j)
    const float rho_j = pj.w, p_j = q.w;
    float eta_j = 0.f, V2_j, hx = 0.f, hy = 0.f, hz = 0.f;
    if (TVF) {
      const float4 r = sq[2 * cap + j];
      hx = r.x; hy = r.y; hz = r.z;
      V2_j = r.w;
      if (FEAT == FORCE_TVF) eta_j = reinterpret_cast<const float*>(sq + 3 * cap)[j];
    } else {
      const float4 ev = sq[2 * cap + j];  // duo layout of FORCE_PLAIN: (eta, (m/rho)^2, -, -)
      eta_j = ev.x;
      V2_j = ev.y;
    }
    // (three operations fewer per trip than the literal forms, same values to the last bit or
    // two: ((m/rho)_i^2 + (m/rho)_j^2) / m_i as one FMA, EPS folded into the own viscosity,
    // sigma / h folded into the polynomial's coefficients)
    const F2 dist = fsqrt2(d2);
    const F2 gw = kernel_gw2_folded<KERN>(c, dist);
    const F2 id = frcp2(add2(dist, f2(c.eps)));
    const F2 wv = fma2(f2(V2_j), o.inv_m, o.V2m);                           // :205 / :247
    const F2 cc = sel2(v0, v1, mul2(mul2(wv, gw), id));                     // :206 / :248
    const F2 eta_ij = FEAT == FORCE_TVF_U ? o.eta2
                                          : mul2(mul2(o.eta2, f2(eta_j)),
                                                 frcp2(add2(o.eta_e, f2(eta_j))));  // :243
    const F2 p_ij = mul2(fma2(f2(rho_j), o.p, mul2(o.rho, f2(p_j))),
                         frcp2(add2(o.rho, f2(rho_j))));                     // :244
    const F2 ncp = mul2(cc, neg2(p_ij)), ce = mul2(cc, eta_ij);
    F2 cj = f2(0.0f);
    if (TVF) {  // c (A_j r)_k = cj u_j[k]   (:250-251)
      F2 d = mul2(f2(hx), dr[0]);
      d = fma2(f2(hy), dr[1], d);
      if (DIM == 3) d = fma2(f2(hz), dr[2], d);
      cj = mul2(cc, d);
    }
    const float uj[3] = {q.x, q.y, q.z};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      fma2_into(a.tv[k], cc, dr[k]);                       // sum_j c r  (:199-213, :912-921)
      fma2_into(a.a[k], ncp, dr[k]);                       // -c p_ij r
      fma2_into(a.a[k], ce, sub2(o.u[k], f2(uj[k])));      // c eta_ij u_ij
      if (TVF) fma2_into(a.a[k], cj, f2(uj[k]));
    }
  }
  __device__ static void fold_duo(const AccD& d, Acc& a0, Acc& a1) {
    init(a0);
    init(a1);
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      a0.a[k] = lo(d.a[k]); a1.a[k] = hi(d.a[k]);
      a0.tv[k] = lo(d.tv[k]); a1.tv[k] = hi(d.tv[k]);
    }
  }
  template <class F>
  __device__ static void each_acc_duo(AccD& a, F f) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      f(a.a[k]); f(a.tv[k]);
    }
  }
  static constexpr int CQ = FEAT == FORCE_TVF ? 3 : 2;  // quads of the compact record
  // bytes staged per particle (engine.cu sizes the staging buffer with it)
  __host__ __device__ static int stage_bytes(int nq_generic) {
    return COMPACT ? (FEAT == FORCE_TVF ? 52 : 40) : 16 * nq_generic;
  }
  __device__ static int qv(const Extra& ex) { return ex.q_v; }
  __device__ static void stage(const Consts&, const Frame& f, const Extra& ex, int gp, float4* sq,
                               int cap, int d) {
    const float4 pt = f.pt[gp], um = f.um[gp], st = f.st[gp], vv = f.vv[gp];
    const float vol = um.w / st.x;
    if (COMPACT) {
      sq[d] = make_float4(pt.x, pt.y, pt.z, st.x);
      sq[cap + d] = make_float4(um.x, um.y, um.z, st.y);
      if (FEAT == FORCE_TVF) {
        const float hr = 0.5f * st.x;
        sq[2 * cap + d] = make_float4(hr * (vv.x - um.x), hr * (vv.y - um.y), hr * (vv.z - um.z),
                                      vol * vol);
        reinterpret_cast<float*>(sq + 3 * cap)[d] = vv.w;
      } else {
        reinterpret_cast<float2*>(sq + 2 * cap)[d] = make_float2(vv.w, vol * vol);
      }
      return;
    }
    sq[d] = pt;
    sq[cap + d] = make_float4(um.x, um.y, um.z, st.x);
    sq[2 * cap + d] = make_float4(st.y, vv.w, um.w, vol * vol);
    if (qv(ex) >= 0) {
      const float hr = 0.5f * st.x;
      sq[qv(ex) * cap + d] = make_float4(hr * (vv.x - um.x), hr * (vv.y - um.y), hr * (vv.z - um.z), 0.f);
    }
    if (FEAT == FORCE_GENERIC) {
      if (ex.q_h >= 0) sq[ex.q_h * cap + d] = make_float4(st.z, f.kc[gp].x, 0.f, 0.f);
      if (ex.q_nw >= 0) sq[ex.q_nw * cap + d] = f.nw ? f.nw[gp] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (ex.q_ut >= 0) sq[ex.q_ut * cap + d] = f.ut[gp];
    }
  }
  __device__ static void load_own(const Consts& c, const Frame& f, const Extra& ex, int p,
                                  float4 pt, Own& o) {
    const float4 um = f.um[p], st = f.st[p], vv = f.vv[p];
    o.u[0] = um.x; o.u[1] = um.y; o.u[2] = um.z;
    o.dvu[0] = vv.x - um.x; o.dvu[1] = vv.y - um.y; o.dvu[2] = vv.z - um.z;
    o.eta = vv.w;
    o.eta2 = 2.0f * vv.w;
    o.rho = st.x; o.p = st.y; o.T = st.z;
    const float vol = um.w / st.x;
    o.V2 = vol * vol;
    o.inv_m = 1.0f / um.w;
    o.tag = __float_as_int(pt.w);
    o.kappa = 0.f; o.Cp = 1.f;
    o.g[0] = o.g[1] = o.g[2] = 0.f;
    if (FEAT == FORCE_GENERIC) {
      if (ex.heat) {
        float2 kc = f.kc[p];
        o.kappa = kc.x; o.Cp = kc.y;
      }
      if (SOLVER == SPHB200_SOLVER_RIE) {
        float r[3] = {pt.x, pt.y, pt.z};
        g_ext_of<DIM>(c, f, p, r, o.g);
      }
    }
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a.a[k] = a.tv[k] = a.av[k] = 0.f;
    a.dT = 0.f;
  }
  __device__ static void pair(const Consts& c, const Extra& ex, const Own& o, Acc& a,
                              const float4* sq, int cap, int j, float4 pj, const float (&dr)[3],
                              float d2) {
    const float4 q1 = sq[cap + j];
    const float uj[3] = {q1.x, q1.y, q1.z};
    float rho_j, p_j, eta_j, m_j = 0.f, V2_j;
    float hdv[3] = {0.f, 0.f, 0.f};  // rho_j / 2 (v_j - u_j)
    bool have_dv = false;
    int tag_j = SPHB200_TAG_FLUID;
    if (COMPACT) {
      rho_j = pj.w;
      p_j = q1.w;
      if (FEAT == FORCE_TVF) {
        const float4 q2 = sq[2 * cap + j];
        hdv[0] = q2.x; hdv[1] = q2.y; hdv[2] = q2.z;
        V2_j = q2.w;
        eta_j = reinterpret_cast<const float*>(sq + 3 * cap)[j];
        have_dv = true;
      } else {
        const float2 ev = reinterpret_cast<const float2*>(sq + 2 * cap)[j];
        eta_j = ev.x;
        V2_j = ev.y;
      }
    } else {
      const float4 q2 = sq[2 * cap + j];
      rho_j = q1.w; p_j = q2.x; eta_j = q2.y; m_j = q2.z; V2_j = q2.w;
      tag_j = __float_as_int(pj.w);
      if (SOLVER == SPHB200_SOLVER_SPH && qv(ex) >= 0) {
        const float4 dj = sq[qv(ex) * cap + j];
        hdv[0] = dj.x; hdv[1] = dj.y; hdv[2] = dj.z;
        have_dv = true;
      }
    }
    const float dist = fsqrt(d2);
    const float gw = kernel_gw<KERN>(c, dist);
    const float id = frcp(dist + c.eps);
    const float wv = (o.V2 + V2_j) * o.inv_m;                        // :205 / :247
    const float cc = wv * gw * id;                                   // :206 / :248
    const float eta_ij = fdiv(o.eta2 * eta_j, o.eta + eta_j + c.eps);  // :243
    // sum_j c r: the transport-velocity acceleration / p_bg, always computed (:912-921)
#pragma unroll
    for (int k = 0; k < DIM; ++k) a.tv[k] = fmaf(cc, dr[k], a.tv[k]);

    if (SOLVER == SPHB200_SOLVER_SPH) {
      const float p_ij = fdiv(rho_j * o.p + o.rho * p_j, o.rho + rho_j);  // :244
      const float cp = cc * p_ij, ce = cc * eta_ij;
      float cj = 0.f;
      if (have_dv) cj = cc * dot3(hdv, dr, DIM);  // c (A_j r)_k = cj u_j[k]   (:250-251)
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        float t = fmaf(-cp, dr[k], a.a[k]);       // -c p_ij r
        t = fmaf(ce, o.u[k] - uj[k], t);          // c eta_ij u_ij
        a.a[k] = have_dv ? fmaf(cj, uj[k], t) : t;
      }
    } else {
      float e[3] = {dr[0] * id, dr[1] * id, DIM == 3 ? dr[2] * id : 0.f};
      float ne[3] = {-e[0], -e[1], -e[2]};
      const bool is_w = is_wall_tag(tag_j);
      float nwj[3] = {0.f, 0.f, 0.f}, ud[3] = {uj[0], uj[1], uj[2]};
      if (is_w && ex.q_nw >= 0) {
        const float4 n4 = sq[ex.q_nw * cap + j];
        nwj[0] = n4.x; nwj[1] = n4.y; nwj[2] = n4.z;
      }
      if (is_w && ex.q_ut >= 0) {
        const float4 t4 = sq[ex.q_ut * cap + j];
        ud[0] = 2.0f * uj[0] - t4.x; ud[1] = 2.0f * uj[1] - t4.y; ud[2] = 2.0f * uj[2] - t4.z;
      }
      float nn[3] = {-nwj[0], -nwj[1], -nwj[2]}, ndr[3] = {-dr[0], -dr[1], -dr[2]};
      const float u_L = is_w ? dot3(o.u, nn, DIM) : dot3(o.u, ne, DIM);  // :346-350
      const float p_L = o.p, rho_L = o.rho;
      const float u_R = is_w ? (-u_L + 2.0f * dot3(uj, nwj, DIM)) : dot3(uj, ne, DIM);
      const float p_R = is_w ? (p_L + rho_L * dot3(o.g, ndr, DIM)) : p_j;
      const float rho_R = is_w ? eos_rho(c, p_R) : rho_j;
      const float P_avg = (p_L + p_R) / 2.0f;
      const float rho_avg = (rho_L + rho_R) / 2.0f;
      float beta = c.c_ref;  // :574-588
      if (c.use_lim) beta = fminf(c.eta_lim * fmaxf(u_L - u_R, 0.0f), c.c_ref);
      const float P_star = P_avg + 0.5f * rho_avg * (u_L - u_R) * beta;  // :371
      const float irr = frcp(o.rho * rho_j);
      const float c9 = -2.0f * m_j * (P_star * irr);  // :374
      const float c6 = 2.0f * m_j * eta_ij * irr;     // :380-386
      float mask = 1.0f;
      if (ex.bc_trick && ex.free_slip) mask = (o.tag == SPHB200_TAG_FLUID) ? 1.0f : 0.0f;
      const float gm = gw * mask;
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        const float vij = is_w ? (o.u[k] - ud[k]) : (o.u[k] - uj[k]);
        a.a[k] += c9 * (gw * e[k]) + (c6 * vij * id) * gm;
      }
    }
    if (FEAT == FORCE_GENERIC) {
      if (SOLVER == SPHB200_SOLVER_SPH && ex.delta) {  // acceleration_delta_fn, solver.py:297-311
        if (tag_j == SPHB200_TAG_FLUID) {
          const float du[3] = {o.u[0] - uj[0], o.u[1] - uj[1], o.u[2] - uj[2]};
          const float idd = frcp((dist + c.eps) * (dist + c.eps));
          const float pi_ij = dot3(du, dr, DIM) * idd;  // dot(u_j - u_i, -r_ij) / (d + EPS)^2
          const float s = fdiv(m_j, rho_j) * pi_ij * c.delta_coef * fdiv(1.0f, o.rho);
#pragma unroll
          for (int k = 0; k < DIM; ++k) a.av[k] += s * (gw * (dr[k] * id));
        }
      }
      if (ex.av) {  // :404-428
        if (o.tag == SPHB200_TAG_FLUID && tag_j == SPHB200_TAG_FLUID) {
          const float rho_ab = (o.rho + rho_j) / 2.0f;
          const float du[3] = {o.u[0] - uj[0], o.u[1] - uj[1], o.u[2] - uj[2]};
          const float num = (m_j * c.av_coef) * dot3(du, dr, DIM);
          const float idd = frcp(rho_ab * (dist * dist + c.av_eps));
#pragma unroll
          for (int k = 0; k < DIM; ++k) a.av[k] += num * (gw * (dr[k] * id)) * idd;
        }
      }
      if (ex.heat) {  // :591-610
        const float4 hj = sq[ex.q_h * cap + j];
        const float eff = fdiv(o.kappa * hj.y, o.kappa + hj.y);
        float rk = 0.f;
#pragma unroll
        for (int k = 0; k < DIM; ++k) rk += dr[k] * (gw * (dr[k] * id));
        const float F = fdiv(rk, dist * dist + c.eps);
        a.dT += fdiv(4.0f * m_j * eff * (o.T - hj.x) * F, o.Cp * o.rho * rho_j);
      }
    }
  }
  __device__ static void finish(const Consts& c, const Frame& f, const Extra& ex, int p,
                                const Own& o, const Acc& a) {
    const float4 pt = f.pt[p];
    float r[3] = {pt.x, pt.y, pt.z}, g[3];
    g_ext_of<DIM>(c, f, p, r, g);
    float du[3], dv[3];
    // A_i term of the standard acceleration, factored out of the pair sum (see the header)
    float ai = 0.f;
    if (SOLVER == SPHB200_SOLVER_SPH && (COMPACT ? TVF : qv(ex) >= 0))
      ai = 0.5f * o.rho * dot3(o.dvu, a.tv, DIM);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float s = a.a[k];
      if (SOLVER == SPHB200_SOLVER_SPH) s = fmaf(ai, o.u[k], s);
      if (FEAT == FORCE_GENERIC && (ex.av || ex.delta)) s = s + a.av[k];  // :925-928 / :311
      du[k] = s + g[k];                                     // :936
      dv[k] = a.tv[k] * c.p_bg_tvf;                         // :205-211
    }
    float dTdt = a.dT;
    // case bc_fn, derivative part (u, v, p, T are patched by k_bc after the sweep)
    if (ex.bc_on && o.tag >= 0 && o.tag <= 3) {
      const uint32_t fl = c.bc[o.tag].flags;
      if (fl & SPHB200_BC_ZERO_DUDT) du[0] = du[1] = du[2] = 0.f;
      if (fl & SPHB200_BC_ZERO_DVDT) dv[0] = dv[1] = dv[2] = 0.f;
      if (fl & SPHB200_BC_ZERO_DTDT) dTdt = 0.f;
      if (o.tag == SPHB200_TAG_FLUID) {
        if (c.inflow_on && pt.x < c.inflow_x) dTdt = 0.f;
        if (c.outflow_on && pt.x > c.outflow_x) dTdt = 0.f;
      }
    }
    const float drhodt = f.du[p].w;
    f.du[p] = make_float4(du[0], du[1], DIM == 3 ? du[2] : 0.f, drhodt);
    f.dv[p] = make_float4(dv[0], dv[1], DIM == 3 ? dv[2] : 0.f, 0.f);
    if (FEAT == FORCE_GENERIC && ex.heat) reinterpret_cast<float*>(&f.st[p])[3] = dTdt;
  }
};

// bc_fn, value part: u, v, p, T overwritten per tag (after the force sweep has
// consumed the wall values the solver computed).
template <int DIM>
__global__ void __launch_bounds__(256) k_bc(int n, Slab sl, Consts c, Frame f) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (sl.dn ? sl.dn[DN_OWN] : n)) return;
  const int p = sl.base + t;
  const float4 pt = f.pt[p];
  const int tag = __float_as_int(pt.w);
  if (tag < 0 || tag > 3) return;
  const sphb200_bc_rule r = c.bc[tag];
  if (r.flags & SPHB200_BC_SET_U) {
    float4 q = f.um[p];
    f.um[p] = make_float4(r.u[0], r.u[1], DIM == 3 ? r.u[2] : 0.f, q.w);
  }
  if (r.flags & SPHB200_BC_SET_V) {
    float4 q = f.vv[p];
    f.vv[p] = make_float4(r.v[0], r.v[1], DIM == 3 ? r.v[2] : 0.f, q.w);
  }
  const bool inflow = c.inflow_on && tag == SPHB200_TAG_FLUID && pt.x < c.inflow_x;
  if ((r.flags & (SPHB200_BC_SET_P | SPHB200_BC_SET_T)) || inflow) {
    float4 q = f.st[p];
    if (inflow) q.z = c.inflow_T;
    if (r.flags & SPHB200_BC_SET_P) q.y = r.p;
    if (r.flags & SPHB200_BC_SET_T) q.z = r.T;
    f.st[p] = q;
  }
}

// ---------------------------------------------------------------------------
// Sparse neighbour list in original indices.  The thread's own particle is the
// SENDER (row 1); receivers (row 0) are written ascending.  Two launches: count,
// then (after an exclusive scan over senders in original order) fill.
template <int DIM>
struct PhysNeighbors {
  static constexpr int MINB = 2;
  static constexpr bool SENDER_VIEW = true;
  static constexpr bool SPARSE = false;
  // builder phase 2 takes two survivors per trip (sweep.cuh); heavy pair() bodies opt out
  static constexpr bool PAIR2 = true;
  struct Own {
    int id;
    long long off;
  };
  struct Acc {
    int n;
  };
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;  // arrays the duo sweeps stage by bulk copy (sweep2.cuh)
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  template <class F>
  __device__ static void each_acc(Acc&, F) {}
  __device__ static void stage(const Consts&, const Frame& f, const Extra&, int gp, float4* sq,
                               int cap, int d) {
    sq[d] = f.pt[gp];
    sq[cap + d] = make_float4(__int_as_float(f.id[gp]), 0.f, 0.f, 0.f);
  }
  __device__ static void load_own(const Consts&, const Frame& f, const Extra& ex, int p, float4,
                                  Own& o) {
    o.id = f.id[p];
    o.off = ex.nl_fill ? (long long)(unsigned)ex.nl_offsets[o.id] : 0;
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) { a.n = 0; }
  __device__ static void pair(const Consts&, const Extra& ex, const Own& o, Acc& a,
                              const float4* sq, int cap, int j, float4, const float (&)[3], float) {
    const int jid = __float_as_int(sq[cap + j].x);
    if (ex.nl_mask_self && jid == o.id) return;
    if (ex.nl_fill) {
      // insertion into the ascending run [off, off + n) of row 0
      long long pos = o.off + a.n;
      if (pos < ex.nl_capacity) {
        int* row0 = ex.nl_idx;
        long long k = pos;
        while (k > o.off && row0[k - 1] > jid) {
          row0[k] = row0[k - 1];
          --k;
        }
        row0[k] = jid;
        ex.nl_idx[ex.nl_capacity + pos] = o.id;
      }
    }
    ++a.n;
  }
  __device__ static void finish(const Consts&, const Frame&, const Extra& ex, int, const Own& o,
                                const Acc& a) {
    if (!ex.nl_fill) ex.nl_counts[o.id] = a.n;
  }
};

}  // namespace sphb200
