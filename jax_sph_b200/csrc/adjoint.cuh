// adjoint.cuh -- vector-Jacobian product of one advance() of the SPH solver (row f4 of
// the reference differentiates through `advance` with jax.grad,
// notebooks/iclr24_grads.ipynb cell 5, jax_sph/integrator.py:22-56).
//
// Scope: solver SPH, density by summation (solver.py:792-802), standard acceleration
// (solver.py:221-256) without transport velocity (tvf = 0, the notebook's setting: A = 0),
// no wall / bc sweep, constant external force.  Reverse mode by hand, as two gather sweeps over
// the same cell-sorted particles (k_sweep<..., LIST_NONE>: search + exact membership + pair):
//
//   a_i   = sum_j S_ij / m_i G(d) s_ij + g,     s_ij = -p_ij r_ij + eta_ij (u_i - u_j),
//           S_ij = (m_i/rho_i)^2 + (m_j/rho_j)^2,  G(d) = w'(d) / (d + EPS),
//           p_ij = (rho_j p_i + rho_i p_j) / (rho_i + rho_j)
//   rho_i = m_i sum_j w(d_ij),   p_i = P(rho_i)
//
// With the cotangent abar_i of a_i and b_ij = abar_i / m_i - abar_j / m_j, the scalar
// L = sum_i abar_i . a_i is the sum over UNORDERED pairs of S G (s_ij . b_ij) (s_ji = -s_ij,
// b_ji = -b_ij), so every derivative is a gather over the neighbours of i:
//
//   rbar_i   = sum_j S [ G'(d) / d (s_ij . b_ij) r_ij - G p_ij b_ij ]            (PhysForceAdj)
//   ubar_i   = sum_j S G eta_ij b_ij
//   pbar_i   = sum_j S G (-(r_ij . b_ij)) rho_j / (rho_i + rho_j)
//   rhobar_i = sum_j [ -2 (m_i/rho_i)^2 / rho_i G (s_ij . b_ij)
//                      + S G (-(r_ij . b_ij)) (p_j - p_ij) / (rho_i + rho_j) ]
//   q_i      = m_i (rhobar_i + rhobar_ext_i + (pbar_i + pbar_ext_i) P'(rho_i))   (k_adj_eos, fluid)
//   rbar_i  += sum_j (q_i + q_j) w'(d) / d r_ij                                  (PhysDensAdj)
//
// and the integrator (u1 = u0 + dt a0, v1 = u1, r1 = wrap(r0 + dt v1)) backwards (k_adj_finish).
// Membership of the neighbour set is piecewise constant and every kernel reaches the cutoff with
// zero value and slope, so the product is the derivative the reference's autodiff gives.
#pragma once
#include "cells.cuh"
#include "common.cuh"

namespace sphb200 {

// d^2 w / d r^2 of the two compile-time kernels (kernel.py:51-103)
template <int KERN>
__device__ __forceinline__ float kernel_ggw(const Consts& c, float r) {
  const float q = r * c.ooh;
  if (KERN == SPHB200_KERNEL_QSK) {
    const float q1 = fmaxf(0.0f, 1.0f - q), q2 = fmaxf(0.0f, 2.0f - q), q3 = fmaxf(0.0f, 3.0f - q);
    return c.sigma_ooh * c.ooh * ((20.0f * (q3 * q3 * q3) - 120.0f * (q2 * q2 * q2)) + 300.0f * (q1 * q1 * q1));
  } else {
    const float q1 = fmaxf(0.0f, 1.0f - 0.5f * q);
    return c.sigma_ooh * c.ooh * (-5.0f * (q1 * q1 * q1) + 7.5f * q * (q1 * q1));
  }
}

// ---------------------------------------------------------------------------
template <int DIM, int KERN>
struct PhysForceAdj {
  static constexpr int MINB = 1;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;
  static constexpr bool PAIR2 = false;
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  struct Own {
    float u[3], am[3];
    float rho, p, eta, V2;
  };
  struct Acc {
    float rbar[3], ubar[3];
    float rhobar, pbar;
  };
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f(a.rbar[k]); f(a.ubar[k]);
    }
    f(a.rhobar); f(a.pbar);
  }
  // staged: (x, y, z, rho) (u, p) (abar / m, (m/rho)^2) (eta, -, -, -)
  __device__ static void stage(const Consts&, const Frame& f, const Extra& ex, int gp, float4* sq,
                               int cap, int d) {
    const float4 pt = f.pt[gp], um = f.um[gp], st = f.st[gp], am = ex.adj.am[gp];
    const float vol = um.w / st.x;
    sq[d] = make_float4(pt.x, pt.y, pt.z, st.x);
    sq[cap + d] = make_float4(um.x, um.y, um.z, st.y);
    sq[2 * cap + d] = make_float4(am.x, am.y, am.z, vol * vol);
    sq[3 * cap + d] = make_float4(f.vv[gp].w, 0.f, 0.f, 0.f);
  }
  __device__ static void load_own(const Consts&, const Frame& f, const Extra& ex, int p, float4,
                                  Own& o) {
    const float4 um = f.um[p], st = f.st[p], am = ex.adj.am[p];
    o.u[0] = um.x; o.u[1] = um.y; o.u[2] = um.z;
    o.am[0] = am.x; o.am[1] = am.y; o.am[2] = am.z;
    o.rho = st.x; o.p = st.y;
    o.eta = f.vv[p].w;
    const float vol = um.w / st.x;
    o.V2 = vol * vol;
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a.rbar[k] = a.ubar[k] = 0.f;
    a.rhobar = a.pbar = 0.f;
  }
  __device__ static void pair(const Consts& c, const Extra&, const Own& o, Acc& a, const float4* sq,
                              int cap, int j, float4 pj, const float (&dr)[3], float d2) {
    const float4 q1 = sq[cap + j], q2 = sq[2 * cap + j];
    const float eta_j = sq[3 * cap + j].x;
    const float rho_j = pj.w, p_j = q1.w, V2_j = q2.w;
    const float uj[3] = {q1.x, q1.y, q1.z}, amj[3] = {q2.x, q2.y, q2.z};
    const float dist = fsqrt(d2);
    const float gw = kernel_gw<KERN>(c, dist), ggw = kernel_ggw<KERN>(c, dist);
    const float idp = 1.0f / (dist + c.eps);
    const float G = gw * idp;
    const float Gp = ggw * idp - gw * idp * idp;
    const float inv_d = 1.0f / fmaxf(dist, 1e-30f);  // r / d; the self pair has r = 0
    const float S = o.V2 + V2_j;
    const float eta_ij = 2.0f * o.eta * eta_j / (o.eta + eta_j + c.eps);
    const float irr = 1.0f / (o.rho + rho_j);
    const float p_ij = (rho_j * o.p + o.rho * p_j) * irr;
    float s[3], b[3], sb = 0.f, rb = 0.f;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      s[k] = -p_ij * dr[k] + eta_ij * (o.u[k] - uj[k]);
      b[k] = o.am[k] - amj[k];
      sb += s[k] * b[k];
      rb += dr[k] * b[k];
    }
    const float SG = S * G;
    const float radial = S * Gp * sb * inv_d;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      a.rbar[k] += radial * dr[k] - SG * p_ij * b[k];
      a.ubar[k] += SG * eta_ij * b[k];
    }
    a.pbar += SG * (-rb) * rho_j * irr;
    a.rhobar += (-2.0f * o.V2 / o.rho) * G * sb + SG * (-rb) * (p_j - p_ij) * irr;
  }
  __device__ static void finish(const Consts&, const Frame&, const Extra& ex, int p, const Own&,
                                const Acc& a) {
    const float4 r0 = ex.adj.rbar[p];
    ex.adj.rbar[p] = make_float4(r0.x + a.rbar[0], r0.y + a.rbar[1], r0.z + a.rbar[2], 0.f);
    ex.adj.ubar[p] = make_float4(a.ubar[0], a.ubar[1], a.ubar[2], 0.f);
    ex.adj.rp[p] = make_float4(a.rhobar, a.pbar, 0.f, 0.f);
  }
};

// rbar_i += sum_j (q_i + q_j) w'(d) / d r_ij  (density by summation, solver.py:792-796)
template <int DIM, int KERN>
struct PhysDensAdj {
  static constexpr int MINB = 2;
  static constexpr bool SENDER_VIEW = false;
  static constexpr bool SPARSE = false;
  static constexpr bool PAIR2 = false;
  static constexpr bool HAS_PAIR2 = false;
  static constexpr bool HAS_DUO = false;
  static constexpr int DUO_MINB = 1;
  static constexpr int DUO_COPIES = 0;
  static constexpr bool DUO_REST = false;
  __device__ static void duo_sources(const Frame&, const Extra&, const float4* (&)[1]) {}
  __device__ static void stage_rest(const Consts&, const Frame&, const Extra&, int, float4*, int, int) {}
  struct Own {
    float q;
  };
  struct Acc {
    float rbar[3];
  };
  template <class F>
  __device__ static void each_acc(Acc& a, F f) {
    f(a.rbar[0]); f(a.rbar[1]); f(a.rbar[2]);
  }
  __device__ static void stage(const Consts&, const Frame& f, const Extra& ex, int gp, float4* sq,
                               int, int d) {
    const float4 pt = f.pt[gp];
    sq[d] = make_float4(pt.x, pt.y, pt.z, ex.adj.rp[gp].z);
  }
  __device__ static void load_own(const Consts&, const Frame&, const Extra& ex, int p, float4, Own& o) {
    o.q = ex.adj.rp[p].z;
  }
  __device__ static bool active(const Consts&, const Own&) { return true; }
  __device__ static void init(Acc& a) { a.rbar[0] = a.rbar[1] = a.rbar[2] = 0.f; }
  __device__ static void pair(const Consts& c, const Extra&, const Own& o, Acc& a, const float4*, int,
                              int, float4 pj, const float (&dr)[3], float d2) {
    const float dist = fsqrt(d2);
    const float f = (o.q + pj.w) * kernel_gw<KERN>(c, dist) / fmaxf(dist, 1e-30f);
#pragma unroll
    for (int k = 0; k < DIM; ++k) a.rbar[k] += f * dr[k];
  }
  __device__ static void finish(const Consts&, const Frame&, const Extra& ex, int p, const Own&,
                                const Acc& a) {
    const float4 r0 = ex.adj.rbar[p];
    ex.adj.rbar[p] = make_float4(r0.x + a.rbar[0], r0.y + a.rbar[1], r0.z + a.rbar[2], 0.f);
  }
};

// ---------------------------------------------------------------------------
// Cotangents in the caller's particle order <-> slot order (through the slot's particle id).
struct CotPtrs {
  const float *r, *u, *v, *dudt, *rho, *p;
};
struct CotOut {
  float *r, *u, *v, *dudt, *dvdt, *rho, *p;
};

template <int DIM>
__device__ __forceinline__ float4 cot_vec(const float* a, int i) {
  if (a == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
  return make_float4(a[DIM * i], a[DIM * i + 1], DIM == 3 ? a[DIM * i + 2] : 0.f, 0.f);
}

// am = dudt cotangent / m, rbar = r cotangent (the sweeps add to it)
template <int DIM>
__global__ void __launch_bounds__(256) k_adj_begin(int n, Frame f, AdjBufs ab, CotPtrs ct) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = f.id[p];
  const float im = 1.0f / f.um[p].w;
  const float4 a = cot_vec<DIM>(ct.dudt, i);
  ab.am[p] = make_float4(a.x * im, a.y * im, a.z * im, 0.f);
  ab.rbar[p] = cot_vec<DIM>(ct.r, i);
}

// q = m (rhobar + pbar P'(rho)) for the particles whose density the sweep computes (fluid)
__global__ void __launch_bounds__(256) k_adj_eos(int n, Consts c, Frame f, AdjBufs ab, CotPtrs ct) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = f.id[p];
  float4 rp = ab.rp[p];
  const float rho = f.st[p].x;
  float dp;  // d p / d rho (eos.py:33-38 / :53-57)
  if (c.eos == SPHB200_EOS_TAIT) {
    dp = c.p_ref / c.rho_ref;
    if (c.gamma != 1.0f) dp *= c.gamma * powf(rho / c.rho_ref, c.gamma - 1.0f);
  } else {
    dp = c.c100;
  }
  const float rhobar = rp.x + (ct.rho ? ct.rho[i] : 0.f) + (rp.y + (ct.p ? ct.p[i] : 0.f)) * dp;
  const int tag = __float_as_int(f.pt[p].w);
  rp.z = tag == SPHB200_TAG_FLUID ? f.um[p].w * rhobar : 0.f;
  ab.rp[p] = rp;
}

// the integrator backwards (integrator.py:26-30 with tvf = 0), scattered to the caller's order
template <int DIM>
__global__ void __launch_bounds__(256) k_adj_finish(int n, Frame f, AdjBufs ab, CotPtrs ct,
                                                    CotOut out, float dt, int integrate) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int i = f.id[p];
  const float4 rbar = ab.rbar[p], uf = ab.ubar[p];
  const float4 ue = cot_vec<DIM>(ct.u, i), ve = cot_vec<DIM>(ct.v, i);
  float rb[3] = {rbar.x, rbar.y, rbar.z};
  float vb[3] = {ve.x, ve.y, ve.z}, ub[3] = {ue.x + uf.x, ue.y + uf.y, ue.z + uf.z};
  if (integrate) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vb[k] += dt * rb[k];  // r1 = r0 + dt v1
      ub[k] += vb[k];       // v1 = u1
    }
  }
#pragma unroll
  for (int k = 0; k < DIM; ++k) {
    if (out.r) out.r[DIM * i + k] = rb[k];
    if (out.u) out.u[DIM * i + k] = ub[k];
    if (out.v) out.v[DIM * i + k] = integrate ? 0.f : vb[k];  // advance() overwrites v
    if (out.dudt) out.dudt[DIM * i + k] = integrate ? dt * ub[k] : 0.f;  // u1 = u0 + dt dudt0
    if (out.dvdt) out.dvdt[DIM * i + k] = 0.f;
  }
  if (out.rho) out.rho[i] = 0.f;  // the summation overwrites rho and p
  if (out.p) out.p[i] = 0.f;
}

}  // namespace sphb200
