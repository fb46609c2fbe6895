// common.cuh -- shared device structs and bit-faithful arithmetic helpers.
//
// sm_100a only.  Everything here is new code; comments cite the reference
// expressions whose *values* the helpers reproduce (paths relative to the
// reference tree).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sphb200.h"

namespace sphb200 {

constexpr unsigned FULL_MASK = 0xffffffffu;

// ---------------------------------------------------------------------------
// Resident particle frame: cell-sorted, float4-packed so that every pass moves
// 16-byte vectors (coalesced LDG.128 / STG.128 and LDS.128 when staged).
//   pt = (x, y, z, tag-bits)        um = (ux, uy, uz, mass)
//   vv = (vx, vy, vz, eta)          st = (rho, p, T, dTdt)
//   du = (dudt xyz, drhodt)         dv = (dvdt xyz, unused)
//   kc = (kappa, Cp)                nw = (nw xyz, 0)        ut = (u_tilde xyz, 0)  [RIE walls]
//   ge = (g_ext xyz, 0)  [only g_mode == ARRAY]             id = original particle index
struct Frame {
  float4 *pt, *um, *vv, *st, *du, *dv, *nw, *ut, *ge;
  // Delta-SPH density diffusion: rows of the renormalisation matrix L, gradient terms G (own
  // use) and H (read from neighbours), one plane of n quads each
  float4 *dl0, *dl1, *dl2, *dg0, *dg1;
  float2 *kc;
  int *id;
};

// Cell grid + stencil/tiling plan (host-computed, passed by value).
struct Grid {
  int n[3];      // cells per axis (n[2] = 1 in 2D)
  int S[3];      // stencil half width in cells
  int W[3];      // per-particle window length  = min(2S+1, n)
  int T[3];      // tile size in cells
  int nt[3];     // tiles per axis
  int ncells;
  int exact_all; // small-box mode: no fast reject, window covers the whole axis
  float inv_cell[3];
  float box[3], half[3];
  float c2;      // f32(cutoff)^2 rounded to f32  (jax_md/partition.py:820-822)
  float c2_hi;   // threshold of the search: (cutoff + skin)^2 plus the rounding of its cheap test
  float c2_lo;   // below this the pair is inside for every rounding
  float c2_fb;   // cheap-reject threshold of the fall-back search inside list consumers: c2_hi,
                 // or +inf when the cell table may be frozen (a particle that drifted across the
                 // periodic seam since the last sort is not where the image shift expects it)
  // Slab decomposition (one rank's view; single GPU: goff = 0, ng = n, own = [0, n)).
  // n[] are the LOCAL cells this rank indexes (own layers + S halo layers each side along the
  // slab axis), positions stay global, so the periodic image of a staged cell is decided by
  // its GLOBAL index goff + local.
  int goff[3];   // global cell index of local cell 0 (negative for the rank at the box bottom)
  int ng[3];     // global cells per axis
  int own_lo[3], own_hi[3];  // local cell range this rank owns (sweeps tile exactly this range)
  int block0;    // first tile of this launch (a sweep may be launched over a sub-range of tiles)
  int ntl;       // tiles of this launch: [block0, block0 + ntl)
  int imargin;   // 1: a tile is "interior" only if its stencil keeps S more cells away from the
                 // periodic seam (engines whose re-sort criterion is the RELATIVE drift of
                 // neighbours, cells.cuh k_drift_box: particles may be far from where they were sorted)
};

// Derived float32 constants, rounded where the reference rounds them.
struct Consts {
  int dim, solver, kernel, eos;
  uint32_t flags;
  float eps;        // finfo(float32).eps, solver.py:21
  float ooh;        // f32(1/h)               kernel.py:55
  float sigma;      // f32(sigma)             kernel.py:56-62 / :80-86
  float sigma_ooh;  // f32(sigma) * f32(1/h)  (jax.grad of w)
  float gwk[3];     // QSK: sigma_ooh * (-5, 30, -75), the coefficients of d w / d r (kernel.py:51-77)
  float dt_s;       // f32(WCSPH.dt)
  float p_ref, rho_ref, p_bg, gamma, inv_gamma, c100;
  float p_bg_tvf;   // eos.p_fn(0), solver.py:802
  float c_ref, eta_lim;
  int use_lim;
  float av_coef;    // f32(alpha * h_ab * c_ab), solver.py:413-415
  float av_eps;     // f32(0.01 * h_ab^2)
  float delta_coef; // DELTA: f32(alpha * support * c_ref * rho_ref), solver.py:303-308
  float delta_rho;  // DELTA: f32(c_ref * delta * support), solver.py:100
  int g_mode, g_axis;
  float g[3], g_lo, g_hi;
  sphb200_bc_rule bc[4];
  int inflow_on, outflow_on;
  float inflow_x, inflow_T, outflow_x;
  int any_walls;    // 0 when no particle carries a wall tag (host-known hint)
};

// Work arrays of the vector-Jacobian product (adjoint.cuh), one entry per slot.
struct AdjBufs {
  float4* am;    // abar / m (xyz): the cotangent of dudt over the particle's mass
  float4* rbar;  // position cotangent, accumulated by the two adjoint sweeps
  float4* ubar;  // velocity cotangent from the force sweep
  float4* rp;    // (rhobar_force, pbar_force, q = m (rhobar + pbar dp/drho), -)
};

// Per-launch switches of the sweep policies (phys.cuh).
struct Extra {
  float4* st_out;  // destination of the (rho, p, T, dTdt) quad (ping-pong, see engine.cu)
  int nq;          // quads staged per particle
  int sb;          // bytes staged per particle when the policy stages a compact record, else 0
  int q_v, q_h, q_nw, q_ut;  // optional staged quads (force sweep), -1 = absent
  int utilde;      // RIE & bc_trick & !free_slip: write u_tilde
  int wallT;       // RIE & bc_trick & heat: Shepard wall temperature
  int finalT;      // this sweep integrates T (no wall sweep follows)
  int heat, av, bc_on, free_slip, bc_trick;
  int delta;       // DELTA solver: velocity diffusion term of acceleration_delta_fn
  // compact force records of every slot (duo force sweep: staged by bulk copies, sweep2.cuh)
  float4 *rec0, *rec1, *rec2;
  float* rec_e;
  const float4* eta_ref;  // SPHB200_HINT_UNIFORM_ETA: the (v, eta) quad of the first own slot
  unsigned* err_word;     // device error word (a broken hint)
  AdjBufs adj;     // adjoint sweeps
  // neighbour-list materialiser
  int* nl_counts;
  const int* nl_offsets;
  int* nl_idx;
  long long nl_capacity;
  int nl_mask_self, nl_fill, nl_n;
};

// ---------------------------------------------------------------------------
// jnp.mod(t, side) for float32 (fmod + sign fix-up toward the divisor), the
// arithmetic of space.py:181 and :209.  fmodf is exact, so the only rounding is
// the fix-up add, exactly as in XLA / NumPy.
__device__ __forceinline__ float mod_side(float t, float side) {
  float m;
  if (t >= 0.0f && t < side) {
    m = t;
  } else if (t >= side && t < 2.0f * side) {
    m = __fsub_rn(t, side);  // exact (Sterbenz)
  } else if (t < 0.0f && t > -side) {
    m = __fadd_rn(t, side);  // fmod(t, side) == t, then + side
  } else {
    m = fmodf(t, side);
    if (m != 0.0f && (m < 0.0f)) m = __fadd_rn(m, side);
  }
  return m;
}

// space.py:170-181  periodic_displacement(side, a - b)
// Both positions lie in [0, side] (k_hash checks and flags SPHB200_ERR_OUTSIDE_BOX
// otherwise), so t = d + half is in (-side/2, 3 side/2) and jnp.mod reduces to two
// selects: t - side is exact (Sterbenz), t + side rounds exactly as the fix-up add.
__device__ __forceinline__ float disp1(float a, float b, float half, float side) {
  const float d = __fsub_rn(a, b);
  const float t = __fadd_rn(d, half);
  float m = t;
  if (t >= side) m = __fsub_rn(t, side);
  if (t < 0.0f) m = __fadd_rn(t, side);
  return __fsub_rn(m, half);
}

// Same value when the minimum image is known to be the direct difference
// (|d| < half, no periodic wrap between the two particles): 0 <= t < side.
__device__ __forceinline__ float disp1_nowrap(float a, float b, float half) {
  return __fsub_rn(__fadd_rn(__fsub_rn(a, b), half), half);
}

// space.py:184-192 sum of squares, left to right, no FMA contraction.
template <int DIM>
__device__ __forceinline__ float sumsq(const float (&d)[3]) {
  float acc = __fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1]));
  if (DIM == 3) acc = __fadd_rn(acc, __fmul_rn(d[2], d[2]));
  return acc;
}

// integrator.py:26-30 + space.py:207-209 for one particle; used by the hashing
// pass and again by the reorder pass, so both see bit-identical positions.
struct Kick {
  float dt;  // f32(dt)
  float c2;  // f32(f32(tvf*0.5) * f32(dt))
  int on;
};

template <int DIM>
__device__ __forceinline__ void integrate_one(const Kick k, const Grid& g, float (&r)[3],
                                              float (&u)[3], float (&v)[3], const float4 du,
                                              const float4 dv) {
  if (!k.on) return;
  const float a[3] = {du.x, du.y, du.z};
  const float b[3] = {dv.x, dv.y, dv.z};
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    u[d] = __fadd_rn(u[d], __fmul_rn(k.dt, a[d]));
    v[d] = __fadd_rn(u[d], __fmul_rn(k.c2, b[d]));
    float t = __fadd_rn(r[d], __fmul_rn(k.dt, v[d]));
    r[d] = mod_side(t, g.box[d]);
  }
}

// Cell coordinates of a position (clamped; the reference truncates
// position / cell_size, jax_md/partition.py:367).
// Returns the LOCAL cell index, or -1 when the position lies outside this rank's local
// cells (slab mode only; c[] then still holds the wrapped offset from local cell 0).
template <int DIM>
__device__ __forceinline__ int cell_of(const Grid& g, const float (&r)[3], int (&c)[3]) {
  bool local = true;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d < DIM) {
      int ci = (int)(r[d] * g.inv_cell[d]);
      ci = ci < 0 ? 0 : (ci >= g.ng[d] ? g.ng[d] - 1 : ci);
      ci -= g.goff[d];
      ci = ci < 0 ? ci + g.ng[d] : (ci >= g.ng[d] ? ci - g.ng[d] : ci);
      local = local && ci < g.n[d];
      c[d] = ci;
    } else {
      c[d] = 0;
    }
  }
  return local ? (c[2] * g.n[1] + c[1]) * g.n[0] + c[0] : -1;
}

// ---------------------------------------------------------------------------
// Smoothing kernels, evaluated inline (kernel.py:51-103).  x^5 = ((x^2)^2)*x as
// lax.integer_pow lowers it.
// Template slot of the run-time-switched kernels (every SPHB200_KERNEL_* other than QSK / WC2K)
constexpr int KERN_ANY = 7;

// The five kernels outside the headline pair, switched on c.kernel at run time.  which = 0: w,
// which = 1: d w / d r as jax.grad differentiates the reference expressions (the jnp.where
// masks of the Cubic / Gaussian kernels carry no gradient; d max(0, x) / dx = [x > 0]).
__device__ __forceinline__ float kernel_any(const Consts& c, float q, int which) {
  switch (c.kernel) {
    case SPHB200_KERNEL_CSK: {  // kernel.py:40-48
      const float c1 = (1.0f - q >= 0.0f) ? 1.0f : 0.0f;
      const float c2 = (2.0f - q < 1.0f && 2.0f - q >= 0.0f) ? 1.0f : 0.0f;
      const float t = 2.0f - q;
      if (which == 0)
        return (1.0f - 1.5f * (q * q) * (1.0f - q / 2.0f)) * c1 + (0.25f * ((t * t) * t)) * c2;
      return (-3.0f * q + 2.25f * (q * q)) * c1 + (-0.75f * (t * t)) * c2;
    }
    case SPHB200_KERNEL_WC4K: {  // kernel.py:130-134: q1^6 (35/12 q^2 + 3 q + 1)
      const float q1 = fmaxf(0.0f, 1.0f - 0.5f * q);
      const float a2 = q1 * q1, a4 = a2 * a2;
      const float poly = (35.0f / 12.0f) * (q * q) + 3.0f * q + 1.0f;
      if (which == 0) return (a2 * a4) * poly;
      return -3.0f * (q1 * a4) * poly + (a2 * a4) * ((35.0f / 6.0f) * q + 3.0f);
    }
    case SPHB200_KERNEL_WC6K: {  // kernel.py:161-165: q1^8 (4 q^3 + 6.25 q^2 + 4 q + 1)
      const float q1 = fmaxf(0.0f, 1.0f - 0.5f * q);
      const float a2 = q1 * q1, a4 = a2 * a2, a8 = a4 * a4;
      const float poly = 4.0f * ((q * q) * q) + 6.25f * (q * q) + 4.0f * q + 1.0f;
      if (which == 0) return a8 * poly;
      return -4.0f * ((q1 * a2) * a4) * poly + a8 * (12.0f * (q * q) + 12.5f * q + 4.0f);
    }
    case SPHB200_KERNEL_GK: {  // kernel.py:178-182
      const float m = (3.0f - q >= 0.0f) ? 1.0f : 0.0f;
      const float e = expf(-(q * q));
      return which == 0 ? m * e : m * ((-2.0f * q) * e);
    }
    default: {  // SPHB200_KERNEL_SGK, kernel.py:197-201
      const float m = (3.0f - q >= 0.0f) ? 1.0f : 0.0f;
      const float e = expf(-(q * q));
      const float a = 0.5f * (float)c.dim + 1.0f;
      return which == 0 ? m * e * (a - q * q) : m * ((-2.0f * q) * e) * (a + 1.0f - q * q);
    }
  }
}

template <int KERN>
__device__ __forceinline__ float kernel_w(const Consts& c, float r) {
  float q = r * c.ooh;
  if (KERN == SPHB200_KERNEL_QSK) {
    float q1 = fmaxf(0.0f, 1.0f - q), q2 = fmaxf(0.0f, 2.0f - q), q3 = fmaxf(0.0f, 3.0f - q);
    float a1 = q1 * q1, a2 = q2 * q2, a3 = q3 * q3;
    float p1 = (a1 * a1) * q1, p2 = (a2 * a2) * q2, p3 = (a3 * a3) * q3;
    return c.sigma * ((p3 - 6.0f * p2) + 15.0f * p1);
  } else if (KERN == SPHB200_KERNEL_WC2K) {
    float q1 = fmaxf(0.0f, 1.0f - 0.5f * q);
    float a = q1 * q1;
    return c.sigma * ((a * a) * (2.0f * q + 1.0f));
  } else {
    return c.sigma * kernel_any(c, q, 0);
  }
}

template <int KERN>
__device__ __forceinline__ float kernel_gw(const Consts& c, float r) {
  float q = r * c.ooh;
  if (KERN == SPHB200_KERNEL_QSK) {
    float q1 = fmaxf(0.0f, 1.0f - q), q2 = fmaxf(0.0f, 2.0f - q), q3 = fmaxf(0.0f, 3.0f - q);
    float a1 = q1 * q1, a2 = q2 * q2, a3 = q3 * q3;
    float poly = (-5.0f * (a3 * a3) + 30.0f * (a2 * a2)) - 75.0f * (a1 * a1);
    return c.sigma_ooh * poly;
  } else if (KERN == SPHB200_KERNEL_WC2K) {
    float q1 = fmaxf(0.0f, 1.0f - 0.5f * q);
    return c.sigma_ooh * ((-5.0f * q) * ((q1 * q1) * q1));
  } else {
    return c.sigma_ooh * kernel_any(c, q, 1);
  }
}

// Pair-level division: one MUFU.RCP (rcp.approx.ftz, <= 1 ulp) and a multiply, well inside
// the float32 summation-order noise of the sweeps; -DSPHB200_PRECISE
// restores IEEE division and square root.
//
// The distance: rsqrt.approx seeds one Newton step whose residual d2 - s*s is formed exactly
// by an FMA, so the result is the correctly rounded square root except for rare half-ulp
// ties -- as good as the IEEE routine, branch-free and half its instructions.  (A bare
// sqrt.approx, 2 ulp, is NOT enough: near the edge of the kernel support w ~ (3 - q)^5
// amplifies an error of q by 5 / (3 - q), which the Shepard wall averages of
// solver.py:505-526 expose for wall particles whose only fluid neighbours sit at the
// cutoff.)  d2 == 0 (the self pair) is lifted to 1e-30: the kernels see q = 0 either way.
#ifdef SPHB200_PRECISE
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float frcp(float a) { return 1.0f / a; }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
#else
__device__ __forceinline__ float frcp(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float fdiv(float a, float b) { return a * frcp(b); }
__device__ __forceinline__ float fsqrt(float a) {
  a = fmaxf(a, 1e-30f);
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  const float s = a * r;
  const float e = __fmaf_rn(-s, s, a);
  return __fmaf_rn(e, 0.5f * r, s);
}
#endif

// ---------------------------------------------------------------------------
// Packed float32 pairs: sm_100 executes add / mul / fma on two float32 lanes of a 64-bit register
// pair in one instruction (PTX add.rn.f32x2 ..., SASS FADD2 / FMUL2 / FFMA2), each lane rounded
// exactly like the scalar operation.  The list consumers keep two pairs in flight per thread
// (sweep.cuh) and evaluate them in the two halves: half the issue slots of the FP32 work.
// Packing and unpacking are register-pair moves (free).  ptxas contracts mul.f32x2 + add.f32x2
// into FFMA2 even with .rn: where the reference's value needs the unfused product (sumsq2) the
// adds are scalar __fadd_rn.
struct F2 {
  unsigned long long v;
};
__device__ __forceinline__ F2 f2(float a, float b) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ F2 f2(float a) { return f2(a, a); }
__device__ __forceinline__ float lo(F2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)y;
  return x;
}
__device__ __forceinline__ float hi(F2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  (void)x;
  return y;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b) {
  F2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// acc += a * b / acc += a with the accumulator tied to the output register pair (a loop-carried
// accumulator written by a fresh asm output costs two moves per trip otherwise)
__device__ __forceinline__ void fma2_into(F2& acc, F2 a, F2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v));
}
__device__ __forceinline__ void add2_into(F2& acc, F2 a) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.v) : "l"(a.v));
}
__device__ __forceinline__ F2 neg2(F2 a) { return f2(-lo(a), -hi(a)); }
__device__ __forceinline__ F2 max2(F2 a, float b) { return f2(fmaxf(lo(a), b), fmaxf(hi(a), b)); }
__device__ __forceinline__ F2 sel2(bool v0, bool v1, F2 a) { return f2(v0 ? lo(a) : 0.0f, v1 ? hi(a) : 0.0f); }

// disp1_nowrap of two neighbours at once (same value per half as the scalar form)
__device__ __forceinline__ F2 disp2_nowrap(float a, F2 b, float half) {
  return sub2(add2(sub2(f2(a), b), f2(half)), f2(half));
}
// ... and of two own particles against one neighbour coordinate
__device__ __forceinline__ F2 disp2_nowrap2(F2 a, float b, float half) {
  return sub2(add2(sub2(a, f2(b)), f2(half)), f2(half));
}
// sumsq of two displacements: products packed, the two adds scalar and unfused (space.py:184-192)
template <int DIM>
__device__ __forceinline__ F2 sumsq2(const F2 (&d)[3]) {
  const F2 x = mul2(d[0], d[0]), y = mul2(d[1], d[1]);
  float s0 = __fadd_rn(lo(x), lo(y)), s1 = __fadd_rn(hi(x), hi(y));
  if (DIM == 3) {
    const F2 z = mul2(d[2], d[2]);
    s0 = __fadd_rn(s0, lo(z));
    s1 = __fadd_rn(s1, hi(z));
  }
  return f2(s0, s1);
}

#ifdef SPHB200_PRECISE
__device__ __forceinline__ F2 frcp2(F2 a) { return f2(1.0f / lo(a), 1.0f / hi(a)); }
__device__ __forceinline__ F2 fsqrt2(F2 a) { return f2(__fsqrt_rn(lo(a)), __fsqrt_rn(hi(a))); }
#else
__device__ __forceinline__ F2 frcp2(F2 a) { return f2(frcp(lo(a)), frcp(hi(a))); }
// fsqrt of both halves: the rsqrt seeds are scalar MUFU ops, the Newton step is packed
__device__ __forceinline__ F2 fsqrt2(F2 a) {
  a = max2(a, 1e-30f);
  float r0, r1;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(lo(a)));
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(hi(a)));
  const F2 r = f2(r0, r1);
  const F2 s = mul2(a, r);
  // s + (a - s s) r / 2 as s - (s s - a)(r / 2): the residual is one FFMA2 (ptxas folds the
  // subtraction into the product), no separate negation
  const F2 t = sub2(mul2(s, s), a);
  return fma2(t, mul2(r, f2(-0.5f)), s);
}
#endif

// kernel_w / kernel_gw of two distances (same arithmetic per half as the scalar forms)
template <int KERN>
__device__ __forceinline__ F2 kernel_w2(const Consts& c, F2 r) {
  if (KERN == SPHB200_KERNEL_QSK) {
    // k - r / h with one rounding, as the compiler contracts the scalar form
    const F2 no = f2(-c.ooh);
    const F2 q1 = max2(fma2(r, no, f2(1.0f)), 0.0f), q2 = max2(fma2(r, no, f2(2.0f)), 0.0f),
             q3 = max2(fma2(r, no, f2(3.0f)), 0.0f);
    const F2 a1 = mul2(q1, q1), a2 = mul2(q2, q2), a3 = mul2(q3, q3);
    const F2 p1 = mul2(mul2(a1, a1), q1), p2 = mul2(mul2(a2, a2), q2), p3 = mul2(mul2(a3, a3), q3);
    return mul2(f2(c.sigma), fma2(f2(15.0f), p1, fma2(f2(-6.0f), p2, p3)));
  } else if (KERN == SPHB200_KERNEL_WC2K) {
    const F2 q = mul2(r, f2(c.ooh));
    const F2 q1 = max2(fma2(q, f2(-0.5f), f2(1.0f)), 0.0f);
    const F2 a = mul2(q1, q1);
    return mul2(f2(c.sigma), mul2(mul2(a, a), fma2(q, f2(2.0f), f2(1.0f))));
  } else {
    return f2(kernel_w<KERN>(c, lo(r)), kernel_w<KERN>(c, hi(r)));
  }
}

template <int KERN>
__device__ __forceinline__ F2 kernel_gw2(const Consts& c, F2 r) {
  if (KERN == SPHB200_KERNEL_QSK) {
    // k - r / h with one rounding, as the compiler contracts the scalar form
    const F2 no = f2(-c.ooh);
    const F2 q1 = max2(fma2(r, no, f2(1.0f)), 0.0f), q2 = max2(fma2(r, no, f2(2.0f)), 0.0f),
             q3 = max2(fma2(r, no, f2(3.0f)), 0.0f);
    const F2 a1 = mul2(q1, q1), a2 = mul2(q2, q2), a3 = mul2(q3, q3);
    const F2 poly = fma2(f2(-75.0f), mul2(a1, a1), fma2(f2(30.0f), mul2(a2, a2), mul2(f2(-5.0f), mul2(a3, a3))));
    return mul2(f2(c.sigma_ooh), poly);
  } else if (KERN == SPHB200_KERNEL_WC2K) {
    const F2 q = mul2(r, f2(c.ooh));
    const F2 q1 = max2(fma2(q, f2(-0.5f), f2(1.0f)), 0.0f);
    return mul2(f2(c.sigma_ooh), mul2(mul2(q, f2(-5.0f)), mul2(mul2(q1, q1), q1)));
  } else {
    return f2(kernel_gw<KERN>(c, lo(r)), kernel_gw<KERN>(c, hi(r)));
  }
}

// ---------------------------------------------------------------------------
// Bulk asynchronous copies global -> shared (the 1-D form of TMA: cp.async.bulk, SASS UBLKCP) that
// signal an mbarrier in shared memory with the bytes they delivered.  Addresses and sizes are
// multiples of 16 bytes.
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this thread's earlier generic-proxy accesses of shared memory (and, through a preceding
// barrier, the block's) before its later asynchronous copies into the same memory
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one arrival + `bytes` more expected from the copies of the current phase
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// asynchronous 4-byte copy global -> shared (LDGSTS) and the wait for all of this thread's
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// kernel_gw2 with sigma / h folded into the coefficients of the quintic spline (one multiply
// fewer per evaluation; the other kernels as they are)
template <int KERN>
__device__ __forceinline__ F2 kernel_gw2_folded(const Consts& c, F2 r) {
  if (KERN == SPHB200_KERNEL_QSK) {
    const F2 no = f2(-c.ooh);
    const F2 q1 = max2(fma2(r, no, f2(1.0f)), 0.0f), q2 = max2(fma2(r, no, f2(2.0f)), 0.0f),
             q3 = max2(fma2(r, no, f2(3.0f)), 0.0f);
    const F2 a1 = mul2(q1, q1), a2 = mul2(q2, q2), a3 = mul2(q3, q3);
    return fma2(f2(c.gwk[2]), mul2(a1, a1), fma2(f2(c.gwk[1]), mul2(a2, a2), mul2(f2(c.gwk[0]), mul2(a3, a3))));
  } else {
    return kernel_gw2<KERN>(c, r);
  }
}

// eos.py:33-38 / :53-57
__device__ __forceinline__ float eos_p(const Consts& c, float rho) {
  if (c.eos == SPHB200_EOS_TAIT) {
    float x = rho / c.rho_ref;
    if (c.gamma != 1.0f) x = powf(x, c.gamma);
    return c.p_ref * (x - 1.0f) + c.p_bg;
  }
  return c.c100 * (rho - c.rho_ref) + c.p_bg;
}

__device__ __forceinline__ float eos_rho(const Consts& c, float p) {
  if (c.eos == SPHB200_EOS_TAIT) {
    float x = ((p + c.p_ref) - c.p_bg) / c.p_ref;
    if (c.gamma != 1.0f) x = powf(x, c.inv_gamma);
    return c.rho_ref * x;
  }
  return (p - c.p_bg) / c.c100 + c.rho_ref;
}

// g_ext_fn(r) in table form.
template <int DIM>
__device__ __forceinline__ void g_ext_of(const Consts& c, const Frame& f, int p,
                                         const float (&r)[3], float (&g)[3]) {
  g[0] = g[1] = g[2] = 0.0f;
  if (c.g_mode == SPHB200_G_CONST) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) g[d] = c.g[d];
  } else if (c.g_mode == SPHB200_G_BAND) {
    float x = r[c.g_axis];
    bool in = (x < c.g_hi) && (x > c.g_lo);
#pragma unroll
    for (int d = 0; d < DIM; ++d) g[d] = in ? c.g[d] : 0.0f;
  } else if (c.g_mode == SPHB200_G_ARRAY) {
    float4 q = f.ge[p];
    g[0] = q.x; g[1] = q.y; g[2] = q.z;
  }
}

__device__ __forceinline__ bool is_wall_tag(int tag) { return tag >= 1 && tag <= 3; }

}  // namespace sphb200
