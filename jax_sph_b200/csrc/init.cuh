// init.cuh -- on-device case initialisation.
//
// Replaces, for the regular-lattice cases, the host-side generators of the reference:
//   pos_init_cartesian_2d / _3d      jax_sph/utils.py:35-54   (meshgrid "xy" + ravel + (i + 0.5) dx)
//   TGV._init_velocity2D / 3D        cases/tgv.py:37-51
//   uniform fields and tags          jax_sph/case_setup.py:152-181, cases/ht.py:90-97
// One write-only pass: every output array is written once, coalesced (a warp's rows are
// adjacent in every array), nothing is read -- bound by HBM write bandwidth.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sphb200.h"
#include "common.cuh"

namespace sphb200 {

// ---- velocity field of a case at given positions (cases/tgv.py:37-51) ----------------------
template <int DIM>
__device__ __forceinline__ void case_velocity(int kind, const float (&x)[3], float (&u)[3]) {
  u[0] = u[1] = u[2] = 0.f;
  if (kind == SPHB200_VEL_TGV2D) {
    float sx, cx, sy, cy;
    sincosf(6.2831853071795864769f * x[0], &sx, &cx);
    sincosf(6.2831853071795864769f * x[1], &sy, &cy);
    u[0] = -1.0f * cx * sy;
    u[1] = sx * cy;
  } else if (kind == SPHB200_VEL_TGV3D) {
    float sx, cx, sy, cy;
    sincosf(x[0], &sx, &cx);
    sincosf(x[1], &sy, &cy);
    const float cz = cosf(x[2]);
    u[0] = sx * cy * cz;
    u[1] = -cx * sy * cz;
  }
}

struct LatticeArgs {
  sphb200_lattice l;
  sphb200_state out;
  int32_t* ids;
  long long rows;
};

template <int DIM>
__global__ void __launch_bounds__(256) k_init_lattice(const LatticeArgs a) {
  const sphb200_lattice& l = a.l;
  const int n0 = l.n[0];
  const int nk = l.k_hi - l.k_lo;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < a.rows; row += stride) {
    int ic[3];
    long long id;
    if (DIM == 3) {
      const long long col = row / nk;  // iy * n0 + ix
      ic[2] = l.k_lo + (int)(row - col * nk);
      ic[1] = (int)(col / n0);
      ic[0] = (int)(col - (long long)ic[1] * n0);
      id = col * l.n[2] + ic[2];
    } else {
      ic[1] = l.k_lo + (int)(row / n0);
      ic[0] = (int)(row - (long long)(ic[1] - l.k_lo) * n0);
      ic[2] = 0;
      id = (long long)ic[1] * n0 + ic[0];
    }
    float x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) x[d] = ((float)ic[d] + 0.5f) * l.dx;  // utils.py:42,53

    int tag = SPHB200_TAG_FLUID;
    float T = l.T;
    if (l.wall_axis >= 0) {
      const int iw = ic[l.wall_axis];
      const bool lower = iw < l.n_walls;
      if (lower || iw >= l.n[l.wall_axis] - l.n_walls) tag = SPHB200_TAG_SOLID_WALL;
      if (lower && x[0] < l.hot_hi && x[0] > l.hot_lo) {  // cases/ht.py:90-97
        tag = SPHB200_TAG_DIRICHLET_WALL;
        T = l.T_hot;
      }
    }

    float u[3];
    case_velocity<DIM>(l.velocity, x, u);

    const sphb200_state& o = a.out;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (o.r) o.r[row * DIM + d] = x[d];
      if (o.u) o.u[row * DIM + d] = u[d];
      if (o.v) o.v[row * DIM + d] = u[d];
    }
    if (o.rho) o.rho[row] = l.rho;
    if (o.p) o.p[row] = l.p;
    if (o.mass) o.mass[row] = l.mass;
    if (o.eta) o.eta[row] = l.eta;
    if (o.T) o.T[row] = T;
    if (o.kappa) o.kappa[row] = l.kappa;
    if (o.Cp) o.Cp[row] = l.Cp;
    if (o.tag) o.tag[row] = tag;
    if (a.ids) a.ids[row] = (int32_t)id;
  }
}


template <int DIM>
__global__ void __launch_bounds__(256) k_eval_velocity(long long n, int kind,
                                                       const float* __restrict__ r,
                                                       float* __restrict__ u, float* __restrict__ v) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    float x[3] = {0.f, 0.f, 0.f}, w[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d) x[d] = r[p * DIM + d];
    case_velocity<DIM>(kind, x, w);
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (u) u[p * DIM + d] = w[d];
      if (v) v[p * DIM + d] = w[d];
    }
  }
}

// ---- position noise (case_setup.py:138-144, utils.py:120-125) --------------------------------
// Philox-4x32-10 (Salmon et al. 2011): counter = (row, 0, 0, 0), key = seed.  Counter-based, so
// a particle's deviates depend on its ROW in the full lattice only -- the same start for any
// slab decomposition.  Not jax.random's threefry stream: same distribution, other numbers.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

template <int DIM>
__global__ void __launch_bounds__(256) k_add_noise(long long n, float* __restrict__ r,
                                                   const int32_t* __restrict__ tag,
                                                   const int32_t* __restrict__ ids, float std,
                                                   unsigned long long seed, float bx, float by,
                                                   float bz) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float box[3] = {bx, by, bz};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    if (tag && tag[p] != SPHB200_TAG_FLUID) continue;  // get_noise_masked: fluid only
    const unsigned long long row = ids ? (unsigned long long)(unsigned)ids[p] : (unsigned long long)p;
    const uint4 x = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), 0u, 0u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    // Box-Muller on two pairs: u1 in (0, 1], u2 in [0, 1)
    const float a0 = ((float)(x.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float a1 = (float)(x.y >> 8) * (1.0f / 16777216.0f);
    const float b0 = ((float)(x.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float b1 = (float)(x.w >> 8) * (1.0f / 16777216.0f);
    float s0, c0, s1, c1;
    sincosf(6.2831853071795864769f * a1, &s0, &c0);
    sincosf(6.2831853071795864769f * b1, &s1, &c1);
    const float m0 = sqrtf(-2.0f * logf(a0)), m1 = sqrtf(-2.0f * logf(b0));
    const float z[3] = {m0 * c0, m0 * s0, m1 * c1};
    (void)s1;
#pragma unroll
    for (int d = 0; d < DIM; ++d)  // shift_fn: jnp.mod(r + dr, side), space.py:207-209
      r[p * DIM + d] = mod_side(__fadd_rn(r[p * DIM + d], __fmul_rn(std, z[d])), box[d]);
  }
}

}  // namespace sphb200
