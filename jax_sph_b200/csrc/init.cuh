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

namespace sphb200 {

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

    float u[3] = {0.f, 0.f, 0.f};
    if (l.velocity == SPHB200_VEL_TGV2D) {  // cases/tgv.py:37-43
      float sx, cx, sy, cy;
      sincosf(6.2831853071795864769f * x[0], &sx, &cx);
      sincosf(6.2831853071795864769f * x[1], &sy, &cy);
      u[0] = -1.0f * cx * sy;
      u[1] = sx * cy;
    } else if (l.velocity == SPHB200_VEL_TGV3D) {  // cases/tgv.py:45-51
      float sx, cx, sy, cy;
      sincosf(x[0], &sx, &cx);
      sincosf(x[1], &sy, &cy);
      const float cz = cosf(x[2]);
      u[0] = sx * cy * cz;
      u[1] = -cx * sy * cz;
    }

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

}  // namespace sphb200
