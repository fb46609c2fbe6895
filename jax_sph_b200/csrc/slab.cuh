// slab.cuh -- device side of the 1-D slab decomposition.
//
// The reference has no multi-device path (jax_sph/simulate.py drives one device); this is
// the decomposition north_star prescribes for it: the periodic box is cut into slabs along
// the slowest-varying cell axis (z in 3D, y in 2D), one slab per GPU.  Because particles are
// kept sorted by (z, y, x) cell, the S boundary layers a neighbour needs are CONTIGUOUS slot
// ranges of the sorted arrays, and a rank's halo layers are contiguous ranges before / after
// its own particles:
//
//     slots:  [base - n_halo_lo, base)   [base, base + n_own)   [base + n_own, + n_halo_hi)
//     cells:  layers [0, S)              layers [S, S + nz_own)  layers [S + nz_own, + S)
//
// Per step:   k_hash (emigrants -> send buffers)            -> exchange 1 (migrants)
//             k_immigrate, scan, reorder, k_halo_pack(A)    -> exchange 2 (halo: positions, state)
//             k_halo_cells + k_halo_unpack, density sweep, k_halo_pack(st) -> exchange 3 ...
// All counts stay on the device (Slab::dn); messages have fixed capacities and carry their
// real counts in a header, so a step never synchronises with the host.  The transport is the
// caller's (NCCL send/recv over NVLink from jax_sph_b200/slab.py).
#pragma once
#include "cells.cuh"
#include "common.cuh"

namespace sphb200 {

// which arrays a halo message carries
enum {
  HX_PT = 1, HX_UM = 2, HX_VV = 4, HX_ST = 8, HX_NW = 16, HX_GE = 32, HX_UT = 64,
  HX_DL0 = 128, HX_DL1 = 256, HX_DL2 = 512,  // Delta-SPH renormalisation matrix rows
  HX_DG1 = 1024,                              // Delta-SPH gradient term H (read from neighbours)
  HX_KC = 2048, HX_CELLS = 4096
};
constexpr int HX_NQ = 11;  // quad arrays a message can carry (bits 0 .. HX_NQ-1)

struct HaloView {
  int* hdr;     // [count, ...]
  int* cells;   // per-cell particle counts of the S layers (HX_CELLS)
  float4* q[HX_NQ]; // pt um vv st nw ge ut dl0 dl1 dl2 dg1, nullptr when not carried
  float2* kc;
};

__host__ __device__ inline size_t halo_bytes(int mask, int cap, int ncl) {
  size_t b = 16;
  if (mask & HX_CELLS) b += ((size_t)ncl * 4 + 15) / 16 * 16;
  for (int i = 0; i < HX_NQ; ++i)
    if (mask & (1 << i)) b += (size_t)cap * 16;
  if (mask & HX_KC) b += (size_t)cap * 8;
  return b;
}

__host__ __device__ inline HaloView halo_view(char* b, int mask, int cap, int ncl) {
  HaloView v;
  v.hdr = reinterpret_cast<int*>(b);
  char* p = b + 16;
  v.cells = nullptr;
  if (mask & HX_CELLS) {
    v.cells = reinterpret_cast<int*>(p);
    p += ((size_t)ncl * 4 + 15) / 16 * 16;
  }
  for (int i = 0; i < HX_NQ; ++i) {
    v.q[i] = nullptr;
    if (mask & (1 << i)) {
      v.q[i] = reinterpret_cast<float4*>(p);
      p += (size_t)cap * 16;
    }
  }
  v.kc = (mask & HX_KC) ? reinterpret_cast<float2*>(p) : nullptr;
  return v;
}

__device__ __forceinline__ float4* frame_quad(const Frame& f, int i) {
  switch (i) {
    case 0: return f.pt;
    case 1: return f.um;
    case 2: return f.vv;
    case 3: return f.st;
    case 4: return f.nw;
    case 5: return f.ge;
    case 6: return f.ut;
    case 7: return f.dl0;
    case 8: return f.dl1;
    case 9: return f.dl2;
    default: return f.dg1;
  }
}

struct SlabGeom {
  int halo_cap;
  int ncl;       // cells in S layers
  int c_own_lo;  // first cell of the first own layer
  int c_own_hi;  // first cell past the last own layer
  int own_cap, cap_total;
};

// header of the emigrant buffers (after k_hash)
__global__ void k_mig_header(Slab sl) {
  mig_view(sl.mig_lo, sl.mig_cap).hdr[0] = min(sl.dn[DN_EMIG_LO], sl.mig_cap);
  mig_view(sl.mig_hi, sl.mig_cap).hdr[0] = min(sl.dn[DN_EMIG_HI], sl.mig_cap);
}

// Immigrants: append the received records after the own particles of the current frame and
// hash them (they were integrated by the rank they come from).  blockIdx.y: 0 = records from
// the lower neighbour, 1 = from the upper one.
template <int DIM>
__global__ void __launch_bounds__(256) k_immigrate(Grid g, Slab sl, SlabGeom sg, Frame f,
                                                   const char* from_lo, const char* from_hi,
                                                   int* __restrict__ key, int* __restrict__ rnk,
                                                   int* __restrict__ count,
                                                   unsigned* __restrict__ err) {
  const MigView m0 = mig_view(const_cast<char*>(from_lo), sl.mig_cap);
  const MigView m1 = mig_view(const_cast<char*>(from_hi), sl.mig_cap);
  const int n0 = min(m0.hdr[0], sl.mig_cap), n1 = min(m1.hdr[0], sl.mig_cap);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool second = blockIdx.y == 1;
  if (i == 0 && !second) sl.dn[DN_IN] = n0 + n1;
  if (i >= (second ? n1 : n0)) return;
  const MigView& m = second ? m1 : m0;
  const int n_own = sl.dn[DN_OWN];
  const int p = sl.base + n_own + (second ? n0 : 0) + i;
  if (p >= sg.cap_total) {
    atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
    return;
  }
  const float4 pt = m.pt[i];
  f.pt[p] = pt;
  f.um[p] = m.um[i];
  f.vv[p] = m.vv[i];
  f.st[p] = m.st[i];
  f.du[p] = m.du[i];
  f.id[p] = m.id[i];
  if (f.kc) f.kc[p] = m.kc[i];
  if (f.nw) f.nw[p] = m.nw[i];
  if (f.ge) f.ge[p] = m.ge[i];
  float r[3] = {pt.x, pt.y, pt.z};
  int c[3];
  const int cell = cell_of<DIM>(g, r, c);
  const int ax = DIM - 1;
  if (cell < 0 || c[ax] < g.own_lo[ax] || c[ax] >= g.own_hi[ax]) {
    atomicOr(err, SPHB200_ERR_SLAB_MIGRATION);
    key[p] = -1;
    return;
  }
  key[p] = cell;
  rnk[p] = atomicAdd(&count[cell], 1);
}

// after the sort: the scan total is the new own count; reset the per-step counters
__global__ void k_slab_after_sort(Grid g, Slab sl, SlabGeom sg, const int* __restrict__ start,
                                  unsigned* __restrict__ err, const int* __restrict__ gate) {
  if (gate != nullptr && *gate == 0) return;
  const int n_new = start[g.ncells] - sl.base;
  if (n_new > sg.own_cap) atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
  sl.dn[DN_OWN] = n_new;
  sl.dn[DN_IN] = 0;
  sl.dn[DN_EMIG_LO] = 0;
  sl.dn[DN_EMIG_HI] = 0;
}

// Boundary layers -> message.  blockIdx.y: 0 = to the lower neighbour (my first S own
// layers), 1 = to the upper neighbour (my last S own layers).
__global__ void __launch_bounds__(256) k_halo_pack(Slab sl, SlabGeom sg, Frame f, int mask,
                                                   const int* __restrict__ start, char* to_lo,
                                                   char* to_hi, unsigned* __restrict__ err) {
  const bool up = blockIdx.y == 1;
  const int c0 = up ? sg.c_own_hi - sg.ncl : sg.c_own_lo;
  const int s0 = start[c0];
  int n = start[c0 + sg.ncl] - s0;
  if (n > sg.halo_cap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
    n = sg.halo_cap;
  }
  const HaloView v = halo_view(up ? to_hi : to_lo, mask, sg.halo_cap, sg.ncl);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (tid == 0) {
    v.hdr[0] = n;
    sl.dn[up ? DN_SEND_HI : DN_SEND_LO] = n;
  }
  if (mask & HX_CELLS)
    for (int j = tid; j < sg.ncl; j += nth) v.cells[j] = start[c0 + j + 1] - start[c0 + j];
#pragma unroll
  for (int a = 0; a < HX_NQ; ++a) {
    if (!(mask & (1 << a))) continue;
    const float4* src = frame_quad(f, a) + s0;
    for (int i = tid; i < n; i += nth) v.q[a][i] = src[i];
  }
  if (mask & HX_KC)
    for (int i = tid; i < n; i += nth) v.kc[i] = f.kc[s0 + i];
}

// Cell table of the halo layers from the received per-cell counts (one block per side).
// blockIdx.x: 0 = message from the lower neighbour -> my layers [0, S), right-aligned so that
// it ends at `base`; 1 = from the upper neighbour -> my layers past the own range.
__global__ void __launch_bounds__(1024) k_halo_cells(Grid g, Slab sl, SlabGeom sg, int mask,
                                                     const char* from_lo, const char* from_hi,
                                                     int* __restrict__ start,
                                                     unsigned* __restrict__ err) {
  __shared__ int sh[32];
  __shared__ int carry_s;
  const bool up = blockIdx.x == 1;
  const HaloView v = halo_view(const_cast<char*>(up ? from_hi : from_lo), mask, sg.halo_cap, sg.ncl);
  int n = v.hdr[0];
  if (n < 0 || n > sg.halo_cap) {
    if (threadIdx.x == 0) atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
    n = max(0, min(n, sg.halo_cap));
  }
  const int n_own = sl.dn[DN_OWN];
  const int dst0 = up ? sl.base + n_own : sl.base - n;
  const int c0 = up ? sg.c_own_hi : 0;
  if (up && dst0 + n > sg.cap_total && threadIdx.x == 0) atomicOr(err, SPHB200_ERR_SLAB_OVERFLOW);
  if (threadIdx.x == 0) {
    carry_s = 0;
    sl.dn[up ? DN_HALO_HI : DN_HALO_LO] = n;
    if (up) start[g.ncells] = dst0 + n;
  }
  __syncthreads();
  for (int b = 0; b < sg.ncl; b += blockDim.x) {
    const int j = b + threadIdx.x;
    const int cnt = j < sg.ncl ? v.cells[j] : 0;
    int tot;
    const int inc = block_incl_scan(cnt, sh, &tot);
    const int carry = carry_s;
    if (j < sg.ncl) start[c0 + j] = dst0 + carry + inc - cnt;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

// Message -> halo slots.  blockIdx.y as in k_halo_cells (0 = from lower, 1 = from upper).
__global__ void __launch_bounds__(256) k_halo_unpack(Slab sl, SlabGeom sg, Frame f, int mask,
                                                     const char* from_lo, const char* from_hi) {
  const bool up = blockIdx.y == 1;
  const HaloView v = halo_view(const_cast<char*>(up ? from_hi : from_lo), mask, sg.halo_cap, sg.ncl);
  const int n = max(0, min(v.hdr[0], sg.halo_cap));
  const int dst0 = up ? sl.base + sl.dn[DN_OWN] : sl.base - n;
  if (up && dst0 + n > sg.cap_total) return;  // flagged by k_halo_cells
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
#pragma unroll
  for (int a = 0; a < HX_NQ; ++a) {
    if (!(mask & (1 << a))) continue;
    float4* dst = frame_quad(f, a) + dst0;
    for (int i = tid; i < n; i += nth) dst[i] = v.q[a][i];
  }
  if (mask & HX_KC)
    for (int i = tid; i < n; i += nth) f.kc[dst0 + i] = v.kc[i];
}

}  // namespace sphb200
