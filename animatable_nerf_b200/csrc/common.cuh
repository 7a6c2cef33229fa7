// Shared helpers of libaninerf_b200: error reporting, launch accounting, small device math.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/aninerf_b200.h"

namespace aninerf {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}

#define ANI_CHECK_ARG(cond)                                                          \
  do {                                                                               \
    if (!(cond)) return ::aninerf::fail(ANINERF_EINVAL, "%s: invalid argument: %s", __func__, #cond); \
  } while (0)

#define ANI_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) return ::aninerf::fail(ANINERF_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// call right after a <<<>>> launch
#define ANI_LAUNCHED()                                                               \
  do {                                                                               \
    ::aninerf::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) return ::aninerf::fail(ANINERF_ECUDA, "%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Device pieces shared by the stage kernels and the fused kernels.  All "bit-exact" arithmetic
// is spelled with _rn intrinsics so that nvcc can neither contract a*b+c into an FMA nor
// reassociate; the op order follows the CPU reference (torch elementwise ops, ATen grid_sampler_3d).
// ---------------------------------------------------------------------------------------------

// z = near*(1-t) + far*t  (tpose_renderer.py:27) -- `one_minus_t` is the fp32 value (1 - t).
__device__ __forceinline__ float z_lerp(float near, float far, float t, float one_minus_t) {
  return __fadd_rn(__fmul_rn(near, one_minus_t), __fmul_rn(far, t));
}

// stratified jitter (tpose_renderer.py:30-37): mids/upper/lower then lower + (upper-lower)*u
__device__ __forceinline__ float z_jitter(float z_prev, float z, float z_next, bool first, bool last, float u) {
  float upper = last ? z : __fmul_rn(0.5f, __fadd_rn(z_next, z));
  float lower = first ? z : __fmul_rn(0.5f, __fadd_rn(z, z_prev));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
}

struct RigidFrame {   // world->pose: (p - Th) @ R
  float R[9];
  float Th[3];
};

// torch.matmul((1,n,3),(1,3,3)) on CPU is bit-equal to this FMA chain (probe P9)
__device__ __forceinline__ void world_to_pose(const RigidFrame &f, float x, float y, float z, float &px, float &py,
                                              float &pz) {
  float dx = __fsub_rn(x, f.Th[0]), dy = __fsub_rn(y, f.Th[1]), dz = __fsub_rn(z, f.Th[2]);
  px = __fmaf_rn(dz, f.R[6], __fmaf_rn(dy, f.R[3], __fmul_rn(dx, f.R[0])));
  py = __fmaf_rn(dz, f.R[7], __fmaf_rn(dy, f.R[4], __fmul_rn(dx, f.R[1])));
  pz = __fmaf_rn(dz, f.R[8], __fmaf_rn(dy, f.R[5], __fmul_rn(dx, f.R[2])));
}

struct VolumeGrid {   // blend-weight volume (X,Y,Z,C) with the reference's normalisation bounds
  float lo[3];
  float ext[3];       // max - min  (blend_utils.py:133)
  int dim[3];         // X, Y, Z
};

// Trilinear corner set of ATen grid_sampler_3d (align_corners=True, padding border) for a point p.
// Axis naming: volume dim 0 (X) is ATen's depth "z", dim 1 (Y) its "y", dim 2 (Z) its "x"
// (blend_utils.py:139 reorders xyz -> zyx).  w[8], off[8] in ATen's accumulation order
// tnw,tne,tsw,tse,bnw,bne,bsw,bse; off = voxel index (x*Y+y)*Z+z or -1 if that corner is outside.
__device__ __forceinline__ void trilinear_corners(const VolumeGrid &g, float px, float py, float pz, float w[8],
                                                  int off[8]) {
  float c[3] = {px, py, pz};
  float fl[3], fr[3];
  int i0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float n = __fdiv_rn(__fsub_rn(c[a], g.lo[a]), g.ext[a]);          // (p - min) / (max - min)
    n = __fsub_rn(__fmul_rn(n, 2.0f), 1.0f);                          // * 2 - 1
    float lim = (float)(g.dim[a] - 1);
    float u = __fmul_rn(__fmul_rn(__fadd_rn(n, 1.0f), 0.5f), lim);    // ((c+1)/2) * (size-1); x*0.5 == x/2 exactly
    u = fminf(lim, fmaxf(u, 0.0f));                                   // border clip
    float f = floorf(u);
    i0[a] = (int)f;
    fl[a] = __fsub_rn(__fadd_rn(f, 1.0f), u);   // weight of the lower index:  (i+1) - u
    fr[a] = __fsub_rn(u, f);                    // weight of the upper index:   u - i
  }
  // ATen axes: ix <-> volume Z (a=2), iy <-> Y (a=1), iz <-> X (a=0).  weight = (wx*wy)*wz
  const int Y = g.dim[1], Z = g.dim[2];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int ex = k & 1, sy = (k >> 1) & 1, bz = (k >> 2) & 1;   // east (x+1), south (y+1), bottom (z+1)
    float wx = ex ? fr[2] : fl[2];
    float wy = sy ? fr[1] : fl[1];
    float wz = bz ? fr[0] : fl[0];
    w[k] = __fmul_rn(__fmul_rn(wx, wy), wz);
    int xi = i0[0] + bz, yi = i0[1] + sy, zi = i0[2] + ex;
    bool in = xi < g.dim[0] && yi < Y && zi < Z;            // lower bounds hold after the clip
    off[k] = in ? (xi * Y + yi) * Z + zi : -1;
  }
}

// The same corners for the distance-plane mask passes (FINITE volumes): an outside corner is not flagged but CLAMPED to the last
// voxel of its axis.  A corner is outside only when floor(u) == size - 1, i.e. u == size - 1 exactly, and then its weight u - floor(u)
// is exactly 0: `acc + v * 0` adds +0 to a sum that starts at +0, so the result is bit-identical to skipping the corner (as ATen
// does) while the per-corner bounds tests, selects and predicated loads -- a quarter of the mask kernel's instructions -- go away.
__device__ __forceinline__ void trilinear_corners_clamped(const VolumeGrid &g, float px, float py, float pz, float w[8], int off[8]) {
  float c[3] = {px, py, pz};
  float fl[3], fr[3];
  int i0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float n = __fdiv_rn(__fsub_rn(c[a], g.lo[a]), g.ext[a]);
    n = __fsub_rn(__fmul_rn(n, 2.0f), 1.0f);
    float lim = (float)(g.dim[a] - 1);
    float u = __fmul_rn(__fmul_rn(__fadd_rn(n, 1.0f), 0.5f), lim);
    u = fminf(lim, fmaxf(u, 0.0f));
    float f = floorf(u);
    i0[a] = (int)f;
    fl[a] = __fsub_rn(__fadd_rn(f, 1.0f), u);
    fr[a] = __fsub_rn(u, f);
  }
  const int Y = g.dim[1], Z = g.dim[2];
  const int base = (i0[0] * Y + i0[1]) * Z + i0[2];
  const int dx = i0[0] + 1 < g.dim[0] ? Y * Z : 0, dy = i0[1] + 1 < Y ? Z : 0, dz = i0[2] + 1 < Z ? 1 : 0;
  // weight = (wx * wy) * wz with ATen's association; x <-> volume Z (a = 2), y <-> Y (a = 1), z <-> X (a = 0)
  const float wxy[4] = {__fmul_rn(fl[2], fl[1]), __fmul_rn(fr[2], fl[1]), __fmul_rn(fl[2], fr[1]), __fmul_rn(fr[2], fr[1])};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ex = k & 1, sy = (k >> 1) & 1, bz = (k >> 2) & 1;
    w[k] = __fmul_rn(wxy[k & 3], bz ? fr[0] : fl[0]);
    off[k] = base + (bz ? dx : 0) + (sy ? dy : 0) + (ex ? dz : 0);
  }
}

}  // namespace aninerf
