// Device helpers shared by the warp kernels (warp.cu: generic + full-z fast path, warp_tile.cu:
// TMA-staged cubic tiles).
#pragma once
#include "common.cuh"

namespace fsg {

#ifndef FSG_WARP_WX
#define FSG_WARP_WX 8
#endif
#ifndef FSG_WARP_WY
#define FSG_WARP_WY 4
#endif
constexpr int WX = FSG_WARP_WX, WY = FSG_WARP_WY;  // tile of (x, y) rows a block stages the control grids for
constexpr int WARP_THREADS = 256;
constexpr int MAX_FZ = 32;  // control-grid extent along z kept in smem (reference: <= 0.06*S)
constexpr int MAX_BZ = 16;  // bias-grid extent along z (reference: <= 0.02*S)
constexpr float MAGIC = 8388608.0f;  // 2^23: x + MAGIC has a unit ulp for 0 <= x < 2^23

enum { PASS_SHIFT = 0, PASS_WARP = 1, PASS_COORDS = 2 };

__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lerp_fma(float a, float b, float w) { return __fmaf_rn(w, __fsub_rn(b, a), a); }

// Affine part of the deformation, copied out of the kernel-parameter space once per thread.
struct Affine {
  float a[9], c[3];
  __device__ __forceinline__ explicit Affine(const fsg_warp_job& job) {
#pragma unroll
    for (int q = 0; q < 9; ++q) a[q] = job.A[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) c[q] = job.c2[q];
  }
};
// Clamped (not yet shifted) sample coordinate of one voxel from its centred position + field.
__device__ __forceinline__ void affine_clamp(const Affine& t, float x1, float y1, float z1, float mx, float my, float mz, float& ii, float& jj, float& kk) {
  ii = add_rn(add_rn(add_rn(mul_rn(t.a[0], x1), mul_rn(t.a[1], y1)), mul_rn(t.a[2], z1)), t.c[0]);
  jj = add_rn(add_rn(add_rn(mul_rn(t.a[3], x1), mul_rn(t.a[4], y1)), mul_rn(t.a[5], z1)), t.c[1]);
  kk = add_rn(add_rn(add_rn(mul_rn(t.a[6], x1), mul_rn(t.a[7], y1)), mul_rn(t.a[8], z1)), t.c[2]);
  ii = fminf(fmaxf(ii, 0.f), mx);
  jj = fminf(fmaxf(jj, 0.f), my);
  kk = fminf(fmaxf(kk, 0.f), mz);
}

}  // namespace fsg
