// K5-motion, full native contract (r02).  The generator only ever calls the reference's extension one way
// (acquisition: interp_psf = false, no masks, no weights; reconstruction: interp_psf = true, equalize, no masks —
// motion.cu holds the tuned kernels for exactly that).  The pybind modules the reference builds expose more
// (SURVEY.md 8(b)): optional volume / slice masks, the per-pixel weight output, both PSF modes in both directions
// (svort/slice_acquisition/slice_acq_cuda.cpp:61-79,105-124 -> slice_acq_cuda_kernel.cu:17-171,472-693) and the two
// rigid-transform conversions (svort/transform/transform_convert_cuda.cpp:27-51 -> transform_convert_cuda_kernel.cu:
// 14-65,190-264).  This file provides them behind the C-ABI so that the extension can be replaced as a whole:
//   fsg_slice_acq_forward_ex / fsg_slice_acq_adjoint_ex   any combination of the options, thread per slice pixel over
//                                                         the compact list of non-zero PSF taps (rotated offsets
//                                                         staged once per 16x16 pixel tile, pixels farther than the
//                                                         PSF radius from the volume culled)
//   fsg_axisangle2mat / fsg_mat2axisangle                 (n, 6) <-> (n, 3, 4)
// Semantics kept from the reference: a pixel is written only when its weight is > 0; the acquisition's bounds test
// is half-open [0, S-1); `round` is C round (half away from zero); in the reconstruction the normalisation weight
// ignores the volume mask and pixels whose weight is < 0.5 are dropped; equalisation divides where the weight is > 0.
#include <stdlib.h>

#include "common.cuh"

namespace fsg {

constexpr int EX_TILE = 16;

struct ExGeom {
  float r11, r12, r13, r21, r22, r23, r31, r32, r33;
  float xc, yc, zc;
};

// centre of pixel (ix, iy) in voxel coordinates; the reference mixes double constants ((w - 1) / 2.) into float
// expressions, which promotes those sub-expressions to double
__device__ __forceinline__ ExGeom ex_geom(const float* __restrict__ t, int ix, int iy, int h, int w, int D, int H, int W, float res) {
  ExGeom g;
  g.r11 = t[0]; g.r12 = t[1]; g.r13 = t[2];
  g.r21 = t[4]; g.r22 = t[5]; g.r23 = t[6];
  g.r31 = t[8]; g.r32 = t[9]; g.r33 = t[10];
  const float px = (float)(((double)ix - (w - 1) / 2.) * (double)res + (double)t[3]);
  const float py = (float)(((double)iy - (h - 1) / 2.) * (double)res + (double)t[7]);
  const float pz = t[11];
  g.xc = (float)((double)(g.r11 * px + g.r12 * py + g.r13 * pz) + (W - 1) / 2.);
  g.yc = (float)((double)(g.r21 * px + g.r22 * py + g.r23 * pz) + (H - 1) / 2.);
  g.zc = (float)((double)(g.r31 * px + g.r32 * py + g.r33 * pz) + (D - 1) / 2.);
  return g;
}

// PSF value interpolated at the position of voxel (xr, yr, zr) rotated back into the PSF grid; false when it falls
// off the grid (half-open bounds like the volume test)
__device__ __forceinline__ bool ex_psf_at(const ExGeom& g, const float* __restrict__ psf, int dp, int hp, int wp, float xr, float yr, float zr, float& out) {
  const float dx = xr - g.xc, dy = yr - g.yc, dz = zr - g.zc;
  const float xp = (g.r11 * dx + g.r21 * dy + g.r31 * dz) + 0.5f * (float)(wp - 1);
  const float yp = (g.r12 * dx + g.r22 * dy + g.r32 * dz) + 0.5f * (float)(hp - 1);
  const float zp = (g.r13 * dx + g.r23 * dy + g.r33 * dz) + 0.5f * (float)(dp - 1);
  if (xp < 0.f || yp < 0.f || zp < 0.f || xp >= (float)(wp - 1) || yp >= (float)(hp - 1) || zp >= (float)(dp - 1)) return false;
  const float fx = floorf(xp), fy = floorf(yp), fz = floorf(zp);
  const float wx = xp - fx, wy = yp - fy, wz = zp - fz;
  const float* q = psf + ((int)fz * wp * hp + (int)fy * wp + (int)fx);
  float v = 0.f;
  v += (1 - wx) * (1 - wy) * (1 - wz) * q[0];
  v += wx * (1 - wy) * (1 - wz) * q[1];
  v += (1 - wx) * wy * (1 - wz) * q[wp];
  v += (1 - wx) * (1 - wy) * wz * q[wp * hp];
  v += wx * wy * (1 - wz) * q[1 + wp];
  v += wx * (1 - wy) * wz * q[1 + wp * hp];
  v += (1 - wx) * wy * wz * q[wp + wp * hp];
  v += wx * wy * wz * q[wp + wp * hp + 1];
  out = v;
  return true;
}

// rotated tap offsets of slice `in` into shared memory: (R p, psf value)
__device__ __forceinline__ void ex_stage_taps(const float* __restrict__ t, const float4* __restrict__ taps, int ntaps, float4* s_tap) {
  for (int p = threadIdx.y * EX_TILE + threadIdx.x; p < ntaps; p += EX_TILE * EX_TILE) {
    const float4 q = taps[p];
    float x = t[0] * q.x;
    x = x + t[1] * q.y;
    x = x + t[2] * q.z;
    float y = t[4] * q.x;
    y = y + t[5] * q.y;
    y = y + t[6] * q.z;
    float z = t[8] * q.x;
    z = z + t[9] * q.y;
    z = z + t[10] * q.z;
    s_tap[p] = make_float4(x, y, z, q.w);
  }
}

struct ExCorner {
  int off;
  float w;
};
// the eight (offset, weight) pairs of a trilinear footprint in the reference's accumulation order
__device__ __forceinline__ void ex_corners(float wx, float wy, float wz, int Sy, int Sz, ExCorner (&c)[8]) {
  c[0] = {0, (1 - wx) * (1 - wy) * (1 - wz)};
  c[1] = {1, wx * (1 - wy) * (1 - wz)};
  c[2] = {Sy, (1 - wx) * wy * (1 - wz)};
  c[3] = {Sz, (1 - wx) * (1 - wy) * wz};
  c[4] = {1 + Sy, wx * wy * (1 - wz)};
  c[5] = {1 + Sz, wx * (1 - wy) * wz};
  c[6] = {Sy + Sz, (1 - wx) * wy * wz};
  c[7] = {Sy + Sz + 1, wx * wy * wz};
}

template <bool INTERP_PSF>
__global__ void __launch_bounds__(EX_TILE* EX_TILE) slice_fwd_ex_kernel(const float* __restrict__ transforms, const float* __restrict__ vol, const uint8_t* __restrict__ vol_mask,
                                                                        const uint8_t* __restrict__ slices_mask, const float* __restrict__ psf, int dp, int hp, int wp,
                                                                        const float4* __restrict__ taps, int ntaps, float radius, float* __restrict__ slices,
                                                                        float* __restrict__ slices_weight, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;
  ex_stage_taps(t, taps, ntaps, s_tap);
  __syncthreads();
  const int ix = blockIdx.x * EX_TILE + threadIdx.x, iy = blockIdx.y * EX_TILE + threadIdx.y;
  if (ix >= w || iy >= h) return;
  const size_t idx = ((size_t)in * h + iy) * w + ix;
  if (slices_mask != nullptr && !slices_mask[idx]) return;
  const ExGeom g = ex_geom(t, ix, iy, h, w, D, H, W, res);
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  if (g.xc + radius < 0.f || g.yc + radius < 0.f || g.zc + radius < 0.f || g.xc - radius >= mx || g.yc - radius >= my || g.zc - radius >= mz) return;
  const int Sy = W, Sz = H * W;
  float val = 0.f, weight = 0.f;
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    if (INTERP_PSF) {  // nearest voxel, PSF interpolated at that voxel
      const float xr = roundf(x), yr = roundf(y), zr = roundf(z);
      const int iv = (int)zr * Sz + (int)yr * Sy + (int)xr;
      if (vol_mask != nullptr && !vol_mask[iv]) continue;
      float pv;
      if (!ex_psf_at(g, psf, dp, hp, wp, xr, yr, zr, pv)) continue;
      val += pv * vol[iv];
      weight += pv;
    } else {  // trilinear sample weighted by the tap
      const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
      const int iv = (int)fz * Sz + (int)fy * Sy + (int)fx;
      ExCorner c[8];
      ex_corners(x - fx, y - fy, z - fz, Sy, Sz, c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (vol_mask == nullptr || vol_mask[iv + c[k].off]) {
          const float pw = c[k].w * q.w;
          val += pw * vol[iv + c[k].off];
          weight += pw;
        }
      }
    }
  }
  if (weight > 0.f) {
    slices[idx] = __fdiv_rn(val, weight);
    if (slices_weight != nullptr) slices_weight[idx] = weight;
  }
}

template <bool INTERP_PSF>
__global__ void __launch_bounds__(EX_TILE* EX_TILE) slice_adj_ex_kernel(const float* __restrict__ transforms, const float* __restrict__ psf, int dp, int hp, int wp,
                                                                        const float4* __restrict__ taps, int ntaps, float radius, const float* __restrict__ slices,
                                                                        const uint8_t* __restrict__ slices_mask, const uint8_t* __restrict__ vol_mask, float* __restrict__ vol,
                                                                        float* __restrict__ vol_weight, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;
  ex_stage_taps(t, taps, ntaps, s_tap);
  __syncthreads();
  const int ix = blockIdx.x * EX_TILE + threadIdx.x, iy = blockIdx.y * EX_TILE + threadIdx.y;
  if (ix >= w || iy >= h) return;
  const size_t idx = ((size_t)in * h + iy) * w + ix;
  if (slices_mask != nullptr && !slices_mask[idx]) return;
  const ExGeom g = ex_geom(t, ix, iy, h, w, D, H, W, res);
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  if (g.xc + radius < 0.f || g.yc + radius < 0.f || g.zc + radius < 0.f || g.xc - radius >= mx || g.yc - radius >= my || g.zc - radius >= mz) return;
  const float s = slices[idx];
  const int Sy = W, Sz = H * W;
  // pass 1: normalisation weight of the pixel (the volume mask plays no role here)
  float weight = 0.f;
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    if (INTERP_PSF) {
      float pv;
      if (ex_psf_at(g, psf, dp, hp, wp, roundf(x), roundf(y), roundf(z), pv)) weight += pv;
    } else {
      weight += q.w;
    }
  }
  if (weight < 0.5f) return;  // border
  // pass 2: scatter psf / weight * s
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    if (INTERP_PSF) {
      const float xr = roundf(x), yr = roundf(y), zr = roundf(z);
      float pv;
      if (!ex_psf_at(g, psf, dp, hp, wp, xr, yr, zr, pv)) continue;
      pv = __fdiv_rn(pv, weight);
      const int iv = (int)zr * Sz + (int)yr * Sy + (int)xr;
      if (vol_mask != nullptr && !vol_mask[iv]) continue;
      atomicAdd(vol + iv, pv * s);
      if (vol_weight != nullptr) atomicAdd(vol_weight + iv, pv);
    } else {
      const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
      const int iv = (int)fz * Sz + (int)fy * Sy + (int)fx;
      const float pn = __fdiv_rn(q.w, weight);
      ExCorner c[8];
      ex_corners(x - fx, y - fy, z - fz, Sy, Sz, c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (vol_mask == nullptr || vol_mask[iv + c[k].off]) {
          const float pw = c[k].w * pn;
          atomicAdd(vol + iv + c[k].off, pw * s);
          if (vol_weight != nullptr) atomicAdd(vol_weight + iv + c[k].off, pw);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) equalize_ex_kernel(float* __restrict__ vol, const float* __restrict__ wgt, unsigned n) {
  for (unsigned v = blockIdx.x * 256 + threadIdx.x; v < n; v += gridDim.x * 256) {
    const float a = wgt[v];
    if (a > 0.f) vol[v] = __fdiv_rn(vol[v], a);
  }
}

// ---------------------------------------------------------------------------------- rigid transform conversions
constexpr float EX_EPS = 1e-6f;

// (n, 6) axis-angle (radians) + translation -> (n, 3, 4): Rodrigues' formula, first-order form below the threshold
__global__ void __launch_bounds__(128) axisangle2mat_kernel(const float* __restrict__ ax, float* __restrict__ mat, int n) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  const float* a = ax + 6 * i;
  float* m = mat + 12 * i;
  float x = a[0], y = a[1], z = a[2];
  const float th2 = x * x + y * y + z * z;
  if (th2 > EX_EPS) {
    const float th = sqrtf(th2);
    x /= th;
    y /= th;
    z /= th;
    const float s = sinf(th), c = cosf(th), o = 1.f - c;
    m[0] = c + x * x * o;
    m[1] = x * y * o - z * s;
    m[2] = y * s + x * z * o;
    m[4] = z * s + x * y * o;
    m[5] = c + y * y * o;
    m[6] = -x * s + y * z * o;
    m[8] = -y * s + x * z * o;
    m[9] = x * s + y * z * o;
    m[10] = c + z * z * o;
  } else {
    m[0] = 1.f; m[1] = -z; m[2] = y;
    m[4] = z; m[5] = 1.f; m[6] = -x;
    m[8] = -y; m[9] = x; m[10] = 1.f;
  }
  m[3] = a[3];
  m[7] = a[4];
  m[11] = a[5];
}

// (n, 3, 4) -> (n, 6) through the unit quaternion with the largest well-conditioned component
__global__ void __launch_bounds__(128) mat2axisangle_kernel(const float* __restrict__ mat, float* __restrict__ ax, int n) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= n) return;
  const float* m = mat + 12 * i;
  const float r00 = m[0], r01 = m[1], r02 = m[2], r10 = m[4], r11 = m[5], r12 = m[6], r20 = m[8], r21 = m[9], r22 = m[10];
  const bool d2 = r22 < EX_EPS, d01 = r00 > r11, d0n1 = r00 < -r11;
  float qw, qx, qy, qz;
  if (!d2 && !d0n1) {
    const float s = 2.f * sqrtf(r00 + r11 + r22 + 1.f);
    qw = 0.25f * s; qx = (r21 - r12) / s; qy = (r02 - r20) / s; qz = (r10 - r01) / s;
  } else if (d2 && d01) {
    const float s = 2.f * sqrtf(r00 - r11 - r22 + 1.f);
    qw = (r21 - r12) / s; qx = 0.25f * s; qy = (r01 + r10) / s; qz = (r02 + r20) / s;
  } else if (d2 && !d01) {
    const float s = 2.f * sqrtf(r11 - r00 - r22 + 1.f);
    qw = (r02 - r20) / s; qx = (r01 + r10) / s; qy = 0.25f * s; qz = (r12 + r21) / s;
  } else {
    const float s = 2.f * sqrtf(r22 - r00 - r11 + 1.f);
    qw = (r10 - r01) / s; qx = (r02 + r20) / s; qy = (r12 + r21) / s; qz = 0.25f * s;
  }
  if (qw < 0.f) {
    qw = -qw; qx = -qx; qy = -qy; qz = -qz;
  }
  const float v2 = qx * qx + qy * qy + qz * qz;
  const float si = sqrtf(v2);
  const float theta = 2.f * atan2f(si, qw);
  const float fac = v2 > EX_EPS ? theta / si : 2.0f / qw;
  float* a = ax + 6 * i;
  a[0] = qx * fac;
  a[1] = qy * fac;
  a[2] = qz * fac;
  a[3] = m[3];
  a[4] = m[7];
  a[5] = m[11];
}

static int check_ex(const char* who, int ntaps, int n, int h, int w, int D, int H, int W, int dp, int hp, int wp) {
  FSG_REQUIRE(n >= 1 && n <= 65535 && h >= 1 && w >= 1 && D >= 2 && H >= 2 && W >= 2, "%s: bad extents", who);
  FSG_REQUIRE(ntaps >= 1 && ntaps <= 4096, "%s: %d PSF taps outside [1, 4096]", who, ntaps);
  FSG_REQUIRE(dp >= 1 && hp >= 1 && wp >= 1, "%s: bad PSF shape", who);
  FSG_REQUIRE((int64_t)D * H * W < ((int64_t)1 << 31) && (int64_t)n * h * w < ((int64_t)1 << 31), "%s: index range exceeds int32", who);
  return 0;
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_slice_acq_forward_ex(const float* transforms, const float* vol, const uint8_t* vol_mask, const uint8_t* slices_mask, const float* psf, int dp, int hp, int wp,
                                        const float* taps, int ntaps, float radius, float* slices, float* slices_weight, int n, int h, int w, int D, int H, int W, float res_slice,
                                        int interp_psf, void* stream) {
  const char* who = "fsg_slice_acq_forward_ex";
  if (int rc = check_ex(who, ntaps, n, h, w, D, H, W, dp, hp, wp)) return rc;
  FSG_REQUIRE(transforms && vol && psf && taps && slices, "%s: NULL pointer", who);
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(taps) & 15) == 0, "%s: taps must be 16-byte aligned", who);
  cudaStream_t s = as_stream(stream);
  const size_t count = (size_t)n * h * w;
  cudaMemsetAsync(slices, 0, sizeof(float) * count, s);
  if (slices_weight) cudaMemsetAsync(slices_weight, 0, sizeof(float) * count, s);
  const dim3 grid((w + EX_TILE - 1) / EX_TILE, (h + EX_TILE - 1) / EX_TILE, n), block(EX_TILE, EX_TILE);
  const size_t smem = sizeof(float4) * ntaps;
  const float4* taps4 = reinterpret_cast<const float4*>(taps);
  if (interp_psf) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(slice_fwd_ex_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    slice_fwd_ex_kernel<true><<<grid, block, smem, s>>>(transforms, vol, vol_mask, slices_mask, psf, dp, hp, wp, taps4, ntaps, radius, slices, slices_weight, h, w, D, H, W, res_slice);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(slice_fwd_ex_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    slice_fwd_ex_kernel<false><<<grid, block, smem, s>>>(transforms, vol, vol_mask, slices_mask, psf, dp, hp, wp, taps4, ntaps, radius, slices, slices_weight, h, w, D, H, W, res_slice);
  }
  return check_launch(who);
}

extern "C" int fsg_slice_acq_adjoint_ex(const float* transforms, const float* psf, int dp, int hp, int wp, const float* taps, int ntaps, float radius, const float* slices,
                                        const uint8_t* slices_mask, const uint8_t* vol_mask, float* vol, float* vol_weight, int n, int h, int w, int D, int H, int W, float res_slice,
                                        int interp_psf, int equalize, void* stream) {
  const char* who = "fsg_slice_acq_adjoint_ex";
  if (int rc = check_ex(who, ntaps, n, h, w, D, H, W, dp, hp, wp)) return rc;
  FSG_REQUIRE(transforms && psf && taps && slices && vol, "%s: NULL pointer", who);
  FSG_REQUIRE(!equalize || vol_weight, "%s: equalize needs the vol_weight buffer", who);
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(taps) & 15) == 0, "%s: taps must be 16-byte aligned", who);
  cudaStream_t s = as_stream(stream);
  const size_t nv = (size_t)D * H * W;
  cudaMemsetAsync(vol, 0, sizeof(float) * nv, s);
  if (vol_weight) cudaMemsetAsync(vol_weight, 0, sizeof(float) * nv, s);
  const dim3 grid((w + EX_TILE - 1) / EX_TILE, (h + EX_TILE - 1) / EX_TILE, n), block(EX_TILE, EX_TILE);
  const size_t smem = sizeof(float4) * ntaps;
  const float4* taps4 = reinterpret_cast<const float4*>(taps);
  if (interp_psf) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(slice_adj_ex_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    slice_adj_ex_kernel<true><<<grid, block, smem, s>>>(transforms, psf, dp, hp, wp, taps4, ntaps, radius, slices, slices_mask, vol_mask, vol, vol_weight, h, w, D, H, W, res_slice);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(slice_adj_ex_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    slice_adj_ex_kernel<false><<<grid, block, smem, s>>>(transforms, psf, dp, hp, wp, taps4, ntaps, radius, slices, slices_mask, vol_mask, vol, vol_weight, h, w, D, H, W, res_slice);
  }
  if (equalize) {
    const size_t want = (nv + 255) / 256;
    equalize_ex_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, s>>>(vol, vol_weight, (unsigned)nv);
  }
  return check_launch(who);
}

extern "C" int fsg_axisangle2mat(const float* axisangle, float* mat, int n, void* stream) {
  FSG_REQUIRE(axisangle && mat && n >= 1, "fsg_axisangle2mat: bad arguments");
  axisangle2mat_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(axisangle, mat, n);
  return check_launch("fsg_axisangle2mat");
}

extern "C" int fsg_mat2axisangle(const float* mat, float* axisangle, int n, void* stream) {
  FSG_REQUIRE(axisangle && mat && n >= 1, "fsg_mat2axisangle: bad arguments");
  mat2axisangle_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(mat, axisangle, n);
  return check_launch("fsg_mat2axisangle");
}
