// K5-motion — slice acquisition (forward) and PSF reconstruction (adjoint) of SimulateMotion.
//
// Replaces the reference's JIT-built extension for the two call sites of the generation path
// (svort/slice_acquisition/slice_acq_cuda_kernel.cu:17-171 forward with interp_psf = false,
// called by Scanner.scan, simulate_reco.py:386-407; :472-693 adjoint with interp_psf = true and
// equalize, called by PSFreconstruction, simulate_reco.py:38-54).  The backward kernels are not
// on this path (nothing is differentiated).
//
// Geometry (identical for both): a slice pixel (ix, iy) of slice `in` sits at
//   c = R * ((ix-(w-1)/2)*res + tx, (iy-(h-1)/2)*res + ty, tz) + (vol centre)
// and every PSF tap p adds R * p.  Volume axes: x = fastest (W), y (H), z (D).
//
// forward : slice = sum_p psf_p * trilinear(vol, c + R p) / sum_p psf_p * (in-bounds weight)
//           - only the non-zero taps are visited (compact list built on the host),
//           - a block covers a 16x16 tile of ONE slice, so R p is staged once per block in smem,
//           - tiles whose centre +- PSF radius misses the volume exit immediately.
// adjoint : two passes over the taps per pixel (normalisation weight, then scatter of
//           psf/weight * s to the nearest voxel with red.global.add) + equalisation vol /= weight.
//           Summation order of the atomics is not deterministic (same as the reference).
#include <stdlib.h>

#include "common.cuh"

namespace fsg {

constexpr int ACQ_TILE = 16;
constexpr int ACQ_MAX_TAPS = 4096;

struct SliceGeom {
  float r11, r12, r13, r21, r22, r23, r31, r32, r33;
  float xc, yc, zc;
};

// centre of pixel (ix, iy): double intermediate like the reference's "(w - 1) / 2." expressions
__device__ __forceinline__ SliceGeom slice_geom(const float* __restrict__ t, int ix, int iy, int h, int w, int D, int H, int W, float res) {
  SliceGeom g;
  g.r11 = t[0]; g.r12 = t[1]; g.r13 = t[2];
  g.r21 = t[4]; g.r22 = t[5]; g.r23 = t[6];
  g.r31 = t[8]; g.r32 = t[9]; g.r33 = t[10];
  const float _x = (float)(((double)ix - (w - 1) / 2.) * (double)res + (double)t[3]);
  const float _y = (float)(((double)iy - (h - 1) / 2.) * (double)res + (double)t[7]);
  const float _z = t[11];
  float xc = g.r11 * _x + g.r12 * _y + g.r13 * _z;
  float yc = g.r21 * _x + g.r22 * _y + g.r23 * _z;
  float zc = g.r31 * _x + g.r32 * _y + g.r33 * _z;
  g.xc = (float)((double)xc + (W - 1) / 2.);
  g.yc = (float)((double)yc + (H - 1) / 2.);
  g.zc = (float)((double)zc + (D - 1) / 2.);
  return g;
}

// The eight corners of a trilinear sample.  PAIRS: `vol` is the x-pair volume of fsg_volume_xpairs
// ((v[i], v[i+1]) per voxel), so the two x corners of a row arrive in one 8-byte load — half the gather
// instructions of the acquisition kernels, which are bound by L1 wavefronts, not by bytes.
template <int PAIRS>
__device__ __forceinline__ void load_corners(const float* __restrict__ vol, int idx, int Sy, int Sz, float& c000, float& c100, float& c010, float& c110, float& c001, float& c101,
                                             float& c011, float& c111) {
  if (PAIRS == 4) {  // xy-quad volume: (v[i], v[i+1], v[i+Sy], v[i+Sy+1]) per voxel, two 16-byte loads
    const float4* v4 = reinterpret_cast<const float4*>(vol) + idx;
    const float4 a = __ldg(v4), b = __ldg(v4 + Sz);
    c000 = a.x; c100 = a.y; c010 = a.z; c110 = a.w; c001 = b.x; c101 = b.y; c011 = b.z; c111 = b.w;
  } else if (PAIRS == 2) {
    const float2* v2 = reinterpret_cast<const float2*>(vol) + idx;
    const float2 a = __ldg(v2), b = __ldg(v2 + Sy), c = __ldg(v2 + Sz), d = __ldg(v2 + Sy + Sz);
    c000 = a.x; c100 = a.y; c010 = b.x; c110 = b.y; c001 = c.x; c101 = c.y; c011 = d.x; c111 = d.y;
  } else {
    const float* v = vol + idx;
    c000 = __ldg(v); c100 = __ldg(v + 1); c010 = __ldg(v + Sy); c110 = __ldg(v + 1 + Sy);
    c001 = __ldg(v + Sz); c101 = __ldg(v + 1 + Sz); c011 = __ldg(v + Sy + Sz); c111 = __ldg(v + Sy + Sz + 1);
  }
}

// Lean accumulation of one tap (FSG_FWD_LEAN=1, opt-in): the trilinear sample as seven nested lerps with
// contracted multiply-adds instead of eight four-factor products, and — the eight weights of a tap sum
// to its PSF value — the normalisation weight accumulates that value directly.  Agrees with the product
// form to float rounding (the reference's extension is itself compiled with FMA contraction).  With the
// packed volumes the acquisition is issue-bound (ncu: 85 % issue utilisation), so instructions count.
__device__ __forceinline__ void lean_accumulate(float wx, float wy, float wz, float qw, float c000, float c100, float c010, float c110, float c001, float c101, float c011, float c111,
                                                float& val, float& weight) {
  const float a00 = __fmaf_rn(wx, c100 - c000, c000), a01 = __fmaf_rn(wx, c110 - c010, c010);
  const float a10 = __fmaf_rn(wx, c101 - c001, c001), a11 = __fmaf_rn(wx, c111 - c011, c011);
  const float b0 = __fmaf_rn(wy, a01 - a00, a00), b1 = __fmaf_rn(wy, a11 - a10, a10);
  val = __fmaf_rn(qw, __fmaf_rn(wz, b1 - b0, b0), val);
  weight += qw;
}

__global__ void __launch_bounds__(256) xyquads_kernel(const float* __restrict__ vol, float4* __restrict__ quads, int64_t n, int Sy) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    quads[i] = make_float4(vol[i], i + 1 < n ? vol[i + 1] : 0.f, i + Sy < n ? vol[i + Sy] : 0.f, i + Sy + 1 < n ? vol[i + Sy + 1] : 0.f);
}

__global__ void __launch_bounds__(256) xpairs_kernel(const float* __restrict__ vol, float2* __restrict__ pairs, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) pairs[i] = make_float2(vol[i], i + 1 < n ? vol[i + 1] : 0.f);
}

// ---------------------------------------------------------------------------------- forward
template <int PAIRS, bool LEAN>
__global__ void __launch_bounds__(ACQ_TILE* ACQ_TILE) slice_fwd_kernel(const float* __restrict__ transforms, const float* __restrict__ vol, const float4* __restrict__ taps, int ntaps,
                                                                        float radius, float* __restrict__ slices, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];  // R * tap offset (x, y, z) and weight
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;
  for (int p = threadIdx.y * ACQ_TILE + threadIdx.x; p < ntaps; p += ACQ_TILE * ACQ_TILE) {
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
  __syncthreads();
  const int ix = blockIdx.x * ACQ_TILE + threadIdx.x, iy = blockIdx.y * ACQ_TILE + threadIdx.y;
  if (ix >= w || iy >= h) return;
  const SliceGeom g = slice_geom(t, ix, iy, h, w, D, H, W, res);
  // no tap of this pixel can land inside [0, W-1) x [0, H-1) x [0, D-1): the slice keeps its zero
  if (g.xc + radius < 0.f || g.yc + radius < 0.f || g.zc + radius < 0.f || g.xc - radius >= (float)(W - 1) || g.yc - radius >= (float)(H - 1) || g.zc - radius >= (float)(D - 1)) return;
  const int Sy = W, Sz = H * W;
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  float val = 0.f, weight = 0.f;
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const float wx = x - fx, wy = y - fy, wz = z - fz;
    float c000, c100, c010, c110, c001, c101, c011, c111;
    load_corners<PAIRS>(vol, (int)fz * Sz + (int)fy * Sy + (int)fx, Sy, Sz, c000, c100, c010, c110, c001, c101, c011, c111);
    if (LEAN) {
      lean_accumulate(wx, wy, wz, q.w, c000, c100, c010, c110, c001, c101, c011, c111, val, weight);
      continue;
    }
    // the eight weights share their (y, z, PSF) factors: 14 multiplies instead of 24 (the kernel is issue-bound
    // once the gathers are packed); the reference multiplies left to right, this is its value to float rounding
    const float ux = 1.f - wx, uy = 1.f - wy, uz = 1.f - wz;
    const float uzq = uz * q.w, wzq = wz * q.w;
    const float t00 = uy * uzq, t10 = wy * uzq, t01 = uy * wzq, t11 = wy * wzq;
    const float p000 = ux * t00, p100 = wx * t00, p010 = ux * t10, p001 = ux * t01;
    const float p110 = wx * t10, p101 = wx * t01, p011 = ux * t11, p111 = wx * t11;
    // contracted multiply-adds like the reference's own build; the eight weights of a tap sum to its PSF value
    val = __fmaf_rn(p000, c000, val);
    val = __fmaf_rn(p100, c100, val);
    val = __fmaf_rn(p010, c010, val);
    val = __fmaf_rn(p001, c001, val);
    val = __fmaf_rn(p110, c110, val);
    val = __fmaf_rn(p101, c101, val);
    val = __fmaf_rn(p011, c011, val);
    val = __fmaf_rn(p111, c111, val);
    weight += q.w;
  }
  if (weight > 0.f) slices[((size_t)in * h + iy) * w + ix] = __fdiv_rn(val, weight);  // one per pixel: keep the IEEE divide
}

// Warp-per-pixel acquisition for large PSFs: the lanes split the taps of one pixel (consecutive taps
// sample neighbouring voxels: coalesced gathers), partial sums are reduced with shuffles.  Summation
// order differs from the sequential tap loop (float tolerance).
template <int PAIRS, bool LEAN>
__global__ void __launch_bounds__(ACQ_TILE* ACQ_TILE) slice_fwd_warp_kernel(const float* __restrict__ transforms, const float* __restrict__ vol, const float4* __restrict__ taps, int ntaps,
                                                                             float radius, float* __restrict__ slices, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;
  const int tid = threadIdx.y * ACQ_TILE + threadIdx.x;
  for (int p = tid; p < ntaps; p += ACQ_TILE * ACQ_TILE) {
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
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int Sy = W, Sz = H * W;
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  const int my_ix = blockIdx.x * ACQ_TILE + (lane & 15), my_iy = blockIdx.y * ACQ_TILE + 2 * warp + (lane >> 4);
  const bool inside = my_ix < w && my_iy < h;
  const SliceGeom mine = slice_geom(t, inside ? my_ix : 0, inside ? my_iy : 0, h, w, D, H, W, res);
  const bool keep = inside && !(mine.xc + radius < 0.f || mine.yc + radius < 0.f || mine.zc + radius < 0.f || mine.xc - radius >= mx || mine.yc - radius >= my || mine.zc - radius >= mz);
  unsigned todo = __ballot_sync(0xffffffffu, keep);
  float my_out = 0.f;
  bool my_set = false;
  while (todo) {
    const int q = __ffs(todo) - 1;
    todo &= todo - 1;
    const float xc = __shfl_sync(0xffffffffu, mine.xc, q), yc = __shfl_sync(0xffffffffu, mine.yc, q), zc = __shfl_sync(0xffffffffu, mine.zc, q);
    float val = 0.f, weight = 0.f;
    for (int p = lane; p < ntaps; p += 32) {
      const float4 o = s_tap[p];
      const float x = xc + o.x, y = yc + o.y, z = zc + o.z;
      if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
      const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
      const float wx = x - fx, wy = y - fy, wz = z - fz;
      float c000, c100, c010, c110, c001, c101, c011, c111;
      load_corners<PAIRS>(vol, (int)fz * Sz + (int)fy * Sy + (int)fx, Sy, Sz, c000, c100, c010, c110, c001, c101, c011, c111);
      if (LEAN) {
        lean_accumulate(wx, wy, wz, o.w, c000, c100, c010, c110, c001, c101, c011, c111, val, weight);
        continue;
      }
      const float ux = 1.f - wx, uy = 1.f - wy, uz = 1.f - wz;
      const float uzq = uz * o.w, wzq = wz * o.w;
      const float t00 = uy * uzq, t10 = wy * uzq, t01 = uy * wzq, t11 = wy * wzq;
      const float p000 = ux * t00, p100 = wx * t00, p010 = ux * t10, p001 = ux * t01;
      const float p110 = wx * t10, p101 = wx * t01, p011 = ux * t11, p111 = wx * t11;
      // contracted multiply-adds like the reference's own build; the eight weights of a tap sum to its PSF value
      val = __fmaf_rn(p000, c000, val);
      val = __fmaf_rn(p100, c100, val);
      val = __fmaf_rn(p010, c010, val);
      val = __fmaf_rn(p001, c001, val);
      val = __fmaf_rn(p110, c110, val);
      val = __fmaf_rn(p101, c101, val);
      val = __fmaf_rn(p011, c011, val);
      val = __fmaf_rn(p111, c111, val);
      weight += o.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      val += __shfl_xor_sync(0xffffffffu, val, o);
      weight += __shfl_xor_sync(0xffffffffu, weight, o);
    }
    if (lane == q && weight > 0.f) {
      my_out = __fdiv_rn(val, weight);
      my_set = true;
    }
  }
  if (my_set) slices[((size_t)in * h + my_iy) * w + my_ix] = my_out;  // one coalesced store per warp
}

// ---------------------------------------------------------------------------------- adjoint
// interpolated PSF value at the voxel nearest to (x, y, z), or -1 when it falls off the PSF grid
__device__ __forceinline__ float psf_at_voxel(const SliceGeom& g, const float* __restrict__ psf, int dp, int hp, int wp, float xr, float yr, float zr) {
  const float dx = xr - g.xc, dy = yr - g.yc, dz = zr - g.zc;
  // the reference adds the double constant (w_p-1)/2.; it is an integer or a half, exactly
  // representable in float, so the float add rounds to the same value without FP64 conversions
  const float xp = (g.r11 * dx + g.r21 * dy + g.r31 * dz) + 0.5f * (float)(wp - 1);
  const float yp = (g.r12 * dx + g.r22 * dy + g.r32 * dz) + 0.5f * (float)(hp - 1);
  const float zp = (g.r13 * dx + g.r23 * dy + g.r33 * dz) + 0.5f * (float)(dp - 1);
  if (xp < 0.f || yp < 0.f || zp < 0.f || xp >= (float)(wp - 1) || yp >= (float)(hp - 1) || zp >= (float)(dp - 1)) return -1.f;
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
  return v;
}

// Lean form for the warp kernel: the same quantity with contracted multiply-adds (the reference's
// extension is compiled with FMA contraction as well), the PSF-grid centre folded into the FMA chain,
// the trilinear blend as seven nested lerps instead of eight three-factor products, PSF in shared
// memory.  Agrees with psf_at_voxel to float rounding; `ok` false when the voxel falls off the PSF grid.
struct PsfGrid {
  const float* q;  // shared memory
  int wp, hpwp;
  float cx, cy, cz, mx, my, mz;  // (n - 1) / 2 and n - 1 per axis
};
__device__ __forceinline__ float psf_lerp(const SliceGeom& g, const PsfGrid& P, float xr, float yr, float zr, bool& ok) {
  const float dx = xr - g.xc, dy = yr - g.yc, dz = zr - g.zc;
  const float xp = __fmaf_rn(g.r11, dx, __fmaf_rn(g.r21, dy, __fmaf_rn(g.r31, dz, P.cx)));
  const float yp = __fmaf_rn(g.r12, dx, __fmaf_rn(g.r22, dy, __fmaf_rn(g.r32, dz, P.cy)));
  const float zp = __fmaf_rn(g.r13, dx, __fmaf_rn(g.r23, dy, __fmaf_rn(g.r33, dz, P.cz)));
  ok = !(xp < 0.f || yp < 0.f || zp < 0.f || xp >= P.mx || yp >= P.my || zp >= P.mz);
  if (!ok) return 0.f;
  const float fx = floorf(xp), fy = floorf(yp), fz = floorf(zp);
  const float wx = xp - fx, wy = yp - fy, wz = zp - fz;
  const float* q = P.q + ((int)fz * P.hpwp + (int)fy * P.wp + (int)fx);
  const float* q1 = q + P.hpwp;
  const float a00 = __fmaf_rn(wx, q[1] - q[0], q[0]), a01 = __fmaf_rn(wx, q[P.wp + 1] - q[P.wp], q[P.wp]);
  const float a10 = __fmaf_rn(wx, q1[1] - q1[0], q1[0]), a11 = __fmaf_rn(wx, q1[P.wp + 1] - q1[P.wp], q1[P.wp]);
  const float b0 = __fmaf_rn(wy, a01 - a00, a00), b1 = __fmaf_rn(wy, a11 - a10, a10);
  return fmaxf(__fmaf_rn(wz, b1 - b0, b0), 0.f);
}
// round half away from zero for x >= 0 (C round(), as the reference uses), exact
__device__ __forceinline__ float round_nonneg(float x) {
  const float f = floorf(x);
  return (x - f >= 0.5f) ? f + 1.0f : f;
}

__global__ void __launch_bounds__(ACQ_TILE* ACQ_TILE) slice_adj_kernel(const float* __restrict__ transforms, const float* __restrict__ psf, int dp, int hp, int wp,
                                                                        const float4* __restrict__ taps, int ntaps, float radius, const float* __restrict__ slices,
                                                                        const int* __restrict__ slice_idx, float2* __restrict__ acc, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;  // transforms are already gathered: [n][3][4]
  for (int p = threadIdx.y * ACQ_TILE + threadIdx.x; p < ntaps; p += ACQ_TILE * ACQ_TILE) {
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
  __syncthreads();
  const int ix = blockIdx.x * ACQ_TILE + threadIdx.x, iy = blockIdx.y * ACQ_TILE + threadIdx.y;
  if (ix >= w || iy >= h) return;
  const SliceGeom g = slice_geom(t, ix, iy, h, w, D, H, W, res);
  if (g.xc + radius < 0.f || g.yc + radius < 0.f || g.zc + radius < 0.f || g.xc - radius >= (float)(W - 1) || g.yc - radius >= (float)(H - 1) || g.zc - radius >= (float)(D - 1)) return;
  const float s = slices[((size_t)(slice_idx ? slice_idx[in] : in) * h + iy) * w + ix];
  const int Sy = W, Sz = H * W;
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  float weight = 0.f;
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    const float pv = psf_at_voxel(g, psf, dp, hp, wp, roundf(x), roundf(y), roundf(z));
    if (pv >= 0.f) weight += pv;
  }
  if (weight < 0.5f) return;  // border
  for (int p = 0; p < ntaps; ++p) {
    const float4 q = s_tap[p];
    const float x = g.xc + q.x, y = g.yc + q.y, z = g.zc + q.z;
    if (x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz) continue;
    const float xr = roundf(x), yr = roundf(y), zr = roundf(z);
    float pv = psf_at_voxel(g, psf, dp, hp, wp, xr, yr, zr);
    if (pv < 0.f) continue;
    pv = __fdiv_rn(pv, weight);
    const int iv = (int)zr * Sz + (int)yr * Sy + (int)xr;
    // value and weight of a voxel are interleaved: ONE 64-bit reduction per tap instead of two 32-bit
    // ones (the kernel is bound by the L2 reduction rate, ~2.6 G taps per reconstruction)
    atomicAdd(acc + iv, make_float2(pv * s, pv));
  }
}

// Warp-per-pixel reconstruction: the 32 lanes of a warp split the PSF taps of ONE slice pixel, keep
// their interpolated PSF values and target voxels in registers, reduce the normalisation weight with
// shuffles and scatter straight away — every tap is evaluated once (the thread-per-pixel kernel above,
// like the reference, evaluates every tap twice), lanes of a warp never diverge on the pixel cull and
// consecutive taps land on neighbouring voxels.  A block still owns a 16x16 pixel tile of one slice
// (rotated tap offsets staged once); each warp walks 32 of its pixels.  TPL = taps per lane.
template <int TPL, bool LEAN>
__global__ void __launch_bounds__(ACQ_TILE* ACQ_TILE) slice_adj_warp_kernel(const float* __restrict__ transforms, const float* __restrict__ psf, int dp, int hp, int wp,
                                                                             const float4* __restrict__ taps, int ntaps, float radius, const float* __restrict__ slices,
                                                                             const int* __restrict__ slice_idx, float2* __restrict__ acc, int h, int w, int D, int H, int W, float res) {
  extern __shared__ float4 s_tap[];
  const int in = blockIdx.z;
  const float* t = transforms + in * 12;
  const int tid = threadIdx.y * ACQ_TILE + threadIdx.x;
  for (int p = tid; p < ntaps; p += ACQ_TILE * ACQ_TILE) {
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
  float* s_psf = reinterpret_cast<float*>(s_tap + ntaps);
  if (LEAN)
    for (int p = tid; p < dp * hp * wp; p += ACQ_TILE * ACQ_TILE) s_psf[p] = psf[p];
  __syncthreads();
  PsfGrid P;
  P.q = s_psf;
  P.wp = wp;
  P.hpwp = hp * wp;
  P.cx = 0.5f * (float)(wp - 1);
  P.cy = 0.5f * (float)(hp - 1);
  P.cz = 0.5f * (float)(dp - 1);
  P.mx = (float)(wp - 1);
  P.my = (float)(hp - 1);
  P.mz = (float)(dp - 1);
  const int lane = tid & 31, warp = tid >> 5;
  const int Sy = W, Sz = H * W;
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  const float* __restrict__ srow = slices + (size_t)(slice_idx ? slice_idx[in] : in) * h * w;
  // warp `warp` owns tile rows 2*warp and 2*warp+1 (32 pixels).  Lane q first evaluates the geometry and
  // the cull of pixel q; the warp then walks the surviving pixels, broadcasting their geometry.
  const int my_ix = blockIdx.x * ACQ_TILE + (lane & 15), my_iy = blockIdx.y * ACQ_TILE + 2 * warp + (lane >> 4);
  const bool inside = my_ix < w && my_iy < h;
  const SliceGeom mine = slice_geom(t, inside ? my_ix : 0, inside ? my_iy : 0, h, w, D, H, W, res);
  const bool keep = inside && !(mine.xc + radius < 0.f || mine.yc + radius < 0.f || mine.zc + radius < 0.f || mine.xc - radius >= mx || mine.yc - radius >= my || mine.zc - radius >= mz);
  const float my_s = keep ? srow[my_iy * w + my_ix] : 0.f;
  unsigned todo = __ballot_sync(0xffffffffu, keep);
  while (todo) {
    const int q = __ffs(todo) - 1;
    todo &= todo - 1;
    SliceGeom g = mine;  // the rotation is the slice's: identical in every lane
    g.xc = __shfl_sync(0xffffffffu, mine.xc, q);
    g.yc = __shfl_sync(0xffffffffu, mine.yc, q);
    g.zc = __shfl_sync(0xffffffffu, mine.zc, q);
    const float s = __shfl_sync(0xffffffffu, my_s, q);
    float pv[TPL];
    int iv[TPL];
    float weight = 0.f;
#pragma unroll
    for (int u = 0; u < TPL; ++u) {
      const int p = lane + 32 * u;
      pv[u] = -1.f;
      iv[u] = 0;
      if (p < ntaps) {
        const float4 o = s_tap[p];
        const float x = g.xc + o.x, y = g.yc + o.y, z = g.zc + o.z;
        if (!(x < 0.f || y < 0.f || z < 0.f || x >= mx || y >= my || z >= mz)) {
          if (LEAN) {
            const float xr = round_nonneg(x), yr = round_nonneg(y), zr = round_nonneg(z);
            bool ok;
            const float v = psf_lerp(g, P, xr, yr, zr, ok);
            iv[u] = (int)zr * Sz + (int)yr * Sy + (int)xr;
            if (ok) {
              pv[u] = v;
              weight += v;
            }
          } else {
            const float xr = roundf(x), yr = roundf(y), zr = roundf(z);
            pv[u] = psf_at_voxel(g, psf, dp, hp, wp, xr, yr, zr);
            iv[u] = (int)zr * Sz + (int)yr * Sy + (int)xr;
            if (pv[u] >= 0.f) weight += pv[u];
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) weight += __shfl_xor_sync(0xffffffffu, weight, o);
    if (weight < 0.5f) continue;  // border
    const float inv = __frcp_rn(weight);
#pragma unroll
    for (int u = 0; u < TPL; ++u) {
      if (pv[u] >= 0.f) {
        const float v = LEAN ? pv[u] * inv : __fdiv_rn(pv[u], weight);
        atomicAdd(acc + iv[u], make_float2(v * s, v));
      }
    }
  }
}

// de-interleave the accumulator; equalize: vol /= weight where the weight is positive (kernel :672-693)
__global__ void __launch_bounds__(256) equalize_kernel(const float2* __restrict__ acc, float* __restrict__ vol, float* __restrict__ wgt, unsigned n, int equalize) {
  for (unsigned v = blockIdx.x * 256 + threadIdx.x; v < n; v += gridDim.x * 256) {
    const float2 a = acc[v];
    vol[v] = (equalize && a.y > 0.f) ? __fdiv_rn(a.x, a.y) : a.x;
    if (wgt) wgt[v] = a.y;
  }
}

// 3^3 mean with zero padding (PSFReconstructor.smooth_volume, simulate_reco.py:584-595) followed by
// the merge with the clean volume: out = wgt * rec + (1 - wgt) * gt (merge_volumes, :692-709),
// wgt = clamp((perlin + increase - pmin) / (pmax - pmin), 0, 1) or a ready MoG weight.
__global__ void __launch_bounds__(256) recon_merge_kernel(const float* __restrict__ rec, const float* __restrict__ gt, const float* __restrict__ wraw, const float* __restrict__ mm,
                                                          float increase, int smooth, int D, int H, int W, float* __restrict__ out) {
  const unsigned n = (unsigned)D * H * W;
  const float pmin = mm ? mm[0] : 0.f, prange = mm ? mm[1] - mm[0] : 1.f;
  for (unsigned v = blockIdx.x * 256 + threadIdx.x; v < n; v += gridDim.x * 256) {
    float r = rec[v];
    if (smooth) {
      const int x = (int)(v % (unsigned)W), y = (int)((v / (unsigned)W) % (unsigned)H), z = (int)(v / ((unsigned)W * H));
      float acc = 0.f;
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx, yy = y + dy, zz = z + dz;
            if (xx < 0 || yy < 0 || zz < 0 || xx >= W || yy >= H || zz >= D) continue;
            acc += rec[(zz * H + yy) * W + xx] * (1.0f / 27.0f);
          }
      r = acc;
    }
    float wv = 0.f;
    if (wraw) wv = fminf(fmaxf(__fdividef(wraw[v] + increase - pmin, prange), 0.f), 1.f);
    out[v] = wraw ? wv * r + (1.f - wv) * gt[v] : r;
  }
}


// ---------------------------------------------------------------------------------- slice stack ops
// per-slice sums (Scanner.scan, simulate_reco.py:408: nnz = slices_no_psf.sum((1,2,3)))
__global__ void __launch_bounds__(256) slice_sums_kernel(const float* __restrict__ slices, int hw, float* __restrict__ sums) {
  const float* s = slices + (size_t)blockIdx.x * hw;
  float acc = 0.f;
  for (int i = threadIdx.x; i < hw; i += 256) acc += s[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k];
    sums[blockIdx.x] = t;
  }
}

// Scanner.random_gamma (:225-236): s = 300 (s/300)^gamma, then s / max(s).  Pass 1 writes the power
// and reduces the max (values are >= 0: the int view of a non-negative float is order-preserving).
__global__ void __launch_bounds__(256) slice_gamma_kernel(float* __restrict__ s, unsigned n, float gamma, int* __restrict__ max_bits) {
  float m = 0.f;
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float v = 300.0f * powf(s[i] / 300.0f, gamma);
    s[i] = v;
    m = fmaxf(m, v);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_int(m));
}
__global__ void __launch_bounds__(256) slice_div_kernel(float* __restrict__ s, unsigned n, const int* __restrict__ max_bits) {
  const float m = __int_as_float(*max_bits);
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) s[i] = __fdiv_rn(s[i], m);
}

// Scanner.add_noise (:248-257): Rician noise where the slice exceeds the threshold.  The reference
// draws randn only for the masked pixels; here every pixel owns two counter-based draws
// (block = pixel / 2) or reads the injected full-size arrays.
__global__ void __launch_bounds__(256) slice_rician_kernel(float* __restrict__ s, unsigned n, float threshold, float sigma, const float* __restrict__ n1, const float* __restrict__ n2,
                                                           fsg_rng rng) {
  for (unsigned i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float v = s[i];
    if (!(v > threshold)) continue;
    float a, b;
    if (n1) {
      a = n1[i];
      b = n2[i];
    } else {
      const float4 g = philox_normal4(rng, i >> 1);
      a = (i & 1) ? g.z : g.x;
      b = (i & 1) ? g.w : g.y;
    }
    const float p = v + a * sigma, q = b * sigma;
    s[i] = sqrtf(p * p + q * q);
  }
}

// Scanner.signal_void (:273-297): slice idx[k] *= 1 - A exp(sx x'^2 + sy y'^2) in a rotated frame.
// params[k] = (yc, xc, theta, a, A, sx) as drawn by the reference.
__global__ void __launch_bounds__(256) slice_void_kernel(float* __restrict__ slices, int h, int w, const int* __restrict__ idx, const float* __restrict__ params) {
  const int k = blockIdx.y;
  const float* p = params + 6 * k;
  float* s = slices + (size_t)idx[k] * h * w;
  const float yc = p[0], xc = p[1], a = p[3], A = p[4], sx0 = p[5];
  float sn, cs;
  sincosf(p[2], &sn, &cs);
  const float sy0 = a * a / sx0;
  const float gx = -0.5f / (sx0 * sx0), gy = -0.5f / (sy0 * sy0);
  const float y0 = -(float)(h - 1) / 2, x0 = -(float)(w - 1) / 2;
  const float ystep = h > 1 ? (float)(h - 1) / (float)(h - 1) : 0.f, xstep = w > 1 ? 1.f : 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < h * w; i += gridDim.x * 256) {
    const int iy = i / w, ix = i - iy * w;
    const float y = (y0 + ystep * iy) - yc, x = (x0 + xstep * ix) - xc;
    const float xr = cs * x - sn * y, yr = sn * x + cs * y;
    s[i] = s[i] * (1.0f - A * expf(gx * xr * xr + gy * yr * yr));
  }
}

}  // namespace fsg

using namespace fsg;

static int check_acq(const char* who, int ntaps, int n, int h, int w, int D, int H, int W) {
  FSG_REQUIRE(n >= 1 && n <= 65535 && h >= 1 && w >= 1, "%s: bad slice stack shape", who);
  FSG_REQUIRE(D >= 2 && H >= 2 && W >= 2 && (int64_t)D * H * W < ((int64_t)1 << 31), "%s: bad volume shape", who);
  FSG_REQUIRE(ntaps >= 1 && ntaps <= ACQ_MAX_TAPS, "%s: ntaps=%d outside [1,%d]", who, ntaps, ACQ_MAX_TAPS);
  return 0;
}

static int acq_forward(const char* who, int pairs, const float* transforms, const float* vol, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D,
                       int H, int W, float res_slice, void* stream) {
  if (int rc = check_acq(who, ntaps, n, h, w, D, H, W)) return rc;
  FSG_REQUIRE(transforms && vol && taps && slices, "%s: NULL pointer", who);
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(taps) & 15) == 0, "%s: taps must be 16-byte aligned", who);
  FSG_REQUIRE(pairs == 1 || (reinterpret_cast<uintptr_t>(vol) & (4 * pairs - 1)) == 0, "%s: the packed volume must be %d-byte aligned", who, 4 * pairs);
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(slices, 0, sizeof(float) * (size_t)n * h * w, s);
  dim3 grid((w + ACQ_TILE - 1) / ACQ_TILE, (h + ACQ_TILE - 1) / ACQ_TILE, n);
  const dim3 block(ACQ_TILE, ACQ_TILE);
  const size_t smem = sizeof(float4) * ntaps;
  const float4* taps4 = reinterpret_cast<const float4*>(taps);
  // large PSFs: lanes over taps (coalesced gathers); small ones (the 1-tap mask acquisition): thread per pixel
  const int warp_min_taps = config().fwd_warp_min_taps;
#define FSG_FWD_LAUNCH(KERNEL)                                                                                  \
  do {                                                                                                          \
    if (smem > 48 * 1024) cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    KERNEL<<<grid, block, smem, s>>>(transforms, vol, taps4, ntaps, radius, slices, h, w, D, H, W, res_slice);  \
  } while (0)
  const bool lean = config().fwd_lean;
#define FSG_FWD_PICK(KERNEL)                              \
  do {                                                    \
    if (lean) {                                           \
      if (pairs == 4) FSG_FWD_LAUNCH((KERNEL<4, true>));  \
      else if (pairs == 2) FSG_FWD_LAUNCH((KERNEL<2, true>)); \
      else FSG_FWD_LAUNCH((KERNEL<1, true>));             \
    } else {                                              \
      if (pairs == 4) FSG_FWD_LAUNCH((KERNEL<4, false>)); \
      else if (pairs == 2) FSG_FWD_LAUNCH((KERNEL<2, false>)); \
      else FSG_FWD_LAUNCH((KERNEL<1, false>));            \
    }                                                     \
  } while (0)
  if (ntaps >= warp_min_taps)
    FSG_FWD_PICK(slice_fwd_warp_kernel);
  else
    FSG_FWD_PICK(slice_fwd_kernel);
#undef FSG_FWD_PICK
#undef FSG_FWD_LAUNCH
  return check_launch(who);
}

extern "C" int fsg_slice_acq_forward(const float* transforms, const float* vol, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D, int H, int W,
                                     float res_slice, void* stream) {
  return acq_forward("fsg_slice_acq_forward", 1, transforms, vol, taps, ntaps, radius, slices, n, h, w, D, H, W, res_slice, stream);
}

extern "C" int fsg_slice_acq_forward_xpairs(const float* transforms, const float* vol_pairs, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D,
                                            int H, int W, float res_slice, void* stream) {
  return acq_forward("fsg_slice_acq_forward_xpairs", 2, transforms, vol_pairs, taps, ntaps, radius, slices, n, h, w, D, H, W, res_slice, stream);
}

extern "C" int fsg_slice_acq_forward_xyquads(const float* transforms, const float* vol_quads, const float* taps, int ntaps, float radius, float* slices, int n, int h, int w, int D,
                                             int H, int W, float res_slice, void* stream) {
  return acq_forward("fsg_slice_acq_forward_xyquads", 4, transforms, vol_quads, taps, ntaps, radius, slices, n, h, w, D, H, W, res_slice, stream);
}

extern "C" int fsg_volume_xyquads(const float* vol, float* quads, int64_t nvox, int row_len, void* stream) {
  FSG_REQUIRE(vol && quads && nvox >= 1 && row_len >= 1, "fsg_volume_xyquads: bad arguments");
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(quads) & 15) == 0, "fsg_volume_xyquads: quads must be 16-byte aligned");
  const int64_t want = (nvox + 255) / 256;
  xyquads_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, as_stream(stream)>>>(vol, reinterpret_cast<float4*>(quads), nvox, row_len);
  return check_launch("fsg_volume_xyquads");
}

extern "C" int fsg_volume_xpairs(const float* vol, float* pairs, int64_t nvox, void* stream) {
  FSG_REQUIRE(vol && pairs && nvox >= 1, "fsg_volume_xpairs: bad arguments");
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(pairs) & 7) == 0, "fsg_volume_xpairs: pairs must be 8-byte aligned");
  const int64_t want = (nvox + 255) / 256;
  xpairs_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, as_stream(stream)>>>(vol, reinterpret_cast<float2*>(pairs), nvox);
  return check_launch("fsg_volume_xpairs");
}

extern "C" int fsg_slice_acq_adjoint(const float* transforms, const float* psf, int dp, int hp, int wp, const float* taps, int ntaps, float radius, const float* slices,
                                     const int32_t* slice_idx, float* vol, float* vol_weight, float* workspace, int n, int h, int w, int D, int H, int W, float res_slice,
                                     int equalize, void* stream) {
  if (int rc = check_acq("fsg_slice_acq_adjoint", ntaps, n, h, w, D, H, W)) return rc;
  FSG_REQUIRE(transforms && psf && taps && slices && vol && workspace, "fsg_slice_acq_adjoint: NULL pointer");
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "fsg_slice_acq_adjoint: workspace must be 8-byte aligned");
  FSG_REQUIRE(dp >= 2 && hp >= 2 && wp >= 2, "fsg_slice_acq_adjoint: PSF must be at least 2 voxels wide per axis");
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(taps) & 15) == 0, "fsg_slice_acq_adjoint: taps must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  const size_t nv = (size_t)D * H * W;
  float2* acc = reinterpret_cast<float2*>(workspace);
  cudaMemsetAsync(acc, 0, sizeof(float2) * nv, s);
  dim3 grid((w + ACQ_TILE - 1) / ACQ_TILE, (h + ACQ_TILE - 1) / ACQ_TILE, n);
  const size_t smem = sizeof(float4) * ntaps;
  if (smem > 48 * 1024) cudaFuncSetAttribute(slice_adj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const float4* taps4 = reinterpret_cast<const float4*>(taps);
  const bool lean = config().adj_lean;  // A/B: false = the reference's operation order (uncontracted, eight-product blend, IEEE divide)
  const size_t smem_w = smem + (lean ? sizeof(float) * (size_t)dp * hp * wp : 0);
  FSG_REQUIRE(smem_w <= 200 * 1024, "fsg_slice_acq_adjoint: %d taps + a %dx%dx%d PSF do not fit in shared memory", ntaps, dp, hp, wp);
#define FSG_ADJ_WARP(TPL)                                                                                                                                                  \
  do {                                                                                                                                                                     \
    if (lean) {                                                                                                                                                            \
      if (smem_w > 48 * 1024) cudaFuncSetAttribute(slice_adj_warp_kernel<TPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);                            \
      slice_adj_warp_kernel<TPL, true><<<grid, dim3(ACQ_TILE, ACQ_TILE), smem_w, s>>>(transforms, psf, dp, hp, wp, taps4, ntaps, radius, slices, slice_idx, acc, h, w, D, H, \
                                                                                      W, res_slice);                                                                      \
    } else {                                                                                                                                                               \
      if (smem_w > 48 * 1024) cudaFuncSetAttribute(slice_adj_warp_kernel<TPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);                           \
      slice_adj_warp_kernel<TPL, false><<<grid, dim3(ACQ_TILE, ACQ_TILE), smem_w, s>>>(transforms, psf, dp, hp, wp, taps4, ntaps, radius, slices, slice_idx, acc, h, w, D,  \
                                                                                       H, W, res_slice);                                                                  \
    }                                                                                                                                                                      \
  } while (0)
  const int tpl = (ntaps + 31) / 32;
  const bool per_thread = config().adj_thread;  // A/B: the thread-per-pixel kernel
  if (per_thread || tpl > 32)
    slice_adj_kernel<<<grid, dim3(ACQ_TILE, ACQ_TILE), smem, s>>>(transforms, psf, dp, hp, wp, taps4, ntaps, radius, slices, slice_idx, acc, h, w, D, H, W, res_slice);
  else if (tpl <= 2)
    FSG_ADJ_WARP(2);
  else if (tpl <= 4)
    FSG_ADJ_WARP(4);
  else if (tpl <= 8)
    FSG_ADJ_WARP(8);
  else if (tpl <= 12)
    FSG_ADJ_WARP(12);
  else if (tpl <= 16)
    FSG_ADJ_WARP(16);
  else if (tpl <= 24)
    FSG_ADJ_WARP(24);
  else
    FSG_ADJ_WARP(32);
#undef FSG_ADJ_WARP
  {
    const size_t want = (nv + 255) / 256;
    equalize_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, s>>>(acc, vol, vol_weight, (unsigned)nv, equalize);
  }
  return check_launch("fsg_slice_acq_adjoint");
}

static unsigned grid_for(size_t n) {
  const size_t want = (n + 255) / 256;
  return (unsigned)(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
}

extern "C" int fsg_slice_sums(const float* slices, int n, int hw, float* sums, void* stream) {
  FSG_REQUIRE(slices && sums && n >= 1 && hw >= 1, "fsg_slice_sums: bad arguments");
  slice_sums_kernel<<<n, 256, 0, as_stream(stream)>>>(slices, hw, sums);
  return check_launch("fsg_slice_sums");
}

extern "C" int fsg_slice_gamma(float* slices, int64_t count, float gamma, float* workspace, void* stream) {
  FSG_REQUIRE(slices && workspace && count >= 1 && count < ((int64_t)1 << 32), "fsg_slice_gamma: bad arguments");
  cudaStream_t s = as_stream(stream);
  cudaMemsetAsync(workspace, 0, sizeof(float), s);
  slice_gamma_kernel<<<grid_for(count), 256, 0, s>>>(slices, (unsigned)count, gamma, reinterpret_cast<int*>(workspace));
  slice_div_kernel<<<grid_for(count), 256, 0, s>>>(slices, (unsigned)count, reinterpret_cast<const int*>(workspace));
  return check_launch("fsg_slice_gamma");
}

extern "C" int fsg_slice_rician(float* slices, int64_t count, float threshold, float sigma, const float* noise1, const float* noise2, fsg_rng rng, void* stream) {
  FSG_REQUIRE(slices && count >= 1 && count < ((int64_t)1 << 32) && (!noise1 == !noise2), "fsg_slice_rician: bad arguments");
  slice_rician_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(slices, (unsigned)count, threshold, sigma, noise1, noise2, rng);
  return check_launch("fsg_slice_rician");
}

extern "C" int fsg_slice_void(float* slices, int h, int w, const int32_t* idx, const float* params, int nvoid, void* stream) {
  FSG_REQUIRE(slices && idx && params && h >= 1 && w >= 1 && nvoid >= 1 && nvoid <= 65535, "fsg_slice_void: bad arguments");
  dim3 grid(grid_for((size_t)h * w) < 64 ? grid_for((size_t)h * w) : 64, nvoid);
  slice_void_kernel<<<grid, 256, 0, as_stream(stream)>>>(slices, h, w, idx, params);
  return check_launch("fsg_slice_void");
}

extern "C" int fsg_recon_merge(const float* rec, const float* gt, const float* weight_raw, const float* minmax, float increase, int smooth, int D, int H, int W, float* out,
                               void* stream) {
  FSG_REQUIRE(rec && out && rec != out && (!weight_raw || gt), "fsg_recon_merge: bad pointers");
  FSG_REQUIRE(D >= 1 && H >= 1 && W >= 1 && (int64_t)D * H * W < ((int64_t)1 << 31), "fsg_recon_merge: bad shape");
  const size_t want = ((size_t)D * H * W + 255) / 256;
  recon_merge_kernel<<<(unsigned)(want < 148 * 16 ? want : 148 * 16), 256, 0, as_stream(stream)>>>(rec, gt, weight_raw, minmax, increase, smooth, D, H, W, out);
  return check_launch("fsg_recon_merge");
}
