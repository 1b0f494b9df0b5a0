// K5 — SR-artifact simulators: building blocks of BlurCortex, StructNoise and
// SimulatedBoundaries (reference: generator/augmentation/artifacts.py:24-133, 136-342, 428-604;
// generator/artifacts/utils.py:125-388).
//
//   fsg_mog            mixture of anisotropic Gaussians, separable exp tables, per-row culling of
//                      blobs that cannot reach the row; optional fused BlurCortex blend
//   fsg_sample_voxels  weighted sampling without replacement of voxel centres (exponential-race
//                      keys = the distribution of torch.multinomial / randperm[:k])
//   fsg_perlin         fractal Perlin noise (tileable, quintic fade) + min/max
//   fsg_struct_blend   StructNoise merge (noise normalisation, clamp, Perlin weight, seg mask)
//   fsg_morph_*        exact binary morphology on uint8 masks: box dilate/erode, 3^3 count
//                      threshold, ball dilation as a squared-distance test, L1 distance, ring
//                      sub-sampling, final boundary selection
// Integer work is bit-exact against the reference; float work is held to the 1e-4 tolerance.
#include "common.cuh"

namespace fsg {

constexpr int ART_THREADS = 256;
constexpr int MOG_MAX = 1024;  // blobs per call

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float NEG_HALF_LOG2E = -0.7213475204444817f;  // exp(-d/2) = 2^(d * this)

// ------------------------------------------------------------------------------------ MoG
// One block per (x, y) row.  Phase 1: c_k = exp(-dx^2/2) * exp(-dy^2/2) for every blob, blobs with
// c_k below 2^-40 are dropped (they add < 1e-12 each, the sum is clamped to [0, 1] and compared
// at 1e-5), survivors are compacted in blob order (deterministic sums).  Phase 2: every thread
// walks the survivors for its z positions.
__global__ void __launch_bounds__(ART_THREADS) mog_kernel(const float* __restrict__ cen, const float* __restrict__ sig, int n, int sx, int sy, int sz, float* __restrict__ out,
                                                          const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst) {
  __shared__ float s_c[MOG_MAX], s_z0[MOG_MAX], s_iz[MOG_MAX];
  __shared__ int s_cnt[ART_THREADS / 32 + 1];
  __shared__ int s_total;
  for (int row = blockIdx.x; row < sx * sy; row += gridDim.x) {
    const int i = row / sy, j = row - i * sy;
    __syncthreads();
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += ART_THREADS) {
      const int k = k0 + threadIdx.x;
      float c = 0.f, z0 = 0.f, iz = 0.f;
      bool keep = false;
      if (k < n) {
        const float dx = __fdividef((float)i - cen[3 * k + 0], sig[3 * k + 0]);
        const float dy = __fdividef((float)j - cen[3 * k + 1], sig[3 * k + 1]);
        const float e = NEG_HALF_LOG2E * (dx * dx + dy * dy);
        keep = e > -40.f;
        c = ex2a(e);
        z0 = cen[3 * k + 2];
        iz = __frcp_rn(sig[3 * k + 2]);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      if (lane == 0) s_cnt[w] = __popc(bal);
      __syncthreads();
      int off = base;
      for (int q = 0; q < w; ++q) off += s_cnt[q];
      if (keep) {
        const int p = off + __popc(bal & ((1u << lane) - 1));
        s_c[p] = c;
        s_z0[p] = z0;
        s_iz[p] = iz;
      }
      if (threadIdx.x == 0) {
        int t = base;
        for (int q = 0; q < ART_THREADS / 32; ++q) t += s_cnt[q];
        s_total = t;
      }
      __syncthreads();
      base = s_total;
    }
    const int m = base;
    for (int l = threadIdx.x; l < sz; l += ART_THREADS) {
      float acc = 0.f;
      for (int q = 0; q < m; ++q) {
        const float d = ((float)l - s_z0[q]) * s_iz[q];
        acc = __fmaf_rn(s_c[q], ex2a(NEG_HALF_LOG2E * d * d), acc);
      }
      acc = fminf(fmaxf(acc, 0.f), 1.f);
      const size_t o = (size_t)row * sz + l;
      if (out) out[o] = acc;
      if (dst) dst[o] = __fmaf_rn(b[o], acc, a[o] * (1.0f - acc));
    }
  }
}

// ------------------------------------------------------------------------------------ voxel sampling
// weight of voxel (i,j,l): label predicate times an optional small MoG (the frontal-lobe prior of
// BlurCortex.blur_proba, evaluated like mog_kernel but directly).
struct SampleSpec {
  const uint8_t* lab;   // label / mask volume
  const uint8_t* lab2;  // optional second mask: predicate becomes lab2 > 0 && lab == 0 (surface)
  int match;            // lab == match, or (match < 0) lab > 0
  int nprior;           // 0 = uniform weights
  float pc[4][3], ps[4][3];
};

__device__ __forceinline__ bool sample_pred(const SampleSpec& sp, size_t v) {
  if (sp.lab2) return sp.lab2[v] > 0 && sp.lab[v] == 0;
  return sp.match < 0 ? sp.lab[v] > 0 : sp.lab[v] == (uint8_t)sp.match;
}
__device__ __forceinline__ float sample_weight(const SampleSpec& sp, int i, int j, int l) {
  if (sp.nprior == 0) return 1.f;
  float acc = 0.f;
  for (int k = 0; k < sp.nprior; ++k) {
    const float dx = ((float)i - sp.pc[k][0]) / sp.ps[k][0], dy = ((float)j - sp.pc[k][1]) / sp.ps[k][1], dz = ((float)l - sp.pc[k][2]) / sp.ps[k][2];
    acc += ex2a(NEG_HALF_LOG2E * (dx * dx + dy * dy + dz * dz));
  }
  return fminf(acc, 1.f);
}

// pass 1: total weight (double) and candidate count
__global__ void __launch_bounds__(ART_THREADS) sample_total_kernel(const __grid_constant__ SampleSpec sp, int sx, int sy, int sz, double* total, unsigned* count) {
  double acc = 0.0;
  unsigned cnt = 0;
  const unsigned n = (unsigned)sx * sy * sz;
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    if (!sample_pred(sp, v)) continue;
    const unsigned q = v / (unsigned)sz;
    acc += (double)sample_weight(sp, (int)(q / (unsigned)sy), (int)(q % (unsigned)sy), (int)(v - q * (unsigned)sz));
    ++cnt;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(total, acc);
    atomicAdd(count, cnt);
  }
}

// pass 2: exponential-race keys key = -ln(u)/w; voxels whose key is below the threshold
// tau = (k + 6 sqrt(k) + 12) / total are appended to the candidate list (expected size ~ k + 6 sqrt k).
__global__ void __launch_bounds__(ART_THREADS) sample_keys_kernel(const __grid_constant__ SampleSpec sp, int sx, int sy, int sz, fsg_rng rng, int k, const double* total,
                                                                 float2* cand, unsigned* ncand, unsigned cap) {
  const double tot = *total;
  if (!(tot > 0.0)) return;
  const float tau = (float)(((double)k + 6.0 * sqrt((double)k) + 12.0) / tot);
  const unsigned n = (unsigned)sx * sy * sz;
  const Philox ph(rng.seed);
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    if (!sample_pred(sp, v)) continue;
    const unsigned q = v / (unsigned)sz;
    const float w = sample_weight(sp, (int)(q / (unsigned)sy), (int)(q % (unsigned)sy), (int)(v - q * (unsigned)sz));
    if (!(w > 0.f)) continue;
    const uint4 r = ph(v, rng.stage, (uint32_t)rng.sample, (uint32_t)(rng.sample >> 32));
    const float u = 2.0f - __uint_as_float(0x3f800000u | (r.x >> 9));  // (0, 1]
    const float key = -0.6931471805599453f * lg2_approx_ftz(u) / w;
    if (key < tau) {
      const unsigned p = atomicAdd(ncand, 1u);
      if (p < cap) cand[p] = make_float2(key, __uint_as_float(v));
    }
  }
}

// pass 3 (one block): sort the candidates by (key, voxel) and emit the first k as centres.
// out[c][0..2] = (axis0, axis1, axis2) voxel coordinates as floats; n_out = number written.
__global__ void __launch_bounds__(1024) sample_pick_kernel(float2* cand, const unsigned* ncand, unsigned cap, int k, int sy, int sz, int transpose, float* out, int* n_out) {
  __shared__ float s_key[2048];
  __shared__ unsigned s_vox[2048];
  const unsigned m = min(*ncand, min(cap, 2048u));
  for (unsigned t = threadIdx.x; t < 2048; t += 1024) {
    s_key[t] = t < m ? cand[t].x : __int_as_float(0x7f800000);
    s_vox[t] = t < m ? __float_as_uint(cand[t].y) : 0xffffffffu;
  }
  __syncthreads();
  for (unsigned size = 2; size <= 2048; size <<= 1) {
    for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
      for (unsigned t = threadIdx.x; t < 1024; t += 1024) {
        const unsigned lo = 2 * t - (t & (stride - 1));
        const unsigned hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const bool gt = s_key[lo] > s_key[hi] || (s_key[lo] == s_key[hi] && s_vox[lo] > s_vox[hi]);
        if (gt == up) {
          const float fk = s_key[lo];
          s_key[lo] = s_key[hi];
          s_key[hi] = fk;
          const unsigned fv = s_vox[lo];
          s_vox[lo] = s_vox[hi];
          s_vox[hi] = fv;
        }
      }
      __syncthreads();
    }
  }
  const int take = min((unsigned)k, m);
  for (int c = threadIdx.x; c < take; c += 1024) {
    const unsigned v = s_vox[c];
    const unsigned q = v / (unsigned)sz;
    // transpose: the reference hands voxel indices (i0,i1,i2) to mog_3d_tensor, which unpacks them
    // as (x0,y0,z0) with x = LAST axis, so the blob sits at axis-order position (i2, i1, i0)
    out[3 * c + (transpose ? 2 : 0)] = (float)(q / (unsigned)sy);
    out[3 * c + 1] = (float)(q % (unsigned)sy);
    out[3 * c + (transpose ? 0 : 2)] = (float)(v - q * (unsigned)sz);
  }
  if (threadIdx.x == 0) *n_out = take;
}

// ------------------------------------------------------------------------------------ Perlin
struct PerlinOct {
  const float* grad;    // [(r0+1)][(r1+1)][(r2+1)][3] unit gradients, tile wrap already applied
  const float* lin[3];  // per-axis lattice coordinate of every voxel (torch.linspace(0, res, S))
  int res[3];
  float amp;
};
struct PerlinArgs {
  PerlinOct o[8];
  int noct;
};

__device__ __forceinline__ float fade(float t) { return t * t * t * (t * (t * 6.f - 15.f) + 10.f); }

__device__ __forceinline__ float perlin_at(const PerlinOct& p, int i, int j, int l) {
  const float gx = __ldg(p.lin[0] + i), gy = __ldg(p.lin[1] + j), gz = __ldg(p.lin[2] + l);
  const float fx = floorf(gx), fy = floorf(gy), fz = floorf(gz);
  const int cx = (int)fx, cy = (int)fy, cz = (int)fz;
  const float lx = gx - fx, ly = gy - fy, lz = gz - fz;
  const int n1 = p.res[1] + 1, n2 = p.res[2] + 1;
  const int x0 = min(cx, p.res[0]), x1 = min(cx + 1, p.res[0]);
  const int y0 = min(cy, p.res[1]), y1 = min(cy + 1, p.res[1]);
  const int z0 = min(cz, p.res[2]), z1 = min(cz + 1, p.res[2]);
  auto dot = [&](int a, int b, int c, float ox, float oy, float oz) {
    const float* g = p.grad + ((a * n1 + b) * n2 + c) * 3;
    return __ldg(g) * (lx - ox) + __ldg(g + 1) * (ly - oy) + __ldg(g + 2) * (lz - oz);
  };
  const float n000 = dot(x0, y0, z0, 0, 0, 0), n100 = dot(x1, y0, z0, 1, 0, 0);
  const float n010 = dot(x0, y1, z0, 0, 1, 0), n110 = dot(x1, y1, z0, 1, 1, 0);
  const float n001 = dot(x0, y0, z1, 0, 0, 1), n101 = dot(x1, y0, z1, 1, 0, 1);
  const float n011 = dot(x0, y1, z1, 0, 1, 1), n111 = dot(x1, y1, z1, 1, 1, 1);
  const float tx = fade(lx), ty = fade(ly), tz = fade(lz);
  const float n00 = n000 * (1.f - tx) + tx * n100, n10 = n010 * (1.f - tx) + tx * n110;
  const float n01 = n001 * (1.f - tx) + tx * n101, n11 = n011 * (1.f - tx) + tx * n111;
  const float n0 = n00 * (1.f - ty) + ty * n10, n1v = n01 * (1.f - ty) + ty * n11;
  return n0 * (1.f - tz) + tz * n1v;
}

__global__ void __launch_bounds__(ART_THREADS) perlin_kernel(const __grid_constant__ PerlinArgs pa, int sx, int sy, int sz, float* __restrict__ out, float* mm) {
  const unsigned n = (unsigned)sx * sy * sz;
  const float inf = __int_as_float(0x7f800000);
  float lo = inf, hi = -inf;
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    const unsigned q = v / (unsigned)sz;
    const int i = (int)(q / (unsigned)sy), j = (int)(q % (unsigned)sy), l = (int)(v - q * (unsigned)sz);
    float acc = 0.f;
    for (int o = 0; o < pa.noct; ++o) acc += pa.o[o].amp * perlin_at(pa.o[o], i, j, l);
    out[v] = acc;
    lo = fminf(lo, acc);
    hi = fmaxf(hi, acc);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<int*>(mm), float_to_ordered(lo));
    atomicMax(reinterpret_cast<int*>(mm) + 1, float_to_ordered(hi));
  }
}

// StructNoise merge (artifacts.py:322-339).  scal = device floats:
//   [0..1] min/max of the multi-scale noise, [2..3] min/max of the image, [4..5] min/max of the
//   raw fractal noise.  weight = clamp((p + increase - pmin) / (pmax - pmin), 0, 1).
__global__ void __launch_bounds__(ART_THREADS) struct_blend_kernel(const float* __restrict__ x, const uint8_t* __restrict__ seg, const float* __restrict__ lr,
                                                                   const float* __restrict__ perlin, const float* __restrict__ scal, float noise_std, float increase,
                                                                   float* __restrict__ out, unsigned n) {
  const float maxabs = fmaxf(fabsf(scal[0]), fabsf(scal[1]));
  const float xhi = scal[3] * 2.f;
  const float pmin = scal[4], prange = scal[5] - scal[4];
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    const float xv = x[v];
    float o = xv;
    if (seg[v] > 0) {
      const float noisy = fminf(fmaxf(xv + noise_std * __fdividef(lr[v], maxabs), 0.f), xhi);
      const float w = fminf(fmaxf(__fdividef(perlin[v] + increase - pmin, prange), 0.f), 1.f);
      o = (1.f - w) * xv + w * noisy;
    }
    out[v] = o;
  }
}

// ------------------------------------------------------------------------------------ morphology
enum { MORPH_MAX = 0, MORPH_MIN = 1, MORPH_SUM = 2 };

// 1-D zero-padded window op of half-width r along one axis on uint8 volumes.
template <int OP>
__global__ void __launch_bounds__(ART_THREADS) morph_axis_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sx, int sy, int sz, int axis, int r) {
  const unsigned n = (unsigned)sx * sy * sz;
  const int stride = axis == 0 ? sy * sz : (axis == 1 ? sz : 1);
  const int len = axis == 0 ? sx : (axis == 1 ? sy : sz);
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    const int pos = axis == 2 ? (int)(v % (unsigned)sz) : (axis == 1 ? (int)((v / (unsigned)sz) % (unsigned)sy) : (int)(v / ((unsigned)sy * sz)));
    const int t0 = max(-r, -pos), t1 = min(r, len - 1 - pos);
    int acc = OP == MORPH_MIN ? 255 : 0;
    if (OP == MORPH_MIN && (t0 > -r || t1 < r)) acc = 0;  // zero padding enters the window
    for (int t = t0; t <= t1; ++t) {
      const int s = src[(int)v + t * stride];
      acc = OP == MORPH_MAX ? max(acc, s) : (OP == MORPH_MIN ? min(acc, s) : acc + s);
    }
    dst[v] = (uint8_t)acc;
  }
}

// Vector forms of the window op (bit-identical to morph_axis_kernel): SIMD byte max / min / add on
// packed voxels.  Zero padding needs no special case: out-of-range words are 0, which is the identity
// of max / add and the padding value of the zero-padded erosion.
template <int OP>
__device__ __forceinline__ uint32_t morph_op4(uint32_t a, uint32_t b) {
  return OP == MORPH_MAX ? __vmaxu4(a, b) : (OP == MORPH_MIN ? __vminu4(a, b) : __vadd4(a, b));
}

// axis 0 / 1 (stride >= one z row): 16 consecutive z voxels per thread, one 16-byte load per tap.
template <int OP>
__global__ void __launch_bounds__(ART_THREADS) morph_axis_vec_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int sx, int sy, int sz, int axis, int r) {
  const unsigned n16 = (unsigned)sx * sy * sz / 16u;
  const int stride = axis == 0 ? sy * sz : sz;
  const int len = axis == 0 ? sx : sy;
  for (unsigned g = blockIdx.x * ART_THREADS + threadIdx.x; g < n16; g += gridDim.x * ART_THREADS) {
    const unsigned v = g * 16u;
    const int pos = axis == 1 ? (int)((v / (unsigned)sz) % (unsigned)sy) : (int)(v / ((unsigned)sy * sz));
    const int t0 = max(-r, -pos), t1 = min(r, len - 1 - pos);
    const uint32_t init = (OP == MORPH_MIN && t0 == -r && t1 == r) ? 0xffffffffu : 0u;
    uint4 acc = make_uint4(init, init, init, init);
    for (int t = t0; t <= t1; ++t) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (int)v + t * stride));
      acc.x = morph_op4<OP>(acc.x, q.x);
      acc.y = morph_op4<OP>(acc.y, q.y);
      acc.z = morph_op4<OP>(acc.z, q.z);
      acc.w = morph_op4<OP>(acc.w, q.w);
    }
    *reinterpret_cast<uint4*>(dst + v) = acc;
  }
}

// axis 2 (along z), half-width R <= 3: 4 consecutive voxels per thread; the shifted windows come from
// the word itself and its two neighbours by funnel shifts.
template <int OP, int R>
__global__ void __launch_bounds__(ART_THREADS) morph_z_vec_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, unsigned n4, int sz) {
  const uint32_t* __restrict__ s32 = reinterpret_cast<const uint32_t*>(src);
  const unsigned wz = (unsigned)sz / 4u;  // words per row
  for (unsigned g = blockIdx.x * ART_THREADS + threadIdx.x; g < n4; g += gridDim.x * ART_THREADS) {
    const unsigned col = g % wz;
    const uint32_t w0 = __ldg(s32 + g);
    const uint32_t wm = col > 0 ? __ldg(s32 + g - 1) : 0u;
    const uint32_t wp = col + 1 < wz ? __ldg(s32 + g + 1) : 0u;
    uint32_t acc = w0;
#pragma unroll
    for (int t = 1; t <= R; ++t) {
      acc = morph_op4<OP>(acc, __funnelshift_r(w0, wp, 8 * t));        // voxels z + t
      acc = morph_op4<OP>(acc, __funnelshift_r(wm, w0, 32 - 8 * t));   // voxels z - t
    }
    reinterpret_cast<uint32_t*>(dst)[g] = acc;
  }
}

// squared / L1 distance to the nearest set voxel, separable min-plus passes with window r.
// pass over `axis`: dst = min_t src[v + t] + cost(t), cost = t^2 (SQ) or |t|; INF = 65535.
template <bool SQ, bool FIRST>
__global__ void __launch_bounds__(ART_THREADS) dist_axis_kernel(const void* __restrict__ src_, uint16_t* __restrict__ dst, int sx, int sy, int sz, int axis, int r) {
  const unsigned n = (unsigned)sx * sy * sz;
  const int stride = axis == 0 ? sy * sz : (axis == 1 ? sz : 1);
  const int len = axis == 0 ? sx : (axis == 1 ? sy : sz);
  const uint8_t* m8 = reinterpret_cast<const uint8_t*>(src_);
  const uint16_t* d16 = reinterpret_cast<const uint16_t*>(src_);
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    const int pos = axis == 2 ? (int)(v % (unsigned)sz) : (axis == 1 ? (int)((v / (unsigned)sz) % (unsigned)sy) : (int)(v / ((unsigned)sy * sz)));
    const int t0 = max(-r, -pos), t1 = min(r, len - 1 - pos);
    int best = 65535;
    for (int t = t0; t <= t1; ++t) {
      const int c = SQ ? t * t : abs(t);
      const int s = FIRST ? (m8[(int)v + t * stride] ? 0 : 65535) : (int)d16[(int)v + t * stride];
      best = min(best, s + c);
    }
    dst[v] = (uint16_t)min(best, 65535);
  }
}

__global__ void __launch_bounds__(ART_THREADS) thresh_u16_kernel(const uint16_t* __restrict__ d, uint8_t* __restrict__ out, int thr, unsigned n) {
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) out[v] = d[v] <= thr;
}

// fuzzy-boundary ring: ring = dil - mask, sub-sampled (keep-mask injected, or Bernoulli(p) from Philox)
__global__ void __launch_bounds__(ART_THREADS) ring_kernel(const uint8_t* __restrict__ dil, const uint8_t* __restrict__ mask, const uint8_t* __restrict__ keep, fsg_rng rng,
                                                           float p, uint8_t* __restrict__ out, unsigned n) {
  const Philox ph(rng.seed);
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    uint8_t o = 0;
    if (dil[v] && !mask[v]) {
      if (keep)
        o = keep[v] != 0;
      else {
        const uint4 w = ph(v, rng.stage, (uint32_t)rng.sample, (uint32_t)(rng.sample >> 32));
        o = (__uint_as_float(0x3f800000u | (w.x >> 9)) - 1.0f) < p;
      }
    }
    out[v] = o;
  }
}
// out = clamp(mask + (count > thr), 0, 1)
__global__ void __launch_bounds__(ART_THREADS) count_merge_kernel(const uint8_t* __restrict__ cnt, const uint8_t* __restrict__ mask, int thr, uint8_t* __restrict__ out, unsigned n) {
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) out[v] = (mask[v] || cnt[v] > thr) ? 1 : 0;
}

// SimulatedBoundaries final selection (artifacts.py:563-604):
//   p = (modif - halo > 0) ? mog : 0 ; idx = max(rint(p * len - 1), 0)
//   mask = modif && l1dist(halo) <= max(idx - 1, 0) ; out = x * mask
__global__ void __launch_bounds__(ART_THREADS) boundary_select_kernel(const float* __restrict__ x, const uint8_t* __restrict__ halo, const uint8_t* __restrict__ modif,
                                                                      const uint16_t* __restrict__ l1, const float* __restrict__ mog, int len, float* __restrict__ out,
                                                                      uint8_t* __restrict__ mask_out, unsigned n) {
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) {
    const bool surf = modif[v] && !halo[v];
    const float p = surf ? mog[v] : 0.f;
    const int idx = max((int)rintf(__fsub_rn(__fmul_rn(p, (float)len), 1.0f)), 0);
    const bool m = modif[v] && ((int)l1[v] <= max(idx - 1, 0));
    if (mask_out) mask_out[v] = m;
    if (out) out[v] = m ? x[v] : 0.f;
  }
}
__global__ void __launch_bounds__(ART_THREADS) mask_mul_kernel(const float* __restrict__ x, const uint8_t* __restrict__ m, float* __restrict__ out, unsigned n) {
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) out[v] = m[v] ? x[v] : 0.f;
}
__global__ void __launch_bounds__(ART_THREADS) label_mask_kernel(const uint8_t* __restrict__ lab, int match, uint8_t* __restrict__ out, unsigned n) {
  for (unsigned v = blockIdx.x * ART_THREADS + threadIdx.x; v < n; v += gridDim.x * ART_THREADS) out[v] = match < 0 ? (lab[v] > 0) : (lab[v] == match);
}

static unsigned art_grid(int64_t n) {
  const int64_t want = (n + ART_THREADS - 1) / ART_THREADS;
  return (unsigned)(want < 148 * 16 ? (want < 1 ? 1 : want) : 148 * 16);
}
static int check_shape(const char* who, int sx, int sy, int sz) {
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1 && (int64_t)sx * sy * sz < ((int64_t)1 << 31), "%s: bad shape", who);
  return 0;
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_mog(const float* centers, const float* sigmas, int n, int sx, int sy, int sz, float* out, const float* a, const float* b, float* dst, void* stream) {
  if (int rc = check_shape("fsg_mog", sx, sy, sz)) return rc;
  FSG_REQUIRE(n >= 0 && n <= MOG_MAX && (n == 0 || (centers && sigmas)), "fsg_mog: n=%d outside [0,%d] or NULL tables", n, MOG_MAX);
  FSG_REQUIRE(out || dst, "fsg_mog: no output");
  FSG_REQUIRE(!dst || (a && b), "fsg_mog: blend needs both inputs");
  const int rows = sx * sy;
  mog_kernel<<<rows < 148 * 8 ? rows : 148 * 8, ART_THREADS, 0, as_stream(stream)>>>(centers, sigmas, n, sx, sy, sz, out, a, b, dst);
  return check_launch("fsg_mog");
}

extern "C" int fsg_sample_voxels(const fsg_sample_job* job, int sx, int sy, int sz, void* stream) {
  if (int rc = check_shape("fsg_sample_voxels", sx, sy, sz)) return rc;
  FSG_REQUIRE(job && job->labels && job->centers_out && job->count_out && job->workspace, "fsg_sample_voxels: NULL pointer");
  FSG_REQUIRE(job->k >= 1 && job->k <= 1024, "fsg_sample_voxels: k=%d outside [1,1024]", job->k);
  FSG_REQUIRE(job->nprior >= 0 && job->nprior <= 4, "fsg_sample_voxels: nprior outside [0,4]");
  FSG_REQUIRE(job->workspace_bytes >= 64 + 2048 * sizeof(float2), "fsg_sample_voxels: workspace too small");
  SampleSpec sp;
  memset(&sp, 0, sizeof(sp));
  sp.lab = job->labels;
  sp.lab2 = job->labels2;
  sp.match = job->match;
  sp.nprior = job->nprior;
  memcpy(sp.pc, job->prior_centers, sizeof(sp.pc));
  memcpy(sp.ps, job->prior_sigmas, sizeof(sp.ps));
  cudaStream_t s = as_stream(stream);
  char* ws = reinterpret_cast<char*>(job->workspace);
  double* total = reinterpret_cast<double*>(ws);
  unsigned* count = reinterpret_cast<unsigned*>(ws + 8);
  unsigned* ncand = reinterpret_cast<unsigned*>(ws + 12);
  float2* cand = reinterpret_cast<float2*>(ws + 64);
  cudaMemsetAsync(ws, 0, 64, s);
  const unsigned g = art_grid((int64_t)sx * sy * sz);
  sample_total_kernel<<<g, ART_THREADS, 0, s>>>(sp, sx, sy, sz, total, count);
  sample_keys_kernel<<<g, ART_THREADS, 0, s>>>(sp, sx, sy, sz, job->rng, job->k, total, cand, ncand, 2048u);
  sample_pick_kernel<<<1, 1024, 0, s>>>(cand, ncand, 2048u, job->k, sy, sz, job->_pad, job->centers_out, job->count_out);
  return check_launch("fsg_sample_voxels");
}

extern "C" int fsg_perlin(const fsg_perlin_octave* octs, int noct, int sx, int sy, int sz, float* out, float* minmax, void* stream) {
  if (int rc = check_shape("fsg_perlin", sx, sy, sz)) return rc;
  FSG_REQUIRE(octs && out && minmax && noct >= 1 && noct <= 8, "fsg_perlin: bad arguments");
  PerlinArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.noct = noct;
  for (int o = 0; o < noct; ++o) {
    FSG_REQUIRE(octs[o].grad && octs[o].lin[0] && octs[o].lin[1] && octs[o].lin[2], "fsg_perlin: octave %d has a NULL table", o);
    pa.o[o].grad = octs[o].grad;
    for (int a = 0; a < 3; ++a) {
      pa.o[o].lin[a] = octs[o].lin[a];
      pa.o[o].res[a] = octs[o].res[a];
      FSG_REQUIRE(octs[o].res[a] >= 1, "fsg_perlin: octave %d res must be >= 1", o);
    }
    pa.o[o].amp = octs[o].amp;
  }
  cudaStream_t s = as_stream(stream);
  minmax_init_kernel<<<1, 32, 0, s>>>(minmax, 1);
  perlin_kernel<<<art_grid((int64_t)sx * sy * sz), ART_THREADS, 0, s>>>(pa, sx, sy, sz, out, minmax);
  minmax_final_kernel<<<1, 32, 0, s>>>(minmax, 1);
  return check_launch("fsg_perlin");
}

extern "C" int fsg_struct_blend(const float* x, const uint8_t* seg, const float* lr, const float* perlin, const float* scal, float noise_std, float increase, float* out,
                                int64_t n, void* stream) {
  FSG_REQUIRE(x && seg && lr && perlin && scal && out && n > 0 && n < ((int64_t)1 << 31), "fsg_struct_blend: bad arguments");
  struct_blend_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(x, seg, lr, perlin, scal, noise_std, increase, out, (unsigned)n);
  return check_launch("fsg_struct_blend");
}

// Packed form of dist_axis_kernel for axis 0 / 1 (bit-identical): 8 consecutive z voxels per thread,
// saturating halfword add + unsigned halfword min (min(s + c, 65535) == saturating add).
template <bool SQ, bool FIRST>
__global__ void __launch_bounds__(ART_THREADS) dist_axis_vec_kernel(const void* __restrict__ src_, uint16_t* __restrict__ dst, int sx, int sy, int sz, int axis, int r) {
  const unsigned n8 = (unsigned)sx * sy * sz / 8u;
  const int stride = axis == 0 ? sy * sz : sz;
  const int len = axis == 0 ? sx : sy;
  const uint8_t* m8 = reinterpret_cast<const uint8_t*>(src_);
  const uint16_t* d16 = reinterpret_cast<const uint16_t*>(src_);
  for (unsigned g = blockIdx.x * ART_THREADS + threadIdx.x; g < n8; g += gridDim.x * ART_THREADS) {
    const unsigned v = g * 8u;
    const int pos = axis == 1 ? (int)((v / (unsigned)sz) % (unsigned)sy) : (int)(v / ((unsigned)sy * sz));
    const int t0 = max(-r, -pos), t1 = min(r, len - 1 - pos);
    uint4 best = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    for (int t = t0; t <= t1; ++t) {
      const uint32_t c = (uint32_t)(SQ ? t * t : abs(t));
      const uint32_t c2 = c | (c << 16);
      uint4 q;
      if (FIRST) {
        const uint2 m = __ldg(reinterpret_cast<const uint2*>(m8 + (int)v + t * stride));
        const uint32_t z0 = __vcmpeq4(m.x, 0u), z1 = __vcmpeq4(m.y, 0u);  // 0xff where the mask is clear -> distance 65535
        q = make_uint4(__byte_perm(z0, 0u, 0x1100), __byte_perm(z0, 0u, 0x3322), __byte_perm(z1, 0u, 0x1100), __byte_perm(z1, 0u, 0x3322));
      } else {
        q = __ldg(reinterpret_cast<const uint4*>(d16 + (int)v + t * stride));
      }
      best.x = __vminu2(best.x, __vaddus2(q.x, c2));
      best.y = __vminu2(best.y, __vaddus2(q.y, c2));
      best.z = __vminu2(best.z, __vaddus2(q.z, c2));
      best.w = __vminu2(best.w, __vaddus2(q.w, c2));
    }
    *reinterpret_cast<uint4*>(dst + v) = best;
  }
}

// op: 0 = box dilation (max), 1 = box erosion (min, zero padded), 2 = box count (sum, k^3 <= 255)
extern "C" int fsg_morph_box(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int k, int op, int sx, int sy, int sz, void* stream) {
  if (int rc = check_shape("fsg_morph_box", sx, sy, sz)) return rc;
  FSG_REQUIRE(src && dst && tmp && src != dst && src != tmp && dst != tmp, "fsg_morph_box: needs three distinct buffers");
  FSG_REQUIRE(k >= 1 && k % 2 == 1 && k <= 31 && op >= 0 && op <= 2 && (op != 2 || k * k * k <= 255), "fsg_morph_box: bad kernel size / op");
  cudaStream_t s = as_stream(stream);
  const unsigned g = art_grid((int64_t)sx * sy * sz);
  const int r = k / 2;
  const uint8_t* in[3] = {src, dst, tmp};
  uint8_t* outp[3] = {dst, tmp, dst};
  // packed-voxel kernels when rows are whole 16-byte words (bit-identical results)
  const bool vec = sz % 16 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(tmp)) & 15) == 0;
  const int64_t nvox = (int64_t)sx * sy * sz;
#define FSG_MORPH_OPS(CALL)                 \
  do {                                      \
    if (op == 0) { CALL(MORPH_MAX); }       \
    else if (op == 1) { CALL(MORPH_MIN); }  \
    else { CALL(MORPH_SUM); }               \
  } while (0)
  for (int a = 0; a < 3; ++a) {
    if (vec && a < 2) {
      const unsigned gv = art_grid(nvox / 16);
#define FSG_CALL(OP) morph_axis_vec_kernel<OP><<<gv, ART_THREADS, 0, s>>>(in[a], outp[a], sx, sy, sz, a, r)
      FSG_MORPH_OPS(FSG_CALL);
#undef FSG_CALL
    } else if (vec && r >= 1 && r <= 3) {
      const unsigned gv = art_grid(nvox / 4);
      const unsigned n4 = (unsigned)(nvox / 4);
#define FSG_CALL(OP)                                                                              \
  do {                                                                                            \
    if (r == 1) morph_z_vec_kernel<OP, 1><<<gv, ART_THREADS, 0, s>>>(in[a], outp[a], n4, sz);      \
    else if (r == 2) morph_z_vec_kernel<OP, 2><<<gv, ART_THREADS, 0, s>>>(in[a], outp[a], n4, sz); \
    else morph_z_vec_kernel<OP, 3><<<gv, ART_THREADS, 0, s>>>(in[a], outp[a], n4, sz);             \
  } while (0)
      FSG_MORPH_OPS(FSG_CALL);
#undef FSG_CALL
    } else {
#define FSG_CALL(OP) morph_axis_kernel<OP><<<g, ART_THREADS, 0, s>>>(in[a], outp[a], sx, sy, sz, a, r)
      FSG_MORPH_OPS(FSG_CALL);
#undef FSG_CALL
    }
  }
#undef FSG_MORPH_OPS
  return check_launch("fsg_morph_box");
}

// metric 0: squared Euclidean distance (ball dilation: dist <= r^2), 1: L1 distance; both capped by
// the search window r (voxels farther than r along an axis read 65535).  dist_out/tmp: uint16 volumes.
extern "C" int fsg_morph_dist(const uint8_t* mask, uint16_t* dist_out, uint16_t* tmp, int r, int metric, int sx, int sy, int sz, void* stream) {
  if (int rc = check_shape("fsg_morph_dist", sx, sy, sz)) return rc;
  FSG_REQUIRE(mask && dist_out && tmp && dist_out != tmp, "fsg_morph_dist: bad buffers");
  FSG_REQUIRE(r >= 1 && r <= 120 && (metric == 0 || metric == 1), "fsg_morph_dist: bad radius / metric");
  cudaStream_t s = as_stream(stream);
  const unsigned g = art_grid((int64_t)sx * sy * sz);
  // packed halfword kernels for the two strided passes when rows are whole 16-byte words
  const bool vec = sz % 8 == 0 && ((reinterpret_cast<uintptr_t>(mask) & 7) | ((reinterpret_cast<uintptr_t>(dist_out) | reinterpret_cast<uintptr_t>(tmp)) & 15)) == 0;
  const unsigned gv = art_grid((int64_t)sx * sy * sz / 8);
  if (metric == 0) {
    if (vec) {
      dist_axis_vec_kernel<true, true><<<gv, ART_THREADS, 0, s>>>(mask, dist_out, sx, sy, sz, 0, r);
      dist_axis_vec_kernel<true, false><<<gv, ART_THREADS, 0, s>>>(dist_out, tmp, sx, sy, sz, 1, r);
    } else {
      dist_axis_kernel<true, true><<<g, ART_THREADS, 0, s>>>(mask, dist_out, sx, sy, sz, 0, r);
      dist_axis_kernel<true, false><<<g, ART_THREADS, 0, s>>>(dist_out, tmp, sx, sy, sz, 1, r);
    }
    dist_axis_kernel<true, false><<<g, ART_THREADS, 0, s>>>(tmp, dist_out, sx, sy, sz, 2, r);
  } else {
    if (vec) {
      dist_axis_vec_kernel<false, true><<<gv, ART_THREADS, 0, s>>>(mask, dist_out, sx, sy, sz, 0, r);
      dist_axis_vec_kernel<false, false><<<gv, ART_THREADS, 0, s>>>(dist_out, tmp, sx, sy, sz, 1, r);
    } else {
      dist_axis_kernel<false, true><<<g, ART_THREADS, 0, s>>>(mask, dist_out, sx, sy, sz, 0, r);
      dist_axis_kernel<false, false><<<g, ART_THREADS, 0, s>>>(dist_out, tmp, sx, sy, sz, 1, r);
    }
    dist_axis_kernel<false, false><<<g, ART_THREADS, 0, s>>>(tmp, dist_out, sx, sy, sz, 2, r);
  }
  return check_launch("fsg_morph_dist");
}

extern "C" int fsg_morph_thresh(const uint16_t* dist, uint8_t* out, int thr, int64_t n, void* stream) {
  FSG_REQUIRE(dist && out && n > 0 && n < ((int64_t)1 << 31), "fsg_morph_thresh: bad arguments");
  thresh_u16_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(dist, out, thr, (unsigned)n);
  return check_launch("fsg_morph_thresh");
}

extern "C" int fsg_morph_ring(const uint8_t* dilated, const uint8_t* mask, const uint8_t* keep, fsg_rng rng, float p, uint8_t* out, int64_t n, void* stream) {
  FSG_REQUIRE(dilated && mask && out && n > 0 && n < ((int64_t)1 << 31), "fsg_morph_ring: bad arguments");
  ring_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(dilated, mask, keep, rng, p, out, (unsigned)n);
  return check_launch("fsg_morph_ring");
}

extern "C" int fsg_morph_count_merge(const uint8_t* count, const uint8_t* mask, int thr, uint8_t* out, int64_t n, void* stream) {
  FSG_REQUIRE(count && mask && out && n > 0 && n < ((int64_t)1 << 31), "fsg_morph_count_merge: bad arguments");
  count_merge_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(count, mask, thr, out, (unsigned)n);
  return check_launch("fsg_morph_count_merge");
}

extern "C" int fsg_boundary_select(const float* x, const uint8_t* halo, const uint8_t* modif, const uint16_t* l1, const float* mog, int len, float* out, uint8_t* mask_out,
                                   int64_t n, void* stream) {
  FSG_REQUIRE(halo && modif && l1 && mog && (out || mask_out) && (!out || x) && n > 0 && n < ((int64_t)1 << 31) && len >= 1, "fsg_boundary_select: bad arguments");
  boundary_select_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(x, halo, modif, l1, mog, len, out, mask_out, (unsigned)n);
  return check_launch("fsg_boundary_select");
}

extern "C" int fsg_mask_mul(const float* x, const uint8_t* mask, float* out, int64_t n, void* stream) {
  FSG_REQUIRE(x && mask && out && n > 0 && n < ((int64_t)1 << 31), "fsg_mask_mul: bad arguments");
  mask_mul_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(x, mask, out, (unsigned)n);
  return check_launch("fsg_mask_mul");
}

extern "C" int fsg_label_mask(const uint8_t* labels, int match, uint8_t* out, int64_t n, void* stream) {
  FSG_REQUIRE(labels && out && n > 0 && n < ((int64_t)1 << 31), "fsg_label_mask: bad arguments");
  label_mask_kernel<<<art_grid(n), ART_THREADS, 0, as_stream(stream)>>>(labels, match, out, (unsigned)n);
  return check_launch("fsg_label_mask");
}
