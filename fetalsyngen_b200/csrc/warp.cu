// K2 — fused spatial deformation (+ gamma / bias-field epilogues).
//
// One launch does what the reference spreads over ~60 torch kernels and 1.5 GB of
// transients (generator/deformation/affine_nonrigid.py:64-84, 164-193, 299-366 and
// utils/generation.py:204-288, 310-397):
//   F      = separable linear up-sampling of the control grid Fsmall (x, then y, then z;
//            every w_f*a + w_c*b rounded exactly like myzoom_torch's three loops)
//   coord  = ((A_r0*(xc+Fx) + A_r1*(yc+Fy)) + A_r2*(zc+Fz)) + c2_r, clamp [0,S-1], -shift
//   image  = trilinear gather (blend x, y, z; 0 where any coord <= 0)
//   seg    = nearest gather (round-half-even)
//   flip   = sources mirrored along x before sampling
//   image  = 300*(image/300)^gamma ; image *= exp(zoom(bf_low))       (optional epilogues)
// No coordinate / field volume ever exists in HBM.  Algorithmic HBM bytes: read img 4 + seg 1,
// write img 4 + seg 1 = 10 B/voxel.
//
// Work decomposition (round-1 ncu: the first version was issue-bound at ~550 instr/voxel):
//   * a block owns WX x-planes x WY rows x the whole z extent, so the x/y blends of the two
//     control grids (phase A) are amortised over S_z voxels per row instead of 32;
//   * coordinates use separately rounded mul/add (bit-exact segmentation), but floor/round are
//     magic-number adds on the FP32 pipe instead of F2I/I2F conversions, the flip is folded
//     into the x stride, and the float image path uses FMA lerps and ex2/lg2 approximations
//     (image parity is a tolerance, not bit-exactness);
//   * floor(min coordinate) (the reference's crop-shift quirk) is resolved by a pre-pass over
//     the twelve edges of the volume; the full-volume pre-pass only runs for jobs whose edges do
//     not already prove the shift to be 0.
#include "common.cuh"

namespace fsg {

constexpr int WX = 8, WY = 4;
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

// Exact control-grid value at voxel (i,j,k): x-, y-, z-blend in myzoom_torch's order.  Used by
// the edge pre-pass only (the main kernel stages the x/y blends in shared memory).
__device__ __forceinline__ void field_at(const fsg_warp_job& job, int i, int j, int k, float& fx, float& fy, float& fz) {
  const Tab tx = load_tab(job.ftab[0], i), ty = load_tab(job.ftab[1], j), tz = load_tab(job.ftab[2], k);
  const int n1 = job.fs[1], n2 = job.fs[2];
  float out[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float* g = job.fsmall + ch;
    auto at = [&](int a, int b, int c) { return __ldg(g + ((size_t)(a * n1 + b) * n2 + c) * 3); };
    const float a_ff = blend(tx.wf, at(tx.f, ty.f, tz.f), tx.wc, at(tx.c, ty.f, tz.f));
    const float a_fc = blend(tx.wf, at(tx.f, ty.f, tz.c), tx.wc, at(tx.c, ty.f, tz.c));
    const float a_cf = blend(tx.wf, at(tx.f, ty.c, tz.f), tx.wc, at(tx.c, ty.c, tz.f));
    const float a_cc = blend(tx.wf, at(tx.f, ty.c, tz.c), tx.wc, at(tx.c, ty.c, tz.c));
    const float b_f = blend(ty.wf, a_ff, ty.wc, a_cf);
    const float b_c = blend(ty.wf, a_fc, ty.wc, a_cc);
    out[ch] = blend(tz.wf, b_f, tz.wc, b_c);
  }
  fx = out[0];
  fy = out[1];
  fz = out[2];
}

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

// ---------------------------------------------------------------------------------- edge pre-pass
// One thread per voxel of the twelve edges of the volume; atomicMin of the clamped coordinates
// into job.shift.  The minimum of an affine map over a box sits on its boundary, and with the
// smooth control-grid field added it stays next to it, so an edge voxel with a coordinate < 1
// proves floor(min) == 0 for that axis (coordinates are clamped at 0).  This is only a
// sufficient test: jobs it does not resolve run the exact full-volume pass.
__global__ void __launch_bounds__(128) shift_edges_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int sx, int sy, int sz) {
  const fsg_warp_job& job = batch.j[blockIdx.y];
  const int total = 4 * (sx + sy + sz);
  const float inf = __int_as_float(0x7f800000);
  float mnx = inf, mny = inf, mnz = inf;
  const Affine aff(job);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    int i, j, k, r = t;
    if (r < 4 * sx) {
      i = r >> 2;
      j = (r & 1) ? sy - 1 : 0;
      k = (r & 2) ? sz - 1 : 0;
    } else if ((r -= 4 * sx) < 4 * sy) {
      j = r >> 2;
      i = (r & 1) ? sx - 1 : 0;
      k = (r & 2) ? sz - 1 : 0;
    } else {
      r -= 4 * sy;
      k = r >> 2;
      i = (r & 1) ? sx - 1 : 0;
      j = (r & 2) ? sy - 1 : 0;
    }
    float x1 = sub_rn((float)i, job.center[0]), y1 = sub_rn((float)j, job.center[1]), z1 = sub_rn((float)k, job.center[2]);
    if (job.fsmall != nullptr) {
      float fx, fy, fz;
      field_at(job, i, j, k, fx, fy, fz);
      x1 = add_rn(x1, fx);
      y1 = add_rn(y1, fy);
      z1 = add_rn(z1, fz);
    }
    float ii, jj, kk;
    affine_clamp(aff, x1, y1, z1, (float)(sx - 1), (float)(sy - 1), (float)(sz - 1), ii, jj, kk);
    mnx = fminf(mnx, ii);
    mny = fminf(mny, jj);
    mnz = fminf(mnz, kk);
  }
  mnx = warp_min(mnx);
  mny = warp_min(mny);
  mnz = warp_min(mnz);
  if ((threadIdx.x & 31) == 0) {
    // clamped coordinates are >= 0, so the int view of the float orders correctly
    int* sh = reinterpret_cast<int*>(const_cast<float*>(job.shift));
    atomicMin(sh + 0, __float_as_int(mnx));
    atomicMin(sh + 1, __float_as_int(mny));
    atomicMin(sh + 2, __float_as_int(mnz));
  }
}

// ---------------------------------------------------------------------------------- main kernel
struct ZTab {  // per-thread z table entry of a control grid, pre-scaled for float4 rows
  int f, c;
  float wc, wf;
};

template <int PASS, bool IMG2>
__global__ void __launch_bounds__(WARP_THREADS) warp_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int sx, int sy, int sz, float* __restrict__ dbg_x,
                                                             float* __restrict__ dbg_y, float* __restrict__ dbg_z) {
  const fsg_warp_job& job = batch.j[blockIdx.z];
  const int tid = threadIdx.x;

  extern __shared__ float4 s_dyn[];
  float4* s_f = s_dyn;  // [WX*WY][fz_n] control grid blended along x and y (xyz + pad)
  __shared__ float s_b[WX * WY][MAX_BZ];  // bias grid blended along x and y
  __shared__ float s_red[3][WARP_THREADS / 32];

  const bool deform = job.mode == 1;
  const bool has_field = deform && job.fsmall != nullptr;
  const bool has_bias = (PASS == PASS_WARP) && job.bf_low != nullptr && job.dst_img != nullptr;

  if (PASS == PASS_SHIFT) {
    // the edge pre-pass already proved floor(min) == 0 on every axis: nothing to do
    const float* sh = job.shift;
    if (sh[0] < 1.f && sh[1] < 1.f && sh[2] < 1.f) return;
  }

  const int fz_n = has_field ? job.fs[2] : 0;
  const float inf = __int_as_float(0x7f800000);
  float mnx = inf, mny = inf, mnz = inf;

  const float shx = (PASS != PASS_SHIFT && deform) ? job.shift[0] : 0.f;
  const float shy = (PASS != PASS_SHIFT && deform) ? job.shift[1] : 0.f;
  const float shz = (PASS != PASS_SHIFT && deform) ? job.shift[2] : 0.f;
  const bool has_shift = (shx != 0.f) || (shy != 0.f) || (shz != 0.f);
  const float mx = (float)(sx - 1), my = (float)(sy - 1), mz = (float)(sz - 1);
  // flip folded into the x stride: element offset of source plane p is xb + p * xs
  const int plane = sy * sz;
  const int xs = job.flip ? -plane : plane;
  const int xb = job.flip ? (sx - 1) * plane : 0;
  const float gamma = job.gamma;
  const bool has_gamma = (PASS == PASS_WARP) && job.has_gamma;
  const Affine aff(job);
  const float cen_x = job.center[0], cen_y = job.center[1], cen_z = job.center[2];
  const float* __restrict__ const src_img = job.src_img;
  const float* __restrict__ const src_img2 = IMG2 ? job.src_img2 : nullptr;
  const uint8_t* __restrict__ const src_seg = job.src_seg;
  float* __restrict__ const dst_img = job.dst_img;
  float* __restrict__ const dst_img2 = IMG2 ? job.dst_img2 : nullptr;
  uint8_t* __restrict__ const dst_seg = job.dst_seg;
  const bool do_img = dst_img != nullptr, do_seg = dst_seg != nullptr, do_img2 = IMG2 && dst_img2 != nullptr;

  const int ntile_y = (sy + WY - 1) / WY;
  const int ntiles = ntile_y * ((sx + WX - 1) / WX);
  // PASS_SHIFT runs a bounded grid and strides over the tiles; the other passes use one tile per block
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int x0 = (tile / ntile_y) * WX, y0 = (tile % ntile_y) * WY;
    if (tile != (int)blockIdx.x) __syncthreads();  // previous tile's readers are done

    // ---- phase A: x- and y-blends of the low-resolution grids for the tile's WX*WY rows
    if (has_field) {
      const int fy_n = job.fs[1];
      const int per_row = fz_n * 3;
      float* sf = reinterpret_cast<float*>(s_f);
      for (int e = tid; e < WX * WY * per_row; e += WARP_THREADS) {
        const int row = e / per_row, rem = e - row * per_row;
        const int rx = row / WY, ry = row - rx * WY;
        const int i = min(x0 + rx, sx - 1), j = min(y0 + ry, sy - 1);
        const Tab tx = load_tab(job.ftab[0], i), ty = load_tab(job.ftab[1], j);
        const float* g = job.fsmall + rem;  // rem = zc*3 + ch
        const int sxs = fy_n * per_row;
        const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * per_row), tx.wc, __ldg(g + tx.c * sxs + ty.f * per_row));
        const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * per_row), tx.wc, __ldg(g + tx.c * sxs + ty.c * per_row));
        const int zc = rem / 3, ch = rem - zc * 3;
        sf[(row * fz_n + zc) * 4 + ch] = blend(ty.wf, t1f, ty.wc, t1c);
      }
    }
    if (has_bias) {
      const int by_n = job.bs[1], bz_n = job.bs[2];
      for (int e = tid; e < WX * WY * bz_n; e += WARP_THREADS) {
        const int row = e / bz_n, zc = e - row * bz_n;
        const int rx = row / WY, ry = row - rx * WY;
        const int i = min(x0 + rx, sx - 1), j = min(y0 + ry, sy - 1);
        const Tab tx = load_tab(job.btab[0], i), ty = load_tab(job.btab[1], j);
        const float* g = job.bf_low + zc;
        const int sxs = by_n * bz_n;
        const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.f * bz_n));
        const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.c * bz_n));
        s_b[row][zc] = blend(ty.wf, t1f, ty.wc, t1c);
      }
    }
    __syncthreads();

    // ---- phase B: thread = z column (strided by the block size), marching over the tile's rows
    for (int k = tid; k < sz; k += WARP_THREADS) {
      ZTab tfz = {0, 0, 0.f, 1.f}, tbz = {0, 0, 0.f, 1.f};
      if (has_field) {
        const Tab t = load_tab(job.ftab[2], k);
        tfz.f = t.f; tfz.c = t.c; tfz.wc = t.wc; tfz.wf = t.wf;
      }
      if (has_bias) {
        const Tab t = load_tab(job.btab[2], k);
        tbz.f = t.f; tbz.c = t.c; tbz.wc = t.wc; tbz.wf = t.wf;
      }
      const float zc = sub_rn((float)k, cen_z);
#pragma unroll 1
      for (int ry = 0; ry < WY; ++ry) {
        const int j = y0 + ry;
        if (j >= sy) break;
        const float yc = sub_rn((float)j, cen_y);
#pragma unroll 2
        for (int rx = 0; rx < WX; ++rx) {
          const int i = x0 + rx;
          if (i >= sx) break;
          const int row = rx * WY + ry;
          const unsigned o = (unsigned)((i * sy + j) * sz + k);
          float ii, jj, kk;
          if (deform) {
            float x1 = sub_rn((float)i, cen_x), y1 = yc, z1 = zc;
            if (has_field) {
              const float4 f0 = s_f[row * fz_n + tfz.f], f1 = s_f[row * fz_n + tfz.c];
              x1 = add_rn(x1, blend(tfz.wf, f0.x, tfz.wc, f1.x));
              y1 = add_rn(y1, blend(tfz.wf, f0.y, tfz.wc, f1.y));
              z1 = add_rn(z1, blend(tfz.wf, f0.z, tfz.wc, f1.z));
            }
            affine_clamp(aff, x1, y1, z1, mx, my, mz, ii, jj, kk);
            if (PASS == PASS_SHIFT) {
              mnx = fminf(mnx, ii);
              mny = fminf(mny, jj);
              mnz = fminf(mnz, kk);
              continue;
            }
            if (has_shift) {
              ii = sub_rn(ii, shx);
              jj = sub_rn(jj, shy);
              kk = sub_rn(kk, shz);
            }
          } else {
            ii = (float)i;
            jj = (float)j;
            kk = (float)k;
          }
          if (PASS == PASS_COORDS) {
            dbg_x[o] = ii;
            dbg_y[o] = jj;
            dbg_z[o] = kk;
            continue;
          }
          if (PASS == PASS_WARP) {
            if (!deform) {
              // identity sampling: only flip + epilogues (deformation gate off)
              const unsigned so = (unsigned)(xb + i * xs + j * sz + k);
              if (do_img) {
                float v = __ldg(src_img + so);
                if (has_gamma) v = mul_rn(300.0f, ex2_approx(mul_rn(gamma, lg2_approx(mul_rn(v, 1.0f / 300.0f)))));
                if (has_bias) v = mul_rn(v, ex2_approx(mul_rn(1.4426950408889634f, blend(tbz.wf, s_b[row][tbz.f], tbz.wc, s_b[row][tbz.c]))));
                dst_img[o] = v;
              }
              if (do_seg) dst_seg[o] = __ldg(src_seg + so);
              if (do_img2) dst_img2[o] = __ldg(src_img2 + so);
              continue;
            }
            if (do_img || do_img2) {
              // coordinates are in [0, S-1]: floor via a round-toward-zero magic add
              const float tx_ = __fadd_rz(ii, MAGIC), ty_ = __fadd_rz(jj, MAGIC), tz_ = __fadd_rz(kk, MAGIC);
              const int fx = __float_as_int(tx_) - 0x4B000000, fy = __float_as_int(ty_) - 0x4B000000, fz = __float_as_int(tz_) - 0x4B000000;
              const float wcx = sub_rn(ii, sub_rn(tx_, MAGIC)), wcy = sub_rn(jj, sub_rn(ty_, MAGIC)), wcz = sub_rn(kk, sub_rn(tz_, MAGIC));
              const int cx = min(fx + 1, sx - 1), cy = min(fy + 1, sy - 1), cz = min(fz + 1, sz - 1);
              const bool ok = fminf(fminf(ii, jj), kk) > 0.f;
              const int of = xb + fx * xs, oc = xb + cx * xs;
              const int rf = fy * sz, rc = cy * sz;
              // all element offsets are non-negative: unsigned indices keep the address math 32-bit
              const unsigned i000 = (unsigned)(of + rf + fz), i001 = (unsigned)(of + rf + cz);
              const unsigned i100 = (unsigned)(oc + rf + fz), i101 = (unsigned)(oc + rf + cz);
              const unsigned i010 = (unsigned)(of + rc + fz), i011 = (unsigned)(of + rc + cz);
              const unsigned i110 = (unsigned)(oc + rc + fz), i111 = (unsigned)(oc + rc + cz);
              if (do_img) {
                const float c000 = __ldg(src_img + i000), c001 = __ldg(src_img + i001);
                const float c100 = __ldg(src_img + i100), c101 = __ldg(src_img + i101);
                const float c010 = __ldg(src_img + i010), c011 = __ldg(src_img + i011);
                const float c110 = __ldg(src_img + i110), c111 = __ldg(src_img + i111);
                const float c00 = lerp_fma(c000, c100, wcx), c01 = lerp_fma(c001, c101, wcx);
                const float c10 = lerp_fma(c010, c110, wcx), c11 = lerp_fma(c011, c111, wcx);
                float v = lerp_fma(lerp_fma(c00, c10, wcy), lerp_fma(c01, c11, wcy), wcz);
                v = ok ? v : 0.f;
                if (has_gamma) v = mul_rn(300.0f, ex2_approx(mul_rn(gamma, lg2_approx(mul_rn(v, 1.0f / 300.0f)))));
                if (has_bias) v = mul_rn(v, ex2_approx(mul_rn(1.4426950408889634f, blend(tbz.wf, s_b[row][tbz.f], tbz.wc, s_b[row][tbz.c]))));
                dst_img[o] = v;
              }
              if (do_img2) {
                const float c000 = __ldg(src_img2 + i000), c001 = __ldg(src_img2 + i001);
                const float c100 = __ldg(src_img2 + i100), c101 = __ldg(src_img2 + i101);
                const float c010 = __ldg(src_img2 + i010), c011 = __ldg(src_img2 + i011);
                const float c110 = __ldg(src_img2 + i110), c111 = __ldg(src_img2 + i111);
                const float c00 = lerp_fma(c000, c100, wcx), c01 = lerp_fma(c001, c101, wcx);
                const float c10 = lerp_fma(c010, c110, wcx), c11 = lerp_fma(c011, c111, wcx);
                const float v = lerp_fma(lerp_fma(c00, c10, wcy), lerp_fma(c01, c11, wcy), wcz);
                dst_img2[o] = ok ? v : 0.f;
              }
            }
            if (do_seg) {
              // round-half-even of a coordinate in [0, S-1] via a round-to-nearest magic add
              const int ir = __float_as_int(add_rn(ii, MAGIC)) - 0x4B000000;
              const int jr = __float_as_int(add_rn(jj, MAGIC)) - 0x4B000000;
              const int kr = __float_as_int(add_rn(kk, MAGIC)) - 0x4B000000;
              dst_seg[o] = __ldg(src_seg + (unsigned)(xb + ir * xs + jr * sz + kr));
            }
          }
        }
      }
    }
  }

  if (PASS == PASS_SHIFT) {
    mnx = warp_min(mnx);
    mny = warp_min(mny);
    mnz = warp_min(mnz);
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) {
      s_red[0][w] = mnx;
      s_red[1][w] = mny;
      s_red[2][w] = mnz;
    }
    __syncthreads();
    if (tid < 3) {
      float m = inf;
      for (int q = 0; q < WARP_THREADS / 32; ++q) m = fminf(m, s_red[tid][q]);
      atomicMin(reinterpret_cast<int*>(const_cast<float*>(job.shift)) + tid, __float_as_int(m));
    }
  }
}

__global__ void shift_init_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int njobs) {
  const int t = threadIdx.x;
  if (t < njobs * 3) const_cast<float*>(batch.j[t / 3].shift)[t % 3] = __int_as_float(0x7f800000);
}
__global__ void shift_final_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int njobs) {
  const int t = threadIdx.x;
  if (t < njobs * 3) {
    float* p = const_cast<float*>(batch.j[t / 3].shift) + t % 3;
    *p = floorf(*p);
  }
}

static int validate(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, bool need_io, const char* who) {
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1 && sx <= 32767 && sy <= 32767 && sz <= 32767, "%s: bad shape %dx%dx%d", who, sx, sy, sz);
  FSG_REQUIRE((int64_t)sx * sy * sz < ((int64_t)1 << 31), "%s: volume exceeds 2^31 voxels", who);
  for (int n = 0; n < njobs; ++n) {
    const fsg_warp_job& j = jobs[n];
    FSG_REQUIRE(j.mode == 0 || j.mode == 1, "%s: job %d mode must be 0 or 1", who, n);
    if (j.mode == 1) FSG_REQUIRE(j.shift != nullptr, "%s: job %d needs a shift buffer", who, n);
    if (j.fsmall) {
      FSG_REQUIRE(j.ftab[0] && j.ftab[1] && j.ftab[2], "%s: job %d control grid without zoom tables", who, n);
      FSG_REQUIRE(j.fs[0] >= 1 && j.fs[1] >= 1 && j.fs[2] >= 1 && j.fs[2] <= MAX_FZ, "%s: job %d control grid z-extent %d outside [1,%d]", who, n, j.fs[2], MAX_FZ);
    }
    if (j.bf_low) {
      FSG_REQUIRE(j.btab[0] && j.btab[1] && j.btab[2], "%s: job %d bias grid without zoom tables", who, n);
      FSG_REQUIRE(j.bs[0] >= 1 && j.bs[1] >= 1 && j.bs[2] >= 1 && j.bs[2] <= MAX_BZ, "%s: job %d bias grid z-extent %d outside [1,%d]", who, n, j.bs[2], MAX_BZ);
    }
    if (need_io) {
      FSG_REQUIRE(j.dst_img || j.dst_seg || j.dst_img2, "%s: job %d has no output", who, n);
      FSG_REQUIRE(!j.dst_img || j.src_img, "%s: job %d dst_img without src_img", who, n);
      FSG_REQUIRE(!j.dst_seg || j.src_seg, "%s: job %d dst_seg without src_seg", who, n);
      FSG_REQUIRE(!j.dst_img2 || j.src_img2, "%s: job %d dst_img2 without src_img2", who, n);
      FSG_REQUIRE(j.dst_img != j.src_img || !j.dst_img || (j.mode == 0 && !j.flip), "%s: job %d in-place warp is not allowed", who, n);
    }
  }
  return 0;
}

static size_t field_smem(const fsg_warp_job* jobs, int njobs) {
  int fz = 1;
  for (int n = 0; n < njobs; ++n)
    if (jobs[n].fsmall && jobs[n].fs[2] > fz) fz = jobs[n].fs[2];
  return (size_t)WX * WY * fz * sizeof(float4);
}

static dim3 warp_grid(int njobs, int sx, int sy) { return dim3(((sy + WY - 1) / WY) * ((sx + WX - 1) / WX), 1, njobs); }

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_warp_shift(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  if (int rc = validate(jobs, njobs, sx, sy, sz, false, "fsg_warp_shift")) return rc;
  for (int n = 0; n < njobs; ++n) FSG_REQUIRE(jobs[n].mode == 1, "fsg_warp_shift: job %d is not a deformation job", n);
  cudaStream_t s = as_stream(stream);
  shift_init_kernel<<<1, 64, 0, s>>>(b, njobs);
  const int edges = 4 * (sx + sy + sz);
  shift_edges_kernel<<<dim3((edges + 127) / 128, njobs), 128, 0, s>>>(b, sx, sy, sz);
  // full-volume pass on a bounded grid; blocks of jobs already resolved by the faces return at once
  const int ntiles = ((sy + WY - 1) / WY) * ((sx + WX - 1) / WX);
  const int gx = ntiles < 148 * 2 ? ntiles : 148 * 2;
  warp_kernel<PASS_SHIFT, false><<<dim3(gx, 1, njobs), WARP_THREADS, field_smem(jobs, njobs), s>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  shift_final_kernel<<<1, 64, 0, s>>>(b, njobs);
  return check_launch("fsg_warp_shift");
}

extern "C" int fsg_warp(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  if (int rc = validate(jobs, njobs, sx, sy, sz, true, "fsg_warp")) return rc;
  bool img2 = false;
  for (int n = 0; n < njobs; ++n) img2 = img2 || jobs[n].dst_img2 != nullptr;
  if (img2)
    warp_kernel<PASS_WARP, true><<<warp_grid(njobs, sx, sy), WARP_THREADS, field_smem(jobs, njobs), as_stream(stream)>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  else
    warp_kernel<PASS_WARP, false><<<warp_grid(njobs, sx, sy), WARP_THREADS, field_smem(jobs, njobs), as_stream(stream)>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  return check_launch("fsg_warp");
}

extern "C" int fsg_warp_coords(const fsg_warp_job* job, int sx, int sy, int sz, float* xx, float* yy, float* zz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, job, 1)) return rc;
  if (int rc = validate(job, 1, sx, sy, sz, false, "fsg_warp_coords")) return rc;
  FSG_REQUIRE(xx && yy && zz, "fsg_warp_coords: NULL output");
  warp_kernel<PASS_COORDS, false><<<warp_grid(1, sx, sy), WARP_THREADS, field_smem(job, 1), as_stream(stream)>>>(b, sx, sy, sz, xx, yy, zz);
  return check_launch("fsg_warp_coords");
}
