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
#include <stdlib.h>

#include "warp_common.cuh"

namespace fsg {

// warp_tile.cu: TMA-staged variant; returns 0 when it launched, -1 when the batch is not eligible
int launch_warp_tile(const fsg_warp_job* jobs, int njobs, bool epi, int sx, int sy, int sz, cudaStream_t stream);
// warp_pipe.cu: pipelined TMA variant (persistent blocks, producer / consumer warps); same return convention
int launch_warp_pipe(const fsg_warp_job* jobs, int njobs, bool epi, int sx, int sy, int sz, cudaStream_t stream);

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

// ---------------------------------------------------------------------------------- fast path
// Same arithmetic as warp_kernel<PASS_WARP> for the production case (deformation with a control
// grid, image + segmentation, no second image, extents that are multiples of the tile), with the
// per-voxel instruction count cut from ~200 to ~130 (r01d ncu: the generic kernel is issue-bound,
// 72 % issue-slot utilisation at 15 % of DRAM bandwidth, 32 IMAD + 20 IADD3 + 9 MOV + 8 BRA per
// voxel of pure overhead):
//   * no per-voxel runtime flags: every job of the launch has the same feature set (the host
//     partitions a batch) and EPI is a template parameter; the crop shift is subtracted
//     unconditionally (x - 0 is exact);
//   * ONE gather base index per voxel.  The floor indices are clamped to S-2 in the float domain
//     (min with the largest float below S-1 before the magic add), so the +1 neighbours are always
//     in bounds and become fixed pointer offsets (+1 element, +sz, +/-plane); where the reference
//     clamps the ceil index instead (coordinate == S-1) the weight is exactly 1 and the lerp returns
//     the same corner value to within an ulp of the image path's tolerance;
//   * the 2^23 magic-number biases of the six float->int conversions fold into one per-thread
//     constant; shared-memory and output addresses advance by constant strides;
//   * gamma and bias share one exponential: v' = 2^(gamma*lg2 v + (1-gamma) lg2 300 + bf*lg2 e).
// Coordinates (and therefore the nearest-neighbour segmentation) stay bit-exact: the chain of
// separately rounded mul/add is unchanged.
// FSG_WARP_PIN: loop invariants kept in ordinary registers through an opaque asm (ptxas otherwise re-derives them
// from the constant bank inside the inner loop): bit 0 the affine (12 registers), bit 1 the strides / texture
// handle, bit 2 the clamp bounds.  r02, texture variant: inner loop of two voxels 238 -> 208 instructions with 6.
#ifndef FSG_WARP_PIN
#define FSG_WARP_PIN 6
#endif
#ifndef FSG_WARP_MINBLOCKS
#define FSG_WARP_MINBLOCKS 3  // <= 85 registers: 3 blocks of 256 threads per SM
#endif
// The texture variant is bound by texture latency and issue slots together: 4 resident blocks (64 registers, no
// spills with FSG_WARP_PIN=6) measured 0.655 ms against 0.70 ms with 3 (r02, 8 volumes of 256^3); the linear
// variants spill at 64 registers and stay at 3.
#ifndef FSG_WARP_MINBLOCKS_TEX
#define FSG_WARP_MINBLOCKS_TEX 4
#endif
// PAIRS: the source image is fsg_gmm's out_pairs volume — 16-bit fixed point (I[z] | I[z+1] << 16), so
// ONE 32-bit gather brings both z corners of an (x, y) row: four gather instructions per voxel instead
// of eight (the kernel is bound by L1 wavefronts per gather, not by bytes).
// PAIRS == 3: the source image is a block-linear fsg_texvol (job.src_tex): the 2x2 (y, z) footprint of each
// of the two x layers comes from ONE texture gather (tld4), so a voxel costs two texture instructions instead
// of eight global loads.  The gather returns the raw float32 texels: same values, same blend, same result.
__device__ __forceinline__ float4 tex_gather_yz(unsigned long long tex, int layer, float u, float v) {
  float4 r;  // (z0,y1), (z1,y1), (z1,y0), (z0,y0) for the footprint z0 = floor(u - 0.5), y0 = floor(v - 0.5)
  asm("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%8}];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(tex), "r"(layer), "f"(u), "f"(v), "f"(0.f));
  return r;
}

// Trilinear blend of one voxel from its two gathers (x layers l and l+1), in the order of the packed code below
// (x, then y, then z; every step a fused multiply-add), so the result is bit-identical to it.  The x blend runs
// as two packed FMAs on the register pairs as the texture unit returns them: (x, y) = row y1 at (z0, z1),
// (z, w) = row y0 at (z1, z0).
__device__ __forceinline__ float tex_blend(const float4 g0, const float4 g1, float wx, float wy, float wz) {
  const P2 w = pk(wx, wx), lo1 = pk(g0.x, g0.y), lo0 = pk(g0.z, g0.w);
  float c10, c11, c01, c00;
  upk(fma2(w, sub2(pk(g1.x, g1.y), lo1), lo1), c10, c11);
  upk(fma2(w, sub2(pk(g1.z, g1.w), lo0), lo0), c01, c00);
  const float c0_ = __fmaf_rn(wy, __fsub_rn(c10, c00), c00), c1_ = __fmaf_rn(wy, __fsub_rn(c11, c01), c01);
  return __fmaf_rn(wz, __fsub_rn(c1_, c0_), c0_);
}

template <bool EPI, int PAIRS>
__global__ void __launch_bounds__(WARP_THREADS, PAIRS == 3 ? FSG_WARP_MINBLOCKS_TEX : FSG_WARP_MINBLOCKS) warp_fast_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int sx, int sy, int sz) {
  const fsg_warp_job& job = batch.j[blockIdx.z];
  const int tid = threadIdx.x;
  extern __shared__ float4 s_dyn[];
  float4* s_f = s_dyn;                    // [WY*WX][fz_n]: row = ry*WX + rx
  __shared__ float s_b[WX * WY][MAX_BZ];  // bias rows, pre-scaled by log2(e)

  const int fz_n = job.fs[2];
  const bool has_bias = EPI && job.bf_low != nullptr;
  const int ntile_y = sy / WY;
  const int x0 = ((int)blockIdx.x / ntile_y) * WX, y0 = ((int)blockIdx.x % ntile_y) * WY;

  // ---- phase A: x/y blends of the control grids for the tile's rows (exact, as in the generic kernel)
  {
    const int fy_n = job.fs[1];
    const int per_row = fz_n * 3;
    float* sf = reinterpret_cast<float*>(s_f);
    for (int e = tid; e < WX * WY * per_row; e += WARP_THREADS) {
      const int row = e / per_row, rem = e - row * per_row;
      const int ry = row / WX, rx = row - ry * WX;
      const Tab tx = load_tab(job.ftab[0], x0 + rx), ty = load_tab(job.ftab[1], y0 + ry);
      const float* g = job.fsmall + rem;
      const int sxs = fy_n * per_row;
      const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * per_row), tx.wc, __ldg(g + tx.c * sxs + ty.f * per_row));
      const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * per_row), tx.wc, __ldg(g + tx.c * sxs + ty.c * per_row));
      const int zc = rem / 3, ch = rem - zc * 3;
      sf[(row * fz_n + zc) * 4 + ch] = blend(ty.wf, t1f, ty.wc, t1c);
    }
  }
  if (EPI) {
    const int bz_n = has_bias ? job.bs[2] : 1;
    for (int e = tid; e < WX * WY * bz_n; e += WARP_THREADS) {
      const int row = e / bz_n, zc = e - row * bz_n;
      float v = 0.f;
      if (has_bias) {
        const int ry = row / WX, rx = row - ry * WX;
        const int by_n = job.bs[1];
        const Tab tx = load_tab(job.btab[0], x0 + rx), ty = load_tab(job.btab[1], y0 + ry);
        const float* g = job.bf_low + zc;
        const int sxs = by_n * bz_n;
        const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.f * bz_n));
        const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.c * bz_n));
        v = blend(ty.wf, t1f, ty.wc, t1c) * 1.4426950408889634f;
      }
      s_b[row][zc] = v;
    }
  }
  __syncthreads();

  Affine aff(job);
#if FSG_WARP_PIN & 1
  // keep the affine in ordinary registers: ptxas otherwise re-reads the twelve values from the constant bank
  // (eleven LDCU + the block-index arithmetic) in every iteration of the inner loop
#pragma unroll
  for (int q = 0; q < 9; ++q) asm volatile("" : "+f"(aff.a[q]));
#pragma unroll
  for (int q = 0; q < 3; ++q) asm volatile("" : "+f"(aff.c[q]));
#endif
  const float cen_x = job.center[0], cen_y = job.center[1], cen_z = job.center[2];
  float mx = (float)(sx - 1), my = (float)(sy - 1), mz = (float)(sz - 1);
  // largest floats below S-1: floor(min(c, that)) <= S-2
  float lx = __int_as_float(__float_as_int(mx) - 1), ly = __int_as_float(__float_as_int(my) - 1), lz = __int_as_float(__float_as_int(mz) - 1);
  const int plane = sy * sz;
  int xs = job.flip ? -plane : plane;
#if FSG_WARP_PIN & 2
  asm volatile("" : "+r"(xs));
#endif
#if FSG_WARP_PIN & 4
  asm volatile("" : "+f"(mx), "+f"(my), "+f"(mz), "+f"(lx), "+f"(ly), "+f"(lz));
#endif
  // index = fx*xs + fy*sz + fz with all three taken as raw float bits of (value + 2^23)
  const unsigned kbias = (unsigned)(job.flip ? (sx - 1) * plane : 0) - 0x4B000000u * (unsigned)(xs + sz + 1);
  unsigned long long src_tex = job.src_tex;
  int lay0 = job.flip ? sx - 1 : 0, lstep = job.flip ? -1 : 1;  // texture layer of floor index l: lay0 + lstep * l
  int fzn = fz_n;
#if FSG_WARP_PIN & 2
  asm volatile("" : "+l"(src_tex));
  asm volatile("" : "+r"(lay0));
  asm volatile("" : "+r"(lstep));
  asm volatile("" : "+r"(fzn));
#endif
  const float* __restrict__ const src_img = PAIRS ? reinterpret_cast<const float*>(job.src_pairs) : job.src_img;
  const uint8_t* __restrict__ const src_seg = job.src_seg;
  float* __restrict__ const dst_img = job.dst_img;
  uint8_t* __restrict__ const dst_seg = job.dst_seg;
  const float shx = job.shift[0], shy = job.shift[1], shz = job.shift[2];
  const float gam = (EPI && job.has_gamma) ? job.gamma : 1.0f;
  // lg2(300) (1 - gamma); the fixed-point scale 2^-7 of the pairs format folds in as -7 gamma
  const float c0 = ((EPI && job.has_gamma) ? 8.22881869049588f * (1.0f - job.gamma) : 0.f) - (PAIRS == 1 ? 7.0f * gam : 0.f);

  const P2 cen2 = pk(cen_x, cen_x), sh2x = pk(shx, shx), sh2y = pk(shy, shy), sh2z = pk(shz, shz), magic2 = pk(MAGIC, MAGIC);
  const P2 c2x = pk(aff.c[0], aff.c[0]), c2y = pk(aff.c[1], aff.c[1]), c2z = pk(aff.c[2], aff.c[2]);
  const char* const img_b = reinterpret_cast<const char*>(src_img);
  constexpr int ESZ = PAIRS == 2 ? 8 : 4;  // bytes per source voxel (float2 pairs: 8)
  const ptrdiff_t by_row = (ptrdiff_t)sz * ESZ, by_plane = (ptrdiff_t)xs * ESZ;

  for (int k = tid; k < sz; k += WARP_THREADS) {
    const Tab tf = load_tab(job.ftab[2], k);
    Tab tb = {0, 0, 0.f, 1.f};
    if (has_bias) tb = load_tab(job.btab[2], k);
    const float zc = sub_rn((float)k, cen_z);
    const P2 zc2 = pk(zc, zc), bwf2 = pk(tb.wf, tb.wf), bwc2 = pk(tb.wc, tb.wc), gam2 = pk(gam, gam), c02 = pk(c0, c0);
    const float4* pf = s_f + tf.f;
    const float4* pc = s_f + tf.c;
    unsigned o_row = (unsigned)((x0 * sy + y0) * sz + k);
    int row = 0, roff = 0;  // roff = row * fzn
#pragma unroll 1
    for (int ry = 0; ry < WY; ++ry, o_row += sz) {
      const float yc = sub_rn((float)(y0 + ry), cen_y);
      const P2 yc2 = pk(yc, yc);
      P2 xi2 = pk((float)x0, (float)(x0 + 1));
      unsigned o = o_row;
#pragma unroll 1
      for (int rx = 0; rx < WX; rx += 2, row += 2, roff += 2 * fzn, o += 2 * plane, xi2 = add2(xi2, pk(2.0f, 2.0f))) {
        // two voxels (i, j, k) and (i+1, j, k): products stay scalar FMULs (separately rounded, ptxas
        // would contract a packed mul feeding a packed add), every add is one packed FADD2
        const float4 f0a = pf[roff], f1a = pc[roff], f0b = pf[roff + fzn], f1b = pc[roff + fzn];
        const P2 fx = add2(pk(mul_rn(tf.wf, f0a.x), mul_rn(tf.wf, f0b.x)), pk(mul_rn(tf.wc, f1a.x), mul_rn(tf.wc, f1b.x)));
        const P2 fy = add2(pk(mul_rn(tf.wf, f0a.y), mul_rn(tf.wf, f0b.y)), pk(mul_rn(tf.wc, f1a.y), mul_rn(tf.wc, f1b.y)));
        const P2 fz = add2(pk(mul_rn(tf.wf, f0a.z), mul_rn(tf.wf, f0b.z)), pk(mul_rn(tf.wc, f1a.z), mul_rn(tf.wc, f1b.z)));
        float x1a, x1b, y1a, y1b, z1a, z1b;
        upk(add2(sub2(xi2, cen2), fx), x1a, x1b);
        upk(add2(yc2, fy), y1a, y1b);
        upk(add2(zc2, fz), z1a, z1b);
        float iia, iib, jja, jjb, kka, kkb;
        upk(add2(add2(add2(pk(mul_rn(aff.a[0], x1a), mul_rn(aff.a[0], x1b)), pk(mul_rn(aff.a[1], y1a), mul_rn(aff.a[1], y1b))), pk(mul_rn(aff.a[2], z1a), mul_rn(aff.a[2], z1b))), c2x), iia, iib);
        upk(add2(add2(add2(pk(mul_rn(aff.a[3], x1a), mul_rn(aff.a[3], x1b)), pk(mul_rn(aff.a[4], y1a), mul_rn(aff.a[4], y1b))), pk(mul_rn(aff.a[5], z1a), mul_rn(aff.a[5], z1b))), c2y), jja, jjb);
        upk(add2(add2(add2(pk(mul_rn(aff.a[6], x1a), mul_rn(aff.a[6], x1b)), pk(mul_rn(aff.a[7], y1a), mul_rn(aff.a[7], y1b))), pk(mul_rn(aff.a[8], z1a), mul_rn(aff.a[8], z1b))), c2z), kka, kkb);
        const P2 ii = sub2(pk(fminf(fmaxf(iia, 0.f), mx), fminf(fmaxf(iib, 0.f), mx)), sh2x);
        const P2 jj = sub2(pk(fminf(fmaxf(jja, 0.f), my), fminf(fmaxf(jjb, 0.f), my)), sh2y);
        const P2 kk = sub2(pk(fminf(fmaxf(kka, 0.f), mz), fminf(fmaxf(kkb, 0.f), mz)), sh2z);
        upk(ii, iia, iib);
        upk(jj, jja, jjb);
        upk(kk, kka, kkb);
        // ---- floor (clamped to S-2) and weights
        const P2 tx2 = add2_rz(pk(fminf(iia, lx), fminf(iib, lx)), magic2), ty2 = add2_rz(pk(fminf(jja, ly), fminf(jjb, ly)), magic2), tz2 = add2_rz(pk(fminf(kka, lz), fminf(kkb, lz)), magic2);
        float txa, txb, tya, tyb, tza, tzb;
        upk(tx2, txa, txb);
        upk(ty2, tya, tyb);
        upk(tz2, tza, tzb);
        const unsigned ba = (unsigned)__float_as_int(txa) * (unsigned)xs + (unsigned)__float_as_int(tya) * (unsigned)sz + (unsigned)__float_as_int(tza) + kbias;
        const unsigned bb = (unsigned)__float_as_int(txb) * (unsigned)xs + (unsigned)__float_as_int(tyb) * (unsigned)sz + (unsigned)__float_as_int(tzb) + kbias;
        const P2 wx2 = sub2(ii, sub2(tx2, magic2)), wy2 = sub2(jj, sub2(ty2, magic2)), wz2 = sub2(kk, sub2(tz2, magic2));
        const char* a00 = img_b + (size_t)ba * ESZ;
        const char* a01 = a00 + by_row;
        const char* a10 = a00 + by_plane;
        const char* a11 = a10 + by_row;
        const char* b00 = img_b + (size_t)bb * ESZ;
        const char* b01 = b00 + by_row;
        const char* b10 = b00 + by_plane;
        const char* b11 = b10 + by_row;
        P2 c000, c001, c010, c011, c100, c101, c110, c111;
        float4 tex_a0, tex_a1, tex_b0, tex_b1;
        if (PAIRS == 3) {
          // floor indices back from the magic-biased floats: layer = x (mirrored when flipped), texel centre
          // of the footprint's far corner = floor + 1 (the gather takes floor(u - 0.5) and the next texel)
          const int la = __float_as_int(txa) - 0x4B000000, lb = __float_as_int(txb) - 0x4B000000;
          const int la0 = lay0 + lstep * la, lb0 = lay0 + lstep * lb;
          float ua, ub, va_, vb_;
          upk(add2(sub2(tz2, magic2), pk(1.0f, 1.0f)), ua, ub);
          upk(add2(sub2(ty2, magic2), pk(1.0f, 1.0f)), va_, vb_);
          const float4 ga0 = tex_gather_yz(src_tex, la0, ua, va_), ga1 = tex_gather_yz(src_tex, la0 + lstep, ua, va_);
          const float4 gb0 = tex_gather_yz(src_tex, lb0, ub, vb_), gb1 = tex_gather_yz(src_tex, lb0 + lstep, ub, vb_);
          // blended per voxel (below) on the register pairs the gathers return: no repacking across the two voxels
          tex_a0 = ga0, tex_a1 = ga1, tex_b0 = gb0, tex_b1 = gb1;
        } else if (PAIRS == 2) {  // float2 (I[z], I[z+1]): both z corners of a row in one 8-byte gather, exact
#define LD2(p) __ldg(reinterpret_cast<const float2*>(p))
          const float2 qa00 = LD2(a00), qa01 = LD2(a01), qa10 = LD2(a10), qa11 = LD2(a11);
          const float2 qb00 = LD2(b00), qb01 = LD2(b01), qb10 = LD2(b10), qb11 = LD2(b11);
#undef LD2
          c000 = pk(qa00.x, qb00.x);
          c001 = pk(qa00.y, qb00.y);
          c010 = pk(qa01.x, qb01.x);
          c011 = pk(qa01.y, qb01.y);
          c100 = pk(qa10.x, qb10.x);
          c101 = pk(qa10.y, qb10.y);
          c110 = pk(qa11.x, qb11.x);
          c111 = pk(qa11.y, qb11.y);
        } else if (PAIRS == 1) {
#define LDU(p) __ldg(reinterpret_cast<const unsigned*>(p))
          const unsigned ua00 = LDU(a00), ua01 = LDU(a01), ua10 = LDU(a10), ua11 = LDU(a11);
          const unsigned ub00 = LDU(b00), ub01 = LDU(b01), ub10 = LDU(b10), ub11 = LDU(b11);
#undef LDU
          // 16-bit field -> float: (0x4B000000 | field) is 2^23 + field exactly; one packed subtract per corner pair
#define LO(u) __uint_as_float(__byte_perm((u), 0x4B000000u, 0x7610))
#define HI(u) __uint_as_float(__byte_perm((u), 0x4B000000u, 0x7632))
          c000 = sub2(pk(LO(ua00), LO(ub00)), magic2);
          c001 = sub2(pk(HI(ua00), HI(ub00)), magic2);
          c010 = sub2(pk(LO(ua01), LO(ub01)), magic2);
          c011 = sub2(pk(HI(ua01), HI(ub01)), magic2);
          c100 = sub2(pk(LO(ua10), LO(ub10)), magic2);
          c101 = sub2(pk(HI(ua10), HI(ub10)), magic2);
          c110 = sub2(pk(LO(ua11), LO(ub11)), magic2);
          c111 = sub2(pk(HI(ua11), HI(ub11)), magic2);
#undef LO
#undef HI
        } else {
#define LDF(p, off) __ldg(reinterpret_cast<const float*>(p) + (off))
          c000 = pk(LDF(a00, 0), LDF(b00, 0));
          c001 = pk(LDF(a00, 1), LDF(b00, 1));
          c010 = pk(LDF(a01, 0), LDF(b01, 0));
          c011 = pk(LDF(a01, 1), LDF(b01, 1));
          c100 = pk(LDF(a10, 0), LDF(b10, 0));
          c101 = pk(LDF(a10, 1), LDF(b10, 1));
          c110 = pk(LDF(a11, 0), LDF(b11, 0));
          c111 = pk(LDF(a11, 1), LDF(b11, 1));
#undef LDF
        }
        // ---- nearest segmentation gather (round-half-even by the magic add)
        float sxa, sxb, sya, syb, sza, szb;
        upk(add2(ii, magic2), sxa, sxb);
        upk(add2(jj, magic2), sya, syb);
        upk(add2(kk, magic2), sza, szb);
        const uint8_t laba = __ldg(src_seg + ((unsigned)__float_as_int(sxa) * (unsigned)xs + (unsigned)__float_as_int(sya) * (unsigned)sz + (unsigned)__float_as_int(sza) + kbias));
        const uint8_t labb = __ldg(src_seg + ((unsigned)__float_as_int(sxb) * (unsigned)xs + (unsigned)__float_as_int(syb) * (unsigned)sz + (unsigned)__float_as_int(szb) + kbias));
        // ---- trilinear blend (image path: fused, packed)
        float va, vb;
        if (PAIRS == 3) {
          float wxa, wxb, wya, wyb, wza, wzb;
          upk(wx2, wxa, wxb);
          upk(wy2, wya, wyb);
          upk(wz2, wza, wzb);
          va = tex_blend(tex_a0, tex_a1, wxa, wya, wza);
          vb = tex_blend(tex_b0, tex_b1, wxb, wyb, wzb);
        } else {
          const P2 c00 = fma2(wx2, sub2(c100, c000), c000), c01 = fma2(wx2, sub2(c101, c001), c001);
          const P2 c10 = fma2(wx2, sub2(c110, c010), c010), c11 = fma2(wx2, sub2(c111, c011), c011);
          const P2 c0_ = fma2(wy2, sub2(c10, c00), c00), c1_ = fma2(wy2, sub2(c11, c01), c01);
          upk(fma2(wz2, sub2(c1_, c0_), c0_), va, vb);
        }
        va = fminf(fminf(iia, jja), kka) > 0.f ? va : 0.f;
        vb = fminf(fminf(iib, jjb), kkb) > 0.f ? vb : 0.f;
        if (PAIRS == 1 && !EPI) {
          va *= 0.0078125f;
          vb *= 0.0078125f;
        }
        if (EPI) {
          const P2 bias = fma2(bwf2, pk(s_b[row][tb.f], s_b[row + 1][tb.f]), mul2(bwc2, pk(s_b[row][tb.c], s_b[row + 1][tb.c])));
          float ea, eb;
          upk(add2(fma2(gam2, pk(lg2_approx(va), lg2_approx(vb)), c02), bias), ea, eb);
          va = ex2_approx(ea);
          vb = ex2_approx(eb);
        }
        dst_img[o] = va;
        dst_img[o + plane] = vb;
        dst_seg[o] = laba;
        dst_seg[o + plane] = labb;
      }
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
      FSG_REQUIRE(!j.dst_img || j.src_img || j.src_pairs || j.src_tex, "%s: job %d dst_img without src_img", who, n);
      FSG_REQUIRE(!j.src_tex || (!j.src_img && !j.src_pairs), "%s: job %d has both a texture and a linear source image", who, n);
      FSG_REQUIRE(!j.dst_seg || j.src_seg, "%s: job %d dst_seg without src_seg", who, n);
      FSG_REQUIRE(!j.dst_img2 || j.src_img2, "%s: job %d dst_img2 without src_img2", who, n);
      FSG_REQUIRE(j.dst_img != j.src_img || !j.dst_img || (j.mode == 0 && !j.flip), "%s: job %d in-place warp is not allowed", who, n);
      FSG_REQUIRE(!j.src_pairs || static_cast<const void*>(j.src_pairs) != static_cast<const void*>(j.dst_img), "%s: job %d in-place warp is not allowed", who, n);
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

// A job takes the fast kernel when it is the production case; `epi` = it has a gamma or bias epilogue.
static bool fast_eligible(const fsg_warp_job& j, int sx, int sy, int sz) {
  return j.mode == 1 && j.fsmall && (j.src_img || j.src_pairs || j.src_tex) && j.dst_img && j.src_seg && j.dst_seg && !j.dst_img2 && sx % WX == 0 && sy % WY == 0 && sx >= 2 && sy >= 2 &&
         sz >= 2 && (!j.bf_low || j.dst_img);
}

extern "C" int fsg_warp(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  FSG_REQUIRE(jobs != nullptr && njobs >= 1 && njobs <= FSG_MAX_JOBS, "fsg_warp: njobs=%d outside [1,%d]", njobs, FSG_MAX_JOBS);
  if (int rc = validate(jobs, njobs, sx, sy, sz, true, "fsg_warp")) return rc;
  // partition the batch: fast kernel with / without epilogue, generic kernel for everything else
  // groups: 0/1 = fast kernel on a float source with / without epilogue, 2 = generic, 3/4 = fast kernel on
  // the fixed-point pairs source with / without epilogue
  fsg_warp_job part[9][FSG_MAX_JOBS];  // 5/6: fast kernel on the float2 pairs source, 7/8: on the block-linear texture source
  int cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = 0; n < njobs; ++n) {
    const fsg_warp_job& j = jobs[n];
    const bool fast = fast_eligible(j, sx, sy, sz);
    FSG_REQUIRE(!j.src_pairs || (fast && (reinterpret_cast<uintptr_t>(j.src_pairs) & (j.pairs_float ? 7 : 3)) == 0), "fsg_warp: job %d: src_pairs is only read by the fast path", n);
    FSG_REQUIRE(!j.src_tex || fast, "fsg_warp: job %d: a texture source (src_tex) is only read by the fast path (deformation with a control grid, tile-aligned extents)", n);
    const int g = !fast ? 2 : ((j.src_tex ? 7 : (j.src_pairs ? (j.pairs_float ? 5 : 3) : 0)) + ((j.has_gamma || j.bf_low) ? 0 : 1));
    part[g][cnt[g]++] = j;
  }
  cudaStream_t s = as_stream(stream);
  Batch<fsg_warp_job> b;
  for (int g = 0; g < 9; ++g) {
    if (!cnt[g] || g == 2) continue;
    // FSG_WARP_TILE=1 selects the TMA-staged cubic-tile variant (warp_tile.cu).  It is parity-green
    // but measured 2.4x slower than the full-z kernel at 256^3 (r01e: 2.17 ms vs 0.90 ms per 8
    // volumes; one 163 KB block per SM, no overlap of the box load with the gathers), so it is
    // opt-in until it is double-buffered.
    if (config().warp_pipe && g < 2) {
      const int rc = launch_warp_pipe(part[g], cnt[g], g == 0, sx, sy, sz, s);
      if (rc == 0) continue;
      if (rc > 0) return rc;
    }
    if (config().warp_tile && g < 2) {
      const int rc = launch_warp_tile(part[g], cnt[g], g == 0, sx, sy, sz, s);
      if (rc == 0) continue;
      if (rc > 0) return rc;
    }
    if (int rc = fill_batch(b, part[g], cnt[g])) return rc;
    const dim3 grid((sy / WY) * (sx / WX), 1, cnt[g]);
    const size_t smem = field_smem(part[g], cnt[g]);
    if (g == 0)
      warp_fast_kernel<true, 0><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 1)
      warp_fast_kernel<false, 0><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 3)
      warp_fast_kernel<true, 1><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 4)
      warp_fast_kernel<false, 1><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 5)
      warp_fast_kernel<true, 2><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 6)
      warp_fast_kernel<false, 2><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else if (g == 7)
      warp_fast_kernel<true, 3><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
    else
      warp_fast_kernel<false, 3><<<grid, WARP_THREADS, smem, s>>>(b, sx, sy, sz);
  }
  if (cnt[2]) {
    if (int rc = fill_batch(b, part[2], cnt[2])) return rc;
    bool img2 = false;
    for (int n = 0; n < cnt[2]; ++n) img2 = img2 || part[2][n].dst_img2 != nullptr;
    if (img2)
      warp_kernel<PASS_WARP, true><<<warp_grid(cnt[2], sx, sy), WARP_THREADS, field_smem(part[2], cnt[2]), s>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
    else
      warp_kernel<PASS_WARP, false><<<warp_grid(cnt[2], sx, sy), WARP_THREADS, field_smem(part[2], cnt[2]), s>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  }
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
