// K2 — fused spatial deformation (+ gamma / bias-field epilogues).
//
// One launch does what the reference spreads over ~60 torch kernels and 1.5 GB of
// transients (generator/deformation/affine_nonrigid.py:64-84, 164-193, 299-366 and
// utils/generation.py:204-288, 310-397):
//   F      = separable linear up-sampling of the control grid Fsmall (x, then y, then z;
//            every w_f*a + w_c*b rounded exactly like myzoom_torch's three loops)
//   coord  = ((A_r0*(xc+Fx) + A_r1*(yc+Fy)) + A_r2*(zc+Fz)) + c2_r, clamp [0,S-1], -shift
//   image  = trilinear gather (blend x, y, z; 0 where any coord <= 0 or > S-1)
//   seg    = nearest gather (rint = half-to-even, clamp)
//   flip   = sources mirrored along x before sampling
//   image  = 300*(image/300)^gamma ; image *= exp(zoom(bf_low))       (optional epilogues)
// No coordinate / field volume ever exists in HBM; the control grids are staged per tile in
// shared memory.  Algorithmic HBM bytes: read img 4 + seg 1, write img 4 + seg 1 = 10 B/voxel.
#include "common.cuh"

namespace fsg {

constexpr int WT_X = 8, WT_Y = 8, WT_Z = 32;
constexpr int WARP_THREADS = WT_Y * WT_Z;
constexpr int MAX_FZ = 32;  // control-grid extent along z kept in smem (reference: <= 0.06*S)
constexpr int MAX_BZ = 16;  // bias-grid extent along z (reference: <= 0.02*S)

enum { PASS_SHIFT = 0, PASS_WARP = 1, PASS_COORDS = 2 };

struct Coord3 {
  float x, y, z;
};

// Trilinear sample of `src` at (ii,jj,kk); arithmetic order of utils/generation.py:227-285.
__device__ __forceinline__ float sample_linear(const float* __restrict__ src, int sx, int sy, int sz, bool flip, float ii, float jj, float kk) {
  const bool ok = (ii > 0.f) && (jj > 0.f) && (kk > 0.f) && (ii <= (float)(sx - 1)) && (jj <= (float)(sy - 1)) && (kk <= (float)(sz - 1));
  if (!ok) return 0.f;
  const float ffx = floorf(ii), ffy = floorf(jj), ffz = floorf(kk);
  int fx = (int)ffx, fy = (int)ffy, fz = (int)ffz;
  int cx = min(fx + 1, sx - 1);
  const int cy = min(fy + 1, sy - 1), cz = min(fz + 1, sz - 1);
  const float wcx = sub_rn(ii, ffx), wcy = sub_rn(jj, ffy), wcz = sub_rn(kk, ffz);
  const float wfx = sub_rn(1.f, wcx), wfy = sub_rn(1.f, wcy), wfz = sub_rn(1.f, wcz);
  if (flip) {
    fx = sx - 1 - fx;
    cx = sx - 1 - cx;
  }
  const size_t rf = (size_t)fx * sy, rc = (size_t)cx * sy;
  const float* pff = src + (rf + fy) * sz;
  const float* pcf = src + (rc + fy) * sz;
  const float* pfc = src + (rf + cy) * sz;
  const float* pcc = src + (rc + cy) * sz;
  const float c000 = __ldg(pff + fz), c001 = __ldg(pff + cz);
  const float c100 = __ldg(pcf + fz), c101 = __ldg(pcf + cz);
  const float c010 = __ldg(pfc + fz), c011 = __ldg(pfc + cz);
  const float c110 = __ldg(pcc + fz), c111 = __ldg(pcc + cz);
  const float c00 = lerp2(c000, wfx, c100, wcx);
  const float c01 = lerp2(c001, wfx, c101, wcx);
  const float c10 = lerp2(c010, wfx, c110, wcx);
  const float c11 = lerp2(c011, wfx, c111, wcx);
  const float c0 = lerp2(c00, wfy, c10, wcy);
  const float c1 = lerp2(c01, wfy, c11, wcy);
  return lerp2(c0, wfz, c1, wcz);
}

__device__ __forceinline__ uint8_t sample_nearest(const uint8_t* __restrict__ src, int sx, int sy, int sz, bool flip, float ii, float jj, float kk) {
  int ir = min(max((int)rintf(ii), 0), sx - 1);
  const int jr = min(max((int)rintf(jj), 0), sy - 1);
  const int kr = min(max((int)rintf(kk), 0), sz - 1);
  if (flip) ir = sx - 1 - ir;
  return __ldg(src + ((size_t)ir * sy + jr) * sz + kr);
}

template <int PASS>
__global__ void __launch_bounds__(WARP_THREADS) warp_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, int sx, int sy, int sz, float* __restrict__ dbg_x,
                                                             float* __restrict__ dbg_y, float* __restrict__ dbg_z) {
  const int ntx = (sx + WT_X - 1) / WT_X;
  const fsg_warp_job& job = batch.j[blockIdx.z / ntx];
  const int x0 = (blockIdx.z % ntx) * WT_X, y0 = blockIdx.y * WT_Y, z0 = blockIdx.x * WT_Z;
  const int tid = threadIdx.y * WT_Z + threadIdx.x;

  __shared__ float s_f[WT_X][WT_Y][MAX_FZ][3];  // control grid blended along x and y
  __shared__ float s_b[WT_X][WT_Y][MAX_BZ];     // bias grid blended along x and y
  __shared__ float s_red[3][WARP_THREADS / 32];

  const bool deform = job.mode == 1;
  const bool has_field = deform && job.fsmall != nullptr;
  const bool has_bias = (PASS == PASS_WARP) && job.bf_low != nullptr && job.dst_img != nullptr;

  // ---- phase A: x- and y-blends of the low-resolution grids for the tile's 64 (x,y) rows
  if (has_field) {
    const int fy_n = job.fs[1], fz_n = job.fs[2];
    const int per_row = fz_n * 3;
    for (int e = tid; e < WT_X * WT_Y * per_row; e += WARP_THREADS) {
      const int row = e / per_row, rem = e - row * per_row;
      const int rx = row / WT_Y, ry = row - rx * WT_Y;
      const int i = min(x0 + rx, sx - 1), j = min(y0 + ry, sy - 1);
      const Tab tx = load_tab(job.ftab[0], i), ty = load_tab(job.ftab[1], j);
      const float* g = job.fsmall + rem;  // rem = zc*3 + ch
      const size_t sxs = (size_t)fy_n * per_row;
      const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + (size_t)ty.f * per_row), tx.wc, __ldg(g + tx.c * sxs + (size_t)ty.f * per_row));
      const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + (size_t)ty.c * per_row), tx.wc, __ldg(g + tx.c * sxs + (size_t)ty.c * per_row));
      (&s_f[rx][ry][0][0])[rem] = blend(ty.wf, t1f, ty.wc, t1c);
    }
  }
  if (has_bias) {
    const int by_n = job.bs[1], bz_n = job.bs[2];
    for (int e = tid; e < WT_X * WT_Y * bz_n; e += WARP_THREADS) {
      const int row = e / bz_n, zc = e - row * bz_n;
      const int rx = row / WT_Y, ry = row - rx * WT_Y;
      const int i = min(x0 + rx, sx - 1), j = min(y0 + ry, sy - 1);
      const Tab tx = load_tab(job.btab[0], i), ty = load_tab(job.btab[1], j);
      const float* g = job.bf_low + zc;
      const size_t sxs = (size_t)by_n * bz_n;
      const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + (size_t)ty.f * bz_n), tx.wc, __ldg(g + tx.c * sxs + (size_t)ty.f * bz_n));
      const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + (size_t)ty.c * bz_n), tx.wc, __ldg(g + tx.c * sxs + (size_t)ty.c * bz_n));
      s_b[rx][ry][zc] = blend(ty.wf, t1f, ty.wc, t1c);
    }
  }
  __syncthreads();

  // ---- phase B: one thread per (y,z) column of the tile, marching over x
  const int k = z0 + threadIdx.x, j = y0 + threadIdx.y;
  const bool in_yz = (k < sz) && (j < sy);
  const int kc = min(k, sz - 1), jc = min(j, sy - 1);
  Tab tfz = {0, 0, 0.f, 1.f}, tbz = {0, 0, 0.f, 1.f};
  if (has_field) tfz = load_tab(job.ftab[2], kc);
  if (has_bias) tbz = load_tab(job.btab[2], kc);
  const float yc = sub_rn((float)jc, job.center[1]);
  const float zc = sub_rn((float)kc, job.center[2]);
  const float shx = (PASS != PASS_SHIFT && deform) ? job.shift[0] : 0.f;
  const float shy = (PASS != PASS_SHIFT && deform) ? job.shift[1] : 0.f;
  const float shz = (PASS != PASS_SHIFT && deform) ? job.shift[2] : 0.f;
  const bool flip = job.flip != 0;
  const float inf = __int_as_float(0x7f800000);
  float mnx = inf, mny = inf, mnz = inf;

#pragma unroll 2
  for (int rx = 0; rx < WT_X; ++rx) {
    const int i = x0 + rx;
    if (i >= sx) break;
    float ii, jj, kk;
    if (deform) {
      float x1 = sub_rn((float)i, job.center[0]), y1 = yc, z1 = zc;
      if (has_field) {
        const float* f0 = &s_f[rx][threadIdx.y][tfz.f][0];
        const float* f1 = &s_f[rx][threadIdx.y][tfz.c][0];
        x1 = add_rn(x1, blend(tfz.wf, f0[0], tfz.wc, f1[0]));
        y1 = add_rn(y1, blend(tfz.wf, f0[1], tfz.wc, f1[1]));
        z1 = add_rn(z1, blend(tfz.wf, f0[2], tfz.wc, f1[2]));
      }
      ii = add_rn(add_rn(add_rn(mul_rn(job.A[0], x1), mul_rn(job.A[1], y1)), mul_rn(job.A[2], z1)), job.c2[0]);
      jj = add_rn(add_rn(add_rn(mul_rn(job.A[3], x1), mul_rn(job.A[4], y1)), mul_rn(job.A[5], z1)), job.c2[1]);
      kk = add_rn(add_rn(add_rn(mul_rn(job.A[6], x1), mul_rn(job.A[7], y1)), mul_rn(job.A[8], z1)), job.c2[2]);
      ii = ii < 0.f ? 0.f : ii;
      jj = jj < 0.f ? 0.f : jj;
      kk = kk < 0.f ? 0.f : kk;
      ii = ii > (float)(sx - 1) ? (float)(sx - 1) : ii;
      jj = jj > (float)(sy - 1) ? (float)(sy - 1) : jj;
      kk = kk > (float)(sz - 1) ? (float)(sz - 1) : kk;
      if (PASS == PASS_SHIFT) {
        if (in_yz) {
          mnx = fminf(mnx, ii);
          mny = fminf(mny, jj);
          mnz = fminf(mnz, kk);
        }
        continue;
      }
      ii = sub_rn(ii, shx);
      jj = sub_rn(jj, shy);
      kk = sub_rn(kk, shz);
    } else {
      ii = (float)i;
      jj = (float)jc;
      kk = (float)kc;
    }
    if (!in_yz) continue;
    const size_t o = ((size_t)i * sy + j) * sz + k;
    if (PASS == PASS_COORDS) {
      dbg_x[o] = ii;
      dbg_y[o] = jj;
      dbg_z[o] = kk;
      continue;
    }
    if (PASS == PASS_WARP) {
      if (job.dst_img) {
        float v;
        if (deform)
          v = sample_linear(job.src_img, sx, sy, sz, flip, ii, jj, kk);
        else
          v = __ldg(job.src_img + ((size_t)(flip ? sx - 1 - i : i) * sy + j) * sz + k);
        if (job.has_gamma) v = mul_rn(300.0f, powf(__fdiv_rn(v, 300.0f), job.gamma));
        if (has_bias) v = mul_rn(v, expf(blend(tbz.wf, s_b[rx][threadIdx.y][tbz.f], tbz.wc, s_b[rx][threadIdx.y][tbz.c])));
        job.dst_img[o] = v;
      }
      if (job.dst_seg) {
        uint8_t l;
        if (deform)
          l = sample_nearest(job.src_seg, sx, sy, sz, flip, ii, jj, kk);
        else
          l = __ldg(job.src_seg + ((size_t)(flip ? sx - 1 - i : i) * sy + j) * sz + k);
        job.dst_seg[o] = l;
      }
      if (job.dst_img2) {
        float v;
        if (deform)
          v = sample_linear(job.src_img2, sx, sy, sz, flip, ii, jj, kk);
        else
          v = __ldg(job.src_img2 + ((size_t)(flip ? sx - 1 - i : i) * sy + j) * sz + k);
        job.dst_img2[o] = v;
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
      // clamped coordinates are >= 0, so the int view of the float orders correctly
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

static dim3 warp_grid(int njobs, int sx, int sy, int sz) {
  return dim3((sz + WT_Z - 1) / WT_Z, (sy + WT_Y - 1) / WT_Y, ((sx + WT_X - 1) / WT_X) * njobs);
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_warp_shift(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  if (int rc = validate(jobs, njobs, sx, sy, sz, false, "fsg_warp_shift")) return rc;
  for (int n = 0; n < njobs; ++n) FSG_REQUIRE(jobs[n].mode == 1, "fsg_warp_shift: job %d is not a deformation job", n);
  cudaStream_t s = as_stream(stream);
  shift_init_kernel<<<1, 64, 0, s>>>(b, njobs);
  warp_kernel<PASS_SHIFT><<<warp_grid(njobs, sx, sy, sz), dim3(WT_Z, WT_Y), 0, s>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  shift_final_kernel<<<1, 64, 0, s>>>(b, njobs);
  return check_launch("fsg_warp_shift");
}

extern "C" int fsg_warp(const fsg_warp_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  if (int rc = validate(jobs, njobs, sx, sy, sz, true, "fsg_warp")) return rc;
  warp_kernel<PASS_WARP><<<warp_grid(njobs, sx, sy, sz), dim3(WT_Z, WT_Y), 0, as_stream(stream)>>>(b, sx, sy, sz, nullptr, nullptr, nullptr);
  return check_launch("fsg_warp");
}

extern "C" int fsg_warp_coords(const fsg_warp_job* job, int sx, int sy, int sz, float* xx, float* yy, float* zz, void* stream) {
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, job, 1)) return rc;
  if (int rc = validate(job, 1, sx, sy, sz, false, "fsg_warp_coords")) return rc;
  FSG_REQUIRE(xx && yy && zz, "fsg_warp_coords: NULL output");
  warp_kernel<PASS_COORDS><<<warp_grid(1, sx, sy, sz), dim3(WT_Z, WT_Y), 0, as_stream(stream)>>>(b, sx, sy, sz, xx, yy, zz);
  return check_launch("fsg_warp_coords");
}
