// K2, TMA-staged variant of the fused deformation kernel.
//
// r01e ncu of the full-z fast kernel (profiles/): 94 % of the L1 data-pipe wavefront peak at
// 15 % of DRAM bandwidth — a warp's 32 lanes walk along output z, the rotated source line they
// sample crosses ~5.4 distinct 128-byte lines per gather instruction, and every voxel issues 8 of
// them.  Global gathers cannot get cheaper than that, so this variant takes the gathers off the
// global-memory path:
//   * a block owns a 16^3 output tile.  The image of that tile under the deformation is bounded
//     by interval arithmetic (affine map of the tile box + the exact min/max of the control-grid
//     nodes that reach the tile), giving an axis-aligned source box of ~26^3 voxels;
//   * one elected thread issues two TMA tiled loads (cp.async.bulk.tensor.3d, float image box and
//     uint8 label box) that land in shared memory through the async proxy without touching the
//     LSU pipe; out-of-volume parts of the box are zero-filled by the TMA unit;
//   * the eight trilinear corners and the nearest label are then shared-memory loads
//     (lanes along z -> consecutive banks).
// Coordinates are computed exactly as in the other kernels (bit-exact segmentation).  A tile whose
// source box exceeds the tensor-map box (strong local field gradients) falls back to global
// gathers inside the same kernel.
#include <cuda.h>
#include <stdlib.h>

#include "warp_common.cuh"

namespace fsg {

constexpr int TT = 16;  // tile edge
constexpr int TILE_THREADS = 256;

struct TileParams {
  CUtensorMap img[FSG_MAX_JOBS];
  CUtensorMap seg[FSG_MAX_JOBS];
  // ex, ey, nz = extents a tile may need; ez / ezs = z extents of the float / uint8 boxes.  The TMA unit
  // wants the global address of a box start 16-byte aligned, i.e. the innermost coordinate a multiple
  // of 4 floats / 16 labels: the box origins are rounded down and ez / ezs carry the slack.
  int box[FSG_MAX_JOBS][5];
  int fnodes, bnodes;        // control-grid z-nodes staged per (x,y) row
  int box_floats, box_bytes; // capacity reserved in shared memory
  int debug;                 // FSG_TILE_DEBUG: bit 0 = force the global-gather path, bit 1 = check box indices
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

struct TileShared {
  unsigned long long mbar;
  int origin[3];  // memory-space index of the needed region's start (x: plane index, after the flip)
  int fit;
  float red[TILE_THREADS / 32][6];
};

// Gather + blend + epilogue + store for the 16 voxels of one thread (x loop, two at a time).
// SMEM: corners come from the staged box (index strides ez / ey*ez), else from global memory.
template <bool EPI, bool SMEM>
__device__ __forceinline__ void tile_voxels(const fsg_warp_job& job, const Affine& aff, const float4* __restrict__ s_f, const float* __restrict__ s_b, int fnodes, int bnodes, int fzlo,
                                            int bzlo, const float* __restrict__ img_base, const uint8_t* __restrict__ seg_base, unsigned kimg, unsigned kseg, int xs_img, int ys_img,
                                            int xs_seg, int ys_seg, int x0, int y0, int z0, int sx, int sy, int sz, unsigned dbg_limit_img = 0,
                                            unsigned dbg_limit_seg = 0) {
  const int tid = threadIdx.x;
  const int yl = tid >> 4, zl = tid & 15;
  const int j = y0 + yl, k = z0 + zl;
  const float mx = (float)(sx - 1), my = (float)(sy - 1), mz = (float)(sz - 1);
  const float lx = __int_as_float(__float_as_int(mx) - 1), ly = __int_as_float(__float_as_int(my) - 1), lz = __int_as_float(__float_as_int(mz) - 1);
  const Tab tf = load_tab(job.ftab[2], k);
  const bool has_bias = EPI && job.bf_low != nullptr;
  Tab tb = {0, 0, 0.f, 1.f};
  if (has_bias) tb = load_tab(job.btab[2], k);
  const float gam = (EPI && job.has_gamma) ? job.gamma : 1.0f;
  const float c0 = (EPI && job.has_gamma) ? 8.22881869049588f * (1.0f - job.gamma) : 0.f;
  const float zc = sub_rn((float)k, job.center[2]), yc = sub_rn((float)j, job.center[1]);
  const P2 zc2 = pk(zc, zc), yc2 = pk(yc, yc), cen2 = pk(job.center[0], job.center[0]), magic2 = pk(MAGIC, MAGIC);
  const P2 sh2x = pk(job.shift[0], job.shift[0]), sh2y = pk(job.shift[1], job.shift[1]), sh2z = pk(job.shift[2], job.shift[2]);
  const P2 c2x = pk(aff.c[0], aff.c[0]), c2y = pk(aff.c[1], aff.c[1]), c2z = pk(aff.c[2], aff.c[2]);
  const P2 bwf2 = pk(tb.wf, tb.wf), bwc2 = pk(tb.wc, tb.wc), gam2 = pk(gam, gam), c02 = pk(c0, c0);
  // staged control-grid rows: s_f[(xl*TT + yl)*fnodes + node - fzlo]
  const float4* pf = s_f + yl * fnodes + (tf.f - fzlo);
  const float4* pc = s_f + yl * fnodes + (tf.c - fzlo);
  const float* bf = s_b + yl * bnodes + (tb.f - bzlo);
  const float* bc = s_b + yl * bnodes + (tb.c - bzlo);
  const int frow = TT * fnodes, brow = TT * bnodes;
  const int plane = sy * sz;
  unsigned o = (unsigned)((x0 * sy + j) * sz + k);
  float* __restrict__ const dst_img = job.dst_img;
  uint8_t* __restrict__ const dst_seg = job.dst_seg;
  P2 xi2 = pk((float)x0, (float)(x0 + 1));
  const char* const img_b = reinterpret_cast<const char*>(img_base);
  const int by_row = ys_img * 4, by_plane = xs_img * 4;
#pragma unroll 1
  for (int xl = 0; xl < TT; xl += 2, o += 2 * plane, xi2 = add2(xi2, pk(2.0f, 2.0f)), pf += 2 * frow, pc += 2 * frow, bf += 2 * brow, bc += 2 * brow) {
    const float4 f0a = pf[0], f1a = pc[0], f0b = pf[frow], f1b = pc[frow];
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
    const P2 tx2 = add2_rz(pk(fminf(iia, lx), fminf(iib, lx)), magic2), ty2 = add2_rz(pk(fminf(jja, ly), fminf(jjb, ly)), magic2), tz2 = add2_rz(pk(fminf(kka, lz), fminf(kkb, lz)), magic2);
    float txa, txb, tya, tyb, tza, tzb;
    upk(tx2, txa, txb);
    upk(ty2, tya, tyb);
    upk(tz2, tza, tzb);
    const unsigned ba = (unsigned)__float_as_int(txa) * (unsigned)xs_img + (unsigned)__float_as_int(tya) * (unsigned)ys_img + (unsigned)__float_as_int(tza) + kimg;
    const unsigned bb = (unsigned)__float_as_int(txb) * (unsigned)xs_img + (unsigned)__float_as_int(tyb) * (unsigned)ys_img + (unsigned)__float_as_int(tzb) + kimg;
    const P2 wx2 = sub2(ii, sub2(tx2, magic2)), wy2 = sub2(jj, sub2(ty2, magic2)), wz2 = sub2(kk, sub2(tz2, magic2));
    float sxa, sxb, sya, syb, sza, szb;
    upk(add2(ii, magic2), sxa, sxb);
    upk(add2(jj, magic2), sya, syb);
    upk(add2(kk, magic2), sza, szb);
    const unsigned sa = (unsigned)__float_as_int(sxa) * (unsigned)xs_seg + (unsigned)__float_as_int(sya) * (unsigned)ys_seg + (unsigned)__float_as_int(sza) + kseg;
    const unsigned sb = (unsigned)__float_as_int(sxb) * (unsigned)xs_seg + (unsigned)__float_as_int(syb) * (unsigned)ys_seg + (unsigned)__float_as_int(szb) + kseg;
    P2 c000, c001, c010, c011, c100, c101, c110, c111;
    uint8_t laba, labb;
    if (SMEM) {
      if (dbg_limit_img) {
        const unsigned hi_i = (unsigned)(abs(xs_img) + ys_img + 1);
        if (ba + (xs_img < 0 ? xs_img : 0) >= dbg_limit_img || ba + hi_i - (xs_img < 0 ? -xs_img : 0) >= dbg_limit_img + hi_i || sa >= dbg_limit_seg || bb >= dbg_limit_img || sb >= dbg_limit_seg) {
          printf("tile (%d,%d,%d) thr %d xl %d: ba=%u bb=%u sa=%u sb=%u lim=%u/%u ii=(%f,%f,%f)\n", x0, y0, z0, tid, xl, ba, bb, sa, sb, dbg_limit_img, dbg_limit_seg, iia, jja, kka);
          continue;
        }
      }
      const float* a00 = img_base + ba;
      const float* b00 = img_base + bb;
      c000 = pk(a00[0], b00[0]);
      c001 = pk(a00[1], b00[1]);
      c010 = pk(a00[ys_img], b00[ys_img]);
      c011 = pk(a00[ys_img + 1], b00[ys_img + 1]);
      c100 = pk(a00[xs_img], b00[xs_img]);
      c101 = pk(a00[xs_img + 1], b00[xs_img + 1]);
      c110 = pk(a00[xs_img + ys_img], b00[xs_img + ys_img]);
      c111 = pk(a00[xs_img + ys_img + 1], b00[xs_img + ys_img + 1]);
      laba = seg_base[sa];
      labb = seg_base[sb];
    } else {
      const char* a00 = img_b + (size_t)ba * 4;
      const char* a01 = a00 + by_row;
      const char* a10 = a00 + by_plane;
      const char* a11 = a10 + by_row;
      const char* b00 = img_b + (size_t)bb * 4;
      const char* b01 = b00 + by_row;
      const char* b10 = b00 + by_plane;
      const char* b11 = b10 + by_row;
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
      laba = __ldg(seg_base + sa);
      labb = __ldg(seg_base + sb);
    }
    const P2 c00 = fma2(wx2, sub2(c100, c000), c000), c01 = fma2(wx2, sub2(c101, c001), c001);
    const P2 c10 = fma2(wx2, sub2(c110, c010), c010), c11 = fma2(wx2, sub2(c111, c011), c011);
    const P2 c0_ = fma2(wy2, sub2(c10, c00), c00), c1_ = fma2(wy2, sub2(c11, c01), c01);
    float va, vb;
    upk(fma2(wz2, sub2(c1_, c0_), c0_), va, vb);
    va = fminf(fminf(iia, jja), kka) > 0.f ? va : 0.f;
    vb = fminf(fminf(iib, jjb), kkb) > 0.f ? vb : 0.f;
    if (EPI) {
      const P2 bias = fma2(bwf2, pk(bf[0], bf[brow]), mul2(bwc2, pk(bc[0], bc[brow])));
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

template <bool EPI>
__global__ void __launch_bounds__(TILE_THREADS, 2) warp_tile_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, const __grid_constant__ TileParams tp, int sx, int sy, int sz) {
  const int jb = blockIdx.z;
  const fsg_warp_job& job = batch.j[jb];
  const int tid = threadIdx.x;
  extern __shared__ __align__(128) unsigned char s_raw[];
  // layout: [image box | label box | control rows (float4) | bias rows | TileShared]
  float* s_img = reinterpret_cast<float*>(s_raw);
  uint8_t* s_seg = s_raw + (size_t)tp.box_floats * 4;
  float4* s_f = reinterpret_cast<float4*>(s_seg + tp.box_bytes);
  float* s_b = reinterpret_cast<float*>(s_f + TT * TT * tp.fnodes);
  TileShared* sh = reinterpret_cast<TileShared*>(s_b + TT * TT * tp.bnodes);

  const int ntz = sz / TT, nty = sy / TT;
  const int tile = blockIdx.x;
  const int z0 = (tile % ntz) * TT, y0 = ((tile / ntz) % nty) * TT, x0 = (tile / (ntz * nty)) * TT;
  const bool has_bias = EPI && job.bf_low != nullptr;

  if (tid == 0) {
    mbar_init(smem_u32(&sh->mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- phase A: x/y blends of the control grid for the z-nodes that reach this tile (exact)
  const int fzlo = job.ftab[2][z0].f;
  const int fzn = min((int)job.ftab[2][z0 + TT - 1].c - fzlo + 1, tp.fnodes);
  const float inf = __int_as_float(0x7f800000);
  float lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
  {
    const int fy_n = job.fs[1], fz_n = job.fs[2];
    const int per = fzn * 3;
    float* sf = reinterpret_cast<float*>(s_f);
    for (int e = tid; e < TT * TT * per; e += TILE_THREADS) {
      const int row = e / per, rem = e - row * per;
      const int xl = row / TT, yl = row - xl * TT;
      const int zn = rem / 3, ch = rem - zn * 3;
      const Tab tx = load_tab(job.ftab[0], x0 + xl), ty = load_tab(job.ftab[1], y0 + yl);
      const float* g = job.fsmall + (fzlo + zn) * 3 + ch;
      const int sys = fz_n * 3, sxs = fy_n * sys;
      const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * sys), tx.wc, __ldg(g + tx.c * sxs + ty.f * sys));
      const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * sys), tx.wc, __ldg(g + tx.c * sxs + ty.c * sys));
      const float v = blend(ty.wf, t1f, ty.wc, t1c);
      sf[(row * tp.fnodes + zn) * 4 + ch] = v;
      if (ch == 0) {
        lo0 = fminf(lo0, v);
        hi0 = fmaxf(hi0, v);
      } else if (ch == 1) {
        lo1 = fminf(lo1, v);
        hi1 = fmaxf(hi1, v);
      } else {
        lo2 = fminf(lo2, v);
        hi2 = fmaxf(hi2, v);
      }
    }
  }
  int bzlo = 0;
  if (EPI) {
    int bzn = 1;
    if (has_bias) {
      bzlo = job.btab[2][z0].f;
      bzn = min((int)job.btab[2][z0 + TT - 1].c - bzlo + 1, tp.bnodes);
    }
    for (int e = tid; e < TT * TT * bzn; e += TILE_THREADS) {
      const int row = e / bzn, zn = e - row * bzn;
      float v = 0.f;
      if (has_bias) {
        const int xl = row / TT, yl = row - xl * TT;
        const int by_n = job.bs[1], bz_n = job.bs[2];
        const Tab tx = load_tab(job.btab[0], x0 + xl), ty = load_tab(job.btab[1], y0 + yl);
        const float* g = job.bf_low + bzlo + zn;
        const int sxs = by_n * bz_n;
        const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.f * bz_n));
        const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.c * bz_n));
        v = blend(ty.wf, t1f, ty.wc, t1c) * 1.4426950408889634f;
      }
      s_b[row * tp.bnodes + zn] = v;
    }
  }
  // ---- block min/max of the staged displacements (the z-blend is a convex combination of them)
  lo0 = warp_min(lo0);
  lo1 = warp_min(lo1);
  lo2 = warp_min(lo2);
  hi0 = warp_max(hi0);
  hi1 = warp_max(hi1);
  hi2 = warp_max(hi2);
  if ((tid & 31) == 0) {
    float* r = sh->red[tid >> 5];
    r[0] = lo0;
    r[1] = lo1;
    r[2] = lo2;
    r[3] = hi0;
    r[4] = hi1;
    r[5] = hi2;
  }
  __syncthreads();

  const Affine aff(job);
  const int ex = tp.box[jb][0], ey = tp.box[jb][1], nz = tp.box[jb][2], ez = tp.box[jb][3], ezs = tp.box[jb][4];
  if (tid == 0) {
    float flo[3], fhi[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      flo[c] = sh->red[0][c];
      fhi[c] = sh->red[0][3 + c];
      for (int w = 1; w < TILE_THREADS / 32; ++w) {
        flo[c] = fminf(flo[c], sh->red[w][c]);
        fhi[c] = fmaxf(fhi[c], sh->red[w][3 + c]);
      }
    }
    // interval of (voxel - center + F) per axis, then of each affine row; margins absorb rounding
    const float plo[3] = {(float)x0 - job.center[0] + flo[0] - 1e-3f, (float)y0 - job.center[1] + flo[1] - 1e-3f, (float)z0 - job.center[2] + flo[2] - 1e-3f};
    const float phi[3] = {(float)(x0 + TT - 1) - job.center[0] + fhi[0] + 1e-3f, (float)(y0 + TT - 1) - job.center[1] + fhi[1] + 1e-3f,
                          (float)(z0 + TT - 1) - job.center[2] + fhi[2] + 1e-3f};
    const int S[3] = {sx, sy, sz};
    const int E[3] = {ex, ey, nz};
    int org[3];
    bool fit = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float l = aff.c[r], h = aff.c[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float a = aff.a[3 * r + c];
        l += fminf(a * plo[c], a * phi[c]);
        h += fmaxf(a * plo[c], a * phi[c]);
      }
      l = fminf(fmaxf(l - 1e-2f, 0.f), (float)(S[r] - 1)) - job.shift[r];
      h = fminf(fmaxf(h + 1e-2f, 0.f), (float)(S[r] - 1)) - job.shift[r];
      // the floor index is clamped to S-2 (tile_voxels), so a tile sitting on the clamped face starts there
      const int i0 = min((int)floorf(l), S[r] - 2), i1 = min((int)floorf(h) + 1, S[r] - 1);
      org[r] = i0;
      fit = fit && (i1 - i0 + 1 <= E[r]);
    }
    // x: the flip mirrors the memory plane index; the box covers planes [m0, m0 + ex)
    int m0 = org[0];
    if (job.flip) m0 = sx - 1 - (org[0] + ex - 1);
    sh->origin[0] = m0;
    sh->origin[1] = org[1];
    sh->origin[2] = org[2];
    sh->fit = fit ? 1 : 0;
    if (fit) {
      const uint32_t bar = smem_u32(&sh->mbar);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, (uint32_t)(ex * ey * ez * 4 + ex * ey * ezs));
      tma_load_3d(smem_u32(s_img), &tp.img[jb], bar, org[2] & ~3, org[1], m0);
      tma_load_3d(smem_u32(s_seg), &tp.seg[jb], bar, org[2] & ~15, org[1], m0);
    }
  }
  __syncthreads();
  const int m0 = sh->origin[0], oy = sh->origin[1], oz = sh->origin[2] & ~3, ozs = sh->origin[2] & ~15;
  const unsigned bias3 = 0x4B000000u;
  if (sh->fit && !(tp.debug & 1)) {
    // box-local index = (sgn*fx + cx)*P + (fy - oy)*ez + (fz - oz), sgn = -1 / cx = sx-1-m0 when flipped
    const int sgn = job.flip ? -1 : 1;
    const int cx = job.flip ? sx - 1 - m0 : -m0;
    const int xs_i = sgn * ey * ez, xs_s = sgn * ey * ezs;
    const unsigned kimg = (unsigned)(cx * ey * ez - oy * ez - oz) - bias3 * (unsigned)(xs_i + ez + 1);
    const unsigned kseg = (unsigned)(cx * ey * ezs - oy * ezs - ozs) - bias3 * (unsigned)(xs_s + ezs + 1);
    mbar_wait(smem_u32(&sh->mbar), 0);
    tile_voxels<EPI, true>(job, aff, s_f, s_b, tp.fnodes, tp.bnodes, fzlo, bzlo, s_img, s_seg, kimg, kseg, xs_i, ez, xs_s, ezs, x0, y0, z0, sx, sy, sz,
                           (tp.debug & 2) ? (unsigned)(ex * ey * ez) : 0u, (unsigned)(ex * ey * ezs));
  } else {
    const int plane = sy * sz;
    const int xs = job.flip ? -plane : plane;
    const unsigned kb = (unsigned)(job.flip ? (sx - 1) * plane : 0) - bias3 * (unsigned)(xs + sz + 1);
    tile_voxels<EPI, false>(job, aff, s_f, s_b, tp.fnodes, tp.bnodes, fzlo, bzlo, job.src_img, job.src_seg, kb, kb, xs, sz, xs, sz, x0, y0, z0, sx, sy, sz);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
  }
  return fn;
}

static bool encode_box(EncodeTiledFn fn, CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* base, int sx, int sy, int sz, int bx, int by, int bz) {
  const cuuint64_t dims[3] = {(cuuint64_t)sz, (cuuint64_t)sy, (cuuint64_t)sx};
  const cuuint64_t strides[2] = {(cuuint64_t)sz * esize, (cuuint64_t)sy * sz * esize};
  const cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx};
  const cuuint32_t es[3] = {1, 1, 1};
  return fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launches the TMA-staged kernel for a batch of fast-eligible jobs; -1 = not applicable (the
// caller uses the full-z kernel), >0 = error.
int launch_warp_tile(const fsg_warp_job* jobs, int njobs, bool epi, int sx, int sy, int sz, cudaStream_t stream) {
  if (sx % TT || sy % TT || sz % TT || sx > 32767 || sy > 32767 || sz > 32767) return -1;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  static thread_local TileParams tp;  // 64-byte aligned tensor maps
  memset(&tp, 0, sizeof(tp));
  int box_floats = 0, box_bytes = 0, fnodes = 2, bnodes = 1;
  for (int n = 0; n < njobs; ++n) {
    const fsg_warp_job& j = jobs[n];
    if ((reinterpret_cast<uintptr_t>(j.src_img) & 15) || (reinterpret_cast<uintptr_t>(j.src_seg) & 15)) return -1;
    int e[3];
    for (int r = 0; r < 3; ++r) {
      float ext = 0.f, l1 = 0.f;
      for (int c = 0; c < 3; ++c) {
        ext += fabsf(j.A[3 * r + c]) * (TT - 1);
        l1 += fabsf(j.A[3 * r + c]);
      }
      // + displacement range of the control nodes that reach a tile (allowance: 3 voxels per axis;
      // tiles that need more fall back to global gathers) + floor/ceil neighbours
      e[r] = (int)ceilf(ext + 3.0f * l1) + 3;
    }
    const int S[3] = {sx, sy, sz};
    for (int r = 0; r < 3; ++r) e[r] = e[r] < S[r] ? e[r] : S[r];
    const int ez = (e[2] + 3 + 3) / 4 * 4, ezs = (e[2] + 15 + 15) / 16 * 16;
    if (e[0] > 256 || e[1] > 256 || ez > 256 || ezs > 256) return -1;
    tp.box[n][0] = e[0];
    tp.box[n][1] = e[1];
    tp.box[n][2] = e[2];
    tp.box[n][3] = ez;
    tp.box[n][4] = ezs;
    box_floats = box_floats > e[0] * e[1] * ez ? box_floats : e[0] * e[1] * ez;
    box_bytes = box_bytes > e[0] * e[1] * ezs ? box_bytes : e[0] * e[1] * ezs;
    if (!encode_box(fn, &tp.img[n], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, j.src_img, sx, sy, sz, e[0], e[1], ez)) return -1;
    if (!encode_box(fn, &tp.seg[n], CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, j.src_seg, sx, sy, sz, e[0], e[1], ezs)) return -1;
    // z-nodes of the control grids that can reach one tile: spacing = sz / n nodes
    const int fn_ = (int)ceilf((float)(TT - 1) * j.fs[2] / sz) + 2;
    fnodes = fnodes > fn_ ? fnodes : fn_;
    if (j.bf_low) {
      const int bn_ = (int)ceilf((float)(TT - 1) * j.bs[2] / sz) + 2;
      bnodes = bnodes > bn_ ? bnodes : bn_;
    }
  }
  box_bytes = (box_bytes + 127) / 128 * 128;
  tp.debug = config().tile_debug;
  tp.fnodes = fnodes;
  tp.bnodes = bnodes;
  tp.box_floats = (box_floats + 31) / 32 * 32;
  tp.box_bytes = box_bytes;
  const size_t smem = (size_t)tp.box_floats * 4 + box_bytes + (size_t)TT * TT * fnodes * 16 + (size_t)TT * TT * bnodes * 4 + sizeof(TileShared) + 16;
  if (smem > 200 * 1024) return -1;
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  const dim3 grid((sx / TT) * (sy / TT) * (sz / TT), 1, njobs);
  if (epi) {
    cudaFuncSetAttribute(warp_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    warp_tile_kernel<true><<<grid, TILE_THREADS, smem, stream>>>(b, tp, sx, sy, sz);
  } else {
    cudaFuncSetAttribute(warp_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    warp_tile_kernel<false><<<grid, TILE_THREADS, smem, stream>>>(b, tp, sx, sy, sz);
  }
  return 0;
}

}  // namespace fsg
