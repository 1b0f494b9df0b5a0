// K2, pipelined TMA variant of the fused deformation kernel (r02).
//
// r01 ncu of the full-z gather kernel: 94 % of the L1 data-pipe wavefront peak (a warp's rotated source line
// crosses ~5.4 128-byte lines per gather, 9 gathers per voxel), issue slots half idle.  The r01 TMA tile kernel
// (warp_tile.cu) took the gathers off the LSU path but ran one 163 KB block per SM with nothing overlapped
// (2.17 ms vs 0.90 ms).  This kernel keeps the idea and pipelines it:
//   * one persistent block per SM, 16 consumer warps + 4 producer warps, two stages of
//     {image box, staged control rows, z tables, tile meta} in shared memory, full / empty mbarriers per stage;
//   * output tiles of 8 x 16 x 16 voxels (image box ~58 KB on average instead of 137 KB for 16^3 + labels), walked
//     in z-fastest order by all SMs so that neighbouring boxes hit L2;
//   * producers (tile n+1, while the consumers blend tile n): exact x/y blends of the control grids for the tile's
//     128 rows and the z nodes that reach it (the r01 phase A), their min / max bound the tile's source box by
//     interval arithmetic, one thread arms the stage's `full` barrier and issues ONE cp.async.bulk.tensor.3d for
//     the float image box (out-of-volume parts zero-filled by the TMA unit);
//   * consumers: the eight trilinear corners come from the staged box (shared-memory loads at fixed offsets from
//     one base index), the nearest label is one global byte gather; coordinates are computed exactly as in the
//     other kernels (separately rounded mul / add chain: bit-exact segmentation), two voxels per thread-iteration
//     with packed FADD2 / FFMA2; a warp's elected lane releases the stage on the `empty` barrier.
// A tile whose source box exceeds the tensor-map box (strong local field gradients) is blended from global memory
// by the same consumer code.
#include <cuda.h>
#include <stdlib.h>

#include "warp_common.cuh"

namespace fsg {

constexpr int PT_X = 8, PT_Y = 16, PT_Z = 16;  // output tile
constexpr int PC_WARPS = 16, PP_WARPS = 4;
constexpr int PC_THREADS = PC_WARPS * 32, PP_THREADS = PP_WARPS * 32, P_THREADS = PC_THREADS + PP_THREADS;
constexpr int P_ROWS = PT_X * PT_Y;  // (x, y) rows of a tile

struct PipeParams {
  CUtensorMap img[FSG_MAX_JOBS];
  int box[FSG_MAX_JOBS][4];  // ex, ey, nz = extents a tile may need; ez = z extent of the box (16-byte aligned start + slack)
  int fnodes, bnodes;        // control-grid z-nodes staged per (x, y) row
  int box_floats;            // capacity of one stage's image box (multiple of 32)
  int ntx, nty, ntz, njobs;
  int debug;                 // FSG_TILE_DEBUG bit 0: force the global-gather path
};

struct PipeMeta {
  int origin[3];  // memory-space index of the box start (x: plane index after the flip; z rounded down to 4)
  int fit;
  int fzlo, bzlo;
  int pad[2];
};

__device__ __forceinline__ uint32_t p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void p_mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void p_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void p_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void p_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "P_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra P_WAIT_DONE;\n"
      "bra P_WAIT_LOOP;\n"
      "P_WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void p_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void p_bar_producers() { asm volatile("bar.sync 1, %0;" ::"n"(PP_THREADS) : "memory"); }

struct PipeTabZ {  // staged per-tile z tables (one entry per z of the tile)
  int f, c;
  float wc, wf;
};

// Gather + blend + epilogue + store for the 4 voxels of one consumer thread (two x-pairs).
// SMEM: image corners come from the staged box (strides ys_img / xs_img floats), else from global memory.
template <bool EPI, bool SMEM>
__device__ __forceinline__ void pipe_voxels(const fsg_warp_job& job, const Affine& aff, const float4* __restrict__ s_f, const float* __restrict__ s_b, int fnodes, int bnodes, const PipeTabZ tf,
                                            const PipeTabZ tb, int fzlo, int bzlo, const float* __restrict__ img_base, unsigned kimg, int xs_img, int ys_img, int x0, int y0, int z0,
                                            int sx, int sy, int sz, unsigned dbg_limit = 0) {
  const int ctid = threadIdx.x;  // consumer thread 0..511
  const int zl = ctid & 15, yl = (ctid >> 4) & 15, xh = ctid >> 8;
  const int j = y0 + yl, k = z0 + zl;
  const float mx = (float)(sx - 1), my = (float)(sy - 1), mz = (float)(sz - 1);
  const float lx = __int_as_float(__float_as_int(mx) - 1), ly = __int_as_float(__float_as_int(my) - 1), lz = __int_as_float(__float_as_int(mz) - 1);
  const bool has_bias = EPI && job.bf_low != nullptr;
  const float gam = (EPI && job.has_gamma) ? job.gamma : 1.0f;
  const float c0 = (EPI && job.has_gamma) ? 8.22881869049588f * (1.0f - job.gamma) : 0.f;
  const float zc = sub_rn((float)k, job.center[2]), yc = sub_rn((float)j, job.center[1]);
  const P2 zc2 = pk(zc, zc), yc2 = pk(yc, yc), cen2 = pk(job.center[0], job.center[0]), magic2 = pk(MAGIC, MAGIC);
  const P2 sh2x = pk(job.shift[0], job.shift[0]), sh2y = pk(job.shift[1], job.shift[1]), sh2z = pk(job.shift[2], job.shift[2]);
  const P2 c2x = pk(aff.c[0], aff.c[0]), c2y = pk(aff.c[1], aff.c[1]), c2z = pk(aff.c[2], aff.c[2]);
  const P2 bwf2 = pk(tb.wf, tb.wf), bwc2 = pk(tb.wc, tb.wc), gam2 = pk(gam, gam), c02 = pk(c0, c0);
  const int xl0 = 4 * xh;
  // staged control-grid rows: s_f[(xl*PT_Y + yl)*fnodes + node - fzlo]
  const float4* pf = s_f + (xl0 * PT_Y + yl) * fnodes + (tf.f - fzlo);
  const float4* pc = s_f + (xl0 * PT_Y + yl) * fnodes + (tf.c - fzlo);
  const float* bf = s_b + (xl0 * PT_Y + yl) * bnodes + (has_bias ? tb.f - bzlo : 0);
  const float* bc = s_b + (xl0 * PT_Y + yl) * bnodes + (has_bias ? tb.c - bzlo : 0);
  const int frow = PT_Y * fnodes, brow = PT_Y * bnodes;
  const int plane = sy * sz;
  unsigned o = (unsigned)(((x0 + xl0) * sy + j) * sz + k);
  float* __restrict__ const dst_img = job.dst_img;
  uint8_t* __restrict__ const dst_seg = job.dst_seg;
  const uint8_t* __restrict__ const src_seg = job.src_seg;
  const int xs_seg = job.flip ? -plane : plane;
  const unsigned kseg = (unsigned)(job.flip ? (sx - 1) * plane : 0) - 0x4B000000u * (unsigned)(xs_seg + sz + 1);
  P2 xi2 = pk((float)(x0 + xl0), (float)(x0 + xl0 + 1));
  const char* const img_b = reinterpret_cast<const char*>(img_base);
  const int by_row = ys_img * 4, by_plane = xs_img * 4;
#pragma unroll
  for (int it = 0; it < 2; ++it, o += 2 * plane, xi2 = add2(xi2, pk(2.0f, 2.0f)), pf += 2 * frow, pc += 2 * frow, bf += 2 * brow, bc += 2 * brow) {
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
    // ---- floor (clamped to S-2) and weights
    const P2 tx2 = add2_rz(pk(fminf(iia, lx), fminf(iib, lx)), magic2), ty2 = add2_rz(pk(fminf(jja, ly), fminf(jjb, ly)), magic2), tz2 = add2_rz(pk(fminf(kka, lz), fminf(kkb, lz)), magic2);
    float txa, txb, tya, tyb, tza, tzb;
    upk(tx2, txa, txb);
    upk(ty2, tya, tyb);
    upk(tz2, tza, tzb);
    const unsigned ba = (unsigned)__float_as_int(txa) * (unsigned)xs_img + (unsigned)__float_as_int(tya) * (unsigned)ys_img + (unsigned)__float_as_int(tza) + kimg;
    const unsigned bb = (unsigned)__float_as_int(txb) * (unsigned)xs_img + (unsigned)__float_as_int(tyb) * (unsigned)ys_img + (unsigned)__float_as_int(tzb) + kimg;
    const P2 wx2 = sub2(ii, sub2(tx2, magic2)), wy2 = sub2(jj, sub2(ty2, magic2)), wz2 = sub2(kk, sub2(tz2, magic2));
    // ---- nearest segmentation gather from global memory (round-half-even by the magic add)
    float sxa, sxb, sya, syb, sza, szb;
    upk(add2(ii, magic2), sxa, sxb);
    upk(add2(jj, magic2), sya, syb);
    upk(add2(kk, magic2), sza, szb);
    const uint8_t laba = __ldg(src_seg + ((unsigned)__float_as_int(sxa) * (unsigned)xs_seg + (unsigned)__float_as_int(sya) * (unsigned)sz + (unsigned)__float_as_int(sza) + kseg));
    const uint8_t labb = __ldg(src_seg + ((unsigned)__float_as_int(sxb) * (unsigned)xs_seg + (unsigned)__float_as_int(syb) * (unsigned)sz + (unsigned)__float_as_int(szb) + kseg));
    P2 c000, c001, c010, c011, c100, c101, c110, c111;
    if (SMEM) {
      const float* a00 = img_base + (dbg_limit ? min(ba, dbg_limit) : ba);  // dbg_limit != 0: timing experiments with meaningless boxes
      const float* b00 = img_base + (dbg_limit ? min(bb, dbg_limit) : bb);
      c000 = pk(a00[0], b00[0]);
      c001 = pk(a00[1], b00[1]);
      c010 = pk(a00[ys_img], b00[ys_img]);
      c011 = pk(a00[ys_img + 1], b00[ys_img + 1]);
      c100 = pk(a00[xs_img], b00[xs_img]);
      c101 = pk(a00[xs_img + 1], b00[xs_img + 1]);
      c110 = pk(a00[xs_img + ys_img], b00[xs_img + ys_img]);
      c111 = pk(a00[xs_img + ys_img + 1], b00[xs_img + ys_img + 1]);
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
__global__ void __launch_bounds__(P_THREADS, 1) warp_pipe_kernel(const __grid_constant__ Batch<fsg_warp_job> batch, const __grid_constant__ PipeParams tp, int sx, int sy, int sz) {
  const int tid = threadIdx.x;
  extern __shared__ __align__(128) unsigned char s_raw[];
  // layout: [image box] x 2 | control rows (float4) x 2 | bias rows x 2 | z tables x 2 x 2 | meta x 2 | reduction x 2 | barriers
  float* s_img = reinterpret_cast<float*>(s_raw);
  float4* s_f = reinterpret_cast<float4*>(s_raw + (size_t)2 * tp.box_floats * 4);
  float* s_b = reinterpret_cast<float*>(s_f + 2 * P_ROWS * tp.fnodes);
  PipeTabZ* s_tz = reinterpret_cast<PipeTabZ*>(s_b + ((2 * P_ROWS * tp.bnodes + 3) & ~3));  // [2 stages][2 (field, bias)][PT_Z]
  PipeMeta* s_meta = reinterpret_cast<PipeMeta*>(s_tz + 2 * 2 * PT_Z);
  float* s_red = reinterpret_cast<float*>(s_meta + 2);  // [2][PP_WARPS][6]
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_red + 2 * PP_WARPS * 6 + 2);  // full[2], empty[2]

  if (tp.debug & 4) {  // timing experiments: the staged rows are never written; keep them finite
    for (int e = tid; e < 2 * P_ROWS * tp.fnodes; e += P_THREADS) s_f[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = tid; e < 2 * P_ROWS * tp.bnodes; e += P_THREADS) s_b[e] = 0.f;
  }
  if (tid == 0) {
    for (int s_ = 0; s_ < 2; ++s_) {
      p_mbar_init(p_smem_u32(s_bar + s_), 1);
      p_mbar_init(p_smem_u32(s_bar + 2 + s_), PC_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ntiles = tp.ntx * tp.nty * tp.ntz;
  const int total = ntiles * tp.njobs;
  const float inf = __int_as_float(0x7f800000);

  if (tid >= PC_THREADS) {
    // ================================================================ producers
    const int ptid = tid - PC_THREADS, pw = ptid >> 5, pl = ptid & 31;
    int n = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++n) {
      const int b = n & 1;
      p_mbar_wait(p_smem_u32(s_bar + 2 + b), ((n >> 1) & 1) ^ 1);  // the consumers released this stage
      const int jb = t / ntiles, tile = t - jb * ntiles;
      const fsg_warp_job& job = batch.j[jb];
      const int z0 = (tile % tp.ntz) * PT_Z, y0 = ((tile / tp.ntz) % tp.nty) * PT_Y, x0 = (tile / (tp.ntz * tp.nty)) * PT_X;
      const bool has_bias = EPI && job.bf_low != nullptr;
      float4* sf = s_f + b * P_ROWS * tp.fnodes;
      float* sb = s_b + b * P_ROWS * tp.bnodes;
      // ---- x/y blends of the control grid for the z-nodes that reach this tile (exact), with their min / max
      const int fzlo = job.ftab[2][z0].f;
      const int fzn = min((int)job.ftab[2][z0 + PT_Z - 1].c - fzlo + 1, tp.fnodes);
      float lo0 = inf, lo1 = inf, lo2 = inf, hi0 = -inf, hi1 = -inf, hi2 = -inf;
      {
        const int fy_n = job.fs[1], fz_n = job.fs[2];
        const int per = fzn * 3;
        float* sff = reinterpret_cast<float*>(sf);
        for (int e = ptid; e < ((tp.debug & 4) ? 0 : P_ROWS * per); e += PP_THREADS) {  // debug bit 2: timing experiment without the staging
          const int row = e / per, rem = e - row * per;
          const int xl = row / PT_Y, yl = row - xl * PT_Y;
          const int zn = rem / 3, ch = rem - zn * 3;
          const Tab tx = load_tab(job.ftab[0], x0 + xl), ty = load_tab(job.ftab[1], y0 + yl);
          const float* g = job.fsmall + (fzlo + zn) * 3 + ch;
          const int sys = fz_n * 3, sxs = fy_n * sys;
          const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * sys), tx.wc, __ldg(g + tx.c * sxs + ty.f * sys));
          const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * sys), tx.wc, __ldg(g + tx.c * sxs + ty.c * sys));
          const float v = blend(ty.wf, t1f, ty.wc, t1c);
          sff[(row * tp.fnodes + zn) * 4 + ch] = v;
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
          bzn = min((int)job.btab[2][z0 + PT_Z - 1].c - bzlo + 1, tp.bnodes);
        }
        for (int e = ptid; e < ((tp.debug & 4) ? 0 : P_ROWS * bzn); e += PP_THREADS) {
          const int row = e / bzn, zn = e - row * bzn;
          float v = 0.f;
          if (has_bias) {
            const int xl = row / PT_Y, yl = row - xl * PT_Y;
            const int by_n = job.bs[1], bz_n = job.bs[2];
            const Tab tx = load_tab(job.btab[0], x0 + xl), ty = load_tab(job.btab[1], y0 + yl);
            const float* g = job.bf_low + bzlo + zn;
            const int sxs = by_n * bz_n;
            const float t1f = blend(tx.wf, __ldg(g + tx.f * sxs + ty.f * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.f * bz_n));
            const float t1c = blend(tx.wf, __ldg(g + tx.f * sxs + ty.c * bz_n), tx.wc, __ldg(g + tx.c * sxs + ty.c * bz_n));
            v = blend(ty.wf, t1f, ty.wc, t1c) * 1.4426950408889634f;
          }
          sb[row * tp.bnodes + zn] = v;
        }
      }
      // ---- z tables of the tile (consumers read them from shared memory: no global-load latency at the tile start)
      if (ptid < PT_Z) {
        const Tab t_ = load_tab(job.ftab[2], z0 + ptid);
        s_tz[(b * 2 + 0) * PT_Z + ptid] = PipeTabZ{t_.f, t_.c, t_.wc, t_.wf};
      } else if (ptid < 2 * PT_Z) {
        PipeTabZ e = {0, 0, 0.f, 1.f};
        if (has_bias) {
          const Tab t_ = load_tab(job.btab[2], z0 + ptid - PT_Z);
          e = PipeTabZ{t_.f, t_.c, t_.wc, t_.wf};
        }
        s_tz[(b * 2 + 1) * PT_Z + ptid - PT_Z] = e;
      }
      lo0 = warp_min(lo0);
      lo1 = warp_min(lo1);
      lo2 = warp_min(lo2);
      hi0 = warp_max(hi0);
      hi1 = warp_max(hi1);
      hi2 = warp_max(hi2);
      if (pl == 0) {
        float* r = s_red + (b * PP_WARPS + pw) * 6;
        r[0] = lo0;
        r[1] = lo1;
        r[2] = lo2;
        r[3] = hi0;
        r[4] = hi1;
        r[5] = hi2;
      }
      p_bar_producers();  // staged rows, tables and partial bounds of all producer warps are written
      if (ptid == 0) {
        const Affine aff(job);
        float flo[3], fhi[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          flo[c] = s_red[(b * PP_WARPS) * 6 + c];
          fhi[c] = s_red[(b * PP_WARPS) * 6 + 3 + c];
          for (int w = 1; w < PP_WARPS; ++w) {
            flo[c] = fminf(flo[c], s_red[(b * PP_WARPS + w) * 6 + c]);
            fhi[c] = fmaxf(fhi[c], s_red[(b * PP_WARPS + w) * 6 + 3 + c]);
          }
        }
        // interval of (voxel - center + F) per axis, then of each affine row; margins absorb rounding
        const float plo[3] = {(float)x0 - job.center[0] + flo[0] - 1e-3f, (float)y0 - job.center[1] + flo[1] - 1e-3f, (float)z0 - job.center[2] + flo[2] - 1e-3f};
        const float phi[3] = {(float)(x0 + PT_X - 1) - job.center[0] + fhi[0] + 1e-3f, (float)(y0 + PT_Y - 1) - job.center[1] + fhi[1] + 1e-3f,
                              (float)(z0 + PT_Z - 1) - job.center[2] + fhi[2] + 1e-3f};
        const int S[3] = {sx, sy, sz};
        const int E[3] = {tp.box[jb][0], tp.box[jb][1], tp.box[jb][2]};
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
          // the floor index is clamped to S-2 (pipe_voxels), so a tile sitting on the clamped face starts there
          const int i0 = min((int)floorf(l), S[r] - 2), i1 = min((int)floorf(h) + 1, S[r] - 1);
          org[r] = i0;
          fit = fit && (i1 - i0 + 1 <= E[r]);
        }
        // x: the flip mirrors the memory plane index; the box covers planes [m0, m0 + ex)
        int m0 = org[0];
        if (job.flip) m0 = sx - 1 - (org[0] + E[0] - 1);
        fit = fit && !(tp.debug & 1);
        PipeMeta* me = s_meta + b;
        me->origin[0] = m0;
        me->origin[1] = org[1];
        me->origin[2] = org[2] & ~3;
        me->fit = fit ? 1 : 0;
        me->fzlo = fzlo;
        me->bzlo = bzlo;
        const uint32_t bar = p_smem_u32(s_bar + b);
        if (tp.debug & 4) {  // timing experiment: bounds are meaningless without the staged rows; aim the box at the volume centre
          fit = true;
          me->fit = 1;
          me->origin[0] = sx / 2 - E[0] / 2;
          me->origin[1] = sy / 2 - E[1] / 2;
          me->origin[2] = (sz / 2 - E[2] / 2) & ~3;
        }
        if (fit && (tp.debug & 8)) {  // timing experiment: no TMA at all
          p_mbar_arrive(bar);
        } else if (fit) {
          p_mbar_expect_tx(bar, (uint32_t)(E[0] * E[1] * tp.box[jb][3] * 4));
          p_tma_load_3d(p_smem_u32(s_img + (size_t)b * tp.box_floats), &tp.img[jb], bar, me->origin[2], me->origin[1], me->origin[0]);
        } else {
          p_mbar_arrive(bar);
        }
      }
    }
  } else {
    // ================================================================ consumers
    const int lane = tid & 31;
    const int zl = tid & 15;
    int n = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++n) {
      const int b = n & 1;
      const int jb = t / ntiles, tile = t - jb * ntiles;
      const fsg_warp_job& job = batch.j[jb];
      const int z0 = (tile % tp.ntz) * PT_Z, y0 = ((tile / tp.ntz) % tp.nty) * PT_Y, x0 = (tile / (tp.ntz * tp.nty)) * PT_X;
      const Affine aff(job);
      p_mbar_wait(p_smem_u32(s_bar + b), (n >> 1) & 1);  // box landed, rows / tables / meta staged
      const PipeMeta me = s_meta[b];
      const PipeTabZ tf = s_tz[(b * 2 + 0) * PT_Z + zl], tb = s_tz[(b * 2 + 1) * PT_Z + zl];
      const float4* sf = s_f + b * P_ROWS * tp.fnodes;
      const float* sb = s_b + b * P_ROWS * tp.bnodes;
      if (me.fit) {
        // box-local index = (sgn*fx + cx)*P + (fy - oy)*ez + (fz - oz), sgn = -1 / cx = sx-1-m0 when flipped
        const int ey = tp.box[jb][1], ez = tp.box[jb][3];
        const int sgn = job.flip ? -1 : 1;
        const int cx = job.flip ? sx - 1 - me.origin[0] : -me.origin[0];
        const int xs_i = sgn * ey * ez;
        const unsigned kimg = (unsigned)(cx * ey * ez - me.origin[1] * ez - me.origin[2]) - 0x4B000000u * (unsigned)(xs_i + ez + 1);
        const unsigned dbg_limit = (tp.debug & 4) ? (unsigned)(tp.box[jb][0] * ey * ez - (job.flip ? 0 : ey * ez) - ez - 2) : 0u;
        pipe_voxels<EPI, true>(job, aff, sf, sb, tp.fnodes, tp.bnodes, tf, tb, me.fzlo, me.bzlo, s_img + (size_t)b * tp.box_floats + ((tp.debug & 4) && job.flip ? ey * ez : 0), kimg, xs_i, ez, x0, y0,
                               z0, sx, sy, sz, dbg_limit);
      } else {
        const int plane = sy * sz;
        const int xs = job.flip ? -plane : plane;
        const unsigned kb = (unsigned)(job.flip ? (sx - 1) * plane : 0) - 0x4B000000u * (unsigned)(xs + sz + 1);
        pipe_voxels<EPI, false>(job, aff, sf, sb, tp.fnodes, tp.bnodes, tf, tb, me.fzlo, me.bzlo, job.src_img, kb, xs, sz, x0, y0, z0, sx, sy, sz);
      }
      __syncwarp();
      if (lane == 0) p_mbar_arrive(p_smem_u32(s_bar + 2 + b));  // this warp is done with the stage
    }
  }
}

typedef CUresult (*PipeEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PipeEncodeFn pipe_encode_fn() {
  static PipeEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PipeEncodeFn>(p);
    cudaGetLastError();
  }
  return fn;
}

// Launches the pipelined TMA kernel for a batch of fast-eligible jobs; -1 = not applicable (the caller uses the
// full-z gather kernel), > 0 = error.
int launch_warp_pipe(const fsg_warp_job* jobs, int njobs, bool epi, int sx, int sy, int sz, cudaStream_t stream) {
  if (sx % PT_X || sy % PT_Y || sz % PT_Z || sx > 32767 || sy > 32767 || sz > 32767) return -1;
  PipeEncodeFn fn = pipe_encode_fn();
  if (!fn) return -1;
  static thread_local PipeParams tp;  // 64-byte aligned tensor maps
  memset(&tp, 0, sizeof(tp));
  int box_floats = 0, fnodes = 2, bnodes = 1;
  const int T[3] = {PT_X, PT_Y, PT_Z};
  for (int n = 0; n < njobs; ++n) {
    const fsg_warp_job& j = jobs[n];
    if (!j.src_img || (reinterpret_cast<uintptr_t>(j.src_img) & 15)) return -1;
    int e[3];
    for (int r = 0; r < 3; ++r) {
      float ext = 0.f, l1 = 0.f;
      for (int c = 0; c < 3; ++c) {
        ext += fabsf(j.A[3 * r + c]) * (T[c] - 1);
        l1 += fabsf(j.A[3 * r + c]);
      }
      // + displacement range of the control nodes that reach a tile (allowance: 3 voxels per axis;
      // tiles that need more are blended from global memory) + floor/ceil neighbours
      e[r] = (int)ceilf(ext + 3.0f * l1) + 3;
    }
    const int S[3] = {sx, sy, sz};
    for (int r = 0; r < 3; ++r) e[r] = e[r] < S[r] ? e[r] : S[r];
    const int ez = (e[2] + 3 + 3) / 4 * 4;
    if (e[0] > 256 || e[1] > 256 || ez > 256) return -1;
    tp.box[n][0] = e[0];
    tp.box[n][1] = e[1];
    tp.box[n][2] = e[2];
    tp.box[n][3] = ez;
    box_floats = box_floats > e[0] * e[1] * ez ? box_floats : e[0] * e[1] * ez;
    const cuuint64_t dims[3] = {(cuuint64_t)sz, (cuuint64_t)sy, (cuuint64_t)sx};
    const cuuint64_t strides[2] = {(cuuint64_t)sz * 4, (cuuint64_t)sy * sz * 4};
    const cuuint32_t box[3] = {(cuuint32_t)ez, (cuuint32_t)e[1], (cuuint32_t)e[0]};
    const cuuint32_t es[3] = {1, 1, 1};
    if (fn(&tp.img[n], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(j.src_img), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return -1;
    // z-nodes of the control grids that can reach one tile: spacing = sz / n nodes
    const int fn_ = (int)ceilf((float)(PT_Z - 1) * j.fs[2] / sz) + 2;
    fnodes = fnodes > fn_ ? fnodes : fn_;
    if (j.bf_low) {
      const int bn_ = (int)ceilf((float)(PT_Z - 1) * j.bs[2] / sz) + 2;
      bnodes = bnodes > bn_ ? bnodes : bn_;
    }
  }
  tp.debug = config().tile_debug;
  tp.fnodes = fnodes;
  tp.bnodes = bnodes;
  tp.box_floats = (box_floats + 31) / 32 * 32;
  tp.ntx = sx / PT_X;
  tp.nty = sy / PT_Y;
  tp.ntz = sz / PT_Z;
  tp.njobs = njobs;
  const size_t smem = (size_t)2 * tp.box_floats * 4 + (size_t)2 * P_ROWS * fnodes * 16 + (size_t)((2 * P_ROWS * bnodes + 3) & ~3) * 4 + 2 * 2 * PT_Z * sizeof(PipeTabZ) + 2 * sizeof(PipeMeta) +
                      (2 * PP_WARPS * 6 + 2) * 4 + 4 * 8 + 128;
  if (smem > 225 * 1024) return -1;
  Batch<fsg_warp_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  const int total = tp.ntx * tp.nty * tp.ntz * njobs;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const dim3 grid(total < sms ? total : sms, 1, 1);
  if (epi) {
    cudaFuncSetAttribute(warp_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    warp_pipe_kernel<true><<<grid, P_THREADS, smem, stream>>>(b, tp, sx, sy, sz);
  } else {
    cudaFuncSetAttribute(warp_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    warp_pipe_kernel<false><<<grid, P_THREADS, smem, stream>>>(b, tp, sx, sy, sz);
  }
  return 0;
}

}  // namespace fsg
