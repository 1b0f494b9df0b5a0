// K4c — separable linear zoom (myzoom_torch, utils/generation.py:310-397) with the global
// reductions that follow it on this path: /max of RandResample.resize_back
// (augmentation/synthseg.py:109-114) and ScaleIntensity(0,1) (data/datasets.py:311).
//
// Separable in the reference's order (x, then y, then z; w_f*X[f] + w_c*X[c] per axis): a block stages the
// x-blend of the coarse rows it needs in shared memory, a warp blends two of them along y into its own row and
// every lane then blends along z (zoom_plane_kernel below).
// The global max/min need every up-sampled value, so the zoom runs twice over the (L2-resident)
// coarse volume: a reduce pass, then the write pass with the normalisation fused.
#include "common.cuh"

namespace fsg {

constexpr int ZM_THREADS = 256;
constexpr int ZM_WARPS = ZM_THREADS / 32;

// a / b with one Newton correction on top of the reciprocal: correctly rounded for all but a
// vanishing fraction of inputs, and exactly 1 for a == b (the image maximum must map to 1).
__device__ __forceinline__ float div_nr(float a, float b, float rb) {
  const float y = __fmul_rn(a, rb);
  const float e = __fmaf_rn(-b, y, a);
  return __fmaf_rn(e, rb, y);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Block = one output x-plane i (r02: a block per 32 rows spent 40 % of its instructions on its prologue, table
// loads and reduction epilogue; ncu 38 instructions per output voxel).  The plane is produced in chunks of RJ rows:
//   phase 1: the coarse rows the chunk reaches (NR <= cap of them, contiguous in memory) are blended along x ONCE
//            into shared memory (two coalesced loads per coarse element; the r01 kernel re-did this for every
//            output row: 4 loads per coarse element and row);
//   phase 2: a warp takes an output row: y-blend of two staged rows into its private row (lane = K, conflict
//            free), then the z-blend with lane = k mod 32 (neighbouring lanes read the same or neighbouring
//            words: no bank conflicts) and coalesced 128-byte stores.
// Several blocks are resident per SM, so one block's phase-1 loads overlap the others' phase 2.
// NE > 0: sz == 32*NE, the lane's z-table entries live in registers as shared-memory addresses.  NE == 0: generic
// extents, table in smem.  REDUCE computes exactly the same values (same operations in the same order) and only
// keeps min / max.
template <bool REDUCE, int NE>
__global__ void __launch_bounds__(ZM_THREADS) zoom_plane_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int sx, int sy, int sz, int rj, int cap, int pitch) {
  const fsg_zoom_job& job = batch.j[blockIdx.y];
  const int n1 = job.n[1], n2 = job.n[2];
  const float* __restrict__ const src = job.src;
  float* __restrict__ const dst = job.dst;
  const fsg_tab* __restrict__ const taby = job.tab[1];
  const fsg_tab* __restrict__ const tabz = job.tab[2];
  extern __shared__ __align__(16) float s_zoom[];
  // layout: staged x-blended rows [cap][pitch] | one row per warp [ZM_WARPS][pitch] | y table [sy] (int2) | generic: z table [sz] (int2)
  float* s_a = s_zoom;
  float* s_row = s_zoom + (size_t)cap * pitch + (threadIdx.x >> 5) * pitch;
  int2* s_ty = reinterpret_cast<int2*>(s_zoom + (((size_t)(cap + ZM_WARPS) * pitch + 1) & ~(size_t)1));  // 8-byte aligned
  int2* s_tz = s_ty + sy;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x;
  const uint32_t row_s = (uint32_t)__cvta_generic_to_shared(s_row);

  constexpr int NK = NE > 0 ? NE : 1;
  uint32_t af[NK], ac[NK];  // shared-memory addresses of the two z neighbours in this warp's row
  float twc[NK], twf[NK];
  if (NE > 0) {
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const fsg_tab e = tabz[lane + 32 * q];
      af[q] = row_s + (uint32_t)e.f * 4u;
      ac[q] = row_s + (uint32_t)e.c * 4u;
      twc[q] = e.wc;
      twf[q] = sub_rn(1.0f, e.wc);
    }
  } else {
    for (int k = threadIdx.x; k < sz; k += ZM_THREADS) {
      const fsg_tab e = tabz[k];
      s_tz[k] = make_int2(((int)e.f * 4) | (((int)e.c * 4) << 16), __float_as_int(e.wc));
    }
  }
  for (int j = threadIdx.x; j < sy; j += ZM_THREADS) {
    const fsg_tab e = taby[j];
    s_ty[j] = make_int2((int)e.f | ((int)e.c << 16), __float_as_int(e.wc));
  }

  const float inf = __int_as_float(0x7f800000);
  float lo_ = inf, hi_ = -inf;
  // post 1: v / max (exactly 1 at the maximum).  post 2: ((v / max) - min / max) / (1 - min / max)
  // folded into one FMA, clamped to [0, 1] (float path: tolerance, not bit parity).
  float vmax = 1.f, rmax = 1.f, pa = 1.f, pb = 0.f;
  float vmax_eq = __int_as_float(0x7fc00000);  // value that maps to exactly 1 under post 2 (NaN: none, for a constant image)
  const int post = job.post;
  if (!REDUCE && post > 0) {
    vmax = job.minmax[1];
    rmax = __frcp_rn(vmax);
    const float qmin = div_nr(job.minmax[0], vmax, rmax);
    const float den = sub_rn(div_nr(vmax, vmax, rmax), qmin);
    const bool flat = den == 0.f;  // constant image: ScaleIntensity returns x * minv = 0
    pa = flat ? 0.f : __fdiv_rn(rmax, den);
    pb = flat ? 0.f : -__fdiv_rn(qmin, den);
    if (!flat) vmax_eq = vmax;
  }
  const Tab tx = load_tab(job.tab[0], i);
  const float* __restrict__ const plane_f = src + (size_t)tx.f * n1 * n2;
  const float* __restrict__ const plane_c = src + (size_t)tx.c * n1 * n2;
  const bool vec_ok = (n2 & 3) == 0 && ((reinterpret_cast<uintptr_t>(plane_f) | reinterpret_cast<uintptr_t>(plane_c)) & 15) == 0;  // pitch % 4 == 0 always
  __syncthreads();

  for (int j0 = 0; j0 < sy; j0 += rj) {
    const int j1 = min(j0 + rj, sy);
    // ---- phase 1: x-blend of the coarse rows [r0, r0 + nr) of planes tx.f / tx.c
    const int r0 = s_ty[j0].x & 0xffff;
    const int nr = min((s_ty[j1 - 1].x >> 16) - r0 + 1, cap);
    for (int r = warp; r < nr; r += ZM_WARPS) {
      // a warp per coarse row (no index division; r02 ncu: the flat scalar loop with e / n2 cost more than phase 2)
      const float* __restrict__ pf = plane_f + (size_t)(r0 + r) * n2;
      const float* __restrict__ pc = plane_c + (size_t)(r0 + r) * n2;
      float* __restrict__ sa = s_a + r * pitch;
      if (vec_ok) {
        const P2 wf2 = pk(tx.wf, tx.wf), wc2 = pk(tx.wc, tx.wc);
        for (int K4 = lane; K4 < (n2 >> 2); K4 += 32) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(pf) + K4), b = __ldg(reinterpret_cast<const float4*>(pc) + K4);
          float4 o;
          upk(fma2(wf2, pk(a.x, a.y), mul2(wc2, pk(b.x, b.y))), o.x, o.y);
          upk(fma2(wf2, pk(a.z, a.w), mul2(wc2, pk(b.z, b.w))), o.z, o.w);
          reinterpret_cast<float4*>(sa)[K4] = o;
        }
      } else {
#pragma unroll 4
        for (int K = lane; K < n2; K += 32) sa[K] = __fmaf_rn(tx.wf, __ldg(pf + K), __fmul_rn(tx.wc, __ldg(pc + K)));
      }
    }
    __syncthreads();

    // ---- phase 2: one output row per warp trip
    for (int j = j0 + warp; j < j1; j += ZM_WARPS) {
      const int2 tyj = s_ty[j];
      const float ywc = __int_as_float(tyj.y), ywf = sub_rn(1.0f, ywc);
      const float* __restrict__ ra = s_a + min((tyj.x & 0xffff) - r0, cap - 1) * pitch;
      const float* __restrict__ rb = s_a + min((tyj.x >> 16) - r0, cap - 1) * pitch;
      __syncwarp();  // previous row's readers are done
      for (int K4 = lane; K4 < ((n2 + 3) >> 2); K4 += 32) {  // pitch % 4 == 0: the <= 3 pad columns are computed and never read
        const float4 a = reinterpret_cast<const float4*>(ra)[K4], b = reinterpret_cast<const float4*>(rb)[K4];
        float4 o;
        o.x = __fmaf_rn(ywf, a.x, __fmul_rn(ywc, b.x));
        o.y = __fmaf_rn(ywf, a.y, __fmul_rn(ywc, b.y));
        o.z = __fmaf_rn(ywf, a.z, __fmul_rn(ywc, b.z));
        o.w = __fmaf_rn(ywf, a.w, __fmul_rn(ywc, b.w));
        reinterpret_cast<float4*>(s_row)[K4] = o;
      }
      __syncwarp();
      float* __restrict__ out = REDUCE ? nullptr : dst + ((size_t)i * sy + j) * sz + lane;
      if (NE > 0) {
        float v[NK];
#pragma unroll
        for (int q = 0; q < NK; ++q) v[q] = __fmaf_rn(twf[q], lds_f32(af[q]), __fmul_rn(twc[q], lds_f32(ac[q])));
#pragma unroll
        for (int q = 0; q < NK; ++q) {
          if (REDUCE) {
            lo_ = fminf(lo_, v[q]);
            hi_ = fmaxf(hi_, v[q]);
          } else {
            float val = v[q];
            if (post == 1) val = div_nr(val, vmax, rmax);
            if (post == 2) val = val == vmax_eq ? 1.f : __saturatef(__fmaf_rn(val, pa, pb));  // the maximum maps to exactly 1 (ScaleIntensity)
            __stcs(out + 32 * q, val);
          }
        }
      } else {
        for (int k = lane; k < sz; k += 32) {
          const int2 e = s_tz[k];
          const float wc = __int_as_float(e.y), wf = sub_rn(1.0f, wc);
          float val = __fmaf_rn(wf, lds_f32(row_s + (uint32_t)(e.x & 0xffff)), __fmul_rn(wc, lds_f32(row_s + (uint32_t)(e.x >> 16))));
          if (REDUCE) {
            lo_ = fminf(lo_, val);
            hi_ = fmaxf(hi_, val);
          } else {
            if (post == 1) val = div_nr(val, vmax, rmax);
            if (post == 2) val = val == vmax_eq ? 1.f : __saturatef(__fmaf_rn(val, pa, pb));
            out[k - lane] = val;
          }
        }
      }
    }
    __syncthreads();  // the next chunk's phase 1 overwrites the staged rows
  }
  if (REDUCE) {
    float lo = warp_min(lo_), hi = warp_max(hi_);
    __shared__ float slo[ZM_WARPS], shi[ZM_WARPS];
    if (lane == 0) {
      slo[warp] = lo;
      shi[warp] = hi;
    }
    __syncthreads();
    if (warp == 0) {
      lo = lane < ZM_WARPS ? slo[lane] : inf;
      hi = lane < ZM_WARPS ? shi[lane] : -inf;
      lo = warp_min(lo);
      hi = warp_max(hi);
      if (lane == 0) {
        atomicMin(reinterpret_cast<int*>(job.minmax), float_to_ordered(lo));
        atomicMax(reinterpret_cast<int*>(job.minmax) + 1, float_to_ordered(hi));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ walk kernel
// Production kernel (sz in {128, 256, 512}).  r02 ncu of zoom_plane_kernel: 38 (reduce) / 53 (write) instructions
// per output voxel, issue-bound at 60-65 % with three smem round trips per output row; every output plane redid the
// gather-heavy z stage.  Here the axes are blended in the order z, y, x (the interpolation is separable and the image
// path is a tolerance, not bit parity: the sums differ from the reference's x, y, z order by rounding only), so that
//   * the gathers (z) run once per COARSE plane and only on the coarse rows a block's row chunk reaches;
//   * the stage that runs once per output voxel (x) is a register-only blend of two planes of the y/z-upsampled
//     chunk held in registers (prev, cur) and one 16-byte store: ~6 instructions per output voxel.
// Block = (job, chunk of RJ output rows, segment of output planes), 256 threads, 8 float4 slots each
// (RJ * sz / 4 = 2048 slots).  It walks the coarse planes X its segment needs:
//   cp.async of the chunk's coarse rows of plane X+1 (double buffered)  ||  z stage of X: coarse rows -> s_tz[r][sz]
//   y stage: cur[m] = w_f * s_tz[f_j] + w_c * s_tz[c_j]   (float4, registers)
//   x stage: every output plane i with c_i == X: out = w_f * prev + w_c * cur (+ /max, ScaleIntensity | min/max)
constexpr int ZW_THREADS = 256;
#ifndef FSG_ZW_M
#define FSG_ZW_M 8
#endif
#ifndef FSG_ZW_MINB
#define FSG_ZW_MINB 2
#endif
constexpr int ZW_M = FSG_ZW_M;  // float4 slots per thread and plane (A/B: -DFSG_ZW_M=4 -DFSG_ZW_MINB=3)

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

template <bool REDUCE, int SZ4>
__global__ void __launch_bounds__(ZW_THREADS, FSG_ZW_MINB) zoom_walk_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int sx, int sy, int cap, int n2max, int xsegs) {
  constexpr int SZ = 4 * SZ4, RJ = ZW_THREADS * ZW_M / SZ4, RSTEP = ZW_THREADS / SZ4;  // rows per chunk, row step between a thread's slots
  const fsg_zoom_job& job = batch.j[blockIdx.y];
  const int n0 = job.n[0], n1 = job.n[1], n2 = job.n[2];
  const float* __restrict__ const src = job.src;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = blockIdx.x % xsegs, chunk = blockIdx.x / xsegs;
  const int j0 = chunk * RJ;
  if (j0 >= sy) return;
  const int j1 = min(j0 + RJ, sy);
  const int i0 = (int)(((int64_t)sx * seg) / xsegs), i1 = (int)(((int64_t)sx * (seg + 1)) / xsegs);

  extern __shared__ __align__(16) float s_zoom[];
  // layout: coarse rows [2][cap][n2max] | z-upsampled rows [cap][SZ] | y table of the chunk [RJ] (int2) | x table [sx] (int2)
  float* s_c = s_zoom;
  float* s_tz = s_zoom + (((size_t)2 * cap * n2max + 3) & ~(size_t)3);
  int2* s_ty = reinterpret_cast<int2*>(s_tz + (size_t)cap * SZ);
  int2* s_tx = s_ty + RJ;

  const int k4 = tid % SZ4, jrow0 = tid / SZ4;
  // z stage mapping: thread = one output column k (lanes = consecutive k: neighbouring lanes read the same or the
  // neighbouring coarse word, no bank conflicts; r02a ncu of the float4-per-thread mapping: 2-way conflicts on every
  // gather, 11 M conflict wavefronts per launch), walking the chunk's coarse rows
  constexpr int NKZ = SZ > ZW_THREADS ? SZ / ZW_THREADS : 1, ZROWS = SZ < ZW_THREADS ? ZW_THREADS / SZ : 1;
  const int zk = tid % SZ, zr0 = tid / SZ;  // SZ >= 256: zk = tid, zr0 = 0
  int zf[NKZ], zc[NKZ];
  float zwc[NKZ], zwf[NKZ];
#pragma unroll
  for (int e = 0; e < NKZ; ++e) {
    const fsg_tab t = job.tab[2][zk + e * ZW_THREADS];
    zf[e] = (int)t.f * 4;
    zc[e] = (int)t.c * 4;
    zwc[e] = t.wc;
    zwf[e] = sub_rn(1.0f, t.wc);
  }
  for (int j = tid; j < RJ; j += ZW_THREADS) {
    const fsg_tab t = job.tab[1][min(j0 + j, sy - 1)];
    s_ty[j] = make_int2((int)t.f | ((int)t.c << 16), __float_as_int(t.wc));
  }
  for (int i = tid; i < sx; i += ZW_THREADS) {
    const fsg_tab t = job.tab[0][i];
    s_tx[i] = make_int2((int)t.f | ((int)t.c << 16), __float_as_int(t.wc));
  }
  const int r0 = job.tab[1][j0].f;
  const int nr = min((int)job.tab[1][j1 - 1].c - r0 + 1, cap);

  const float inf = __int_as_float(0x7f800000);
  float lo_ = inf, hi_ = -inf;
  // post 1: v / max (exactly 1 at the maximum).  post 2: ((v / max) - min / max) / (1 - min / max)
  // folded into one FMA, clamped to [0, 1] (float path: tolerance, not bit parity).
  float vmax = 1.f, rmax = 1.f, pa = 1.f, pb = 0.f;
  float vmax_eq = __int_as_float(0x7fc00000);  // value that maps to exactly 1 under post 2 (NaN: none, for a constant image)
  const int post = job.post;
  if (!REDUCE && post > 0) {
    vmax = job.minmax[1];
    rmax = __frcp_rn(vmax);
    const float qmin = div_nr(job.minmax[0], vmax, rmax);
    const float den = sub_rn(div_nr(vmax, vmax, rmax), qmin);
    const bool flat = den == 0.f;  // constant image: ScaleIntensity returns x * minv = 0
    pa = flat ? 0.f : __fdiv_rn(rmax, den);
    pb = flat ? 0.f : -__fdiv_rn(qmin, den);
    if (!flat) vmax_eq = vmax;
  }
  __syncthreads();

  const int x_first = s_tx[i0].x & 0xffff, x_last = s_tx[i1 - 1].x >> 16;
  const int ncopy = nr * n2;  // the chunk's coarse rows of one plane are contiguous in memory
  auto issue_load = [&](int X, int b) {
    const float* __restrict__ g = src + ((size_t)X * n1 + r0) * n2;
    float* d = s_c + (size_t)b * cap * n2max;
    for (int e = tid; e < ncopy; e += ZW_THREADS) cp_async4(d + e, g + e);
  };
  issue_load(x_first, 0);

  float4 prev[ZW_M], cur[ZW_M];
#pragma unroll
  for (int m = 0; m < ZW_M; ++m) prev[m] = cur[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* __restrict__ const out0 = REDUCE ? nullptr : job.dst + ((size_t)j0 + jrow0) * SZ + 4 * k4;
  int i = i0, b = 0;
  for (int X = x_first; X <= x_last; ++X, b ^= 1) {
    cp_async_commit_wait_all();
    __syncthreads();  // plane X's coarse rows are in s_c[b]; the y stage of plane X-1 is done with s_tz
    // ---- z stage: s_tz[r][zk] for the rows r = zr0, zr0 + ZROWS, ...
    {
      const char* cb = reinterpret_cast<const char*>(s_c + (size_t)b * cap * n2max);
#pragma unroll 2
      for (int r = zr0; r < nr; r += ZROWS) {
        const char* row = cb + (size_t)r * n2 * 4;
#pragma unroll
        for (int e = 0; e < NKZ; ++e)
          s_tz[(size_t)r * SZ + zk + e * ZW_THREADS] = __fmaf_rn(zwf[e], *reinterpret_cast<const float*>(row + zf[e]), __fmul_rn(zwc[e], *reinterpret_cast<const float*>(row + zc[e])));
      }
    }
    __syncthreads();
    if (X + 1 <= x_last) issue_load(X + 1, b ^ 1);  // s_c[b ^ 1] was last read by the z stage of plane X-1
    // ---- y stage (registers)
#pragma unroll
    for (int m = 0; m < ZW_M; ++m) {
      prev[m] = cur[m];
      const int2 ty = s_ty[jrow0 + RSTEP * m];
      const float wc = __int_as_float(ty.y), wf = sub_rn(1.0f, wc);
      const float4 a = reinterpret_cast<const float4*>(s_tz + (size_t)min((ty.x & 0xffff) - r0, cap - 1) * SZ)[k4];
      const float4 c = reinterpret_cast<const float4*>(s_tz + (size_t)min((ty.x >> 16) - r0, cap - 1) * SZ)[k4];
      cur[m].x = __fmaf_rn(wf, a.x, __fmul_rn(wc, c.x));
      cur[m].y = __fmaf_rn(wf, a.y, __fmul_rn(wc, c.y));
      cur[m].z = __fmaf_rn(wf, a.z, __fmul_rn(wc, c.z));
      cur[m].w = __fmaf_rn(wf, a.w, __fmul_rn(wc, c.w));
    }
    // ---- x stage: output planes whose upper neighbour is plane X (block-uniform loop)
    while (i < i1) {
      const int2 tx = s_tx[i];
      const int fx = tx.x & 0xffff, cx = tx.x >> 16;
      if (cx > X) break;
      if (cx == X) {
        const float wc = __int_as_float(tx.y), wf = sub_rn(1.0f, wc);
        const bool lower_is_cur = fx == X;  // clamped end of the table: both neighbours are plane X
        float* __restrict__ out = REDUCE ? nullptr : out0 + (size_t)i * sy * SZ;
#pragma unroll
        for (int m = 0; m < ZW_M; ++m) {
          const float4 a = lower_is_cur ? cur[m] : prev[m];
          float v[4];
          v[0] = __fmaf_rn(wf, a.x, __fmul_rn(wc, cur[m].x));
          v[1] = __fmaf_rn(wf, a.y, __fmul_rn(wc, cur[m].y));
          v[2] = __fmaf_rn(wf, a.z, __fmul_rn(wc, cur[m].z));
          v[3] = __fmaf_rn(wf, a.w, __fmul_rn(wc, cur[m].w));
          if (REDUCE) {
            if (j0 + jrow0 + RSTEP * m < sy) {
              lo_ = fminf(fminf(lo_, v[0]), fminf(fminf(v[1], v[2]), v[3]));
              hi_ = fmaxf(fmaxf(hi_, v[0]), fmaxf(fmaxf(v[1], v[2]), v[3]));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (post == 1) v[e] = div_nr(v[e], vmax, rmax);
              if (post == 2) v[e] = v[e] == vmax_eq ? 1.f : __saturatef(__fmaf_rn(v[e], pa, pb));  // the maximum maps to exactly 1 (ScaleIntensity)
            }
            if (j0 + jrow0 + RSTEP * m < sy) __stcs(reinterpret_cast<float4*>(out + (size_t)RSTEP * m * SZ), make_float4(v[0], v[1], v[2], v[3]));
          }
        }
      }
      ++i;
    }
  }
  if (REDUCE) {
    float lo = warp_min(lo_), hi = warp_max(hi_);
    __shared__ float slo[ZW_THREADS / 32], shi[ZW_THREADS / 32];
    if (lane == 0) {
      slo[warp] = lo;
      shi[warp] = hi;
    }
    __syncthreads();
    if (warp == 0) {
      lo = lane < ZW_THREADS / 32 ? slo[lane] : inf;
      hi = lane < ZW_THREADS / 32 ? shi[lane] : -inf;
      lo = warp_min(lo);
      hi = warp_max(hi);
      if (lane == 0) {
        atomicMin(reinterpret_cast<int*>(job.minmax), float_to_ordered(lo));
        atomicMax(reinterpret_cast<int*>(job.minmax) + 1, float_to_ordered(hi));
      }
    }
  }
}

__global__ void zoom_mm_init_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int njobs) {
  const int t = threadIdx.x;
  if (t < njobs) {
    int* p = reinterpret_cast<int*>(batch.j[t].minmax);
    p[0] = float_to_ordered(__int_as_float(0x7f800000));
    p[1] = float_to_ordered(__int_as_float(0xff800000));
  }
}
__global__ void zoom_mm_final_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int njobs) {
  const int t = threadIdx.x;
  if (t < njobs * 2) {
    float* p = batch.j[t / 2].minmax + (t % 2);
    *p = ordered_to_float(*reinterpret_cast<int*>(p));
  }
}

static int check_zoom(const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, bool need_dst, const char* who, int* max_n2) {
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1 && (int64_t)sx * sy < ((int64_t)1 << 31), "%s: bad shape", who);
  *max_n2 = 1;
  for (int i = 0; i < njobs; ++i) {
    const fsg_zoom_job& j = jobs[i];
    FSG_REQUIRE(j.n[0] >= 1 && j.n[1] >= 1 && j.n[2] >= 1, "%s: job %d bad source shape", who, i);
    FSG_REQUIRE(j.n[0] <= 32767 && j.n[1] <= 32767 && j.n[2] <= 32767, "%s: source extent exceeds the int16 table range", who);
    FSG_REQUIRE(j.src && j.tab[0] && j.tab[1] && j.tab[2], "%s: job %d has a NULL src/table", who, i);
    FSG_REQUIRE(!need_dst || j.dst, "%s: job %d has a NULL dst", who, i);
    FSG_REQUIRE(j.post >= 0 && j.post <= 2, "%s: job %d post must be 0..2", who, i);
    FSG_REQUIRE((j.post == 0 && need_dst) || j.minmax, "%s: job %d needs a minmax buffer", who, i);
    if (j.n[2] > *max_n2) *max_n2 = j.n[2];
  }
  return 0;
}

constexpr int ZM_RJ = 32;  // output rows per block

template <bool REDUCE, int NE>
static void launch_zoom_ne(const Batch<fsg_zoom_job>& b, int njobs, int sx, int sy, int sz, int cap, int pitch, size_t smem, cudaStream_t s) {
  auto k = zoom_plane_kernel<REDUCE, NE>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<dim3((unsigned)sx, njobs), ZM_THREADS, smem, s>>>(b, sx, sy, sz, ZM_RJ, cap, pitch);
}

template <bool REDUCE, int SZ4>
static bool launch_zoom_walk(const Batch<fsg_zoom_job>& b, const fsg_zoom_job* jobs, int njobs, int sx, int sy, int max_n2, cudaStream_t s) {
  constexpr int RJ = ZW_THREADS * ZW_M / SZ4;
  int cap = 2;
  for (int i = 0; i < njobs; ++i) {
    const int c = (int)(((int64_t)(RJ - 1) * jobs[i].n[1]) / sy) + 3;
    const int cc = c < jobs[i].n[1] ? c : jobs[i].n[1];
    cap = cap > cc ? cap : cc;
    if (REDUCE == false && (reinterpret_cast<uintptr_t>(jobs[i].dst) & 15)) return false;
  }
  const size_t smem = ((((size_t)2 * cap * max_n2 + 3) & ~(size_t)3) + (size_t)cap * 4 * SZ4 + 2 * (size_t)RJ + 2 * (size_t)sx) * sizeof(float);
  if (smem > (FSG_ZW_MINB >= 3 ? 72 : 110) * 1024) return false;  // FSG_ZW_MINB blocks per SM
  const int nchunk = (sy + RJ - 1) / RJ;
  // segments of output planes: enough blocks for >= 2 waves of 2 blocks per SM, each segment at least 8 planes long
  int xsegs = (148 * 4 + nchunk * njobs - 1) / (nchunk * njobs);
  xsegs = xsegs < 1 ? 1 : xsegs;
  const int seg_cap = sx / 8 > 1 ? sx / 8 : 1;
  xsegs = xsegs < seg_cap ? xsegs : seg_cap;
  auto k = zoom_walk_kernel<REDUCE, SZ4>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<dim3((unsigned)(nchunk * xsegs), njobs), ZW_THREADS, smem, s>>>(b, sx, sy, cap, max_n2, xsegs);
  return true;
}

template <bool REDUCE>
static int launch_zoom(const Batch<fsg_zoom_job>& b, const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, int max_n2, cudaStream_t s, const char* who) {
  if (sz == 256 && launch_zoom_walk<REDUCE, 64>(b, jobs, njobs, sx, sy, max_n2, s)) return 0;
  if (sz == 128 && launch_zoom_walk<REDUCE, 32>(b, jobs, njobs, sx, sy, max_n2, s)) return 0;
  if (sz == 512 && launch_zoom_walk<REDUCE, 128>(b, jobs, njobs, sx, sy, max_n2, s)) return 0;
  // coarse rows one block can reach: the y table advances by n1 / sy per output row (myzoom_torch positions)
  int cap = 2;
  for (int i = 0; i < njobs; ++i) {
    const int c = (int)(((int64_t)(ZM_RJ - 1) * jobs[i].n[1]) / sy) + 3;
    const int cc = c < jobs[i].n[1] ? c : jobs[i].n[1];
    cap = cap > cc ? cap : cc;
  }
  const int pitch = (max_n2 + 3) & ~3;  // rows of the staged planes are 16-byte aligned
  const size_t smem = ((size_t)(cap + ZM_WARPS) * pitch + 2 + 2 * (size_t)sy + 2 * (size_t)sz) * sizeof(float);
  FSG_REQUIRE(smem <= 200 * 1024, "%s: rows of %d / %d voxels do not fit in shared memory", who, sz, max_n2);
  const int ne = (sz % 32 == 0 && sz <= 384) ? sz / 32 : 0;
  switch (ne) {
    case 4: launch_zoom_ne<REDUCE, 4>(b, njobs, sx, sy, sz, cap, pitch, smem, s); break;
    case 8: launch_zoom_ne<REDUCE, 8>(b, njobs, sx, sy, sz, cap, pitch, smem, s); break;
    case 12: launch_zoom_ne<REDUCE, 12>(b, njobs, sx, sy, sz, cap, pitch, smem, s); break;
    default: launch_zoom_ne<REDUCE, 0>(b, njobs, sx, sy, sz, cap, pitch, smem, s); break;
  }
  return 0;
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_zoom_minmax(const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_zoom_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  int max_n2;
  if (int rc = check_zoom(jobs, njobs, sx, sy, sz, false, "fsg_zoom_minmax", &max_n2)) return rc;
  cudaStream_t s = as_stream(stream);
  zoom_mm_init_kernel<<<1, 32, 0, s>>>(b, njobs);
  if (int rc = launch_zoom<true>(b, jobs, njobs, sx, sy, sz, max_n2, s, "fsg_zoom_minmax")) return rc;
  zoom_mm_final_kernel<<<1, 32, 0, s>>>(b, njobs);
  return check_launch("fsg_zoom_minmax");
}

extern "C" int fsg_zoom(const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_zoom_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  int max_n2;
  if (int rc = check_zoom(jobs, njobs, sx, sy, sz, true, "fsg_zoom", &max_n2)) return rc;
  if (int rc = launch_zoom<false>(b, jobs, njobs, sx, sy, sz, max_n2, as_stream(stream), "fsg_zoom")) return rc;
  return check_launch("fsg_zoom");
}
