// K4c — separable linear zoom (myzoom_torch, utils/generation.py:310-397) with the global
// reductions that follow it on this path: /max of RandResample.resize_back
// (augmentation/synthseg.py:109-114) and ScaleIntensity(0,1) (data/datasets.py:311).
//
// A warp produces one output row (i, j, :).  It first blends the four coarse rows
// (fx|cx, fy|cy) along x and then y into a shared-memory row of n2 values (coalesced loads,
// the reference's rounding order: w_f*X[f] + w_c*X[c] per axis), then every lane blends along
// z from shared memory.  ~20 instructions per output voxel instead of 8 scattered gathers and
// 64-bit index arithmetic (round-1 ncu: 228 instr/voxel, issue-bound).
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

// KG > 0: sz == 128*KG and every lane owns the same KG groups of 4 consecutive z positions for
// all rows, so its z-table entries live in registers and the stores are 16-byte vectors.
// KG == 0: generic extents, z table read from shared memory per voxel.
template <bool REDUCE, int KG>
__global__ void __launch_bounds__(ZM_THREADS, 3) zoom_rows_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int sx, int sy, int sz) {
  const fsg_zoom_job& job = batch.j[blockIdx.y];
  const int n1 = job.n[1], n2 = job.n[2];
  extern __shared__ float s_zoom[];
  // layout: z table [sz] as (int f | c<<16, float wc), then one coarse row per warp
  int2* s_tz = reinterpret_cast<int2*>(s_zoom);
  float* s_row = s_zoom + 2 * sz + (threadIdx.x >> 5) * n2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NK = KG > 0 ? KG * 4 : 1;
  int tfc[NK];  // f | c << 16
  float twc[NK], twf[NK];
  if (KG > 0) {
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      const fsg_tab e = job.tab[2][128 * (q >> 2) + 4 * lane + (q & 3)];
      tfc[q] = (int)e.f | ((int)e.c << 16);
      twc[q] = e.wc;
      twf[q] = sub_rn(1.0f, e.wc);
    }
  } else {
    for (int k = threadIdx.x; k < sz; k += ZM_THREADS) {
      const fsg_tab e = job.tab[2][k];
      s_tz[k] = make_int2((int)e.f | ((int)e.c << 16), __float_as_int(e.wc));
    }
    __syncthreads();
  }

  const float* __restrict__ src = job.src;
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
  const int nrows = sx * sy;
  for (int row = blockIdx.x * ZM_WARPS + warp; row < nrows; row += gridDim.x * ZM_WARPS) {
    const int i = row / sy, j = row - i * sy;
    const Tab tx = load_tab(job.tab[0], i), ty = load_tab(job.tab[1], j);
    const float* pff = src + ((size_t)tx.f * n1 + ty.f) * n2;
    const float* pcf = src + ((size_t)tx.c * n1 + ty.f) * n2;
    const float* pfc = src + ((size_t)tx.f * n1 + ty.c) * n2;
    const float* pcc = src + ((size_t)tx.c * n1 + ty.c) * n2;
    __syncwarp();  // previous row's readers are done
    for (int K0 = lane; K0 < n2; K0 += 128) {
      // four K positions per lane with all 16 loads issued before the first blend
      float vff[4], vcf[4], vfc[4], vcc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int K = K0 + 32 * u;
        const bool in = K < n2;
        vff[u] = in ? __ldg(pff + K) : 0.f;
        vcf[u] = in ? __ldg(pcf + K) : 0.f;
        vfc[u] = in ? __ldg(pfc + K) : 0.f;
        vcc[u] = in ? __ldg(pcc + K) : 0.f;
      }
      // image path (tolerance, not bit parity): the x / y blends run as packed FP32x2 FMAs on two K
      // positions at a time
      const P2 wfx = pk(tx.wf, tx.wf), wcx = pk(tx.wc, tx.wc), wfy = pk(ty.wf, ty.wf), wcy = pk(ty.wc, ty.wc);
#pragma unroll
      for (int u = 0; u < 4; u += 2) {
        const P2 a_f = fma2(wfx, pk(vff[u], vff[u + 1]), mul2(wcx, pk(vcf[u], vcf[u + 1])));  // tmp1[y=f]
        const P2 a_c = fma2(wfx, pk(vfc[u], vfc[u + 1]), mul2(wcx, pk(vcc[u], vcc[u + 1])));  // tmp1[y=c]
        float r0, r1;
        upk(fma2(wfy, a_f, mul2(wcy, a_c)), r0, r1);                                           // tmp2
        if (K0 + 32 * u < n2) s_row[K0 + 32 * u] = r0;
        if (K0 + 32 * (u + 1) < n2) s_row[K0 + 32 * (u + 1)] = r1;
      }
    }
    __syncwarp();
    if (KG > 0) {
#pragma unroll
      for (int m = 0; m < (KG > 0 ? KG : 1); ++m) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int q = 4 * m + e;
          const P2 lo = pk(s_row[tfc[q] & 0xffff], s_row[tfc[q + 1] & 0xffff]), hi = pk(s_row[tfc[q] >> 16], s_row[tfc[q + 1] >> 16]);
          upk(fma2(pk(twf[q], twf[q + 1]), lo, mul2(pk(twc[q], twc[q + 1]), hi)), v[e], v[e + 1]);
        }
        if (REDUCE) {
          lo_ = fminf(fminf(lo_, v[0]), fminf(fminf(v[1], v[2]), v[3]));
          hi_ = fmaxf(fmaxf(hi_, v[0]), fmaxf(fmaxf(v[1], v[2]), v[3]));
        } else {
          if (post == 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = div_nr(v[e], vmax, rmax);
          }
          if (post == 2) {
            const P2 pa2 = pk(pa, pa), pb2 = pk(pb, pb);
            float n0, n1, n2_, n3;
            upk(fma2(pk(v[0], v[1]), pa2, pb2), n0, n1);
            upk(fma2(pk(v[2], v[3]), pa2, pb2), n2_, n3);
            v[0] = v[0] == vmax_eq ? 1.f : fminf(fmaxf(n0, 0.f), 1.f);  // the maximum maps to exactly 1 (ScaleIntensity)
            v[1] = v[1] == vmax_eq ? 1.f : fminf(fmaxf(n1, 0.f), 1.f);
            v[2] = v[2] == vmax_eq ? 1.f : fminf(fmaxf(n2_, 0.f), 1.f);
            v[3] = v[3] == vmax_eq ? 1.f : fminf(fmaxf(n3, 0.f), 1.f);
          }
          __stcs(reinterpret_cast<float4*>(job.dst + (size_t)row * sz + 128 * m + 4 * lane), make_float4(v[0], v[1], v[2], v[3]));
        }
      }
    } else {
      float* __restrict__ out = REDUCE ? nullptr : job.dst + (size_t)row * sz;
      for (int k = lane; k < sz; k += 32) {
        const int2 e = s_tz[k];
        const float wc = __int_as_float(e.y), wf = sub_rn(1.0f, wc);
        float val = blend(wf, s_row[e.x & 0xffff], wc, s_row[e.x >> 16]);
        if (REDUCE) {
          lo_ = fminf(lo_, val);
          hi_ = fmaxf(hi_, val);
        } else {
          if (post == 1) val = div_nr(val, vmax, rmax);
          if (post == 2) val = val == vmax_eq ? 1.f : fminf(fmaxf(__fmaf_rn(val, pa, pb), 0.f), 1.f);
          out[k] = val;
        }
      }
    }
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

template <bool REDUCE, int KG>
static void launch_zoom_kg(const Batch<fsg_zoom_job>& b, int njobs, int sx, int sy, int sz, size_t smem, cudaStream_t s) {
  auto k = zoom_rows_kernel<REDUCE, KG>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t want = ((int64_t)sx * sy + ZM_WARPS - 1) / ZM_WARPS;
  const int cap = 148 * 8;
  k<<<dim3((unsigned)(want < cap ? want : cap), njobs), ZM_THREADS, smem, s>>>(b, sx, sy, sz);
}

template <bool REDUCE>
static int launch_zoom(const Batch<fsg_zoom_job>& b, const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, int max_n2, cudaStream_t s, const char* who) {
  const size_t smem = ((size_t)2 * sz + (size_t)ZM_WARPS * max_n2) * sizeof(float);
  FSG_REQUIRE(smem <= 200 * 1024, "%s: rows of %d / %d voxels do not fit in shared memory", who, sz, max_n2);
  bool vec = (sz % 128 == 0) && sz <= 512;
  for (int i = 0; i < njobs && vec && !REDUCE; ++i) vec = (reinterpret_cast<uintptr_t>(jobs[i].dst) & 15) == 0;
  const int kg = vec ? sz / 128 : 0;
  switch (kg) {
    case 1: launch_zoom_kg<REDUCE, 1>(b, njobs, sx, sy, sz, smem, s); break;
    case 2: launch_zoom_kg<REDUCE, 2>(b, njobs, sx, sy, sz, smem, s); break;
    case 3: launch_zoom_kg<REDUCE, 3>(b, njobs, sx, sy, sz, smem, s); break;
    case 4: launch_zoom_kg<REDUCE, 4>(b, njobs, sx, sy, sz, smem, s); break;
    default: launch_zoom_kg<REDUCE, 0>(b, njobs, sx, sy, sz, smem, s); break;
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
