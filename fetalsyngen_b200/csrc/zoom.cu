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

template <bool REDUCE>
__global__ void __launch_bounds__(ZM_THREADS) zoom_rows_kernel(const __grid_constant__ Batch<fsg_zoom_job> batch, int sx, int sy, int sz) {
  const fsg_zoom_job& job = batch.j[blockIdx.y];
  const int n1 = job.n[1], n2 = job.n[2];
  extern __shared__ float s_zoom[];
  // layout: z table [sz] as (int f | c<<16, float wc), then one coarse row per warp
  int2* s_tz = reinterpret_cast<int2*>(s_zoom);
  float* s_row = s_zoom + 2 * sz + (threadIdx.x >> 5) * n2;
  for (int k = threadIdx.x; k < sz; k += ZM_THREADS) {
    const fsg_tab e = job.tab[2][k];
    s_tz[k] = make_int2((int)e.f | ((int)e.c << 16), __float_as_int(e.wc));
  }
  __syncthreads();

  const float* __restrict__ src = job.src;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inf = __int_as_float(0x7f800000);
  float lo = inf, hi = -inf;
  float vmax = 1.f, rmax = 1.f, qmin = 0.f, den = 1.f, rden = 1.f;
  const int post = job.post;
  if (!REDUCE && post > 0) {
    vmax = job.minmax[1];
    rmax = __frcp_rn(vmax);
    qmin = div_nr(job.minmax[0], vmax, rmax);
    den = sub_rn(div_nr(vmax, vmax, rmax), qmin);
    rden = __frcp_rn(den);
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
    for (int K = lane; K < n2; K += 32) {
      const float a_f = blend(tx.wf, __ldg(pff + K), tx.wc, __ldg(pcf + K));  // tmp1[y=f]
      const float a_c = blend(tx.wf, __ldg(pfc + K), tx.wc, __ldg(pcc + K));  // tmp1[y=c]
      s_row[K] = blend(ty.wf, a_f, ty.wc, a_c);                               // tmp2
    }
    __syncwarp();
    float* __restrict__ out = REDUCE ? nullptr : job.dst + (size_t)row * sz;
    for (int k = lane; k < sz; k += 32) {
      const int2 e = s_tz[k];
      const float wc = __int_as_float(e.y), wf = sub_rn(1.0f, wc);
      float val = blend(wf, s_row[e.x & 0xffff], wc, s_row[e.x >> 16]);
      if (REDUCE) {
        lo = fminf(lo, val);
        hi = fmaxf(hi, val);
      } else {
        if (post >= 1) val = div_nr(val, vmax, rmax);
        if (post >= 2) val = div_nr(sub_rn(val, qmin), den, rden);
        out[k] = val;
      }
    }
  }
  if (REDUCE) {
    lo = warp_min(lo);
    hi = warp_max(hi);
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

template <bool REDUCE>
static int launch_zoom(const Batch<fsg_zoom_job>& b, int njobs, int sx, int sy, int sz, int max_n2, cudaStream_t s, const char* who) {
  const size_t smem = ((size_t)2 * sz + (size_t)ZM_WARPS * max_n2) * sizeof(float);
  FSG_REQUIRE(smem <= 200 * 1024, "%s: rows of %d / %d voxels do not fit in shared memory", who, sz, max_n2);
  auto k = zoom_rows_kernel<REDUCE>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t want = ((int64_t)sx * sy + ZM_WARPS - 1) / ZM_WARPS;
  const int cap = 148 * 8;
  k<<<dim3((unsigned)(want < cap ? want : cap), njobs), ZM_THREADS, smem, s>>>(b, sx, sy, sz);
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
  if (int rc = launch_zoom<true>(b, njobs, sx, sy, sz, max_n2, s, "fsg_zoom_minmax")) return rc;
  zoom_mm_final_kernel<<<1, 32, 0, s>>>(b, njobs);
  return check_launch("fsg_zoom_minmax");
}

extern "C" int fsg_zoom(const fsg_zoom_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_zoom_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  int max_n2;
  if (int rc = check_zoom(jobs, njobs, sx, sy, sz, true, "fsg_zoom", &max_n2)) return rc;
  if (int rc = launch_zoom<false>(b, njobs, sx, sy, sz, max_n2, as_stream(stream), "fsg_zoom")) return rc;
  return check_launch("fsg_zoom");
}
