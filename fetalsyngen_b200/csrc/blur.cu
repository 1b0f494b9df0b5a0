// K4a (stand-alone stage API) — separable zero-padded Gaussian blur in the reference's exact
// arithmetic order (utils/generation.py:84-110): one pass per axis, x then y then z like the
// three conv3d calls, taps accumulated in index order without FMA.  Taps are built on the host
// with the reference's expressions (make_gaussian_kernel, utils/generation.py:74-81).
// The fused production path (blur composed with the down-sampling) is sepconv.cu.
#include "common.cuh"

namespace fsg {

constexpr int BLUR_THREADS = 256;

struct BlurPass {
  const float* src[FSG_MAX_JOBS];
  float* dst[FSG_MAX_JOBS];
  const float* taps[FSG_MAX_JOBS];
  int ntaps[FSG_MAX_JOBS];
};

// Accumulates taps in index order, each product and sum rounded (no FMA).
template <int AXIS>
__global__ void __launch_bounds__(BLUR_THREADS) blur_axis_kernel(const __grid_constant__ BlurPass p, int sx, int sy, int sz) {
  const int jb = blockIdx.y;
  const int nt = p.ntaps[jb];
  __shared__ float s_w[FSG_MAX_TAPS];
  for (int t = threadIdx.x; t < nt; t += BLUR_THREADS) s_w[t] = p.taps[jb][t];
  __syncthreads();
  const float* __restrict__ src = p.src[jb];
  float* __restrict__ dst = p.dst[jb];
  const int r = nt / 2;
  const unsigned n = (unsigned)sx * sy * sz;  // < 2^31 (checked by the caller)
  const unsigned stride = gridDim.x * BLUR_THREADS;
  const int astride = AXIS == 0 ? sy * sz : (AXIS == 1 ? sz : 1);
  const int alen = AXIS == 0 ? sx : (AXIS == 1 ? sy : sz);
  for (unsigned v = blockIdx.x * BLUR_THREADS + threadIdx.x; v < n; v += stride) {
    int pos;
    if (AXIS == 2)
      pos = (int)(v % (unsigned)sz);
    else if (AXIS == 1)
      pos = (int)((v / (unsigned)sz) % (unsigned)sy);
    else
      pos = (int)(v / ((unsigned)sy * sz));
    float acc = 0.f;
    const int t0 = max(0, r - pos), t1 = min(nt, alen + r - pos);
    const float* __restrict__ q = src + (int)v - r * astride;
    for (int t = t0; t < t1; ++t) acc = add_rn(acc, mul_rn(s_w[t], __ldg(q + t * astride)));
    dst[v] = acc;
  }
}

__global__ void __launch_bounds__(256) copy_kernel(const float* __restrict__ a, float* __restrict__ b, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_blur3d(const fsg_blur_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  FSG_REQUIRE(jobs != nullptr, "fsg_blur3d: jobs pointer is NULL");
  FSG_REQUIRE(njobs >= 1 && njobs <= FSG_MAX_JOBS, "fsg_blur3d: njobs=%d outside [1,%d]", njobs, FSG_MAX_JOBS);
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1 && (int64_t)sx * sy * sz < ((int64_t)1 << 31), "fsg_blur3d: bad shape");
  cudaStream_t s = as_stream(stream);
  const int64_t n = (int64_t)sx * sy * sz;
  const int64_t want = (n + BLUR_THREADS - 1) / BLUR_THREADS;
  const unsigned gx = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  // Jobs may blur different subsets of axes; each job ping-pongs src -> {tmp,dst} so that its
  // last active pass lands in dst.  Launch per axis over the jobs active on that axis.
  int active[FSG_MAX_JOBS], done[FSG_MAX_JOBS];
  for (int i = 0; i < njobs; ++i) {
    const fsg_blur_job& j = jobs[i];
    FSG_REQUIRE(j.src && j.dst, "fsg_blur3d: job %d has NULL src/dst", i);
    active[i] = 0;
    done[i] = 0;
    for (int a = 0; a < 3; ++a) {
      const bool on = j.taps[a] != nullptr && j.ntaps[a] > 0;
      FSG_REQUIRE(!on || (j.ntaps[a] % 2 == 1 && j.ntaps[a] <= FSG_MAX_TAPS), "fsg_blur3d: job %d axis %d needs an odd tap count <= %d", i, a, FSG_MAX_TAPS);
      active[i] += on;
    }
    FSG_REQUIRE(active[i] < 2 || j.tmp, "fsg_blur3d: job %d needs a tmp volume", i);
    FSG_REQUIRE(active[i] == 0 || j.src != j.dst, "fsg_blur3d: job %d cannot blur in place", i);
    if (active[i] == 0 && j.src != j.dst) copy_kernel<<<gx, 256, 0, s>>>(j.src, j.dst, n);
  }
  for (int a = 0; a < 3; ++a) {
    BlurPass p;
    memset(&p, 0, sizeof(p));
    int m = 0;
    for (int i = 0; i < njobs; ++i) {
      const fsg_blur_job& j = jobs[i];
      if (!(j.taps[a] != nullptr && j.ntaps[a] > 0)) continue;
      const int left = active[i] - done[i];  // passes left including this one
      // buffers: pass lands in dst when `left` is odd, in tmp when even
      float* out = (left % 2 == 1) ? j.dst : j.tmp;
      const float* in = done[i] == 0 ? j.src : ((left % 2 == 1) ? j.tmp : j.dst);
      p.src[m] = in;
      p.dst[m] = out;
      p.taps[m] = j.taps[a];
      p.ntaps[m] = j.ntaps[a];
      ++m;
      ++done[i];
    }
    if (m == 0) continue;
    dim3 grid(gx, (unsigned)m);
    if (a == 0)
      blur_axis_kernel<0><<<grid, BLUR_THREADS, 0, s>>>(p, sx, sy, sz);
    else if (a == 1)
      blur_axis_kernel<1><<<grid, BLUR_THREADS, 0, s>>>(p, sx, sy, sz);
    else
      blur_axis_kernel<2><<<grid, BLUR_THREADS, 0, s>>>(p, sx, sy, sz);
  }
  return check_launch("fsg_blur3d");
}
