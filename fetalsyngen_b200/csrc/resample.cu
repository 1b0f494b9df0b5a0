// K4b — regular-grid resampling kernels (stand-alone stage API; the fused production path is
// sepconv.cu).
//  fsg_resample : trilinear down-sampling onto the coarse grid + additive noise, bit-exact
//                 arithmetic order of the reference
//                 (augmentation/synthseg.py:84-107 -> utils/generation.py:227-285; :217-235)
//  fsg_add_noise: elementwise noise at any resolution.
#include "common.cuh"

namespace fsg {

constexpr int RS_THREADS = 256;

template <bool INJECT>
__global__ void __launch_bounds__(RS_THREADS) resample_kernel(const __grid_constant__ Batch<fsg_resample_job> batch, int sx, int sy, int sz) {
  const fsg_resample_job& job = batch.j[blockIdx.y];
  const int n0 = job.n[0], n1 = job.n[1], n2 = job.n[2];
  const float* __restrict__ src = job.src;
  const int64_t n = (int64_t)n0 * n1 * n2;
  const int64_t ngroups = (n + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
  for (int64_t g = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; g < ngroups; g += stride) {
    float nz[4] = {0.f, 0.f, 0.f, 0.f};
    if (job.has_noise) {
      if (INJECT) {
        for (int e = 0; e < 4; ++e)
          if (g * 4 + e < n) nz[e] = job.noise[g * 4 + e];
      } else {
        const float4 q = philox_normal4(job.rng, (uint32_t)g);
        nz[0] = q.x; nz[1] = q.y; nz[2] = q.z; nz[3] = q.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t v = g * 4 + e;
      if (v >= n) break;
      const unsigned vv = (unsigned)v, q = vv / (unsigned)n2;
      const int k = (int)(vv - q * (unsigned)n2);
      const int i = (int)(q / (unsigned)n1), j = (int)(q - (unsigned)i * (unsigned)n1);
      const fsg_tab ex = job.tab[0][i], ey = job.tab[1][j], ez = job.tab[2][k];
      float val = 0.f;
      if (ex.f >= 0 && ey.f >= 0 && ez.f >= 0) {
        const float wcx = ex.wc, wcy = ey.wc, wcz = ez.wc;
        const float wfx = sub_rn(1.f, wcx), wfy = sub_rn(1.f, wcy), wfz = sub_rn(1.f, wcz);
        const float* pff = src + ((size_t)ex.f * sy + ey.f) * sz;
        const float* pcf = src + ((size_t)ex.c * sy + ey.f) * sz;
        const float* pfc = src + ((size_t)ex.f * sy + ey.c) * sz;
        const float* pcc = src + ((size_t)ex.c * sy + ey.c) * sz;
        const float c00 = lerp2(__ldg(pff + ez.f), wfx, __ldg(pcf + ez.f), wcx);
        const float c01 = lerp2(__ldg(pff + ez.c), wfx, __ldg(pcf + ez.c), wcx);
        const float c10 = lerp2(__ldg(pfc + ez.f), wfx, __ldg(pcc + ez.f), wcx);
        const float c11 = lerp2(__ldg(pfc + ez.c), wfx, __ldg(pcc + ez.c), wcx);
        const float c0 = lerp2(c00, wfy, c10, wcy);
        const float c1 = lerp2(c01, wfy, c11, wcy);
        val = lerp2(c0, wfz, c1, wcz);
      }
      if (job.has_noise) {
        val = add_rn(val, mul_rn(job.noise_std, nz[e]));
        val = val < 0.f ? 0.f : val;
      }
      job.dst[v] = val;
    }
  }
}

template <bool INJECT>
__global__ void __launch_bounds__(RS_THREADS) noise_kernel(const __grid_constant__ Batch<fsg_noise_job> batch, int64_t n) {
  const fsg_noise_job& job = batch.j[blockIdx.y];
  const int64_t ngroups = (n + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
  for (int64_t g = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; g < ngroups; g += stride) {
    float nz[4];
    if (INJECT) {
      for (int e = 0; e < 4; ++e) nz[e] = (g * 4 + e < n) ? job.noise[g * 4 + e] : 0.f;
    } else {
      const float4 q = philox_normal4(job.rng, (uint32_t)g);
      nz[0] = q.x; nz[1] = q.y; nz[2] = q.z; nz[3] = q.w;
    }
    for (int e = 0; e < 4 && g * 4 + e < n; ++e) {
      const float v = add_rn(job.src[g * 4 + e], mul_rn(job.noise_std, nz[e]));
      job.dst[g * 4 + e] = (v < 0.f && !(job.flags & 1)) ? 0.f : v;
    }
  }
}

static unsigned grid_x(int64_t work) {
  const int64_t want = (work + RS_THREADS - 1) / RS_THREADS;
  return (unsigned)(want < 148 * 32 ? (want < 1 ? 1 : want) : 148 * 32);
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_resample(const fsg_resample_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  Batch<fsg_resample_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1, "fsg_resample: bad shape");
  int64_t n = 0;
  FSG_REQUIRE(sx <= 32767 && sy <= 32767 && sz <= 32767, "fsg_resample: source extent exceeds the int16 table range");
  bool inject = false;
  for (int i = 0; i < njobs; ++i) {
    const fsg_resample_job& j = jobs[i];
    FSG_REQUIRE(j.src && j.dst && j.tab[0] && j.tab[1] && j.tab[2], "fsg_resample: job %d has a NULL pointer", i);
    FSG_REQUIRE(j.n[0] >= 1 && j.n[1] >= 1 && j.n[2] >= 1, "fsg_resample: job %d bad coarse shape", i);
    const int64_t nj = (int64_t)j.n[0] * j.n[1] * j.n[2];
    n = nj > n ? nj : n;
    if (j.has_noise && j.noise) inject = true;
  }
  for (int i = 0; i < njobs; ++i) FSG_REQUIRE(!jobs[i].has_noise || ((jobs[i].noise != nullptr) == inject), "fsg_resample: jobs mix injected and Philox noise");
  FSG_REQUIRE(n < ((int64_t)1 << 31), "fsg_resample: volume too large");
  dim3 grid(grid_x((n + 3) / 4), (unsigned)njobs);
  if (inject)
    resample_kernel<true><<<grid, RS_THREADS, 0, as_stream(stream)>>>(b, sx, sy, sz);
  else
    resample_kernel<false><<<grid, RS_THREADS, 0, as_stream(stream)>>>(b, sx, sy, sz);
  return check_launch("fsg_resample");
}

extern "C" int fsg_add_noise(const fsg_noise_job* jobs, int njobs, int64_t nvox, void* stream) {
  Batch<fsg_noise_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  FSG_REQUIRE(nvox > 0 && nvox / 4 < (int64_t)1 << 32, "fsg_add_noise: bad size");
  const bool inject = jobs[0].noise != nullptr;
  for (int i = 0; i < njobs; ++i) {
    FSG_REQUIRE(jobs[i].src && jobs[i].dst, "fsg_add_noise: job %d has a NULL pointer", i);
    FSG_REQUIRE((jobs[i].noise != nullptr) == inject, "fsg_add_noise: jobs mix injected and Philox noise");
  }
  dim3 grid(grid_x((nvox + 3) / 4), (unsigned)njobs);
  if (inject)
    noise_kernel<true><<<grid, RS_THREADS, 0, as_stream(stream)>>>(b, nvox);
  else
    noise_kernel<false><<<grid, RS_THREADS, 0, as_stream(stream)>>>(b, nvox);
  return check_launch("fsg_add_noise");
}

