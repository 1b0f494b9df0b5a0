// K1 — GMM intensity synthesis fused with the seed sum and counter-based Philox noise.
//   L = sum_m seed[m][v];  I = max(0, mus[L] + sigmas[L] * N(v))
// Follows generator/intensity/rand_gmm.py:90-97 (sum) and :146-149 (lookup, noise, clamp).
// HBM-bound: reads 1 byte per seed volume per voxel (+4 B when noise is injected), writes 4 B.
// Each thread owns 4 consecutive voxels = one Philox block = one 16-byte store.
#include "common.cuh"

namespace fsg {

constexpr int GMM_THREADS = 256;
constexpr int GMM_MAX_LABELS = 256;

template <bool INJECT>
__global__ void __launch_bounds__(GMM_THREADS) gmm_kernel(const __grid_constant__ Batch<fsg_gmm_job> batch, int64_t nvox) {
  const fsg_gmm_job& job = batch.j[blockIdx.y];
  __shared__ float s_mu[GMM_MAX_LABELS], s_sg[GMM_MAX_LABELS];
  for (int i = threadIdx.x; i < GMM_MAX_LABELS; i += GMM_THREADS) {
    const bool in = i < job.nlabels;
    s_mu[i] = in ? job.mus[i] : 0.f;
    s_sg[i] = in ? job.sigmas[i] : 0.f;
  }
  __syncthreads();

  const int64_t ngroups = (nvox + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * GMM_THREADS;
  for (int64_t g = (int64_t)blockIdx.x * GMM_THREADS + threadIdx.x; g < ngroups; g += stride) {
    const int64_t v0 = g * 4;
    const bool full = v0 + 4 <= nvox;
    int lab[4] = {0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int8_t* sp = job.seed[m];
      if (sp == nullptr) continue;
      if (full) {
        const char4 c = __ldg(reinterpret_cast<const char4*>(sp + v0));
        lab[0] += c.x; lab[1] += c.y; lab[2] += c.z; lab[3] += c.w;
      } else {
        for (int e = 0; e < 4 && v0 + e < nvox; ++e) lab[e] += sp[v0 + e];
      }
    }
    float n[4];
    if (INJECT) {
      if (full) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(job.noise + v0));
        n[0] = q.x; n[1] = q.y; n[2] = q.z; n[3] = q.w;
      } else {
        for (int e = 0; e < 4; ++e) n[e] = (v0 + e < nvox) ? job.noise[v0 + e] : 0.f;
      }
    } else {
      const float4 q = philox_normal4(job.rng, (uint32_t)g);
      n[0] = q.x; n[1] = q.y; n[2] = q.z; n[3] = q.w;
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int l = lab[e] & (GMM_MAX_LABELS - 1);
      const float v = add_rn(s_mu[l], mul_rn(s_sg[l], n[e]));
      o[e] = v < 0.f ? 0.f : v;
    }
    if (full) {
      *reinterpret_cast<float4*>(job.out + v0) = make_float4(o[0], o[1], o[2], o[3]);
      if (job.labels_out) *reinterpret_cast<uchar4*>(job.labels_out + v0) = make_uchar4(lab[0], lab[1], lab[2], lab[3]);
    } else {
      for (int e = 0; e < 4 && v0 + e < nvox; ++e) {
        job.out[v0 + e] = o[e];
        if (job.labels_out) job.labels_out[v0 + e] = (uint8_t)lab[e];
      }
    }
  }
}

template <bool RAW>
__global__ void __launch_bounds__(256) philox_fill_kernel(fsg_rng rng, float* out, int64_t n) {
  const int64_t ngroups = (n + 3) / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    float v[4];
    if (RAW) {
      const Philox ph(rng.seed);
      const uint4 w = ph((uint32_t)g, rng.stage, (uint32_t)rng.sample, (uint32_t)(rng.sample >> 32));
      v[0] = __uint_as_float(w.x); v[1] = __uint_as_float(w.y); v[2] = __uint_as_float(w.z); v[3] = __uint_as_float(w.w);
    } else {
      const float4 q = philox_normal4(rng, (uint32_t)g);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
    for (int e = 0; e < 4 && g * 4 + e < n; ++e) out[g * 4 + e] = v[e];
  }
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_gmm(const fsg_gmm_job* jobs, int njobs, int64_t nvox, void* stream) {
  Batch<fsg_gmm_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  FSG_REQUIRE(nvox > 0, "fsg_gmm: nvox must be positive");
  FSG_REQUIRE(nvox / 4 < (int64_t)1 << 32, "fsg_gmm: volume too large for the 32-bit Philox block counter");
  bool inject = jobs[0].noise != nullptr;
  for (int i = 0; i < njobs; ++i) {
    const fsg_gmm_job& j = jobs[i];
    FSG_REQUIRE(j.out && j.mus && j.sigmas, "fsg_gmm: job %d has a NULL out/mus/sigmas", i);
    FSG_REQUIRE(j.seed[0] || j.seed[1] || j.seed[2] || j.seed[3], "fsg_gmm: job %d has no seed volume", i);
    FSG_REQUIRE(j.nlabels >= 1 && j.nlabels <= GMM_MAX_LABELS, "fsg_gmm: nlabels=%d outside [1,%d]", j.nlabels, GMM_MAX_LABELS);
    FSG_REQUIRE((j.noise != nullptr) == inject, "fsg_gmm: jobs mix injected and Philox noise");
    for (int m = 0; m < 4; ++m) FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.seed[m]) & 3) == 0, "fsg_gmm: seed pointers must be 4-byte aligned");
    FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.noise) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.labels_out) & 3) == 0,
                "fsg_gmm: out/noise must be 16-byte aligned");
  }
  const int64_t ngroups = (nvox + 3) / 4;
  int64_t want = (ngroups + GMM_THREADS - 1) / GMM_THREADS;
  const int64_t cap = 148 * 32;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)njobs);
  if (inject)
    gmm_kernel<true><<<grid, GMM_THREADS, 0, as_stream(stream)>>>(b, nvox);
  else
    gmm_kernel<false><<<grid, GMM_THREADS, 0, as_stream(stream)>>>(b, nvox);
  return check_launch("fsg_gmm");
}

extern "C" int fsg_philox_fill(fsg_rng rng, float* out, int64_t n, int raw, void* stream) {
  FSG_REQUIRE(out && n > 0, "fsg_philox_fill: bad arguments");
  const int64_t want = ((n + 3) / 4 + 255) / 256;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  if (raw)
    philox_fill_kernel<true><<<blocks, 256, 0, as_stream(stream)>>>(rng, out, n);
  else
    philox_fill_kernel<false><<<blocks, 256, 0, as_stream(stream)>>>(rng, out, n);
  return check_launch("fsg_philox_fill");
}
