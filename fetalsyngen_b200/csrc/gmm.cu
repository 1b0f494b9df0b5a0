// K1 — GMM intensity synthesis fused with the seed sum and counter-based Philox noise.
//   L = sum_m seed[m][v];  I = max(0, mus[L] + sigmas[L] * N(v))
// Follows generator/intensity/rand_gmm.py:90-97 (sum) and :146-149 (lookup, noise, clamp).
// HBM-bound: reads 1 byte per seed volume per voxel (+4 B when noise is injected), writes 4 B.
// Each thread owns 4 consecutive voxels = one Philox block = one 16-byte store.
#include "common.cuh"

namespace fsg {

constexpr int GMM_THREADS = 256;
constexpr int GMM_MAX_LABELS = 256;

// NSEED = number of leading non-NULL seed pointers (the host entry compacts them).
template <bool INJECT, int NSEED>
__global__ void __launch_bounds__(GMM_THREADS) gmm_kernel(const __grid_constant__ Batch<fsg_gmm_job> batch, int64_t nvox) {
  const fsg_gmm_job& job = batch.j[blockIdx.y];
  __shared__ float2 s_ms[GMM_MAX_LABELS];  // (mu, sigma) per label
  for (int i = threadIdx.x; i < GMM_MAX_LABELS; i += GMM_THREADS) {
    const bool in = i < job.nlabels;
    s_ms[i] = make_float2(in ? job.mus[i] : 0.f, in ? job.sigmas[i] : 0.f);
  }
  __syncthreads();
  __shared__ float s_first[2][GMM_THREADS];  // first value of every thread's group (pairs output), double-buffered

  const int8_t* __restrict__ sp[4] = {job.seed[0], job.seed[1], job.seed[2], job.seed[3]};
  const float* __restrict__ noise = job.noise;
  float* __restrict__ out = job.out;
  uint8_t* __restrict__ lab_out = job.labels_out;
  const fsg_rng rng = job.rng;
  uint32_t* __restrict__ pairs = job.out_pairs;
  const int row_len = job.row_len;

  const int64_t ngroups = nvox >> 2;  // whole groups of 4 voxels; the tail is handled below
  const int64_t stride = (int64_t)gridDim.x * GMM_THREADS;
  int it = 0;
  for (int64_t gbase = (int64_t)blockIdx.x * GMM_THREADS; gbase < ngroups; gbase += stride, it ^= 1) {
    const int64_t g = gbase + threadIdx.x;
    const bool active = g < ngroups;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
    const int64_t v0 = g * 4;
    // labels of the seed volumes are disjoint small non-negative codes: one byte-wise SIMD add sums 4 voxels
    uint32_t lab4 = 0;
#pragma unroll
    for (int m = 0; m < NSEED; ++m) lab4 = __vadd4(lab4, __ldcs(reinterpret_cast<const uint32_t*>(sp[m] + v0)));
    float4 n;
    if (INJECT)
      n = __ldcs(reinterpret_cast<const float4*>(noise + v0));
    else
      n = philox_normal4(rng, (uint32_t)g);
    const float2 m0 = s_ms[lab4 & 0xff], m1 = s_ms[(lab4 >> 8) & 0xff], m2 = s_ms[(lab4 >> 16) & 0xff], m3 = s_ms[lab4 >> 24];
    o.x = fmaxf(add_rn(m0.x, mul_rn(m0.y, n.x)), 0.f);
    o.y = fmaxf(add_rn(m1.x, mul_rn(m1.y, n.y)), 0.f);
    o.z = fmaxf(add_rn(m2.x, mul_rn(m2.y, n.z)), 0.f);
    o.w = fmaxf(add_rn(m3.x, mul_rn(m3.y, n.w)), 0.f);
    if (out) *reinterpret_cast<float4*>(out + v0) = o;
    if (lab_out) *reinterpret_cast<uint32_t*>(lab_out + v0) = lab4;
    }
    if (pairs) {  // block-uniform
      // the pair of voxel v needs I[v+1]: the next thread's first value (shared memory), or — for the
      // block's last thread when its group does not end a row — one extra evaluation
      // (one barrier per iteration: the buffer written now is next written two iterations later)
      s_first[it][threadIdx.x] = o.x;
      __syncthreads();
      if (active) {
        const int64_t v0 = g * 4;
        float nxt = 0.f;
        const int64_t v4 = v0 + 4;
        if (v4 < nvox && (v4 % row_len) != 0) {
          if (threadIdx.x + 1 < GMM_THREADS) {
            nxt = s_first[it][threadIdx.x + 1];
          } else {
            int l = 0;
            for (int m = 0; m < NSEED; ++m) l += sp[m][v4];
            const float nz = INJECT ? noise[v4] : philox_normal4(rng, (uint32_t)(g + 1)).x;
            const float2 ms = s_ms[l & (GMM_MAX_LABELS - 1)];
            nxt = fmaxf(add_rn(ms.x, mul_rn(ms.y, nz)), 0.f);
          }
        }
        // round(I * 128) by the 2^23 magic add (round-to-nearest-even in the FMA); the low 16 bits of the
        // float's mantissa are the fixed-point value (I <= 511.99 after the clamp)
        const float cap = 511.9921875f;
        const uint32_t q0 = __float_as_uint(__fmaf_rn(fminf(o.x, cap), 128.0f, 8388608.0f)), q1 = __float_as_uint(__fmaf_rn(fminf(o.y, cap), 128.0f, 8388608.0f));
        const uint32_t q2 = __float_as_uint(__fmaf_rn(fminf(o.z, cap), 128.0f, 8388608.0f)), q3 = __float_as_uint(__fmaf_rn(fminf(o.w, cap), 128.0f, 8388608.0f));
        const uint32_t q4 = __float_as_uint(__fmaf_rn(fminf(nxt, cap), 128.0f, 8388608.0f));
        // pair = low halves of (q_k, q_{k+1})
        *reinterpret_cast<uint4*>(pairs + v0) = make_uint4(__byte_perm(q0, q1, 0x5410), __byte_perm(q1, q2, 0x5410), __byte_perm(q2, q3, 0x5410), __byte_perm(q3, q4, 0x5410));
      }
    }
  }
  // tail (nvox % 4 voxels): one thread, scalar
  if (blockIdx.x == 0 && threadIdx.x == 0 && (nvox & 3)) {
    const int64_t v0 = ngroups * 4;
    float nn[4] = {0.f, 0.f, 0.f, 0.f};
    if (!INJECT) {
      const float4 q = philox_normal4(rng, (uint32_t)ngroups);
      nn[0] = q.x; nn[1] = q.y; nn[2] = q.z; nn[3] = q.w;
    }
    for (int e = 0; v0 + e < nvox; ++e) {
      int l = 0;
      for (int m = 0; m < NSEED; ++m) l += sp[m][v0 + e];
      const float nz = INJECT ? noise[v0 + e] : nn[e];
      const float2 ms = s_ms[l & (GMM_MAX_LABELS - 1)];
      if (out) out[v0 + e] = fmaxf(add_rn(ms.x, mul_rn(ms.y, nz)), 0.f);
      if (lab_out) lab_out[v0 + e] = (uint8_t)l;
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

// Control grids drawn on the device (batched generation): out[i] = scale * N(0,1), one Philox
// stream per job.  Replaces torch.randn on the host + an upload per sample
// (affine_nonrigid.py:312-316 Fsmall = nonlin_std * randn, synthseg.py:170-172 bf = bf_std * randn).
__global__ void __launch_bounds__(256) grid_draw_kernel(const __grid_constant__ Batch<fsg_grid_job> batch) {
  const fsg_grid_job& job = batch.j[blockIdx.y];
  const int ngroups = (job.n + 3) / 4;
  for (int g = blockIdx.x * 256 + threadIdx.x; g < ngroups; g += gridDim.x * 256) {
    const float4 q = philox_normal4(job.rng, (uint32_t)g);
    const float v[4] = {q.x, q.y, q.z, q.w};
    for (int e = 0; e < 4 && g * 4 + e < job.n; ++e) job.out[g * 4 + e] = __fmul_rn(job.scale, v[e]);
  }
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_draw_grids(const fsg_grid_job* jobs, int njobs, void* stream) {
  Batch<fsg_grid_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  int nmax = 0;
  for (int n = 0; n < njobs; ++n) {
    FSG_REQUIRE(jobs[n].out && jobs[n].n >= 1, "fsg_draw_grids: job %d has no output", n);
    nmax = jobs[n].n > nmax ? jobs[n].n : nmax;
  }
  grid_draw_kernel<<<dim3(((nmax + 3) / 4 + 255) / 256, njobs), 256, 0, as_stream(stream)>>>(b);
  return check_launch("fsg_draw_grids");
}

template <bool INJECT>
static void launch_gmm(const Batch<fsg_gmm_job>& b, int nseed, dim3 grid, int64_t nvox, cudaStream_t s) {
  switch (nseed) {
    case 1: gmm_kernel<INJECT, 1><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    case 2: gmm_kernel<INJECT, 2><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    case 3: gmm_kernel<INJECT, 3><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    default: gmm_kernel<INJECT, 4><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
  }
}

extern "C" int fsg_gmm(const fsg_gmm_job* jobs, int njobs, int64_t nvox, void* stream) {
  Batch<fsg_gmm_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  FSG_REQUIRE(nvox > 0, "fsg_gmm: nvox must be positive");
  FSG_REQUIRE(nvox / 4 < (int64_t)1 << 32, "fsg_gmm: volume too large for the 32-bit Philox block counter");
  bool inject = jobs[0].noise != nullptr;
  int nseed = -1;
  for (int i = 0; i < njobs; ++i) {
    const fsg_gmm_job& j = jobs[i];
    FSG_REQUIRE((j.out || j.out_pairs) && j.mus && j.sigmas, "fsg_gmm: job %d has a NULL out/mus/sigmas", i);
    if (j.out_pairs)
      FSG_REQUIRE(j.row_len >= 1 && nvox % j.row_len == 0 && nvox % 4 == 0 && (reinterpret_cast<uintptr_t>(j.out_pairs) & 15) == 0,
                  "fsg_gmm: job %d: out_pairs needs row_len dividing nvox, nvox %% 4 == 0 and a 16-byte aligned buffer", i);
    FSG_REQUIRE(j.nlabels >= 1 && j.nlabels <= GMM_MAX_LABELS, "fsg_gmm: nlabels=%d outside [1,%d]", j.nlabels, GMM_MAX_LABELS);
    FSG_REQUIRE((j.noise != nullptr) == inject, "fsg_gmm: jobs mix injected and Philox noise");
    // compact the seed pointers of the launch copy to the front
    int n = 0;
    for (int m = 0; m < 4; ++m)
      if (j.seed[m]) b.j[i].seed[n++] = j.seed[m];
    for (int m = n; m < 4; ++m) b.j[i].seed[m] = nullptr;
    FSG_REQUIRE(n >= 1, "fsg_gmm: job %d has no seed volume", i);
    FSG_REQUIRE(nseed < 0 || nseed == n, "fsg_gmm: jobs of one launch must have the same number of seed volumes");
    nseed = n;
    for (int m = 0; m < 4; ++m) FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.seed[m]) & 3) == 0, "fsg_gmm: seed pointers must be 4-byte aligned");
    FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.noise) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.labels_out) & 3) == 0,
                "fsg_gmm: out/noise must be 16-byte aligned");
  }
  const int64_t ngroups = (nvox + 3) / 4;
  int64_t want = (ngroups + GMM_THREADS - 1) / GMM_THREADS;
  const int64_t cap = 148 * 16;
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)njobs);
  if (inject)
    launch_gmm<true>(b, nseed, grid, nvox, as_stream(stream));
  else
    launch_gmm<false>(b, nseed, grid, nvox, as_stream(stream));
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
