// K1 — GMM intensity synthesis fused with the seed sum and counter-based Philox noise.
//   L = sum_m seed[m][v];  I = max(0, mus[L] + sigmas[L] * N(v))
// Follows generator/intensity/rand_gmm.py:90-97 (sum) and :146-149 (lookup, noise, clamp).
// HBM-bound: reads 1 byte per seed volume per voxel (+4 B when noise is injected), writes 4 B.
// Each thread owns 4 consecutive voxels = one Philox block = one 16-byte store.
#include "common.cuh"

namespace fsg {

constexpr int GMM_THREADS = 256;
constexpr int GMM_MAX_LABELS = 256;

// Label of one packed seed word (fsg_unpack_job): tab[m] = shift | mask << 8 | base << 16, selected by the 3-bit meta field.
__device__ __forceinline__ uint32_t decode_word(uint32_t w, const uint32_t (&tab)[5]) {
  const uint32_t m = w & 7u;
  const uint32_t t = m == 1 ? tab[1] : (m == 2 ? tab[2] : (m == 3 ? tab[3] : (m == 4 ? tab[4] : 0u)));
  return (t >> 16) + ((w >> (t & 0xffu)) & ((t >> 8) & 0xffu));
}

// NSEED = number of leading non-NULL seed pointers (the host entry compacts them); NSEED == 0: packed mode, the
// labels are decoded from job.words (uint16 when job.word_bytes == 2, else uint32).
// SURF_ONLY: every job of the launch writes through its surface (the production hand-over to the deformation): the
// pairs hand-over (a barrier and a shared-memory exchange per iteration) and the linear store are compiled out —
// 0.272 -> 0.259 ms per 8 volumes (r02).
template <bool INJECT, int NSEED, bool SURF_ONLY, bool W16 = false>
__global__ void __launch_bounds__(GMM_THREADS) gmm_kernel(const __grid_constant__ Batch<fsg_gmm_job> batch, int64_t nvox) {
  const fsg_gmm_job& job = batch.j[blockIdx.y];
  __shared__ float2 s_ms[GMM_MAX_LABELS];  // (mu, sigma) per label
  for (int i = threadIdx.x; i < GMM_MAX_LABELS; i += GMM_THREADS) {
    const bool in = i < job.nlabels;
    s_ms[i] = make_float2(in ? job.mus[i] : 0.f, in ? job.sigmas[i] : 0.f);
  }
  __syncthreads();
  __shared__ float s_first[2][GMM_THREADS];  // first value of every thread's group (pairs output), double-buffered
  __shared__ uint2 s_dec[8];                  // packed mode: (shift, mask | base << 8) per 3-bit meta field (0, 5..7: label 0)
  if (NSEED == 0 && threadIdx.x < 8) {
    const int m = threadIdx.x;
    s_dec[m] = (m >= 1 && m <= 4) ? make_uint2((uint32_t)job.shift[m - 1], (uint32_t)job.mask[m - 1] | ((uint32_t)(10 * m) << 8)) : make_uint2(0u, 0u);
    // visibility: the __syncthreads above ran before; a second one follows below for packed launches
  }
  if (NSEED == 0) __syncthreads();

  const int8_t* __restrict__ sp[4] = {job.seed[0], job.seed[1], job.seed[2], job.seed[3]};
  uint32_t wtab[5] = {0, 0, 0, 0, 0};
  if (NSEED == 0) {
#pragma unroll
    for (int m = 1; m <= 4; ++m) wtab[m] = (uint32_t)job.shift[m - 1] | ((uint32_t)job.mask[m - 1] << 8) | ((uint32_t)(10 * m) << 16);
  }
  const bool w16 = W16 || job.word_bytes == 2;  // W16: every job of the launch has 16-bit words (the 32-bit decode is compiled out)
  auto label_at = [&](int64_t v) -> int {  // scalar path (block tails)
    if (NSEED == 0) return (int)decode_word(w16 ? (uint32_t)static_cast<const uint16_t*>(job.words)[v] : static_cast<const uint32_t*>(job.words)[v], wtab);
    int l = 0;
    for (int m = 0; m < NSEED; ++m) l += sp[m][v];
    return l;
  };
  const float* __restrict__ noise = job.noise;
  float* __restrict__ out = job.out;
  uint8_t* __restrict__ lab_out = job.labels_out;
  const fsg_rng rng = job.rng;
  uint32_t* __restrict__ pairs = job.out_pairs;
  const int row_len = job.row_len;
  // block-linear output (fsg_texvol): group g = 4 voxels of row (x, y) at column 4 * z4
  const unsigned long long surf = job.out_surf;
  const uint32_t rq = surf ? (uint32_t)row_len >> 2 : 1u, sny = surf ? (uint32_t)job.surf_ny : 1u;
  const bool pow2 = ((rq & (rq - 1)) | (sny & (sny - 1))) == 0;
  const int rq_sh = __ffs(rq) - 1, ny_sh = __ffs(sny) - 1;
  // Thread -> group mapping of the surface mode: a warp owns one 64-byte x 8-row tile (a GOB of the block-linear
  // layout: 4 float4 groups along z times 8 rows), so its 32 stores fill one contiguous 512-byte block of the
  // array (measured on B200: 5.4 TB/s, the same as linear stores; a warp spread along one row writes eight
  // 64-byte pieces of eight tiles and reaches 3.4 TB/s).  The group index g — Philox counter and voxel offset —
  // of a voxel does not depend on which thread handles it.
  const bool gob = surf && (sny & 7u) == 0 && (rq & 3u) == 0;
  const uint32_t tiles_z = rq >> 2, tiles_y = sny >> 3;

  const int64_t ngroups = nvox >> 2;  // whole groups of 4 voxels; the tail is handled below
  const int64_t stride = (int64_t)gridDim.x * GMM_THREADS;
  int it = 0;
  for (int64_t gbase = (int64_t)blockIdx.x * GMM_THREADS; gbase < ngroups; gbase += stride, it ^= 1) {
    int64_t g = gbase + threadIdx.x;
    const bool active = g < ngroups;
    uint32_t sx_ = 0, sy_ = 0, sz4 = 0;  // surface coordinates of the group
    if ((SURF_ONLY || surf) && active) {
      const uint32_t gi = (uint32_t)g;
      if (gob) {
        const uint32_t lane = gi & 31u, tile = gi >> 5;
        uint32_t zc, t2, y8;
        if (pow2) {
          zc = tile & (tiles_z - 1), t2 = tile >> (rq_sh - 2), y8 = t2 & (tiles_y - 1), sx_ = t2 >> (ny_sh - 3);
        } else {
          t2 = tile / tiles_z, zc = tile - t2 * tiles_z, sx_ = t2 / tiles_y, y8 = t2 - sx_ * tiles_y;
        }
        sy_ = y8 * 8u + (lane >> 2), sz4 = zc * 4u + (lane & 3u);
        g = (int64_t)((sx_ * sny + sy_) * rq + sz4);
      } else if (pow2) {
        const uint32_t row = gi >> rq_sh;
        sz4 = gi & (rq - 1), sx_ = row >> ny_sh, sy_ = row & (sny - 1);
      } else {
        const uint32_t row = gi / rq;
        sz4 = gi - row * rq, sx_ = row / sny, sy_ = row - sx_ * sny;
      }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
    const int64_t v0 = g * 4;
    // labels of the seed volumes are disjoint small non-negative codes: one byte-wise SIMD add sums 4 voxels
    uint32_t lab4 = 0;
    if (NSEED == 0) {
      if (w16) {
        const uint2 q = __ldcs(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(job.words) + v0));
        auto dec = [&](uint32_t w) -> uint32_t {  // meta ? 10 meta + ((w >> shift) & mask) : 0, tables in shared memory
          const uint2 t = s_dec[w & 7u];
          return (t.y >> 8) + ((w >> t.x) & (t.y & 0xffu));
        };
        lab4 = dec(q.x & 0xffffu) | (dec(q.x >> 16) << 8) | (dec(q.y & 0xffffu) << 16) | (dec(q.y >> 16) << 24);
      } else {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(static_cast<const uint32_t*>(job.words) + v0));
        lab4 = decode_word(q.x, wtab) | (decode_word(q.y, wtab) << 8) | (decode_word(q.z, wtab) << 16) | (decode_word(q.w, wtab) << 24);
      }
    }
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
    if (SURF_ONLY || surf) {
      surf2DLayeredwrite<float4>(o, (cudaSurfaceObject_t)surf, (int)(sz4 * 16u), (int)sy_, (int)sx_);
    } else if (out) {
      *reinterpret_cast<float4*>(out + v0) = o;
    }
    if (lab_out) *reinterpret_cast<uint32_t*>(lab_out + v0) = lab4;
    }
    if (!SURF_ONLY && pairs) {  // block-uniform
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
            const int l = label_at(v4);
            const float nz = INJECT ? noise[v4] : philox_normal4(rng, (uint32_t)(g + 1)).x;
            const float2 ms = s_ms[l & (GMM_MAX_LABELS - 1)];
            nxt = fmaxf(add_rn(ms.x, mul_rn(ms.y, nz)), 0.f);
          }
        }
        if (job.pairs_float) {  // lossless float2 pairs: (I[v], I[v+1]) per voxel, two 16-byte stores
          float4* o2 = reinterpret_cast<float4*>(reinterpret_cast<float2*>(pairs) + v0);
          o2[0] = make_float4(o.x, o.y, o.y, o.z);
          o2[1] = make_float4(o.z, o.w, o.w, nxt);
          continue;
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
      const int l = label_at(v0 + e);
      const float nz = INJECT ? noise[v0 + e] : nn[e];
      const float2 ms = s_ms[l & (GMM_MAX_LABELS - 1)];
      if (out) out[v0 + e] = fmaxf(add_rn(ms.x, mul_rn(ms.y, nz)), 0.f);
      if (lab_out) lab_out[v0 + e] = (uint8_t)l;
    }
  }
}

// Bit-packed seed cache -> label volume of one sample (see fsg.h).  HBM-bound: 2 or 4 bytes in, 1 out.
template <typename Word>
__global__ void __launch_bounds__(256) unpack_seeds_kernel(const __grid_constant__ Batch<fsg_unpack_job> batch, int64_t nvox) {
  const fsg_unpack_job& job = batch.j[blockIdx.y];
  const Word* __restrict__ words = static_cast<const Word*>(job.words);
  // per meta-label: shift | mask << 8 | base << 16 in one register, selected by the 3-bit meta field
  uint32_t tab[5];
  tab[0] = 0;
#pragma unroll
  for (int m = 1; m <= 4; ++m) tab[m] = (uint32_t)job.shift[m - 1] | ((uint32_t)job.mask[m - 1] << 8) | ((uint32_t)(10 * m) << 16);
  constexpr int PER = 16 / sizeof(Word);  // words per 16-byte load
  const int64_t ngroups = nvox / PER;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto decode = [&](uint32_t w) -> uint32_t {
    const uint32_t m = w & 7u;
    const uint32_t t = m == 1 ? tab[1] : (m == 2 ? tab[2] : (m == 3 ? tab[3] : (m == 4 ? tab[4] : 0u)));
    return (t >> 16) + ((w >> (t & 0xffu)) & ((t >> 8) & 0xffu));
  };
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += stride) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(words) + g);
    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
    if (sizeof(Word) == 2) {
      uint32_t o[2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
        o[h] = decode(q[2 * h] & 0xffffu) | (decode(q[2 * h] >> 16) << 8) | (decode(q[2 * h + 1] & 0xffffu) << 16) | (decode(q[2 * h + 1] >> 16) << 24);
      *reinterpret_cast<uint2*>(job.out + g * PER) = make_uint2(o[0], o[1]);
    } else {
      *reinterpret_cast<uint32_t*>(job.out + g * PER) = decode(q[0]) | (decode(q[1]) << 8) | (decode(q[2]) << 16) | (decode(q[3]) << 24);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t v = ngroups * PER; v < nvox; ++v) job.out[v] = (uint8_t)decode((uint32_t)words[v]);
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

template <bool INJECT, bool SURF_ONLY>
static void launch_gmm(const Batch<fsg_gmm_job>& b, int nseed, bool w16, dim3 grid, int64_t nvox, cudaStream_t s) {
  switch (nseed) {
    case 0:
      if (w16)
        gmm_kernel<INJECT, 0, SURF_ONLY, true><<<grid, GMM_THREADS, 0, s>>>(b, nvox);
      else
        gmm_kernel<INJECT, 0, SURF_ONLY><<<grid, GMM_THREADS, 0, s>>>(b, nvox);
      break;
    case 1: gmm_kernel<INJECT, 1, SURF_ONLY><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    case 2: gmm_kernel<INJECT, 2, SURF_ONLY><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    case 3: gmm_kernel<INJECT, 3, SURF_ONLY><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
    default: gmm_kernel<INJECT, 4, SURF_ONLY><<<grid, GMM_THREADS, 0, s>>>(b, nvox); break;
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
    FSG_REQUIRE((j.out || j.out_pairs || j.out_surf) && j.mus && j.sigmas, "fsg_gmm: job %d has a NULL out/mus/sigmas", i);
    if (j.out_surf)
      FSG_REQUIRE(!j.out && !j.out_pairs && j.row_len >= 4 && j.row_len % 4 == 0 && j.surf_ny >= 1 && nvox % ((int64_t)j.row_len * j.surf_ny) == 0,
                  "fsg_gmm: job %d: out_surf needs row_len (nz, a multiple of 4) and surf_ny with nz * ny dividing nvox, and no other output", i);
    if (j.out_pairs)
      FSG_REQUIRE(j.row_len >= 1 && nvox % j.row_len == 0 && nvox % 4 == 0 && (reinterpret_cast<uintptr_t>(j.out_pairs) & 15) == 0,
                  "fsg_gmm: job %d: out_pairs needs row_len dividing nvox, nvox %% 4 == 0 and a 16-byte aligned buffer", i);
    FSG_REQUIRE(j.nlabels >= 1 && j.nlabels <= GMM_MAX_LABELS, "fsg_gmm: nlabels=%d outside [1,%d]", j.nlabels, GMM_MAX_LABELS);
    FSG_REQUIRE((j.noise != nullptr) == inject, "fsg_gmm: jobs mix injected and Philox noise");
    // compact the seed pointers of the launch copy to the front
    int n = 0;
    if (j.words) {  // packed mode: labels decoded from the subject's seed words
      FSG_REQUIRE(j.word_bytes == 2 || j.word_bytes == 4, "fsg_gmm: job %d: word_bytes must be 2 or 4", i);
      FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.words) & 15) == 0, "fsg_gmm: job %d: words must be 16-byte aligned", i);
      for (int m = 0; m < 4; ++m)
        FSG_REQUIRE(j.shift[m] >= 0 && j.shift[m] < 8 * j.word_bytes && j.mask[m] >= 0 && j.mask[m] <= 15, "fsg_gmm: job %d: bad field for meta-label %d", i, m + 1);
      for (int m = 0; m < 4; ++m) b.j[i].seed[m] = nullptr;
    } else {
      for (int m = 0; m < 4; ++m)
        if (j.seed[m]) b.j[i].seed[n++] = j.seed[m];
      for (int m = n; m < 4; ++m) b.j[i].seed[m] = nullptr;
      FSG_REQUIRE(n >= 1, "fsg_gmm: job %d has no seed volume", i);
    }
    FSG_REQUIRE(nseed < 0 || nseed == n, "fsg_gmm: jobs of one launch must have the same number of seed volumes (or all be packed)");
    nseed = n;
    for (int m = 0; m < 4 && !j.words; ++m) FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.seed[m]) & 3) == 0, "fsg_gmm: seed pointers must be 4-byte aligned");
    FSG_REQUIRE((reinterpret_cast<uintptr_t>(j.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.noise) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.labels_out) & 3) == 0,
                "fsg_gmm: out/noise must be 16-byte aligned");
  }
  const int64_t ngroups = (nvox + 3) / 4;
  int64_t want = (ngroups + GMM_THREADS - 1) / GMM_THREADS;
  const int64_t cap = 148 * 16;  // (r02: a single wave of 148 * 8 blocks over the launch measured 0.248 vs 0.230 ms)
  dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)njobs);
  bool surf_only = true, w16 = true;
  for (int i = 0; i < njobs; ++i) surf_only = surf_only && jobs[i].out_surf != 0, w16 = w16 && jobs[i].words && jobs[i].word_bytes == 2;
  if (inject) {
    if (surf_only) launch_gmm<true, true>(b, nseed, w16, grid, nvox, as_stream(stream));
    else launch_gmm<true, false>(b, nseed, w16, grid, nvox, as_stream(stream));
  } else {
    if (surf_only) launch_gmm<false, true>(b, nseed, w16, grid, nvox, as_stream(stream));
    else launch_gmm<false, false>(b, nseed, w16, grid, nvox, as_stream(stream));
  }
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

extern "C" int fsg_unpack_seeds(const fsg_unpack_job* jobs, int njobs, int64_t nvox, void* stream) {
  Batch<fsg_unpack_job> b;
  if (int rc = fill_batch(b, jobs, njobs)) return rc;
  FSG_REQUIRE(nvox >= 1, "fsg_unpack_seeds: nvox must be positive");
  const int wb = jobs[0].word_bytes;
  FSG_REQUIRE(wb == 2 || wb == 4, "fsg_unpack_seeds: word_bytes must be 2 or 4");
  for (int i = 0; i < njobs; ++i) {
    const fsg_unpack_job& j = jobs[i];
    FSG_REQUIRE(j.words && j.out, "fsg_unpack_seeds: job %d has a NULL pointer", i);
    FSG_REQUIRE(j.word_bytes == wb, "fsg_unpack_seeds: jobs of one call must share word_bytes");
    FSG_REQUIRE(((reinterpret_cast<uintptr_t>(j.words) & 15) | (reinterpret_cast<uintptr_t>(j.out) & 7)) == 0, "fsg_unpack_seeds: job %d: words must be 16-byte and out 8-byte aligned", i);
    for (int m = 0; m < 4; ++m)
      FSG_REQUIRE(j.shift[m] >= 0 && j.shift[m] < 8 * wb && j.mask[m] >= 0 && j.mask[m] <= 15, "fsg_unpack_seeds: job %d: bad field for meta-label %d", i, m + 1);
  }
  const int64_t want = (nvox / (16 / wb) + 255) / 256;
  const unsigned gx = (unsigned)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
  if (wb == 2)
    unpack_seeds_kernel<uint16_t><<<dim3(gx, njobs), 256, 0, as_stream(stream)>>>(b, nvox);
  else
    unpack_seeds_kernel<uint32_t><<<dim3(gx, njobs), 256, 0, as_stream(stream)>>>(b, nvox);
  return check_launch("fsg_unpack_seeds");
}
