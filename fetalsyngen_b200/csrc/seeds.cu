// K6 — seed generation (scripts/generate_seeds.py:133-211): label fusion + ordered partition of the
// image voxels by meta-label, then sklearn's GaussianMixture(n_init=5, init_params="k-means++")
// .fit_predict for one feature, restated for the device:
//   * every (meta-label, number of components, initialisation) is one job; all jobs advance in the
//     same launches (grid.y = job), so a subject's ~100-180 fits cost one kernel sequence;
//   * greedy k-means++ needs no distance array in one dimension (the closest-centre distance is
//     recomputed from <= 16 centres in registers / shared memory) and no prefix sum: D^2 sampling
//     is an exponential race (argmax of d_i / E_i, E_i ~ Exp(1) from Philox), one reduction pass;
//   * E step + sufficient statistics are one pass over the (L2-resident) values in float64 — the
//     reference's sklearn call runs in float64 — with per-block partials summed in a fixed order
//     by the one-block M step, which also tests convergence and raises the job's `converged` flag:
//     blocks of finished jobs exit at once, the host never synchronises inside the EM loop.
// Everything here is tolerance-checked against sklearn (tests/golden/seeds_*.npz): identical
// iteration counts, parameters to ~1e-9 relative, identical labels.
#include <math.h>

#include "common.cuh"

namespace fsg {

constexpr int EM_THREADS = 256;
constexpr int EM_WARPS = EM_THREADS / 32;
constexpr int EM_BLOCKS = 128;  // blocks per job
constexpr int EM_ACC = 3 * FSG_EM_MAXK + 1;  // s0[16] | s1[16] | s2[16] | log-likelihood
constexpr int EM_ACC_STRIDE = 52;
constexpr int KPP_T = 4;  // most local trials: 2 + int(ln 16)
constexpr double LOG_2PI = 1.8378770664093453;
constexpr double EPS10 = 10 * 2.220446049250313e-16;  // 10 * finfo(float64).eps (_gaussian_mixture.py:312)

struct EmScratch {
  double acc[EM_BLOCKS][EM_ACC_STRIDE];
  double pot[EM_BLOCKS][KPP_T];
  float race_val[EM_BLOCKS][KPP_T];
  int32_t race_idx[EM_BLOCKS][KPP_T];
  double centre[FSG_EM_MAXK];
  double cand_x[KPP_T];
  int32_t cand[KPP_T];
  int32_t _pad[KPP_T];
};

__device__ __forceinline__ int kpp_trials(int k) { return k < 3 ? 2 : (k < 8 ? 3 : 4); }  // 2 + int(ln k), k <= 16

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ k-means++ seeding
__global__ void __launch_bounds__(EM_THREADS) kpp_race_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int c) {
  const fsg_em_job job = jobs[blockIdx.y];
  if (c >= job.k) return;
  EmScratch& S = scratch[blockIdx.y];
  __shared__ double sc[FSG_EM_MAXK];
  __shared__ float sv[EM_WARPS][KPP_T];
  __shared__ int si[EM_WARPS][KPP_T];
  if (threadIdx.x < c) sc[threadIdx.x] = S.centre[threadIdx.x];
  __syncthreads();
  const int T = c == 0 ? 1 : kpp_trials(job.k);
  float bv[KPP_T];
  int bi[KPP_T];
#pragma unroll
  for (int t = 0; t < KPP_T; ++t) {
    bv[t] = -1.f;
    bi[t] = 0;
  }
  const Philox ph(job.rng_seed);
  for (int64_t i = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x; i < job.n; i += (int64_t)EM_BLOCKS * EM_THREADS) {
    const double x = (double)job.x[i];
    double d = 1.0;  // first centre: uniform draw
    if (c > 0) {
      d = INFINITY;
      for (int k = 0; k < c; ++k) {
        const double e = x - sc[k];
        d = fmin(d, e * e);
      }
    }
    const uint4 w = ph((uint32_t)i, (uint32_t)c, (uint32_t)job.rng_stream, (uint32_t)(job.rng_stream >> 32));
    const uint32_t word[4] = {w.x, w.y, w.z, w.w};
    const float df = (float)d;
#pragma unroll
    for (int t = 0; t < KPP_T; ++t) {
      if (t < T) {
        const float u = ((float)(word[t] >> 9) + 0.5f) * (1.0f / 8388608.0f);  // (0, 1), exact in float32
        const float v = df / -__logf(u);
        if (v > bv[t]) {
          bv[t] = v;
          bi[t] = (int)i;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < KPP_T; ++t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv[t], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi[t], o);
      if (ov > bv[t] || (ov == bv[t] && oi < bi[t])) {
        bv[t] = ov;
        bi[t] = oi;
      }
    }
    if (lane == 0) {
      sv[warp][t] = bv[t];
      si[warp][t] = bi[t];
    }
  }
  __syncthreads();
  if (threadIdx.x < KPP_T) {
    const int t = threadIdx.x;
    float v = sv[0][t];
    int ix = si[0][t];
    for (int w2 = 1; w2 < EM_WARPS; ++w2)
      if (sv[w2][t] > v || (sv[w2][t] == v && si[w2][t] < ix)) {
        v = sv[w2][t];
        ix = si[w2][t];
      }
    S.race_val[blockIdx.x][t] = v;
    S.race_idx[blockIdx.x][t] = ix;
  }
}

__global__ void kpp_pick_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int c) {
  const fsg_em_job job = jobs[blockIdx.x];
  if (c >= job.k) return;
  EmScratch& S = scratch[blockIdx.x];
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = c == 0 ? 1 : kpp_trials(job.k);
  if (t >= T) return;
  float v = -2.f;
  int ix = 0x7fffffff;
  for (int b = lane; b < EM_BLOCKS; b += 32) {
    const float ov = S.race_val[b][t];
    const int oi = S.race_idx[b][t];
    if (ov > v || (ov == v && oi < ix)) {
      v = ov;
      ix = oi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    if (ov > v || (ov == v && oi < ix)) {
      v = ov;
      ix = oi;
    }
  }
  if (lane == 0) {
    const double xv = (double)job.x[ix];
    S.cand[t] = ix;
    S.cand_x[t] = xv;
    if (c == 0) {
      S.centre[0] = xv;
      job.seeds[0] = ix;
    }
  }
}

__global__ void __launch_bounds__(EM_THREADS) kpp_pot_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int c) {
  const fsg_em_job job = jobs[blockIdx.y];
  if (c == 0 || c >= job.k) return;
  EmScratch& S = scratch[blockIdx.y];
  __shared__ double sc[FSG_EM_MAXK];
  __shared__ double scand[KPP_T];
  __shared__ double red[EM_WARPS][KPP_T];
  const int T = kpp_trials(job.k);
  if (threadIdx.x < c) sc[threadIdx.x] = S.centre[threadIdx.x];
  if (threadIdx.x < KPP_T) scand[threadIdx.x] = threadIdx.x < T ? S.cand_x[threadIdx.x] : 0.0;
  __syncthreads();
  double p[KPP_T] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x; i < job.n; i += (int64_t)EM_BLOCKS * EM_THREADS) {
    const double x = (double)job.x[i];
    double d = INFINITY;
    for (int k = 0; k < c; ++k) {
      const double e = x - sc[k];
      d = fmin(d, e * e);
    }
#pragma unroll
    for (int t = 0; t < KPP_T; ++t) {
      const double e = x - scand[t];
      p[t] += fmin(d, e * e);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t < KPP_T; ++t) {
    const double s = warp_sum(p[t]);
    if (lane == 0) red[warp][t] = s;
  }
  __syncthreads();
  if (threadIdx.x < KPP_T) {
    double s = 0.0;
    for (int w = 0; w < EM_WARPS; ++w) s += red[w][threadIdx.x];
    S.pot[blockIdx.x][threadIdx.x] = s;
  }
}

__global__ void kpp_choose_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int c) {
  const fsg_em_job job = jobs[blockIdx.x];
  if (c == 0 || c >= job.k) return;
  EmScratch& S = scratch[blockIdx.x];
  __shared__ double pot[KPP_T];
  const int T = kpp_trials(job.k);
  if (threadIdx.x < T) {
    double s = 0.0;
    for (int b = 0; b < EM_BLOCKS; ++b) s += S.pot[b][threadIdx.x];
    pot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int best = 0;
    for (int t = 1; t < T; ++t)
      if (pot[t] < pot[best]) best = t;  // np.argmin: first minimum
    S.centre[c] = S.cand_x[best];
    job.seeds[c] = S.cand[best];
  }
}

// ------------------------------------------------------------------ EM
// _initialize (mixture/_gaussian_mixture.py:849-877) with the one-hot responsibilities of the
// k-means++ branch (mixture/_base.py:149-158): component j owns the single sample seeds[j].
__global__ void em_init_kernel(const fsg_em_job* __restrict__ jobs, double reg_covar) {
  const fsg_em_job job = jobs[blockIdx.x];
  const int k = threadIdx.x;
  if (k < job.k) {
    const double nk = 1.0 + EPS10;
    const double xs = (double)job.x[job.seeds[k]];
    const double mean = xs / nk;
    const double diff = xs - mean;
    job.params[k] = nk / (double)job.n;
    job.params[FSG_EM_MAXK + k] = mean;
    job.params[2 * FSG_EM_MAXK + k] = diff * diff / nk + reg_covar;
  }
  if (k == 0) {
    job.state[0] = 0;
    job.state[1] = 0;
    job.trace[0] = -INFINITY;
  }
}

// Per-component constants of _estimate_log_gaussian_prob (:521-553) + log weights, in shared memory.
struct EmParams {
  double pc[FSG_EM_MAXK];   // precision Cholesky factor 1 / sqrt(cov)
  double mpc[FSG_EM_MAXK];  // mean * pc
  double ld[FSG_EM_MAXK];   // log det = log(pc)
  double lw[FSG_EM_MAXK];   // log weight
};
__device__ __forceinline__ void load_params(EmParams& P, const fsg_em_job& job) {
  const int k = threadIdx.x;
  if (k < job.k) {
    const double pc = 1.0 / sqrt(job.params[2 * FSG_EM_MAXK + k]);
    P.pc[k] = pc;
    P.mpc[k] = job.params[FSG_EM_MAXK + k] * pc;
    P.ld[k] = log(pc);
    P.lw[k] = log(job.params[k]);
  }
  __syncthreads();
}
__device__ __forceinline__ double wlp(const EmParams& P, int k, double x) {
  const double y = x * P.pc[k] - P.mpc[k];
  return (-0.5 * (LOG_2PI + y * y) + P.ld[k]) + P.lw[k];
}

template <int KMAX>
__global__ void __launch_bounds__(EM_THREADS) em_estep_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int max_iter) {
  const fsg_em_job job = jobs[blockIdx.y];
  if (job.state[1] != 0 || job.state[0] >= max_iter) return;
  EmScratch& S = scratch[blockIdx.y];
  __shared__ EmParams P;
  __shared__ double red[EM_WARPS][EM_ACC];
  load_params(P, job);
  const int K = job.k;
  double s0[KMAX], s1[KMAX], s2[KMAX], ll = 0.0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) s0[k] = s1[k] = s2[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x; i < job.n; i += (int64_t)EM_BLOCKS * EM_THREADS) {
    const double x = (double)job.x[i];
    double t[KMAX], m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        t[k] = wlp(P, k, x);
        m = fmax(m, t[k]);
      }
    double s = 0.0;  // logsumexp (mixture/_base.py:573)
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) s += exp(t[k] - m);
    const double lpn = log(s) + m;
    ll += lpn;
    const double xx = x * x;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) {
        const double r = exp(t[k] - lpn);  // responsibility = exp(log_resp)
        s0[k] += r;
        s1[k] += r * x;
        s2[k] += r * xx;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const double a = warp_sum(s0[k]), b = warp_sum(s1[k]), c = warp_sum(s2[k]);
    if (lane == 0) {
      red[warp][k] = a;
      red[warp][FSG_EM_MAXK + k] = b;
      red[warp][2 * FSG_EM_MAXK + k] = c;
    }
  }
  ll = warp_sum(ll);
  if (lane == 0) red[warp][3 * FSG_EM_MAXK] = ll;
  __syncthreads();
  if (threadIdx.x < EM_ACC) {
    const int v = threadIdx.x, k = v % FSG_EM_MAXK;
    double s = 0.0;
    if (v == 3 * FSG_EM_MAXK || k < KMAX)
      for (int w = 0; w < EM_WARPS; ++w) s += red[w][v];
    S.acc[blockIdx.x][v] = s;
  }
}

// _m_step (:883-901) from the block partials, the lower bound and the convergence test of
// BaseMixture.fit_predict (mixture/_base.py:265-278).
__global__ void em_mstep_kernel(const fsg_em_job* __restrict__ jobs, EmScratch* __restrict__ scratch, int max_iter, double tol, double reg_covar) {
  const fsg_em_job job = jobs[blockIdx.x];
  const int n_iter = job.state[0], conv = job.state[1];
  if (conv != 0 || n_iter >= max_iter) return;
  EmScratch& S = scratch[blockIdx.x];
  __shared__ double snk[FSG_EM_MAXK];
  const int k = threadIdx.x;
  double nk = 0.0, mean = 0.0, cov = 0.0;
  if (k < job.k) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < EM_BLOCKS; ++b) {
      s0 += S.acc[b][k];
      s1 += S.acc[b][FSG_EM_MAXK + k];
      s2 += S.acc[b][2 * FSG_EM_MAXK + k];
    }
    nk = s0 + EPS10;
    mean = s1 / nk;
    // sum r (x - mean)^2 expanded around the new mean
    const double ss = fmax((s2 - 2.0 * mean * s1) + mean * mean * s0, 0.0);
    cov = ss / nk + reg_covar;
    snk[k] = nk;
  }
  __syncthreads();
  if (k < job.k) {
    double tot = 0.0;
    for (int q = 0; q < job.k; ++q) tot += snk[q];
    job.params[k] = nk / tot;
    job.params[FSG_EM_MAXK + k] = mean;
    job.params[2 * FSG_EM_MAXK + k] = cov;
  }
  if (k == 32) {
    double ll = 0.0;
    for (int b = 0; b < EM_BLOCKS; ++b) ll += S.acc[b][3 * FSG_EM_MAXK];
    const double lower = ll / (double)job.n;
    const double prev = job.trace[n_iter];
    job.trace[n_iter + 1] = lower;
    job.state[0] = n_iter + 1;
    job.state[1] = fabs(lower - prev) < tol ? 1 : 0;
  }
}

// Final E step of fit_predict (mixture/_base.py:307-312): argmax of the weighted log probabilities.
__global__ void __launch_bounds__(EM_THREADS) em_predict_kernel(const fsg_em_job* __restrict__ jobs) {
  const fsg_em_job job = jobs[blockIdx.y];
  __shared__ EmParams P;
  const int K = job.k;
  if (K > 1) load_params(P, job);
  for (int64_t i = (int64_t)blockIdx.x * EM_THREADS + threadIdx.x; i < job.n; i += (int64_t)EM_BLOCKS * EM_THREADS) {
    int best = 0;
    if (K > 1) {
      const double x = (double)job.x[i];
      double bv = wlp(P, 0, x);
      for (int k = 1; k < K; ++k) {
        const double v = wlp(P, k, x);
        if (v > bv) {
          bv = v;
          best = k;
        }
      }
    }
    if (job.labels) job.labels[i] = (uint8_t)best;
    if (job.out) job.out[job.index[i]] = (int8_t)(job.label_base + best);
  }
}

// ------------------------------------------------------------------ label fusion + ordered partition
constexpr int PT_THREADS = 256;
constexpr int PT_PER = 16;
constexpr int PT_TILE = PT_THREADS * PT_PER;

struct Lut {
  uint8_t v[256];
};

__device__ __forceinline__ int meta_of(float img, uint8_t seg, const uint8_t* lut) {
  const int v = lut[seg];
  if (v != 4) return v;
  return (img != 0.f && img == img) ? 4 : 0;  // background label with signal -> non-brain tissue; NaN counts as 0
}

// loads the thread's PT_PER consecutive voxels; returns the number valid
__device__ __forceinline__ int load_run(const float* __restrict__ image, const uint8_t* __restrict__ seg, int64_t n, int64_t base, float* img, uint8_t* sg) {
  if (base + PT_PER <= n) {
#pragma unroll
    for (int q = 0; q < PT_PER / 4; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(image + base) + q);
      img[4 * q] = v.x;
      img[4 * q + 1] = v.y;
      img[4 * q + 2] = v.z;
      img[4 * q + 3] = v.w;
    }
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(seg + base));
    const uint32_t w[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int q = 0; q < PT_PER; ++q) sg[q] = (uint8_t)(w[q >> 2] >> (8 * (q & 3)));
    return PT_PER;
  }
  int cnt = 0;
  for (int q = 0; q < PT_PER; ++q) {
    const bool in = base + q < n;
    img[q] = in ? image[base + q] : 0.f;
    sg[q] = in ? seg[base + q] : 0;
    cnt += in;
  }
  return cnt;
}

// four 16-bit counters (meta-labels 1..4) in one word: a thread holds <= 16, a block <= 4096 per field
__device__ __forceinline__ unsigned long long count_run(const float* img, const uint8_t* sg, int valid, const uint8_t* lut) {
  unsigned long long c = 0;
#pragma unroll
  for (int q = 0; q < PT_PER; ++q) {
    const int m = q < valid ? meta_of(img[q], sg[q], lut) : 0;
    if (m) c += 1ull << (16 * (m - 1));
  }
  return c;
}

__global__ void __launch_bounds__(PT_THREADS) part_count_kernel(const float* __restrict__ image, const uint8_t* __restrict__ seg, const __grid_constant__ Lut lut, int64_t n,
                                                                unsigned long long* __restrict__ blockcnt) {
  __shared__ uint8_t slut[256];
  __shared__ unsigned long long red[PT_THREADS / 32];
  slut[threadIdx.x] = lut.v[threadIdx.x];
  __syncthreads();
  float img[PT_PER];
  uint8_t sg[PT_PER];
  const int64_t base = ((int64_t)blockIdx.x * PT_THREADS + threadIdx.x) * PT_PER;
  const int valid = base < n ? load_run(image, seg, n, base, img, sg) : 0;
  unsigned long long c = count_run(img, sg, valid, slut);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int w = 0; w < PT_THREADS / 32; ++w) s += red[w];
    blockcnt[blockIdx.x] = s;
  }
}

// exclusive scan of the per-block counts (one block): offsets[b][m] = partition base of m + members
// of m in blocks < b; counts[m] = partition sizes
__global__ void __launch_bounds__(1024) part_scan_kernel(const unsigned long long* __restrict__ blockcnt, int nblocks, int4* __restrict__ offsets, int64_t* __restrict__ counts) {
  __shared__ int4 wsum[32];
  __shared__ int4 total;
  const int chunk = (nblocks + 1023) / 1024;
  const int b0 = threadIdx.x * chunk, b1 = min(b0 + chunk, nblocks);
  int4 loc = make_int4(0, 0, 0, 0);
  for (int b = b0; b < b1; ++b) {
    const unsigned long long c = blockcnt[b];
    loc.x += (int)(c & 0xffff);
    loc.y += (int)((c >> 16) & 0xffff);
    loc.z += (int)((c >> 32) & 0xffff);
    loc.w += (int)((c >> 48) & 0xffff);
  }
  int4 inc = loc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o);
    const int z = __shfl_up_sync(0xffffffffu, inc.z, o), w = __shfl_up_sync(0xffffffffu, inc.w, o);
    if (lane >= o) {
      inc.x += x;
      inc.y += y;
      inc.z += z;
      inc.w += w;
    }
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int4 run = make_int4(0, 0, 0, 0);
    for (int w = 0; w < 32; ++w) {
      const int4 t = wsum[w];
      wsum[w] = run;
      run.x += t.x;
      run.y += t.y;
      run.z += t.z;
      run.w += t.w;
    }
    total = run;
    counts[0] = run.x;
    counts[1] = run.y;
    counts[2] = run.z;
    counts[3] = run.w;
  }
  __syncthreads();
  const int4 wb = wsum[warp], tot = total;
  // exclusive prefix of this thread's chunk + partition bases
  int4 run = make_int4(wb.x + inc.x - loc.x, tot.x + wb.y + inc.y - loc.y, tot.x + tot.y + wb.z + inc.z - loc.z, tot.x + tot.y + tot.z + wb.w + inc.w - loc.w);
  for (int b = b0; b < b1; ++b) {
    const unsigned long long c = blockcnt[b];
    offsets[b] = run;
    run.x += (int)(c & 0xffff);
    run.y += (int)((c >> 16) & 0xffff);
    run.z += (int)((c >> 32) & 0xffff);
    run.w += (int)((c >> 48) & 0xffff);
  }
}

__global__ void __launch_bounds__(PT_THREADS) part_scatter_kernel(const float* __restrict__ image, const uint8_t* __restrict__ seg, const __grid_constant__ Lut lut, int64_t n,
                                                                  const int4* __restrict__ offsets, float* __restrict__ x, int32_t* __restrict__ index) {
  __shared__ uint8_t slut[256];
  __shared__ unsigned long long wtot[PT_THREADS / 32];
  slut[threadIdx.x] = lut.v[threadIdx.x];
  __syncthreads();
  float img[PT_PER];
  uint8_t sg[PT_PER];
  const int64_t base = ((int64_t)blockIdx.x * PT_THREADS + threadIdx.x) * PT_PER;
  const int valid = base < n ? load_run(image, seg, n, base, img, sg) : 0;
  const unsigned long long mine = count_run(img, sg, valid, slut);
  unsigned long long inc = mine;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  unsigned long long pre = inc - mine;
  for (int w = 0; w < warp; ++w) pre += wtot[w];
  const int4 off = offsets[blockIdx.x];
  int pos[4] = {off.x + (int)(pre & 0xffff), off.y + (int)((pre >> 16) & 0xffff), off.z + (int)((pre >> 32) & 0xffff), off.w + (int)((pre >> 48) & 0xffff)};
#pragma unroll
  for (int q = 0; q < PT_PER; ++q) {
    const int m = q < valid ? meta_of(img[q], sg[q], slut) : 0;
    if (m) {
      const int p = pos[m - 1]++;
      x[p] = img[q] == img[q] ? img[q] : 0.f;
      index[p] = (int32_t)(base + q);
    }
  }
}

// ------------------------------------------------------------------ host side
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int stage_jobs(const fsg_em_job* jobs, int njobs, void* workspace, int64_t bytes, cudaStream_t s, const char* who, const fsg_em_job** table, EmScratch** scratch, int* kmax) {
  FSG_REQUIRE(jobs && njobs >= 1 && njobs <= 65535, "%s: njobs=%d outside [1,65535]", who, njobs);
  FSG_REQUIRE(workspace && bytes >= fsg_em_workspace(njobs), "%s: workspace of %lld bytes is smaller than fsg_em_workspace(%d)", who, (long long)bytes, njobs);
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  *kmax = 1;
  for (int i = 0; i < njobs; ++i) {
    const fsg_em_job& j = jobs[i];
    FSG_REQUIRE(j.x && j.n >= 1 && j.n < ((int64_t)1 << 31), "%s: job %d has no values (n=%lld)", who, i, (long long)j.n);
    FSG_REQUIRE(j.k >= 1 && j.k <= FSG_EM_MAXK, "%s: job %d: k=%d outside [1,%d]", who, i, j.k, FSG_EM_MAXK);
    FSG_REQUIRE(j.n >= j.k, "%s: job %d: Expected n_samples >= n_components but got n_components = %d, n_samples = %lld", who, i, j.k, (long long)j.n);
    if (j.k > *kmax) *kmax = j.k;
  }
  if (cudaMemcpyAsync(workspace, jobs, sizeof(fsg_em_job) * njobs, cudaMemcpyHostToDevice, s) != cudaSuccess) {
    check_launch(who);
    return 2;
  }
  *table = static_cast<const fsg_em_job*>(workspace);
  *scratch = reinterpret_cast<EmScratch*>(static_cast<char*>(workspace) + align_up(sizeof(fsg_em_job) * njobs, 256));
  return 0;
}

}  // namespace fsg

using namespace fsg;

extern "C" int64_t fsg_em_workspace(int njobs) {
  if (njobs < 1) return 0;
  return (int64_t)(align_up(sizeof(fsg_em_job) * (size_t)njobs, 256) + sizeof(EmScratch) * (size_t)njobs);
}

extern "C" int fsg_em_seed(const fsg_em_job* jobs, int njobs, void* workspace, int64_t bytes, void* stream) {
  cudaStream_t s = as_stream(stream);
  const fsg_em_job* table;
  EmScratch* scratch;
  int kmax;
  if (int rc = stage_jobs(jobs, njobs, workspace, bytes, s, "fsg_em_seed", &table, &scratch, &kmax)) return rc;
  for (int i = 0; i < njobs; ++i) FSG_REQUIRE(jobs[i].seeds, "fsg_em_seed: job %d has a NULL seeds buffer", i);
  const dim3 grid(EM_BLOCKS, njobs);
  for (int c = 0; c < kmax; ++c) {
    kpp_race_kernel<<<grid, EM_THREADS, 0, s>>>(table, scratch, c);
    kpp_pick_kernel<<<njobs, 32 * KPP_T, 0, s>>>(table, scratch, c);
    if (c > 0) {
      kpp_pot_kernel<<<grid, EM_THREADS, 0, s>>>(table, scratch, c);
      kpp_choose_kernel<<<njobs, 32, 0, s>>>(table, scratch, c);
    }
  }
  return check_launch("fsg_em_seed");
}

extern "C" int fsg_em_fit(const fsg_em_job* jobs, int njobs, int max_iter, double tol, double reg_covar, void* workspace, int64_t bytes, void* stream) {
  cudaStream_t s = as_stream(stream);
  const fsg_em_job* table;
  EmScratch* scratch;
  int kmax;
  FSG_REQUIRE(max_iter >= 1 && tol >= 0 && reg_covar >= 0, "fsg_em_fit: bad max_iter / tol / reg_covar");
  if (int rc = stage_jobs(jobs, njobs, workspace, bytes, s, "fsg_em_fit", &table, &scratch, &kmax)) return rc;
  for (int i = 0; i < njobs; ++i) FSG_REQUIRE(jobs[i].seeds && jobs[i].params && jobs[i].trace && jobs[i].state, "fsg_em_fit: job %d has a NULL seeds/params/trace/state buffer", i);
  em_init_kernel<<<njobs, 32, 0, s>>>(table, reg_covar);
  const dim3 grid(EM_BLOCKS, njobs);
  for (int it = 0; it < max_iter; ++it) {
    if (kmax <= 4)
      em_estep_kernel<4><<<grid, EM_THREADS, 0, s>>>(table, scratch, max_iter);
    else if (kmax <= 8)
      em_estep_kernel<8><<<grid, EM_THREADS, 0, s>>>(table, scratch, max_iter);
    else
      em_estep_kernel<16><<<grid, EM_THREADS, 0, s>>>(table, scratch, max_iter);
    em_mstep_kernel<<<njobs, 64, 0, s>>>(table, scratch, max_iter, tol, reg_covar);
  }
  return check_launch("fsg_em_fit");
}

extern "C" int fsg_em_predict(const fsg_em_job* jobs, int njobs, void* workspace, int64_t bytes, void* stream) {
  cudaStream_t s = as_stream(stream);
  const fsg_em_job* table;
  EmScratch* scratch;
  int kmax;
  if (int rc = stage_jobs(jobs, njobs, workspace, bytes, s, "fsg_em_predict", &table, &scratch, &kmax)) return rc;
  for (int i = 0; i < njobs; ++i) {
    FSG_REQUIRE(jobs[i].k == 1 || jobs[i].params, "fsg_em_predict: job %d has a NULL params buffer", i);
    FSG_REQUIRE(jobs[i].labels || jobs[i].out, "fsg_em_predict: job %d has neither labels nor out", i);
    FSG_REQUIRE(!jobs[i].out || jobs[i].index, "fsg_em_predict: job %d writes a volume but has no index", i);
  }
  em_predict_kernel<<<dim3(EM_BLOCKS, njobs), EM_THREADS, 0, s>>>(table);
  return check_launch("fsg_em_predict");
}

extern "C" int64_t fsg_seed_partition_workspace(int64_t n) {
  if (n < 1) return 0;
  const int64_t nblocks = (n + PT_TILE - 1) / PT_TILE;
  return (int64_t)(align_up((size_t)nblocks * sizeof(unsigned long long), 256) + (size_t)nblocks * sizeof(int4));
}

extern "C" int fsg_seed_partition(const float* image, const uint8_t* seg, const uint8_t* lut_host, int64_t n, float* x, int32_t* index, int64_t* counts, void* workspace,
                                  int64_t bytes, void* stream) {
  FSG_REQUIRE(image && seg && lut_host && x && index && counts, "fsg_seed_partition: NULL pointer");
  FSG_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), "fsg_seed_partition: n=%lld outside [1, 2^31)", (long long)n);
  FSG_REQUIRE(workspace && bytes >= fsg_seed_partition_workspace(n), "fsg_seed_partition: workspace too small");
  FSG_REQUIRE(((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(seg) | reinterpret_cast<uintptr_t>(workspace)) & 15) == 0,
              "fsg_seed_partition: image, seg and workspace must be 16-byte aligned");
  Lut lut;
  for (int i = 0; i < 256; ++i) {
    FSG_REQUIRE(lut_host[i] <= 4, "fsg_seed_partition: lut[%d]=%d is not a meta-label 0..4", i, lut_host[i]);
    lut.v[i] = lut_host[i];
  }
  cudaStream_t s = as_stream(stream);
  const int nblocks = (int)((n + PT_TILE - 1) / PT_TILE);
  unsigned long long* blockcnt = static_cast<unsigned long long*>(workspace);
  int4* offsets = reinterpret_cast<int4*>(static_cast<char*>(workspace) + align_up((size_t)nblocks * sizeof(unsigned long long), 256));
  part_count_kernel<<<nblocks, PT_THREADS, 0, s>>>(image, seg, lut, n, blockcnt);
  part_scan_kernel<<<1, 1024, 0, s>>>(blockcnt, nblocks, offsets, counts);
  part_scatter_kernel<<<nblocks, PT_THREADS, 0, s>>>(image, seg, lut, n, offsets, x, index);
  return check_launch("fsg_seed_partition");
}
