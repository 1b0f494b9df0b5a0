// K4ab — separable banded resampling: out = (Rx (x) Ry (x) Rz) in.
//
// The reference blurs the full-resolution volume with three zero-padded 1-D Gaussian
// convolutions and then samples it trilinearly on the coarse grid
// (augmentation/synthseg.py:63-107 -> utils/generation.py:84-110, 227-285).  Both steps are
// linear and separable, so per axis they compose into ONE banded matrix R_a (n_out x n_in)
// whose row I holds  w_f*taps(. - f_I) + w_c*taps(. - c_I)  (zero rows where the reference's
// linear sampler returns 0, zero padding = dropped taps).  fsg_sep_compose builds the rows on the
// device from the host's 1-D position tables and taps; fsg_sepconv applies them axis by axis,
// shrinking the volume at every pass:
//     x: [sx][sy][sz] -> [n0][sy][sz]     streaming, register sliding window, 4N + 4fN bytes
//     y: [n0][sy][sz] -> [n0][n1][sz]     streaming, register sliding window
//     z: [n0][n1][sz] -> [n0][n1][n2]     + RandNoise epilogue.  sz <= 256 (production): a warp per row, the row in
//                                         registers (8 samples per lane, halo by shuffles), 13-tap blur with static
//                                         register indexing, then the coarse samples (sep_zrow_kernel); otherwise
//                                         rows staged in shared memory with the composed windows (sep_z_kernel)
// instead of 3 full-resolution blur passes + a gather (24 N + 4(N+n^3) bytes before).
// With identity positions the same kernels are a plain separable blur (BlurCortex).
// Noise: Philox block = flat output index / 4, component = index % 4, in every kernel.
// r02 measurement that shaped this file: a schedule that ran z + y fused on the full-resolution volume and x last
// (4N(1 + 2f^2 + f^3) bytes instead of 4N(1 + 2f + 2f^2 + f^3)) was 2x SLOWER (0.82 vs 0.43 ms per 8 volumes): the
// z blur at full resolution costs ~27 instructions per voxel against ~10 for the streaming x pass, which already
// runs at 5.2 TB/s of its own traffic.  Cheap-first (x, y streaming) and the instruction-heavy axis last wins.
// Float image path: FMA accumulation, parity is the 1e-4 tolerance (tests/test_gpu_base.py).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace fsg {

constexpr int SEP_THREADS = 128;
#ifndef FSG_SEP_MINBLOCKS
#define FSG_SEP_MINBLOCKS 5
#endif
constexpr int SEPZ_THREADS = 256;
constexpr int SEPZ_ROWS = 32;

struct SepPass {
  const float* src[FSG_MAX_JOBS];
  float* dst[FSG_MAX_JOBS];
  const int16_t* q0[FSG_MAX_JOBS];
  const float* w[FSG_MAX_JOBS];
  int n_out[FSG_MAX_JOBS];
  int width[FSG_MAX_JOBS];
  int outer[FSG_MAX_JOBS];  // slow dimension count of this pass (1 for x, n0 for y, n0*n1 for z)
};

struct SepNoise {
  const float* noise[FSG_MAX_JOBS];
  fsg_rng rng[FSG_MAX_JOBS];
  float std[FSG_MAX_JOBS];
  int has[FSG_MAX_JOBS];
};

template <int VEC>
struct VecT;
template <>
struct VecT<1> {
  using T = float;
  static __device__ __forceinline__ T zero() { return 0.f; }
  static __device__ __forceinline__ void fma(T& a, float w, const T& v) { a = __fmaf_rn(w, v, a); }
};
template <>
struct VecT<2> {
  using T = float2;
  static __device__ __forceinline__ T zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ void fma(T& a, float w, const T& v) {
    a.x = __fmaf_rn(w, v.x, a.x);
    a.y = __fmaf_rn(w, v.y, a.y);
  }
};

// RandNoise epilogue (synthseg.py:217-235) of the last pass: out = max(0, out + std * N).  Injected draws are
// indexed by the flat output index; the Philox stream is laid out per output row so that a block of four normals
// never straddles rows: block = row * ceil(n_out / 4) + K / 4, component K % 4.
__device__ __forceinline__ float noise_normal(const SepNoise& nz, int jb, int64_t row, int K, int n_out, bool inject) {
  if (inject) return __ldg(nz.noise[jb] + row * n_out + K);
  const float4 q = philox_normal4(nz.rng[jb], (uint32_t)row * (uint32_t)((n_out + 3) >> 2) + (uint32_t)(K >> 2));
  const int comp = K & 3;
  return comp == 0 ? q.x : (comp == 1 ? q.y : (comp == 2 ? q.z : q.w));
}
__device__ __forceinline__ float noise_apply(float v, float std, float n) {
  v = __fmaf_rn(std, n, v);
  return v < 0.f ? 0.f : v;
}

// Streaming pass along a slow axis.  The volume is viewed as [outer][a_in][inner]; a thread owns
// VEC consecutive inner elements of one `outer` slab and walks the axis once, keeping the last W
// planes in a rotating register window (slot = plane index mod W, static after unrolling).
// Output I is emitted when its last source plane has been loaded; its weights are stored
// right-aligned in a W-wide row, so the tap loop is W FMAs on statically indexed registers.
template <int W, int VEC>
__global__ void __launch_bounds__(SEP_THREADS, FSG_SEP_MINBLOCKS) sep_stream_kernel(const __grid_constant__ SepPass p, int a_in, int inner) {
  using V = VecT<VEC>;
  using T = typename V::T;
  const int jb = blockIdx.y;
  const int n_out = p.n_out[jb], width = p.width[jb], outer = p.outer[jb];
  extern __shared__ float s_mem[];
  float* s_w = s_mem;                                       // [n_out][W], right-aligned
  int* s_last = reinterpret_cast<int*>(s_mem + n_out * W);  // [n_out] last source plane of each output
  for (int e = threadIdx.x; e < n_out * W; e += SEP_THREADS) {
    const int I = e / W, t = e - I * W - (W - width);
    s_w[e] = t >= 0 ? p.w[jb][I * width + t] : 0.f;
  }
  for (int I = threadIdx.x; I < n_out; I += SEP_THREADS) s_last[I] = (int)p.q0[jb][I] + width - 1;
  __syncthreads();

  const int cols = inner / VEC;
  const int chunks = (cols + SEP_THREADS - 1) / SEP_THREADS;
  for (int chunk = blockIdx.x; chunk < outer * chunks; chunk += gridDim.x) {
    const int o = chunk / chunks, c = (chunk - o * chunks) * SEP_THREADS + threadIdx.x;
    if (c >= cols) continue;
    const T* __restrict__ in = reinterpret_cast<const T*>(p.src[jb] + (size_t)o * a_in * inner) + c;
    T* __restrict__ out = reinterpret_cast<T*>(p.dst[jb] + (size_t)o * n_out * inner) + c;
    T win[W];
#pragma unroll
    for (int u = 0; u < W; ++u) win[u] = V::zero();
    int I = 0;
    int next_last = s_last[0];
    for (int qb = 0; qb < a_in; qb += W) {
      T cur[W];
#pragma unroll
      for (int u = 0; u < W; ++u) cur[u] = (qb + u < a_in) ? __ldcs(in + (size_t)(qb + u) * cols) : V::zero();
#pragma unroll
      for (int u = 0; u < W; ++u) {
        win[u] = cur[u];
        const int q = qb + u;
        while (next_last == q) {
          const float4* wr = reinterpret_cast<const float4*>(s_w + I * W);
          T acc = V::zero();
#pragma unroll
          for (int t4 = 0; t4 < W / 4; ++t4) {
            const float4 w4 = wr[t4];
            V::fma(acc, w4.x, win[(u + 1 + 4 * t4 + 0) % W]);
            V::fma(acc, w4.y, win[(u + 1 + 4 * t4 + 1) % W]);
            V::fma(acc, w4.z, win[(u + 1 + 4 * t4 + 2) % W]);
            V::fma(acc, w4.w, win[(u + 1 + 4 * t4 + 3) % W]);
          }
          out[(size_t)I * cols] = acc;
          ++I;
          next_last = I < n_out ? s_last[I] : -1;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- fused x + y
// The x and y passes in ONE kernel (r02): the x-pass result (4 f N bytes written and read back) never exists.
// A WARP owns a 32-float z chunk, a tile of XY_TJ outputs along y and a segment of the x outputs, and walks the
// source planes of the segment once:
//   per plane q: the y pass of the tile's rows with the rotating register window of sep_stream_kernel (lane = z,
//     one coalesced 128-byte load per row), each emitted y output stored in slot q mod 16 of a per-warp ring in
//     shared memory [16 planes][XY_TJ][32];
//   when q completes the window of x output I: out[I][J][z] = sum_t wx[I][t] * ring[q - 15 + t][J][z] for the
//     tile's J (weights right-aligned in 16 slots), one coalesced store per (I, J).
// Warps never synchronise with each other; rows shared by neighbouring y tiles (the 13-row halo) are re-read
// through L1 / L2.  Result = the same separable sum in the order y, x (x, y in the two-pass schedule): equal to
// rounding.
constexpr int XY_W = 16, XY_TJ = 8, XY_WARPS = 4;
struct SepXY {
  const float* src[FSG_MAX_JOBS];
  float* dst[FSG_MAX_JOBS];
  const int16_t* q0x[FSG_MAX_JOBS];
  const float* wx[FSG_MAX_JOBS];
  const int16_t* q0y[FSG_MAX_JOBS];
  const float* wy[FSG_MAX_JOBS];
  int n0[FSG_MAX_JOBS], n1[FSG_MAX_JOBS], widthx[FSG_MAX_JOBS], widthy[FSG_MAX_JOBS];
};

__global__ void __launch_bounds__(XY_WARPS * 32) sep_xy_kernel(const __grid_constant__ SepXY p, int sx, int sy, int sz, int xsegs, int seg_cap) {
  const int jb = blockIdx.y;
  const int n0 = p.n0[jb], n1 = p.n1[jb], widthx = p.widthx[jb], widthy = p.widthy[jb];
  const int nzc = sz / 32, nyg = (n1 + XY_TJ * XY_WARPS - 1) / (XY_TJ * XY_WARPS);
  int bid = blockIdx.x;
  const int zc = bid % nzc;
  bid /= nzc;
  const int yg = bid % nyg, xs = bid / nyg;
  if (xs >= xsegs) return;  // uniform per block: jobs with fewer y tiles than the largest
  const int seg_len = (n0 + xsegs - 1) / xsegs;
  const int Ia = xs * seg_len, Ib = min(n0, Ia + seg_len);
  if (Ia >= Ib) return;

  extern __shared__ __align__(16) float s_xy[];
  float* s_ring = s_xy;                                                 // [XY_WARPS][16][XY_TJ][32]
  float* s_wx = s_ring + XY_WARPS * XY_W * XY_TJ * 32;                  // [seg_cap][16] right-aligned
  int* s_lastx = reinterpret_cast<int*>(s_wx + seg_cap * XY_W);         // [seg_cap]
  float* s_wy = reinterpret_cast<float*>(s_lastx + seg_cap);            // [XY_TJ * XY_WARPS][16] right-aligned
  int* s_lasty = reinterpret_cast<int*>(s_wy + XY_TJ * XY_WARPS * XY_W);  // [XY_TJ * XY_WARPS]
  const int Jb0 = yg * XY_TJ * XY_WARPS;
  for (int e = threadIdx.x; e < (Ib - Ia) * XY_W; e += XY_WARPS * 32) {
    const int I = Ia + e / XY_W, t = e % XY_W - (XY_W - widthx);
    s_wx[e] = t >= 0 ? p.wx[jb][I * widthx + t] : 0.f;
  }
  for (int i = threadIdx.x; i < Ib - Ia; i += XY_WARPS * 32) s_lastx[i] = (int)p.q0x[jb][Ia + i] + widthx - 1;
  for (int e = threadIdx.x; e < XY_TJ * XY_WARPS * XY_W; e += XY_WARPS * 32) {
    const int J = Jb0 + e / XY_W, t = e % XY_W - (XY_W - widthy);
    s_wy[e] = (J < n1 && t >= 0) ? p.wy[jb][J * widthy + t] : 0.f;
  }
  for (int i = threadIdx.x; i < XY_TJ * XY_WARPS; i += XY_WARPS * 32) s_lasty[i] = Jb0 + i < n1 ? (int)p.q0y[jb][Jb0 + i] + widthy - 1 : -1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* ring = s_ring + warp * (XY_W * XY_TJ * 32) + lane;
#pragma unroll 4
  for (int e = 0; e < XY_W * XY_TJ; ++e) ring[e * 32] = 0.f;
  __syncthreads();

  const int Jl0 = warp * XY_TJ, J0 = Jb0 + Jl0;
  if (J0 >= n1) return;
  const int nJ = min(XY_TJ, n1 - J0);
  const int r_lo = (int)p.q0y[jb][J0], r_hi = s_lasty[Jl0 + nJ - 1];
  const int qa = (int)p.q0x[jb][Ia], qb = s_lastx[Ib - Ia - 1];
  const float* __restrict__ src = p.src[jb] + (size_t)zc * 32 + lane;
  float* __restrict__ dst = p.dst[jb] + (size_t)zc * 32 + lane;
  const size_t plane = (size_t)sy * sz;

  int Il = 0;  // x output index within the segment
  int next_last_x = s_lastx[0];
  for (int q = qa; q <= qb; ++q) {
    const float* __restrict__ in = src + (size_t)q * plane;
    float* const slot = ring + (q & (XY_W - 1)) * (XY_TJ * 32);
    // ---- y pass of the tile's rows of plane q
    float win[XY_W];
#pragma unroll
    for (int u = 0; u < XY_W; ++u) win[u] = 0.f;
    int Jl = 0;
    int next_last = s_lasty[Jl0];
    for (int rb = r_lo; rb <= r_hi; rb += XY_W) {
      float cur[XY_W];
#pragma unroll
      for (int u = 0; u < XY_W; ++u) cur[u] = (rb + u <= r_hi) ? __ldcs(in + (size_t)(rb + u) * sz) : 0.f;
#pragma unroll
      for (int u = 0; u < XY_W; ++u) {
        win[u] = cur[u];
        const int r = rb + u;
        while (next_last == r) {
          const float4* wr = reinterpret_cast<const float4*>(s_wy + (Jl0 + Jl) * XY_W);
          float acc = 0.f;
#pragma unroll
          for (int t4 = 0; t4 < XY_W / 4; ++t4) {
            const float4 w4 = wr[t4];
            acc = __fmaf_rn(w4.x, win[(u + 1 + 4 * t4 + 0) % XY_W], acc);
            acc = __fmaf_rn(w4.y, win[(u + 1 + 4 * t4 + 1) % XY_W], acc);
            acc = __fmaf_rn(w4.z, win[(u + 1 + 4 * t4 + 2) % XY_W], acc);
            acc = __fmaf_rn(w4.w, win[(u + 1 + 4 * t4 + 3) % XY_W], acc);
          }
          slot[Jl * 32] = acc;
          ++Jl;
          next_last = Jl < nJ ? s_lasty[Jl0 + Jl] : -1;
        }
      }
    }
    // ---- x outputs completed by plane q
    while (next_last_x == q) {
      const float4* wr = reinterpret_cast<const float4*>(s_wx + Il * XY_W);
      float wxr[XY_W];
#pragma unroll
      for (int t4 = 0; t4 < XY_W / 4; ++t4) {
        const float4 w4 = wr[t4];
        wxr[4 * t4] = w4.x, wxr[4 * t4 + 1] = w4.y, wxr[4 * t4 + 2] = w4.z, wxr[4 * t4 + 3] = w4.w;
      }
      float acc[XY_TJ];
#pragma unroll
      for (int j = 0; j < XY_TJ; ++j) acc[j] = 0.f;
#pragma unroll
      for (int e = 0; e < XY_W; ++e) {
        const float* rp = ring + ((q + 1 + e) & (XY_W - 1)) * (XY_TJ * 32);
#pragma unroll
        for (int j = 0; j < XY_TJ; ++j) acc[j] = __fmaf_rn(wxr[e], rp[j * 32], acc[j]);
      }
      float* out = dst + ((size_t)(Ia + Il) * n1 + J0) * sz;
#pragma unroll
      for (int j = 0; j < XY_TJ; ++j)
        if (j < nJ) out[(size_t)j * sz] = acc[j];
      ++Il;
      next_last_x = Ia + Il < Ib ? s_lastx[Il] : -1;
    }
  }
}

// Last pass: rows of the fast axis staged in shared memory.  Thread = one output position K
// (its W weights live in registers) looping over the block's rows; lanes are consecutive K,
// so the stores are coalesced and the shared-memory reads stride by ~1/factor.
// RandNoise epilogue (synthseg.py:217-235): out = max(0, out + std*N(flat output index)).
template <int W, bool INJECT>
__global__ void __launch_bounds__(SEPZ_THREADS) sep_z_kernel(const __grid_constant__ SepPass p, const __grid_constant__ SepNoise nz, int a_in) {
  const int jb = blockIdx.y;
  const int n_out = p.n_out[jb], width = p.width[jb], rows = p.outer[jb];
  // rows are stored with W zero floats in front (pitch a_in + W), so a right-aligned window that
  // starts before the row reads zeros and the tap loop needs no bounds logic
  extern __shared__ float s_rows[];  // [SEPZ_ROWS][W + a_in]
  const int pitch = a_in + W;
  const int row0 = blockIdx.x * SEPZ_ROWS;
  if (row0 >= rows) return;
  const int nrow = min(SEPZ_ROWS, rows - row0);
  {
    const float* __restrict__ src = p.src[jb] + (size_t)row0 * a_in;
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (a_in % 4 == 0)) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      const int a4 = a_in / 4;
      for (int e = threadIdx.x; e < nrow * a4; e += SEPZ_THREADS) {
        const int r = e / a4, c = e - r * a4;
        *reinterpret_cast<float4*>(s_rows + r * pitch + W + 4 * c) = __ldcs(s4 + e);
      }
    } else {
      for (int e = threadIdx.x; e < nrow * a_in; e += SEPZ_THREADS) {
        const int r = e / a_in, c = e - r * a_in;
        s_rows[r * pitch + W + c] = __ldcs(src + e);
      }
    }
    for (int e = threadIdx.x; e < nrow * W; e += SEPZ_THREADS) s_rows[(e / W) * pitch + (e % W)] = 0.f;
  }
  __syncthreads();

  const int kw = (n_out + 31) & ~31;                        // K extent rounded to whole warps
  const int groups = kw <= SEPZ_THREADS ? SEPZ_THREADS / kw : 1;  // row groups working in parallel
  const int g = threadIdx.x / kw;
  const bool has_noise = nz.has[jb] != 0;
  const float nstd = nz.std[jb];
  if (g >= groups) return;  // whole warps only (kw is a multiple of 32)
  for (int kb = threadIdx.x - g * kw; kb < kw; kb += (groups == 1 ? SEPZ_THREADS : kw)) {
    const int K = kb;
    const bool live = K < n_out;
    float wreg[W];
    int q0 = 0;
    if (live) {
      q0 = (int)p.q0[jb][K] - (W - width);  // right-aligned window start (may be < 0: zero weights there)
#pragma unroll
      for (int t = 0; t < W; ++t) wreg[t] = (t >= W - width) ? __ldg(p.w[jb] + K * width + (t - (W - width))) : 0.f;
    } else {
#pragma unroll
      for (int t = 0; t < W; ++t) wreg[t] = 0.f;
    }
    for (int r = g; r < nrow; r += groups) {
      float acc = 0.f;
      {
        const float* row = s_rows + r * pitch + W + q0;  // q0 >= -W (dead lanes: q0 = 0, zero weights)
        float acc2 = 0.f;
#pragma unroll
        for (int t = 0; t < W; t += 2) {
          acc = __fmaf_rn(wreg[t], row[t], acc);
          acc2 = __fmaf_rn(wreg[t + 1], row[t + 1], acc2);
        }
        acc = __fadd_rn(acc, acc2);
      }
      if (has_noise && live) acc = noise_apply(acc, nstd, noise_normal(nz, jb, row0 + r, K, n_out, INJECT));
      if (live) p.dst[jb][(size_t)(row0 + r) * n_out + K] = acc;
    }
  }
}

// Last pass for rows of at most 256 samples (the production extents): one warp per row, the row in registers.
// Lane l holds its 8 consecutive samples (two 16-byte loads, the next row already in flight), fetches the 6 + 6
// halo samples from its neighbours by shuffles and evaluates the zero-padded 13-tap Gaussian for its 8 positions
// with statically indexed registers (the reference's conv along z, utils/generation.py:84-110); the blurred row
// goes to shared memory and the lanes sample it at the coarse positions (w_f*B[f] + w_c*B[c], 0 where the
// reference's sampler returns 0; utils/generation.py:227-285).  The outputs are transposed through shared memory
// so that a lane owns four consecutive K: one Philox block and one 16-byte store per group.
// r02 ncu of sep_z_kernel at these extents: bound by shared-memory wavefronts (W loads per output, 2-way bank
// conflicts at stride 1/f); here a row costs 2 + 16 shared-memory instructions per lane instead of 16 per output.
constexpr int ZR_THREADS = 256;
#ifndef FSG_ZR_MINBLOCKS
#define FSG_ZR_MINBLOCKS 3
#endif

constexpr int ZR_R = 6;      // blur radius held in registers (13 taps)
constexpr int ZR_ROW = 256;  // row / coarse row limit (8 samples per lane)
struct SepZRow {
  const float* src[FSG_MAX_JOBS];
  float* dst[FSG_MAX_JOBS];
  const fsg_tab* pos[FSG_MAX_JOBS];
  const float* taps[FSG_MAX_JOBS];
  int ntaps[FSG_MAX_JOBS];
  int n_out[FSG_MAX_JOBS];
  int rows[FSG_MAX_JOBS];
};

__device__ __forceinline__ void zr_load_row(const float* __restrict__ src, int row, int rows, int sz, int lane, float (&v)[8]) {
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = 0.f;
  if (row < rows && 8 * lane < sz) {
    const float4* r4 = reinterpret_cast<const float4*>(src + (size_t)row * sz + 8 * lane);
    const float4 a = __ldcs(r4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    if (8 * lane + 4 < sz) {
      const float4 b = __ldcs(r4 + 1);
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  }
}

template <bool INJECT>
__global__ void __launch_bounds__(ZR_THREADS, FSG_ZR_MINBLOCKS) sep_zrow_kernel(const __grid_constant__ SepZRow p, const __grid_constant__ SepNoise nz, int sz) {
  const int jb = blockIdx.y;
  const int n2 = p.n_out[jb], rows = p.rows[jb];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ __align__(16) float s_b[ZR_THREADS / 32][ZR_ROW];  // blurred row of each warp
  __shared__ __align__(16) float s_o[ZR_THREADS / 32][ZR_ROW];  // coarse row of each warp (transposition)
  float* const my_b = s_b[warp];
  float* const my_o = s_o[warp];
  const char* const my_bb = reinterpret_cast<const char*>(my_b);

  float tp[2 * ZR_R + 1];
  {
    const int nt = p.taps[jb] ? p.ntaps[jb] : 1, r = nt / 2;
#pragma unroll
    for (int s_ = 0; s_ < 2 * ZR_R + 1; ++s_) {
      const int t = s_ - ZR_R + r;
      tp[s_] = (t >= 0 && t < nt) ? (p.taps[jb] ? __ldg(p.taps[jb] + t) : 1.0f) : 0.f;
    }
  }
  // coarse positions sampled by this lane: K = lane + 32 e (neighbouring lanes read neighbouring words)
  int zfc[8];  // byte offsets into the blurred row: f*4 | c*4 << 16, -1 = the sampler returns 0
  float zwc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int K = lane + 32 * e;
    zfc[e] = -1;
    zwc[e] = 0.f;
    if (K < n2) {
      if (p.pos[jb]) {
        const fsg_tab t = p.pos[jb][K];
        if (t.f >= 0) zfc[e] = ((int)t.f * 4) | (((int)t.c * 4) << 16);
        zwc[e] = t.wc;
      } else {
        zfc[e] = (K * 4) | ((K * 4) << 16);
      }
    }
  }
  const bool has_noise = nz.has[jb] != 0;
  const float nstd = nz.std[jb];
  const uint32_t kgroups = (uint32_t)((n2 + 3) >> 2);
  const float* __restrict__ src = p.src[jb];
  float* __restrict__ dst = p.dst[jb];
  const bool dst16 = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  const int stride = gridDim.x * (ZR_THREADS / 32);

  int row = blockIdx.x * (ZR_THREADS / 32) + warp;
  float cur[8];
  zr_load_row(src, row, rows, sz, lane, cur);
  for (; row < rows; row += stride) {
    float nxt[8];
    zr_load_row(src, row + stride, rows, sz, lane, nxt);
    float w20[8 + 2 * ZR_R];
#pragma unroll
    for (int m = 0; m < ZR_R; ++m) {
      const float l = __shfl_up_sync(0xffffffffu, cur[8 - ZR_R + m], 1), r = __shfl_down_sync(0xffffffffu, cur[m], 1);
      w20[m] = lane == 0 ? 0.f : l;
      w20[8 + ZR_R + m] = lane == 31 ? 0.f : r;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) w20[ZR_R + m] = cur[m];
    float b[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      float acc = 0.f, acc2 = 0.f;
#pragma unroll
      for (int s_ = 0; s_ < 2 * ZR_R; s_ += 2) {
        acc = __fmaf_rn(tp[s_], w20[m + s_], acc);
        acc2 = __fmaf_rn(tp[s_ + 1], w20[m + s_ + 1], acc2);
      }
      b[m] = __fadd_rn(__fmaf_rn(tp[2 * ZR_R], w20[m + 2 * ZR_R], acc), acc2);
    }
    __syncwarp();  // the previous row's readers of my_b / my_o are done
    reinterpret_cast<float4*>(my_b)[2 * lane] = make_float4(b[0], b[1], b[2], b[3]);
    reinterpret_cast<float4*>(my_b)[2 * lane + 1] = make_float4(b[4], b[5], b[6], b[7]);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
      if (zfc[e] >= 0) {
        const float lo = *reinterpret_cast<const float*>(my_bb + (zfc[e] & 0xffff)), hi = *reinterpret_cast<const float*>(my_bb + (zfc[e] >> 16));
        v = __fmaf_rn(__fsub_rn(1.0f, zwc[e]), lo, __fmul_rn(zwc[e], hi));
      }
      my_o[lane + 32 * e] = v;
    }
    __syncwarp();
    // ---- lane owns K = 4 lane + 128 h + (0..3): noise + store
    float* __restrict__ orow = dst + (size_t)row * n2;
    const bool vec = dst16 && (((size_t)row * n2) & 3) == 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int K0 = 4 * lane + 128 * h;
      if (K0 < n2) {
        float4 o = reinterpret_cast<const float4*>(my_o)[lane + 32 * h];
        if (has_noise) {
          float4 nq;
          if (INJECT) {
            const float* nrow = nz.noise[jb] + (size_t)row * n2 + K0;
            nq.x = __ldg(nrow);
            nq.y = K0 + 1 < n2 ? __ldg(nrow + 1) : 0.f;
            nq.z = K0 + 2 < n2 ? __ldg(nrow + 2) : 0.f;
            nq.w = K0 + 3 < n2 ? __ldg(nrow + 3) : 0.f;
          } else {
            nq = philox_normal4(nz.rng[jb], (uint32_t)row * kgroups + (uint32_t)(K0 >> 2));
          }
          o.x = noise_apply(o.x, nstd, nq.x);
          o.y = noise_apply(o.y, nstd, nq.y);
          o.z = noise_apply(o.z, nstd, nq.z);
          o.w = noise_apply(o.w, nstd, nq.w);
        }
        if (vec && K0 + 3 < n2) {
          *reinterpret_cast<float4*>(orow + K0) = o;
        } else {
          orow[K0] = o.x;
          if (K0 + 1 < n2) orow[K0 + 1] = o.y;
          if (K0 + 2 < n2) orow[K0 + 2] = o.z;
          if (K0 + 3 < n2) orow[K0 + 3] = o.w;
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) cur[m] = nxt[m];
  }
}

// Generic fallback for windows wider than the register variants (very wide blurs): one thread
// per output element, taps read through L1.  View [outer][a_in][inner] -> [outer][n_out][inner].
template <bool INJECT>
__global__ void __launch_bounds__(256) sep_generic_kernel(const __grid_constant__ SepPass p, const __grid_constant__ SepNoise nz, int a_in, int inner, int last) {
  const int jb = blockIdx.y;
  const int n_out = p.n_out[jb], width = p.width[jb], outer = p.outer[jb];
  const int64_t total = (int64_t)outer * n_out * inner;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % inner);
    const int64_t oi = e / inner;
    const int I = (int)(oi % n_out), o = (int)(oi / n_out);
    const float* __restrict__ in = p.src[jb] + ((size_t)o * a_in + p.q0[jb][I]) * inner + c;
    const float* __restrict__ w = p.w[jb] + I * width;
    float acc = 0.f;
    for (int t = 0; t < width; ++t) acc = __fmaf_rn(__ldg(w + t), __ldg(in + (size_t)t * inner), acc);
    if (last && nz.has[jb]) acc = noise_apply(acc, nz.std[jb], noise_normal(nz, jb, o, I, n_out, INJECT));  // last pass: inner == 1, o = row
    p.dst[jb][e] = acc;
  }
}

// Builds the banded rows on the device: row I = w_f*taps(. - f_I) + w_c*taps(. - c_I) restricted
// to the window [q0, q0+width) (taps falling outside [0, n_in) are the zero padding).
struct ComposeBatch {
  fsg_sepcompose_job j[3 * FSG_MAX_JOBS];
};
__global__ void __launch_bounds__(128) sep_compose_kernel(const __grid_constant__ ComposeBatch b) {
  const fsg_sepcompose_job& job = b.j[blockIdx.y];
  const int nt = job.taps ? job.ntaps : 1, r = nt / 2, width = job.width;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < job.n_out; I += gridDim.x * blockDim.x) {
    int f = I, c = I;
    float wc = 0.f;
    if (job.pos) {
      const fsg_tab e = job.pos[I];
      f = e.f;
      c = e.c;
      wc = e.wc;
    }
    const bool outside = f < 0;
    const float wf = __fsub_rn(1.0f, wc);
    int q0 = outside ? (2 * I < job.n_out ? 0 : job.n_in - width) : min(max(f - r, 0), job.n_in - width);
    job.q0_out[I] = (int16_t)q0;
    for (int s = 0; s < width; ++s) {
      const int q = q0 + s;
      float w = 0.f;
      if (!outside) {
        const int t1 = q - (f - r), t2 = q - (c - r);
        if (t1 >= 0 && t1 < nt) w = __fmul_rn(wf, job.taps ? job.taps[t1] : 1.0f);
        if (t2 >= 0 && t2 < nt) w = __fmaf_rn(wc, job.taps ? job.taps[t2] : 1.0f, w);
      }
      job.w_out[I * width + s] = w;
    }
  }
}

template <int W, int VEC>
static void launch_stream(const SepPass& p, int njobs, int a_in, int inner, int max_nout, int max_outer, cudaStream_t s) {
  const size_t smem = (size_t)max_nout * (W + 1) * sizeof(float);
  auto k = sep_stream_kernel<W, VEC>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int chunks = ((inner / VEC) + SEP_THREADS - 1) / SEP_THREADS;
  int64_t want = (int64_t)max_outer * chunks;
  const int cap = 148 * 12;
  const unsigned gx = (unsigned)(want < cap ? want : cap);
  k<<<dim3(gx, njobs), SEP_THREADS, smem, s>>>(p, a_in, inner);
}

template <int W>
static void launch_z(const SepPass& p, const SepNoise& nz, int njobs, int a_in, int max_rows, bool inject, cudaStream_t s) {
  const size_t smem = (size_t)SEPZ_ROWS * (a_in + W) * sizeof(float);
  const unsigned gx = (unsigned)((max_rows + SEPZ_ROWS - 1) / SEPZ_ROWS);
  if (inject) {
    auto k = sep_z_kernel<W, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<dim3(gx, njobs), SEPZ_THREADS, smem, s>>>(p, nz, a_in);
  } else {
    auto k = sep_z_kernel<W, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<dim3(gx, njobs), SEPZ_THREADS, smem, s>>>(p, nz, a_in);
  }
}

}  // namespace fsg

using namespace fsg;

extern "C" int fsg_sep_compose(const fsg_sepcompose_job* jobs, int njobs, void* stream) {
  FSG_REQUIRE(jobs != nullptr, "fsg_sep_compose: jobs pointer is NULL");
  FSG_REQUIRE(njobs >= 1 && njobs <= 3 * FSG_MAX_JOBS, "fsg_sep_compose: njobs=%d outside [1,%d]", njobs, 3 * FSG_MAX_JOBS);
  ComposeBatch b;
  memset(&b, 0, sizeof(b));
  int max_n = 1;
  for (int i = 0; i < njobs; ++i) {
    const fsg_sepcompose_job& j = jobs[i];
    FSG_REQUIRE(j.q0_out && j.w_out, "fsg_sep_compose: job %d has a NULL output", i);
    FSG_REQUIRE(j.n_in >= 1 && j.n_in <= 32767 && j.n_out >= 1 && j.n_out <= 32767, "fsg_sep_compose: job %d bad extents", i);
    FSG_REQUIRE(j.pos || j.n_out == j.n_in, "fsg_sep_compose: job %d identity positions need n_out == n_in", i);
    const int nt = j.taps ? j.ntaps : 1;
    FSG_REQUIRE(nt >= 1 && nt % 2 == 1 && nt <= FSG_MAX_TAPS, "fsg_sep_compose: job %d needs an odd tap count <= %d", i, FSG_MAX_TAPS);
    const int need = nt + (j.pos ? 1 : 0);
    FSG_REQUIRE(j.width == (need < j.n_in ? need : j.n_in), "fsg_sep_compose: job %d width must be min(n_in, ntaps + (pos != NULL))", i);
    FSG_REQUIRE(j.cap_q0 >= j.n_out && (int64_t)j.cap_w >= (int64_t)j.n_out * j.width, "fsg_sep_compose: job %d: tables of %d x %d entries do not fit the workspace (%d / %d)", i, j.n_out,
                j.width, j.cap_q0, j.cap_w);
    b.j[i] = j;
    if (j.n_out > max_n) max_n = j.n_out;
  }
  sep_compose_kernel<<<dim3((max_n + 127) / 128, njobs), 128, 0, as_stream(stream)>>>(b);
  return check_launch("fsg_sep_compose");
}

extern "C" int fsg_sepconv(const fsg_sepconv_job* jobs, int njobs, int sx, int sy, int sz, void* stream) {
  FSG_REQUIRE(jobs != nullptr, "fsg_sepconv: jobs pointer is NULL");
  FSG_REQUIRE(njobs >= 1 && njobs <= FSG_MAX_JOBS, "fsg_sepconv: njobs=%d outside [1,%d]", njobs, FSG_MAX_JOBS);
  FSG_REQUIRE(sx >= 1 && sy >= 1 && sz >= 1 && sx <= 32767 && sy <= 32767 && sz <= 32767, "fsg_sepconv: bad shape");
  cudaStream_t s = as_stream(stream);
  const int n_in[3] = {sx, sy, sz};
  bool inject = false;
  for (int i = 0; i < njobs; ++i) {
    const fsg_sepconv_job& j = jobs[i];
    FSG_REQUIRE(j.src && j.dst && j.tmp1 && j.tmp2, "fsg_sepconv: job %d has a NULL buffer", i);
    FSG_REQUIRE(j.src != j.dst && j.src != j.tmp1 && j.tmp1 != j.tmp2 && j.tmp2 != j.dst, "fsg_sepconv: job %d: adjacent passes must not alias", i);
    // an axis may be up-sampled (n_out > n_in: simulated spacing finer than the input resolution), so every pass is checked against its buffer
    FSG_REQUIRE((int64_t)j.ax[0].n_out * sy * sz <= j.cap_tmp1, "fsg_sepconv: job %d: the x pass result %dx%dx%d does not fit tmp1 (%lld floats)", i, j.ax[0].n_out, sy, sz, (long long)j.cap_tmp1);
    FSG_REQUIRE((int64_t)j.ax[0].n_out * j.ax[1].n_out * sz <= j.cap_tmp2, "fsg_sepconv: job %d: the y pass result %dx%dx%d does not fit tmp2 (%lld floats)", i, j.ax[0].n_out, j.ax[1].n_out, sz,
                (long long)j.cap_tmp2);
    FSG_REQUIRE((int64_t)j.ax[0].n_out * j.ax[1].n_out * j.ax[2].n_out <= j.cap_dst, "fsg_sepconv: job %d: the output %dx%dx%d does not fit dst (%lld floats)", i, j.ax[0].n_out, j.ax[1].n_out,
                j.ax[2].n_out, (long long)j.cap_dst);
    for (int a = 0; a < 3; ++a) {
      const fsg_sepaxis& ax = j.ax[a];
      FSG_REQUIRE(ax.q0 && ax.w, "fsg_sepconv: job %d axis %d has a NULL table", i, a);
      FSG_REQUIRE(ax.n_out >= 1 && ax.n_out <= 32767, "fsg_sepconv: job %d axis %d bad n_out", i, a);
      FSG_REQUIRE(ax.width >= 1 && ax.width <= n_in[a] && ax.width <= 2 * FSG_MAX_TAPS, "fsg_sepconv: job %d axis %d width %d outside [1,min(%d,%d)]", i, a, ax.width, n_in[a], 2 * FSG_MAX_TAPS);
    }
    if (j.has_noise && j.noise) inject = true;
  }
  for (int i = 0; i < njobs; ++i) FSG_REQUIRE(!jobs[i].has_noise || ((jobs[i].noise != nullptr) == inject), "fsg_sepconv: jobs mix injected and Philox noise");

  SepNoise nz;
  memset(&nz, 0, sizeof(nz));
  for (int i = 0; i < njobs; ++i) {
    nz.noise[i] = jobs[i].noise;
    nz.rng[i] = jobs[i].rng;
    nz.std[i] = jobs[i].noise_std;
    nz.has[i] = jobs[i].has_noise;
  }
  SepNoise none;
  memset(&none, 0, sizeof(none));

  // ---- x and y passes fused (sep_xy_kernel) when every window fits 16 slots and a row is a whole number of 32-float chunks
  int first_axis = 0;
  {
    bool fuse = config().sep_xy && sz % 32 == 0;
    int max_n0 = 1, max_n1 = 1;
    for (int i = 0; i < njobs && fuse; ++i) {
      const fsg_sepconv_job& j = jobs[i];
      fuse = j.ax[0].width <= XY_W && j.ax[1].width <= XY_W;
      max_n0 = std::max(max_n0, j.ax[0].n_out);
      max_n1 = std::max(max_n1, j.ax[1].n_out);
    }
    if (fuse) {
      SepXY q;
      memset(&q, 0, sizeof(q));
      int64_t tiles = 0;
      for (int i = 0; i < njobs; ++i) {
        const fsg_sepconv_job& j = jobs[i];
        q.src[i] = j.src, q.dst[i] = j.tmp2;
        q.q0x[i] = j.ax[0].q0, q.wx[i] = j.ax[0].w, q.n0[i] = j.ax[0].n_out, q.widthx[i] = j.ax[0].width;
        q.q0y[i] = j.ax[1].q0, q.wy[i] = j.ax[1].w, q.n1[i] = j.ax[1].n_out, q.widthy[i] = j.ax[1].width;
        tiles += (int64_t)(sz / 32) * ((j.ax[1].n_out + XY_TJ * XY_WARPS - 1) / (XY_TJ * XY_WARPS));
      }
      // x segments: enough blocks for whole waves of 3 resident blocks per SM; every segment re-reads 15 planes
      const int slots = 148 * 3;
      int xsegs = 1;
      double best = 1e30;
      for (int c = 1; c <= 8 && c <= max_n0; ++c) {
        const double waves = std::ceil((double)tiles * c / slots);
        const double cost = waves * ((double)sx / c + (XY_W - 1));
        if (cost < best) best = cost, xsegs = c;
      }
      const int seg_cap = ((max_n0 + xsegs - 1) / xsegs + 3) / 4 * 4;  // multiple of 4: keeps the tables behind it 16-byte aligned
      const size_t smem = sizeof(float) * ((size_t)XY_WARPS * XY_W * XY_TJ * 32 + (size_t)seg_cap * (XY_W + 1) + (size_t)XY_TJ * XY_WARPS * (XY_W + 1));
      if (smem <= 200 * 1024) {
        const int nyg = (max_n1 + XY_TJ * XY_WARPS - 1) / (XY_TJ * XY_WARPS);
        cudaFuncSetAttribute(sep_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sep_xy_kernel<<<dim3((unsigned)((sz / 32) * nyg * xsegs), njobs), XY_WARPS * 32, smem, s>>>(q, sx, sy, sz, xsegs, seg_cap);
        first_axis = 2;
      }
    }
  }

  for (int a = first_axis; a < 3; ++a) {
    SepPass p;
    memset(&p, 0, sizeof(p));
    int maxw = 0, max_nout = 0, max_outer = 0;
    for (int i = 0; i < njobs; ++i) {
      const fsg_sepconv_job& j = jobs[i];
      p.src[i] = a == 0 ? j.src : (a == 1 ? j.tmp1 : j.tmp2);
      p.dst[i] = a == 0 ? j.tmp1 : (a == 1 ? j.tmp2 : j.dst);
      p.q0[i] = j.ax[a].q0;
      p.w[i] = j.ax[a].w;
      p.n_out[i] = j.ax[a].n_out;
      p.width[i] = j.ax[a].width;
      p.outer[i] = a == 0 ? 1 : (a == 1 ? j.ax[0].n_out : j.ax[0].n_out * j.ax[1].n_out);
      maxw = p.width[i] > maxw ? p.width[i] : maxw;
      max_nout = p.n_out[i] > max_nout ? p.n_out[i] : max_nout;
      max_outer = p.outer[i] > max_outer ? p.outer[i] : max_outer;
    }
    const int a_in = n_in[a];
    if (a < 2) {
      const int inner = a == 0 ? sy * sz : sz;
      bool vec2 = (inner % 2 == 0);
      for (int i = 0; i < njobs && vec2; ++i) vec2 = ((reinterpret_cast<uintptr_t>(p.src[i]) | reinterpret_cast<uintptr_t>(p.dst[i])) & 7) == 0;
      if (maxw <= 8) {
        if (vec2) launch_stream<8, 2>(p, njobs, a_in, inner, max_nout, max_outer, s);
        else launch_stream<8, 1>(p, njobs, a_in, inner, max_nout, max_outer, s);
      } else if (maxw <= 16) {
        if (vec2) launch_stream<16, 2>(p, njobs, a_in, inner, max_nout, max_outer, s);
        else launch_stream<16, 1>(p, njobs, a_in, inner, max_nout, max_outer, s);
      } else if (maxw <= 32) {
        launch_stream<32, 1>(p, njobs, a_in, inner, max_nout, max_outer, s);
      } else {
        int64_t tot = (int64_t)max_outer * max_nout * inner;
        int64_t want = (tot + 255) / 256;
        sep_generic_kernel<false><<<dim3((unsigned)(want < 148 * 32 ? want : 148 * 32), njobs), 256, 0, s>>>(p, none, a_in, inner, 0);
      }
    } else {
      // rows of <= 256 samples whose blur fits 13 taps, described by `pos` / `taps`: warp-per-row register kernel
      bool zrow = sz <= ZR_ROW && (sz & 3) == 0;
      for (int i = 0; i < njobs && zrow; ++i) {
        const fsg_sepaxis& z = jobs[i].ax[2];
        zrow = z.ntaps >= 1 && z.ntaps <= 2 * ZR_R + 1 && (z.ntaps & 1) && z.n_out <= ZR_ROW && (z.pos != nullptr || z.n_out == sz) && (reinterpret_cast<uintptr_t>(p.src[i]) & 15) == 0;
      }
      if (zrow) {
        SepZRow zr;
        memset(&zr, 0, sizeof(zr));
        for (int i = 0; i < njobs; ++i) {
          zr.src[i] = p.src[i];
          zr.dst[i] = p.dst[i];
          zr.pos[i] = jobs[i].ax[2].pos;
          zr.taps[i] = jobs[i].ax[2].taps;
          zr.ntaps[i] = jobs[i].ax[2].ntaps;
          zr.n_out[i] = p.n_out[i];
          zr.rows[i] = p.outer[i];
        }
        const int want = (max_outer + ZR_THREADS / 32 - 1) / (ZR_THREADS / 32);
        const int cap = 148 * 3 * 2;  // 3 resident blocks per SM, two waves
        const dim3 grid((unsigned)(want < cap ? want : cap), njobs);
        if (inject)
          sep_zrow_kernel<true><<<grid, ZR_THREADS, 0, s>>>(zr, nz, sz);
        else
          sep_zrow_kernel<false><<<grid, ZR_THREADS, 0, s>>>(zr, nz, sz);
      } else if (maxw <= 8 && a_in >= 8) {
        launch_z<8>(p, nz, njobs, a_in, max_outer, inject, s);
      } else if (maxw <= 16 && a_in >= 16) {
        launch_z<16>(p, nz, njobs, a_in, max_outer, inject, s);
      } else if (maxw <= 32 && a_in >= 32) {
        launch_z<32>(p, nz, njobs, a_in, max_outer, inject, s);
      } else {
        int64_t tot = (int64_t)max_outer * max_nout;
        int64_t want = (tot + 255) / 256;
        dim3 grid((unsigned)(want < 148 * 32 ? want : 148 * 32), njobs);
        if (inject)
          sep_generic_kernel<true><<<grid, 256, 0, s>>>(p, nz, a_in, 1, 1);
        else
          sep_generic_kernel<false><<<grid, 256, 0, s>>>(p, nz, a_in, 1, 1);
      }
    }
  }
  return check_launch("fsg_sepconv");
}
