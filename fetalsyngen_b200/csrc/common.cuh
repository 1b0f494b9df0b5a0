// Shared device/host helpers for libfsg (sm_100a).  Compiled with -fmad=false: the parity
// contract needs every float32 multiply/add rounded separately, as torch eager does.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "fsg.h"

namespace fsg {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define FSG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      fsg::set_error(__VA_ARGS__);        \
      return 1;                           \
    }                                     \
  } while (0)

// A/B switches of DESIGN.md section 4, read from the environment ONCE when the library is loaded (core.cu);
// launch wrappers only look at this struct.
struct Config {
  bool warp_tile;         // FSG_WARP_TILE=1: r01 TMA-staged 16^3 tile kernel for the warp (single buffered)
  bool warp_pipe;         // FSG_WARP_PIPE=0/1: pipelined TMA kernel (persistent, producer / consumer warps)
  int fwd_warp_min_taps;  // FSG_FWD_WARP_MIN_TAPS: PSF size from which the acquisition splits taps over lanes
  bool fwd_lean;          // FSG_FWD_LEAN=1: nested-lerp accumulation in the acquisition
  bool adj_lean;          // FSG_ADJ_LEAN=0: reference operation order in the PSF reconstruction
  bool adj_thread;        // FSG_ADJ_THREAD=1: thread-per-pixel PSF reconstruction
  int tile_debug;         // FSG_TILE_DEBUG
  bool sep_xy;            // FSG_SEP_XY=1: x and y passes of fsg_sepconv fused in one kernel (sep_xy_kernel)
};
const Config& config();

// Jobs travel by value in the kernel parameter space (<= 32 KB on sm_70+ with CUDA 12.1+).
template <typename Job>
struct Batch {
  Job j[FSG_MAX_JOBS];
};

template <typename Job>
inline int fill_batch(Batch<Job>& b, const Job* jobs, int njobs) {
  FSG_REQUIRE(jobs != nullptr, "jobs pointer is NULL");
  FSG_REQUIRE(njobs >= 1 && njobs <= FSG_MAX_JOBS, "njobs=%d outside [1,%d]", njobs, FSG_MAX_JOBS);
  memset(&b, 0, sizeof(b));
  memcpy(b.j, jobs, sizeof(Job) * njobs);
  return 0;
}

// ------------------------------------------------------------------ exact float32 helpers
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
// w_f*a + w_c*b with three roundings (utils/generation.py:270-278, :377-386)
__device__ __forceinline__ float blend(float wf, float a, float wc, float b) {
  return __fadd_rn(__fmul_rn(wf, a), __fmul_rn(wc, b));
}
// a*w_f + b*w_c: same value, operand order of fast_3D_interp_torch
__device__ __forceinline__ float lerp2(float a, float wf, float b, float wc) {
  return __fadd_rn(__fmul_rn(a, wf), __fmul_rn(b, wc));
}

// Packed FP32x2 (sm_100: FADD2 / FMUL2 / FFMA2 take one issue slot for two lanes).  add/sub are
// individually rounded like their scalar forms.  ptxas contracts mul.f32x2 feeding add.f32x2 into
// FFMA2 even under --fmad=false, so the bit-exact coordinate chain keeps its products scalar
// (FMUL writes straight into the halves of a register pair) and packs only the adds; mul2 / fma2
// are used on the tolerance-checked image path.
typedef unsigned long long P2;
__device__ __forceinline__ P2 pk(float a, float b) {
  P2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk(P2 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ P2 add2(P2 a, P2 b) {
  P2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ P2 add2_rz(P2 a, P2 b) {
  P2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ P2 sub2(P2 a, P2 b) {
  P2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ P2 mul2(P2 a, P2 b) {
  P2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) {
  P2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ------------------------------------------------------------------ ordered float atomics
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ Philox4x32-10
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t lo0 = 0xD2511F53u * c0, hi0 = __umulhi(0xD2511F53u, c0);
      const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = __umulhi(0xCD9E8D57u, c2);
      const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

__device__ __forceinline__ float lg2_approx_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx_ftz(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Four standard normals for counter block `blk` of stream (stage, sample): Box-Muller on the
// two word pairs.  Uniforms are built from mantissa bits (no int->float conversion):
// u1 = 2 - [1,2) in (0,1] with 23 bits, u2 = [1,2) - 1 in [0,1); ln via lg2.approx.
__device__ __forceinline__ float4 philox_normal4(const fsg_rng& r, uint32_t blk) {
  const Philox ph(r.seed);
  const uint4 w = ph(blk, r.stage, (uint32_t)r.sample, (uint32_t)(r.sample >> 32));
  const float u1a = 2.0f - __uint_as_float(0x3f800000u | (w.x >> 9));
  const float u1b = 2.0f - __uint_as_float(0x3f800000u | (w.z >> 9));
  const float u2a = __uint_as_float(0x3f800000u | (w.y >> 9)) - 1.0f;
  const float u2b = __uint_as_float(0x3f800000u | (w.w >> 9)) - 1.0f;
  // -2 ln u = (-2 ln 2) * log2 u
  const float ra = sqrt_approx_ftz(-1.3862943611198906f * lg2_approx_ftz(u1a));
  const float rb = sqrt_approx_ftz(-1.3862943611198906f * lg2_approx_ftz(u1b));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u2a, &sa, &ca);
  __sincosf(6.283185307179586f * u2b, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// min/max accumulators live as order-preserving ints while atomics run
static __global__ void minmax_init_kernel(float* mm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    reinterpret_cast<int*>(mm)[2 * i] = float_to_ordered(__int_as_float(0x7f800000));      // +inf
    reinterpret_cast<int*>(mm)[2 * i + 1] = float_to_ordered(__int_as_float(0xff800000));  // -inf
  }
}
static __global__ void minmax_final_kernel(float* mm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) mm[i] = ordered_to_float(reinterpret_cast<int*>(mm)[i]);
}

struct Tab {
  int f, c;
  float wc, wf;
};
__device__ __forceinline__ Tab load_tab(const fsg_tab* t, int i) {
  const fsg_tab e = t[i];
  Tab r;
  r.f = e.f;
  r.c = e.c;
  r.wc = e.wc;
  r.wf = __fsub_rn(1.0f, e.wc);
  return r;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace fsg
