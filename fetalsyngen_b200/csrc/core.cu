// libfsg core: error plumbing, ABI introspection, dtype plumbing, min/max + ScaleIntensity.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace fsg {

static thread_local char g_err[512] = "";

static Config read_config() {
  auto flag = [](const char* name, bool dflt) {
    const char* e = getenv(name);
    return e ? e[0] == '1' : dflt;
  };
  auto num = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
  };
  Config c;
  c.warp_tile = flag("FSG_WARP_TILE", false);
  c.warp_pipe = flag("FSG_WARP_PIPE", false);
  // r01: 215 taps 2.5 ms (thread) vs 3.9 ms (warp); 729 taps 12.7 vs 9.3 ms
  c.fwd_warp_min_taps = num("FSG_FWD_WARP_MIN_TAPS", 400);
  // opt-in: measured with the xy-quad volume 1.59 vs 1.27 ms (215 taps), 4.46 vs 4.62 ms (729 taps), 0.49 vs
  // 0.65 ms (37 taps) — the nested lerps are a longer dependent chain than the eight independent products
  c.fwd_lean = flag("FSG_FWD_LEAN", false);
  c.adj_lean = flag("FSG_ADJ_LEAN", true);
  c.adj_thread = flag("FSG_ADJ_THREAD", false);
  c.tile_debug = num("FSG_TILE_DEBUG", 0);
  c.sep_xy = num("FSG_SEP_XY", 0) != 0;
  return c;
}
static const Config g_config = read_config();  // static initialiser: runs at dlopen
const Config& config() { return g_config; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

// ------------------------------------------------------------------ elementwise converters
template <typename In, typename Out>
__global__ void __launch_bounds__(256) convert_kernel(const In* __restrict__ x, Out* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (Out)x[i];
}

template <typename In, typename Out>
static int convert(const In* x, Out* out, int64_t n, void* stream, const char* name) {
  FSG_REQUIRE(x && out, "%s: NULL pointer", name);
  FSG_REQUIRE(n >= 0, "%s: negative size", name);
  if (n == 0) return 0;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  convert_kernel<In, Out><<<blocks, 256, 0, as_stream(stream)>>>(x, out, n);
  return check_launch(name);
}

// ------------------------------------------------------------------ min/max + ScaleIntensity
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, int64_t n, float* mm) {
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(x4 + i);
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    lo = fminf(lo, x[i]);
    hi = fmaxf(hi, x[i]);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  __shared__ float slo[8], shi[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    slo[w] = lo;
    shi[w] = hi;
  }
  __syncthreads();
  if (w == 0) {
    lo = l < 8 ? slo[l] : __int_as_float(0x7f800000);
    hi = l < 8 ? shi[l] : __int_as_float(0xff800000);
    lo = warp_min(lo);
    hi = warp_max(hi);
    if (l == 0) {
      atomicMin(reinterpret_cast<int*>(mm), float_to_ordered(lo));
      atomicMax(reinterpret_cast<int*>(mm) + 1, float_to_ordered(hi));
    }
  }
}

// (x - min) / (max - min) * 1 + 0, monai ScaleIntensity(minv=0, maxv=1); constant image -> 0.
__global__ void __launch_bounds__(256) scale_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, const float* __restrict__ mm) {
  const float lo = mm[0], hi = mm[1];
  const float den = __fsub_rn(hi, lo);
  const bool flat = (lo == hi);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    out[i] = flat ? __fmul_rn(v, 0.0f) : __fdiv_rn(__fsub_rn(v, lo), den);
  }
}

}  // namespace fsg

using namespace fsg;

extern "C" {

int fsg_version(void) { return FSG_VERSION; }
const char* fsg_last_error(void) { return g_err; }

int fsg_sizeof(const char* name) {
  if (!name) return -1;
#define FSG_SZ(T) \
  if (strcmp(name, #T) == 0) return (int)sizeof(T);
  FSG_SZ(fsg_tab)
  FSG_SZ(fsg_rng)
  FSG_SZ(fsg_gmm_job)
  FSG_SZ(fsg_warp_job)
  FSG_SZ(fsg_blur_job)
  FSG_SZ(fsg_resample_job)
  FSG_SZ(fsg_noise_job)
  FSG_SZ(fsg_zoom_job)
  FSG_SZ(fsg_sepaxis)
  FSG_SZ(fsg_sepconv_job)
  FSG_SZ(fsg_sepcompose_job)
  FSG_SZ(fsg_sample_job)
  FSG_SZ(fsg_perlin_octave)
  FSG_SZ(fsg_grid_job)
  FSG_SZ(fsg_em_job)
  FSG_SZ(fsg_unpack_job)
  FSG_SZ(fsg_texvol)
  FSG_SZ(fsg_draw_config)
  FSG_SZ(fsg_draw_out)
  FSG_SZ(fsg_step_inputs)
  FSG_SZ(fsg_step_sample)
  FSG_SZ(fsg_step)
  FSG_SZ(fsg_step_jobs)
#undef FSG_SZ
  return -1;
}

// Small-parameter fetch: an SM copy from (mapped, pinned) host memory into the device ring.  A
// cudaMemcpyAsync would queue on the H2D copy engine behind the bulk input transfers of the next
// pipeline steps (FIFO across streams) and stall this step's kernels for tens of milliseconds.
__global__ void __launch_bounds__(256) fetch_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int n4) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n4; i += gridDim.x * 256) dst[i] = src[i];
}

int fsg_fetch_params(const float* src_host_mapped, float* dst, int64_t nfloats, void* stream) {
  FSG_REQUIRE(src_host_mapped && dst && nfloats >= 0 && nfloats < ((int64_t)1 << 30), "fsg_fetch_params: bad arguments");
  FSG_REQUIRE(((reinterpret_cast<uintptr_t>(src_host_mapped) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "fsg_fetch_params: buffers must be 16-byte aligned");
  if (nfloats == 0) return 0;
  const int n4 = (int)((nfloats + 3) / 4);
  const int want = (n4 + 255) / 256;
  fetch_kernel<<<want < 148 ? want : 148, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src_host_mapped), reinterpret_cast<float4*>(dst), n4);
  return check_launch("fsg_fetch_params");
}

int fsg_f32_to_u8(const float* x, uint8_t* out, int64_t n, void* stream) { return convert(x, out, n, stream, "fsg_f32_to_u8"); }
int fsg_u8_to_f32(const uint8_t* x, float* out, int64_t n, void* stream) { return convert(x, out, n, stream, "fsg_u8_to_f32"); }
int fsg_u8_to_i64(const uint8_t* x, int64_t* out, int64_t n, void* stream) { return convert(x, out, n, stream, "fsg_u8_to_i64"); }

int fsg_minmax(const float* x, int64_t n, float* minmax, void* stream) {
  FSG_REQUIRE(x && minmax, "fsg_minmax: NULL pointer");
  FSG_REQUIRE(n > 0, "fsg_minmax: empty input");
  FSG_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "fsg_minmax: x must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  minmax_init_kernel<<<1, 32, 0, s>>>(minmax, 1);
  const int64_t want = (n / 4 + 255) / 256 + 1;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  minmax_kernel<<<blocks, 256, 0, s>>>(x, n, minmax);
  minmax_final_kernel<<<1, 32, 0, s>>>(minmax, 1);
  return check_launch("fsg_minmax");
}

int fsg_scale_intensity(const float* x, float* out, int64_t n, const float* minmax, void* stream) {
  FSG_REQUIRE(x && out && minmax, "fsg_scale_intensity: NULL pointer");
  if (n <= 0) return 0;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
  scale_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, out, n, minmax);
  return check_launch("fsg_scale_intensity");
}

}  // extern "C"
