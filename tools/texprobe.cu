// Probe (not part of the product): how fast can the eight trilinear corners of a rotated 256^3 resample be
// fetched on B200 through (A) plain global loads, (B) point-sampled tex3D on a block-linear 3-D array,
// (C) two 2x2 texture gathers (tld4) on a layered 2-D array, (D) one hardware-filtered tex3D (9-bit weights: a
// rate reference only, not accurate enough for the parity tolerance), (E) tld4 on pitch-linear 2-D textures.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/texprobe tools/texprobe.cu ; run on the GPU box.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

constexpr int S = 256;

struct Xf {
  float m[9];
  float c;
  float amp;
};

__device__ __forceinline__ void coords(const Xf& t, int x, int y, int z, float& px, float& py, float& pz) {
  float fx = x - t.c, fy = y - t.c, fz = z - t.c;
  px = t.m[0] * fx + t.m[1] * fy + t.m[2] * fz + t.c + t.amp * __sinf(0.05f * y);
  py = t.m[3] * fx + t.m[4] * fy + t.m[5] * fz + t.c + t.amp * __sinf(0.04f * z);
  pz = t.m[6] * fx + t.m[7] * fy + t.m[8] * fz + t.c + t.amp * __sinf(0.03f * x);
  px = fminf(fmaxf(px, 0.f), S - 1.001f);
  py = fminf(fmaxf(py, 0.f), S - 1.001f);
  pz = fminf(fmaxf(pz, 0.f), S - 1.001f);
}

__device__ __forceinline__ float blend(float v000, float v001, float v010, float v011, float v100, float v101, float v110,
                                       float v111, float tx, float ty, float tz) {
  float a = v000 + tz * (v001 - v000), b = v010 + tz * (v011 - v010);
  float c = v100 + tz * (v101 - v100), d = v110 + tz * (v111 - v110);
  float e = a + ty * (b - a), f = c + ty * (d - c);
  return e + tx * (f - e);
}

__global__ void __launch_bounds__(256) k_ldg(const float* __restrict__ src, float* __restrict__ out, Xf t) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float px, py, pz;
  coords(t, x, y, z, px, py, pz);
  int ix = (int)px, iy = (int)py, iz = (int)pz;
  float tx = px - ix, ty = py - iy, tz = pz - iz;
  const float* p = src + ((size_t)ix * S + iy) * S + iz;
  out[i] = blend(p[0], p[1], p[S], p[S + 1], p[S * S], p[S * S + 1], p[S * S + S], p[S * S + S + 1], tx, ty, tz);
}

__global__ void __launch_bounds__(256) k_tex3d_point(cudaTextureObject_t tex, float* __restrict__ out, Xf t) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float px, py, pz;
  coords(t, x, y, z, px, py, pz);
  int ix = (int)px, iy = (int)py, iz = (int)pz;
  float tx = px - ix, ty = py - iy, tz = pz - iz;
  // texture axes: width = z, height = y, depth = x
  float u = iz + 0.5f, v = iy + 0.5f, w = ix + 0.5f;
  float v000 = tex3D<float>(tex, u, v, w), v001 = tex3D<float>(tex, u + 1, v, w);
  float v010 = tex3D<float>(tex, u, v + 1, w), v011 = tex3D<float>(tex, u + 1, v + 1, w);
  float v100 = tex3D<float>(tex, u, v, w + 1), v101 = tex3D<float>(tex, u + 1, v, w + 1);
  float v110 = tex3D<float>(tex, u, v + 1, w + 1), v111 = tex3D<float>(tex, u + 1, v + 1, w + 1);
  out[i] = blend(v000, v001, v010, v011, v100, v101, v110, v111, tx, ty, tz);
}

__global__ void __launch_bounds__(256) k_tex3d_linear(cudaTextureObject_t tex, float* __restrict__ out, Xf t) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float px, py, pz;
  coords(t, x, y, z, px, py, pz);
  out[i] = tex3D<float>(tex, pz + 0.5f, py + 0.5f, px + 0.5f);
}

// tld4 on a layered 2-D array: layer = x plane, (u, v) = (z, y). Returns the 2x2 footprint around (u, v):
// order (x,y+1), (x+1,y+1), (x+1,y), (x,y) in texel space = w, z(=...) per the PTX manual: comps are
// (i0,j1), (i1,j1), (i1,j0), (i0,j0).
__device__ __forceinline__ float4 gather_a2d(cudaTextureObject_t tex, int layer, float u, float v) {
  float4 r;
  asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%8}];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(tex), "r"(layer), "f"(u), "f"(v), "f"(0.f));
  return r;
}

__global__ void __launch_bounds__(256) k_gather_layered(cudaTextureObject_t tex, float* __restrict__ out, Xf t) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float px, py, pz;
  coords(t, x, y, z, px, py, pz);
  int ix = (int)px, iy = (int)py, iz = (int)pz;
  float tx = px - ix, ty = py - iy, tz = pz - iz;
  float u = iz + 1.0f, v = iy + 1.0f;  // footprint (iz, iz+1) x (iy, iy+1)
  float4 a = gather_a2d(tex, ix, u, v), b = gather_a2d(tex, ix + 1, u, v);
  // a.w = (iz, iy), a.z = (iz+1, iy), a.x = (iz, iy+1), a.y = (iz+1, iy+1)
  out[i] = blend(a.w, a.z, a.x, a.y, b.w, b.z, b.x, b.y, tx, ty, tz);
}

__global__ void __launch_bounds__(256) k_gather_pitch(cudaTextureObject_t t0, cudaTextureObject_t t1, float* __restrict__ out, Xf t) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float px, py, pz;
  coords(t, x, y, z, px, py, pz);
  int ix = (int)px, iy = (int)py, iz = (int)pz;
  float tx = px - ix, ty = py - iy, tz = pz - iz;
  // two textures of S/2 planes each, rows = plane * S + y
  float u = iz + 1.0f;
  int r0 = ix * S + iy, r1 = (ix + 1) * S + iy;
  const int half = S / 2 * S;
  float4 a = r0 < half ? tex2Dgather<float4>(t0, u, r0 + 1.0f, 0) : tex2Dgather<float4>(t1, u, r0 - half + 1.0f, 0);
  float4 b = r1 < half ? tex2Dgather<float4>(t0, u, r1 + 1.0f, 0) : tex2Dgather<float4>(t1, u, r1 - half + 1.0f, 0);
  out[i] = blend(a.w, a.z, a.x, a.y, b.w, b.z, b.x, b.y, tx, ty, tz);
}

// ---- store side: what fsg_gmm would do (16 bytes per thread along z) into linear memory or into the layered array
__global__ void __launch_bounds__(256) k_store_lin(float* __restrict__ out, float seed) {
  int i = blockIdx.x * 256 + threadIdx.x;  // float4 index
  float b = seed + i;
  reinterpret_cast<float4*>(out)[i] = make_float4(b, b + 1, b + 2, b + 3);
}
__global__ void __launch_bounds__(256) k_store_surf(cudaSurfaceObject_t surf, float seed) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z4 = i % (S / 4), y = (i / (S / 4)) % S, x = i / (S / 4 * S);
  float b = seed + i;
  surf2DLayeredwrite<float4>(make_float4(b, b + 1, b + 2, b + 3), surf, z4 * 16, y, x);
}
// same bytes, but each warp writes one whole 64-byte x 8-row tile (a GOB of the block-linear layout)
__global__ void __launch_bounds__(256) k_store_surf_gob(cudaSurfaceObject_t surf, float seed) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int lane = i & 31, tile = i >> 5;
  int zc = tile % (S / 16), y8 = (tile / (S / 16)) % (S / 8), x = tile / (S / 16 * (S / 8));
  int y = y8 * 8 + (lane >> 2), z4 = zc * 4 + (lane & 3);
  int g = (x * S + y) * (S / 4) + z4;  // the float4 index k_store_surf gives this voxel group
  float b = seed + g;
  surf2DLayeredwrite<float4>(make_float4(b, b + 1, b + 2, b + 3), surf, z4 * 16, y, x);
}
__global__ void __launch_bounds__(256) k_check_surf(cudaTextureObject_t tex, float seed, int* bad) {
  int i = blockIdx.x * 256 + threadIdx.x;
  int z = i % S, y = (i / S) % S, x = i / (S * S);
  float v = tex2DLayered<float>(tex, z + 0.5f, y + 0.5f, x);
  int i4 = i / 4;
  if (v != seed + i4 + (i & 3)) atomicAdd(bad, 1);
}

static Xf make_xf(float deg) {
  float a = deg * 3.14159265f / 180.f, b = 0.7f * a, c = -0.5f * a;
  float ca = cosf(a), sa = sinf(a), cb = cosf(b), sb = sinf(b), cc = cosf(c), sc = sinf(c);
  // Rz(c) * Ry(b) * Rx(a)
  float rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca}, ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb}, rz[9] = {cc, -sc, 0, sc, cc, 0, 0, 0, 1};
  float t1[9], m[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      t1[i * 3 + j] = 0;
      for (int k = 0; k < 3; k++) t1[i * 3 + j] += ry[i * 3 + k] * rx[k * 3 + j];
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      m[i * 3 + j] = 0;
      for (int k = 0; k < 3; k++) m[i * 3 + j] += rz[i * 3 + k] * t1[k * 3 + j];
    }
  Xf t;
  for (int i = 0; i < 9; i++) t.m[i] = m[i];
  t.c = (S - 1) / 2.f;
  t.amp = 3.f;
  return t;
}

template <class F>
static float time_ms(F f, int reps = 20) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; i++) f();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

static double max_diff(const float* da, const float* db, size_t n) {
  std::vector<float> a(n), b(n);
  cudaMemcpy(a.data(), da, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(b.data(), db, n * 4, cudaMemcpyDeviceToHost);
  double m = 0;
  for (size_t i = 0; i < n; i++) m = fmax(m, fabs((double)a[i] - b[i]));
  return m;
}

int main() {
  const size_t N = (size_t)S * S * S;
  std::vector<float> h(N);
  srand(1);
  for (size_t i = 0; i < N; i++) h[i] = (float)(rand() & 0xffff) / 256.f;
  // NV volumes round-robin so that the working set exceeds L2 like the production batch (8 x 64 MB)
  constexpr int NV = 4;
  float* src[NV];
  float *out, *ref;
  for (int v = 0; v < NV; v++) {
    CK(cudaMalloc(&src[v], N * 4));
    CK(cudaMemcpy(src[v], h.data(), N * 4, cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&out, N * 4));
  CK(cudaMalloc(&ref, N * 4));
  int grid = (int)(N / 256);

  cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
  cudaArray_t arr3[NV], arrl[NV];
  cudaTextureObject_t tex3p[NV], tex3l[NV], texl[NV], texp0[NV], texp1[NV];
  bool layered_ok = true, pitch_ok = true;
  for (int v = 0; v < NV; v++) {
    CK(cudaMalloc3DArray(&arr3[v], &cd, make_cudaExtent(S, S, S), 0));
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr(src[v], S * 4, S, S);
    p.dstArray = arr3[v];
    p.extent = make_cudaExtent(S, S, S);
    p.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&p));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr3[v];
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&tex3p[v], &rd, &td, nullptr));
    td.filterMode = cudaFilterModeLinear;
    CK(cudaCreateTextureObject(&tex3l[v], &rd, &td, nullptr));
    // layered 2-D + gather
    cudaError_t e = cudaMalloc3DArray(&arrl[v], &cd, make_cudaExtent(S, S, S), cudaArrayLayered | cudaArrayTextureGather | cudaArraySurfaceLoadStore);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc3DArray(&arrl[v], &cd, make_cudaExtent(S, S, S), cudaArrayLayered | cudaArraySurfaceLoadStore);
      if (v == 0) printf("layered+gather flag refused, plain layered: %s\n", cudaGetErrorString(e));
    }
    if (e != cudaSuccess) {
      layered_ok = false;
      cudaGetLastError();
    } else {
      p.dstArray = arrl[v];
      CK(cudaMemcpy3D(&p));
      rd.res.array.array = arrl[v];
      td.filterMode = cudaFilterModePoint;
      e = cudaCreateTextureObject(&texl[v], &rd, &td, nullptr);
      if (e != cudaSuccess) {
        layered_ok = false;
        printf("layered texture object: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
    // pitch-linear halves
    cudaResourceDesc rp = {};
    rp.resType = cudaResourceTypePitch2D;
    rp.res.pitch2D.desc = cd;
    rp.res.pitch2D.width = S;
    rp.res.pitch2D.height = S / 2 * S;
    rp.res.pitch2D.pitchInBytes = S * 4;
    rp.res.pitch2D.devPtr = src[v];
    td.filterMode = cudaFilterModePoint;
    e = cudaCreateTextureObject(&texp0[v], &rp, &td, nullptr);
    rp.res.pitch2D.devPtr = src[v] + N / 2;
    if (e == cudaSuccess) e = cudaCreateTextureObject(&texp1[v], &rp, &td, nullptr);
    if (e != cudaSuccess) {
      pitch_ok = false;
      if (v == 0) printf("pitch2D texture: %s\n", cudaGetErrorString(e));
      cudaGetLastError();
    }
  }

  if (layered_ok) {
    cudaSurfaceObject_t surf[NV];
    bool ok = true;
    for (int v = 0; v < NV; v++) {
      cudaResourceDesc rd = {};
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = arrl[v];
      cudaError_t e = cudaCreateSurfaceObject(&surf[v], &rd);
      if (e != cudaSuccess) {
        printf("surface object: %s (array needs cudaArraySurfaceLoadStore)\n", cudaGetErrorString(e));
        cudaGetLastError();
        ok = false;
        break;
      }
    }
    if (ok) {
      int v = 0;
      int g4 = (int)(N / 4 / 256);
      float ms = time_ms([&] { k_store_lin<<<g4, 256>>>(src[v++ % NV], 1.f); });
      printf("store linear float4  %.4f ms (%.0f GB/s)\n", ms, N * 4 / ms * 1e-6);
      ms = time_ms([&] { k_store_surf<<<g4, 256>>>(surf[v++ % NV], 1.f); });
      cudaError_t e = cudaDeviceSynchronize();
      printf("store surface float4 %.4f ms (%.0f GB/s) %s\n", ms, N * 4 / ms * 1e-6, cudaGetErrorString(e));
      ms = time_ms([&] { k_store_surf_gob<<<g4, 256>>>(surf[v++ % NV], 1.f); });
      e = cudaDeviceSynchronize();
      printf("store surface float4, one GOB per warp %.4f ms (%.0f GB/s) %s\n", ms, N * 4 / ms * 1e-6, cudaGetErrorString(e));
      int* bad;
      cudaMalloc(&bad, 4);
      cudaMemset(bad, 0, 4);
      k_store_surf_gob<<<g4, 256>>>(surf[0], 7.f);
      k_check_surf<<<grid, 256>>>(texl[0], 7.f, bad);
      int hb = -1;
      cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
      printf("surface store readback mismatches: %d\n", hb);
      // restore the data the gather tests expect
      for (int q = 0; q < NV; q++) {
        cudaMemcpy(src[q], h.data(), N * 4, cudaMemcpyHostToDevice);
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(src[q], S * 4, S, S);
        p.dstArray = arrl[q];
        p.extent = make_cudaExtent(S, S, S);
        p.kind = cudaMemcpyDeviceToDevice;
        cudaMemcpy3D(&p);
      }
    }
  }

  for (float deg : {0.f, 5.f, 10.f, 20.f}) {
    Xf t = make_xf(deg);
    int v = 0;
    k_ldg<<<grid, 256>>>(src[0], ref, t);
    CK(cudaDeviceSynchronize());
    float ms = time_ms([&] { k_ldg<<<grid, 256>>>(src[v++ % NV], out, t); });
    printf("deg %4.1f  ldg            %.4f ms  (%.1f Gvox/s)\n", deg, ms, N / ms * 1e-6);
    ms = time_ms([&] { k_tex3d_point<<<grid, 256>>>(tex3p[v++ % NV], out, t); });
    CK(cudaDeviceSynchronize());
    printf("deg %4.1f  tex3D point x8 %.4f ms  (%.1f Gvox/s)  maxdiff %.3g\n", deg, ms, N / ms * 1e-6, max_diff(out, ref, N));
    ms = time_ms([&] { k_tex3d_linear<<<grid, 256>>>(tex3l[v++ % NV], out, t); });
    CK(cudaDeviceSynchronize());
    printf("deg %4.1f  tex3D linear   %.4f ms  (%.1f Gvox/s)  maxdiff %.3g\n", deg, ms, N / ms * 1e-6, max_diff(out, ref, N));
    if (layered_ok) {
      ms = time_ms([&] { k_gather_layered<<<grid, 256>>>(texl[v++ % NV], out, t); });
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("layered gather failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      printf("deg %4.1f  tld4 layered   %.4f ms  (%.1f Gvox/s)  maxdiff %.3g\n", deg, ms, N / ms * 1e-6, max_diff(out, ref, N));
    }
    if (pitch_ok) {
      ms = time_ms([&] { int k = v++ % NV; k_gather_pitch<<<grid, 256>>>(texp0[k], texp1[k], out, t); });
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("pitch gather failed: %s\n", cudaGetErrorString(e));
        return 1;
      }
      printf("deg %4.1f  tld4 pitch2D   %.4f ms  (%.1f Gvox/s)  maxdiff %.3g\n", deg, ms, N / ms * 1e-6, max_diff(out, ref, N));
    }
  }
  return 0;
}
