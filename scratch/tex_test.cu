// Micro-benchmark: trilinear 8-corner gather of a rotated grid at 256^3 x B volumes through
//  (A) LDG from linear memory, (B) tex3D point fetch from a 3-D cudaArray, (C) tex1Dfetch on linear memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch/tex_test scratch/tex_test.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

struct Aff { float a[9], c[3], ctr[3]; };

template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ src, cudaTextureObject_t tex, float* __restrict__ dst, int S, Aff A) {
  // block = 8 x 4 (x,y) columns, warp = 32 consecutive z
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int z = blockIdx.x * 32 + lane;
  const int y = blockIdx.y * 4 + (w & 3);
  const int x = blockIdx.z * 2 + (w >> 2);
  if (x >= S || y >= S || z >= S) return;
  const float xc = x - A.ctr[0], yc = y - A.ctr[1], zc = z - A.ctr[2];
  float fx = A.a[0] * xc + A.a[1] * yc + A.a[2] * zc + A.c[0];
  float fy = A.a[3] * xc + A.a[4] * yc + A.a[5] * zc + A.c[1];
  float fz = A.a[6] * xc + A.a[7] * yc + A.a[8] * zc + A.c[2];
  const float hi = (float)(S - 1);
  fx = fminf(fmaxf(fx, 0.f), hi); fy = fminf(fmaxf(fy, 0.f), hi); fz = fminf(fmaxf(fz, 0.f), hi);
  int ix = min((int)floorf(fx), S - 2), iy = min((int)floorf(fy), S - 2), iz = min((int)floorf(fz), S - 2);
  const float wx = fx - ix, wy = fy - iy, wz = fz - iz;
  float c[8];
  if (MODE == 0) {
    const float* p = src + ((size_t)ix * S + iy) * S + iz;
    c[0] = __ldg(p); c[1] = __ldg(p + 1); c[2] = __ldg(p + S); c[3] = __ldg(p + S + 1);
    p += (size_t)S * S;
    c[4] = __ldg(p); c[5] = __ldg(p + 1); c[6] = __ldg(p + S); c[7] = __ldg(p + S + 1);
  } else if (MODE == 1) {
    // cudaArray extent (width = z, height = y, depth = x), unnormalised, point sampling: texel centre at +0.5
    const float tz = iz + 0.5f, ty = iy + 0.5f, tx = ix + 0.5f;
    c[0] = tex3D<float>(tex, tz, ty, tx); c[1] = tex3D<float>(tex, tz + 1, ty, tx);
    c[2] = tex3D<float>(tex, tz, ty + 1, tx); c[3] = tex3D<float>(tex, tz + 1, ty + 1, tx);
    c[4] = tex3D<float>(tex, tz, ty, tx + 1); c[5] = tex3D<float>(tex, tz + 1, ty, tx + 1);
    c[6] = tex3D<float>(tex, tz, ty + 1, tx + 1); c[7] = tex3D<float>(tex, tz + 1, ty + 1, tx + 1);
  } else {
    const int p = (ix * S + iy) * S + iz;
    c[0] = tex1Dfetch<float>(tex, p); c[1] = tex1Dfetch<float>(tex, p + 1); c[2] = tex1Dfetch<float>(tex, p + S); c[3] = tex1Dfetch<float>(tex, p + S + 1);
    const int q = p + S * S;
    c[4] = tex1Dfetch<float>(tex, q); c[5] = tex1Dfetch<float>(tex, q + 1); c[6] = tex1Dfetch<float>(tex, q + S); c[7] = tex1Dfetch<float>(tex, q + S + 1);
  }
  const float v00 = c[0] * (1 - wz) + c[1] * wz, v01 = c[2] * (1 - wz) + c[3] * wz;
  const float v10 = c[4] * (1 - wz) + c[5] * wz, v11 = c[6] * (1 - wz) + c[7] * wz;
  const float v0 = v00 * (1 - wy) + v01 * wy, v1 = v10 * (1 - wy) + v11 * wy;
  dst[((size_t)x * S + y) * S + z] = v0 * (1 - wx) + v1 * wx;
}

// surface write of a linear volume into the array (what a producer kernel would do)
__global__ void fill_surf(cudaSurfaceObject_t surf, const float* __restrict__ src, int S) {
  const int z = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), x = blockIdx.z;
  if (z < S && y < S) surf3Dwrite(src[((size_t)x * S + y) * S + z], surf, z * 4, y, x);
}
__global__ void copy_lin(float* __restrict__ dst, const float* __restrict__ src, int S) {
  const int z = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), x = blockIdx.z;
  if (z < S && y < S) dst[((size_t)x * S + y) * S + z] = src[((size_t)x * S + y) * S + z];
}

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 256, B = argc > 2 ? atoi(argv[2]) : 8;
  const float deg = argc > 3 ? atof(argv[3]) : 12.f;
  const size_t N = (size_t)S * S * S;
  std::vector<float> h(N);
  for (size_t i = 0; i < N; ++i) h[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f;
  std::vector<float*> src(B), dst(B);
  std::vector<cudaArray_t> arr(B);
  std::vector<cudaTextureObject_t> t3(B), t1(B);
  std::vector<cudaSurfaceObject_t> sf(B);
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
  for (int b = 0; b < B; ++b) {
    CK(cudaMalloc(&src[b], N * 4)); CK(cudaMalloc(&dst[b], N * 4));
    CK(cudaMemcpy(src[b], h.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc3DArray(&arr[b], &cd, make_cudaExtent(S, S, S), cudaArraySurfaceLoadStore));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr[b];
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&t3[b], &rd, &td, nullptr));
    CK(cudaCreateSurfaceObject(&sf[b], &rd));
    cudaResourceDesc rl = {}; rl.resType = cudaResourceTypeLinear; rl.res.linear.devPtr = src[b]; rl.res.linear.desc = cd; rl.res.linear.sizeInBytes = N * 4;
    cudaTextureDesc tl = {}; tl.readMode = cudaReadModeElementType;
    CK(cudaCreateTextureObject(&t1[b], &rl, &tl, nullptr));
  }
  // rotation about all three axes by `deg`, scale 1.05
  const float r = deg * 3.14159265f / 180.f, cs = cosf(r), sn = sinf(r);
  float Rx[9] = {1, 0, 0, 0, cs, -sn, 0, sn, cs}, Ry[9] = {cs, 0, sn, 0, 1, 0, -sn, 0, cs}, Rz[9] = {cs, -sn, 0, sn, cs, 0, 0, 0, 1}, T[9], M[9];
  auto mm = [](const float* a, const float* b, float* o) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { float s = 0; for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j]; o[3 * i + j] = s; } };
  mm(Rx, Ry, T); mm(T, Rz, M);
  Aff A; for (int i = 0; i < 9; ++i) A.a[i] = M[i] * 1.05f; for (int i = 0; i < 3; ++i) { A.c[i] = (S - 1) / 2.f; A.ctr[i] = (S - 1) / 2.f; }
  dim3 gb(256), gg((S + 31) / 32, (S + 3) / 4, (S + 1) / 2);
  dim3 fb(256), fg((S + 31) / 32, (S + 7) / 8, S);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // producer cost: surface write vs linear copy
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0); for (int b = 0; b < B; ++b) fill_surf<<<fg, fb>>>(sf[b], src[b], S); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("surface write of %d volumes: %.3f ms\n", B, ms);
    cudaEventRecord(e0); for (int b = 0; b < B; ++b) copy_lin<<<fg, fb>>>(dst[b], src[b], S); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    if (rep == 2) printf("linear copy of %d volumes:   %.3f ms\n", B, ms);
  }
  std::vector<float> o0(N), o1(N);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      for (int b = 0; b < B; ++b) {
        if (mode == 0) gather_kernel<0><<<gg, gb>>>(src[b], 0, dst[b], S, A);
        if (mode == 1) gather_kernel<1><<<gg, gb>>>(src[b], t3[b], dst[b], S, A);
        if (mode == 2) gather_kernel<2><<<gg, gb>>>(src[b], t1[b], dst[b], S, A);
      }
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    CK(cudaGetLastError());
    printf("mode %d (%s): %.3f ms for %d volumes of %d^3 (rotation %.0f deg)\n", mode, mode == 0 ? "LDG linear" : mode == 1 ? "tex3D point, cudaArray" : "tex1Dfetch linear", ms, B, S, deg);
    CK(cudaMemcpy(mode == 0 ? o0.data() : o1.data(), dst[0], N * 4, cudaMemcpyDeviceToHost));
    if (mode > 0) { size_t bad = 0; for (size_t i = 0; i < N; ++i) bad += o0[i] != o1[i]; printf("   mismatches vs LDG: %zu\n", bad); }
  }
  return 0;
}
