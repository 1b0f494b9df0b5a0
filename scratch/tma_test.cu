// standalone check of the TMA box load used by warp_tile.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Params { CUtensorMap img[4]; CUtensorMap seg[4]; int box[4]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ Params tp, int jb, int c0, int c1, int c2, float* out, uint8_t* outs, int mode) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int ex = tp.box[0], ey = tp.box[1], ez = tp.box[2], ezs = tp.box[3];
  float* s_img = reinterpret_cast<float*>(s_raw);
  uint8_t* s_seg = s_raw + (size_t)((ex * ey * ez + 31) / 32 * 32) * 4;
  __shared__ unsigned long long mbar;
  const uint32_t bar = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    uint32_t bytes = ex * ey * ez * 4 + (mode >= 2 ? ex * ey * ezs : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(s_img)), "l"(&tp.img[jb]), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    if (mode >= 2)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(s_seg)), "l"(&tp.seg[jb]), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  __syncthreads();
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < ex * ey * ez; i += blockDim.x) out[i] = s_img[i];
  if (mode >= 2) for (int i = threadIdx.x; i < ex * ey * ezs; i += blockDim.x) outs[i] = s_seg[i];
}
int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 2;
  const int S = 32, ex = 28, ey = 28, ez = 28, ezs = 32;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
  EncodeTiledFn fn = (EncodeTiledFn)p;
  std::vector<float> h(S * S * S); std::vector<uint8_t> hs(S * S * S);
  for (int i = 0; i < S * S * S; ++i) { h[i] = (float)i; hs[i] = (uint8_t)(i * 7); }
  float *d, *out; uint8_t *ds, *outs;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&ds, hs.size()); cudaMalloc(&out, ex * ey * ez * 4); cudaMalloc(&outs, ex * ey * ezs);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(ds, hs.data(), hs.size(), cudaMemcpyHostToDevice);
  static Params tp; memset(&tp, 0, sizeof(tp));
  tp.box[0] = ex; tp.box[1] = ey; tp.box[2] = ez; tp.box[3] = ezs;
  int jb = 1;
  {
    cuuint64_t dims[3] = {S, S, S}; cuuint64_t st[2] = {S * 4, S * S * 4}; cuuint32_t box[3] = {ez, ey, ex}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tp.img[jb], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode img rc=%d\n", (int)r);
    cuuint64_t st2[2] = {S, S * S}; cuuint32_t box2[3] = {ezs, ey, ex};
    r = fn(&tp.seg[jb], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ds, dims, st2, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode seg rc=%d\n", (int)r);
  }
  size_t smem = (size_t)((ex * ey * ez + 31) / 32 * 32) * 4 + ex * ey * ezs + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int c0 = argc > 2 ? atoi(argv[2]) : 3, c1 = 2, c2 = -1;
  k<<<1, 256, smem>>>(tp, jb, c0, c1, c2, out, outs, mode);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<float> ho(ex * ey * ez); std::vector<uint8_t> hso(ex * ey * ezs);
  cudaMemcpy(ho.data(), out, ho.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hso.data(), outs, hso.size(), cudaMemcpyDeviceToHost);
  int bad = 0, bads = 0;
  for (int x = 0; x < ex; ++x) for (int y = 0; y < ey; ++y) for (int z = 0; z < ezs; ++z) {
    int gx = c2 + x, gy = c1 + y, gz = c0 + z;
    bool in = gx >= 0 && gx < S && gy >= 0 && gy < S && gz >= 0 && gz < S;
    if (z < ez) { float want = in ? h[(gx * S + gy) * S + gz] : 0.f; if (ho[(x * ey + y) * ez + z] != want) ++bad; }
    if (mode >= 2) { uint8_t want = in ? hs[(gx * S + gy) * S + gz] : 0; if (hso[(x * ey + y) * ezs + z] != want) ++bads; }
  }
  printf("mismatches img=%d seg=%d\n", bad, bads);
  return 0;
}
