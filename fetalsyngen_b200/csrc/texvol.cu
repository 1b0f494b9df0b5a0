// Block-linear intensity volumes (fsg_texvol, see fsg.h): a layered 2-D CUDA array per sample that fsg_gmm
// writes through a surface and fsg_warp's fast path reads with 2x2 texture gathers.  Allocation and object
// creation happen once per engine; nothing here is on the per-step path.
#include "common.cuh"

using namespace fsg;

#define FSG_CUDA(call, what)                                                        \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      fsg::set_error("%s: %s", what, cudaGetErrorString(e_));                       \
      cudaGetLastError();                                                           \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

extern "C" int fsg_texvol_create(int nx, int ny, int nz, fsg_texvol* out) {
  FSG_REQUIRE(out != nullptr, "fsg_texvol_create: NULL output");
  FSG_REQUIRE(nx >= 1 && ny >= 1 && nz >= 4 && nz % 4 == 0 && nx <= 2048 && ny <= 32768 && nz <= 32768, "fsg_texvol_create: bad shape %dx%dx%d (nz must be a multiple of 4)", nx, ny, nz);
  *out = fsg_texvol{};
  const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
  cudaArray_t arr = nullptr;
  // width = z (fastest), height = y, layers = x
  FSG_CUDA(cudaMalloc3DArray(&arr, &desc, make_cudaExtent((size_t)nz, (size_t)ny, (size_t)nx), cudaArrayLayered | cudaArraySurfaceLoadStore), "fsg_texvol_create: cudaMalloc3DArray");
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex = 0;
  cudaSurfaceObject_t surf = 0;
  cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  if (e == cudaSuccess) e = cudaCreateSurfaceObject(&surf, &rd);
  if (e != cudaSuccess) {
    if (tex) cudaDestroyTextureObject(tex);
    cudaFreeArray(arr);
    set_error("fsg_texvol_create: texture/surface object: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 2;
  }
  out->array = reinterpret_cast<uint64_t>(arr);
  out->tex = (uint64_t)tex;
  out->surf = (uint64_t)surf;
  out->nx = nx;
  out->ny = ny;
  out->nz = nz;
  return 0;
}

extern "C" int fsg_texvol_destroy(fsg_texvol* v) {
  FSG_REQUIRE(v != nullptr, "fsg_texvol_destroy: NULL volume");
  if (v->tex) cudaDestroyTextureObject((cudaTextureObject_t)v->tex);
  if (v->surf) cudaDestroySurfaceObject((cudaSurfaceObject_t)v->surf);
  if (v->array) cudaFreeArray(reinterpret_cast<cudaArray_t>(v->array));
  *v = fsg_texvol{};
  cudaGetLastError();
  return 0;
}

extern "C" int fsg_texvol_copy(const fsg_texvol* v, float* linear_dev, int to_linear, void* stream) {
  FSG_REQUIRE(v != nullptr && v->array && linear_dev != nullptr, "fsg_texvol_copy: NULL argument");
  cudaMemcpy3DParms p = {};
  const cudaPitchedPtr lin = make_cudaPitchedPtr(linear_dev, (size_t)v->nz * sizeof(float), (size_t)v->nz, (size_t)v->ny);
  if (to_linear) {
    p.srcArray = reinterpret_cast<cudaArray_t>(v->array);
    p.dstPtr = lin;
  } else {
    p.srcPtr = lin;
    p.dstArray = reinterpret_cast<cudaArray_t>(v->array);
  }
  p.extent = make_cudaExtent((size_t)v->nz, (size_t)v->ny, (size_t)v->nx);
  p.kind = cudaMemcpyDeviceToDevice;
  FSG_CUDA(cudaMemcpy3DAsync(&p, as_stream(stream)), "fsg_texvol_copy");
  return 0;
}
