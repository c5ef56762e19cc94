// Host-side plumbing shared by every entry point: per-thread error string, device queries and
// TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda).
#include "common.cuh"

#include <cstring>
#include <mutex>

namespace isx {

namespace {
thread_local char g_last_error[512] = "";
}

char* last_error_buffer() { return g_last_error; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int device_sm_count(int* out) {
  int dev = 0;
  ISX_CHECK_CUDA(cudaGetDevice(&dev));
  ISX_CHECK_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return ISX_OK;
}

namespace {
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* fn) {
  static EncodeTiledFn cached = nullptr;
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (err == cudaSuccess && qres == cudaDriverEntryPointSuccess) cached = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (cached == nullptr) {
    return set_error(ISX_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (%s)",
                     cudaGetErrorString(err));
  }
  *fn = cached;
  return ISX_OK;
}
}  // namespace

int encode_tmap_2d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base,
                   uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes, uint32_t box_rows,
                   uint32_t box_cols, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc != ISX_OK) return rc;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = fn(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(ISX_ERR_CUDA,
                     "cuTensorMapEncodeTiled(2d) failed with CUresult %d (rows=%llu cols=%llu pitch=%llu "
                     "box=%ux%u)",
                     (int)r, (unsigned long long)rows, (unsigned long long)cols,
                     (unsigned long long)row_pitch_bytes, box_rows, box_cols);
  }
  return ISX_OK;
}

int encode_tmap_3d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base,
                   uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                   uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                   CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc != ISX_OK) return rc;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  (void)elem_bytes;
  CUresult r = fn(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(ISX_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  }
  return ISX_OK;
}

}  // namespace isx

extern "C" {

int isx_abi_version(void) { return ISX_ABI_VERSION; }

const char* isx_last_error(void) { return isx::last_error_buffer(); }

int isx_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  ISX_REQUIRE(sm_count && cc_major && cc_minor, "isx_device_info: null out-param");
  int dev = 0;
  ISX_CHECK_CUDA(cudaGetDevice(&dev));
  ISX_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  ISX_CHECK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  ISX_CHECK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return ISX_OK;
}

}  // extern "C"
