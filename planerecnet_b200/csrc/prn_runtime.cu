// prn_runtime.cu — error strings, driver entry points, device queries for libprn_b200.
#include <stdarg.h>
#include <stdio.h>
#include "prn_internal.h"

namespace prn {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  // No link-time dependency on libcuda: resolve through the runtime.
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         int dtype, uint64_t pitch_elems) {
  if (pitch_elems == 0) pitch_elems = cols;
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(PRN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_elems * 2) % 16 != 0)
    return set_error(PRN_ERR_INVALID, "tensor map base/pitch must be 16-byte aligned");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype == PRN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PRN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PRN_OK;
}

int encode_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes, int dtype) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(PRN_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (rank < 2 || rank > 5) return set_error(PRN_ERR_INVALID, "tensor map rank must be 2..5");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(PRN_ERR_INVALID, "tensor map base must be 16-byte aligned");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] % 16 != 0) return set_error(PRN_ERR_INVALID, "tensor map strides must be multiples of 16 bytes");
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dtype == PRN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(PRN_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
  return PRN_OK;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  n = v;
  return n;
}

}  // namespace prn

extern "C" {
const char* prn_last_error(void) { return prn::g_err; }
int prn_abi_version(void) { return 1; }
#ifndef PRN_BUILD_FP
#define PRN_BUILD_FP "unknown"
#endif
static const char g_build_fp[] = "PRN_FP:" PRN_BUILD_FP;     // the marker lets build.py read it from the file's bytes
const char* prn_build_fingerprint(void) { return g_build_fp + 7; }
int prn_device_sm_count(void) { return prn::sm_count(); }
}
