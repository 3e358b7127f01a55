// prn_pw.cuh — helpers shared by the HBM-bound pointwise / reduction passes (prn_pointwise.cu, prn_train.cu):
// 16-byte (8 channel) loads and stores of 16-bit NHWC rows, grid sizing, dtype dispatch.
#pragma once
#include "prn_internal.h"
#include "prn_ptx.cuh"

namespace prn {

constexpr int kPwThreads = 256;

static inline int pw_grid(long long work_items) {
  long long b = (work_items + kPwThreads - 1) / kPwThreads;
  const long long cap = static_cast<long long>(sm_count()) * 16;  // grid-stride beyond 16 CTAs/SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 v = Pack2<T>::unpack(w[j]);
    f[2 * j] = v.x;
    f[2 * j + 1] = v.y;
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float* f) {
  uint4 o;
  o.x = Pack2<T>::pack(f[0], f[1]);
  o.y = Pack2<T>::pack(f[2], f[3]);
  o.z = Pack2<T>::pack(f[4], f[5]);
  o.w = Pack2<T>::pack(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

}  // namespace prn

#define PRN_DISPATCH(dtype, KERNEL_CALL_BF16, KERNEL_CALL_F16) \
  do {                                                         \
    if ((dtype) == PRN_BF16) { KERNEL_CALL_BF16; }             \
    else if ((dtype) == PRN_F16) { KERNEL_CALL_F16; }          \
    else return set_error(PRN_ERR_INVALID, "bad dtype %d", (int)(dtype)); \
  } while (0)

#define PRN_LAUNCH_CHECK()                                                                        \
  do {                                                                                            \
    cudaError_t _e = cudaGetLastError();                                                          \
    if (_e != cudaSuccess) return set_error(PRN_ERR_CUDA, "%s launch: %s", __func__, cudaGetErrorString(_e)); \
    return PRN_OK;                                                                                \
  } while (0)

