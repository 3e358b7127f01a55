// prn_pw.cuh — helpers shared by the HBM-bound pointwise / reduction passes (prn_pointwise.cu, prn_train.cu):
// 16-byte (8 channel) loads and stores of 16-bit NHWC rows, grid sizing, dtype dispatch.
#pragma once
#include "prn_internal.h"
#include "prn_ptx.cuh"

namespace prn {

constexpr int kPwThreads = 256;

static inline int pw_grid(long long work_items) {
  // the passes decode their element index with 32-bit arithmetic: refuse anything larger (grid 0 = launch error, reported by
  // PRN_LAUNCH_CHECK); the path's largest tensor has 2e7 16-byte groups
  if (work_items >= (1LL << 31)) {
    set_error(PRN_ERR_UNSUPPORTED, "pointwise pass over %lld work items (>= 2^31) is not supported", work_items);
    return 0;
  }
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
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
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

// Block-level fold of per-thread channel partials for the channel-stationary reductions (BatchNorm / GroupNorm backward):
// thread t of a kPwThreads block holds s1[8], s2[8] for the 8 channels of group t % cv.  Plain shared-memory stores into a
// padded [slot][16][cv + 1] array + one summing pass — no shared-memory atomics: fp32 atomicAdd on shared memory is a
// compare-and-swap loop, and in the natural acc[2 * channel] layout the 32 lanes of a warp hit 2 banks (16-way conflict) with
// all 8 warps contending for the same words; tools/probe/reduce_probe.cu measured that tail at ~half of the whole pass.
// Requires cv a power of two <= kPwThreads (so that t % cv is the thread's channel group).  emit(o, total) is called once for
// every output o = 2 * channel + {0: s1, 1: s2}, o in [0, 16 cv), by exactly one thread of the block.
constexpr int kFoldFloats = (kPwThreads / 32) * 16 * 33;   // largest case: cv == 32

__device__ __forceinline__ bool fold_ok(int cv) { return (cv & (cv - 1)) == 0 && cv <= kPwThreads; }

template <typename Emit>
__device__ __forceinline__ void block_fold_chan(float* s1, float* s2, int cv, float* part, Emit emit) {
  const int t = threadIdx.x, lane = t & 31;
  bool writer = true;
  int slot, nslots, cg;
  if (cv < 32) {        // lanes l, l + cv, ... of a warp hold the same channels: fold them with shuffles first
    for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
      }
    }
    writer = lane < cv;
    slot = t >> 5;
    nslots = kPwThreads / 32;
    cg = lane;
  } else {
    slot = t / cv;
    nslots = kPwThreads / cv;
    cg = t % cv;
  }
  const int stride = cv + 1;
  if (writer) {
    float* p = part + slot * 16 * stride + cg;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p[j * stride] = s1[j];
      p[(8 + j) * stride] = s2[j];
    }
  }
  __syncthreads();
  for (int o = t; o < 16 * cv; o += kPwThreads) {
    const int k = (o & 1) * 8 + ((o & 15) >> 1);
    const float* p = part + k * stride + (o >> 4);
    float tot = 0.f;
    for (int s = 0; s < nslots; ++s) tot += p[s * 16 * stride];
    emit(o, tot);
  }
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

