// prn_train.cu — HBM-bound passes of the training step (train-mode BatchNorm forward, and the backward of
// BatchNorm / ReLU / max-pool / strided 1x1 / modulated deformable sampling).  NHWC, 16-bit, 8 channels
// (16 bytes) per thread; per-channel reductions accumulate in shared memory and leave with one fp32 atomic per
// (CTA, channel).  The reference has no hand-written backward: these are autograd's formulas for the operator
// call sites cited per function in include/prn_b200.h.
#include "prn_pw.cuh"

namespace prn {

// ---------------------------------------------------------------- BatchNorm2d, training mode
// stats[c*2 + {0,1}] = {sum, sumsq} over `count` rows (accumulated by the conv epilogue, fp32)
__global__ void bn_finalize_kernel(const float* __restrict__ stats, float* __restrict__ mean_invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int C, float count,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = stats[2 * c] / count;
  const float var = fmaxf(stats[2 * c + 1] / count - mean * mean, 0.f);   // biased: normalisation
  mean_invstd[2 * c] = mean;
  mean_invstd[2 * c + 1] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}


// Channel-stationary threads: a thread keeps its 8 channels (scale/shift in registers) and walks rows with a stride of
// (total threads / (C/8)); total thread count is a multiple of C/8 (chan_grid).  One 16-byte load per operand and row.
template <typename T>
__global__ void __launch_bounds__(kPwThreads) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                            const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const T* __restrict__ res,
                                                            long long rows, int C, int relu) {
  const int cv = C / 8;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rstep = (gridDim.x * blockDim.x) / static_cast<unsigned>(cv);      // 32-bit: grids stay far below 2^31 threads
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(mean_invstd + 2 * (c + j) + 1) * __ldg(gamma + c + j);
    sh[j] = __ldg(beta + c + j) - __ldg(mean_invstd + 2 * (c + j)) * sc[j];
  }
  const float lo = relu ? 0.f : -INFINITY;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  // two rows per iteration, all loads issued before the first use (latency-bound otherwise)
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < rows; m += 2 * rstep) {
    const long long m2 = m + rstep;
    const bool two = m2 < rows;
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(x + m * C + c));
    const uint4 x1 = two ? __ldg(reinterpret_cast<const uint4*>(x + m2 * C + c)) : zero4;
    uint4 r0 = zero4, r1 = zero4;
    if (res != nullptr) {
      r0 = __ldg(reinterpret_cast<const uint4*>(res + m * C + c));
      if (two) r1 = __ldg(reinterpret_cast<const uint4*>(res + m2 * C + c));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float f[8], r[8];
      unpack8<T>(h ? x1 : x0, f);
      unpack8<T>(h ? r1 : r0, r);      // zeros when there is no residual
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]) + r[j], lo);
      store8(out + (h ? m2 : m) * C + c, f);
    }
  }
}

// bn_finalize + bn_apply in one launch (training forward: 113 BatchNorms per step): every thread derives mean / invstd of its
// 8 channels from the batch sums; the first C/8 threads of the grid also publish mean_invstd (the backward reads it) and update
// the running statistics (nn.BatchNorm2d: momentum, unbiased variance).
template <typename T>
__global__ void __launch_bounds__(kPwThreads) bn_finalize_apply_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                     const float* __restrict__ stats, float* __restrict__ mean_invstd,
                                                                     float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                     float count, float eps, float momentum,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const T* __restrict__ res, long long rows, int C, int relu) {
  const int cv = C / 8;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rstep = (gridDim.x * blockDim.x) / static_cast<unsigned>(cv);      // 32-bit: grids stay far below 2^31 threads
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  const bool writer = gtid < cv;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = __ldg(stats + 2 * (c + j)) / count;
    const float var = fmaxf(__ldg(stats + 2 * (c + j) + 1) / count - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    if (writer) {
      mean_invstd[2 * (c + j)] = mean;
      mean_invstd[2 * (c + j) + 1] = invstd;
      if (running_mean != nullptr) {
        const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
        running_mean[c + j] = (1.f - momentum) * running_mean[c + j] + momentum * mean;
        running_var[c + j] = (1.f - momentum) * running_var[c + j] + momentum * unbiased;
      }
    }
    sc[j] = invstd * __ldg(gamma + c + j);
    sh[j] = __ldg(beta + c + j) - mean * sc[j];
  }
  const float lo = relu ? 0.f : -INFINITY;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < rows; m += 2 * rstep) {
    const long long m2 = m + rstep;
    const bool two = m2 < rows;
    const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(x + m * C + c));
    const uint4 x1 = two ? __ldg(reinterpret_cast<const uint4*>(x + m2 * C + c)) : zero4;
    uint4 r0 = zero4, r1 = zero4;
    if (res != nullptr) {
      r0 = __ldg(reinterpret_cast<const uint4*>(res + m * C + c));
      if (two) r1 = __ldg(reinterpret_cast<const uint4*>(res + m2 * C + c));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      float f[8], r[8];
      unpack8<T>(h ? x1 : x0, f);
      unpack8<T>(h ? r1 : r0, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]) + r[j], lo);
      store8(out + (h ? m2 : m) * C + c, f);
    }
  }
}

// sums[c*2] += sum_m g, sums[c*2+1] += sum_m g * xhat with g = dz * (out > 0) (out == NULL: g = dz) and
// xhat = (x - mean) * invstd (x == NULL: second sum skipped).  Total thread count is a multiple of C/8 so that a
// thread keeps its 8 channels for all of its rows.
template <typename T>
__global__ void __launch_bounds__(kPwThreads) chan_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ out,
                                                               const T* __restrict__ x, const float* __restrict__ mean_invstd,
                                                               float* __restrict__ sums, long long rows, int C) {
  const int cv = C / 8;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long tthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  const long long rstep = static_cast<unsigned>(tthreads) / static_cast<unsigned>(cv);
  // s2 accumulates sum g * x; the normalisation is applied once at the end: sum g*xhat = inv * (s2 - mean * s1)
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  // two rows per iteration, all loads issued before the first use (the pass is latency-bound otherwise)
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < rows; m += 2 * rstep) {
    const long long m2 = m + rstep;
    const bool two = m2 < rows;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(dz + m * C + c));
    const uint4 a1 = two ? __ldg(reinterpret_cast<const uint4*>(dz + m2 * C + c)) : zero4;
    uint4 b0 = zero4, b1 = zero4, c0 = zero4, c1 = zero4;
    if (out != nullptr) {
      b0 = __ldg(reinterpret_cast<const uint4*>(out + m * C + c));
      if (two) b1 = __ldg(reinterpret_cast<const uint4*>(out + m2 * C + c));
    }
    if (x != nullptr) {
      c0 = __ldg(reinterpret_cast<const uint4*>(x + m * C + c));
      if (two) c1 = __ldg(reinterpret_cast<const uint4*>(x + m2 * C + c));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float g[8], o[8], xv[8];
      unpack8<T>(r ? a1 : a0, g);
      if (out != nullptr) {
        unpack8<T>(r ? b1 : b0, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
      }
      unpack8<T>(r ? c1 : c0, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += g[j];
        s2[j] = fmaf(g[j], xv[j], s2[j]);
      }
    }
  }
  if (x != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mean = __ldg(mean_invstd + 2 * (c + j)), inv = __ldg(mean_invstd + 2 * (c + j) + 1);
      s2[j] = inv * (s2[j] - mean * s1[j]);
    }
  }
  if (fold_ok(cv)) {     // every BatchNorm of the ResNets: C in {64 ... 2048}
    __shared__ float part[kFoldFloats];
    block_fold_chan(s1, s2, cv, part, [&](int o, float tot) { atomicAdd(sums + o, tot); });
    return;
  }
  // general channel counts: shared-memory atomics
  extern __shared__ float acc[];   // [C][2]
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(acc + 2 * (c + j), s1[j]);
    atomicAdd(acc + 2 * (c + j) + 1, s2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(sums + i, acc[i]);
}

// dx = gamma * invstd * (g - sum_g / count - xhat * sum_gxhat / count); optionally g itself is written too (the
// gradient of the residual branch of a bottleneck: models/backbone.py:69-70)
template <typename T>
__global__ void __launch_bounds__(kPwThreads) bn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ out,
                                                                const T* __restrict__ x, const float* __restrict__ mean_invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ sums,
                                                                T* __restrict__ dx, T* __restrict__ g_out, long long rows, int C,
                                                                float inv_count) {
  const int cv = C / 8;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rstep = (gridDim.x * blockDim.x) / static_cast<unsigned>(cv);      // 32-bit: grids stay far below 2^31 threads
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  // dx = a * g + b * x + k  with  a = gamma*invstd,  b = -a*invstd*sgx,  k = -a*sg - b*mean
  float ka[8], kb[8], kk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float mean = __ldg(mean_invstd + 2 * (c + j)), inv = __ldg(mean_invstd + 2 * (c + j) + 1);
    const float sg = __ldg(sums + 2 * (c + j)) * inv_count, sgx = __ldg(sums + 2 * (c + j) + 1) * inv_count;
    ka[j] = __ldg(gamma + c + j) * inv;
    kb[j] = -ka[j] * inv * sgx;
    kk[j] = -ka[j] * sg - kb[j] * mean;
  }
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  // two rows per iteration, all loads issued before the first use (latency-bound otherwise: most launches move < 30 MB)
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < rows; m += 2 * rstep) {
    const long long m2 = m + rstep;
    const bool two = m2 < rows;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(dz + m * C + c));
    const uint4 a1 = two ? __ldg(reinterpret_cast<const uint4*>(dz + m2 * C + c)) : zero4;
    const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(x + m * C + c));
    const uint4 c1 = two ? __ldg(reinterpret_cast<const uint4*>(x + m2 * C + c)) : zero4;
    uint4 b0 = zero4, b1 = zero4;
    if (out != nullptr) {
      b0 = __ldg(reinterpret_cast<const uint4*>(out + m * C + c));
      if (two) b1 = __ldg(reinterpret_cast<const uint4*>(out + m2 * C + c));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long mm = h ? m2 : m;
      float g[8], o[8], xv[8];
      unpack8<T>(h ? a1 : a0, g);
      unpack8<T>(h ? c1 : c0, xv);
      if (out != nullptr) {
        unpack8<T>(h ? b1 : b0, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
      }
      if (g_out != nullptr) store8(g_out + mm * C + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = fmaf(ka[j], g[j], fmaf(kb[j], xv[j], kk[j]));
      store8(dx + mm * C + c, xv);
    }
  }
}

// g = dz * (out > 0): ReLU backward on its own (layers whose normalisation is not a BatchNorm)
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dz, const T* __restrict__ out, T* __restrict__ g, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a[8], o[8];
    load8(dz + i * 8, a);
    load8(out + i * 8, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = o[j] > 0.f ? a[j] : 0.f;
    store8(g + i * 8, a);
  }
}

// dst[b, s*i, s*j, :] += src[b, i, j, :]: input gradient of a 1x1 stride-s convolution (the bottleneck's downsample
// branch, models/backbone.py:152-167) joined with the gradient already in dst
template <typename T>
__global__ void add_strided_kernel(T* __restrict__ dst, const T* __restrict__ src, int B, int h, int w, int C, int s) {
  const int cv = C / 8;
  const long long total = static_cast<long long>(B) * h * w * cv;
  const int H = h * s, W = w * s;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(w);
    const int x = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(w)), b = static_cast<int>(tu_ / static_cast<unsigned>(h));
    const int y = static_cast<int>(tu_) - b * h;
    T* dp = dst + ((static_cast<long long>(b) * H + y * s) * W + x * s) * C + c;
    float a[8], q[8];
    load8(src + m * C + c, a);
    load8(dp, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += q[j];
    store8(dp, a);
  }
}

// out16 = a32 (+ b16): joins an fp32 scatter buffer with the 16-bit gradient stream
template <typename T>
__global__ void add_f32_kernel(const float* __restrict__ a32, const T* __restrict__ b16, T* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 u = __ldg(reinterpret_cast<const float4*>(a32) + 2 * i), v = __ldg(reinterpret_cast<const float4*>(a32) + 2 * i + 1);
    float f[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    if (b16 != nullptr) {
      float q[8];
      load8(b16 + i * 8, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += q[j];
    }
    store8(out + i * 8, f);
  }
}

// out16 = a16 + b16
template <typename T>
__global__ void add16_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], q[8];
    load8(a + i * 8, f);
    load8(b + i * 8, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += q[j];
    store8(out + i * 8, f);
  }
}

// ---------------------------------------------------------------- MaxPool2d(3, 2, 1) backward (models/backbone.py:104)
// Gather form, deterministic: an input pixel receives the gradient of every window in which it is the FIRST maximum
// in (dy, dx) scan order (torch's strict `>` update rule), so ties resolve like the reference.
// One thread owns a 2x2 block of input pixels (8 channels): the four windows that can select them — (k, l), (k, l+1), (k+1, l),
// (k+1, l+1) — cover a 5x5 input patch, scanned once in row-major order (= every window's own scan order): 25 loads for four
// input pixels instead of 2.25 windows x 9 loads for each one (the per-pixel form ran at 1/13 of the HBM rate).
template <typename T>
__global__ void __launch_bounds__(kPwThreads) maxpool3s2_bwd_kernel(const T* __restrict__ in, const T* __restrict__ dout,
                                                                  T* __restrict__ din, int B, int H, int W, int C) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, cv = C / 8;
  const int Hb = (H + 1) / 2, Wb = (W + 1) / 2;
  const long long total = static_cast<long long>(B) * Hb * Wb * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(Wb);
    const int l = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(Wb)), b = static_cast<int>(tu_ / static_cast<unsigned>(Hb));
    const int k = static_cast<int>(tu_) - b * Hb;
    const T* img = in + static_cast<long long>(b) * H * W * C + c;
    float best[2][2][8];
    int arg[2][2][8];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
      for (int wx = 0; wx < 2; ++wx)
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[wy][wx][j] = -INFINITY; arg[wy][wx][j] = -1; }
#pragma unroll
    for (int py = -1; py <= 3; ++py) {
      const int yy = 2 * k + py;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int px = -1; px <= 3; ++px) {
        const int xx = 2 * l + px;
        if (xx < 0 || xx >= W) continue;
        float f[8];
        load8(img + (static_cast<long long>(yy) * W + xx) * C, f);
#pragma unroll
        for (int wy = 0; wy < 2; ++wy) {
          const int dy = py - (2 * wy - 1);
          if (dy < 0 || dy > 2) continue;
#pragma unroll
          for (int wx = 0; wx < 2; ++wx) {
            const int dx = px - (2 * wx - 1);
            if (dx < 0 || dx > 2) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (f[j] > best[wy][wx][j]) { best[wy][wx][j] = f[j]; arg[wy][wx][j] = dy * 3 + dx; }
          }
        }
      }
    }
    float g[2][2][8];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
      for (int wx = 0; wx < 2; ++wx) {
        const int ho = k + wy, wo = l + wx;
        if (ho < Ho && wo < Wo) load8(dout + ((static_cast<long long>(b) * Ho + ho) * Wo + wo) * C + c, g[wy][wx]);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) { g[wy][wx][j] = 0.f; arg[wy][wx][j] = -1; }
        }
      }
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      const int y = 2 * k + iy;
      if (y >= H) continue;
#pragma unroll
      for (int ix = 0; ix < 2; ++ix) {
        const int x = 2 * l + ix;
        if (x >= W) continue;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int wy = 0; wy <= iy; ++wy)            // even rows / columns lie in one window only, odd ones in two
#pragma unroll
          for (int wx = 0; wx <= ix; ++wx) {
            const int my = (iy - (2 * wy - 1)) * 3 + (ix - (2 * wx - 1));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += arg[wy][wx][j] == my ? g[wy][wx][j] : 0.f;
          }
        store8(din + ((static_cast<long long>(b) * H + y) * W + x) * C + c, acc);
      }
    }
  }
}

// ---------------------------------------------------------------- modulated deformable sampling, training form
// torchvision.ops.deform_conv2d (models/dcn.py:59-66) split into its two halves: the sampled + modulated im2col
// matrix col[m, tap*C + c] (16-bit, saved for the backward) and a plain contraction with the packed weights.
struct DcnGeom {
  int B, H, W, C, Ho, Wo, stride, pad;
};

__device__ __forceinline__ void dcn_corners(const DcnGeom& g, const float* om, int tap, int ho, int wo, float& mk, bool& inside,
                                            float& ly, float& lx, int& y0, int& x0) {
  const int ky = tap / 3, kx = tap - ky * 3;
  const float py = static_cast<float>(ho * g.stride - g.pad + ky) + __ldg(om + 2 * tap);
  const float px = static_cast<float>(wo * g.stride - g.pad + kx) + __ldg(om + 2 * tap + 1);
  mk = __ldg(om + 18 + tap);
  inside = py > -1.f && py < static_cast<float>(g.H) && px > -1.f && px < static_cast<float>(g.W);
  const float fy = floorf(py), fx = floorf(px);
  y0 = static_cast<int>(fy);
  x0 = static_cast<int>(fx);
  ly = py - fy;
  lx = px - fx;
}

template <typename T>
__global__ void dcn_im2col_kernel(const T* __restrict__ x, const float* __restrict__ offmask, T* __restrict__ col, DcnGeom g) {
  const int cv = g.C / 8;
  const long long total = static_cast<long long>(g.B) * g.Ho * g.Wo * 9 * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): six 64-bit divisions per 16 bytes made this pass compute-bound
    const unsigned iu_ = static_cast<unsigned>(i), t1_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - t1_ * static_cast<unsigned>(cv)) * 8;
    const unsigned mu_ = t1_ / 9u;
    const int tap = static_cast<int>(t1_ - mu_ * 9u);
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(g.Wo);
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(g.Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(g.Ho));
    const int ho = static_cast<int>(tu_) - b * g.Ho;
    float mk, ly, lx;
    bool inside;
    int y0, x0;
    dcn_corners(g, offmask + m * 32, tap, ho, wo, mk, inside, ly, lx, y0, x0);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (inside) {
#pragma unroll
      for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
        for (int cx = 0; cx < 2; ++cx) {
          const int yy = y0 + cy, xx = x0 + cx;
          if (yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
          const float wgt = (cy ? ly : 1.f - ly) * (cx ? lx : 1.f - lx) * mk;
          float f[8];
          load8(x + ((static_cast<long long>(b) * g.H + yy) * g.W + xx) * g.C + c, f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, f[j], acc[j]);
        }
      }
    }
    store8(col + m * (9LL * g.C) + static_cast<long long>(tap) * g.C + c, acc);
  }
}

// One warp per output pixel.  dcol = gradient of the im2col matrix.  Produces
//   dx32 (fp32, zero-initialised by the caller) += mask * bilinear weights * dcol      scattered with vector reductions
//   dpre16[m, 0:64]: gradient w.r.t. the PRE-activation output of the fused offset/modulator conv: offsets pass
//   where the clamp is inactive (|off| < bound), modulators through d(2*sigmoid)/dz = m * (1 - m/2); columns >= 27 are 0.
template <typename T>
__global__ void dcn_col2im_bwd_kernel(const T* __restrict__ x, const float* __restrict__ offmask, const T* __restrict__ dcol,
                                      float* __restrict__ dx32, T* __restrict__ dpre16, DcnGeom g, float bound) {
  const int cv = g.C / 8;
  const int lane = threadIdx.x & 31;
  const long long warp_g = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const long long M = static_cast<long long>(g.B) * g.Ho * g.Wo;
  for (long long m = warp_g; m < M; m += nwarps) {
    const unsigned mu_ = static_cast<unsigned>(m), tu_ = mu_ / static_cast<unsigned>(g.Wo);      // m < 2^24 pixels
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(g.Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(g.Ho));
    const int ho = static_cast<int>(tu_) - b * g.Ho;
    const float* om = offmask + m * 32;
    float v0 = 0.f, v1 = 0.f;   // the two dpre columns (2*lane, 2*lane+1) this lane writes
    for (int tap = 0; tap < 9; ++tap) {
      float mk, ly, lx;
      bool inside;
      int y0, x0;
      dcn_corners(g, om, tap, ho, wo, mk, inside, ly, lx, y0, x0);
      float s_m = 0.f, s_y = 0.f, s_x = 0.f;
      if (inside) {
        for (int ch = lane; ch < cv; ch += 32) {
          const int c = ch * 8;
          float gcol[8];
          load8(dcol + m * (9LL * g.C) + static_cast<long long>(tap) * g.C + c, gcol);
#pragma unroll
          for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
            for (int cx = 0; cx < 2; ++cx) {
              const int yy = y0 + cy, xx = x0 + cx;
              if (yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
              const float wy = cy ? ly : 1.f - ly, wxx = cx ? lx : 1.f - lx;
              const long long pix = ((static_cast<long long>(b) * g.H + yy) * g.W + xx) * g.C + c;
              float f[8];
              load8(x + pix, f);
              float dot = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) dot = fmaf(gcol[j], f[j], dot);
              s_m = fmaf(wy * wxx, dot, s_m);
              s_y = fmaf((cy ? 1.f : -1.f) * wxx, dot, s_y);
              s_x = fmaf((cx ? 1.f : -1.f) * wy, dot, s_x);
              const float wgt = wy * wxx * mk;
              red_add_v4_f32(dx32 + pix, wgt * gcol[0], wgt * gcol[1], wgt * gcol[2], wgt * gcol[3]);
              red_add_v4_f32(dx32 + pix + 4, wgt * gcol[4], wgt * gcol[5], wgt * gcol[6], wgt * gcol[7]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_m += __shfl_xor_sync(0xffffffffu, s_m, o);
        s_y += __shfl_xor_sync(0xffffffffu, s_y, o);
        s_x += __shfl_xor_sync(0xffffffffu, s_x, o);
      }
      const float offy = __ldg(om + 2 * tap), offx = __ldg(om + 2 * tap + 1);
      const float d_dy = (offy > -bound && offy < bound) ? s_y * mk : 0.f;
      const float d_dx = (offx > -bound && offx < bound) ? s_x * mk : 0.f;
      const float d_mk = s_m * mk * (1.f - 0.5f * mk);
      if (lane == tap) { v0 = d_dy; v1 = d_dx; }
      if (lane == 9 + (tap >> 1)) { if (tap & 1) v1 = d_mk; else v0 = d_mk; }
    }
    reinterpret_cast<uint32_t*>(dpre16 + m * 64)[lane] = Pack2<T>::pack(v0, v1);
  }
}

}  // namespace prn

// =================================================================================================== C ABI
using namespace prn;

static int chan_reduce_grid(long long rows, int cv, int rows_per_thread = 8, int ctas_per_sm = 8) {
  // total threads must be a multiple of cv: grid multiple of cv / gcd(cv, 256)
  int a = cv, b = kPwThreads;
  while (b) { const int t = a % b; a = b; b = t; }
  const int g0 = cv / a;
  long long want = (rows * cv + kPwThreads * static_cast<long long>(rows_per_thread) - 1) / (kPwThreads * static_cast<long long>(rows_per_thread));
  const long long cap = static_cast<long long>(sm_count()) * ctas_per_sm;
  if (want > cap) want = cap;
  long long grid = want / g0 * g0;
  if (grid < g0) grid = g0;
  return static_cast<int>(grid);
}

extern "C" {

int prn_bn_finalize(const float* stats, float* mean_invstd, float* running_mean, float* running_var, int32_t c, int64_t count,
                    float eps, float momentum, void* stream) {
  PRN_REQUIRE(stats && mean_invstd && c > 0 && count > 0 && (running_mean == nullptr) == (running_var == nullptr),
              "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(stats, mean_invstd, running_mean, running_var, c,
                                                                                   static_cast<float>(count), eps, momentum);
  PRN_LAUNCH_CHECK();
}

int prn_bn_apply(const void* x16, void* out16, const float* mean_invstd, const float* gamma, const float* beta,
                 const void* residual16, int64_t rows, int32_t c, int32_t relu, int32_t dtype, void* stream) {
  PRN_REQUIRE(x16 && out16 && mean_invstd && gamma && beta && rows > 0 && c > 0 && c % 8 == 0, "bn_apply: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = chan_reduce_grid(rows, c / 8, 4);
  PRN_DISPATCH(dtype,
               (bn_apply_kernel<__nv_bfloat16><<<grid, kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x16), static_cast<__nv_bfloat16*>(out16), mean_invstd, gamma, beta, static_cast<const __nv_bfloat16*>(residual16), rows, c, relu)),
               (bn_apply_kernel<__half><<<grid, kPwThreads, 0, st>>>(static_cast<const __half*>(x16), static_cast<__half*>(out16), mean_invstd, gamma, beta, static_cast<const __half*>(residual16), rows, c, relu)));
  PRN_LAUNCH_CHECK();
}

int prn_bn_finalize_apply(const void* x16, void* out16, const float* stats, float* mean_invstd, float* running_mean,
                          float* running_var, int64_t count, float eps, float momentum, const float* gamma, const float* beta,
                          const void* residual16, int64_t rows, int32_t c, int32_t relu, int32_t dtype, void* stream) {
  PRN_REQUIRE(x16 && out16 && stats && mean_invstd && gamma && beta && rows > 0 && count > 0 && c > 0 && c % 8 == 0 &&
                  (running_mean == nullptr) == (running_var == nullptr), "bn_finalize_apply: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = chan_reduce_grid(rows, c / 8, 4);
  PRN_REQUIRE(static_cast<long long>(grid) * kPwThreads >= c / 8, "bn_finalize_apply: grid smaller than the channel groups");
  const float cnt = static_cast<float>(count);
  PRN_DISPATCH(dtype,
               (bn_finalize_apply_kernel<__nv_bfloat16><<<grid, kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x16), static_cast<__nv_bfloat16*>(out16), stats, mean_invstd, running_mean, running_var, cnt, eps, momentum, gamma, beta, static_cast<const __nv_bfloat16*>(residual16), rows, c, relu)),
               (bn_finalize_apply_kernel<__half><<<grid, kPwThreads, 0, st>>>(static_cast<const __half*>(x16), static_cast<__half*>(out16), stats, mean_invstd, running_mean, running_var, cnt, eps, momentum, gamma, beta, static_cast<const __half*>(residual16), rows, c, relu)));
  PRN_LAUNCH_CHECK();
}

int prn_chan_reduce(const void* dz16, const void* out16, const void* x16, const float* mean_invstd, float* sums, int64_t rows,
                    int32_t c, int32_t dtype, void* stream) {
  PRN_REQUIRE(dz16 && sums && rows > 0 && c > 0 && c % 8 == 0 && c <= 4096 && (x16 == nullptr || mean_invstd != nullptr),
              "chan_reduce: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 4 CTAs of this kernel are resident per SM (registers): a larger grid only adds a second, partial wave (reduce_probe.cu)
  const int grid = chan_reduce_grid(rows, c / 8, 8, 4);
  const size_t smem = static_cast<size_t>(c) * 2 * sizeof(float);
  PRN_DISPATCH(dtype,
               (chan_reduce_kernel<__nv_bfloat16><<<grid, kPwThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(dz16), static_cast<const __nv_bfloat16*>(out16), static_cast<const __nv_bfloat16*>(x16), mean_invstd, sums, rows, c)),
               (chan_reduce_kernel<__half><<<grid, kPwThreads, smem, st>>>(static_cast<const __half*>(dz16), static_cast<const __half*>(out16), static_cast<const __half*>(x16), mean_invstd, sums, rows, c)));
  PRN_LAUNCH_CHECK();
}

int prn_bn_bwd_apply(const void* dz16, const void* out16, const void* x16, const float* mean_invstd, const float* gamma,
                     const float* sums, void* dx16, void* g_out16, int64_t rows, int32_t c, int32_t dtype, void* stream) {
  PRN_REQUIRE(dz16 && x16 && mean_invstd && gamma && sums && dx16 && rows > 0 && c > 0 && c % 8 == 0, "bn_bwd_apply: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = chan_reduce_grid(rows, c / 8, 4);
  const float inv_count = 1.0f / static_cast<float>(rows);
  PRN_DISPATCH(dtype,
               (bn_bwd_apply_kernel<__nv_bfloat16><<<grid, kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dz16), static_cast<const __nv_bfloat16*>(out16), static_cast<const __nv_bfloat16*>(x16), mean_invstd, gamma, sums, static_cast<__nv_bfloat16*>(dx16), static_cast<__nv_bfloat16*>(g_out16), rows, c, inv_count)),
               (bn_bwd_apply_kernel<__half><<<grid, kPwThreads, 0, st>>>(static_cast<const __half*>(dz16), static_cast<const __half*>(out16), static_cast<const __half*>(x16), mean_invstd, gamma, sums, static_cast<__half*>(dx16), static_cast<__half*>(g_out16), rows, c, inv_count)));
  PRN_LAUNCH_CHECK();
}

int prn_relu_bwd(const void* dz16, const void* out16, void* g16, int64_t n, int32_t dtype, void* stream) {
  PRN_REQUIRE(dz16 && out16 && g16 && n > 0 && n % 8 == 0, "relu_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (relu_bwd_kernel<__nv_bfloat16><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dz16), static_cast<const __nv_bfloat16*>(out16), static_cast<__nv_bfloat16*>(g16), n / 8)),
               (relu_bwd_kernel<__half><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __half*>(dz16), static_cast<const __half*>(out16), static_cast<__half*>(g16), n / 8)));
  PRN_LAUNCH_CHECK();
}

int prn_add_strided(void* dst16, const void* src16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t stride, int32_t dtype,
                    void* stream) {
  PRN_REQUIRE(dst16 && src16 && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && stride >= 1, "add_strided: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h * w * (c / 8);
  PRN_DISPATCH(dtype,
               (add_strided_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<__nv_bfloat16*>(dst16), static_cast<const __nv_bfloat16*>(src16), batch, h, w, c, stride)),
               (add_strided_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<__half*>(dst16), static_cast<const __half*>(src16), batch, h, w, c, stride)));
  PRN_LAUNCH_CHECK();
}

int prn_add_f32(const float* a32, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream) {
  PRN_REQUIRE(a32 && out16 && n > 0 && n % 8 == 0, "add_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (add_f32_kernel<__nv_bfloat16><<<pw_grid(n / 8), kPwThreads, 0, st>>>(a32, static_cast<const __nv_bfloat16*>(b16), static_cast<__nv_bfloat16*>(out16), n / 8)),
               (add_f32_kernel<__half><<<pw_grid(n / 8), kPwThreads, 0, st>>>(a32, static_cast<const __half*>(b16), static_cast<__half*>(out16), n / 8)));
  PRN_LAUNCH_CHECK();
}

int prn_add16(const void* a16, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream) {
  PRN_REQUIRE(a16 && b16 && out16 && n > 0 && n % 8 == 0, "add16: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (add16_kernel<__nv_bfloat16><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(a16), static_cast<const __nv_bfloat16*>(b16), static_cast<__nv_bfloat16*>(out16), n / 8)),
               (add16_kernel<__half><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __half*>(a16), static_cast<const __half*>(b16), static_cast<__half*>(out16), n / 8)));
  PRN_LAUNCH_CHECK();
}

int prn_maxpool3x3s2_bwd(const void* in16, const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c,
                         int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && dout16 && din16 && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
  PRN_DISPATCH(dtype,
               (maxpool3s2_bwd_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<const __nv_bfloat16*>(dout16), static_cast<__nv_bfloat16*>(din16), batch, h, w, c)),
               (maxpool3s2_bwd_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<const __half*>(dout16), static_cast<__half*>(din16), batch, h, w, c)));
  PRN_LAUNCH_CHECK();
}

static int dcn_geom(DcnGeom* g, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t stride, int32_t pad) {
  PRN_REQUIRE(batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0 && stride >= 1 && pad >= 0, "dcn: bad geometry");
  g->B = batch; g->H = h; g->W = w; g->C = c; g->stride = stride; g->pad = pad;
  g->Ho = (h + 2 * pad - 3) / stride + 1;
  g->Wo = (w + 2 * pad - 3) / stride + 1;
  return PRN_OK;
}

int prn_dcn_im2col(const void* x16, const float* offmask, void* col16, int32_t batch, int32_t h, int32_t w, int32_t c,
                   int32_t stride, int32_t pad, int32_t dtype, void* stream) {
  PRN_REQUIRE(x16 && offmask && col16, "dcn_im2col: NULL argument");
  DcnGeom g;
  int rc = dcn_geom(&g, batch, h, w, c, stride, pad);
  if (rc != PRN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * g.Ho * g.Wo * 9 * (c / 8);
  PRN_DISPATCH(dtype,
               (dcn_im2col_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x16), offmask, static_cast<__nv_bfloat16*>(col16), g)),
               (dcn_im2col_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(x16), offmask, static_cast<__half*>(col16), g)));
  PRN_LAUNCH_CHECK();
}

int prn_dcn_col2im_bwd(const void* x16, const float* offmask, const void* dcol16, float* dx32, void* dpre16, int32_t batch,
                       int32_t h, int32_t w, int32_t c, int32_t stride, int32_t pad, float clamp_bound, int32_t dtype, void* stream) {
  PRN_REQUIRE(x16 && offmask && dcol16 && dx32 && dpre16, "dcn_col2im_bwd: NULL argument");
  DcnGeom g;
  int rc = dcn_geom(&g, batch, h, w, c, stride, pad);
  if (rc != PRN_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * g.Ho * g.Wo * 32;
  PRN_DISPATCH(dtype,
               (dcn_col2im_bwd_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x16), offmask, static_cast<const __nv_bfloat16*>(dcol16), dx32, static_cast<__nv_bfloat16*>(dpre16), g, clamp_bound)),
               (dcn_col2im_bwd_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(x16), offmask, static_cast<const __half*>(dcol16), dx32, static_cast<__half*>(dpre16), g, clamp_bound)));
  PRN_LAUNCH_CHECK();
}

}  // extern "C"

// ---------------------------------------------------------------- weight packing (fp32 OIHW parameters -> 16-bit operands)
// A training step re-packs every weight (the optimizer changed it): one launch per operand instead of a chain of
// permute / pad / cast kernels.  One CTA per output row, the row's source elements staged through shared memory.
namespace prn {

struct PackSplit { int lo, real, pad; };

// forward operand: out[n][tap][off_s + c] = w[n][lo_s + c][tap] (c < real_s), zero elsewhere; rows n >= cout are zero
template <typename T>
__global__ void pack_fwd_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int kk, int cpad_tot,
                                int nsplit, PackSplit s0, PackSplit s1) {
  extern __shared__ float row[];   // [cin][kk]
  const int n = blockIdx.x;
  T* o = out + static_cast<long long>(n) * kk * cpad_tot;
  if (n >= cout) {
    for (int i = threadIdx.x; i < kk * cpad_tot; i += blockDim.x) o[i] = static_cast<T>(0.f);
    return;
  }
  const float* src = w + static_cast<long long>(n) * cin * kk;
  for (int i = threadIdx.x; i < cin * kk; i += blockDim.x) row[i] = __ldg(src + i);
  __syncthreads();
  for (int i = threadIdx.x; i < kk * cpad_tot; i += blockDim.x) {
    const int tap = i / cpad_tot, c = i - tap * cpad_tot;
    float v = 0.f;
    if (c < s0.pad) {
      if (c < s0.real) v = row[(s0.lo + c) * kk + tap];
    } else if (nsplit > 1) {
      const int c1 = c - s0.pad;
      if (c1 < s1.real) v = row[(s1.lo + c1) * kk + tap];
    }
    o[i] = static_cast<T>(v);
  }
}

// input-gradient operand: out[r][tap'][o] = w[o][lo + r][kk - 1 - tap'] (o < cout, r < hi - lo), zero elsewhere
template <typename T>
__global__ void pack_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int cout, int cin, int kk, int lo, int nreal,
                                  int cout_pad) {
  const int r = blockIdx.x;
  T* o = out + static_cast<long long>(r) * kk * cout_pad;
  for (int i = threadIdx.x; i < kk * cout_pad; i += blockDim.x) {
    const int tap = i / cout_pad, oc = i - tap * cout_pad;
    float v = 0.f;
    if (r < nreal && oc < cout) v = __ldg(w + (static_cast<long long>(oc) * cin + lo + r) * kk + (kk - 1 - tap));
    o[i] = static_cast<T>(v);
  }
}

// ---- every packed operand of a training step in ONE launch (the per-conv kernels above cost ~250 launches per step once the
//      optimizer has touched the weights).  recs[r] describes one operand; work[b] = {record, row} for block b.
struct PackRec {
  const float* w;
  void* out;
  int kind;            // 0 forward operand (pack_fwd_kernel), 1 input-gradient operand (pack_dgrad_kernel)
  int cout, cin, kk;
  int a, b, c, d, e, f, g;   // fwd: cpad_tot, nsplit, s0.lo, s0.real, s0.pad, s1.lo, s1.real (s1.pad = cpad_tot - s0.pad)
                              // dgrad: lo, nreal, cout_pad, rows (work row = first of 8 rows)
  int rows;                  // fwd: output rows of the operand (n_pad); a work item covers kPackRows consecutive rows
};
constexpr int kPackRows = 8;  // rows per thread block (64 776 one-row blocks cost 0.4 ms in block scheduling alone)

template <typename T>
__global__ void __launch_bounds__(256) pack_multi_kernel(const PackRec* __restrict__ recs, const int2* __restrict__ work) {
  extern __shared__ float row[];
  const int2 wk = work[blockIdx.x];
  const PackRec r = recs[wk.x];
  const int n = wk.y;
  const int kk = r.kk;
  if (r.kind == 0) {
    const int cpad_tot = r.a, nsplit = r.b;
    const PackSplit s0{r.c, r.d, r.e}, s1{r.f, r.g, cpad_tot - r.e};
    const int n_end = min(n + kPackRows, r.rows);
    for (int nn = n; nn < n_end; ++nn) {
      T* o = static_cast<T*>(r.out) + static_cast<long long>(nn) * kk * cpad_tot;
      if (nn >= r.cout) {
        for (int i = threadIdx.x; i < kk * cpad_tot; i += blockDim.x) o[i] = static_cast<T>(0.f);
        continue;
      }
      const float* src = r.w + static_cast<long long>(nn) * r.cin * kk;
      __syncthreads();                        // the previous row's readers are done with the staging buffer
      for (int i = threadIdx.x; i < r.cin * kk; i += blockDim.x) row[i] = __ldg(src + i);
      __syncthreads();
      for (int i = threadIdx.x; i < kk * cpad_tot; i += blockDim.x) {
        const int tap = i / cpad_tot, c = i - tap * cpad_tot;
        float v = 0.f;
        if (c < s0.pad) {
          if (c < s0.real) v = row[(s0.lo + c) * kk + tap];
        } else if (nsplit > 1) {
          const int c1 = c - s0.pad;
          if (c1 < s1.real) v = row[(s1.lo + c1) * kk + tap];
        }
        o[i] = static_cast<T>(v);
      }
    }
  } else {
    // a block packs 8 consecutive input-channel rows [n, n + 8): a thread's 8 source words w[oc][lo + n .. n + 7][tap] are
    // kk floats apart (one 32-byte sector for a 1x1 kernel), where one row per block read 4 bytes of every sector
    const int lo = r.a, nreal = r.b, cout_pad = r.c, rows = r.d;
    const int nq = min(8, rows - n);
    for (int i = threadIdx.x; i < kk * cout_pad; i += blockDim.x) {
      const int tap = i / cout_pad, oc = i - tap * cout_pad;
      const float* src = r.w + (static_cast<long long>(oc) * r.cin + lo + n) * kk + (kk - 1 - tap);
      T* o = static_cast<T*>(r.out) + static_cast<long long>(n) * kk * cout_pad + i;
      for (int q = 0; q < nq; ++q) {
        float v = 0.f;
        if (n + q < nreal && oc < r.cout) v = __ldg(src + q * kk);
        o[static_cast<long long>(q) * kk * cout_pad] = static_cast<T>(v);
      }
    }
  }
}

// ---- many strided fp32 vectors -> contiguous destinations in one launch (parameter gradients -> the flat all-reduce buffer)
struct CopyRec {
  const float* src;
  float* dst;
  long long n;
  int stride, pad_;
};
constexpr int kCopyChunk = 2048;

__global__ void __launch_bounds__(256) copy_multi_kernel(const CopyRec* __restrict__ recs, const int2* __restrict__ work, float scale) {
  const int2 wk = work[blockIdx.x];
  const CopyRec r = recs[wk.x];
  const long long lo = static_cast<long long>(wk.y) * kCopyChunk;
  const long long hi = min(lo + kCopyChunk, r.n);
  if (r.stride == 1 && ((reinterpret_cast<uintptr_t>(r.src) | reinterpret_cast<uintptr_t>(r.dst)) & 15) == 0) {
    const long long i4 = lo + threadIdx.x * 4LL;                 // lo is a multiple of 2048: 16-byte aligned on both sides
    for (long long i = i4; i < hi; i += 256 * 4) {
      if (i + 4 <= hi) {
        float4 v = __ldg(reinterpret_cast<const float4*>(r.src + i));
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *reinterpret_cast<float4*>(r.dst + i) = v;
      } else {
        for (long long j = i; j < hi; ++j) r.dst[j] = scale * __ldg(r.src + j);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += 256) r.dst[i] = scale * __ldg(r.src + i * r.stride);
  }
}

// ---- weight gradients: accumulator layout [cout rows][kh*kw][cpad] (the packed operand's layout) -> parameter layout
//      [cout][cin][kh*kw], many convs in one launch (a torch permute + contiguous per conv before: ~140 launches per step).
//      A block takes kUnpackRows output channels: each row goes through a padded shared-memory tile so that both the read of the
//      accumulator row and the write of the parameter row are coalesced.
struct UnpackRec {
  const float* src;
  float* dst;
  int cout, cin, kk, cpad, ld, pad_;
};
static_assert(sizeof(UnpackRec) == 40, "UnpackRec is a 40-byte record (ops.UnpackMulti packs it)");
constexpr int kUnpackRows = 8;

__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const UnpackRec* __restrict__ recs, const int2* __restrict__ work) {
  extern __shared__ float tile[];          // [kk][cpad + 1]
  const int2 wk = work[blockIdx.x];
  const UnpackRec r = recs[wk.x];
  const int pitch = r.cpad + 1;
  const int n_end = min(wk.y + kUnpackRows, r.cout);
  for (int n = wk.y; n < n_end; ++n) {
    const float* src = r.src + static_cast<long long>(n) * r.ld;
    __syncthreads();
    for (int i = threadIdx.x; i < r.kk * r.cpad; i += blockDim.x) {
      const int tap = i / r.cpad, c = i - tap * r.cpad;
      tile[tap * pitch + c] = __ldg(src + i);
    }
    __syncthreads();
    float* dst = r.dst + static_cast<long long>(n) * r.cin * r.kk;
    for (int i = threadIdx.x; i < r.cin * r.kk; i += blockDim.x) {
      const int c = i / r.kk, tap = i - c * r.kk;
      dst[i] = tile[tap * pitch + c];
    }
  }
}

}  // namespace prn

extern "C" int prn_unpack_wgrad_multi(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, int32_t smem_bytes, void* stream) {
  using namespace prn;
  PRN_REQUIRE(recs_dev && work_dev && n_blocks > 0 && smem_bytes > 0 && smem_bytes <= 48 * 1024, "unpack_wgrad_multi: bad arguments");
  unpack_wgrad_multi_kernel<<<n_blocks, 256, smem_bytes, static_cast<cudaStream_t>(stream)>>>(static_cast<const UnpackRec*>(recs_dev),
                                                                                             reinterpret_cast<const int2*>(work_dev));
  PRN_LAUNCH_CHECK();
}

extern "C" int prn_copy_multi_f32(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, float scale, void* stream) {
  using namespace prn;
  PRN_REQUIRE(recs_dev && work_dev && n_blocks > 0, "copy_multi_f32: bad arguments");
  copy_multi_kernel<<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const CopyRec*>(recs_dev),
                                                                            reinterpret_cast<const int2*>(work_dev), scale);
  PRN_LAUNCH_CHECK();
}

extern "C" int prn_pack_multi(const void* recs_dev, const int32_t* work_dev, int32_t n_blocks, int32_t smem_bytes, int32_t dtype,
                              void* stream) {
  using namespace prn;
  PRN_REQUIRE(recs_dev && work_dev && n_blocks > 0 && smem_bytes >= 0 && smem_bytes <= 48 * 1024, "pack_multi: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (pack_multi_kernel<__nv_bfloat16><<<n_blocks, 256, smem_bytes, st>>>(static_cast<const PackRec*>(recs_dev), reinterpret_cast<const int2*>(work_dev))),
               (pack_multi_kernel<__half><<<n_blocks, 256, smem_bytes, st>>>(static_cast<const PackRec*>(recs_dev), reinterpret_cast<const int2*>(work_dev))));
  PRN_LAUNCH_CHECK();
}

extern "C" int prn_pack_conv_weight(const float* w, void* out16, int32_t cout, int32_t cin, int32_t ksize, int32_t n_pad,
                                    int32_t nsplit, const int32_t* lo, const int32_t* real, const int32_t* pad, int32_t dtype,
                                    void* stream) {
  using namespace prn;
  PRN_REQUIRE(w && out16 && lo && real && pad && cout > 0 && cin > 0 && ksize >= 1 && ksize <= 7 && n_pad >= cout &&
                  (nsplit == 1 || nsplit == 2), "pack_conv_weight: bad arguments");
  PackSplit s0{lo[0], real[0], pad[0]}, s1{0, 0, 0};
  if (nsplit == 2) s1 = PackSplit{lo[1], real[1], pad[1]};
  PRN_REQUIRE(s0.lo >= 0 && s0.real >= 0 && s0.pad >= s0.real && s0.lo + s0.real <= cin && s1.lo + s1.real <= cin && s1.pad >= s1.real,
              "pack_conv_weight: bad channel split");
  const int kk = ksize * ksize, cpad_tot = s0.pad + s1.pad;
  const size_t smem = static_cast<size_t>(cin) * kk * sizeof(float);
  PRN_REQUIRE(smem <= 48 * 1024, "pack_conv_weight: cin*k*k too large (%d x %d)", cin, kk);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (pack_fwd_kernel<__nv_bfloat16><<<n_pad, 256, smem, st>>>(w, static_cast<__nv_bfloat16*>(out16), cout, cin, kk, cpad_tot, nsplit, s0, s1)),
               (pack_fwd_kernel<__half><<<n_pad, 256, smem, st>>>(w, static_cast<__half*>(out16), cout, cin, kk, cpad_tot, nsplit, s0, s1)));
  PRN_LAUNCH_CHECK();
}

extern "C" int prn_pack_dgrad_weight(const float* w, void* out16, int32_t cout, int32_t cin, int32_t ksize, int32_t lo, int32_t hi,
                                     int32_t rows_pad, int32_t cout_pad, int32_t dtype, void* stream) {
  using namespace prn;
  PRN_REQUIRE(w && out16 && cout > 0 && cin > 0 && ksize >= 1 && ksize <= 7 && lo >= 0 && hi > lo && hi <= cin &&
                  rows_pad >= hi - lo && cout_pad >= cout, "pack_dgrad_weight: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int kk = ksize * ksize;
  PRN_DISPATCH(dtype,
               (pack_dgrad_kernel<__nv_bfloat16><<<rows_pad, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(out16), cout, cin, kk, lo, hi - lo, cout_pad)),
               (pack_dgrad_kernel<__half><<<rows_pad, 256, 0, st>>>(w, static_cast<__half*>(out16), cout, cin, kk, lo, hi - lo, cout_pad)));
  PRN_LAUNCH_CHECK();
}

// ---------------------------------------------------------------- Adam over many tensors in one launch (train.py:251-256)
// torch.optim.Adam semantics (betas, eps, no weight decay, no amsgrad), one learning rate per tensor.  The step counter
// and bias corrections live in device memory so that the update can be replayed from a CUDA graph.
namespace prn {

__global__ void adam_advance_kernel(float* state, float beta1, float beta2, const int* __restrict__ found_inf) {
  // state = {step, 1 - beta1^step, sqrt(1 - beta2^step)}; a step with non-finite gradients is skipped entirely (GradScaler's rule)
  if (found_inf != nullptr && *found_inf != 0) return;
  const float step = state[0] + 1.f;
  state[0] = step;
  state[1] = 1.f - powf(beta1, step);
  state[2] = sqrtf(1.f - powf(beta2, step));
}

constexpr int kAdamChunk = 1 << 16;

// table[t] = {param, grad, exp_avg, exp_avg_sq, grad stride (elements)}; chunks[c] = {tensor, chunk index}
__global__ void __launch_bounds__(256) adam_multi_kernel(const long long* __restrict__ table, const long long* __restrict__ numel,
                                                       const float* __restrict__ lr, const int* __restrict__ chunks,
                                                       const float* __restrict__ state, float beta1, float beta2, float eps,
                                                       float inv_grad_scale, const int* __restrict__ found_inf) {
  if (found_inf != nullptr && *found_inf != 0) return;
  const int t = chunks[2 * blockIdx.x], ck = chunks[2 * blockIdx.x + 1];
  float* p = reinterpret_cast<float*>(table[5 * t]);
  const float* g = reinterpret_cast<const float*>(table[5 * t + 1]);
  float* m = reinterpret_cast<float*>(table[5 * t + 2]);
  float* v = reinterpret_cast<float*>(table[5 * t + 3]);
  const long long gs = table[5 * t + 4];
  const long long n = numel[t];
  const float step_size = __ldg(lr + t) / state[1];
  const float inv_bc2_sqrt = 1.f / state[2];
  const long long lo = static_cast<long long>(ck) * kAdamChunk;
  const long long hi = lo + kAdamChunk < n ? lo + kAdamChunk : n;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float gi = g[i * gs] * inv_grad_scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * inv_bc2_sqrt + eps);
  }
}

// *found_inf = 1 if any gradient element of the table is inf / NaN (the flag is cleared by a memset before this kernel)
__global__ void __launch_bounds__(256) grads_nonfinite_kernel(const long long* __restrict__ table, const long long* __restrict__ numel,
                                                            const int* __restrict__ chunks, int* __restrict__ found_inf) {
  const int t = chunks[2 * blockIdx.x], ck = chunks[2 * blockIdx.x + 1];
  const float* g = reinterpret_cast<const float*>(table[5 * t + 1]);
  const long long gs = table[5 * t + 4];
  const long long n = numel[t];
  const long long lo = static_cast<long long>(ck) * kAdamChunk;
  const long long hi = lo + kAdamChunk < n ? lo + kAdamChunk : n;
  bool bad = false;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float v = g[i * gs];
    bad = bad || !(fabsf(v) <= 3.4028234e38f);            // false for inf and NaN
  }
  if (__syncthreads_or(bad ? 1 : 0) && threadIdx.x == 0) *found_inf = 1;
}

}  // namespace prn

extern "C" int prn_adam_multi_checked(const int64_t* table, const int64_t* numel, const float* lr, const int32_t* chunks,
                                      int32_t n_chunks, float* state3, float beta1, float beta2, float eps, float grad_scale,
                                      int32_t* found_inf, void* stream) {
  using namespace prn;
  PRN_REQUIRE(table && numel && lr && chunks && state3 && found_inf && n_chunks > 0 && grad_scale > 0.f, "adam_multi_checked: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_CUDA(cudaMemsetAsync(found_inf, 0, sizeof(int32_t), st));
  grads_nonfinite_kernel<<<n_chunks, 256, 0, st>>>(reinterpret_cast<const long long*>(table), reinterpret_cast<const long long*>(numel),
                                                   chunks, found_inf);
  adam_advance_kernel<<<1, 1, 0, st>>>(state3, beta1, beta2, found_inf);
  adam_multi_kernel<<<n_chunks, 256, 0, st>>>(reinterpret_cast<const long long*>(table), reinterpret_cast<const long long*>(numel), lr,
                                              chunks, state3, beta1, beta2, eps, 1.0f / grad_scale, found_inf);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "adam_multi_checked launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}

extern "C" int prn_adam_multi(const int64_t* table, const int64_t* numel, const float* lr, const int32_t* chunks, int32_t n_chunks,
                              float* state3, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  using namespace prn;
  PRN_REQUIRE(table && numel && lr && chunks && state3 && n_chunks > 0 && grad_scale > 0.f, "adam_multi: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adam_advance_kernel<<<1, 1, 0, st>>>(state3, beta1, beta2, nullptr);
  adam_multi_kernel<<<n_chunks, 256, 0, st>>>(reinterpret_cast<const long long*>(table), reinterpret_cast<const long long*>(numel), lr,
                                              chunks, state3, beta1, beta2, eps, 1.0f / grad_scale, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "adam_multi launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}
