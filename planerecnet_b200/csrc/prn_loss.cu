// prn_loss.cu — dense parts of PlaneRecNetLoss (models/functions/losses.py), first two terms:
//   * sigmoid focal category loss over all grid cells (losses.py:121-138, 331-352): value and d/d(logits) in one pass
//   * RMSE-log depth loss on the x2 bilinear-upsampled prediction (losses.py:141-147, 371-392): reduce, finalize,
//     and the gradient scattered back through the resampler
// HBM-bound single passes over small tensors; fp32 throughout (the loss is not a 16-bit quantity).
#include <math.h>
#include "prn_pw.cuh"

namespace prn {

// ---------------------------------------------------------------- sigmoid focal loss, reduction = sum
// logits fp32 [n][ld] (first nc columns valid), labels int64 [n] in [0, nc] (nc = background -> all-zero one-hot row).
// loss_sum += sum alpha_t * (1 - p_t)^gamma * (-log p_t);  dlogits [n][nc] = d(loss_sum)/d(logits) (optional).
// With z = +x for the positive class and -x otherwise: p_t = sigmoid(z), s = 1 - p_t = sigmoid(-z),
//   d/dx = sign * alpha_t * (gamma * p_t * s^gamma * log p_t - s^(gamma+1)).
__global__ void focal_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, float alpha, float gamma,
                                  float* __restrict__ loss_sum, float* __restrict__ dlogits, long long n, int ld, int nc) {
  float acc = 0.f;
  const long long total = n * nc;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nc;
    const int c = static_cast<int>(i - row * nc);
    const float x = __ldg(logits + row * ld + c);
    const bool pos = __ldg(labels + row) == c;
    const float z = pos ? x : -x;
    const float logpt = -(fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z))));       // log sigmoid(z)
    const float pt = expf(logpt);
    const float s = 1.f / (1.f + expf(z));                                 // sigmoid(-z) = 1 - p_t
    const float a_t = alpha >= 0.f ? (pos ? alpha : 1.f - alpha) : 1.f;
    const float sg = powf(s, gamma);
    acc += a_t * sg * (-logpt);
    if (dlogits != nullptr) dlogits[i] = (pos ? 1.f : -1.f) * a_t * (gamma * pt * sg * logpt - sg * s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(loss_sum, t);
  }
}

// ---------------------------------------------------------------- RMSE-log depth loss on the x2 upsampled prediction
__device__ __forceinline__ void up2_index(int dst, int in_size, int* i0, int* i1, float* l) {
  float s = (static_cast<float>(dst) + 0.5f) * 0.5f - 0.5f;      // F.interpolate(scale_factor=2, bilinear, align_corners=False)
  s = s < 0.f ? 0.f : s;
  int a = static_cast<int>(s);
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a < in_size - 1 ? a + 1 : a;
  *l = s - static_cast<float>(a);
}

__device__ __forceinline__ float up2_sample(const float* __restrict__ img, int h, int w, int Y, int X, int* y0, int* y1, int* x0,
                                            int* x1, float* ly, float* lx) {
  up2_index(Y, h, y0, y1, ly);
  up2_index(X, w, x0, x1, lx);
  const float a = __ldg(img + *y0 * w + *x0), b = __ldg(img + *y0 * w + *x1);
  const float c = __ldg(img + *y1 * w + *x0), d = __ldg(img + *y1 * w + *x1);
  return (1.f - *ly) * ((1.f - *lx) * a + *lx * b) + *ly * ((1.f - *lx) * c + *lx * d);
}

// sums[b*2 + {0,1}] += {sum over valid pixels of (log d_up - log gt)^2, number of valid pixels}; valid = gt > min_depth
__global__ void depth_rmselog_reduce_kernel(const float* __restrict__ depth, const float* __restrict__ gt, float* __restrict__ sums,
                                            int h, int w, float min_depth, float clamp_val) {
  const int b = blockIdx.y;
  const int H = 2 * h, W = 2 * w;
  const float* img = depth + static_cast<long long>(b) * h * w;
  const float* g = gt + static_cast<long long>(b) * H * W;
  float s = 0.f, v = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const float gv = __ldg(g + i);
    if (gv > min_depth) {
      int y0, y1, x0, x1;
      float ly, lx;
      const float d = up2_sample(img, h, w, i / W, i % W, &y0, &y1, &x0, &x1, &ly, &lx);
      const float df = logf(fmaxf(d, clamp_val)) - logf(fmaxf(gv, clamp_val));
      s += df * df;
      v += 1.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __shared__ float ps[8], pv[8];
  if ((threadIdx.x & 31) == 0) { ps[threadIdx.x >> 5] = s; pv[threadIdx.x >> 5] = v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tv = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { ts += ps[k]; tv += pv[k]; }
    atomicAdd(sums + 2 * b, ts);
    atomicAdd(sums + 2 * b + 1, tv);
  }
}

// loss = weight * mean_b sqrt(S_b / V_b); coef[b] = d loss / d S_b * 2 = weight / (B * sqrt(S_b * V_b))
__global__ void depth_rmselog_finalize_kernel(const float* __restrict__ sums, float* __restrict__ loss, float* __restrict__ coef,
                                              int B, float weight) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float S = sums[2 * b], V = sums[2 * b + 1];
    acc += sqrtf(S / V);
    coef[b] = weight / (static_cast<float>(B) * sqrtf(S * V));
  }
  *loss = weight * acc / static_cast<float>(B);
}

// d_depth (zero-initialised by the caller) += transpose of the x2 bilinear resampler applied to
// coef[b] * (log d_up - log gt) / d_up on the valid pixels (0 where d_up is below the clamp)
__global__ void depth_rmselog_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ gt, const float* __restrict__ coef,
                                         float* __restrict__ d_depth, int h, int w, float min_depth, float clamp_val) {
  const int b = blockIdx.y;
  const int H = 2 * h, W = 2 * w;
  const float* img = depth + static_cast<long long>(b) * h * w;
  const float* g = gt + static_cast<long long>(b) * H * W;
  float* dd = d_depth + static_cast<long long>(b) * h * w;
  const float cf = __ldg(coef + b);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const float gv = __ldg(g + i);
    if (!(gv > min_depth)) continue;
    int y0, y1, x0, x1;
    float ly, lx;
    const float d = up2_sample(img, h, w, i / W, i % W, &y0, &y1, &x0, &x1, &ly, &lx);
    if (!(d > clamp_val)) continue;
    const float gr = cf * (logf(d) - logf(fmaxf(gv, clamp_val))) / d;
    atomicAdd(dd + y0 * w + x0, gr * (1.f - ly) * (1.f - lx));
    atomicAdd(dd + y0 * w + x1, gr * (1.f - ly) * lx);
    atomicAdd(dd + y1 * w + x0, gr * ly * (1.f - lx));
    atomicAdd(dd + y1 * w + x1, gr * ly * lx);
  }
}

}  // namespace prn

using namespace prn;

extern "C" {

int prn_focal_loss(const float* logits, const int64_t* labels, float alpha, float gamma, float* loss_sum, float* dlogits, int64_t n,
                   int32_t ld, int32_t nc, void* stream) {
  PRN_REQUIRE(logits && labels && loss_sum && n > 0 && nc > 0 && ld >= nc, "focal_loss: bad arguments");
  focal_loss_kernel<<<pw_grid(n * nc), kPwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, reinterpret_cast<const long long*>(labels), alpha, gamma, loss_sum, dlogits, n, ld, nc);
  PRN_LAUNCH_CHECK();
}

int prn_depth_rmselog_fwd(const float* depth, const float* gt, float* sums, float* loss, float* coef, int32_t batch, int32_t h,
                          int32_t w, float min_depth, float clamp_val, float weight, void* stream) {
  PRN_REQUIRE(depth && gt && sums && loss && coef && batch > 0 && h > 0 && w > 0, "depth_rmselog_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>(pw_grid(4LL * h * w / 4 + 1)), static_cast<unsigned>(batch));
  depth_rmselog_reduce_kernel<<<grid, kPwThreads, 0, st>>>(depth, gt, sums, h, w, min_depth, clamp_val);
  depth_rmselog_finalize_kernel<<<1, 32, 0, st>>>(sums, loss, coef, batch, weight);
  PRN_LAUNCH_CHECK();
}

int prn_depth_rmselog_bwd(const float* depth, const float* gt, const float* coef, float* d_depth, int32_t batch, int32_t h, int32_t w,
                          float min_depth, float clamp_val, void* stream) {
  PRN_REQUIRE(depth && gt && coef && d_depth && batch > 0 && h > 0 && w > 0, "depth_rmselog_bwd: bad arguments");
  const dim3 grid(static_cast<unsigned>(pw_grid(4LL * h * w / 4 + 1)), static_cast<unsigned>(batch));
  depth_rmselog_bwd_kernel<<<grid, kPwThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, gt, coef, d_depth, h, w, min_depth,
                                                                                     clamp_val);
  PRN_LAUNCH_CHECK();
}

}  // extern "C"
