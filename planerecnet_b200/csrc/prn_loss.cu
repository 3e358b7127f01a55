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

// ---------------------------------------------------------------- dice + depth-gradient ("lava") terms over instance rows
// seg fp32 [rows][P] = sigmoid mask probabilities of the positive cells' dynamic convolutions (one row per instance,
// rows grouped per image: image = row / rows_per_img); target uint8 [rows][P]; gw fp32 [B][P] = per-image pixel weights of
// the lava term (the depth-gradient map pulled back through the bilinear x4 resampler).
//   stats[row] = {a = sum s*t, b = sum s*s, c = sum t*t, lv = sum s*gw}          (losses.py:355-368, 277-286)
namespace prn {

__global__ void __launch_bounds__(256) dice_lava_rows_kernel(const float* __restrict__ seg, const unsigned char* __restrict__ target,
                                                           const float* __restrict__ gw, float* __restrict__ stats, int P,
                                                           int rows_per_img) {
  const int row = blockIdx.x;
  const float* s = seg + static_cast<long long>(row) * P;
  const unsigned char* t = target + static_cast<long long>(row) * P;
  const float* g = gw + static_cast<long long>(row / rows_per_img) * P;
  float a = 0.f, b = 0.f, c = 0.f, lv = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float sv = __ldg(s + i), tv = static_cast<float>(t[i]);
    a = fmaf(sv, tv, a);
    b = fmaf(sv, sv, b);
    c = fmaf(tv, tv, c);
    lv = fmaf(sv, __ldg(g + i), lv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    lv += __shfl_xor_sync(0xffffffffu, lv, o);
  }
  __shared__ float red[8][4];
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a; red[threadIdx.x >> 5][1] = b; red[threadIdx.x >> 5][2] = c; red[threadIdx.x >> 5][3] = lv; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    stats[static_cast<long long>(row) * 4 + threadIdx.x] = v;
  }
}

// dx[row][p] = (ca*t + 2*cb*s + cl*gw) * s * (1 - s), coef[row] = {ca, cb, cl}; 16-bit output (operand of the two gradient
// contractions); rows with all-zero coefficients (padding) produce zeros
template <typename T>
__global__ void dice_lava_bwd_kernel(const float* __restrict__ seg, const unsigned char* __restrict__ target, const float* __restrict__ gw,
                                     const float* __restrict__ coef, T* __restrict__ dx, int P, int rows_per_img, long long total8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = i * 8;
    const int row = static_cast<int>(e / P), p = static_cast<int>(e - static_cast<long long>(row) * P);
    const float ca = __ldg(coef + 3 * row), cb = __ldg(coef + 3 * row + 1), cl = __ldg(coef + 3 * row + 2);
    const float* g = gw + static_cast<long long>(row / rows_per_img) * P + p;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sv = __ldg(seg + e + j), tv = static_cast<float>(target[e + j]);
      o[j] = (ca * tv + 2.f * cb * sv + cl * __ldg(g + j)) * sv * (1.f - sv);
    }
    store8(dx + e, o);
  }
}

// Lava pixel weights: gmap = clamp(sobel(gt)^2-sum / max(gt, res)^2, max 1e-2), zeroed below 1e-4 (losses.py:288-329, 186-188),
// pulled back through F.interpolate(seg, size=(H, W), bilinear, align_corners=False) from [h, w] = [H/4, W/4]:
// gw[b][y][x] += weight of (Y, X) on (y, x) * gmap[b][Y][X];  gsum[b] += gmap.  One thread per full-resolution pixel.
__global__ void lava_weights_kernel(const float* __restrict__ gt, float* __restrict__ gw, float* __restrict__ gsum, int H, int W, int h,
                                    int w, float depth_res) {
  const int b = blockIdx.y;
  const float* d = gt + static_cast<long long>(b) * H * W;
  float* o = gw + static_cast<long long>(b) * h * w;
  const float sy_scale = static_cast<float>(h) / static_cast<float>(H), sx_scale = static_cast<float>(w) / static_cast<float>(W);
  float local = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int Y = i / W, X = i % W;
    auto at = [&](int yy, int xx) {      // reflect padding by 1
      yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
      xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
      return __ldg(d + yy * W + xx);
    };
    const float gx = (at(Y - 1, X - 1) - at(Y - 1, X + 1) + 2.f * (at(Y, X - 1) - at(Y, X + 1)) + at(Y + 1, X - 1) - at(Y + 1, X + 1)) * 0.125f;
    const float gy = (at(Y - 1, X - 1) + 2.f * at(Y - 1, X) + at(Y - 1, X + 1) - at(Y + 1, X - 1) - 2.f * at(Y + 1, X) - at(Y + 1, X + 1)) * 0.125f;
    const float dc = fmaxf(at(Y, X), depth_res);
    float gm = (gx * gx + gy * gy) / (dc * dc);
    gm = fminf(gm, 1e-2f);
    if (gm < 1e-4f) gm = 0.f;
    if (gm == 0.f) continue;
    local += gm;
    // source coordinates of the resampler at (Y, X)
    float fy = (static_cast<float>(Y) + 0.5f) * sy_scale - 0.5f, fx = (static_cast<float>(X) + 0.5f) * sx_scale - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = y0 > h - 1 ? h - 1 : y0;
    x0 = x0 > w - 1 ? w - 1 : x0;
    const int y1 = y0 < h - 1 ? y0 + 1 : y0, x1 = x0 < w - 1 ? x0 + 1 : x0;
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    atomicAdd(o + y0 * w + x0, gm * (1.f - ly) * (1.f - lx));
    atomicAdd(o + y0 * w + x1, gm * (1.f - ly) * lx);
    atomicAdd(o + y1 * w + x0, gm * ly * (1.f - lx));
    atomicAdd(o + y1 * w + x1, gm * ly * lx);
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) local += __shfl_xor_sync(0xffffffffu, local, k);
  if ((threadIdx.x & 31) == 0 && local != 0.f) atomicAdd(gsum + b, local);
}

}  // namespace prn

extern "C" {

int prn_dice_lava_rows(const float* seg, const uint8_t* target, const float* gw, float* stats, int32_t rows, int32_t pixels,
                       int32_t rows_per_img, void* stream) {
  PRN_REQUIRE(seg && target && gw && stats && rows > 0 && pixels > 0 && rows_per_img > 0 && rows % rows_per_img == 0,
              "dice_lava_rows: bad arguments");
  dice_lava_rows_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(seg, target, gw, stats, pixels, rows_per_img);
  PRN_LAUNCH_CHECK();
}

int prn_dice_lava_bwd(const float* seg, const uint8_t* target, const float* gw, const float* coef, void* dx16, int32_t rows,
                      int32_t pixels, int32_t rows_per_img, int32_t dtype, void* stream) {
  PRN_REQUIRE(seg && target && gw && coef && dx16 && rows > 0 && pixels > 0 && pixels % 8 == 0 && rows_per_img > 0 &&
                  rows % rows_per_img == 0, "dice_lava_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total8 = static_cast<long long>(rows) * pixels / 8;
  PRN_DISPATCH(dtype,
               (dice_lava_bwd_kernel<__nv_bfloat16><<<pw_grid(total8), kPwThreads, 0, st>>>(seg, target, gw, coef, static_cast<__nv_bfloat16*>(dx16), pixels, rows_per_img, total8)),
               (dice_lava_bwd_kernel<__half><<<pw_grid(total8), kPwThreads, 0, st>>>(seg, target, gw, coef, static_cast<__half*>(dx16), pixels, rows_per_img, total8)));
  PRN_LAUNCH_CHECK();
}

int prn_lava_weights(const float* gt, float* gw, float* gsum, int32_t batch, int32_t H, int32_t W, int32_t h, int32_t w, float depth_res,
                     void* stream) {
  PRN_REQUIRE(gt && gw && gsum && batch > 0 && H > 2 && W > 2 && h > 0 && w > 0, "lava_weights: bad arguments");
  const dim3 grid(static_cast<unsigned>(pw_grid(static_cast<long long>(H) * W / 4 + 1)), static_cast<unsigned>(batch));
  lava_weights_kernel<<<grid, kPwThreads, 0, static_cast<cudaStream_t>(stream)>>>(gt, gw, gsum, H, W, h, w, depth_res);
  PRN_LAUNCH_CHECK();
}

}  // extern "C"

// ---------------------------------------------------------------- plane surface-normal term: per-triplet geometry
// models/functions/vnl.py:20-165 for ALL sampled triplets of a batch (one thread per triplet): the three 3-D points of the
// prediction and of the ground truth (u0 / v0 = image centre), the selection mask (in front, not colinear, not too close:
// vnl.py:56-98) and the per-triplet loss 1 - |cos| between the triplet's normal and the plane's ground-truth normal (plane
// regions, float64 like the reference's plane parameters) resp. the ground-truth triplet's normal (the non-planar rest, fp32).
// pix int64 [3][T] = global pixel (image * H * W + y * W + x) of the three points; region int64 [T]; rest uint8 [R];
// tgt float64 [R][3]; fxfy fp32 [B][2].  The backward recomputes the geometry from the depth map and adds
// c[t] * d(loss_t)/d(depth) into d_depth with fp32 reductions (three pixels per triplet).
namespace prn {

struct VnlGeom {
  int H, W;
  float delta_z;
};

__device__ __forceinline__ void vnl_point(const float* __restrict__ depth, const float* __restrict__ fxfy, long long g, const VnlGeom& q,
                                          float* p, float* jac) {
  const long long hw = static_cast<long long>(q.H) * q.W;
  const int b = static_cast<int>(g / hw);
  const int r = static_cast<int>(g - b * hw);
  const int y = r / q.W, x = r - y * q.W;
  const float u = static_cast<float>(x - q.W / 2), v = static_cast<float>(y - q.H / 2);
  const float fx = __ldg(fxfy + 2 * b), fy = __ldg(fxfy + 2 * b + 1);
  const float d = __ldg(depth + g);
  p[0] = __fdiv_rn(__fmul_rn(u, fabsf(d)), fx);
  p[1] = __fdiv_rn(__fmul_rn(v, fabsf(d)), fy);
  p[2] = d;
  if (jac != nullptr) {       // d(point)/d(depth)
    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    jac[0] = u * sg / fx;
    jac[1] = v * sg / fy;
    jac[2] = 1.f;
  }
}

__device__ __forceinline__ float sum3(float a, float b, float c) { return __fadd_rn(__fadd_rn(a, b), c); }

// nv = (p1 - p0) x (p2 - p0), unit-ish normal n = nv / (|nv| + (|nv| == 0) * 0.01)      (vnl.py:100-108)
__device__ __forceinline__ void vnl_normal(const float (*p)[3], float* nv, float* n, float* nrm_out, float* s_out) {
  const float a0 = p[1][0] - p[0][0], a1 = p[1][1] - p[0][1], a2 = p[1][2] - p[0][2];
  const float b0 = p[2][0] - p[0][0], b1 = p[2][1] - p[0][1], b2 = p[2][2] - p[0][2];
  nv[0] = __fsub_rn(__fmul_rn(a1, b2), __fmul_rn(a2, b1));
  nv[1] = __fsub_rn(__fmul_rn(a2, b0), __fmul_rn(a0, b2));
  nv[2] = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
  const float nrm = sqrtf(sum3(__fmul_rn(nv[0], nv[0]), __fmul_rn(nv[1], nv[1]), __fmul_rn(nv[2], nv[2])));
  const float s = nrm + (nrm == 0.f ? 0.01f : 0.f);
  n[0] = nv[0] / s; n[1] = nv[1] / s; n[2] = nv[2] / s;
  *nrm_out = nrm;
  *s_out = s;
}

// vnl.py:146 quirk: a predicted point i with z == 0 sets COORDINATE i of all three points to 1e-4 (never true behind a softplus)
__device__ __forceinline__ void vnl_rest_quirk(float (*p)[3], bool* fixed) {
  const bool z0 = p[0][2] == 0.f, z1 = p[1][2] == 0.f, z2 = p[2][2] == 0.f;
  fixed[0] = z0; fixed[1] = z1; fixed[2] = z2;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    if (fixed[c]) { p[0][c] = 0.0001f; p[1][c] = 0.0001f; p[2][c] = 0.0001f; }
}

__global__ void vnl_triplet_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ fxfy,
                                       const long long* __restrict__ pix, const long long* __restrict__ region,
                                       const unsigned char* __restrict__ rest, const double* __restrict__ tgt, double* __restrict__ loss_t,
                                       unsigned char* __restrict__ keep, long long T, VnlGeom q) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < T;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = __ldg(region + t);
    const bool is_rest = __ldg(rest + r) != 0;
    float pp[3][3], pg[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const long long g = __ldg(pix + i * T + t);
      vnl_point(pred, fxfy, g, q, pp[i], nullptr);
      vnl_point(gt, fxfy, g, q, pg[i], nullptr);
    }
    const float (*ts)[3] = is_rest ? pg : pp;            // planes are tested on the prediction, the rest on the ground truth
    // ---- selection (vnl.py:56-98)
    float df[3][3];                                       // [pair][xyz]: p1 - p0, p2 - p0, p2 - p1
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      df[0][c] = ts[1][c] - ts[0][c];
      df[1][c] = ts[2][c] - ts[0][c];
      df[2][c] = ts[2][c] - ts[1][c];
    }
    float qn[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) qn[k] = sqrtf(sum3(__fmul_rn(df[k][0], df[k][0]), __fmul_rn(df[k][1], df[k][1]), __fmul_rn(df[k][2], df[k][2])));
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float dot = sum3(__fmul_rn(df[i][0], df[j][0]), __fmul_rn(df[i][1], df[j][1]), __fmul_rn(df[i][2], df[j][2]));
        const float cs = __fdiv_rn(dot, __fadd_rn(__fmul_rn(qn[i], qn[j]), 1e-8f));
        cnt += (cs > 0.985f || cs < -0.985f) ? 1 : 0;
      }
    const bool colinear = cnt > 3;
    const bool in_front = ts[0][2] > q.delta_z && ts[1][2] > q.delta_z && ts[2][2] > q.delta_z;
    const float dd = is_rest ? 0.1f : 0.005f;
    bool near = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) near = near && (fabsf(df[0][c]) < dd || fabsf(df[1][c]) < dd || fabsf(df[2][c]) < dd);
    keep[t] = (in_front && !(near || colinear)) ? 1 : 0;
    // ---- loss of the triplet
    float nv[3], n[3], nrm, s;
    double l;
    if (!is_rest) {
      vnl_normal(pp, nv, n, &nrm, &s);
      const double t0 = __ldg(tgt + 3 * r), t1 = __ldg(tgt + 3 * r + 1), t2 = __ldg(tgt + 3 * r + 2);
      const double n1 = fmax(sqrt(static_cast<double>(n[0]) * n[0] + static_cast<double>(n[1]) * n[1] + static_cast<double>(n[2]) * n[2]), 1e-8);
      const double n2 = fmax(sqrt(t0 * t0 + t1 * t1 + t2 * t2), 1e-8);
      const double cs = (n[0] / n1) * (t0 / n2) + (n[1] / n1) * (t1 / n2) + (n[2] / n1) * (t2 / n2);
      l = 1.0 - fabs(cs);
    } else {
      bool fixed[3];
      vnl_rest_quirk(pp, fixed);
      vnl_normal(pp, nv, n, &nrm, &s);
      float gv[3], gn[3], gnrm, gs;
      vnl_normal(pg, gv, gn, &gnrm, &gs);
      const float n1 = fmaxf(sqrtf(sum3(n[0] * n[0], n[1] * n[1], n[2] * n[2])), 1e-8f);
      const float n2 = fmaxf(sqrtf(sum3(gn[0] * gn[0], gn[1] * gn[1], gn[2] * gn[2])), 1e-8f);
      const float cs = sum3((n[0] / n1) * (gn[0] / n2), (n[1] / n1) * (gn[1] / n2), (n[2] / n1) * (gn[2] / n2));
      l = static_cast<double>(1.f - fabsf(cs));
    }
    loss_t[t] = l;
  }
}

__global__ void vnl_triplet_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ fxfy,
                                       const long long* __restrict__ pix, const long long* __restrict__ region,
                                       const unsigned char* __restrict__ rest, const double* __restrict__ tgt, const double* __restrict__ coef,
                                       float* __restrict__ d_pred, long long T, VnlGeom q) {
  for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < T;
       t += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double c = __ldg(coef + t);
    if (c == 0.0 || c != c) continue;
    const long long r = __ldg(region + t);
    const bool is_rest = __ldg(rest + r) != 0;
    float pp[3][3], jac[3][3];
    long long gpix[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      gpix[i] = __ldg(pix + i * T + t);
      vnl_point(pred, fxfy, gpix[i], q, pp[i], jac[i]);
    }
    bool fixed[3] = {false, false, false};
    double x2[3];
    if (is_rest) {
      float pg[3][3], gv[3], gn[3], gnrm, gs;
#pragma unroll
      for (int i = 0; i < 3; ++i) vnl_point(gt, fxfy, gpix[i], q, pg[i], nullptr);
      vnl_normal(pg, gv, gn, &gnrm, &gs);
      x2[0] = gn[0]; x2[1] = gn[1]; x2[2] = gn[2];
      vnl_rest_quirk(pp, fixed);
    } else {
      x2[0] = __ldg(tgt + 3 * r); x2[1] = __ldg(tgt + 3 * r + 1); x2[2] = __ldg(tgt + 3 * r + 2);
    }
    float nvf[3], nf[3], nrmf, sf;
    vnl_normal(pp, nvf, nf, &nrmf, &sf);
    const double x1[3] = {nf[0], nf[1], nf[2]};
    const double n1 = sqrt(x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2]), n2 = sqrt(x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2]);
    const double n1c = fmax(n1, 1e-8), n2c = fmax(n2, 1e-8);
    const double cs = (x1[0] / n1c) * (x2[0] / n2c) + (x1[1] / n1c) * (x2[1] / n2c) + (x1[2] / n1c) * (x2[2] / n2c);
    const double sg = cs > 0.0 ? 1.0 : (cs < 0.0 ? -1.0 : 0.0);
    if (sg == 0.0) continue;
    // d(loss)/d(x1): loss = 1 - |cos|, cos = sum (x1 / N1) (x2 / N2) with N1 = |x1| (its clamp carries no gradient)
    double g1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g1[k] = -sg * c * ((x2[k] / n2c) / n1c - (n1 > 0.0 ? cs * x1[k] / (n1c * n1) : 0.0));
    // through n = nv / s, s = |nv| + const
    const double s = sf, nrm = nrmf;
    const double nv[3] = {nvf[0], nvf[1], nvf[2]};
    const double gdotn = g1[0] * nv[0] + g1[1] * nv[1] + g1[2] * nv[2];
    double gnv[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) gnv[k] = g1[k] / s - (nrm > 0.0 ? gdotn / (s * s) * nv[k] / nrm : 0.0);
    // through the cross product nv = A x B, A = p1 - p0, B = p2 - p0
    const double A[3] = {static_cast<double>(pp[1][0]) - pp[0][0], static_cast<double>(pp[1][1]) - pp[0][1], static_cast<double>(pp[1][2]) - pp[0][2]};
    const double Bv[3] = {static_cast<double>(pp[2][0]) - pp[0][0], static_cast<double>(pp[2][1]) - pp[0][1], static_cast<double>(pp[2][2]) - pp[0][2]};
    const double gA[3] = {Bv[1] * gnv[2] - Bv[2] * gnv[1], Bv[2] * gnv[0] - Bv[0] * gnv[2], Bv[0] * gnv[1] - Bv[1] * gnv[0]};
    const double gB[3] = {gnv[1] * A[2] - gnv[2] * A[1], gnv[2] * A[0] - gnv[0] * A[2], gnv[0] * A[1] - gnv[1] * A[0]};
    double gp[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      gp[1][k] = gA[k];
      gp[2][k] = gB[k];
      gp[0][k] = -(gA[k] + gB[k]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double gd = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (!fixed[k]) gd += gp[i][k] * jac[i][k];
      atomicAdd(d_pred + gpix[i], static_cast<float>(gd));
    }
  }
}

}  // namespace prn

extern "C" {

int prn_vnl_triplets_fwd(const float* pred, const float* gt, const float* fxfy, const int64_t* pix, const int64_t* region,
                         const uint8_t* rest, const double* tgt, double* loss_t, uint8_t* keep, int64_t n_triplets, int32_t h, int32_t w,
                         float delta_z, void* stream) {
  using namespace prn;
  PRN_REQUIRE(pred && gt && fxfy && pix && region && rest && tgt && loss_t && keep && n_triplets >= 0 && h > 0 && w > 0,
              "vnl_triplets_fwd: bad arguments");
  if (n_triplets == 0) return PRN_OK;
  VnlGeom q{h, w, delta_z};
  vnl_triplet_fwd_kernel<<<pw_grid(n_triplets), kPwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, gt, fxfy, reinterpret_cast<const long long*>(pix), reinterpret_cast<const long long*>(region), rest, tgt, loss_t, keep,
      n_triplets, q);
  PRN_LAUNCH_CHECK();
}

int prn_vnl_triplets_bwd(const float* pred, const float* gt, const float* fxfy, const int64_t* pix, const int64_t* region,
                         const uint8_t* rest, const double* tgt, const double* coef, float* d_pred, int64_t n_triplets, int32_t h,
                         int32_t w, float delta_z, void* stream) {
  using namespace prn;
  PRN_REQUIRE(pred && gt && fxfy && pix && region && rest && tgt && coef && d_pred && n_triplets >= 0 && h > 0 && w > 0,
              "vnl_triplets_bwd: bad arguments");
  if (n_triplets == 0) return PRN_OK;
  VnlGeom q{h, w, delta_z};
  vnl_triplet_bwd_kernel<<<pw_grid(n_triplets), kPwThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, gt, fxfy, reinterpret_cast<const long long*>(pix), reinterpret_cast<const long long*>(region), rest, tgt, coef, d_pred,
      n_triplets, q);
  PRN_LAUNCH_CHECK();
}

}  // extern "C"
