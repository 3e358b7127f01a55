// prn_pointwise.cu — the HBM-bound passes between the tensor-core contractions: layout changes,
// pooling, bilinear resampling, GroupNorm application.  NHWC, 16-bit, 8 channels (16 bytes) per
// thread, fully coalesced; grids sized in whole waves where the tensor is large enough.
#include "prn_internal.h"
#include "prn_ptx.cuh"
#include "prn_pw.cuh"

namespace prn {

// ---------------------------------------------------------------- stem im2col (models/backbone.py:101,200)
// x NCHW fp32 [B,3,H,W] -> A [B*Ho*Wo, 192] 16-bit, k = (ky*7 + kx)*3 + c for the 7x7/s2/p3 conv, zero
// padded to 192 so the stem runs as a K=192 GEMM on the tensor cores.
// One CTA = 64 consecutive output pixels of one output row.  Stage 1 copies the 3 x 7 x 133 fp32 input patch
// they read into shared memory with coalesced row loads; stage 2 emits the [64][192] rows with 16-byte
// stores that are contiguous per pixel (384 B), so both sides of the pass stream at full sector width.
constexpr int kStemPx = 64;
constexpr int kStemPatchW = 2 * kStemPx + 5;   // 133 input columns
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int H, int W) {
  __shared__ float patch[21 * (kStemPatchW + 3)];
  __shared__ __align__(16) short koff[192];      // k -> offset of tap (c, ky, kx) inside the patch, -1 for the zero padding
  constexpr int kPW = kStemPatchW + 3;
  const int Ho = H / 2, Wo = W / 2;
  const int strips = (Wo + kStemPx - 1) / kStemPx;
  const int strip = blockIdx.x % strips;
  const int ho = (blockIdx.x / strips) % Ho;
  const int b = blockIdx.x / (strips * Ho);
  const int wo0 = strip * kStemPx;
  const int x0 = 2 * wo0 - 3, y0 = 2 * ho - 3;
  if (threadIdx.x < 192) {
    const int k = threadIdx.x;
    const int c = k % 3, tap = k / 3;
    koff[k] = k < 147 ? static_cast<short>((c * 7 + tap / 7) * kPW + tap % 7) : static_cast<short>(-1);
  }
  for (int i = threadIdx.x; i < 21 * kStemPatchW; i += blockDim.x) {
    const int col = i % kStemPatchW, rc = i / kStemPatchW;
    const int ky = rc % 7, c = rc / 7;
    const int yy = y0 + ky, xx = x0 + col;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(x + ((static_cast<long long>(b) * 3 + c) * H + yy) * W + xx);
    patch[rc * kPW + col] = v;
  }
  __syncthreads();
  const int npx = min(kStemPx, Wo - wo0);
  for (int i = threadIdx.x; i < npx * 24; i += blockDim.x) {
    const int kc = i % 24, px = i / 24;
    const uint4 ko = *reinterpret_cast<const uint4*>(koff + kc * 8);
    const short* o = reinterpret_cast<const short*>(&ko);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = o[j] >= 0 ? patch[o[j] + 2 * px] : 0.f;
    const long long m = (static_cast<long long>(b) * Ho + ho) * Wo + wo0 + px;
    store8(out + m * 192 + kc * 8, f);
  }
}

// Same im2col straight from the CAMERA image: BGR, NHWC, uint8 or float [B, H_img, W_img, 3] with values 0..255 — the input
// of FastBaseTransform (data/augmentations.py:496-530).  The transform ((x - MEANS) / STD per BGR channel, then BGR -> RGB) and
// pad_even_divided (models/functions/funcs.py:204-210: zero-valued raw pixels up to the next multiple of 32, i.e.
// -mean/std after normalisation) are folded into the patch load; pixels outside the padded extent are the conv's zero padding.
template <typename T, typename U>
__global__ void __launch_bounds__(256) stem_im2col_image_kernel(const U* __restrict__ img, T* __restrict__ out, int B, int Hi, int Wi,
                                                                int H, int W, float m0, float m1, float m2, float s0, float s1,
                                                                float s2) {
  __shared__ float patch[21 * (kStemPatchW + 3)];
  __shared__ __align__(16) short koff[192];
  constexpr int kPW = kStemPatchW + 3;
  const int Ho = H / 2, Wo = W / 2;
  const int strips = (Wo + kStemPx - 1) / kStemPx;
  const int strip = blockIdx.x % strips;
  const int ho = (blockIdx.x / strips) % Ho;
  const int b = blockIdx.x / (strips * Ho);
  const int wo0 = strip * kStemPx;
  const int x0 = 2 * wo0 - 3, y0 = 2 * ho - 3;
  if (threadIdx.x < 192) {
    const int k = threadIdx.x;
    const int c = k % 3, tap = k / 3;
    koff[k] = k < 147 ? static_cast<short>((c * 7 + tap / 7) * kPW + tap % 7) : static_cast<short>(-1);
  }
  // patch[(c_rgb * 7 + ky)][col]; the three BGR values of a pixel are adjacent in memory: one thread loads one pixel
  for (int i = threadIdx.x; i < 7 * kStemPatchW; i += blockDim.x) {
    const int col = i % kStemPatchW, ky = i / kStemPatchW;
    const int yy = y0 + ky, xx = x0 + col;
    float r = 0.f, g = 0.f, bl = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      float vb = 0.f, vg = 0.f, vr = 0.f;                 // pad_even_divided: raw zeros outside the camera image
      if (yy < Hi && xx < Wi) {
        const U* px = img + ((static_cast<long long>(b) * Hi + yy) * Wi + xx) * 3;
        vb = static_cast<float>(px[0]); vg = static_cast<float>(px[1]); vr = static_cast<float>(px[2]);
      }
      bl = (vb - m0) / s0; g = (vg - m1) / s1; r = (vr - m2) / s2;      // (img - mean) / std in BGR order, then BGR -> RGB
    }
    patch[(0 * 7 + ky) * kPW + col] = r;
    patch[(1 * 7 + ky) * kPW + col] = g;
    patch[(2 * 7 + ky) * kPW + col] = bl;
  }
  __syncthreads();
  const int npx = min(kStemPx, Wo - wo0);
  for (int i = threadIdx.x; i < npx * 24; i += blockDim.x) {
    const int kc = i % 24, px = i / 24;
    const uint4 ko = *reinterpret_cast<const uint4*>(koff + kc * 8);
    const short* o = reinterpret_cast<const short*>(&ko);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = o[j] >= 0 ? patch[o[j] + 2 * px] : 0.f;
    const long long m = (static_cast<long long>(b) * Ho + ho) * Wo + wo0 + px;
    store8(out + m * 192 + kc * 8, f);
  }
}

// ---------------------------------------------------------------- 3x3/s2/p1 max pool (models/backbone.py:104)
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, cv = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(Wo);
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(Ho));
    const int ho = static_cast<int>(tu_) - b * Ho;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int y = ho * 2 - 1 + dy;
      if (y < 0 || y >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = wo * 2 - 1 + dx;
        if (xx < 0 || xx >= W) continue;
        float f[8];
        load8(in + ((static_cast<long long>(b) * H + y) * W + xx) * C + c, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], f[j]);
      }
    }
    store8(out + m * C + c, best);
  }
}

// ---------------------------------------------------------------- 2x2 average (== bilinear x0.5, align_corners=False:
// models/fpn.py:54, planerecnet.py:115)
template <typename T>
__global__ void avgpool2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(Wo);
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(Ho));
    const int ho = static_cast<int>(tu_) - b * Ho;
    const T* p = in + ((static_cast<long long>(b) * H + 2 * ho) * W + 2 * wo) * C + c;
    float a[8], q[8], r[8], s[8], o[8];
    load8(p, a);
    load8(p + C, q);
    load8(p + static_cast<long long>(W) * C, r);
    load8(p + static_cast<long long>(W) * C + C, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.25f * ((a[j] + q[j]) + (r[j] + s[j]));
    store8(out + m * C + c, o);
  }
}

// ---------------------------------------------------------------- bilinear resize (+ coord channels)
// planerecnet.py:370-382: cat[feat, x, y] then F.interpolate(size=S, bilinear, align_corners=False).
// out [B,Ho,Wo,c_out] with c_out >= C (+2 coords); channels beyond are zero-filled.
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int* i0, int* i1, float* l) {
  float s = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;
  s = s < 0.f ? 0.f : s;
  int a = static_cast<int>(s);
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a < in_size - 1 ? a + 1 : a;
  *l = s - static_cast<float>(a);
}

template <typename T>
__global__ void resize_bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C,
                                       int Ho, int Wo, int c_out, int add_coord) {
  const int cv = c_out / 8;
  const float sh = static_cast<float>(H) / static_cast<float>(Ho), sw = static_cast<float>(W) / static_cast<float>(Wo);
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(Wo);
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(Ho));
    const int ho = static_cast<int>(tu_) - b * Ho;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(ho, sh, H, &y0, &y1, &ly);
    src_index(wo, sw, W, &x0, &x1, &lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    float o[8];
    if (c < C) {
      const T* base = in + static_cast<long long>(b) * H * W * C + c;
      float a[8], q[8], r[8], s[8];
      load8(base + (static_cast<long long>(y0) * W + x0) * C, a);
      load8(base + (static_cast<long long>(y0) * W + x1) * C, q);
      load8(base + (static_cast<long long>(y1) * W + x0) * C, r);
      load8(base + (static_cast<long long>(y1) * W + x1) * C, s);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * q[j] + w10 * r[j] + w11 * s[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      if (add_coord && c == C) {
        // linspace(-1, 1, n)[i] = -1 + 2 i / (n - 1)
        const float xs = W > 1 ? 2.f / static_cast<float>(W - 1) : 0.f;
        const float ys = H > 1 ? 2.f / static_cast<float>(H - 1) : 0.f;
        const float cx0 = -1.f + xs * x0, cx1 = -1.f + xs * x1, cy0 = -1.f + ys * y0, cy1 = -1.f + ys * y1;
        o[0] = (w00 + w10) * cx0 + (w01 + w11) * cx1;
        o[1] = (w00 + w01) * cy0 + (w10 + w11) * cy1;
      }
    }
    store8(out + m * c_out + c, o);
  }
}

// ---------------------------------------------------------------- coord channel append without resize
// planerecnet.py:483-490 (mask head level 3): out [B,H,W,c_out] = [in, x, y, 0...]
template <typename T>
__global__ void append_coord_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C,
                                    int c_out) {
  const int cv = c_out / 8;
  const long long total = static_cast<long long>(B) * H * W * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);      // 32-bit decode (see pw_grid)
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    float o[8];
    if (c < C) {
      load8(in + m * C + c, o);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      if (c == C) {
        const unsigned tq_ = mu_ / static_cast<unsigned>(W);
        const int x = static_cast<int>(mu_ - tq_ * static_cast<unsigned>(W)), y = static_cast<int>(tq_ % static_cast<unsigned>(H));
        o[0] = W > 1 ? -1.f + 2.f * x / static_cast<float>(W - 1) : -1.f;
        o[1] = H > 1 ? -1.f + 2.f * y / static_cast<float>(H - 1) : -1.f;
      }
    }
    store8(out + m * c_out + c, o);
  }
}

// ---------------------------------------------------------------- GroupNorm apply (+ReLU) from epilogue sums
// planerecnet.py:341-342, 420-421, 463-464: y = relu((x - mean) * rstd * gamma + beta), statistics over
// (C/G channels x H x W) per sample, eps 1e-5; sums come from the conv epilogue (fp32 accumulators).
template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ in, T* __restrict__ out, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int B, int HW, int C,
                                int cg, float eps, int relu) {
  const int cv = C / 8, G = C / cg;
  const float inv_cnt = 1.f / (static_cast<float>(HW) * cg);
  const long long total = static_cast<long long>(B) * HW * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);      // 32-bit decode (see pw_grid)
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const int b = static_cast<int>(mu_ / static_cast<unsigned>(HW));
    float f[8];
    load8(in + m * C + c, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cg;
      const float s1 = __ldg(stats + (static_cast<long long>(b) * G + g) * 2);
      const float s2 = __ldg(stats + (static_cast<long long>(b) * G + g) * 2 + 1);
      const float mean = s1 * inv_cnt;
      const float var = fmaxf(s2 * inv_cnt - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      float y = (f[j] - mean) * rstd * __ldg(gamma + c + j) + __ldg(beta + c + j);
      f[j] = relu ? fmaxf(y, 0.f) : y;
    }
    store8(out + m * C + c, f);
  }
}

// ---------------------------------------------------------------- bilinear x2 upsample (+ accumulate)
// planerecnet.py:439,453 (nn.Upsample(scale_factor=2, bilinear, align_corners=False)) and the level sum :493.
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C,
                                  int accumulate) {
  const int Ho = 2 * H, Wo = 2 * W, cv = C / 8;
  const long long total = static_cast<long long>(B) * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(Wo);
    const int wo = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(Wo)), b = static_cast<int>(tu_ / static_cast<unsigned>(Ho));
    const int ho = static_cast<int>(tu_) - b * Ho;
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(ho, 0.5f, H, &y0, &y1, &ly);
    src_index(wo, 0.5f, W, &x0, &x1, &lx);
    const T* base = in + static_cast<long long>(b) * H * W * C + c;
    float a[8], q[8], r[8], s[8], o[8];
    load8(base + (static_cast<long long>(y0) * W + x0) * C, a);
    load8(base + (static_cast<long long>(y0) * W + x1) * C, q);
    load8(base + (static_cast<long long>(y1) * W + x0) * C, r);
    load8(base + (static_cast<long long>(y1) * W + x1) * C, s);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * q[j] + w10 * r[j] + w11 * s[j];
    if (accumulate) {
      float prev[8];
      load8(out + m * C + c, prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += prev[j];
    }
    store8(out + m * C + c, o);
  }
}

// ---------------------------------------------------------------- elementwise product (planerecnet.py:600 x * attn)
template <typename T>
__global__ void mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    load8(a + i * 8, x);
    load8(b + i * 8, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= y[j];
    store8(out + i * 8, x);
  }
}

// ---------------------------------------------------------------- plane-prior attention: centre-pixel gather
// planerecnet.py:594 bilinear x0.25 (align_corners=False) of a map == mean of pixels (4i+1,4i+2)x(4j+1,4j+2);
// sigmoid and the 3728->256 conv are per-pixel, so only those pixels are needed.  Row order:
// (image, block row, block col, 2x2 position) so 4 consecutive rows form one output pixel.
template <typename T>
__global__ void ppa_gather_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H, int W, int C) {
  const int Hb = H / 4, Wb = W / 4, cv = C / 8;
  const long long total = static_cast<long long>(B) * Hb * Wb * 4 * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);      // 32-bit decode (see pw_grid)
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const int pos = static_cast<int>(m & 3);
    const long long blk = m >> 2;
    const int bx = static_cast<int>(blk % Wb), by = static_cast<int>((blk / Wb) % Hb);
    const int b = static_cast<int>(blk / (static_cast<long long>(Wb) * Hb));
    const int y = 4 * by + 1 + (pos >> 1), x = 4 * bx + 1 + (pos & 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + y) * W + x) * C + c));
    *reinterpret_cast<uint4*>(out + m * C + c) = v;
  }
}

// ---------------------------------------------------------------- layout conversions at the module boundary
// NHWC (16-bit or fp32, row pitch ld) -> NCHW fp32 contiguous, the layout callers .view() (losses.py:74,130).
template <typename T, bool kSrcF32>
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ src, float* __restrict__ dst, int B, int HW, int C,
                                    int ld, int src_img_rows) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const long long idx = (static_cast<long long>(b) * src_img_rows + p) * ld + c;
      if (kSrcF32) v = __ldg(static_cast<const float*>(src) + idx);
      else v = static_cast<float>(static_cast<const T*>(src)[idx]);
    }
    tile[r][threadIdx.x] = v;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[(static_cast<long long>(b) * C + c) * HW + p] = tile[threadIdx.x][r];
  }
}

// NCHW fp32 -> NHWC 16-bit with channel padding (used by the per-module entry points / tests)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int HW, int C, int c_pad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? __ldg(src + (static_cast<long long>(b) * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < c_pad) dst[(static_cast<long long>(b) * HW + p) * c_pad + c] = static_cast<T>(tile[threadIdx.x][r]);
  }
}

}  // namespace prn

// =================================================================================================== C ABI
using namespace prn;

extern "C" {

int prn_stem_im2col(const float* x_nchw, void* out16, int32_t batch, int32_t h, int32_t w, int32_t dtype, void* stream) {
  PRN_REQUIRE(x_nchw && out16 && batch > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "stem_im2col: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int strips = (w / 2 + kStemPx - 1) / kStemPx;
  const int grid = batch * (h / 2) * strips;
  PRN_DISPATCH(dtype,
               (stem_im2col_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x_nchw, static_cast<__nv_bfloat16*>(out16), batch, h, w)),
               (stem_im2col_kernel<__half><<<grid, 256, 0, st>>>(x_nchw, static_cast<__half*>(out16), batch, h, w)));
  PRN_LAUNCH_CHECK();
}

int prn_stem_im2col_image(const void* img_bgr_nhwc, int32_t is_u8, void* out16, int32_t batch, int32_t h_img, int32_t w_img,
                          int32_t h_pad, int32_t w_pad, const float* mean_bgr3, const float* std_bgr3, int32_t dtype, void* stream) {
  PRN_REQUIRE(img_bgr_nhwc && out16 && mean_bgr3 && std_bgr3 && batch > 0 && h_img > 0 && w_img > 0 && h_pad >= h_img &&
                  w_pad >= w_img && h_pad % 2 == 0 && w_pad % 2 == 0, "stem_im2col_image: bad arguments");
  PRN_REQUIRE(std_bgr3[0] != 0.f && std_bgr3[1] != 0.f && std_bgr3[2] != 0.f, "stem_im2col_image: zero std");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int strips = (w_pad / 2 + kStemPx - 1) / kStemPx;
  const int grid = batch * (h_pad / 2) * strips;
  const float m0 = mean_bgr3[0], m1 = mean_bgr3[1], m2 = mean_bgr3[2], s0 = std_bgr3[0], s1 = std_bgr3[1], s2 = std_bgr3[2];
#define PRN_STEM_IMG(T, U)                                                                                                  \
  stem_im2col_image_kernel<T, U><<<grid, 256, 0, st>>>(static_cast<const U*>(img_bgr_nhwc), static_cast<T*>(out16), batch, h_img, \
                                                       w_img, h_pad, w_pad, m0, m1, m2, s0, s1, s2)
  if (is_u8) {
    PRN_DISPATCH(dtype, (PRN_STEM_IMG(__nv_bfloat16, uint8_t)), (PRN_STEM_IMG(__half, uint8_t)));
  } else {
    PRN_DISPATCH(dtype, (PRN_STEM_IMG(__nv_bfloat16, float)), (PRN_STEM_IMG(__half, float)));
  }
#undef PRN_STEM_IMG
  PRN_LAUNCH_CHECK();
}

int prn_maxpool3x3s2(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && batch > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "maxpool: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * ((h - 1) / 2 + 1) * ((w - 1) / 2 + 1) * (c / 8);
  PRN_DISPATCH(dtype,
               (maxpool3s2_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c)),
               (maxpool3s2_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), batch, h, w, c)));
  PRN_LAUNCH_CHECK();
}

int prn_avgpool2x2(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && batch > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "avgpool2x2: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * (h / 2) * (w / 2) * (c / 8);
  PRN_DISPATCH(dtype,
               (avgpool2_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c)),
               (avgpool2_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), batch, h, w, c)));
  PRN_LAUNCH_CHECK();
}

int prn_resize_bilinear(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t h_out,
                        int32_t w_out, int32_t c_out, int32_t add_coord, int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && batch > 0 && h > 0 && w > 0 && h_out > 0 && w_out > 0 && c % 8 == 0 && c_out % 8 == 0 &&
                  c_out >= c + (add_coord ? 2 : 0), "resize_bilinear: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h_out * w_out * (c_out / 8);
  PRN_DISPATCH(dtype,
               (resize_bilinear_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c, h_out, w_out, c_out, add_coord)),
               (resize_bilinear_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), batch, h, w, c, h_out, w_out, c_out, add_coord)));
  PRN_LAUNCH_CHECK();
}

int prn_append_coord(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t c_out,
                     int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && batch > 0 && h > 0 && w > 0 && c % 8 == 0 && c_out % 8 == 0 && c_out >= c + 2, "append_coord: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h * w * (c_out / 8);
  PRN_DISPATCH(dtype,
               (append_coord_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c, c_out)),
               (append_coord_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), batch, h, w, c, c_out)));
  PRN_LAUNCH_CHECK();
}

int prn_groupnorm_apply(const void* in16, void* out16, const float* stats, const float* gamma, const float* beta,
                        int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps, int32_t relu,
                        int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && stats && gamma && beta && batch > 0 && hw > 0 && c % 8 == 0 && ch_per_group > 0 &&
                  c % ch_per_group == 0, "groupnorm_apply: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * hw * (c / 8);
  PRN_DISPATCH(dtype,
               (gn_apply_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), stats, gamma, beta, batch, hw, c, ch_per_group, eps, relu)),
               (gn_apply_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), stats, gamma, beta, batch, hw, c, ch_per_group, eps, relu)));
  PRN_LAUNCH_CHECK();
}

int prn_upsample2x_bilinear(const void* in16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c,
                            int32_t accumulate, int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && out16 && batch > 0 && h > 0 && w > 0 && c % 8 == 0, "upsample2x: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * 4 * h * w * (c / 8);
  PRN_DISPATCH(dtype,
               (upsample2x_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(in16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c, accumulate)),
               (upsample2x_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(in16), static_cast<__half*>(out16), batch, h, w, c, accumulate)));
  PRN_LAUNCH_CHECK();
}

int prn_mul(const void* a16, const void* b16, void* out16, int64_t n, int32_t dtype, void* stream) {
  PRN_REQUIRE(a16 && b16 && out16 && n > 0 && n % 8 == 0, "mul: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (mul_kernel<__nv_bfloat16><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(a16), static_cast<const __nv_bfloat16*>(b16), static_cast<__nv_bfloat16*>(out16), n / 8)),
               (mul_kernel<__half><<<pw_grid(n / 8), kPwThreads, 0, st>>>(static_cast<const __half*>(a16), static_cast<const __half*>(b16), static_cast<__half*>(out16), n / 8)));
  PRN_LAUNCH_CHECK();
}

int prn_ppa_gather(const void* mask16, void* out16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype, void* stream) {
  PRN_REQUIRE(mask16 && out16 && batch > 0 && h % 4 == 0 && w % 4 == 0 && c % 8 == 0, "ppa_gather: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * (h / 4) * (w / 4) * 4 * (c / 8);
  PRN_DISPATCH(dtype,
               (ppa_gather_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(mask16), static_cast<__nv_bfloat16*>(out16), batch, h, w, c)),
               (ppa_gather_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(mask16), static_cast<__half*>(out16), batch, h, w, c)));
  PRN_LAUNCH_CHECK();
}

int prn_nhwc_to_nchw_f32(const void* src, int32_t src_is_f32, float* dst, int32_t batch, int32_t hw, int32_t c,
                         int32_t ld, int32_t src_img_rows, int32_t dtype, void* stream) {
  PRN_REQUIRE(src && dst && batch > 0 && hw > 0 && c > 0 && ld >= c, "nhwc_to_nchw: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_img_rows == 0) src_img_rows = hw;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch), block(32, 8);
  if (src_is_f32) nhwc_to_nchw_kernel<__half, true><<<grid, block, 0, st>>>(src, dst, batch, hw, c, ld, src_img_rows);
  else if (dtype == PRN_BF16) nhwc_to_nchw_kernel<__nv_bfloat16, false><<<grid, block, 0, st>>>(src, dst, batch, hw, c, ld, src_img_rows);
  else nhwc_to_nchw_kernel<__half, false><<<grid, block, 0, st>>>(src, dst, batch, hw, c, ld, src_img_rows);
  PRN_LAUNCH_CHECK();
}

int prn_nchw_f32_to_nhwc(const float* src, void* dst16, int32_t batch, int32_t hw, int32_t c, int32_t c_pad,
                         int32_t dtype, void* stream) {
  PRN_REQUIRE(src && dst16 && batch > 0 && hw > 0 && c > 0 && c_pad >= c, "nchw_to_nhwc: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((hw + 31) / 32, (c_pad + 31) / 32, batch), block(32, 8);
  PRN_DISPATCH(dtype,
               (nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst16), batch, hw, c, c_pad)),
               (nchw_to_nhwc_kernel<__half><<<grid, block, 0, st>>>(src, static_cast<__half*>(dst16), batch, hw, c, c_pad)));
  PRN_LAUNCH_CHECK();
}

}  // extern "C"

// ================================================================================================
// 3x3 reflect-padded conv to ONE output channel (+bias, softplus): the depth head (planerecnet.py:570-573).
// N = 1 wastes a tensor-core tile (and re-gathers the 64-channel input 9x for one column), so this layer
// runs on the CUDA cores: a CTA stages a 16x16 output tile's 18x18x C halo in shared memory (pixel
// stride padded by 16 B against bank conflicts) and every thread accumulates one pixel in fp32.
namespace prn {

constexpr int kD1Tile = 16;

template <typename T>
__global__ void __launch_bounds__(kD1Tile * kD1Tile)
conv3x3_to1_kernel(const T* __restrict__ in, const float* __restrict__ wgt /*[9][C] fp32*/, float bias_host, const float* __restrict__ bias_dev,
                   float* __restrict__ out, int B, int H, int W, int C, int softplus) {
  extern __shared__ __align__(16) uint8_t sm[];
  const int pstride = C * 2 + 16;                          // bytes per staged pixel
  float* wsm = reinterpret_cast<float*>(sm);               // [9*C]
  uint8_t* tile = sm + 9 * C * 4;                          // [(18*18)][pstride]
  const int tiles_x = (W + kD1Tile - 1) / kD1Tile, tiles_y = (H + kD1Tile - 1) / kD1Tile;
  const int tx0 = (blockIdx.x % tiles_x) * kD1Tile;
  const int ty0 = ((blockIdx.x / tiles_x) % tiles_y) * kD1Tile;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) wsm[i] = __ldg(wgt + i);
  const int cv = C / 8, halo = kD1Tile + 2;
  for (int i = threadIdx.x; i < halo * halo * cv; i += blockDim.x) {
    const int c8 = i % cv, pix = i / cv;
    int y = ty0 - 1 + pix / halo, x = tx0 - 1 + pix % halo;
    y = y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y);         // nn.ReflectionPad2d(1)
    x = x < 0 ? -x : (x >= W ? 2 * W - 2 - x : x);
    y = min(max(y, 0), H - 1);                             // tiles hanging over the border: any valid address
    x = min(max(x, 0), W - 1);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + y) * W + x) * C) + c8);
    *reinterpret_cast<uint4*>(tile + pix * pstride + c8 * 16) = v;
  }
  __syncthreads();
  const int ly = threadIdx.x / kD1Tile, lx = threadIdx.x % kD1Tile;
  const int oy = ty0 + ly, ox = tx0 + lx;
  float acc = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint8_t* px = tile + ((ly + tap / 3) * halo + lx + tap % 3) * pstride;
    const float* wt = wsm + tap * C;
    for (int c8 = 0; c8 < cv; ++c8) {
      const uint4 v = *reinterpret_cast<const uint4*>(px + c8 * 16);
      const float4 w0 = *reinterpret_cast<const float4*>(wt + c8 * 8), w1 = *reinterpret_cast<const float4*>(wt + c8 * 8 + 4);
      const float2 f0 = Pack2<T>::unpack(v.x), f1 = Pack2<T>::unpack(v.y), f2 = Pack2<T>::unpack(v.z), f3 = Pack2<T>::unpack(v.w);
      acc = fmaf(f0.x, w0.x, acc); acc = fmaf(f0.y, w0.y, acc); acc = fmaf(f1.x, w0.z, acc); acc = fmaf(f1.y, w0.w, acc);
      acc = fmaf(f2.x, w1.x, acc); acc = fmaf(f2.y, w1.y, acc); acc = fmaf(f3.x, w1.z, acc); acc = fmaf(f3.y, w1.w, acc);
    }
  }
  if (oy < H && ox < W) {
    float v = acc + (bias_dev != nullptr ? __ldg(bias_dev) : bias_host);
    if (softplus) v = v > 20.f ? v : log1pf(expf(v));
    out[(static_cast<long long>(b) * H + oy) * W + ox] = v;
  }
}

}  // namespace prn

static int conv3x3_to1_launch(const void* in16, const float* weight9c, float bias, const float* bias_dev, float* out, int32_t batch,
                              int32_t h, int32_t w, int32_t c, int32_t softplus, int32_t dtype, void* stream) {
  PRN_REQUIRE(in16 && weight9c && out && batch > 0 && h > 1 && w > 1 && c > 0 && c % 8 == 0 && c <= 128,
              "conv3x3_to1_reflect: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = batch * ((h + kD1Tile - 1) / kD1Tile) * ((w + kD1Tile - 1) / kD1Tile);
  const size_t smem = static_cast<size_t>(9) * c * 4 + static_cast<size_t>(kD1Tile + 2) * (kD1Tile + 2) * (c * 2 + 16);
  if (dtype == PRN_BF16) {
    static bool cfg = false;
    if (!cfg) { PRN_CUDA(cudaFuncSetAttribute(conv3x3_to1_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); cfg = true; }
    conv3x3_to1_kernel<__nv_bfloat16><<<grid, kD1Tile * kD1Tile, smem, st>>>(static_cast<const __nv_bfloat16*>(in16), weight9c, bias, bias_dev, out, batch, h, w, c, softplus);
  } else if (dtype == PRN_F16) {
    static bool cfg = false;
    if (!cfg) { PRN_CUDA(cudaFuncSetAttribute(conv3x3_to1_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); cfg = true; }
    conv3x3_to1_kernel<__half><<<grid, kD1Tile * kD1Tile, smem, st>>>(static_cast<const __half*>(in16), weight9c, bias, bias_dev, out, batch, h, w, c, softplus);
  } else {
    return set_error(PRN_ERR_INVALID, "conv3x3_to1_reflect: bad dtype");
  }
  PRN_LAUNCH_CHECK();
}

extern "C" int prn_conv3x3_to1_reflect(const void* in16, const float* weight9c, float bias, float* out, int32_t batch,
                                       int32_t h, int32_t w, int32_t c, int32_t softplus, int32_t dtype, void* stream) {
  return conv3x3_to1_launch(in16, weight9c, bias, nullptr, out, batch, h, w, c, softplus, dtype, stream);
}

extern "C" int prn_conv3x3_to1_reflect_devbias(const void* in16, const float* weight9c, const float* bias_dev, float* out,
                                               int32_t batch, int32_t h, int32_t w, int32_t c, int32_t softplus, int32_t dtype,
                                               void* stream) {
  if (!bias_dev) return set_error(PRN_ERR_INVALID, "conv3x3_to1_reflect_devbias: NULL bias");
  return conv3x3_to1_launch(in16, weight9c, 0.f, bias_dev, out, batch, h, w, c, softplus, dtype, stream);
}
