// prn_train_heads.cu — backward passes of the head / decoder glue around the contractions: GroupNorm(+ReLU),
// 2x2 mean, bilinear x2, arbitrary bilinear resize, reflection padding + nearest x2 (folded after the input-gradient
// contraction), softplus.  NHWC 16-bit, 8 channels per thread; gather formulations (deterministic) wherever the
// stencil is fixed, fp32 vector reductions only for the arbitrary-ratio resize.  Autograd formulas of the
// operator call sites cited per function in include/prn_b200.h.
#include "prn_pw.cuh"

namespace prn {

__device__ __forceinline__ void bil_index(int dst, float scale, int in_size, int* i0, int* i1, float* l) {
  float s = (static_cast<float>(dst) + 0.5f) * scale - 0.5f;    // same arithmetic as the forward resampler
  s = s < 0.f ? 0.f : s;
  int a = static_cast<int>(s);
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a < in_size - 1 ? a + 1 : a;
  *l = s - static_cast<float>(a);
}

// ---------------------------------------------------------------- GroupNorm(32, C) + ReLU backward
// stats[(b*G + g)*2 + {0,1}] = {sum, sumsq} of x over (pixels, channels of the group), as left by the conv epilogue.
// pass 1: per (image, channel): a = sum_pix g, bq = sum_pix g * xhat with g = dz * (out > 0);
//         sums_bc[(b*C + c)*2 + {0,1}] += {a, bq};  dgb[c*2 + {0,1}] += {a, bq} (= dbeta, dgamma; shared-weight levels
//         keep accumulating into the same dgb).
template <typename T>
__global__ void gn_bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ out, const T* __restrict__ x,
                                     const float* __restrict__ stats, float* __restrict__ sums_bc, float* __restrict__ dgb,
                                     int HW, int C, int cg, float eps) {
  const int cv = C / 8, G = C / cg;
  const int b = blockIdx.y;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long tthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  const long long rstep = static_cast<unsigned>(tthreads) / static_cast<unsigned>(cv);
  const float inv_cnt = 1.f / (static_cast<float>(HW) * cg);
  float mean[8], rstd[8], s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cg;
    const float m1 = __ldg(stats + (static_cast<long long>(b) * G + g) * 2) * inv_cnt;
    const float m2 = __ldg(stats + (static_cast<long long>(b) * G + g) * 2 + 1) * inv_cnt;
    mean[j] = m1;
    rstd[j] = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + eps);
    s1[j] = 0.f;
    s2[j] = 0.f;
  }
  const long long img = static_cast<long long>(b) * HW;
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < HW; m += rstep) {
    float g[8], o[8], xv[8];
    load8(dz + (img + m) * C + c, g);
    load8(out + (img + m) * C + c, o);
    load8(x + (img + m) * C + c, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = o[j] > 0.f ? g[j] : 0.f;
      s1[j] += gg;
      s2[j] = fmaf(gg, (xv[j] - mean[j]) * rstd[j], s2[j]);
    }
  }
  if (fold_ok(cv)) {
    __shared__ float part[kFoldFloats];
    block_fold_chan(s1, s2, cv, part, [&](int o, float tot) {
      atomicAdd(sums_bc + static_cast<long long>(b) * C * 2 + o, tot);
      atomicAdd(dgb + o, tot);
    });
    return;
  }
  extern __shared__ float acc[];   // [C][2]: general channel counts, shared-memory atomics
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(acc + 2 * (c + j), s1[j]);
    atomicAdd(acc + 2 * (c + j) + 1, s2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    atomicAdd(sums_bc + static_cast<long long>(b) * C * 2 + i, acc[i]);
    atomicAdd(dgb + i, acc[i]);
  }
}

// pass 2: dx = rstd * (g*gamma - mean_grp(g*gamma) - xhat * mean_grp(g*gamma*xhat)), group means from sums_bc.
// Channel-stationary threads per image (blockIdx.y): dx = a*g + b*x + k with per-(image, channel) constants in registers.
template <typename T>
__global__ void __launch_bounds__(kPwThreads) gn_bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ out,
                                                                const T* __restrict__ x, const float* __restrict__ stats,
                                                                const float* __restrict__ gamma, const float* __restrict__ sums_bc,
                                                                T* __restrict__ dx, int B, int HW, int C, int cg, float eps) {
  const int cv = C / 8, G = C / cg;
  const int b = blockIdx.y;
  const float inv_cnt = 1.f / (static_cast<float>(HW) * cg);
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rstep = (gridDim.x * blockDim.x) / static_cast<unsigned>(cv);      // 32-bit: grids stay far below 2^31 threads
  const int c = static_cast<int>(static_cast<unsigned>(gtid) % static_cast<unsigned>(cv)) * 8;
  // per-group sums of gamma * {a, bq}: one warp per group, lanes over the group's channels (a per-thread loop over the
  // group's channels was a 2 * cg-deep chain of L2 round trips in every thread: ~30 us of a 38 us launch)
  __shared__ float gsum[2 * 512];
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int grp = warp; grp < G; grp += nwarps) {
      float ga = 0.f, gb = 0.f;
      for (int cc = grp * cg + lane; cc < (grp + 1) * cg; cc += 32) {
        const float gm = __ldg(gamma + cc);
        ga = fmaf(gm, __ldg(sums_bc + (static_cast<long long>(b) * C + cc) * 2), ga);
        gb = fmaf(gm, __ldg(sums_bc + (static_cast<long long>(b) * C + cc) * 2 + 1), gb);
      }
      for (int off = 16; off; off >>= 1) {
        ga += __shfl_xor_sync(0xffffffffu, ga, off);
        gb += __shfl_xor_sync(0xffffffffu, gb, off);
      }
      if (lane == 0) {
        gsum[2 * grp] = ga;
        gsum[2 * grp + 1] = gb;
      }
    }
  }
  __syncthreads();
  float ka[8], kb[8], kk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int grp = (c + j) / cg;
    const float m1 = __ldg(stats + (static_cast<long long>(b) * G + grp) * 2) * inv_cnt;
    const float m2 = __ldg(stats + (static_cast<long long>(b) * G + grp) * 2 + 1) * inv_cnt;
    const float rstd = rsqrtf(fmaxf(m2 - m1 * m1, 0.f) + eps);
    const float ga = gsum[2 * grp], gb = gsum[2 * grp + 1];      // sum over the group's channels of gamma * {a, bq}
    ka[j] = rstd * __ldg(gamma + c + j);
    kb[j] = -rstd * rstd * gb * inv_cnt;
    kk[j] = -rstd * ga * inv_cnt - kb[j] * m1;
  }
  const long long img = static_cast<long long>(b) * HW;
  for (long long m = static_cast<unsigned>(gtid) / static_cast<unsigned>(cv); m < HW; m += rstep) {
    float g[8], o[8], xv[8];
    load8(dz + (img + m) * C + c, g);
    load8(out + (img + m) * C + c, o);
    load8(x + (img + m) * C + c, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) xv[j] = fmaf(ka[j], o[j] > 0.f ? g[j] : 0.f, fmaf(kb[j], xv[j], kk[j]));
    store8(dx + (img + m) * C + c, xv);
  }
}

// ---------------------------------------------------------------- 2x2 mean backward (models/fpn.py:54, planerecnet.py:115)
template <typename T>
__global__ void avgpool2_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, int B, int H, int W, int C, int accumulate) {
  const int cv = C / 8, Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * H * W * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(W);
    const int x = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(W)), b = static_cast<int>(tu_ / static_cast<unsigned>(H));
    const int y = static_cast<int>(tu_) - b * H;
    float g[8];
    load8(dout + ((static_cast<long long>(b) * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c, g);
    if (accumulate) {
      float p[8];
      load8(din + m * C + c, p);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(0.25f, g[j], p[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] *= 0.25f;
    }
    store8(din + m * C + c, g);
  }
}

// ---------------------------------------------------------------- bilinear x2 backward (planerecnet.py:439,453,493), gather form
template <typename T>
__global__ void upsample2x_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din, int B, int H, int W, int C) {
  const int cv = C / 8, Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(B) * H * W * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(W);
    const int x = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(W)), b = static_cast<int>(tu_ / static_cast<unsigned>(H));
    const int y = static_cast<int>(tu_) - b * H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ho = max(2 * y - 2, 0); ho <= min(2 * y + 2, Ho - 1); ++ho) {
      int y0, y1;
      float ly;
      bil_index(ho, 0.5f, H, &y0, &y1, &ly);
      const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int wo = max(2 * x - 2, 0); wo <= min(2 * x + 2, Wo - 1); ++wo) {
        int x0, x1;
        float lx;
        bil_index(wo, 0.5f, W, &x0, &x1, &lx);
        const float wgt = wy * ((x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f));
        if (wgt == 0.f) continue;
        float g[8];
        load8(dout + ((static_cast<long long>(b) * Ho + ho) * Wo + wo) * C + c, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, g[j], acc[j]);
      }
    }
    store8(din + m * C + c, acc);
  }
}

// ---------------------------------------------------------------- arbitrary bilinear resize backward (planerecnet.py:381)
// dout [B,Ho,Wo,ld] 16-bit (first C channels are feature channels; the coord channels carry no parameter gradient)
// -> din32 fp32 [B,H,W,C] += scattered (caller zeroes)
template <typename T>
__global__ void resize_bilinear_bwd_kernel(const T* __restrict__ dout, float* __restrict__ din32, int B, int H, int W, int C,
                                           int Ho, int Wo, int ld) {
  const int cv = C / 8;
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
    bil_index(ho, sh, H, &y0, &y1, &ly);
    bil_index(wo, sw, W, &x0, &x1, &lx);
    float g[8];
    load8(dout + m * ld + c, g);
    float* base = din32 + static_cast<long long>(b) * H * W * C + c;
    const float wts[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float* p = base + (static_cast<long long>(ys[k]) * W + xs[k]) * C;
      red_add_v4_f32(p, wts[k] * g[0], wts[k] * g[1], wts[k] * g[2], wts[k] * g[3]);
      red_add_v4_f32(p + 4, wts[k] * g[4], wts[k] * g[5], wts[k] * g[6], wts[k] * g[7]);
    }
  }
}

// ---------------------------------------------------------------- ReflectionPad2d(1) [+ nearest x2] backward
// dpad [B, He+2, We+2, ld] = gradient w.r.t. the padded (and upsampled) tensor, He = h*up, We = w*up (the "full"
// correlation of dY with the flipped weights).  din[b,y,x] = sum over effective rows ye in {up*y .. up*y+up-1} of
// dpad rows {ye+1} + {0 if ye == 1} + {He+1 if ye == He-2}; same along x.  Gather form.
template <typename T, int kUp>
__global__ void reflect_fold_kernel(const T* __restrict__ dpad, T* __restrict__ din, int B, int h, int w, int C, int ld,
                                    int accumulate) {
  const int cv = C / 8, He = h * kUp, We = w * kUp, Hp = He + 2, Wp = We + 2;
  const long long total = static_cast<long long>(B) * h * w * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // 32-bit index decode (pw_grid refuses >= 2^31 work items): the 64-bit divisions cost more than the pass's arithmetic
    const unsigned iu_ = static_cast<unsigned>(i), mu_ = iu_ / static_cast<unsigned>(cv);
    const int c = static_cast<int>(iu_ - mu_ * static_cast<unsigned>(cv)) * 8;
    const long long m = mu_;
    const unsigned tu_ = mu_ / static_cast<unsigned>(w);
    const int x = static_cast<int>(mu_ - tu_ * static_cast<unsigned>(w)), b = static_cast<int>(tu_ / static_cast<unsigned>(h));
    const int y = static_cast<int>(tu_) - b * h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const T* img = dpad + static_cast<long long>(b) * Hp * Wp * ld + c;
    // the kUp x kUp interior taps first (unconditional: their loads are all in flight together), then the rare border folds;
    // no index arrays (dynamically indexed ones live in local memory)
    uint4 q[kUp][kUp];
#pragma unroll
    for (int dy = 0; dy < kUp; ++dy)
#pragma unroll
      for (int dx = 0; dx < kUp; ++dx)
        q[dy][dx] = __ldg(reinterpret_cast<const uint4*>(img + (static_cast<long long>(y * kUp + dy + 1) * Wp + (x * kUp + dx + 1)) * ld));
    uint4 prev = make_uint4(0, 0, 0, 0);
    if (accumulate) prev = __ldg(reinterpret_cast<const uint4*>(din + m * C + c));
#pragma unroll
    for (int dy = 0; dy < kUp; ++dy)
#pragma unroll
      for (int dx = 0; dx < kUp; ++dx) {
        float g[8];
        unpack8<T>(q[dy][dx], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      }
    if (accumulate) {
      float g[8];
      unpack8<T>(prev, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += g[j];
    }
    const int ye0 = y * kUp, xe0 = x * kUp;
    const bool row_edge = ye0 <= 1 || ye0 + kUp - 1 >= He - 2, col_edge = xe0 <= 1 || xe0 + kUp - 1 >= We - 2;
    if (row_edge || col_edge) {
      auto add = [&](int r, int col) {
        float g[8];
        load8(img + (static_cast<long long>(r) * Wp + col) * ld, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += g[j];
      };
      for (int dy = 0; dy < kUp; ++dy) {
        const int ye = ye0 + dy;
        const int er0 = ye == 1 ? 0 : -1, er1 = ye == He - 2 ? He + 1 : -1;      // extra (reflected) rows of this row
        for (int dx = 0; dx < kUp; ++dx) {
          const int xe = xe0 + dx;
          const int ec0 = xe == 1 ? 0 : -1, ec1 = xe == We - 2 ? We + 1 : -1;
          // (row set) x (column set) minus the interior tap already taken
          if (ec0 >= 0) add(ye + 1, ec0);
          if (ec1 >= 0) add(ye + 1, ec1);
          if (er0 >= 0) { add(er0, xe + 1); if (ec0 >= 0) add(er0, ec0); if (ec1 >= 0) add(er0, ec1); }
          if (er1 >= 0) { add(er1, xe + 1); if (ec0 >= 0) add(er1, ec0); if (ec1 >= 0) add(er1, ec1); }
        }
      }
    }
    store8(din + m * C + c, acc);
  }
}

// ---------------------------------------------------------------- Softplus backward (planerecnet.py:572), single channel
// dpre16[m, 0] = dout[m] * (1 - exp(-out[m])) (= sigmoid(pre) for out = softplus(pre)); columns 1..63 = 0, so that the
// row is a 64-channel operand of the input / weight gradient contractions of the 64 -> 1 depth head.
template <typename T>
__global__ void softplus_bwd_pad_kernel(const float* __restrict__ dout, const float* __restrict__ out, T* __restrict__ dpre16,
                                        long long rows) {
  const long long total = rows * 8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i >> 3;
    const int ch = static_cast<int>(i & 7);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ch == 0) f[0] = __ldg(dout + m) * (1.f - __expf(-__ldg(out + m)));
    store8(dpre16 + m * 64 + ch * 8, f);
  }
}

}  // namespace prn

using namespace prn;

extern "C" {

int prn_gn_bwd_reduce(const void* dz16, const void* out16, const void* x16, const float* stats, float* sums_bc, float* dgb,
                      int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps, int32_t dtype, void* stream) {
  PRN_REQUIRE(dz16 && out16 && x16 && stats && sums_bc && dgb && batch > 0 && hw > 0 && c > 0 && c % 8 == 0 && c <= 4096 &&
                  ch_per_group > 0 && c % ch_per_group == 0, "gn_bwd_reduce: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cv = c / 8;
  int a = cv, b = kPwThreads;
  while (b) { const int t = a % b; a = b; b = t; }
  const int g0 = cv / a;
  long long want = (static_cast<long long>(hw) * cv + kPwThreads * 4LL - 1) / (kPwThreads * 4LL);
  const long long cap = static_cast<long long>(sm_count()) * 8 / batch + 1;
  if (want > cap) want = cap;
  long long gx = want / g0 * g0;
  if (gx < g0) gx = g0;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(batch));
  const size_t smem = static_cast<size_t>(c) * 2 * sizeof(float);
  PRN_DISPATCH(dtype,
               (gn_bwd_reduce_kernel<__nv_bfloat16><<<grid, kPwThreads, smem, st>>>(static_cast<const __nv_bfloat16*>(dz16), static_cast<const __nv_bfloat16*>(out16), static_cast<const __nv_bfloat16*>(x16), stats, sums_bc, dgb, hw, c, ch_per_group, eps)),
               (gn_bwd_reduce_kernel<__half><<<grid, kPwThreads, smem, st>>>(static_cast<const __half*>(dz16), static_cast<const __half*>(out16), static_cast<const __half*>(x16), stats, sums_bc, dgb, hw, c, ch_per_group, eps)));
  PRN_LAUNCH_CHECK();
}

int prn_gn_bwd_apply(const void* dz16, const void* out16, const void* x16, const float* stats, const float* gamma,
                     const float* sums_bc, void* dx16, int32_t batch, int32_t hw, int32_t c, int32_t ch_per_group, float eps,
                     int32_t dtype, void* stream) {
  PRN_REQUIRE(dz16 && out16 && x16 && stats && gamma && sums_bc && dx16 && batch > 0 && hw > 0 && c > 0 && c % 8 == 0 &&
                  ch_per_group > 0 && c % ch_per_group == 0 && c / ch_per_group <= 512, "gn_bwd_apply: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cv = c / 8;
  int a = cv, b = kPwThreads;
  while (b) { const int t = a % b; a = b; b = t; }
  const int g0 = cv / a;
  long long want = (static_cast<long long>(hw) * cv + kPwThreads * 4LL - 1) / (kPwThreads * 4LL);
  const long long cap = static_cast<long long>(sm_count()) * 16 / batch + 1;
  if (want > cap) want = cap;
  long long gx = want / g0 * g0;
  if (gx < g0) gx = g0;
  const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(batch));
  PRN_DISPATCH(dtype,
               (gn_bwd_apply_kernel<__nv_bfloat16><<<grid, kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dz16), static_cast<const __nv_bfloat16*>(out16), static_cast<const __nv_bfloat16*>(x16), stats, gamma, sums_bc, static_cast<__nv_bfloat16*>(dx16), batch, hw, c, ch_per_group, eps)),
               (gn_bwd_apply_kernel<__half><<<grid, kPwThreads, 0, st>>>(static_cast<const __half*>(dz16), static_cast<const __half*>(out16), static_cast<const __half*>(x16), stats, gamma, sums_bc, static_cast<__half*>(dx16), batch, hw, c, ch_per_group, eps)));
  PRN_LAUNCH_CHECK();
}

int prn_avgpool2x2_bwd(const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t accumulate,
                       int32_t dtype, void* stream) {
  PRN_REQUIRE(dout16 && din16 && batch > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "avgpool2x2_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h * w * (c / 8);
  PRN_DISPATCH(dtype,
               (avgpool2_bwd_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dout16), static_cast<__nv_bfloat16*>(din16), batch, h, w, c, accumulate)),
               (avgpool2_bwd_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(dout16), static_cast<__half*>(din16), batch, h, w, c, accumulate)));
  PRN_LAUNCH_CHECK();
}

int prn_upsample2x_bilinear_bwd(const void* dout16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t dtype,
                                void* stream) {
  PRN_REQUIRE(dout16 && din16 && batch > 0 && h > 0 && w > 0 && c % 8 == 0, "upsample2x_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h * w * (c / 8);
  PRN_DISPATCH(dtype,
               (upsample2x_bwd_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dout16), static_cast<__nv_bfloat16*>(din16), batch, h, w, c)),
               (upsample2x_bwd_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(dout16), static_cast<__half*>(din16), batch, h, w, c)));
  PRN_LAUNCH_CHECK();
}

int prn_resize_bilinear_bwd(const void* dout16, float* din32, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t h_out,
                            int32_t w_out, int32_t ld_dout, int32_t dtype, void* stream) {
  PRN_REQUIRE(dout16 && din32 && batch > 0 && h > 0 && w > 0 && h_out > 0 && w_out > 0 && c % 8 == 0 && ld_dout >= c &&
                  ld_dout % 8 == 0, "resize_bilinear_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h_out * w_out * (c / 8);
  PRN_DISPATCH(dtype,
               (resize_bilinear_bwd_kernel<__nv_bfloat16><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(dout16), din32, batch, h, w, c, h_out, w_out, ld_dout)),
               (resize_bilinear_bwd_kernel<__half><<<pw_grid(work), kPwThreads, 0, st>>>(static_cast<const __half*>(dout16), din32, batch, h, w, c, h_out, w_out, ld_dout)));
  PRN_LAUNCH_CHECK();
}

int prn_reflect_fold(const void* dpad16, void* din16, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t ld_dpad,
                     int32_t upsample, int32_t accumulate, int32_t dtype, void* stream) {
  PRN_REQUIRE(dpad16 && din16 && batch > 0 && h > 0 && w > 0 && c % 8 == 0 && ld_dpad >= c && ld_dpad % 8 == 0 &&
                  (upsample == 1 || upsample == 2) && h * upsample >= 3 && w * upsample >= 3, "reflect_fold: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = static_cast<long long>(batch) * h * w * (c / 8);
  const int grid = pw_grid(work);
#define PRN_FOLD(T_, UP_) reflect_fold_kernel<T_, UP_><<<grid, kPwThreads, 0, st>>>(static_cast<const T_*>(dpad16), static_cast<T_*>(din16), batch, h, w, c, ld_dpad, accumulate)
  if (upsample == 2) PRN_DISPATCH(dtype, (PRN_FOLD(__nv_bfloat16, 2)), (PRN_FOLD(__half, 2)));
  else PRN_DISPATCH(dtype, (PRN_FOLD(__nv_bfloat16, 1)), (PRN_FOLD(__half, 1)));
#undef PRN_FOLD
  PRN_LAUNCH_CHECK();
}

int prn_softplus_bwd_pad(const float* dout, const float* out, void* dpre16, int64_t rows, int32_t dtype, void* stream) {
  PRN_REQUIRE(dout && out && dpre16 && rows > 0, "softplus_bwd_pad: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PRN_DISPATCH(dtype,
               (softplus_bwd_pad_kernel<__nv_bfloat16><<<pw_grid(rows * 8), kPwThreads, 0, st>>>(dout, out, static_cast<__nv_bfloat16*>(dpre16), rows)),
               (softplus_bwd_pad_kernel<__half><<<pw_grid(rows * 8), kPwThreads, 0, st>>>(dout, out, static_cast<__half*>(dpre16), rows)));
  PRN_LAUNCH_CHECK();
}

}  // extern "C"
