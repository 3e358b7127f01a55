// prn_post.cu — inference bookkeeping kernels (planerecnet.py:106-107, 182-289; nms.py:8-12).
// Integer / comparison work over small tensors: HBM- and latency-bound, one pass each, no host syncs.
#include "prn_internal.h"
#include "prn_ptx.cuh"

namespace prn {

// ---------------------------------------------------------------- sigmoid + point-NMS (nms.py:8-12)
// logits: fp32 [B, total, ld] (first nc columns valid), rows ordered level-major then (y, x) — the layout the
// instance-head epilogue writes.  scores[b, row, c] = s if s == max(s over the 2x2 window ending at (y,x)) else 0,
// s = sigmoid(logit): max_pool2d(kernel 2, stride 1, padding 1)[:-1,:-1] looks up / left.
__global__ void point_nms_kernel(const float* __restrict__ logits, float* __restrict__ scores, int B, int total, int ld,
                                 int nc, int n_levels, const int* __restrict__ grids) {
  const long long n = static_cast<long long>(B) * total * nc;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nc);
    const int row = static_cast<int>((i / nc) % total);
    const int b = static_cast<int>(i / (static_cast<long long>(nc) * total));
    int off = 0, S = 0, lvl = 0;
    for (; lvl < n_levels; ++lvl) {
      S = grids[lvl];
      if (row < off + S * S) break;
      off += S * S;
    }
    const int y = (row - off) / S, xx = (row - off) % S;
    const float* base = logits + (static_cast<long long>(b) * total + off) * ld + c;
    auto sg = [&](int yy, int xq) { return 1.f / (1.f + expf(-__ldg(base + static_cast<long long>(yy * S + xq) * ld))); };
    const float v = sg(y, xx);
    float m = v;
    if (y > 0) m = fmaxf(m, sg(y - 1, xx));
    if (xx > 0) m = fmaxf(m, sg(y, xx - 1));
    if (y > 0 && xx > 0) m = fmaxf(m, sg(y - 1, xx - 1));
    scores[i] = (m == v) ? v : 0.f;
  }
}

// ---------------------------------------------------------------- per-candidate mask statistics
// seg fp32 [rows][P] (sigmoid mask probabilities).  Per row: area = #(seg > thr), ssum = sum of seg where > thr
// (planerecnet.py:216-232), and the binary mask as 16-bit 0/1 [rows][P] for the matrix-NMS Gram contraction.
template <typename T>
__global__ void mask_stats_kernel(const float* __restrict__ seg, T* __restrict__ mask16, float* __restrict__ area,
                                  float* __restrict__ ssum, int P, float thr) {
  const int row = blockIdx.x;
  const float* s = seg + static_cast<long long>(row) * P;
  T* mo = mask16 + static_cast<long long>(row) * P;
  float cnt = 0.f, acc = 0.f;
  for (int i = threadIdx.x * 4; i < P; i += blockDim.x * 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(s + i));
    const float f[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool on = f[j] > thr;
      o[j] = on ? 1.f : 0.f;
      cnt += o[j];
      acc += on ? f[j] : 0.f;
    }
    uint2 pk;
    pk.x = Pack2<T>::pack(o[0], o[1]);
    pk.y = Pack2<T>::pack(o[2], o[3]);
    *reinterpret_cast<uint2*>(mo + i) = pk;
  }
  __shared__ float sc[32], sa[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if ((threadIdx.x & 31) == 0) { sc[threadIdx.x >> 5] = cnt; sa[threadIdx.x >> 5] = acc; }
  __syncthreads();
  if (threadIdx.x < 32) {
    cnt = threadIdx.x < (blockDim.x >> 5) ? sc[threadIdx.x] : 0.f;
    acc = threadIdx.x < (blockDim.x >> 5) ? sa[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (threadIdx.x == 0) { area[row] = cnt; ssum[row] = acc; }
  }
}

// ---------------------------------------------------------------- final masks: bilinear upsample + threshold + boxes
// planerecnet.py:272-286: F.interpolate(seg, size=(H,W), bilinear, align_corners=False) > thr, then the tight
// box (xmin, ymin, xmax, ymax) of every mask.  seg rows are selected through `sel` (row index into seg).
__global__ void upsample_mask_box_kernel(const float* __restrict__ seg, const int* __restrict__ sel, bool* __restrict__ masks,
                                         int* __restrict__ boxes, int h, int w, int H, int W, float thr) {
  const int inst = blockIdx.y;
  const float* s = seg + static_cast<long long>(sel[inst]) * h * w;
  bool* mo = masks + static_cast<long long>(inst) * H * W;
  const float sh = static_cast<float>(h) / static_cast<float>(H), sw = static_cast<float>(w) / static_cast<float>(W);
  int xmin = W, ymin = H, xmax = -1, ymax = -1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int y = i / W, x = i - y * W;
    float fy = (static_cast<float>(y) + 0.5f) * sh - 0.5f;
    float fx = (static_cast<float>(x) + 0.5f) * sw - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = min(static_cast<int>(fy), h - 1), x0 = min(static_cast<int>(fx), w - 1);
    const int y1 = y0 < h - 1 ? y0 + 1 : y0, x1 = x0 < w - 1 ? x0 + 1 : x0;
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v = hy * (hx * __ldg(s + y0 * w + x0) + lx * __ldg(s + y0 * w + x1)) +
                    ly * (hx * __ldg(s + y1 * w + x0) + lx * __ldg(s + y1 * w + x1));
    const bool on = v > thr;
    mo[i] = on;
    if (on) {
      xmin = min(xmin, x); xmax = max(xmax, x);
      ymin = min(ymin, y); ymax = max(ymax, y);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
    ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
  }
  if ((threadIdx.x & 31) == 0 && xmax >= 0) {
    atomicMin(boxes + inst * 4 + 0, xmin);
    atomicMin(boxes + inst * 4 + 1, ymin);
    atomicMax(boxes + inst * 4 + 2, xmax);
    atomicMax(boxes + inst * 4 + 3, ymax);
  }
}

}  // namespace prn

using namespace prn;

// bool [n] (one byte each, 0/1) -> bits, LSB first: out[i] bit j = in[8 i + j].  16 input bytes per thread.
__global__ void pack_mask_bits_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_pairs) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_pairs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t bits = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t x = w[k] & 0x01010101u;                       // bytes b0..b3 -> bits 0, 8, 16, 24
      const uint32_t nib = (x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xFu;
      bits |= nib << (4 * k);
    }
    reinterpret_cast<uint16_t*>(out)[i] = static_cast<uint16_t>(bits);
  }
}

extern "C" {

int prn_pack_mask_bits(const void* masks_bool, void* out_bits, int64_t n_bool, void* stream) {
  PRN_REQUIRE(masks_bool && out_bits && n_bool > 0 && n_bool % 16 == 0 && (reinterpret_cast<uintptr_t>(masks_bool) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out_bits) & 1) == 0,
              "pack_mask_bits: need a 16-byte aligned input of a multiple of 16 bools");
  const long long n_pairs = n_bool / 16;
  const int threads = 256;
  long long blocks = (n_pairs + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_mask_bits_kernel<<<static_cast<int>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(masks_bool), static_cast<uint8_t*>(out_bits), n_pairs);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "pack_mask_bits launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}

int prn_point_nms_sigmoid(const float* logits, float* scores, int32_t batch, int32_t total, int32_t ld, int32_t nc,
                          int32_t n_levels, const int32_t* grids_dev, void* stream) {
  PRN_REQUIRE(logits && scores && grids_dev && batch > 0 && total > 0 && nc > 0 && ld >= nc && n_levels > 0,
              "point_nms_sigmoid: bad arguments");
  const long long n = static_cast<long long>(batch) * total * nc;
  const int grid = static_cast<int>((n + 255) / 256);
  point_nms_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, scores, batch, total, ld, nc, n_levels, grids_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "point_nms_sigmoid launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}

int prn_mask_stats(const float* seg, void* mask16, float* area, float* ssum, int32_t rows, int32_t pixels, float thr,
                   int32_t dtype, void* stream) {
  PRN_REQUIRE(seg && mask16 && area && ssum && rows > 0 && pixels > 0 && pixels % 4 == 0, "mask_stats: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == PRN_BF16)
    mask_stats_kernel<__nv_bfloat16><<<rows, 256, 0, st>>>(seg, static_cast<__nv_bfloat16*>(mask16), area, ssum, pixels, thr);
  else if (dtype == PRN_F16)
    mask_stats_kernel<__half><<<rows, 256, 0, st>>>(seg, static_cast<__half*>(mask16), area, ssum, pixels, thr);
  else
    return set_error(PRN_ERR_INVALID, "mask_stats: bad dtype");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "mask_stats launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}

int prn_upsample_mask_box(const float* seg, const int32_t* sel, void* masks_bool, int32_t* boxes, int32_t n_inst, int32_t h,
                          int32_t w, int32_t h_out, int32_t w_out, float thr, void* stream) {
  PRN_REQUIRE(seg && sel && masks_bool && boxes && n_inst > 0 && h > 0 && w > 0 && h_out > 0 && w_out > 0,
              "upsample_mask_box: bad arguments");
  dim3 grid((h_out * w_out + 256 * 8 - 1) / (256 * 8), n_inst);
  upsample_mask_box_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(seg, sel, static_cast<bool*>(masks_bool), boxes, h,
                                                                              w, h_out, w_out, thr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "upsample_mask_box launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- greedy mask-NMS (nms.py:53-80, nms_type == 'mask')
// One CTA per image over the score-sorted candidates.  inter [B][n][n] fp32 = exact mask intersections (Gram matrix of
// the 0/1 masks); candidate i, if still kept, removes every later kept j of the same label whose IoU
// inter/(area_i + area_j - inter) exceeds thr (or whose union is 0), in the reference's i-then-j order.
namespace prn {
__global__ void mask_nms_greedy_kernel(const float* __restrict__ inter, const float* __restrict__ area,
                                       const long long* __restrict__ labels, const unsigned char* __restrict__ valid,
                                       unsigned char* __restrict__ keep, int n, float thr) {
  extern __shared__ unsigned char kp[];
  const int b = blockIdx.x;
  const float* I = inter + static_cast<long long>(b) * n * n;
  const float* A = area + static_cast<long long>(b) * n;
  const long long* Lb = labels + static_cast<long long>(b) * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) kp[j] = valid[static_cast<long long>(b) * n + j];
  __syncthreads();
  for (int i = 0; i + 1 < n; ++i) {
    if (kp[i]) {                      // uniform across the CTA (read after the barrier)
      const float ai = A[i];
      const long long li = Lb[i];
      for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
        if (!kp[j] || Lb[j] != li) continue;
        const float in = I[static_cast<long long>(i) * n + j];
        const float un = ai + A[j] - in;
        if (un > 0.f) {
          if (in / un > thr) kp[j] = 0;
        } else {
          kp[j] = 0;
        }
      }
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < n; j += blockDim.x) keep[static_cast<long long>(b) * n + j] = kp[j];
}
}  // namespace prn

extern "C" int prn_mask_nms_greedy(const float* inter, const float* area, const int64_t* labels, const uint8_t* valid,
                                   uint8_t* keep, int32_t batch, int32_t n, float thr, void* stream) {
  using namespace prn;
  PRN_REQUIRE(inter && area && labels && valid && keep && batch > 0 && n > 0 && n <= 48 * 1024, "mask_nms_greedy: bad arguments");
  mask_nms_greedy_kernel<<<batch, 256, n, static_cast<cudaStream_t>(stream)>>>(inter, area, reinterpret_cast<const long long*>(labels),
                                                                               valid, keep, n, thr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(PRN_ERR_CUDA, "mask_nms_greedy launch: %s", cudaGetErrorString(e));
  return PRN_OK;
}
