// prn_conv_common.cuh — pieces shared by the implicit-GEMM convolution kernels (prn_conv.cu: cp.async / deformable gather
// producers; prn_conv_tma.cu: TMA halo-tile producers): launch parameters, fast index decode, the epilogue math of one
// accumulator chunk (bias, residual, GroupNorm / BatchNorm partial sums, activations, 16-bit / fp32 stores).
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "prn_internal.h"
#include "prn_ptx.cuh"

namespace prn {

constexpr int kTileM = 128;
constexpr int kATileBytes = kTileM * 128;  // 128 rows x 64 16-bit elements
// 16 warps = 4 warpgroups: WG0 epilogue A | WG1 epilogue B (plain conv) or A-producer B (deformable) |
// WG2 A-producer A | WG3 warp 12 = TMA weight producer, warp 13 = MMA issuer + TMEM owner, 14-15 idle.
// Registers are re-balanced per warpgroup with setmaxnreg (128/thread at launch).
constexpr int kThreads = 512;
constexpr int kEpiWarps = 4;   // warps per epilogue group (one per TMEM lane quarter)
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kStageOutBytes = 8 * 4096;   // per epilogue warp: one 32-row x 128 B staging tile for TMA stores
constexpr int kCtrlBytes = 2048;   // mbarriers + TMEM slot (first 256 B), bias staging for the epilogue (+1024, 1 KB)

struct ConvKParams {
  PrnConv d;
  int m_group;        // rows per weight group
  int groups;
  int imgs_per_group;
  int m_tiles, n_tiles, total_tiles;
  int cluster;        // 1, or 2 = CTA pairs over consecutive M tiles sharing every weight k-block (TMA multicast)
  int m_ptiles;       // ceil(m_tiles / cluster)
  int total_ptiles;   // groups * m_ptiles * n_tiles: loop count of every CTA (pair)
  int n_tile;
  int stages;
  int tmem_cols;
  int kb_per_tap;
  int num_kb;
  int hw_out;
  int out_img_rows;
  int ld0, ld1;
  float inv_hw_out, inv_w_out;
  int lean_epi;       // 1: lean epilogue instantiation (see epi_chunk)
  int tma_store;      // 1: 16-bit output rows are dense -> epilogue stages 32x64 sub-tiles in smem and TMA-stores them
  uint32_t idesc;
  long long* dbg;   // optional role-level cycle counters of CTA 0 (prn_conv2d_fwd_profile)
  int dcn_dbg;      // timing experiments (env PRN_DCN_DEBUG): 1 = deformable producer without the blend, 2 = without the loads
};

// prn_conv_tma.cu: the TMA-fed kernel (3x3 stride 1 pad 1 through shared-memory halo tiles, 1x1 stride 1)
bool conv_tma_eligible(const PrnConv& d);
int conv_tma_launch(const PrnConv* desc, void* stream, long long* dbg);
int conv_tma_plan_ex(const PrnConv& d, int32_t* out8);

// wait on an mbarrier, optionally accumulating the stall cycles
__device__ __forceinline__ void mbar_wait_acc(uint32_t bar, uint32_t parity, bool prof, long long& acc) {
  if (!prof) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

// q = m / d, r = m % d for 0 <= m < 2^24 with inv = 1.0f / d: one multiply + fix-up instead of a ~40 instruction
// integer division (the producer and the epilogue decode 8 + 1 rows per tile).
__device__ __forceinline__ void fast_divmod(int m, int dv, float inv, int& q, int& r) {
  q = __float2int_rz(__int2float_rz(m) * inv);
  r = m - q * dv;
  if (r < 0) { --q; r += dv; }
  if (r >= dv) { ++q; r -= dv; }
}

template <int W>
__device__ __forceinline__ void act_apply(float* x, int act, float param, int col0) {
  switch (act) {
    case PRN_ACT_RELU:
#pragma unroll
      for (int j = 0; j < W; ++j) x[j] = fmaxf(x[j], 0.f);
      break;
    case PRN_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < W; ++j) x[j] = __fdividef(1.f, 1.f + __expf(-x[j]));
      break;
    case PRN_ACT_SIGMOID_AVG4:
      // sigmoid(x) = 0.5 + 0.5 tanh(x / 2): ONE special-function op (MUFU.TANH) instead of two (EX2 + RCP) — this epilogue is
      // MUFU-bound (plane-prior attention: 143 M sigmoids per batch of 8).  tanh.approx.f32 has a relative error of 2^-11, i.e.
      // <= 2.4e-4 absolute on the sigmoid: the size of the f16 rounding of the stored average
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x[j]));
        x[j] = fmaf(0.5f, t, 0.5f);
      }
      break;
    case PRN_ACT_SOFTPLUS:
#pragma unroll
      for (int j = 0; j < W; ++j) x[j] = x[j] > 20.f ? x[j] : log1pf(expf(x[j]));
      break;
    case PRN_ACT_DCN_OFFMASK:
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int col = col0 + j;
        x[j] = col < 18 ? fminf(fmaxf(x[j], -param), param) : (col < 27 ? 2.f / (1.f + __expf(-x[j])) : 0.f);
      }
      break;
    default: break;
  }
}

// W/8 16-byte residual loads of one row chunk (zeros when the row is out of range)
template <int W>
__device__ __forceinline__ void load_res(uint4* r, const void* ptr, bool on) {
#pragma unroll
  for (int h = 0; h < W / 8; ++h) r[h] = on ? __ldg(reinterpret_cast<const uint4*>(ptr) + h) : make_uint4(0, 0, 0, 0);
}

// Epilogue math for W accumulator columns of one output row (one thread = one TMEM lane).
// kFull = false is the lean instantiation used by most layers (bias, residual, none/ReLU, 16-bit output): the
// generic one (statistics, sigmoid/softplus/DCN activations, fp32 output, row averaging) is ~10x more code and
// thrashes the instruction cache when it sits inside the per-chunk loop.
template <typename T, int W, int kEpi, typename P>
__device__ __forceinline__ void epi_chunk(const P& p, float* x, const uint4* res, bool has_res,
                                          const float* bias_s, int col0, bool valid, bool img_uniform, int img,
                                          int lane, size_t orow, uint32_t stage_row, int jbase, int scol = -1) {
  const PrnConv& d = p.d;
  const int sc = scol >= 0 ? scol : col0;      // column of the direct stores (differs from col0 for pixel-shuffled outputs)
  constexpr bool kFull = kEpi != 0;      // kEpi: 0 lean | 1 full | 2 full + BatchNorm batch statistics (training step only)
  if (bias_s != nullptr) {
#pragma unroll
    for (int j = 0; j < W / 4; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * j);
      x[4 * j] += b.x; x[4 * j + 1] += b.y; x[4 * j + 2] += b.z; x[4 * j + 3] += b.w;
    }
  }
  if (has_res) {
#pragma unroll
    for (int h = 0; h < W / 8; ++h) {
      const uint32_t w4[4] = {res[h].x, res[h].y, res[h].z, res[h].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = Pack2<T>::unpack(w4[j]);
        x[8 * h + 2 * j] += f.x;
        x[8 * h + 2 * j + 1] += f.y;
      }
    }
  }
  if constexpr (!kFull) {
    const float lo = d.act == PRN_ACT_RELU ? 0.f : -INFINITY;
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = fmaxf(x[j], lo);
    if (stage_row != 0) {
#pragma unroll
      for (int h = 0; h < W / 8; ++h) {
        const uint32_t o0 = Pack2<T>::pack(x[8 * h], x[8 * h + 1]), o1 = Pack2<T>::pack(x[8 * h + 2], x[8 * h + 3]);
        const uint32_t o2 = Pack2<T>::pack(x[8 * h + 4], x[8 * h + 5]), o3 = Pack2<T>::pack(x[8 * h + 6], x[8 * h + 7]);
        // jbase >= 0: 128-byte rows (64 columns per staging tile, SWIZZLE_128B); jbase < 0: 64-byte rows (32 columns, SWIZZLE_64B)
        const uint32_t dst = jbase >= 0 ? stage_row + ((((jbase + h) ^ (lane & 7))) << 4)
                                        : stage_row + (((h ^ ((lane >> 1) & 3))) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
    } else if (valid) {
      uint4* op = reinterpret_cast<uint4*>(static_cast<T*>(d.out16) + orow * d.ld_out16 + sc);
#pragma unroll
      for (int h = 0; h < W / 8; ++h) {
        uint4 o;
        o.x = Pack2<T>::pack(x[8 * h], x[8 * h + 1]); o.y = Pack2<T>::pack(x[8 * h + 2], x[8 * h + 3]);
        o.z = Pack2<T>::pack(x[8 * h + 4], x[8 * h + 5]); o.w = Pack2<T>::pack(x[8 * h + 6], x[8 * h + 7]);
        op[h] = o;
      }
    }
    return;
  }
  if (d.stats != nullptr && d.stats_cg > 0) {
    // GroupNorm partial sums of the pre-normalisation conv output (fp32 accumulators).  Per 16 columns every
    // lane holds 4 (sum, sumsq) pairs over column quads; pairs are merged for 8/16-channel groups.  When the whole
    // warp belongs to one image the 8 values are reduce-scattered over the lanes with a butterfly (8+4+2+1
    // shuffles + 2 for the replicated pair instead of 8 x 5), and 8 lanes issue one atomic each.
    const int cg = d.stats_cg;
    const int G = d.n_pad / cg;
#pragma unroll
    for (int hf = 0; hf < W / 16; ++hf) {
      const float* xx = x + 16 * hf;
      float v[8];   // v[2*qd] = sum, v[2*qd+1] = sumsq of quad qd
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        v[2 * qd] = (xx[4 * qd] + xx[4 * qd + 1]) + (xx[4 * qd + 2] + xx[4 * qd + 3]);
        v[2 * qd + 1] = (xx[4 * qd] * xx[4 * qd] + xx[4 * qd + 1] * xx[4 * qd + 1]) +
                        (xx[4 * qd + 2] * xx[4 * qd + 2] + xx[4 * qd + 3] * xx[4 * qd + 3]);
      }
      const int colq = col0 + 16 * hf;
      if (img_uniform) {
        // reduce-scatter: after the 3 halving steps lane bits (4,3,2) select the value index; bits (1,0) replicate
        const bool h4 = (lane & 16) != 0, h3 = (lane & 8) != 0, h2 = (lane & 4) != 0;
        float w4[4], w2[2], w1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float keep = h4 ? v[4 + i] : v[i], send = h4 ? v[i] : v[4 + i];
          w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float keep = h3 ? w4[2 + i] : w4[i], send = h3 ? w4[i] : w4[2 + i];
          w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const float keep = h2 ? w2[1] : w2[0], send = h2 ? w2[0] : w2[1];
          w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
        w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
        if ((lane & 3) == 0) {
          const int vi = (h4 ? 4 : 0) + (h3 ? 2 : 0) + (h2 ? 1 : 0);     // value index 0..7 = quad * 2 + {sum, sumsq}
          const int gi = (colq + (vi >> 1) * 4) / cg;
          atomicAdd(d.stats + (static_cast<size_t>(img) * G + gi) * 2 + (vi & 1), w1);
        }
      } else if (valid) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const int gi = (colq + qd * 4) / cg;
          atomicAdd(d.stats + (static_cast<size_t>(img) * G + gi) * 2, v[2 * qd]);
          atomicAdd(d.stats + (static_cast<size_t>(img) * G + gi) * 2 + 1, v[2 * qd + 1]);
        }
      }
    }
  } else if (kEpi == 2 && d.stats != nullptr) {
    // BatchNorm batch statistics: per-channel {sum, sumsq} over the warp's 32 rows.  Butterfly reduce-scatter over
    // the lanes (W-1 shuffles per quantity instead of 5*W): after the halving steps lane l holds the total of column
    // bits(l), and every lane issues one atomic per quantity.
    float s1[W], s2[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      s1[j] = valid ? x[j] : 0.f;
      s2[j] = s1[j] * s1[j];
    }
    int col = 0;
    int off = 16;
#pragma unroll
    for (int n = W; n > 1; n >>= 1) {
      const bool hi = (lane & off) != 0;
      const int h = n >> 1;
#pragma unroll
      for (int i = 0; i < W / 2; ++i) {
        if (i < h) {
          const float k1 = hi ? s1[h + i] : s1[i], t1 = hi ? s1[i] : s1[h + i];
          const float k2 = hi ? s2[h + i] : s2[i], t2 = hi ? s2[i] : s2[h + i];
          s1[i] = k1 + __shfl_xor_sync(0xffffffffu, t1, off);
          s2[i] = k2 + __shfl_xor_sync(0xffffffffu, t2, off);
        }
      }
      col += hi ? h : 0;
      off >>= 1;
    }
    if constexpr (W == 16) {        // 16 columns over 32 lanes: lane pairs hold the two halves of a column's total
      s1[0] += __shfl_xor_sync(0xffffffffu, s1[0], 1);
      s2[0] += __shfl_xor_sync(0xffffffffu, s2[0], 1);
    }
    if (W == 32 || (lane & 1) == 0) {
      atomicAdd(d.stats + static_cast<size_t>(col0 + col) * 2, s1[0]);
      atomicAdd(d.stats + static_cast<size_t>(col0 + col) * 2 + 1, s2[0]);
    }
  }
  act_apply<W>(x, d.act, d.act_param, col0);
  bool store = valid;
  if (d.act == PRN_ACT_SIGMOID_AVG4) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      float t = x[j];
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      x[j] = 0.25f * t;
    }
    store = valid && (lane & 3) == 0;
  }
  if (stage_row != 0) {
    // 16-byte pieces of this row go to the 128B-swizzled staging tile; a TMA store writes it out coalesced
    // (rows beyond M and columns beyond n_pad are clipped by the tensor map).
#pragma unroll
    for (int h = 0; h < W / 8; ++h) {
      const uint32_t o0 = Pack2<T>::pack(x[8 * h], x[8 * h + 1]), o1 = Pack2<T>::pack(x[8 * h + 2], x[8 * h + 3]);
      const uint32_t o2 = Pack2<T>::pack(x[8 * h + 4], x[8 * h + 5]), o3 = Pack2<T>::pack(x[8 * h + 6], x[8 * h + 7]);
      const uint32_t dst = jbase >= 0 ? stage_row + ((((jbase + h) ^ (lane & 7))) << 4)
                                      : stage_row + (((h ^ ((lane >> 1) & 3))) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
    }
  } else if (store) {
    if (d.out16) {
      uint4* op = reinterpret_cast<uint4*>(static_cast<T*>(d.out16) + orow * d.ld_out16 + sc);
#pragma unroll
      for (int h = 0; h < W / 8; ++h) {
        uint4 o;
        o.x = Pack2<T>::pack(x[8 * h], x[8 * h + 1]); o.y = Pack2<T>::pack(x[8 * h + 2], x[8 * h + 3]);
        o.z = Pack2<T>::pack(x[8 * h + 4], x[8 * h + 5]); o.w = Pack2<T>::pack(x[8 * h + 6], x[8 * h + 7]);
        op[h] = o;
      }
    }
    if (d.out32) {
      float4* op = reinterpret_cast<float4*>(d.out32 + orow * d.ld_out32 + sc);
#pragma unroll
      for (int j = 0; j < W / 4; ++j) op[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    }
  }
}

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

}  // namespace prn
