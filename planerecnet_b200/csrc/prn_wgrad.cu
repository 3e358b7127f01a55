// prn_wgrad.cu — weight gradient of a convolution as a split-K implicit GEMM on tcgen05 (sm_100a).
//
//   dW[n, (ky,kx,c)] += sum_m dY[m, n] * X_im2col[m, (ky,kx,c)]          m = (image, ho, wo)
//
// The contraction runs over output pixels, so both operands are "MN-major" for the tensor core: dY rows
// (pixel-major, Cout contiguous) are the A operand, im2col rows (pixel-major, Cin contiguous) the B operand;
// each 64-pixel k-block is a stack of eight 8-row x 128-byte swizzle atoms per 64-wide M/N atom.
//
//   A (dY)   TMA boxes {64 cout, 64 pixels}, 128B swizzle, m_sub * 2 boxes per k-block (m_sub*128 cout rows per CTA)
//   B (X)    gathered with 16-byte cp.async exactly like the forward A operand (zero / reflect padding, nearest x2,
//            stride, two-source concat are address arithmetic); up to 4 atoms (tap, 64-channel block) = N 256
//   D        fp32 in TMEM: m_sub accumulators of 128 lanes x 256 columns; one work unit (M tile, N tile, K split)
//            per CTA; partial sums leave through vectorised fp32 reductions (red.global.add.v4.f32)
//
// Warp roles (320 threads): warps 0-7 gather B (the cp.async path moves 12 B/clk/SM with 256 threads, 7 with 128:
// profiles/r02_probe.txt), warps 0-3 then drain TMEM; warp 8 = TMA producer of A; warp 9 = MMA issuer.
// The reference has no hand-written backward: this is autograd's conv weight gradient for the nn.Conv2d / F.conv2d
// call sites listed in include/prn_b200.h (models/backbone.py:56-66, models/fpn.py:55,61, planerecnet.py:386-391,
// 478-495, 593-605).
#include <stdio.h>
#include <stdlib.h>
#include "prn_internal.h"
#include "prn_ptx.cuh"

namespace prn {

constexpr int kWgGatherWarps = 16;            // im2col gather warps (the first four also drain TMEM)
constexpr int kWgGatherThreads = kWgGatherWarps * 32;
constexpr int kWgRowsPerThread = 64 * 8 / kWgGatherThreads;   // 16-byte pieces per thread, atom and k-block
constexpr int kWgThreads = kWgGatherThreads + 64;
constexpr int kWgKBlock = 64;                 // pixels per k-block
constexpr int kWgAtomBytes = kWgKBlock * 128; // one [64 pixels][64 channels] 16-bit atom tile
constexpr int kWgMaxAtoms = 4;                // N tile = 4 atoms = 256 columns
constexpr int kWgSmemBudget = 227 * 1024;

struct WgradKParams {
  PrnWgrad d;
  int ld0, ld1;
  int m_rows;
  int hw_out;
  float inv_hw_out, inv_w_out;
  int kb_per_tap;
  int atoms;
  int n_tiles;
  int m_sub;
  int m_tiles;
  int splits;
  int kb_total;
  int kb_per_split;
  int stages;
  uint32_t idesc_base;   // formats + majors + M; N is added per tile
  uint32_t lbo, sbo;     // descriptor byte offsets of the MN-major operand tiles
  int b_tma;             // 1: 1x1 / stride 1 / single source: the B operand is plain pixel rows -> TMA boxes, no LSU gather
};

__device__ __forceinline__ void wg_divmod(int m, int dv, float inv, int& q, int& r) {
  q = __float2int_rz(__int2float_rz(m) * inv);
  r = m - q * dv;
  if (r < 0) { --q; r += dv; }
  if (r >= dv) { ++q; r -= dv; }
}

// MN-major operand, 128-byte swizzle: 8 k-rows x 128 B atoms; SBO = stride between 8-row k groups, LBO = stride
// between 64-element atoms along M/N (cute::UMMA::make_umma_desc<Major::MN>, SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <typename T>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ WgradKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);
  const uint32_t bar_full = base;          // [8] x 8 B
  const uint32_t bar_empty = base + 64;    // [8] x 8 B
  const uint32_t bar_tfull = base + 128;
  const uint32_t tmem_slot = base + 136;
  const uint32_t a_stage_bytes = static_cast<uint32_t>(p.m_sub) * 2u * kWgAtomBytes;
  const uint32_t b_stage_bytes = kWgMaxAtoms * kWgAtomBytes;
  const uint32_t a_base = base + 1024;
  const uint32_t b_base = a_base + static_cast<uint32_t>(p.stages) * a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const PrnWgrad& d = p.d;

  // work unit of this CTA
  const int unit = static_cast<int>(blockIdx.x);
  const int split = unit % p.splits;
  const int nt = (unit / p.splits) % p.n_tiles;
  const int mt = unit / (p.splits * p.n_tiles);
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int na = min(kWgMaxAtoms, p.atoms - nt * kWgMaxAtoms);
  const uint32_t tmem_cols = p.m_sub == 2 ? 512u : 256u;

  if (warp == kWgGatherWarps && lane == 0) {
    tma_prefetch_desc(&tmap_dy);
    if (p.b_tma) tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, p.b_tma ? 2 : kWgGatherThreads + 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_fence_init();
  }
  if (warp == kWgGatherWarps + 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + 136);

  if (warp < kWgGatherWarps) {
   if (p.b_tma) {
    // =========================================================== B producer, 1x1 convs: one TMA box {64 ch, 64 pixels} per atom
    if (threadIdx.x == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * s, static_cast<uint32_t>(na) * kWgAtomBytes);
        const uint32_t b_stage = b_base + static_cast<uint32_t>(s) * b_stage_bytes;
        for (int j = 0; j < na; ++j)
          tma_load_2d(b_stage + static_cast<uint32_t>(j) * kWgAtomBytes, &tmap_x, bar_full + 8 * s, (nt * kWgMaxAtoms + j) * 64,
                      kb * kWgKBlock);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
   } else {
    // =========================================================== B producer: im2col gather of 64 pixels x na atoms
    const int tid = threadIdx.x;
    const int chunk = tid & 7;          // 16-byte chunk (8 channels) of the 128-byte row
    const int r0 = tid >> 3;            // rows r0 + kRowStep * i of the k-block
    constexpr int kRowStep = kWgGatherThreads / 8;
    const uint32_t swz = static_cast<uint32_t>((chunk ^ (r0 & 7)) << 4);
    const int ups_shift = d.upsample == 2 ? 1 : 0;
    const int h_eff = d.h_in << ups_shift, w_eff = d.w_in << ups_shift;
    const uint8_t* a_src[kWgMaxAtoms];
    uint32_t a_pitch[kWgMaxAtoms];
    int a_ky[kWgMaxAtoms], a_kx[kWgMaxAtoms];
#pragma unroll
    for (int j = 0; j < kWgMaxAtoms; ++j) {
      const int ai = min(nt * kWgMaxAtoms + j, p.atoms - 1);
      const int tap = ai / p.kb_per_tap, cc = ai - tap * p.kb_per_tap;
      const int c = cc * 64;
      const bool first = c < d.c0;
      a_src[j] = static_cast<const uint8_t*>(first ? d.src0 : d.src1) + static_cast<size_t>((first ? c : c - d.c0) + chunk * 8) * 2;
      a_pitch[j] = static_cast<uint32_t>(first ? p.ld0 : p.ld1) * 2u;
      a_ky[j] = tap / d.ksize;
      a_kx[j] = tap - a_ky[j] * d.ksize;
    }
    int s = 0;
    uint32_t ph = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      int img_pix[kWgRowsPerThread], hy[kWgRowsPerThread], wx[kWgRowsPerThread];
      uint32_t valid = 0;
#pragma unroll
      for (int i = 0; i < kWgRowsPerThread; ++i) {
        const int m = kb * kWgKBlock + r0 + kRowStep * i;
        const bool v = m < p.m_rows;
        const int mm = v ? m : 0;
        int img, rem, ho, wo;
        wg_divmod(mm, p.hw_out, p.inv_hw_out, img, rem);
        wg_divmod(rem, d.w_out, p.inv_w_out, ho, wo);
        img_pix[i] = img * d.h_in * d.w_in;
        hy[i] = ho * d.stride - d.pad;
        wx[i] = wo * d.stride - d.pad;
        valid |= (v ? 1u : 0u) << i;
      }
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      const uint32_t b_stage = b_base + static_cast<uint32_t>(s) * b_stage_bytes;
#pragma unroll
      for (int j = 0; j < kWgMaxAtoms; ++j) {
        if (j < na) {
#pragma unroll
          for (int i = 0; i < kWgRowsPerThread; ++i) {
            int y = hy[i] + a_ky[j], x = wx[i] + a_kx[j];
            bool in = (valid >> i) & 1u;
            if (d.pad_mode == PRN_PAD_REFLECT) {
              y = y < 0 ? -y : (y >= h_eff ? 2 * h_eff - 2 - y : y);
              x = x < 0 ? -x : (x >= w_eff ? 2 * w_eff - 2 - x : x);
            } else {
              in = in && static_cast<unsigned>(y) < static_cast<unsigned>(h_eff) &&
                   static_cast<unsigned>(x) < static_cast<unsigned>(w_eff);
            }
            const uint32_t pix = in ? static_cast<uint32_t>(img_pix[i] + (y >> ups_shift) * d.w_in + (x >> ups_shift)) : 0u;
            cp_async16(b_stage + static_cast<uint32_t>(j) * kWgAtomBytes + static_cast<uint32_t>(r0 + kRowStep * i) * 128u + swz,
                       a_src[j] + static_cast<size_t>(pix) * a_pitch[j], in ? 16u : 0u);
          }
        }
      }
      cp_async_mbar_arrive_noinc(bar_full + 8 * s);
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
   }

    // =========================================================== drain: TMEM -> fp32 reductions into dW (warps 0-3)
    if (warp < 4) {
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    const int q = warp;   // TMEM lane quarter
    const int ncols = na * 64;
    for (int sub = 0; sub < p.m_sub; ++sub) {
      const int row = (mt * p.m_sub + sub) * 128 + q * 32 + lane;
      const bool valid = row < d.n;
      float* orow = d.dw + static_cast<size_t>(valid ? row : 0) * d.ld_dw + static_cast<size_t>(nt) * (kWgMaxAtoms * 64);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(sub * 256);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t vb[32];
        __syncwarp();
        tmem_ld_x32(t_row + c0, vb);
        tmem_ld_wait();
        tmem_ld_publish16(vb);
        tmem_ld_publish16(vb + 16);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4_f32(orow + c0 + 4 * j, __uint_as_float(vb[4 * j]), __uint_as_float(vb[4 * j + 1]),
                       __uint_as_float(vb[4 * j + 2]), __uint_as_float(vb[4 * j + 3]));
        }
      }
    }
    }
  } else if (warp == kWgGatherWarps) {
    // =========================================================== A producer: TMA boxes of dY
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * s, a_stage_bytes);
        const uint32_t a_stage = a_base + static_cast<uint32_t>(s) * a_stage_bytes;
        for (int h = 0; h < 2 * p.m_sub; ++h)
          tma_load_2d(a_stage + static_cast<uint32_t>(h) * kWgAtomBytes, &tmap_dy, bar_full + 8 * s,
                      (mt * p.m_sub * 2 + h) * 64, kb * kWgKBlock);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =========================================================== MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t idesc = p.idesc_base | ((static_cast<uint32_t>(na * 64) >> 3) << 17);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * s, ph);
        fence_proxy_async_smem();
        tc_fence_after();
        const uint32_t a_stage = a_base + static_cast<uint32_t>(s) * a_stage_bytes;
        const uint32_t b_stage = b_base + static_cast<uint32_t>(s) * b_stage_bytes;
        for (int sub = 0; sub < p.m_sub; ++sub) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixels = two 8-row swizzle atoms = 2048 B further into every 64-wide atom tile
            umma_f16(tmem_base + static_cast<uint32_t>(sub * 256),
                     umma_desc_sw128_mn(a_stage + static_cast<uint32_t>(sub) * 2u * kWgAtomBytes + k * 2048, p.lbo, p.sbo),
                     umma_desc_sw128_mn(b_stage + k * 2048, p.lbo, p.sbo), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      umma_commit(bar_tfull);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kWgGatherWarps + 1) tmem_dealloc(tmem_base, tmem_cols);
}

static int wgrad_plan(const PrnWgrad& d, WgradKParams* p) {
  PRN_REQUIRE(d.src0 != nullptr && d.dy != nullptr && d.dw != nullptr, "wgrad: src0/dy/dw must be non-NULL");
  PRN_REQUIRE(d.c0 > 0 && d.c0 % 64 == 0 && d.c1 >= 0 && d.c1 % 64 == 0, "wgrad: channel counts must be multiples of 64 (c0=%d c1=%d)", d.c0, d.c1);
  PRN_REQUIRE(d.c1 == 0 || d.src1 != nullptr, "wgrad: src1 is NULL but c1=%d", d.c1);
  PRN_REQUIRE(d.batch > 0 && d.h_in > 0 && d.w_in > 0 && d.h_out > 0 && d.w_out > 0, "wgrad: bad spatial dims");
  PRN_REQUIRE(d.upsample == 1 || d.upsample == 2, "wgrad: upsample must be 1 or 2");
  PRN_REQUIRE(d.ksize >= 1 && d.ksize <= 7 && d.stride >= 1 && d.pad >= 0, "wgrad: bad ksize/stride/pad");
  PRN_REQUIRE(d.pad_mode == PRN_PAD_ZERO || d.pad_mode == PRN_PAD_REFLECT, "wgrad: bad pad_mode");
  PRN_REQUIRE(d.dtype == PRN_BF16 || d.dtype == PRN_F16, "wgrad: dtype must be PRN_BF16 or PRN_F16");
  PRN_REQUIRE(d.n > 0 && d.ld_dy >= d.n && d.ld_dy % 8 == 0, "wgrad: need n > 0 and ld_dy >= n, multiple of 8 (n=%d ld_dy=%d)", d.n, d.ld_dy);
  const int h_eff = d.h_in * d.upsample, w_eff = d.w_in * d.upsample;
  PRN_REQUIRE((h_eff + 2 * d.pad - d.ksize) / d.stride + 1 == d.h_out && (w_eff + 2 * d.pad - d.ksize) / d.stride + 1 == d.w_out,
              "wgrad: h_out/w_out inconsistent with input dims (%dx%d -> %dx%d)", h_eff, w_eff, d.h_out, d.w_out);
  PRN_REQUIRE(d.pad_mode != PRN_PAD_REFLECT || (d.pad < h_eff && d.pad < w_eff), "wgrad: reflect pad too large");
  PRN_REQUIRE(d.ld0 == 0 || (d.ld0 >= d.c0 && d.ld0 % 8 == 0), "wgrad: bad ld0");
  PRN_REQUIRE(d.ld1 == 0 || (d.ld1 >= d.c1 && d.ld1 % 8 == 0), "wgrad: bad ld1");
  const int ktot = d.ksize * d.ksize * (d.c0 + d.c1);
  PRN_REQUIRE(d.ld_dw >= ktot && d.ld_dw % 4 == 0 && (reinterpret_cast<uintptr_t>(d.dw) & 15) == 0,
              "wgrad: dw must be 16-byte aligned with ld_dw >= ksize^2*(c0+c1) = %d, multiple of 4 (got %d)", ktot, d.ld_dw);
  p->d = d;
  p->ld0 = d.ld0 ? d.ld0 : d.c0;
  p->ld1 = d.ld1 ? d.ld1 : d.c1;
  p->hw_out = d.h_out * d.w_out;
  const long long m_rows = static_cast<long long>(d.batch) * p->hw_out;
  PRN_REQUIRE(m_rows < (1 << 24), "wgrad: more than 2^24 output pixels is not supported");
  PRN_REQUIRE(static_cast<unsigned long long>(d.batch) * d.h_in * d.w_in * p->ld0 * 2ull < (1ull << 32) &&
                  static_cast<unsigned long long>(d.batch) * d.h_in * d.w_in * (p->ld1 ? p->ld1 : 1) * 2ull < (1ull << 32),
              "wgrad: source tensors of 4 GiB or more are not supported");
  p->m_rows = static_cast<int>(m_rows);
  p->inv_hw_out = 1.0f / static_cast<float>(p->hw_out);
  p->inv_w_out = 1.0f / static_cast<float>(d.w_out);
  p->kb_per_tap = (d.c0 + d.c1) / 64;
  p->atoms = d.ksize * d.ksize * p->kb_per_tap;
  p->n_tiles = ceil_div(p->atoms, kWgMaxAtoms);
  p->m_sub = d.n > 128 ? 2 : 1;
  p->m_tiles = ceil_div(d.n, 128 * p->m_sub);
  p->kb_total = ceil_div(p->m_rows, kWgKBlock);
  // split-K factor: `waves` CTAs per SM in total.  More splits shorten the K loop of a CTA but multiply the fp32 reduction traffic
  // (every split adds its whole 128 x 256 x m_sub accumulator into dW with red.global)
  static const int waves = [] { const char* e = getenv("PRN_WGRAD_WAVES"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : 1; }();
  // measured (profiles/r02_wgrad_waves.txt): one CTA per SM beats two (R101 step 30.7 -> 29.4 ms): the reduction traffic of the
  // extra splits costs more than the shorter K loops save
  static const int min_kb = [] { const char* e = getenv("PRN_WGRAD_MINKB"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 64 ? v : 16; }();   // >= 16 k-blocks per split: 30.7 -> 29.0 ms per R101 step together with waves = 1
  auto splits_for = [&](int tiles_) {
    int s_ = (waves * sm_count()) / tiles_;
    if (s_ > p->kb_total / min_kb) s_ = p->kb_total / min_kb;
    return s_ < 1 ? 1 : s_;
  };
  // 256-row CTA tiles (two 128-row accumulators sharing one gathered operand) halve the gather traffic, but the small maps
  // (30x40, 15x20: few k-blocks to split) then fill a quarter of the SMs with CTAs that issue twice the MMAs each: when the
  // 256-row plan cannot occupy half of the machine, use 128-row tiles (same reduction traffic, twice the CTAs)
  static const bool msub_auto = [] { const char* e = getenv("PRN_WGRAD_MSUB_AUTO"); return !(e != nullptr && e[0] == '0'); }();
  if (msub_auto && p->m_sub == 2) {
    const int tiles2 = p->m_tiles * p->n_tiles;
    if (tiles2 * splits_for(tiles2) * 2 <= sm_count()) {
      p->m_sub = 1;
      p->m_tiles = ceil_div(d.n, 128);
    }
  }
  const int tiles = p->m_tiles * p->n_tiles;
  int splits = splits_for(tiles);
  p->kb_per_split = ceil_div(p->kb_total, splits);
  p->splits = ceil_div(p->kb_total, p->kb_per_split);
  const int stage_bytes = p->m_sub * 2 * kWgAtomBytes + kWgMaxAtoms * kWgAtomBytes;
  int stages = (kWgSmemBudget - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  p->stages = stages;
  const uint32_t fmt = d.dtype == PRN_BF16 ? 1u : 0u;
  p->idesc_base = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 4) << 24);
  p->lbo = (d.flags & 1) ? 1024u : static_cast<uint32_t>(kWgAtomBytes);
  p->sbo = (d.flags & 1) ? static_cast<uint32_t>(kWgAtomBytes) : 1024u;
  static const bool tma_off = [] { const char* e = getenv("PRN_WGRAD_TMA"); return e != nullptr && e[0] == '0'; }();
  p->b_tma = (!tma_off && d.ksize == 1 && d.stride == 1 && d.upsample == 1 && d.pad == 0 && d.c1 == 0 &&
              (reinterpret_cast<uintptr_t>(d.src0) & 15) == 0 && p->ld0 % 8 == 0) ? 1 : 0;
  return PRN_OK;
}

template <typename T>
static int wgrad_launch_t(const CUtensorMap& tm, const CUtensorMap& tmx, const WgradKParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PRN_CUDA(cudaFuncSetAttribute(wgrad_umma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBudget));
    configured = true;
  }
  const size_t smem = 2048 + static_cast<size_t>(p.stages) * (p.m_sub * 2 * kWgAtomBytes + kWgMaxAtoms * kWgAtomBytes);
  const int grid = p.m_tiles * p.n_tiles * p.splits;
  wgrad_umma_kernel<T><<<grid, kWgThreads, smem, st>>>(tm, tmx, p);
  PRN_CUDA(cudaGetLastError());
  return PRN_OK;
}

}  // namespace prn

extern "C" int prn_conv2d_wgrad_plan(const PrnWgrad* desc, int32_t* out8) {
  if (!desc || !out8) return prn::set_error(PRN_ERR_INVALID, "wgrad: NULL argument");
  prn::WgradKParams p;
  int rc = prn::wgrad_plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  out8[0] = p.m_tiles; out8[1] = p.n_tiles; out8[2] = p.splits; out8[3] = p.kb_per_split;
  out8[4] = p.m_sub; out8[5] = p.stages; out8[6] = p.m_tiles * p.n_tiles * p.splits; out8[7] = p.atoms;
  return PRN_OK;
}

extern "C" int prn_conv2d_wgrad(const PrnWgrad* desc, void* stream) {
  using namespace prn;
  if (!desc) return set_error(PRN_ERR_INVALID, "wgrad: NULL descriptor");
  WgradKParams p;
  int rc = wgrad_plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  CUtensorMap tm;
  rc = encode_tmap_2d_sw128(&tm, desc->dy, static_cast<uint64_t>(p.m_rows), static_cast<uint64_t>(desc->n), kWgKBlock,
                            desc->dtype, static_cast<uint64_t>(desc->ld_dy));
  if (rc != PRN_OK) return rc;
  CUtensorMap tmx = tm;
  if (p.b_tma) {
    rc = encode_tmap_2d_sw128(&tmx, desc->src0, static_cast<uint64_t>(p.m_rows), static_cast<uint64_t>(desc->c0), kWgKBlock,
                              desc->dtype, static_cast<uint64_t>(p.ld0));
    if (rc != PRN_OK) return rc;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // timing experiments only (tools/graph_profile.py): PRN_WGRAD_SKIP=1 drops the launch, which shows what the weight-gradient
  // kernels on the side streams cost the main chain (gradients are then wrong, of course)
  static const bool skip = [] { const char* e = getenv("PRN_WGRAD_SKIP"); return e != nullptr && e[0] == '1'; }();
  if (skip) return PRN_OK;
  if (desc->dtype == PRN_BF16) return wgrad_launch_t<__nv_bfloat16>(tm, tmx, p, st);
  return wgrad_launch_t<__half>(tm, tmx, p, st);
}
