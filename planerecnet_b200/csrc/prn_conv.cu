// prn_conv.cu — implicit-GEMM convolution for sm_100a.
//
// out[m, n] = act( sum_k A[m, k] * W[n, k] + bias[n] + residual[m, n] )
//   m = (image, ho, wo) flattened, 128 rows per CTA tile (one TMEM lane per row)
//   k = (ky, kx, c), walked in 64-element blocks (128 B rows, 128B-swizzled K-major smem tiles)
//   n = output channels, n_tile <= 256 columns of fp32 accumulators in TMEM, double buffered
//
// Warp roles (320 threads, persistent CTAs, static round-robin tile schedule):
//   warps 0-3  epilogue   tcgen05.ld -> bias/residual/activation/statistics -> NHWC stores
//   warps 4-7  A producer im2col gather straight from the NHWC activation: one 16-byte cp.async per
//                         (pixel, 8 channels), zero-fill / reflect / nearest-x2 / two-source concat
//                         resolved in the address computation (nothing is materialised in HBM);
//                         with `dcn_offmask` the rows are bilinear-sampled + modulated in registers
//                         (torchvision deform_conv2d semantics) and written with st.shared
//   warp 8     B producer one TMA box {64 k, n_tile rows} of the packed weights per k-block
//   warp 9     MMA issuer one thread issuing tcgen05.mma (M=128, N=n_tile, K=16) x4 per k-block and
//                         tcgen05.commit to recycle the smem stage / publish the accumulator
//
// Replaces the nn.Conv2d / F.conv2d / deform_conv2d call sites listed in include/prn_b200.h.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "prn_internal.h"
#include "prn_ptx.cuh"

#include "prn_conv_common.cuh"

namespace prn {

template <typename T, bool kDCN, int kEpi, int kCl>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                 const __grid_constant__ ConvKParams p) {
  constexpr bool kFull = kEpi != 0;
  constexpr int kEpiGroups = kDCN ? 1 : 2;              // epilogue warpgroups
  constexpr int kProdWarps = kDCN ? 8 : 4;              // A-producer warps
  constexpr int kRows = kDCN ? 4 : 8;                   // A rows per producer thread
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);

  // control block: mbarriers + TMEM base slot, then the epilogue's bias staging area
  const uint32_t bar_full = base;             // [kMaxStages] x 8 B
  const uint32_t bar_empty = base + 64;       // [kMaxStages] x 8 B
  const uint32_t bar_tfull = base + 128;      // [2]
  const uint32_t bar_tempty = base + 144;     // [2]
  const uint32_t tmem_slot = base + 160;
  float* bias_s = reinterpret_cast<float*>(base_ptr + 1024);   // [256]
  const uint32_t stg_base = base + kCtrlBytes;                  // [8 warps] x 4 KB, 1024-aligned
  const uint32_t a_base = stg_base + kStageOutBytes;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.n_tile) * 128u;
  const uint32_t b_base = a_base + static_cast<uint32_t>(p.stages) * kATileBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const PrnConv& d = p.d;
  // Programmatic dependent launch: let the next kernel of the stream start its prologue on SMs this grid has
  // already left; our own reads/writes of activations wait (below) until the previous grid has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    if (p.tma_store) tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, kProdWarps * 32 + 1);
      mbar_init(bar_empty + 8 * s, kCl);     // one tcgen05.commit per CTA of the cluster (they share the B stage)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kEpiGroups * kEpiWarps * 32);
    }
    mbar_fence_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + 160);
  const int crank = kCl > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  if constexpr (kCl > 1) cluster_sync_all();           // peer barriers are initialised before anything targets them
  asm volatile("griddepcontrol.wait;" ::: "memory");   // everything above overlapped the previous kernel's tail
  // tile walk shared by all roles: pair-tile pt -> (group g, M tile mt = mp * kCl + rank, N tile nt)
  const int pt0 = static_cast<int>(blockIdx.x) / kCl, pt_step = static_cast<int>(gridDim.x) / kCl;

  const bool is_prod = wg == 2 || (kDCN && wg == 1);
  const bool is_epi = wg == 0 || (!kDCN && wg == 1);

  if (is_prod) {
    // =========================================================== A producer
    if constexpr (kDCN) reg_inc<144>(); else reg_dec<104>();
    const int pw = kDCN ? ((wg == 2 ? 0 : 4) + (warp & 3)) : (warp & 3);
    const int sub = lane >> 3;    // row within a group of 4
    const int chunk = lane & 7;   // 16-byte chunk (8 channels) within the 128-byte row
    const int ups_shift = d.upsample == 2 ? 1 : 0;
    const int h_eff = d.h_in << ups_shift, w_eff = d.w_in << ups_shift;
    int s = 0;
    uint32_t ph = 0;
    uint32_t it = 0;
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && warp == 8 && lane == 0;
    long long w_empty = 0;
    uint32_t row_off[kRows];   // swizzled position of this thread's 16-byte chunk in each of its rows of an A stage
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int r = pw * (4 * kRows) + i * 4 + sub;
      row_off[i] = static_cast<uint32_t>(r * 128 + ((chunk ^ (r & 7)) << 4));
    }
    const long long t_role0 = prof ? clock64() : 0;
    for (int tile = pt0; tile < p.total_ptiles; tile += pt_step) {
      const int rest = tile / p.n_tiles;
      const int mt = (rest % p.m_ptiles) * kCl + crank;
      const int g = rest / p.m_ptiles;
      int img_pix[kRows], hy[kRows], wx[kRows];
      int m_glob[kRows];
      uint32_t valid = 0;
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int r = pw * (4 * kRows) + i * 4 + sub;
        const int m = mt * kTileM + r;
        const bool v = m < p.m_group;
        const int mm = v ? m : 0;
        int n_local, rem, ho, wo;
        fast_divmod(mm, p.hw_out, p.inv_hw_out, n_local, rem);
        fast_divmod(rem, d.w_out, p.inv_w_out, ho, wo);
        const int img = g * p.imgs_per_group + n_local;
        img_pix[i] = img * d.h_in * d.w_in;
        hy[i] = ho * d.stride - d.pad;
        wx[i] = wo * d.stride - d.pad;
        m_glob[i] = g * p.m_group + mm;
        valid |= (v ? 1u : 0u) << i;
      }
      const int taps = d.ksize * d.ksize;
      int ky = 0, kx = -1;
      for (int tap = 0; tap < taps; ++tap) {
        if (++kx == d.ksize) { kx = 0; ++ky; }
        if constexpr (!kDCN) {
          // per tap: pixel index of every row's source (0 when the tap falls outside: the address stays valid
          // and the copy is issued with src-size 0 = zero fill)
          uint32_t poff[kRows];
          uint32_t ok = 0;
#pragma unroll
          for (int i = 0; i < kRows; ++i) {
            int y = hy[i] + ky, x = wx[i] + kx;
            bool in = (valid >> i) & 1u;
            if (d.pad_mode == PRN_PAD_REFLECT) {
              y = y < 0 ? -y : (y >= h_eff ? 2 * h_eff - 2 - y : y);
              x = x < 0 ? -x : (x >= w_eff ? 2 * w_eff - 2 - x : x);
            } else if (d.pad_mode == PRN_PAD_CLAMP) {
              y = min(max(y, 0), h_eff - 1);
              x = min(max(x, 0), w_eff - 1);
            } else {
              in = in && static_cast<unsigned>(y) < static_cast<unsigned>(h_eff) &&
                   static_cast<unsigned>(x) < static_cast<unsigned>(w_eff);
            }
            const int pix = in ? img_pix[i] + (y >> ups_shift) * d.w_in + (x >> ups_shift) : 0;
            poff[i] = static_cast<uint32_t>(pix);
            ok |= (in ? 1u : 0u) << i;
          }
          for (int cc = 0; cc < p.kb_per_tap; ++cc) {
            const int c = cc * 64;
            const bool first = c < d.c0;
            const uint8_t* srcb = static_cast<const uint8_t*>(first ? d.src0 : d.src1) +
                                  static_cast<size_t>((first ? c : c - d.c0) + chunk * 8) * 2;
            const uint32_t pitch = static_cast<uint32_t>(first ? p.ld0 : p.ld1) * 2u;   // bytes per pixel
            mbar_wait_acc(bar_empty + 8 * s, ph ^ 1u, prof, w_empty);
            const uint32_t a_stage = a_base + static_cast<uint32_t>(s) * kATileBytes;
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
              cp_async16(a_stage + row_off[i], srcb + poff[i] * pitch, ((ok >> i) & 1u) ? 16u : 0u);   // < 4 GiB (host-checked)
            }
            // hardware arrives on the stage's barrier when these copies have landed: the thread never waits
            // for its own loads, so up to `stages` k-blocks of gathers are in flight per thread
            cp_async_mbar_arrive_noinc(bar_full + 8 * s);
            ++it;
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        } else {
          // ---- modulated deformable sampling (models/dcn.py:59-66, torchvision deform_conv2d):
          //      value = mask * bilinear(x, ho*s - pad + ky + dy, wo*s - pad + kx + dx), zero outside
          //      (-1, H) x (-1, W); corners outside the image contribute 0 (weight 0, address clamped).
          //      Coefficients are per (row, tap): computed once and reused for every 64-channel block.
          const uint8_t* src = static_cast<const uint8_t*>(d.src0);
          const uint32_t pitch = static_cast<uint32_t>(p.ld0) * 2u;
          const float fh = static_cast<float>(d.h_in), fw = static_cast<float>(d.w_in);
          float wg4[kRows][4];
          uint32_t po[kRows][4];
#pragma unroll
          for (int i = 0; i < kRows; ++i) {
            const float* om = d.dcn_offmask + static_cast<size_t>(m_glob[i]) * 32;
            const float py = static_cast<float>(hy[i] + ky) + __ldg(om + 2 * tap);
            const float px = static_cast<float>(wx[i] + kx) + __ldg(om + 2 * tap + 1);
            const float mk = __ldg(om + 18 + tap);
            const bool inside = ((valid >> i) & 1u) && py > -1.f && py < fh && px > -1.f && px < fw;
            const float fy = floorf(py), fx = floorf(px);
            const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
            const float ly = py - fy, lx = px - fx;
#pragma unroll
            for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
              for (int cx = 0; cx < 2; ++cx) {
                const int yy = y0 + cy, xx = x0 + cx;
                const bool okc = inside && yy >= 0 && yy < d.h_in && xx >= 0 && xx < d.w_in;
                wg4[i][cy * 2 + cx] = okc ? (cy ? ly : 1.f - ly) * (cx ? lx : 1.f - lx) * mk : 0.f;
                const int yc = min(max(yy, 0), d.h_in - 1), xc = min(max(xx, 0), d.w_in - 1);
                po[i][cy * 2 + cx] = static_cast<uint32_t>(img_pix[i] + yc * d.w_in + xc) * pitch;
              }
            }
          }
          for (int cc = 0; cc < p.kb_per_tap; ++cc) {
            const uint8_t* srcb = src + static_cast<size_t>(cc * 64 + chunk * 8) * 2;
            uint4 q[kRows][4];
            if (p.dcn_dbg != 2) {
#pragma unroll
              for (int i = 0; i < kRows; ++i)
#pragma unroll
                for (int cnr = 0; cnr < 4; ++cnr) q[i][cnr] = __ldg(reinterpret_cast<const uint4*>(srcb + po[i][cnr]));
            } else {
#pragma unroll
              for (int i = 0; i < kRows; ++i)
#pragma unroll
                for (int cnr = 0; cnr < 4; ++cnr) q[i][cnr] = make_uint4(po[i][cnr], cc, tap, i);
            }
            mbar_wait_acc(bar_empty + 8 * s, ph ^ 1u, prof, w_empty);
            const uint32_t a_stage = a_base + static_cast<uint32_t>(s) * kATileBytes;
            if (p.dcn_dbg == 1) {
#pragma unroll
              for (int i = 0; i < kRows; ++i) {
                const uint4 v = make_uint4(q[i][0].x ^ q[i][1].x, q[i][0].y ^ q[i][2].y, q[i][0].z ^ q[i][3].z, q[i][0].w ^ q[i][1].w ^ q[i][2].w ^ q[i][3].w);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_stage + row_off[i]), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
            } else
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
              float acc[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
              for (int cnr = 0; cnr < 4; ++cnr) {
                const uint32_t w4[4] = {q[i][cnr].x, q[i][cnr].y, q[i][cnr].z, q[i][cnr].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 f = Pack2<T>::unpack(w4[j]);
                  acc[2 * j] = fmaf(wg4[i][cnr], f.x, acc[2 * j]);
                  acc[2 * j + 1] = fmaf(wg4[i][cnr], f.y, acc[2 * j + 1]);
                }
              }
              const uint32_t o0 = Pack2<T>::pack(acc[0], acc[1]), o1 = Pack2<T>::pack(acc[2], acc[3]);
              const uint32_t o2 = Pack2<T>::pack(acc[4], acc[5]), o3 = Pack2<T>::pack(acc[6], acc[7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_stage + row_off[i]), "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                           : "memory");
            }
            fence_proxy_async_smem();     // generic-proxy st.shared -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full + 8 * s);
            ++it;
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
    if (prof) { p.dbg[0] = clock64() - t_role0; p.dbg[1] = w_empty; p.dbg[2] = 0; p.dbg[3] = it; }
  } else if (wg == 3) {
    reg_dec<40>();
    if (warp == 12) {
      // =========================================================== B producer (TMA)
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        for (int tile = pt0; tile < p.total_ptiles; tile += pt_step) {
          const int nt = tile % p.n_tiles;
          const int g = (tile / p.n_tiles) / p.m_ptiles;
          const int row0 = g * d.w_group_rows + nt * p.n_tile;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);        // (cluster: both CTAs' MMAs have released the stage)
            mbar_arrive_expect_tx(bar_full + 8 * s, b_stage_bytes);
            if constexpr (kCl == 1) {
              tma_load_2d(b_base + static_cast<uint32_t>(s) * b_stage_bytes, &tmap_w, bar_full + 8 * s, kb * 64, row0);
            } else {
              // this CTA fetches its half of the weight rows and multicasts it into both CTAs' stage
              const uint32_t half = b_stage_bytes / 2;
              tma_load_2d_mcast(b_base + static_cast<uint32_t>(s) * b_stage_bytes + crank * half, &tmap_w, bar_full + 8 * s,
                                kb * 64, row0 + crank * (p.n_tile / 2), static_cast<uint16_t>(0x3));
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else if (warp == 13) {
      // =========================================================== MMA issuer
      if (lane == 0) {
        int s = 0;
        uint32_t ph = 0;
        int acc = 0;
        uint32_t acc_ph = 0;
        const bool prof = p.dbg != nullptr && blockIdx.x == 0;
        long long w_full = 0, w_tempty = 0;
        const long long t_role0 = prof ? clock64() : 0;
        for (int tile = pt0; tile < p.total_ptiles; tile += pt_step) {
          mbar_wait_acc(bar_tempty + 8 * acc, acc_ph ^ 1u, prof, w_tempty);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait_acc(bar_full + 8 * s, ph, prof, w_full);
            fence_proxy_async_smem();   // cp.async-written operand tile -> async proxy (belt and braces)
            tc_fence_after();
            const uint32_t a_addr = a_base + static_cast<uint32_t>(s) * kATileBytes;
            const uint32_t b_addr = b_base + static_cast<uint32_t>(s) * b_stage_bytes;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), p.idesc,
                       (kb | k) != 0 ? 1u : 0u);
            }
            if constexpr (kCl == 1) umma_commit(bar_empty + 8 * s);
            else umma_commit_mcast(bar_empty + 8 * s, static_cast<uint16_t>(0x3));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
          umma_commit(bar_tfull + 8 * acc);
          acc ^= 1;
          if (acc == 0) acc_ph ^= 1u;
        }
        if (prof) { p.dbg[4] = clock64() - t_role0; p.dbg[5] = w_full; p.dbg[6] = w_tempty; }
      }
    }
  } else if (is_epi) {
    // =========================================================== epilogue (WG0, and WG1 for plain convs)
    if constexpr (kDCN) reg_inc<168>(); else reg_inc<184>();
    const int q = warp & 3;         // TMEM lane quarter
    const int ge = wg;              // epilogue group: owns the 64-column groups with index % kEpiGroups == ge
    const int etid = threadIdx.x;   // 0 .. kEpiGroups*128-1
    int acc = 0;
    uint32_t acc_ph = 0;
    const bool avg4 = d.act == PRN_ACT_SIGMOID_AVG4;
    const bool has_res = d.residual != nullptr;
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long w_tfull = 0;
    const long long t_role0 = prof ? clock64() : 0;
    int bias_n0 = -1;
    const uint32_t stg_tile = stg_base + static_cast<uint32_t>(warp) * 4096u;   // warps 0-7
    for (int tile = pt0; tile < p.total_ptiles; tile += pt_step) {
      const int nt = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int mt = (rest % p.m_ptiles) * kCl + crank;
      const int g = rest / p.m_ptiles;
      const int n0 = nt * p.n_tile;
      const int n_valid = min(p.n_tile, d.n_pad - n0);   // multiple of 16
      const int m = mt * kTileM + q * 32 + lane;
      const bool valid = m < p.m_group;
      const int mm = valid ? m : 0;
      int n_local, pp;
      fast_divmod(mm, p.hw_out, p.inv_hw_out, n_local, pp);
      const int img = g * p.imgs_per_group + n_local;
      const size_t orow = avg4 ? static_cast<size_t>(img) * p.out_img_rows + (pp >> 2)
                               : static_cast<size_t>(img) * p.out_img_rows + pp;
      const size_t rrow = static_cast<size_t>(g) * p.m_group + mm;
      // (the shuffle must not sit behind `valid &&`: short-circuiting would leave the invalid lanes out of it)
      const int img_lane0 = __shfl_sync(0xffffffffu, img, 0);
      const bool img_uniform = __all_sync(0xffffffffu, valid && img == img_lane0);
      const T* res_row = has_res ? static_cast<const T*>(d.residual) + rrow * d.ld_res + n0 : nullptr;

      // bias of this tile's columns -> shared memory (overlaps the tile's MMAs); reloaded only when n0 changes
      if (d.bias != nullptr && n0 != bias_n0) {
        named_bar_sync(1, kEpiGroups * kEpiWarps * 32);          // everyone is done with the previous tile's values
        for (int i = etid; i < n_valid; i += kEpiGroups * kEpiWarps * 32) bias_s[i] = __ldg(d.bias + n0 + i);
        named_bar_sync(1, kEpiGroups * kEpiWarps * 32);
        bias_n0 = n0;
      }

      mbar_wait_acc(bar_tfull + 8 * acc, acc_ph, prof, w_tfull);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.n_tile);
      const int row0_out = mt * kTileM + q * 32;          // first output row of this warp's sub-tile

      // accumulator -> registers, software pipelined over this group's column chunks: the tcgen05.ld and the
      // residual loads of the next chunk are in flight while the current one is processed.  tcgen05.ld is
      // .sync.aligned: lanes diverged by per-row predicates must reconverge (__syncwarp) before each one.
      const int n32 = n_valid >> 5;
      const int ntot = n32 + ((kFull && (n_valid & 16)) ? 1 : 0);       // 32-wide chunks (+ one 16-wide tail)
      auto next_of = [&](int c) {
        ++c;
        while (c < ntot && ((c >> 1) % kEpiGroups) != ge) ++c;
        return c;
      };
      int ci = (0 % kEpiGroups) == ge ? 0 : next_of(0);
      uint32_t vb[32];
      uint4 rb[4];
      __syncwarp();
      if (ci < ntot) {
        if (ci < n32) {
          tmem_ld_x32(t_row + ci * 32, vb);
          load_res<32>(rb, res_row + ci * 32, has_res && valid);
        } else {
          tmem_ld_x16(t_row + ci * 32, vb);
          load_res<16>(rb, res_row + ci * 32, has_res && valid);
        }
      }
      while (ci < ntot) {
        tmem_ld_wait();
        tmem_ld_publish16(vb);
        tmem_ld_publish16(vb + 16);
        float x[32];
        uint4 rc[4];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(vb[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) rc[j] = rb[j];
        const int cn = next_of(ci);
        __syncwarp();
        if (cn < ntot) {
          if (cn < n32) {
            tmem_ld_x32(t_row + cn * 32, vb);
            load_res<32>(rb, res_row + cn * 32, has_res && valid);
          } else {
            tmem_ld_x16(t_row + cn * 32, vb);
            load_res<16>(rb, res_row + cn * 32, has_res && valid);
          }
        }
        if (p.tma_store && (ci & 1) == 0) {
          // the staging tile is about to be overwritten: its previous TMA store must have finished reading it
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
        const uint32_t srow = p.tma_store ? stg_tile + static_cast<uint32_t>(lane) * 128u : 0u;
        if (!kFull || ci < n32)
          epi_chunk<T, 32, kEpi>(p, x, rc, has_res, d.bias ? bias_s + ci * 32 : nullptr, n0 + ci * 32, valid, img_uniform,
                                  img, lane, orow, srow, (ci & 1) * 4);
        else
          epi_chunk<T, 16, kEpi>(p, x, rc, has_res, d.bias ? bias_s + ci * 32 : nullptr, n0 + ci * 32, valid, img_uniform,
                                  img, lane, orow, srow, (ci & 1) * 4);
        if (p.tma_store && ((ci & 1) == 1 || ci + 1 == ntot)) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_out, stg_tile, n0 + (ci >> 1) * 64, row0_out);
            bulk_commit();
          }
        }
        ci = cn;
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
    if (p.tma_store && lane == 0) bulk_wait<0>();   // all TMA stores of this warp have completed
    if (prof) { p.dbg[7] = clock64() - t_role0; p.dbg[8] = w_tfull; }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (kCl > 1) cluster_sync_all();   // the peer may still multicast into / arrive on this CTA's shared memory
  if (warp == 13) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// ------------------------------------------------------------------------------------------------ host
static bool cluster_enabled();
static int plan(const PrnConv& d, ConvKParams* p) {
  PRN_REQUIRE(d.src0 != nullptr && d.weight != nullptr, "conv: src0/weight must be non-NULL");
  PRN_REQUIRE(d.c0 > 0 && d.c0 % 64 == 0 && d.c1 >= 0 && d.c1 % 64 == 0, "conv: channel counts must be multiples of 64 (c0=%d c1=%d)", d.c0, d.c1);
  PRN_REQUIRE(d.c1 == 0 || d.src1 != nullptr, "conv: src1 is NULL but c1=%d", d.c1);
  PRN_REQUIRE(d.batch > 0 && d.h_in > 0 && d.w_in > 0 && d.h_out > 0 && d.w_out > 0, "conv: bad spatial dims");
  PRN_REQUIRE(d.upsample == 1 || d.upsample == 2, "conv: upsample must be 1 or 2");
  PRN_REQUIRE(d.ksize >= 1 && d.ksize <= 7 && d.stride >= 1 && d.pad >= 0, "conv: bad ksize/stride/pad");
  PRN_REQUIRE(d.pad_mode == PRN_PAD_ZERO || d.pad_mode == PRN_PAD_REFLECT || d.pad_mode == PRN_PAD_CLAMP, "conv: bad pad_mode");
  PRN_REQUIRE(d.shuffle_n == 0, "conv: pixel-shuffled outputs need the TMA halo kernel (3x3, stride 1, pad 1, PRN_CONV_TMA != 0)");
  PRN_REQUIRE(d.n_pad > 0 && d.n_pad % 16 == 0, "conv: n_pad must be a positive multiple of 16 (got %d)", d.n_pad);
  PRN_REQUIRE(d.dtype == PRN_BF16 || d.dtype == PRN_F16, "conv: dtype must be PRN_BF16 or PRN_F16");
  PRN_REQUIRE(d.out16 != nullptr || d.out32 != nullptr, "conv: no output buffer");
  PRN_REQUIRE(d.out16 == nullptr || d.ld_out16 % 8 == 0, "conv: ld_out16 must be a multiple of 8");
  PRN_REQUIRE(d.out32 == nullptr || d.ld_out32 % 4 == 0, "conv: ld_out32 must be a multiple of 4");
  PRN_REQUIRE(d.residual == nullptr || d.ld_res % 8 == 0, "conv: ld_res must be a multiple of 8");
  PRN_REQUIRE(d.stats == nullptr || d.stats_cg == 0 || d.stats_cg == 4 || d.stats_cg == 8 || d.stats_cg == 16,
              "conv: stats_cg must be 0, 4, 8 or 16");
  const int h_eff = d.h_in * d.upsample, w_eff = d.w_in * d.upsample;
  PRN_REQUIRE((h_eff + 2 * d.pad - d.ksize) / d.stride + 1 == d.h_out && (w_eff + 2 * d.pad - d.ksize) / d.stride + 1 == d.w_out,
              "conv: h_out/w_out inconsistent with input dims (%dx%d -> %dx%d)", h_eff, w_eff, d.h_out, d.w_out);
  PRN_REQUIRE(d.pad_mode != PRN_PAD_REFLECT || (d.pad < h_eff && d.pad < w_eff), "conv: reflect pad too large");
  if (d.dcn_offmask) {
    PRN_REQUIRE(d.c1 == 0 && d.upsample == 1 && d.pad_mode == PRN_PAD_ZERO && d.ksize == 3,
                "conv: deformable sampling needs a single source, 3x3, zero padding");
  }
  PRN_REQUIRE(d.ld0 == 0 || (d.ld0 >= d.c0 && d.ld0 % 8 == 0), "conv: bad ld0");
  PRN_REQUIRE(d.ld1 == 0 || (d.ld1 >= d.c1 && d.ld1 % 8 == 0), "conv: bad ld1");
  p->d = d;
  p->ld0 = d.ld0 ? d.ld0 : d.c0;
  p->ld1 = d.ld1 ? d.ld1 : d.c1;
  p->hw_out = d.h_out * d.w_out;
  if (d.w_group_rows != 0) {
    p->groups = d.batch;
    p->imgs_per_group = 1;
    PRN_REQUIRE(d.w_group_rows >= 0 && (long long)d.w_group_rows * (d.batch - 1) + d.n_pad <= (long long)d.w_rows_total + 256,
                "conv: grouped weights exceed w_rows_total");
  } else {
    p->groups = 1;
    p->imgs_per_group = d.batch;
  }
  p->m_group = p->imgs_per_group * p->hw_out;
  if (d.act == PRN_ACT_SIGMOID_AVG4) {
    PRN_REQUIRE(p->hw_out % 4 == 0, "conv: SIGMOID_AVG4 needs rows per image divisible by 4");
    p->out_img_rows = d.out_img_rows ? d.out_img_rows : p->hw_out / 4;
  } else {
    p->out_img_rows = d.out_img_rows ? d.out_img_rows : p->hw_out;
  }
  PRN_REQUIRE(p->m_group < (1 << 24), "conv: more than 2^24 output rows per group is not supported");
  PRN_REQUIRE(static_cast<unsigned long long>(d.batch) * d.h_in * d.w_in * (d.ld0 ? d.ld0 : d.c0) * 2ull < (1ull << 32) &&
                  static_cast<unsigned long long>(d.batch) * d.h_in * d.w_in * (d.ld1 ? d.ld1 : (d.c1 ? d.c1 : 1)) * 2ull < (1ull << 32),
              "conv: source tensors of 4 GiB or more are not supported (32-bit byte offsets in the gather)");
  p->inv_hw_out = 1.0f / static_cast<float>(p->hw_out);
  p->inv_w_out = 1.0f / static_cast<float>(d.w_out);
  // N tiling.  Candidates: the whole (padded) width if it fits the double-buffered TMEM, else 256/128/64.  Pick the
  // one with the smallest modelled time: waves x (k-blocks x max(MMA issue, L2->SM operand stream) + epilogue).
  // Narrow tiles give more CTAs work but re-gather the A tile once per N tile (9x more expensive for DCN).
  p->m_tiles = ceil_div(p->m_group, kTileM);
  const int sms = sm_count();
  const int kb = d.ksize * d.ksize * ((d.c0 + d.c1) / 64);
  int cands[4];
  int nc = 0;
  if (d.n_pad <= 256) cands[nc++] = d.n_pad;
  else cands[nc++] = (d.n_pad % 256 == 0) ? 256 : 128;
  for (int t = 128; t >= 64; t /= 2)
    if (t < cands[0] && d.n_pad % t == 0) cands[nc++] = t;
  int n_tile = cands[0];
  double best = 1e300;
  for (int i = 0; i < nc; ++i) {
    const int t = cands[i];
    const long long tiles = static_cast<long long>(p->groups) * p->m_tiles * ceil_div(d.n_pad, t);
    const double waves = static_cast<double>((tiles + sms - 1) / sms);
    const double a_cost = d.dcn_offmask ? 3400.0 : 16384.0 / 46.0;          // cycles to produce one A k-block
    const double per_kb = fmax(2.0 * t, a_cost + 128.0 * t / 46.0) + 60.0;
    const double tile_cost = kb * per_kb + 30.0 * t + 2500.0;
    const double cost = waves * tile_cost;
    if (cost < best) { best = cost; n_tile = t; }
  }
  p->n_tile = n_tile;
  p->n_tiles = ceil_div(d.n_pad, n_tile);
  p->total_tiles = p->groups * p->m_tiles * p->n_tiles;
  // CTA pairs sharing the weight stream pay off when the weight k-blocks dominate the L2->SM traffic
  p->cluster = (cluster_enabled() && !d.dcn_offmask && !(d.stats != nullptr && d.stats_cg == 0) && n_tile >= 128 && n_tile % 32 == 0 && p->m_tiles >= 2 && kb >= 4) ? 2 : 1;
  p->m_ptiles = ceil_div(p->m_tiles, p->cluster);
  p->total_ptiles = p->groups * p->m_ptiles * p->n_tiles;
  int cols = 32;
  while (cols < 2 * n_tile) cols *= 2;
  p->tmem_cols = cols;
  const int stage_bytes = kATileBytes + n_tile * 128;
  int stages = (kSmemBudget - 1024 - kCtrlBytes - kStageOutBytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  PRN_REQUIRE(stages >= 2, "conv: not enough shared memory for a 2-stage pipeline");
  p->stages = stages;
  p->kb_per_tap = (d.c0 + d.c1) / 64;
  p->num_kb = d.ksize * d.ksize * p->kb_per_tap;
  p->dbg = nullptr;
  static const int dcn_dbg = [] { const char* e = getenv("PRN_DCN_DEBUG"); return e ? atoi(e) : 0; }();
  p->dcn_dbg = dcn_dbg;
  // dense 16-bit output rows (row index == m) can go through the smem-staged TMA store
  p->tma_store = (d.out16 != nullptr && d.out32 == nullptr && d.act != PRN_ACT_SIGMOID_AVG4 && p->groups == 1 &&
                  p->out_img_rows == p->hw_out && (reinterpret_cast<uintptr_t>(d.out16) & 15) == 0 && d.ld_out16 % 8 == 0)
                     ? 1 : 0;
  p->lean_epi = (d.out16 != nullptr && d.out32 == nullptr && d.stats == nullptr &&
                 (d.act == PRN_ACT_NONE || d.act == PRN_ACT_RELU) && n_tile % 32 == 0 && d.n_pad % 32 == 0) ? 1 : 0;
  p->idesc = umma_idesc(d.dtype == PRN_BF16 ? 1u : 0u, kTileM, static_cast<uint32_t>(n_tile));
  return PRN_OK;
}

static bool cluster_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PRN_CLUSTER");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PRN_PDL");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v == 1;
}

template <typename T, bool kDCN, int kEpi, int kCl>
static int launch(const CUtensorMap& tm, const CUtensorMap& tmo, const ConvKParams& p, int grid, size_t smem,
                  cudaStream_t st) {
  static bool configured = false;  // per instantiation
  if (!configured) {
    PRN_CUDA(cudaFuncSetAttribute(conv_umma_kernel<T, kDCN, kEpi, kCl>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kCl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = kCl;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PRN_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<T, kDCN, kEpi, kCl>, tm, tmo, p));
  return PRN_OK;
}

}  // namespace prn

static int launch_grid(const prn::ConvKParams& p) {
  const int sms = prn::sm_count();
  return p.total_ptiles * p.cluster < sms ? p.total_ptiles * p.cluster : (sms / p.cluster) * p.cluster;
}

extern "C" int prn_conv2d_plan(const PrnConv* desc, int32_t* n_tile, int32_t* stages, int32_t* grid) {
  if (!desc) return prn::set_error(PRN_ERR_INVALID, "conv: NULL descriptor");
  prn::ConvKParams p;
  int rc = prn::plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  if (n_tile) *n_tile = p.n_tile;
  if (stages) *stages = p.stages;
  if (grid) *grid = launch_grid(p);
  return PRN_OK;
}

extern "C" int prn_conv2d_plan_ex(const PrnConv* desc, int32_t* out8) {
  if (!desc || !out8) return prn::set_error(PRN_ERR_INVALID, "conv: NULL argument");
  if (prn::conv_tma_eligible(*desc)) return prn::conv_tma_plan_ex(*desc, out8);
  prn::ConvKParams p;
  int rc = prn::plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  out8[0] = p.n_tile; out8[1] = p.stages; out8[2] = launch_grid(p); out8[3] = p.cluster;
  out8[4] = p.m_tiles; out8[5] = p.n_tiles; out8[6] = p.lean_epi; out8[7] = p.tma_store;
  return PRN_OK;
}

static int conv_launch(const PrnConv* desc, void* stream, long long* dbg);

extern "C" int prn_conv2d_fwd(const PrnConv* desc, void* stream) { return conv_launch(desc, stream, nullptr); }

extern "C" int prn_conv2d_fwd_profile(const PrnConv* desc, void* stream, int64_t* counters16) {
  return conv_launch(desc, stream, reinterpret_cast<long long*>(counters16));
}

static int conv_launch(const PrnConv* desc, void* stream, long long* dbg) {
  using namespace prn;
  if (!desc) return set_error(PRN_ERR_INVALID, "conv: NULL descriptor");
  if (conv_tma_eligible(*desc)) return conv_tma_launch(desc, stream, dbg);     // 3x3/s1/p1 and 1x1/s1: TMA-fed A operand
  ConvKParams p;
  int rc = plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  p.dbg = dbg;
  const PrnConv& d = p.d;
  CUtensorMap tm;
  const uint64_t kdim = static_cast<uint64_t>(d.ksize) * d.ksize * (d.c0 + d.c1);
  const uint64_t rows = d.w_rows_total > 0 ? static_cast<uint64_t>(d.w_rows_total) : static_cast<uint64_t>(d.n_pad);
  rc = encode_tmap_2d_sw128(&tm, d.weight, rows, kdim, static_cast<uint32_t>(p.n_tile / p.cluster), d.dtype);
  if (rc != PRN_OK) return rc;
  const int grid = launch_grid(p);
  CUtensorMap tmo = tm;
  if (p.tma_store) {
    rc = encode_tmap_2d_sw128(&tmo, d.out16, static_cast<uint64_t>(p.m_group), static_cast<uint64_t>(d.n_pad), 32, d.dtype,
                              static_cast<uint64_t>(d.ld_out16));
    if (rc != PRN_OK) return rc;
  }
  const size_t smem = 1024 + kCtrlBytes + kStageOutBytes + static_cast<size_t>(p.stages) * (kATileBytes + p.n_tile * 128);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool dcn = d.dcn_offmask != nullptr;
  const bool full = !p.lean_epi;
  const bool bnstats = d.stats != nullptr && d.stats_cg == 0;     // training step: separate instantiation keeps the
  if (bnstats) {                                                   // statistics code out of the inference kernels
    if (dcn) return set_error(PRN_ERR_UNSUPPORTED, "conv: BatchNorm statistics are not available with deformable sampling");
    if (d.dtype == PRN_BF16) return launch<__nv_bfloat16, false, 2, 1>(tm, tmo, p, grid, smem, st);
    return launch<__half, false, 2, 1>(tm, tmo, p, grid, smem, st);
  }
#define PRN_LAUNCH(T)                                                                                                  \
  (dcn ? (full ? launch<T, true, 1, 1>(tm, tmo, p, grid, smem, st) : launch<T, true, 0, 1>(tm, tmo, p, grid, smem, st)) \
       : (p.cluster == 2                                                                                               \
              ? (full ? launch<T, false, 1, 2>(tm, tmo, p, grid, smem, st) : launch<T, false, 0, 2>(tm, tmo, p, grid, smem, st)) \
              : (full ? launch<T, false, 1, 1>(tm, tmo, p, grid, smem, st) : launch<T, false, 0, 1>(tm, tmo, p, grid, smem, st))))
  if (d.dtype == PRN_BF16) return PRN_LAUNCH(__nv_bfloat16);
  return PRN_LAUNCH(__half);
#undef PRN_LAUNCH
}
