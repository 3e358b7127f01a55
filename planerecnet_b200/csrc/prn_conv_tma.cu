// prn_conv_tma.cu — implicit-GEMM convolution for sm_100a whose A operand arrives by TMA (no LSU gather).
//
//   3x3 / stride 1 / pad 1 ("halo mode"): one M tile = a 16 x 8 patch of output pixels of one image.  Per 64-channel
//       block ONE tiled 4-D TMA box {64 ch, 16 w, 18 h, 1 n} brings the patch's input halo (out-of-image pixels arrive
//       as zeros = the zero padding); the nine taps are then nine tcgen05.mma groups whose A descriptors start at halo
//       pixel (ky*16 + kx): an 8-pixel run of one halo row is one 8-row core-matrix group, consecutive output rows are
//       SBO = 2048 B (one halo row) apart.  The input is read from L2 once per 64-channel block instead of once per tap
//       (9x less A traffic than an im2col gather).  Reflection / replicate padding patch the halo's border rows and
//       columns in shared memory (one warp, border tiles only).
//   1x1 / stride 1 ("linear mode"): one M tile = 128 consecutive pixel rows = one 2-D TMA box per 64-channel block.
//
// Warp roles (512 threads, persistent CTAs, static round-robin tile schedule):
//   WG0-WG2 (warps 0-11)  epilogue: tcgen05.ld -> bias/residual/activation/statistics -> TMA stores through two 2 KB
//                          staging buffers per warp (a store drains while the next chunk is converted), or direct stores
//   warp 12                A producer (TMA), warp 14: B producer (TMA, one {64 k, n_tile} weight box per tap and block)
//   warp 13                MMA issuer + TMEM owner, warp 15: halo border patch (reflect / clamp padding)
//
// Replaces the same nn.Conv2d / F.conv2d call sites as prn_conv.cu (include/prn_b200.h); prn_conv2d_fwd dispatches here
// when the geometry qualifies (conv_tma_eligible) and falls back to the gather kernel otherwise.
#include "prn_conv_common.cuh"

namespace prn {

constexpr int kHaloW = 16;                         // halo pitch in pixels (10 used: 8 + 2)
constexpr int kHaloH = 18;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;  // 36864
constexpr int kPatchH = 16, kPatchW = 8;

struct TmaKParams {
  PrnConv d;
  int halo;               // 1: 3x3 halo mode, 0: 1x1 linear mode
  int groups, imgs_per_group, m_group;
  int tiles_x, tiles_y, tiles_per_img;
  int m_tiles, n_tiles, total_tiles;
  int n_tile, sa, sb, tmem_cols;
  int ncb, cb0, taps, ctot;
  uint32_t a_stage_bytes, b_stage_bytes, sbo;
  int fix;                // 0 none, 1 reflect, 2 clamp
  int hw_out, out_img_rows;
  float inv_hw_out;
  int lean_epi, tma_store;
  int base_off_mode;      // 1: descriptor base-offset field = (start address >> 7) & 7
  uint32_t idesc;
  long long* dbg;
};

constexpr int kTmaCtrlBytes = 2048;     // barriers (first 512 B), bias staging (+1024, 1 KB)
constexpr int kTmaStageOutBytes = 3 * 4 * 4096;   // per epilogue warp (up to 12) two 2 KB staging buffers

template <typename T, int kEpi>
__global__ void __launch_bounds__(kThreads, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                const __grid_constant__ TmaKParams p) {
  constexpr bool kFull = kEpi != 0;
  // epilogue warpgroups: three (WG0-2, 152 registers per thread) for the inference instantiations; the BatchNorm-statistics
  // instantiation of the training step needs 184 registers per thread and keeps two (WG2 idles)
  constexpr int kTmaEpiGroups = kEpi == 2 ? 2 : 3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_u32);

  const uint32_t bar_afull = base;               // [8]
  const uint32_t bar_aempty = base + 64;         // [8]
  const uint32_t bar_aready = base + 128;        // [8]  (after the border patch)
  const uint32_t bar_bfull = base + 192;         // [8]
  const uint32_t bar_bempty = base + 256;        // [8]
  const uint32_t bar_tfull = base + 320;         // [2]
  const uint32_t bar_tempty = base + 336;        // [2]
  const uint32_t tmem_slot = base + 352;
  float* bias_s = reinterpret_cast<float*>(base_ptr + 1024);
  const uint32_t stg_base = base + kTmaCtrlBytes;                 // 12 warps x 2 x 2 KB
  const uint32_t a_base = stg_base + kTmaStageOutBytes;           // 1024-aligned
  const uint32_t b_base = a_base + static_cast<uint32_t>(p.sa) * p.a_stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const PrnConv& d = p.d;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    if (p.cb0 < p.ncb) tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_w);
    if (p.tma_store) tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < 8; ++s) {
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
      mbar_init(bar_aready + 8 * s, 1);
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, kTmaEpiGroups * kEpiWarps * 32);
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
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(base_ptr + 352);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int t0 = static_cast<int>(blockIdx.x), t_step = static_cast<int>(gridDim.x);

  if (wg == 3) {
    reg_dec<56>();
    if (warp == 12 && lane == 0) {
      // =========================================================== A producer (TMA)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t0; tile < p.total_tiles; tile += t_step) {
        const int mt = tile / p.n_tiles;
        int c_w = 0, c_h = 0, c_n = 0, row0 = 0;
        if (p.halo) {
          const int img = mt / p.tiles_per_img, rem = mt - img * p.tiles_per_img;
          const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
          c_w = tx * kPatchW - d.pad; c_h = ty * kPatchH - d.pad; c_n = img;      // pad 1, or 2 (zero padding only)
        } else {
          const int g = mt / p.m_tiles, mm = mt - g * p.m_tiles;
          row0 = g * p.m_group + mm * kTileM;
        }
        for (int cb = 0; cb < p.ncb; ++cb) {
          const bool first = cb < p.cb0;
          const CUtensorMap* tm = first ? &tmap_a0 : &tmap_a1;
          const int cc = (first ? cb : cb - p.cb0) * 64;
          mbar_wait(bar_aempty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(bar_afull + 8 * s, p.a_stage_bytes);
          const uint32_t dst = a_base + static_cast<uint32_t>(s) * p.a_stage_bytes;
          if (p.halo) tma_load_4d(dst, tm, bar_afull + 8 * s, cc, c_w, c_h, c_n);
          else tma_load_2d(dst, tm, bar_afull + 8 * s, cc, row0);
          if (++s == p.sa) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 14 && lane == 0) {
      // =========================================================== B producer (TMA)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t0; tile < p.total_tiles; tile += t_step) {
        const int nt = tile % p.n_tiles;
        const int mt = tile / p.n_tiles;
        const int g = p.halo ? 0 : mt / p.m_tiles;
        const int wrow0 = g * d.w_group_rows + nt * p.n_tile;
        for (int cb = 0; cb < p.ncb; ++cb) {
          for (int tap = 0; tap < p.taps; ++tap) {
            mbar_wait(bar_bempty + 8 * s, ph ^ 1u);
            mbar_arrive_expect_tx(bar_bfull + 8 * s, p.b_stage_bytes);
            tma_load_2d(b_base + static_cast<uint32_t>(s) * p.b_stage_bytes, &tmap_w, bar_bfull + 8 * s,
                        tap * p.ctot + cb * 64, wrow0);
            if (++s == p.sb) { s = 0; ph ^= 1u; }
          }
        }
      }
    } else if (warp == 13 && lane == 0) {
      // =========================================================== MMA issuer
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int acc = 0;
      uint32_t acc_ph = 0;
      const bool prof = p.dbg != nullptr && blockIdx.x == 0;
      long long w_a = 0, w_b = 0, w_tempty = 0;
      const long long t_role0 = prof ? clock64() : 0;
      const uint32_t bar_a = p.fix ? bar_aready : bar_afull;
      for (int tile = t0; tile < p.total_tiles; tile += t_step) {
        mbar_wait_acc(bar_tempty + 8 * acc, acc_ph ^ 1u, prof, w_tempty);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.n_tile);
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait_acc(bar_a + 8 * sa, pha, prof, w_a);
          tc_fence_after();
          const uint32_t a_stage = a_base + static_cast<uint32_t>(sa) * p.a_stage_bytes;
          int ky = 0, kx = -1;
          for (int tap = 0; tap < p.taps; ++tap) {
            if (++kx == 3) { kx = 0; ++ky; }
            mbar_wait_acc(bar_bfull + 8 * sb, phb, prof, w_b);
            tc_fence_after();
            const uint32_t a_addr = a_stage + static_cast<uint32_t>(ky * kHaloW + kx) * 128u;
            const uint32_t b_addr = b_base + static_cast<uint32_t>(sb) * p.b_stage_bytes;
            const uint32_t bo = p.base_off_mode ? ((a_addr >> 7) & 7u) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d_tmem, umma_desc_sw128_ex(a_addr + k * 32, p.sbo, bo), umma_desc_sw128(b_addr + k * 32), p.idesc,
                       (cb | tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(bar_bempty + 8 * sb);
            if (++sb == p.sb) { sb = 0; phb ^= 1u; }
          }
          umma_commit(bar_aempty + 8 * sa);
          if (++sa == p.sa) { sa = 0; pha ^= 1u; }
        }
        umma_commit(bar_tfull + 8 * acc);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1u;
      }
      if (prof) { p.dbg[4] = clock64() - t_role0; p.dbg[5] = w_a; p.dbg[6] = w_tempty; p.dbg[9] = w_b; }
    } else if (warp == 15 && p.fix != 0) {
      // =========================================================== halo border patch (reflect / clamp padding)
      // The TMA box zero-fills pixels outside the image; nn.ReflectionPad2d(1) needs image row 1 at row -1 (row H-2 at
      // row H), replicate padding row 0 (row H-1); same for columns.  Rows first, then columns over all rows (corners).
      int s = 0;
      uint32_t ph = 0;
      const int back = p.fix == 1 ? 2 : 1;       // distance of the source row/column from the padded one
      for (int tile = t0; tile < p.total_tiles; tile += t_step) {
        const int mt = tile / p.n_tiles;
        const int img = mt / p.tiles_per_img, rem = mt - img * p.tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int hb = d.h_in - (ty * kPatchH - 1);      // halo row that holds image row H (first row below the image)
        const int wb = d.w_in - (tx * kPatchW - 1);      // halo column that holds image column W
        const bool top = ty == 0, left = tx == 0, bottom = hb <= kHaloH - 1, right = wb <= kPatchW + 1;
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(bar_afull + 8 * s, ph);
          if (top || left || bottom || right) {
            uint8_t* hp = base_ptr + (a_base - base) + static_cast<size_t>(s) * p.a_stage_bytes;
            auto copy_px = [&](int pd, int ps, int j) {
              const uint4 v = *reinterpret_cast<const uint4*>(hp + ps * 128 + ((j ^ (ps & 7)) << 4));
              *reinterpret_cast<uint4*>(hp + pd * 128 + ((j ^ (pd & 7)) << 4)) = v;
            };
            // rows: 10 used pixels x 8 chunks = 80 copies per row
            for (int i = lane; i < 80; i += 32) {
              const int px = i >> 3, j = i & 7;
              if (top) copy_px(0 * kHaloW + px, back * kHaloW + px, j);
              if (bottom) copy_px(hb * kHaloW + px, (hb - back) * kHaloW + px, j);
            }
            __syncwarp();
            // columns: 18 rows x 8 chunks = 144 copies per column
            for (int i = lane; i < kHaloH * 8; i += 32) {
              const int r = i >> 3, j = i & 7;
              if (left) copy_px(r * kHaloW + 0, r * kHaloW + back, j);
              if (right) copy_px(r * kHaloW + wb, r * kHaloW + wb - back, j);
            }
            fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
          }
          if (lane == 0) mbar_arrive(bar_aready + 8 * s);
          if (++s == p.sa) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (wg >= kTmaEpiGroups) {
    reg_dec<56>();
  } else {
    // =========================================================== epilogue (WG0 + WG1 [+ WG2])
    if constexpr (kTmaEpiGroups == 3) reg_inc<152>(); else reg_inc<184>();
    const int q = warp & 3;         // TMEM lane quarter
    const int ge = wg;              // epilogue group: owns the 32-column chunks with index % 3 == ge
    const int etid = threadIdx.x;   // 0 .. 383
    int acc = 0;
    uint32_t acc_ph = 0;
    const bool avg4 = d.act == PRN_ACT_SIGMOID_AVG4;
    const bool has_res = d.residual != nullptr;
    const bool has_stats = d.stats != nullptr;
    const bool prof = p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long w_tfull = 0;
    const long long t_role0 = prof ? clock64() : 0;
    int bias_n0 = -1;
    const uint32_t stg_tile = stg_base + static_cast<uint32_t>(warp) * 4096u;   // two 2 KB buffers (warps 0-11)
    uint32_t n_store = 0;           // chunks this warp has staged so far (selects the buffer)
    const int m_local = q * 32 + lane;
    for (int tile = t0; tile < p.total_tiles; tile += t_step) {
      const int nt = tile % p.n_tiles;
      const int mt = tile / p.n_tiles;
      const int n0 = nt * p.n_tile;
      const int n_valid = min(p.n_tile, d.n_pad - n0);   // multiple of 16
      bool valid;
      int img, pp, px = 0, py = 0, tx8 = 0, ty16 = 0, row0_out = 0;
      if (p.halo) {
        img = mt / p.tiles_per_img;
        const int rem = mt - img * p.tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        tx8 = tx * kPatchW; ty16 = ty * kPatchH;
        py = ty16 + (m_local >> 3); px = tx8 + (m_local & 7);
        valid = py < d.h_out && px < d.w_out;
        pp = valid ? py * d.w_out + px : 0;
      } else {
        const int g = mt / p.m_tiles, mm = mt - g * p.m_tiles;
        const int m = mm * kTileM + m_local;
        valid = m < p.m_group;
        int n_local;
        fast_divmod(valid ? m : 0, p.hw_out, p.inv_hw_out, n_local, pp);
        img = g * p.imgs_per_group + n_local;
        row0_out = mm * kTileM + q * 32;
      }
      const size_t orow = static_cast<size_t>(img) * p.out_img_rows + (avg4 ? (pp >> 2) : pp);
      const size_t rrow = static_cast<size_t>(img) * p.hw_out + pp;
      const int img_lane0 = __shfl_sync(0xffffffffu, img, 0);
      const bool all_valid_same = __all_sync(0xffffffffu, valid && img == img_lane0);
      // halo tiles never straddle images: with the invalid rows zeroed below, the warp-wide statistics path applies
      const bool img_uniform = p.halo ? true : all_valid_same;
      const T* res_row = has_res ? static_cast<const T*>(d.residual) + rrow * d.ld_res + n0 : nullptr;

      if (d.bias != nullptr && n0 != bias_n0) {
        named_bar_sync(1, kTmaEpiGroups * kEpiWarps * 32);
        for (int i = etid; i < n_valid; i += kTmaEpiGroups * kEpiWarps * 32) bias_s[i] = __ldg(d.bias + n0 + i);
        named_bar_sync(1, kTmaEpiGroups * kEpiWarps * 32);
        bias_n0 = n0;
      }

      mbar_wait_acc(bar_tfull + 8 * acc, acc_ph, prof, w_tfull);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * p.n_tile);

      const int n32 = n_valid >> 5;
      const int ntot = n32 + ((kFull && (n_valid & 16)) ? 1 : 0);
      // chunk c belongs to group c % kTmaEpiGroups (three groups, N = 256: 3 + 3 + 2 chunks)
      auto next_of = [&](int c) { return c + kTmaEpiGroups; };
      int ci = ge;
      uint32_t vb[32];
      __syncwarp();
      if (ci < ntot) {
        if (ci < n32) tmem_ld_x32(t_row + ci * 32, vb);
        else tmem_ld_x16(t_row + ci * 32, vb);
      }
      while (ci < ntot) {
        // residual of THIS chunk: issued first, consumed after the accumulator has been unpacked (three epilogue warps per
        // scheduler cover the latency; prefetching it one chunk ahead would cost 16 more registers per thread)
        uint4 rc[4];
        if (ci < n32) load_res<32>(rc, res_row + ci * 32, has_res && valid);
        else load_res<16>(rc, res_row + ci * 32, has_res && valid);
        tmem_ld_wait();
        tmem_ld_publish16(vb);
        tmem_ld_publish16(vb + 16);
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(vb[j]);
        const int cn = next_of(ci);
        __syncwarp();
        if (cn < ntot) {
          if (cn < n32) tmem_ld_x32(t_row + cn * 32, vb);
          else tmem_ld_x16(t_row + cn * 32, vb);
        }
        if (kFull && p.halo && has_stats && d.stats_cg > 0 && !valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.f;     // rows outside the image must not enter the statistics
        }
        uint32_t srow = 0u;
        if (p.tma_store) {
          // two staging buffers per warp: the store of the chunk before last must have finished READING its buffer
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
          srow = stg_tile + (n_store & 1u) * 2048u + static_cast<uint32_t>(lane) * 64u;
        }
        const int col0 = n0 + ci * 32;
        size_t orow_c = orow;
        int scol = -1;
        if (d.shuffle_n > 0) {
          // sub-pixel phases: columns [phase*shuffle_n, +shuffle_n) of pixel (y,x) -> pixel (2y+a, 2x+b) of the x2 output
          const int phase = col0 / d.shuffle_n;
          scol = col0 - phase * d.shuffle_n;
          orow_c = static_cast<size_t>(img) * p.out_img_rows +
                   static_cast<size_t>(2 * py + (phase >> 1)) * (2 * d.w_out) + (2 * px + (phase & 1));
        }
        if (!kFull || ci < n32)
          epi_chunk<T, 32, kEpi>(p, x, rc, has_res, d.bias ? bias_s + ci * 32 : nullptr, col0, valid, img_uniform, img, lane,
                                  orow_c, srow, -1, scol);
        else
          epi_chunk<T, 16, kEpi>(p, x, rc, has_res, d.bias ? bias_s + ci * 32 : nullptr, col0, valid, img_uniform, img, lane,
                                  orow_c, srow, -1, scol);
        if (p.tma_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            const uint32_t src = stg_tile + (n_store & 1u) * 2048u;
            if (p.halo) tma_store_4d(&tmap_out, src, col0, tx8, ty16 + 4 * q, img);
            else tma_store_2d(&tmap_out, src, col0, row0_out);
            bulk_commit();
          }
          ++n_store;
        }
        ci = cn;
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * acc);
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1u;
    }
    if (p.tma_store && lane == 0) bulk_wait<0>();
    if (prof) { p.dbg[7] = clock64() - t_role0; p.dbg[8] = w_tfull;  }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 13) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
}

// ------------------------------------------------------------------------------------------------ host
static int tma_mode() {           // PRN_CONV_TMA: 0 = off, 1 = on (default); PRN_CONV_TMA_BASEOFF: 0 (default) / 1
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PRN_CONV_TMA");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return v;
}
static int tma_base_off() {
  static int v = -1;
  if (v < 0) {
    // measured (tools/probe/umma_probe.cu, profiles/r02_probe.txt): the tensor core swizzles on absolute shared-memory
    // address bits, so the base-offset field must stay 0 even when an operand starts off a 1024-byte boundary
    const char* e = getenv("PRN_CONV_TMA_BASEOFF");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v;
}

bool conv_tma_eligible(const PrnConv& d) {
  if (!tma_mode()) return false;
  if (d.dcn_offmask != nullptr || d.upsample != 1 || d.stride != 1) return false;
  if (d.dtype != PRN_BF16 && d.dtype != PRN_F16) return false;
  // pad 2 with zero padding = the "full" correlation the input gradient of a reflection-padded 3x3 conv needs (train_engine.py)
  const bool k3 = d.ksize == 3 && (d.pad == 1 || (d.pad == 2 && d.pad_mode == PRN_PAD_ZERO && d.shuffle_n == 0));
  const bool k1 = d.ksize == 1 && d.pad == 0;
  if (!k3 && !k1) return false;
  if (k1 && d.pad_mode != PRN_PAD_ZERO) return false;
  if (k3 && d.w_group_rows != 0) return false;
  if (d.stats != nullptr && d.stats_cg == 0 && d.shuffle_n > 0) return false;
  if (k3 && (d.h_in < 2 || d.w_in < 2)) return false;
  // TMA needs 16-byte aligned bases and pitches
  const int ld0 = d.ld0 ? d.ld0 : d.c0, ld1 = d.ld1 ? d.ld1 : d.c1;
  if ((reinterpret_cast<uintptr_t>(d.src0) & 15) || (ld0 % 8)) return false;
  if (d.c1 > 0 && ((reinterpret_cast<uintptr_t>(d.src1) & 15) || (ld1 % 8))) return false;
  return true;
}

static int tma_plan(const PrnConv& d, TmaKParams* p) {
  PRN_REQUIRE(d.src0 != nullptr && d.weight != nullptr, "conv: src0/weight must be non-NULL");
  PRN_REQUIRE(d.c0 > 0 && d.c0 % 64 == 0 && d.c1 >= 0 && d.c1 % 64 == 0, "conv: channel counts must be multiples of 64 (c0=%d c1=%d)", d.c0, d.c1);
  PRN_REQUIRE(d.c1 == 0 || d.src1 != nullptr, "conv: src1 is NULL but c1=%d", d.c1);
  PRN_REQUIRE(d.batch > 0 && d.h_in > 0 && d.w_in > 0, "conv: bad spatial dims");
  {
    const int grow = d.ksize == 3 ? 2 * d.pad - 2 : 0;
    PRN_REQUIRE(d.h_out == d.h_in + grow && d.w_out == d.w_in + grow, "conv: h_out/w_out inconsistent with input dims");
  }
  PRN_REQUIRE(d.n_pad > 0 && d.n_pad % 16 == 0, "conv: n_pad must be a positive multiple of 16 (got %d)", d.n_pad);
  PRN_REQUIRE(d.out16 != nullptr || d.out32 != nullptr, "conv: no output buffer");
  PRN_REQUIRE(d.out16 == nullptr || d.ld_out16 % 8 == 0, "conv: ld_out16 must be a multiple of 8");
  PRN_REQUIRE(d.out32 == nullptr || d.ld_out32 % 4 == 0, "conv: ld_out32 must be a multiple of 4");
  PRN_REQUIRE(d.residual == nullptr || d.ld_res % 8 == 0, "conv: ld_res must be a multiple of 8");
  PRN_REQUIRE(d.stats == nullptr || d.stats_cg == 0 || d.stats_cg == 4 || d.stats_cg == 8 || d.stats_cg == 16,
              "conv: stats_cg must be 0, 4, 8 or 16");
  PRN_REQUIRE(d.pad_mode == PRN_PAD_ZERO || d.pad_mode == PRN_PAD_REFLECT || d.pad_mode == PRN_PAD_CLAMP, "conv: bad pad_mode");
  PRN_REQUIRE(d.shuffle_n == 0 || (d.ksize == 3 && d.shuffle_n % 32 == 0 && d.n_pad == 4 * d.shuffle_n && d.out16 != nullptr &&
                                   d.out32 == nullptr && d.residual == nullptr && d.act != PRN_ACT_SIGMOID_AVG4),
              "conv: shuffle_n needs a 3x3 conv with n_pad == 4*shuffle_n (multiple of 32), a 16-bit output and no residual");
  p->d = d;
  p->halo = d.ksize == 3 ? 1 : 0;
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
    PRN_REQUIRE(!p->halo && p->hw_out % 4 == 0, "conv: SIGMOID_AVG4 needs a 1x1 conv and rows per image divisible by 4");
    p->out_img_rows = d.out_img_rows ? d.out_img_rows : p->hw_out / 4;
  } else if (d.shuffle_n > 0) {
    p->out_img_rows = d.out_img_rows ? d.out_img_rows : 4 * p->hw_out;
  } else {
    p->out_img_rows = d.out_img_rows ? d.out_img_rows : p->hw_out;
  }
  PRN_REQUIRE(p->m_group < (1 << 24), "conv: more than 2^24 output rows per group is not supported");
  p->inv_hw_out = 1.0f / static_cast<float>(p->hw_out);
  p->ctot = d.c0 + d.c1;
  p->ncb = p->ctot / 64;
  p->cb0 = d.c0 / 64;
  p->taps = p->halo ? 9 : 1;
  p->a_stage_bytes = p->halo ? kHaloBytes : kATileBytes;
  p->sbo = p->halo ? kHaloW * 128 : 1024;
  p->fix = !p->halo ? 0 : (d.pad_mode == PRN_PAD_REFLECT ? 1 : (d.pad_mode == PRN_PAD_CLAMP ? 2 : 0));
  if (p->halo) {
    p->tiles_x = ceil_div(d.w_out, kPatchW);
    p->tiles_y = ceil_div(d.h_out, kPatchH);
    p->tiles_per_img = p->tiles_x * p->tiles_y;
    p->m_tiles = d.batch * p->tiles_per_img;       // all images
  } else {
    p->tiles_x = p->tiles_y = p->tiles_per_img = 0;
    p->m_tiles = ceil_div(p->m_group, kTileM);      // per group
  }
  const int m_tiles_all = p->halo ? p->m_tiles : p->groups * p->m_tiles;
  // N tiling: the whole padded width when it fits the double-buffered TMEM, else 256 / 128; narrower tiles only when that
  // fills more SMs of a sub-wave launch (the A operand is cheap here, the weights are re-streamed per M tile either way)
  const int sms = sm_count();
  int n_tile = d.n_pad <= 256 ? d.n_pad : ((d.n_pad % 256 == 0 || d.n_pad > 1024) ? 256 : 128);
  {
    // Empirical model (tuned on tools/conv_probe.py): waves x (k-blocks x max(MMA, operand stream) + epilogue).  A narrower tile
    // has to win by 15 %: near ties go to the wide tile, whose MMAs run at the better rate (161 cycles per 256 columns against
    // 116 per <= 128, §3.0 of DESIGN.md) — e.g. the 60x80 FPN conv: 320 tiles at N = 256 (46.8 us) against 640 at N = 128 (72.5 us).
    const int kb = p->ncb * p->taps;
    double best = 1e300;
    int best_t = n_tile;
    for (int t = n_tile; t >= 64; t /= 2) {
      if (d.n_pad % t != 0 && t != n_tile) continue;
      const long long tiles = static_cast<long long>(m_tiles_all) * ceil_div(d.n_pad, t);
      const double waves = static_cast<double>((tiles + sms - 1) / sms);
      const double a_bytes = static_cast<double>(p->a_stage_bytes) / p->taps;
      const double per_kb = fmax(2.0 * t, (a_bytes + 128.0 * t) / 60.0) + 20.0;
      const double cost = waves * (kb * per_kb + 12.0 * t + 2500.0);
      if (cost < best * (t == n_tile ? 1.0 : 0.85)) { best = cost; best_t = t; }
      if (t % 2 != 0 || (t / 2) % 16 != 0) break;
    }
    n_tile = best_t;
  }
  p->n_tile = n_tile;
  p->n_tiles = ceil_div(d.n_pad, n_tile);
  p->total_tiles = m_tiles_all * p->n_tiles;
  int cols = 32;
  while (cols < 2 * n_tile) cols *= 2;
  p->tmem_cols = cols;
  p->b_stage_bytes = static_cast<uint32_t>(n_tile) * 128u;
  // shared memory: control + store staging + A ring + B ring
  const int avail = kSmemBudget - 1024 - kTmaCtrlBytes - kTmaStageOutBytes;
  int sa = p->halo ? 2 : 4;
  int sb = (avail - sa * static_cast<int>(p->a_stage_bytes)) / static_cast<int>(p->b_stage_bytes);
  if (p->halo && sb > 8) {           // small weight tiles: a third halo stage instead of more than 8 weight stages
    sa = 3;
    sb = (avail - sa * static_cast<int>(p->a_stage_bytes)) / static_cast<int>(p->b_stage_bytes);
  }
  if (!p->halo) {
    // linear mode consumes one A and one B stage per k-block: balance the rings
    while (sa < 8 && sb > sa + 1) {
      ++sa;
      sb = (avail - sa * static_cast<int>(p->a_stage_bytes)) / static_cast<int>(p->b_stage_bytes);
    }
  }
  if (sb > 8) sb = 8;
  PRN_REQUIRE(sb >= 2, "conv: not enough shared memory for a 2-stage weight pipeline");
  p->sa = sa;
  p->sb = sb;
  p->dbg = nullptr;
  const bool dense16 = d.out16 != nullptr && d.out32 == nullptr && d.act != PRN_ACT_SIGMOID_AVG4 && d.shuffle_n == 0 &&
                       (reinterpret_cast<uintptr_t>(d.out16) & 15) == 0 && d.ld_out16 % 8 == 0;
  static const bool store_off = [] { const char* e = getenv("PRN_TMA_STORE"); return e != nullptr && e[0] == '0'; }();
  p->tma_store = (!store_off && dense16 && (p->halo ? true : (p->groups == 1 && p->out_img_rows == p->hw_out))) ? 1 : 0;
  p->lean_epi = (d.out16 != nullptr && d.out32 == nullptr && d.stats == nullptr &&
                 (d.act == PRN_ACT_NONE || d.act == PRN_ACT_RELU) && n_tile % 32 == 0 && d.n_pad % 32 == 0) ? 1 : 0;
  p->base_off_mode = tma_base_off();
  p->idesc = umma_idesc(d.dtype == PRN_BF16 ? 1u : 0u, kTileM, static_cast<uint32_t>(n_tile));
  return PRN_OK;
}

template <typename T, int kEpi>
static int tma_launch(const CUtensorMap& ta0, const CUtensorMap& ta1, const CUtensorMap& tw, const CUtensorMap& to,
                      const TmaKParams& p, int grid, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PRN_CUDA(cudaFuncSetAttribute(conv_tma_kernel<T, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  const char* e = getenv("PRN_PDL");
  if (e == nullptr || e[0] != '0') {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PRN_CUDA(cudaLaunchKernelEx(&cfg, conv_tma_kernel<T, kEpi>, ta0, ta1, tw, to, p));
  return PRN_OK;
}

int conv_tma_plan_ex(const PrnConv& d, int32_t* out8) {
  TmaKParams p;
  int rc = tma_plan(d, &p);
  if (rc != PRN_OK) return rc;
  const int sms = sm_count();
  out8[0] = p.n_tile; out8[1] = p.sa * 16 + p.sb; out8[2] = p.total_tiles < sms ? p.total_tiles : sms; out8[3] = 1;
  out8[4] = p.halo ? p.m_tiles : p.groups * p.m_tiles; out8[5] = p.n_tiles; out8[6] = p.lean_epi; out8[7] = p.tma_store + 2;
  return PRN_OK;
}

int conv_tma_launch(const PrnConv* desc, void* stream, long long* dbg) {
  TmaKParams p;
  int rc = tma_plan(*desc, &p);
  if (rc != PRN_OK) return rc;
  p.dbg = dbg;
  const PrnConv& d = p.d;
  const int ld0 = d.ld0 ? d.ld0 : d.c0, ld1 = d.ld1 ? d.ld1 : d.c1;
  CUtensorMap ta0, ta1, tw, to;
  // ---- A operand maps
  for (int s = 0; s < 2; ++s) {
    CUtensorMap* tm = s == 0 ? &ta0 : &ta1;
    if (s == 1 && d.c1 == 0) { ta1 = ta0; break; }
    const void* src = s == 0 ? d.src0 : d.src1;
    const uint64_t cc = s == 0 ? d.c0 : d.c1, ld = s == 0 ? ld0 : ld1;
    if (p.halo) {
      const uint64_t dims[4] = {cc, static_cast<uint64_t>(d.w_in), static_cast<uint64_t>(d.h_in), static_cast<uint64_t>(d.batch)};
      const uint64_t strides[3] = {ld * 2, static_cast<uint64_t>(d.w_in) * ld * 2, static_cast<uint64_t>(d.h_in) * d.w_in * ld * 2};
      const uint32_t box[4] = {64, kHaloW, kHaloH, 1};
      rc = encode_tmap_nd(tm, src, 4, dims, strides, box, 128, d.dtype);
    } else {
      const uint64_t dims[2] = {cc, static_cast<uint64_t>(d.batch) * p.hw_out};
      const uint64_t strides[1] = {ld * 2};
      const uint32_t box[2] = {64, kTileM};
      rc = encode_tmap_nd(tm, src, 2, dims, strides, box, 128, d.dtype);
    }
    if (rc != PRN_OK) return rc;
  }
  // ---- weights
  {
    const uint64_t kdim = static_cast<uint64_t>(p.taps) * p.ctot;
    const uint64_t rows = d.w_rows_total > 0 ? static_cast<uint64_t>(d.w_rows_total) : static_cast<uint64_t>(d.n_pad);
    rc = encode_tmap_2d_sw128(&tw, d.weight, rows, kdim, static_cast<uint32_t>(p.n_tile), d.dtype);
    if (rc != PRN_OK) return rc;
  }
  // ---- output (32-column chunks, 64-byte swizzle)
  to = tw;
  if (p.tma_store) {
    if (p.halo) {
      const uint64_t dims[4] = {static_cast<uint64_t>(d.n_pad), static_cast<uint64_t>(d.w_out), static_cast<uint64_t>(d.h_out),
                                static_cast<uint64_t>(d.batch)};
      const uint64_t strides[3] = {static_cast<uint64_t>(d.ld_out16) * 2, static_cast<uint64_t>(d.w_out) * d.ld_out16 * 2,
                                   static_cast<uint64_t>(p.out_img_rows) * d.ld_out16 * 2};
      const uint32_t box[4] = {32, kPatchW, 4, 1};
      rc = encode_tmap_nd(&to, d.out16, 4, dims, strides, box, 64, d.dtype);
    } else {
      const uint64_t dims[2] = {static_cast<uint64_t>(d.n_pad), static_cast<uint64_t>(p.m_group)};
      const uint64_t strides[1] = {static_cast<uint64_t>(d.ld_out16) * 2};
      const uint32_t box[2] = {32, 32};
      rc = encode_tmap_nd(&to, d.out16, 2, dims, strides, box, 64, d.dtype);
    }
    if (rc != PRN_OK) return rc;
  }
  const int sms = sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  const size_t smem = 1024 + kTmaCtrlBytes + kTmaStageOutBytes + static_cast<size_t>(p.sa) * p.a_stage_bytes +
                      static_cast<size_t>(p.sb) * p.b_stage_bytes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bnstats = d.stats != nullptr && d.stats_cg == 0;
  const int epi = bnstats ? 2 : (p.lean_epi ? 0 : 1);
#define PRN_TMA_LAUNCH(T)                                                                      \
  (epi == 2 ? tma_launch<T, 2>(ta0, ta1, tw, to, p, grid, smem, st)                            \
            : (epi == 1 ? tma_launch<T, 1>(ta0, ta1, tw, to, p, grid, smem, st) : tma_launch<T, 0>(ta0, ta1, tw, to, p, grid, smem, st)))
  if (d.dtype == PRN_BF16) return PRN_TMA_LAUNCH(__nv_bfloat16);
  return PRN_TMA_LAUNCH(__half);
#undef PRN_TMA_LAUNCH
}

}  // namespace prn
