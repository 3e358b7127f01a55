// umma_probe.cu — hardware facts the conv kernels are designed around (run on a B200; prints a report).
//
//  T1  tcgen05.mma with a K-major SWIZZLE_128B A operand whose 8-row groups do NOT start on a 1024-byte boundary
//      (descriptor start = base + shift*128 B, SBO in {1024, 1280, 2048, 2304}, base-offset field 0 or (addr>>7)&7):
//      which combinations read "pixel rows" p = shift + g*(SBO/128) + r of an address-swizzled pixel array correctly?
//  T2  TMA tiled 4-D box {64 ch, 16 w, 18 h, 1 n} with negative start coordinates: shared-memory layout + zero fill.
//  T3  3x3 convolution of one 16x8 output patch from that halo box: 9 taps x 4 MMAs with tap-shifted descriptors.
//  T4  SM ingest bandwidth: TMA box loads vs the cp.async (LDGSTS) gather pattern, no MMA, full chip / half chip.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I planerecnet_b200/csrc -o tools/probe/umma_probe tools/probe/umma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "prn_ptx.cuh"

using namespace prn;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);          \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  return reinterpret_cast<EncodeTiledFn>(p);
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ---------------------------------------------------------------------------------------------- T1
// a_pix: [npix][64] f16 "pixel rows"; b: [64][64] f16; out: [128][64] fp32
__global__ void __launch_bounds__(128, 1)
t1_kernel(const __half* a_pix, int npix, const __half* b, float* out, int shift, int sbo, int use_base_off) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_u = smem_u32(raw);
  const uint32_t base = (raw_u + 1023u) & ~1023u;
  uint8_t* bp = raw + (base - raw_u);
  const uint32_t a_s = base;                       // npix * 128 B
  const uint32_t b_s = base + 64 * 1024;           // 8 KB
  const uint32_t bar = base + 72 * 1024;
  const uint32_t slot = bar + 16;
  const int tid = threadIdx.x;
  for (int i = tid; i < npix * 8; i += 128) {
    const int p = i >> 3, j = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(a_pix)[p * 8 + j];
    *reinterpret_cast<uint4*>(bp + p * 128 + ((j ^ (p & 7)) << 4)) = v;
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int n = i >> 3, j = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(b)[n * 8 + j];
    *reinterpret_cast<uint4*>(bp + 64 * 1024 + n * 128 + ((j ^ (n & 7)) << 4)) = v;
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (tid < 32) { tmem_alloc(slot, 64); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bp + 72 * 1024 + 16);
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(0, 128, 64);
    const uint32_t a0 = a_s + shift * 128;
    const uint32_t bo = use_base_off ? ((a0 >> 7) & 7) : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem, desc_sw128(a0 + k * 32, sbo, bo), desc_sw128(b_s + k * 32, 1024, 0), idesc, k ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t v[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_x32(tmem + ((warp * 32u) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------------------- T2 / T3
// halo box {64, 16, 18, 1} at (c0, w0, h0, n0) -> dump (T2); then 3x3 conv of the 16x8 patch with w [64][9*64] (T3)
__global__ void __launch_bounds__(128, 1)
t23_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, uint8_t* dump, float* out,
           int c0, int w0, int h0, int n0, int use_base_off) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_u = smem_u32(raw);
  const uint32_t base = (raw_u + 1023u) & ~1023u;
  uint8_t* bp = raw + (base - raw_u);
  const uint32_t a_s = base;                       // 18*16*128 = 36864
  const uint32_t b_s = base + 40 * 1024;           // 9 taps x 8 KB = 72 KB
  const uint32_t bar = base + 120 * 1024;
  const uint32_t bar2 = bar + 8;
  const uint32_t slot = bar + 16;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); mbar_fence_init(); }
  if (tid < 32) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bp + 120 * 1024 + 16);
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, 36864 + 9 * 8192);
    tma_load_4d(a_s, &tm_x, bar, c0, w0, h0, n0);
    for (int t = 0; t < 9; ++t) tma_load_2d(b_s + t * 8192, &tm_w, bar, t * 64, 0);
  }
  mbar_wait(bar, 0);
  for (int i = tid; i < 36864 / 16; i += 128) reinterpret_cast<uint4*>(dump)[i] = reinterpret_cast<const uint4*>(bp)[i];
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc(0, 128, 64);
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      const uint32_t a0 = a_s + (ky * 16 + kx) * 128;
      const uint32_t bo = use_base_off ? ((a0 >> 7) & 7) : 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem, desc_sw128(a0 + k * 32, 2048, bo), desc_sw128(b_s + t * 8192 + k * 32, 1024, 0), idesc, (t | k) ? 1u : 0u);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t v[32];
  for (int c = 0; c < 2; ++c) {
    tmem_ld_x32(tmem + ((warp * 32u) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------------------- T4
// mode 0: TMA 2-D boxes {64, rows} (rows*128 B per load); mode 1: cp.async gather, `nthreads_prod` producer threads,
// 16 KB stages (128 rows x 128 B), row r of a stage comes from src row (row0 + r*row_stride) (im2col-like scatter)
struct T4Params {
  int mode, iters, stages, box_rows, nprod, region_rows, same_region, row_stride;
  const uint8_t* src;
  long long* cycles;   // per CTA
};
__global__ void __launch_bounds__(320, 1)
t4_kernel(const __grid_constant__ CUtensorMap tm, const T4Params p) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_u = smem_u32(raw);
  const uint32_t base = (raw_u + 1023u) & ~1023u;
  const uint32_t bar_full = base, bar_empty = base + 128;
  const uint32_t data = base + 1024;
  const uint32_t stage_bytes = p.box_rows * 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, p.mode == 0 ? 1 : p.nprod);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const long long region0 = p.same_region ? 0 : static_cast<long long>(blockIdx.x) * p.region_rows;
  const long long t0 = clock64();
  if (warp == 9) {            // consumer: frees the stage as soon as it is full
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < p.iters; ++it) {
        mbar_wait(bar_full + 8 * s, ph);
        mbar_arrive(bar_empty + 8 * s);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      p.cycles[blockIdx.x] = clock64() - t0;
    }
  } else if (p.mode == 0) {
    if (warp == 8 && lane == 0) {
      int s = 0; uint32_t ph = 0;
      int row = 0;
      for (int it = 0; it < p.iters; ++it) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        mbar_arrive_expect_tx(bar_full + 8 * s, stage_bytes);
        tma_load_2d(data + s * stage_bytes, &tm, bar_full + 8 * s, 0, static_cast<int>(region0 + row));
        row += p.box_rows;
        if (row + p.box_rows > p.region_rows) row = 0;
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (tid < p.nprod) {
    const int rows_per_thread = 128 * 8 / p.nprod;      // 16-byte pieces per thread per stage
    int s = 0; uint32_t ph = 0;
    int row = 0;
    const int chunk = tid & 7, sub = tid >> 3;
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      for (int i = 0; i < rows_per_thread; ++i) {
        const int r = i * (p.nprod / 8) + sub;
        const long long srow = region0 + (row + static_cast<long long>(r) * p.row_stride) % p.region_rows;
        cp_async16(data + s * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4), p.src + srow * 128 + chunk * 16, 16);
      }
      cp_async_mbar_arrive_noinc(bar_full + 8 * s);
      row += 128;
      if (row + 128 > p.region_rows) row = 0;
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  }
}

// ---------------------------------------------------------------------------------------------- T5
// tcgen05.mma issue/execute rate with operands resident in shared memory (no loads in flight): cycles per M x N x 16 MMA.
// mode 0: SS (A and B from shared memory), one accumulator; 1: SS, two accumulators alternating; 2: A from TMEM (TS), one
// accumulator.  m = 128 or 64.
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__global__ void __launch_bounds__(128, 1) t5_kernel(long long* out, int mode, int m, int n, int iters) {
  extern __shared__ uint8_t raw[];
  const uint32_t raw_u = smem_u32(raw);
  const uint32_t base = (raw_u + 1023u) & ~1023u;
  uint8_t* bp = raw + (base - raw_u);
  const uint32_t a_s = base, b_s = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
  const int tid = threadIdx.x;
  for (int i = tid; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  if (tid < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(bp + 16384 + 32768 + 16);
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(0, m, n);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int k = it & 3;
      const uint32_t d = tmem + ((mode == 1 && (it & 4)) ? 256u : 0u);
      if (mode == 2) umma_f16_ts(d, tmem + 480u, desc_sw128(b_s + k * 32, 1024, 0), idesc, it > 7 ? 1u : 0u);
      else umma_f16(d, desc_sw128(a_s + k * 32, 1024, 0), desc_sw128(b_s + k * 32, 1024, 0), idesc, it > 7 ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

static float h2f(__half h) { return __half2float(h); }

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  EncodeTiledFn enc = encoder();
  srand(1);

  // ------------------------------------------------------------------ T1
  {
    const int npix = 512;
    std::vector<__half> ha(npix * 64), hb(64 * 64);
    for (auto& v : ha) v = __float2half((rand() % 17 - 8) / 8.0f);
    for (auto& v : hb) v = __float2half((rand() % 17 - 8) / 8.0f);
    __half *da, *db;
    float* dout;
    CK(cudaMalloc(&da, ha.size() * 2));
    CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMalloc(&dout, 128 * 64 * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
    const int shifts[] = {0, 1, 2, 3, 5, 8, 17, 18, 19, 33};
    const int sbos[] = {1024, 2048, 1280, 2304};
    std::vector<float> hout(128 * 64);
    printf("T1: max |err| of D = A_shifted . B^T  (rows p = shift + g*SBO/128 + r), exact integer-ish data\n");
    for (int sbo : sbos)
      for (int ubo = 0; ubo < 2; ++ubo) {
        printf("T1 sbo=%4d base_off=%s :", sbo, ubo ? "addr" : "0   ");
        for (int sh : shifts) {
          if (sh + 15 * (sbo / 128) + 8 > npix) { printf("  sh%-2d skip", sh); continue; }
          CK(cudaMemset(dout, 0xff, 128 * 64 * 4));
          t1_kernel<<<1, 128, 80 * 1024>>>(da, npix, db, dout, sh, sbo, ubo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("  sh%-2d LAUNCH-ERR %s\n", sh, cudaGetErrorString(e)); return 3; }
          CK(cudaMemcpy(hout.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
          double worst = 0;
          for (int m = 0; m < 128; ++m) {
            const int p = sh + (m / 8) * (sbo / 128) + (m % 8);
            for (int n = 0; n < 64; ++n) {
              double acc = 0;
              for (int k = 0; k < 64; ++k) acc += h2f(ha[p * 64 + k]) * h2f(hb[n * 64 + k]);
              worst = fmax(worst, fabs(acc - hout[m * 64 + n]));
            }
          }
          printf("  sh%-2d %s", sh, worst < 1e-3 ? "OK  " : "BAD ");
        }
        printf("\n");
      }
    cudaFree(da); cudaFree(db); cudaFree(dout);
  }

  // ------------------------------------------------------------------ T2 / T3
  {
    const int B = 2, H = 20, W = 24, C = 128;
    std::vector<__half> hx(static_cast<size_t>(B) * H * W * C), hw(64 * 9 * 64);
    for (auto& v : hx) v = __float2half((rand() % 17 - 8) / 8.0f);
    for (auto& v : hw) v = __float2half((rand() % 9 - 4) / 8.0f);
    __half *dx, *dw;
    uint8_t* ddump;
    float* dout;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&ddump, 36864));
    CK(cudaMalloc(&dout, 128 * 64 * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmx, tmw;
    {
      cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {64, 16, 18, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("T2: encode 4d failed %d\n", (int)r); return 4; }
      cuuint64_t gd2[2] = {9 * 64, 64};
      cuuint64_t gs2[1] = {9 * 64 * 2};
      cuuint32_t b2[2] = {64, 64};
      cuuint32_t e2[2] = {1, 1};
      r = enc(&tmw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, gd2, gs2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("T2: encode 2d failed %d\n", (int)r); return 4; }
    }
    CK(cudaFuncSetAttribute(t23_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 124 * 1024));
    struct { int c0, w0, h0, n0; } cases[] = {{64, -1, -1, 1}, {0, 7, 3, 0}, {64, 15, 15, 1}};
    for (auto cs : cases)
      for (int ubo = 0; ubo < 2; ++ubo) {
        t23_kernel<<<1, 128, 124 * 1024>>>(tmx, tmw, ddump, dout, cs.c0, cs.w0, cs.h0, cs.n0, ubo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T2/T3 LAUNCH-ERR %s\n", cudaGetErrorString(e)); return 3; }
        std::vector<__half> hd(36864 / 2);
        std::vector<float> ho(128 * 64);
        CK(cudaMemcpy(hd.data(), ddump, 36864, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ho.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
        auto xin = [&](int n, int y, int x, int c) -> float {
          if (y < 0 || y >= H || x < 0 || x >= W) return 0.f;
          return h2f(hx[((static_cast<size_t>(n) * H + y) * W + x) * C + c]);
        };
        int bad_layout = 0;
        for (int hh = 0; hh < 18; ++hh)
          for (int ww = 0; ww < 16; ++ww) {
            const int p = hh * 16 + ww;
            for (int j = 0; j < 8; ++j)
              for (int e8 = 0; e8 < 8; ++e8) {
                const float got = h2f(hd[(p * 128 + ((j ^ (p & 7)) << 4)) / 2 + e8]);
                if (got != xin(cs.n0, cs.h0 + hh, cs.w0 + ww, cs.c0 + j * 8 + e8)) ++bad_layout;
              }
          }
        double worst = 0;
        for (int m = 0; m < 128; ++m) {
          const int r = m >> 3, c = m & 7;
          for (int n = 0; n < 64; ++n) {
            double acc = 0;
            for (int t = 0; t < 9; ++t)
              for (int k = 0; k < 64; ++k)
                acc += xin(cs.n0, cs.h0 + r + t / 3, cs.w0 + c + t % 3, cs.c0 + k) * h2f(hw[n * 576 + t * 64 + k]);
            worst = fmax(worst, fabs(acc - ho[m * 64 + n]));
          }
        }
        printf("T2 box@(c%d,w%d,h%d,n%d): layout mismatches %d (0 = pixel p=h*16+w at p*128, chunk j at j^(p&7), OOB zero) | "
               "T3 base_off=%s 3x3 patch conv max|err| %.4g %s\n", cs.c0, cs.w0, cs.h0, cs.n0, bad_layout, ubo ? "addr" : "0",
               worst, worst < 1e-2 ? "OK" : "BAD");
      }
    cudaFree(dx); cudaFree(dw); cudaFree(ddump); cudaFree(dout);
  }

  // ------------------------------------------------------------------ T4
  {
    const long long rows_total = 8ll << 20;                 // 8 Mi rows x 128 B = 1 GiB
    uint8_t* dsrc;
    CK(cudaMalloc(&dsrc, rows_total * 128));
    CK(cudaMemset(dsrc, 1, rows_total * 128));
    long long* dcyc;
    CK(cudaMalloc(&dcyc, 256 * 8));
    CK(cudaFuncSetAttribute(t4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Cfg { const char* name; int mode, grid, stages, box_rows, nprod, region_rows, same, row_stride; };
    const Cfg cfgs[] = {
        {"tma 16KB x8 stages, same 2MB region (weights-like), 148 CTAs", 0, 148, 8, 128, 0, 16384, 1, 1},
        {"tma 32KB x4 stages, same 2MB region, 148 CTAs", 0, 148, 4, 256, 0, 16384, 1, 1},
        {"tma 32KB x6 stages, same 2MB region, 148 CTAs", 0, 148, 6, 256, 0, 16384, 1, 1},
        {"tma 16KB x8 stages, per-CTA 256KB regions (L2 resident), 148 CTAs", 0, 148, 8, 128, 0, 2048, 0, 1},
        {"tma 32KB x6 stages, per-CTA 256KB regions (L2 resident), 148 CTAs", 0, 148, 6, 256, 0, 2048, 0, 1},
        {"tma 32KB x6 stages, per-CTA 256KB regions (L2 resident), 74 CTAs", 0, 74, 6, 256, 0, 2048, 0, 1},
        {"tma 32KB x6 stages, per-CTA 256KB regions (L2 resident), 16 CTAs", 0, 16, 6, 256, 0, 2048, 0, 1},
        {"tma 32KB x6 stages, per-CTA 6.9MB regions (HBM stream), 148 CTAs", 0, 148, 6, 256, 0, 56320, 0, 1},
        {"cp.async 128 thr, 16KB x8 stages, per-CTA 256KB regions, contiguous rows, 148 CTAs", 1, 148, 8, 128, 128, 2048, 0, 1},
        {"cp.async 256 thr, 16KB x8 stages, per-CTA 256KB regions, contiguous rows, 148 CTAs", 1, 148, 8, 128, 256, 2048, 0, 1},
        {"cp.async 128 thr, 16KB x8 stages, per-CTA 256KB regions, row stride 3 (gather), 148 CTAs", 1, 148, 8, 128, 128, 2048, 0, 3},
        {"cp.async 256 thr, 16KB x8 stages, per-CTA 256KB regions, row stride 3 (gather), 148 CTAs", 1, 148, 8, 128, 256, 2048, 0, 3},
        {"cp.async 128 thr, 16KB x8 stages, per-CTA 256KB regions, contiguous rows, 16 CTAs", 1, 16, 8, 128, 128, 2048, 0, 1},
    };
    CUtensorMap tm128, tm256;
    for (int which = 0; which < 2; ++which) {
      cuuint64_t gd[2] = {64, (cuuint64_t)rows_total};
      cuuint64_t gs[1] = {128};
      cuuint32_t bx[2] = {64, which ? 256u : 128u};
      cuuint32_t es[2] = {1, 1};
      CUresult r = enc(which ? &tm256 : &tm128, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dsrc, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("T4: encode failed %d\n", (int)r); return 4; }
    }
    for (const Cfg& c : cfgs) {
      T4Params p;
      p.mode = c.mode; p.iters = 4096; p.stages = c.stages; p.box_rows = c.box_rows; p.nprod = c.nprod;
      p.region_rows = c.region_rows; p.same_region = c.same; p.row_stride = c.row_stride; p.src = dsrc; p.cycles = dcyc;
      const size_t smem = 3072 + static_cast<size_t>(c.stages) * c.box_rows * 128;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        t4_kernel<<<c.grid, 320, smem>>>(c.box_rows == 256 ? tm256 : tm128, p);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("T4 LAUNCH-ERR %s (%s)\n", cudaGetErrorString(e), c.name); return 3; }
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cyc(c.grid);
      CK(cudaMemcpy(cyc.data(), dcyc, c.grid * 8, cudaMemcpyDeviceToHost));
      long long cmax = 0; double csum = 0;
      for (long long v : cyc) { cmax = v > cmax ? v : cmax; csum += v; }
      const double bytes = 4096.0 * c.box_rows * 128;
      printf("T4 %-88s: %.1f B/clk/SM (mean), %.1f (slowest CTA); chip %.2f TB/s\n", c.name, bytes / (csum / c.grid), bytes / cmax,
             bytes * c.grid / (ms * 1e-3) / 1e12);
    }
  }
  // ------------------------------------------------------------------ T5
  {
    long long* dout;
    CK(cudaMalloc(&dout, 16));
    CK(cudaFuncSetAttribute(t5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
    const int iters = 512;
    const char* mname[3] = {"SS 1 accumulator ", "SS 2 accumulators", "TS (A in TMEM)   "};
    for (int mode = 0; mode < 3; ++mode)
      for (int m : {128, 64})
        for (int n : {32, 64, 128, 256}) {
          if (mode == 1 && n > 128 && false) continue;
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) {
            t5_kernel<<<1, 128, 52 * 1024>>>(dout, mode, m, n, iters);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("T5 LAUNCH-ERR %s (mode %d m %d n %d)\n", cudaGetErrorString(e), mode, m, n); return 3; }
          }
          CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
          printf("T5 %s M=%3d N=%3d: %.1f cycles per MMA to issue, %.1f to complete  (math floor %d)\n", mname[mode], m, n,
                 (double)h[0] / iters, (double)h[1] / iters, n / 2);
        }
  }
  printf("probe done\n");
  return 0;
}
