// reduce_probe.cu — what bounds the BatchNorm-backward channel reduction (sum g, sum g*x over rows of an NHWC 16-bit tensor):
// grid size, loads in flight per thread, and how the per-block partial sums reach global memory
//   mode 0  scalar red.global.add.f32 from a shared-memory stage (the round-1 kernel)
//   mode 1  red.global.add.v4.f32 (sm_90+), 4x fewer atomic operations
//   mode 2  no global reduction at all (timing only: what the loads alone cost)
//   mode 3  plain stores of per-block partials [grid][2C] + a second kernel that folds them
// Buffers are larger than L2 in rotation (cold) or reused (warm).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/reduce_probe tools/probe/reduce_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                        \
    }                                                                                 \
  } while (0)

typedef __nv_bfloat16 T;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(h[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int R, int kThreads>
__global__ void __launch_bounds__(kThreads) reduce_kernel(const T* __restrict__ dz, const T* __restrict__ out, const T* __restrict__ x,
                                                          float* __restrict__ sums, float* __restrict__ partials, long long rows, int C,
                                                          int mode) {
  extern __shared__ float acc[];   // [C][2]
  const int cv = C / 8;
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long tthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const int c = static_cast<int>(gtid % cv) * 8;
  const long long rstep = tthreads / cv;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  for (long long m = gtid / cv; m < rows; m += R * rstep) {
    uint4 a[R], b[R], cc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long mr = m + r * rstep;
      const bool ok = mr < rows;
      a[r] = ok ? __ldg(reinterpret_cast<const uint4*>(dz + mr * C + c)) : zero4;
      b[r] = ok ? __ldg(reinterpret_cast<const uint4*>(out + mr * C + c)) : zero4;
      cc[r] = ok ? __ldg(reinterpret_cast<const uint4*>(x + mr * C + c)) : zero4;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float g[8], o[8], xv[8];
      unpack8(a[r], g);
      unpack8(b[r], o);
      unpack8(cc[r], xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = o[j] > 0.f ? g[j] : 0.f;
        s1[j] += g[j];
        s2[j] = fmaf(g[j], xv[j], s2[j]);
      }
    }
  }
  const int lane = threadIdx.x & 31;
  bool writer = true;
  if (cv < 32 && (32 % cv) == 0) {
    for (int off = 16; off >= cv; off >>= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
      }
    }
    writer = lane < cv;
  }
  if (writer) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(acc + 2 * (c + j), s1[j]);
      atomicAdd(acc + 2 * (c + j) + 1, s2[j]);
    }
  }
  __syncthreads();
  if (mode == 0) {
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(sums + i, acc[i]);
  } else if (mode == 1) {
    for (int i = threadIdx.x * 4; i < C * 2; i += blockDim.x * 4) red_v4(sums + i, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  } else if (mode == 3) {
    float* p = partials + static_cast<long long>(blockIdx.x) * C * 2;
    for (int i = threadIdx.x * 4; i < C * 2; i += blockDim.x * 4)
      *reinterpret_cast<float4*>(p + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
  } else {
    if (acc[threadIdx.x % (2 * C)] == 123.456f) sums[0] = 1.f;
  }
}

__global__ void fold_kernel(const float* __restrict__ partials, float* __restrict__ sums, int nblk, int C2) {
  // one thread per (output, slice of blocks); 32 slices folded by shuffles
  const int o = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x & 31;
  if (o >= C2) return;
  float s = 0.f;
  for (int b = lane; b < nblk; b += 32) s += partials[static_cast<long long>(b) * C2 + o];
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) sums[o] = s;
}

// the apply pass next to it, for scale: dx = a*g + b*x + k, g = dz * (out > 0)
__global__ void __launch_bounds__(256) apply_kernel(const T* __restrict__ dz, const T* __restrict__ out, const T* __restrict__ x,
                                                    T* __restrict__ dx, long long rows, int C) {
  const int cv = C / 8;
  const long long gtid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long rstep = (static_cast<long long>(gridDim.x) * blockDim.x) / cv;
  const int c = static_cast<int>(gtid % cv) * 8;
  for (long long m = gtid / cv; m < rows; m += rstep) {
    float g[8], o[8], xv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(dz + m * C + c)), g);
    unpack8(__ldg(reinterpret_cast<const uint4*>(out + m * C + c)), o);
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + m * C + c)), xv);
    __nv_bfloat162 r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g0 = o[2 * j] > 0.f ? g[2 * j] : 0.f, g1 = o[2 * j + 1] > 0.f ? g[2 * j + 1] : 0.f;
      r[j] = __floats2bfloat162_rn(fmaf(0.5f, g0, 0.25f * xv[2 * j]), fmaf(0.5f, g1, 0.25f * xv[2 * j + 1]));
    }
    *reinterpret_cast<uint4*>(dx + m * C + c) = *reinterpret_cast<uint4*>(r);
  }
}

static int gcd(int a, int b) { return b ? gcd(b, a % b) : a; }

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  struct Shape { long long rows; int C; const char* what; };
  const Shape shapes[] = {{614400, 64, "stem bn1"},      {153600, 256, "layer1 bn3"}, {153600, 64, "layer1 bn1/2"},
                          {38400, 512, "layer2 bn3"},    {38400, 128, "layer2 bn1/2"}, {9600, 1024, "layer3 bn3"},
                          {9600, 256, "layer3 bn1/2"},   {2400, 2048, "layer4 bn3"},  {2400, 512, "layer4 bn1/2"}};
  const size_t kMaxElems = 614400ll * 64;   // 39.3M elements = 78.6 MB per tensor
  const int kRot = 3;                        // 3 tensors x 3 rotations x 78.6 MB = 707 MB >> L2
  std::vector<T*> bufs(3 * kRot);
  for (auto& p : bufs) {
    CK(cudaMalloc(&p, kMaxElems * sizeof(T)));
    CK(cudaMemset(p, 0x3c, kMaxElems * sizeof(T)));
  }
  T* dxbuf;
  CK(cudaMalloc(&dxbuf, kMaxElems * sizeof(T)));
  float *sums, *partials, *flush;
  CK(cudaMalloc(&sums, 4096 * 2 * sizeof(float)));
  CK(cudaMalloc(&partials, 1184ll * 4096 * 2 * sizeof(float)));
  const size_t kFlush = 256u << 20;
  CK(cudaMalloc(&flush, kFlush));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(reduce_kernel<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(reduce_kernel<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(reduce_kernel<4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  int occ2 = 0, occ4 = 0, occ45 = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, reduce_kernel<2, 256>, 256, 8192));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, reduce_kernel<4, 256>, 256, 8192));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ45, reduce_kernel<4, 512>, 512, 8192));
  printf("SMs %d; resident blocks per SM: R=2/256thr %d, R=4/256thr %d, R=4/512thr %d\n", sms, occ2, occ4, occ45);

  auto time_it = [&](auto launch, bool cold) {
    float best = 1e9f, sum = 0.f;
    const int n = 6;
    for (int it = 0; it < n + 2; ++it) {
      if (cold) CK(cudaMemsetAsync(flush, it, kFlush));
      CK(cudaEventRecord(e0));
      launch(it % kRot);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    return sum / n * 1e3f;
  };

  for (const Shape& s : shapes) {
    const int cv = s.C / 8;
    const double mb = s.rows * (double)s.C * 2 * 3 / 1e6;
    printf("\n== %s: rows %lld C %d, 3 x %.1f MB read (%.1f us at 7.4 TB/s)\n", s.what, s.rows, s.C, mb / 3, mb / 7.4);
    // the apply pass for scale
    {
      const int g0 = cv / gcd(cv, 256);
      long long want = (s.rows * cv + 256 * 4 - 1) / (256 * 4);
      if (want > sms * 8) want = sms * 8;
      int grid = (int)(want / g0 * g0);
      if (grid < g0) grid = g0;
      const float t = time_it([&](int r) { apply_kernel<<<grid, 256>>>(bufs[r * 3], bufs[r * 3 + 1], bufs[r * 3 + 2], dxbuf, s.rows, s.C); }, true);
      printf("  apply pass (3 reads + 1 write), grid %d: %.1f us cold\n", grid, t);
    }
    for (int variant = 0; variant < 3; ++variant) {          // 0: R=2/256, 1: R=4/256, 2: R=4/512
      const int thr = variant == 2 ? 512 : 256;
      const int R = variant == 0 ? 2 : 4;
      const int g0 = cv / gcd(cv, thr);
      for (int gsel = 0; gsel < 5; ++gsel) {
        long long want;
        if (gsel == 0) {                                      // the shipped rule: 8 rows per thread, cap 8 blocks per SM
          want = (s.rows * cv + thr * 8ll - 1) / (thr * 8ll);
          if (want > sms * 8) want = sms * 8;
        } else {
          want = (long long)sms * (1 << (gsel - 1));          // 1, 2, 4, 8 blocks per SM
          const long long maxb = (s.rows * cv + thr - 1) / thr;
          if (want > maxb) want = maxb;
        }
        int grid = (int)(want / g0 * g0);
        if (grid < g0) grid = g0;
        printf("  R=%d thr=%d grid %4d (%s):", R, thr, grid, gsel == 0 ? "shipped rule" : "SMs x k   ");
        for (int mode = 0; mode < 4; ++mode) {
          auto launch = [&](int r) {
            const size_t smem = (size_t)s.C * 2 * sizeof(float);
            if (mode != 3) CK(cudaMemsetAsync(sums, 0, s.C * 2 * sizeof(float)));
            if (variant == 0) reduce_kernel<2, 256><<<grid, 256, smem>>>(bufs[r * 3], bufs[r * 3 + 1], bufs[r * 3 + 2], sums, partials, s.rows, s.C, mode);
            else if (variant == 1) reduce_kernel<4, 256><<<grid, 256, smem>>>(bufs[r * 3], bufs[r * 3 + 1], bufs[r * 3 + 2], sums, partials, s.rows, s.C, mode);
            else reduce_kernel<4, 512><<<grid, 512, smem>>>(bufs[r * 3], bufs[r * 3 + 1], bufs[r * 3 + 2], sums, partials, s.rows, s.C, mode);
            if (mode == 3) fold_kernel<<<(s.C * 2 + 7) / 8, 256>>>(partials, sums, grid, s.C * 2);
          };
          const float tc = time_it(launch, true);
          const float tw = time_it(launch, false);
          static const char* names[] = {"red", "red.v4", "none", "partials+fold"};
          printf("  %s %.1f/%.1f", names[mode], tc, tw);
        }
        printf("  us cold/warm\n");
      }
    }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
