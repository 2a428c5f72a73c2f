// SM partitioning with CUDA green contexts: does a streaming-store kernel on N2 SMs keep HBM busy while a
// shared-memory-latency-bound kernel runs undisturbed on the other N1 SMs?  (cf. tools/contention.cu)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o greenctx greenctx.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DRV(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *s_; cuGetErrorString(r_, &s_); printf("%s: %s\n", #x, s_); return 1; } } while (0)

__global__ void __launch_bounds__(128) fill(float4 *out, unsigned long long n16, unsigned long long warps) {
  const int lane = threadIdx.x & 31;
  for (unsigned long long warp = (unsigned long long)blockIdx.x * 4 + (threadIdx.x >> 5); warp < warps; warp += (unsigned long long)gridDim.x * 4) {
    float4 *dst = out + warp * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (warp * 256 + j * 32 + lane < n16) __stcs(dst + j * 32 + lane, make_float4(0, 0, 0, 0));
  }
}

// TMA variant: one elected thread per CTA streams zeros with cp.async.bulk shared->global (16 KB per op)
constexpr int ZT = 16384;
__global__ void __launch_bounds__(128) fill_tma(uint8_t *out, unsigned long long bytes) {
  extern __shared__ __align__(128) uint8_t z[];
  for (int i = threadIdx.x; i < ZT / 16; i += 128) reinterpret_cast<uint4 *>(z)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned long long n = (bytes + ZT - 1) / ZT;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(z);
  for (unsigned long long c = blockIdx.x; c < n; c += gridDim.x) {
    const unsigned long long left = bytes - c * ZT;
    const uint32_t sz = left < ZT ? (uint32_t)left : ZT;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * ZT), "r"(src), "r"(sz) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(128) co(int iters, int *sink) {
  __shared__ uint8_t sm[21000];
  for (int i = threadIdx.x; i < 21000; i += 128) sm[i] = (uint8_t)(i * 7 + 1);
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
  for (int i = 0; i < iters * 10; ++i) x = x + sm[(x >> 3) % 21000] * 31u + 1u;
  if (x == 0xdeadbeef) *sink = 1;
}

int main() {
  cudaFree(0);
  CUdevice dev;
  DRV(cuDeviceGet(&dev, 0));
  const unsigned long long bytes = 4096ull * 112896ull, n16 = bytes / 16, warps = (n16 + 255) / 256;
  float4 *buf; int *sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  // calibrate the co-kernel to ~30 us on the whole GPU (1024 CTAs, one wave)
  cudaEvent_t e0, e1, h0, h1;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&h0); cudaEventCreate(&h1);
  int iters = 30; float ms = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(h0); co<<<1024, 128>>>(iters, sink); cudaEventRecord(h1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, h0, h1); iters = (int)(iters * 0.030f / ms) + 1;
  }
  printf("co-kernel alone on 148 SMs: %.1f us\n", ms * 1e3);
  CUdevResource all;
  DRV(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("device SMs: %u\n", all.sm.smCount);
  for (unsigned n1 : {32u, 48u, 64u, 80u, 96u}) {
    CUdevResource part[1], rest;
    unsigned groups = 1;
    DRV(cuDevSmResourceSplitByCount(part, &groups, &all, &rest, 0, n1));
    CUdevResourceDesc dA, dB;
    DRV(cuDevResourceGenerateDesc(&dA, &part[0], 1));
    DRV(cuDevResourceGenerateDesc(&dB, &rest, 1));
    CUgreenCtx gA, gB;
    DRV(cuGreenCtxCreate(&gA, dA, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    DRV(cuGreenCtxCreate(&gB, dB, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream sA, sB;
    DRV(cuGreenCtxStreamCreate(&sA, gA, CU_STREAM_NON_BLOCKING, 0));
    DRV(cuGreenCtxStreamCreate(&sB, gB, CU_STREAM_NON_BLOCKING, 0));
    const unsigned nA = part[0].sm.smCount, nB = rest.sm.smCount;
    float fsum = 0, csum = 0, falone = 0;
    const int reps = 20;
    const unsigned grid = (unsigned)((warps + 3) / 4);
    for (int r = 0; r < reps + 3; ++r) {  // fill alone on the B partition
      cudaDeviceSynchronize();
      cudaEventRecord(e0, (cudaStream_t)sB); fill<<<grid, 128, 0, (cudaStream_t)sB>>>(buf, n16, warps); cudaEventRecord(e1, (cudaStream_t)sB);
      cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1); if (r >= 3) falone += ms;
    }
    for (int r = 0; r < reps + 3; ++r) {
      cudaDeviceSynchronize();
      cudaEventRecord(h0, (cudaStream_t)sA); co<<<1024, 128, 0, (cudaStream_t)sA>>>(iters, sink); cudaEventRecord(h1, (cudaStream_t)sA);
      cudaEventRecord(e0, (cudaStream_t)sB); fill<<<grid, 128, 0, (cudaStream_t)sB>>>(buf, n16, warps); cudaEventRecord(e1, (cudaStream_t)sB);
      cudaDeviceSynchronize();
      float f, c; cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&c, h0, h1);
      if (r >= 3) { fsum += f; csum += c; }
    }
    {
      cudaFuncSetAttribute(fill_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, ZT);
      for (int per_sm : {2, 4, 8}) {
        float tsum = 0;
        for (int r = 0; r < reps + 3; ++r) {
          cudaDeviceSynchronize();
          cudaEventRecord(e0, (cudaStream_t)sB); fill_tma<<<nB * per_sm, 128, ZT, (cudaStream_t)sB>>>((uint8_t *)buf, bytes); cudaEventRecord(e1, (cudaStream_t)sB);
          cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1); if (r >= 3) tsum += ms;
        }
        printf("   TMA fill alone on %3u SMs, %d CTAs/SM: %6.1f us (%.0f GB/s)\n", nB, per_sm, tsum / reps * 1e3, bytes / (tsum / reps * 1e-3) / 1e9);
      }
      float tsum = 0, csum2 = 0;
      for (int r = 0; r < reps + 3; ++r) {
        cudaDeviceSynchronize();
        cudaEventRecord(h0, (cudaStream_t)sA); co<<<1024, 128, 0, (cudaStream_t)sA>>>(iters, sink); cudaEventRecord(h1, (cudaStream_t)sA);
        cudaEventRecord(e0, (cudaStream_t)sB); fill_tma<<<nB * 4, 128, ZT, (cudaStream_t)sB>>>((uint8_t *)buf, bytes); cudaEventRecord(e1, (cudaStream_t)sB);
        cudaDeviceSynchronize();
        float f, c; cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&c, h0, h1);
        if (r >= 3) { tsum += f; csum2 += c; }
      }
      printf("   TMA fill + co together: co %6.1f us, fill %6.1f us (%.0f GB/s)\n", csum2 / reps * 1e3, tsum / reps * 1e3, bytes / (tsum / reps * 1e-3) / 1e9);
    }
    printf("co on %3u SMs, fill on %3u SMs: fill alone %6.1f us (%.0f GB/s); together: co %6.1f us, fill %6.1f us (%.0f GB/s)\n", nA, nB,
           falone / reps * 1e3, bytes / (falone / reps * 1e-3) / 1e9, csum / reps * 1e3, fsum / reps * 1e3, bytes / (fsum / reps * 1e-3) / 1e9);
    cuStreamDestroy(sA); cuStreamDestroy(sB); cuGreenCtxDestroy(gA); cuGreenCtxDestroy(gB);
  }
  return 0;
}
