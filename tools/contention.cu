// Which SM resource does a co-running latency/issue-bound kernel take away from a streaming store kernel?
// fill: 28,672 CTAs x 128 threads, 4 KB of st.global.cs per warp (the shape of expand_kernel), low priority.
// co:   1,024 CTAs x 128 threads, 21 KB smem, ~30 us of (1) nanosleep, (2) dependent ALU, (3) shared-memory
//       byte loads in a dependent chain, (4) independent shared-memory loads; high priority.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(128) fill(float4 *out, unsigned long long n16) {
  const unsigned long long warp = (unsigned long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  float4 *dst = out + warp * 256;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (warp * 256 + j * 32 + lane < n16) __stcs(dst + j * 32 + lane, make_float4(0, 0, 0, 0));
}

// persistent variant: `gridDim.x` CTAs of `blockDim.x` threads stride over the 4 KB groups in address order
__global__ void fill_persistent(float4 *out, unsigned long long n16) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const unsigned long long groups = (n16 + 255) / 256, stride = (unsigned long long)gridDim.x * wpb;
  for (unsigned long long warp = (unsigned long long)blockIdx.x * wpb + (threadIdx.x >> 5); warp < groups; warp += stride) {
    float4 *dst = out + warp * 256;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (warp * 256 + j * 32 + lane < n16) __stcs(dst + j * 32 + lane, make_float4(0, 0, 0, 0));
  }
}

__global__ void __launch_bounds__(128) co(int mode, int iters, int *sink) {
  __shared__ uint8_t sm[21000];
  for (int i = threadIdx.x; i < 21000; i += 128) sm[i] = (uint8_t)(i * 7 + 1);
  __syncthreads();
  unsigned x = threadIdx.x * 2654435761u + blockIdx.x;
  if (mode == 1) {
    for (int i = 0; i < iters; ++i) __nanosleep(1000);
  } else if (mode == 2) {
    for (int i = 0; i < iters * 40; ++i) x = x * 1664525u + 1013904223u + (x >> 7);
  } else if (mode == 3) {
    for (int i = 0; i < iters * 10; ++i) x = x + sm[(x >> 3) % 21000] * 31u + 1u;
  } else if (mode == 4) {
    unsigned a = 0, b = 0, c = 0, d = 0;
    for (int i = 0; i < iters * 6; ++i) {
      a += sm[(x + i) % 21000]; b += sm[(x + 2 * i + 5) % 21000]; c += sm[(x + 3 * i + 11) % 21000]; d += sm[(x + 5 * i + 17) % 21000];
    }
    x = a ^ b ^ c ^ d;
  }
  if (x == 0xdeadbeef) *sink = 1;
}

int main() {
  const unsigned long long bytes = 4096ull * 112896ull, n16 = bytes / 16;
  float4 *buf; int *sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  int least, greatest;
  cudaDeviceGetStreamPriorityRange(&least, &greatest);
  cudaStream_t lo, hi;
  cudaStreamCreateWithPriority(&lo, cudaStreamNonBlocking, least);
  cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, greatest);
  cudaEvent_t e0, e1, h0, h1, go;
  cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&h0); cudaEventCreate(&h1); cudaEventCreateWithFlags(&go, cudaEventDisableTiming);
  const unsigned grid = (unsigned)((n16 / 256 + 3) / 4);
  const char *names[] = {"none", "nanosleep (resident, idle)", "dependent ALU", "dependent LDS.U8 chain", "independent LDS"};
  // calibrate iters for ~30 us per mode
  for (int mode = 0; mode <= 4; ++mode) {
    int iters = 30;
    float co_ms = 0;
    if (mode) {
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(mode, iters, sink); cudaEventRecord(h1, hi);
        cudaDeviceSynchronize(); cudaEventElapsedTime(&co_ms, h0, h1);
        iters = (int)(iters * 0.030f / co_ms) + 1;
      }
    }
    float fill_ms = 0, sum = 0, cosum = 0;
    const int reps = 20;
    for (int r = 0; r < reps + 3; ++r) {
      cudaDeviceSynchronize();
      cudaEventRecord(go, lo);                 // both streams start together
      cudaStreamWaitEvent(hi, go, 0);
      if (mode) { cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(mode, iters, sink); cudaEventRecord(h1, hi); }
      cudaEventRecord(e0, lo); fill<<<grid, 128, 0, lo>>>(buf, n16); cudaEventRecord(e1, lo);
      cudaDeviceSynchronize();
      cudaEventElapsedTime(&fill_ms, e0, e1);
      if (mode) cudaEventElapsedTime(&co_ms, h0, h1);
      if (r >= 3) { sum += fill_ms; cosum += co_ms; }
    }
    printf("%-28s co alone ~30 us; together: co %6.1f us, fill %6.1f us (%.0f GB/s)\n", names[mode], mode ? cosum / reps * 1e3 : 0.0f,
           sum / reps * 1e3, bytes / (sum / reps * 1e-3) / 1e9);
  }
  // a persistent fill with few warps per SM keeps the SM's store queue short: what does the LDS chain see?
  {
    int iters = 30;
    float co_ms = 0;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(3, iters, sink); cudaEventRecord(h1, hi);
      cudaDeviceSynchronize(); cudaEventElapsedTime(&co_ms, h0, h1);
      iters = (int)(iters * 0.030f / co_ms) + 1;
    }
    for (int wps : {2, 4, 8, 16, 32}) {
      const int threads = wps >= 8 ? 256 : wps * 32, ctas = 148 * (wps * 32 / threads);
      for (int with_co = 0; with_co < 2; ++with_co) {
        float fill_ms = 0, sum = 0, cosum = 0;
        const int reps = 20;
        for (int r = 0; r < reps + 3; ++r) {
          cudaDeviceSynchronize();
          cudaEventRecord(go, lo);
          cudaStreamWaitEvent(hi, go, 0);
          if (with_co) { cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(3, iters, sink); cudaEventRecord(h1, hi); }
          cudaEventRecord(e0, lo); fill_persistent<<<ctas, threads, 0, lo>>>(buf, n16); cudaEventRecord(e1, lo);
          cudaDeviceSynchronize();
          cudaEventElapsedTime(&fill_ms, e0, e1);
          if (with_co) cudaEventElapsedTime(&co_ms, h0, h1);
          if (r >= 3) { sum += fill_ms; cosum += co_ms; }
        }
        printf("persistent fill %2d warps/SM %s: co %6.1f us, fill %6.1f us (%.0f GB/s)\n", wps, with_co ? "beside the LDS chain" : "alone               ",
               with_co ? cosum / reps * 1e3 : 0.0f, sum / reps * 1e3, bytes / (sum / reps * 1e-3) / 1e9);
      }
    }
  }
  // the same co-kernels beside cudaMemsetAsync (does the driver's memset run on the SMs or on a copy engine?)
  for (int mode = 0; mode <= 4; mode += 3) {
    int iters = 30;
    float co_ms = 0;
    if (mode) {
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(mode, iters, sink); cudaEventRecord(h1, hi);
        cudaDeviceSynchronize(); cudaEventElapsedTime(&co_ms, h0, h1);
        iters = (int)(iters * 0.030f / co_ms) + 1;
      }
    }
    float fill_ms = 0, sum = 0, cosum = 0;
    const int reps = 20;
    for (int r = 0; r < reps + 3; ++r) {
      cudaDeviceSynchronize();
      cudaEventRecord(go, lo);
      cudaStreamWaitEvent(hi, go, 0);
      if (mode) { cudaEventRecord(h0, hi); co<<<1024, 128, 0, hi>>>(mode, iters, sink); cudaEventRecord(h1, hi); }
      cudaEventRecord(e0, lo); cudaMemsetAsync(buf, 0, bytes, lo); cudaEventRecord(e1, lo);
      cudaDeviceSynchronize();
      cudaEventElapsedTime(&fill_ms, e0, e1);
      if (mode) cudaEventElapsedTime(&co_ms, h0, h1);
      if (r >= 3) { sum += fill_ms; cosum += co_ms; }
    }
    printf("cudaMemsetAsync beside %-28s: co %6.1f us, memset %6.1f us (%.0f GB/s)\n", names[mode], mode ? cosum / reps * 1e3 : 0.0f,
           sum / reps * 1e3, bytes / (sum / reps * 1e-3) / 1e9);
  }
  return 0;
}
