// Microbenchmark: ways to zero-fill 462 MB on B200 (write-only HBM stream).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zfill_bench zfill_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int CHUNK, int ISSUERS>
__global__ void tma_fill(uint8_t *a, unsigned long long bytes) {
  extern __shared__ __align__(128) uint8_t zbuf[];
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4 *>(zbuf)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0 || w >= ISSUERS) return;
  const unsigned long long n = (bytes + CHUNK - 1) / CHUNK;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
  for (unsigned long long c = (unsigned long long)blockIdx.x * ISSUERS + w; c < n; c += (unsigned long long)gridDim.x * ISSUERS) {
    unsigned long long left = bytes - c * CHUNK;
    uint32_t sz = left < CHUNK ? (uint32_t)left : CHUNK;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a + c * CHUNK), "r"(src), "r"(sz) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int CHUNK, int MODE>
__global__ void tma_fill_hint(uint8_t *a, unsigned long long bytes) {
  extern __shared__ __align__(128) uint8_t zbuf[];
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4 *>(zbuf)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  uint64_t pol;
  if (MODE == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
  const unsigned long long n = (bytes + CHUNK - 1) / CHUNK;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
  for (unsigned long long c = blockIdx.x; c < n; c += gridDim.x) {
    unsigned long long left = bytes - c * CHUNK;
    uint32_t sz = left < CHUNK ? (uint32_t)left : CHUNK;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(a + c * CHUNK), "r"(src), "r"(sz), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// block-contiguous assignment like the fused kernel: block b owns bytes [b*per, (b+1)*per)
template <int CHUNK>
__global__ void tma_fill_owned(uint8_t *a, unsigned long long bytes) {
  extern __shared__ __align__(128) uint8_t zbuf[];
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) reinterpret_cast<uint4 *>(zbuf)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned long long per = bytes / gridDim.x;  // multiple of 16 by construction
  uint8_t *base = a + per * blockIdx.x;
  const uint32_t src = (uint32_t)__cvta_generic_to_shared(zbuf);
  for (unsigned long long off = 0; off < per; off += CHUNK) {
    unsigned long long left = per - off;
    uint32_t sz = left < CHUNK ? (uint32_t)left : CHUNK;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(src), "r"(sz) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void st128_wt(uint4 *a, unsigned long long n16) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
    __stwt(a + i, make_uint4(0, 0, 0, 0));
}

__global__ void st128_fill(uint4 *a, unsigned long long n16) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
    __stcs(a + i, make_uint4(0, 0, 0, 0));
}
__global__ void st128_fill_plain(uint4 *a, unsigned long long n16) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
    a[i] = make_uint4(0, 0, 0, 0);
}
__global__ void st256_fill(uint8_t *a, unsigned long long n32) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += stride) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(a + i * 32), "f"(0.0f) : "memory");
  }
}

int main() {
  const unsigned long long bytes = 4096ull * 112896ull;
  uint8_t *buf;
  CK(cudaMalloc(&buf, bytes));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto report = [&](const char *name, float ms, int reps) {
    printf("%-34s %8.2f us  %8.1f GB/s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
  };
  const int reps = 20;
  float ms;
#define TIME(name, launch)                                   \
  for (int i = 0; i < 3; ++i) { launch; }                    \
  CK(cudaDeviceSynchronize());                               \
  cudaEventRecord(e0);                                       \
  for (int i = 0; i < reps; ++i) { launch; }                 \
  cudaEventRecord(e1);                                       \
  CK(cudaDeviceSynchronize());                               \
  cudaEventElapsedTime(&ms, e0, e1);                         \
  report(name, ms, reps);

  TIME("cudaMemsetAsync", cudaMemsetAsync(buf, 0, bytes));
  CK(cudaFuncSetAttribute(tma_fill<32768, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(tma_fill<32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(tma_fill<8192, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
  CK(cudaFuncSetAttribute(tma_fill<4096, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096));
  CK(cudaFuncSetAttribute(tma_fill<65536, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  TIME("tma 32K x1 issuer, grid 296", (tma_fill<32768, 1><<<296, 128, 32768>>>(buf, bytes)));
  TIME("tma 32K x4 issuers, grid 296", (tma_fill<32768, 4><<<296, 128, 32768>>>(buf, bytes)));
  TIME("tma 32K x4 issuers, grid 592", (tma_fill<32768, 4><<<592, 128, 32768>>>(buf, bytes)));
  TIME("tma 8K x4 issuers, grid 592", (tma_fill<8192, 4><<<592, 128, 8192>>>(buf, bytes)));
  TIME("tma 8K x4 issuers, grid 1184", (tma_fill<8192, 4><<<1184, 128, 8192>>>(buf, bytes)));
  TIME("tma 4K x4 issuers, grid 1184", (tma_fill<4096, 4><<<1184, 128, 4096>>>(buf, bytes)));
  TIME("tma 64K x4 issuers, grid 296", (tma_fill<65536, 4><<<296, 128, 65536>>>(buf, bytes)));
  CK(cudaFuncSetAttribute(tma_fill_hint<32768, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(tma_fill_hint<32768, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(tma_fill_hint<32768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  TIME("tma 32K evict_first, grid 296", (tma_fill_hint<32768, 0><<<296, 128, 32768>>>(buf, bytes)));
  TIME("tma 32K evict_last, grid 296", (tma_fill_hint<32768, 1><<<296, 128, 32768>>>(buf, bytes)));
  TIME("tma 32K evict_unchanged, grid 296", (tma_fill_hint<32768, 2><<<296, 128, 32768>>>(buf, bytes)));
  TIME("tma owned 16K, grid 1024", (tma_fill_owned<16384><<<1024, 128, 16384>>>(buf, bytes)));
  TIME("tma owned 8K, grid 1024", (tma_fill_owned<8192><<<1024, 128, 8192>>>(buf, bytes)));
  TIME("tma owned 16K, grid 4096", (tma_fill_owned<16384><<<4096, 32, 16384>>>(buf, bytes)));
  TIME("st.wt 128b, grid 2368 x 256", (st128_wt<<<2368, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
  TIME("st.wt 128b, grid 4736 x 256", (st128_wt<<<4736, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
  TIME("st.cs 128b, grid 4736 x 256", (st128_fill<<<4736, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
  TIME("st.cs 128b, grid 9472 x 256", (st128_fill<<<9472, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
  TIME("st.cs 128b, grid 9472 x 128", (st128_fill<<<9472, 128>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
  for (int g : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    char name[64];
    snprintf(name, 64, "st.cs 128b, grid %d x 256", g);
    TIME(name, (st128_fill<<<g, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
    snprintf(name, 64, "st 128b plain, grid %d x 256", g);
    TIME(name, (st128_fill_plain<<<g, 256>>>(reinterpret_cast<uint4 *>(buf), bytes / 16)));
    snprintf(name, 64, "st 256b, grid %d x 256", g);
    TIME(name, (st256_fill<<<g, 256>>>(buf, bytes / 32)));
  }
  return 0;
}
