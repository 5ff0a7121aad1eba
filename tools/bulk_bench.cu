// Microbenchmark: throughput of a cp.async.bulk (global -> shared, mbarrier completion) ring as a function
// of bytes per copy, copies per stage and ring depth -- the weight stream of the streaming conv kernel.
// A producer thread fills stages, a consumer thread waits for each and releases it at once (no MMAs), so
// the result is the ceiling the copy path itself sets.  Source = a 4 MB L2-resident buffer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/bulk_bench tools/bulk_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(64) bench(const uint8_t* src, size_t src_bytes, int copy_bytes, int copies, int depth,
                                            int src_stride, int iters, int mode, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const uint32_t bar0 = smem_u32(bars);
  const int stage_bytes = copy_bytes * copies;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * depth; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    size_t off = (mode & 2) ? 0 : ((size_t)blockIdx.x * 65536) % src_bytes;
    for (int it = 0; it < iters; ++it) {
      const int s = it % depth;
      if ((mode & 1) == 0) mbar_wait(bar0 + 8 * (depth + s), ((it / depth) & 1) ^ 1);
      mbar_expect_tx(bar0 + 8 * s, (uint32_t)stage_bytes);
      for (int c = 0; c < copies; ++c) {
        bulk_load(smem_u32(smem) + s * stage_bytes + c * copy_bytes, src + off, (uint32_t)copy_bytes, bar0 + 8 * s);
        off += src_stride;
        if (off + copy_bytes > src_bytes) off = 0;
      }
    }
  } else if (threadIdx.x == 32 && (mode & 1) == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % depth;
      mbar_wait(bar0 + 8 * s, (it / depth) & 1);
      mbar_arrive(bar0 + 8 * (depth + s));
    }
  }
  if (threadIdx.x == 0 && (mode & 1) == 1) {   // issue-only mode: time the issue loop, then let the last phases land
    out[blockIdx.x] = clock64() - t0;
    long long t1 = clock64();
    while (clock64() - t1 < 400000) {
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && (mode & 1) == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
  const size_t src_bytes = 4 << 20;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  cudaMalloc(&out, 1024 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 400;
  printf("%8s %7s %6s %7s %6s | %10s %12s %12s\n", "copyB", "copies", "depth", "stride", "grid", "cyc/stage", "B/cyc/CTA", "B/cyc chip");
  const int grids[2] = {148, 296};
  struct C { int bytes, copies, depth, stride; } cs[] = {
      {2048, 4, 8, 4096}, {2048, 4, 8, 2048}, {8192, 1, 8, 8192}, {1024, 4, 8, 2048}, {1024, 8, 8, 2048}, {4096, 2, 8, 4096},
      {2048, 4, 4, 4096}, {2048, 4, 2, 4096}, {16384, 1, 4, 16384}, {16384, 1, 8, 16384}, {2048, 2, 8, 4096}, {2048, 1, 8, 4096},
      {512, 4, 8, 1024}, {32768, 1, 4, 32768}};
  uint8_t* flush;
  cudaMalloc(&flush, 256 << 20);
  for (int mode = 0; mode < 7; mode += 2)   // 0 = distinct addresses warm L2, 2 = same addresses, 4 = distinct cold, 6 = same cold
  for (auto c : cs)
    for (int g : grids) {
      if (g != 148) continue;
      if (mode & 4) cudaMemset(flush, 0, 256 << 20);
      size_t smem = (size_t)c.bytes * c.copies * c.depth;
      if (g == 296 && smem > 100 * 1024) continue;
      bench<<<g, 64, smem>>>(src, src_bytes, c.bytes, c.copies, c.depth, c.stride, iters, mode, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      long long h[1024];
      cudaMemcpy(h, out, g * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;
      for (int i = 0; i < g; ++i) avg += (double)h[i];
      avg /= g;
      double per_stage = avg / iters, bpc = (double)c.bytes * c.copies / per_stage;
      printf("m%d %8d %7d %6d %7d %6d | %10.1f %12.2f %12.1f\n", mode, c.bytes, c.copies, c.depth, c.stride, g, per_stage, bpc, bpc * g);
    }
  return 0;
}
