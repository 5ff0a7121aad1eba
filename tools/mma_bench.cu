// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) as a function of
// N, operand major-ness and the un-swizzled descriptor strides (LBO / SBO).  Operands are whatever
// bytes sit in shared memory; only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
               "r"(acc) : "memory");
}

struct Case { int N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, a_layout, b_layout, nacc; };

__global__ void __launch_bounds__(128) bench(Case c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem), c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t b0 = make_desc(smem_u32(smem) + 96 * 1024, c.b_lbo, c.b_sbo, c.b_layout);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) tc_mma(tmem + (j % c.nacc) * c.N, a0 + (uint64_t)(j * 2), b0 + (uint64_t)(j & 1), idesc, 1);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  // {N, a_mn, b_mn, a_lbo, a_sbo, b_lbo, b_sbo, a_layout, b_layout, nacc}
  Case cases[] = {
      {64, 0, 0, 11520, 160, 1024, 128, 0, 0, 2},     // fprop 32->64 as shipped: K-major no swizzle
      {64, 0, 0, 11520 + 16, 160, 1024, 128, 0, 0, 2},   // LBO not a multiple of 128
      {64, 0, 0, 11520 + 64, 160, 1024, 128, 0, 0, 2},
      {64, 0, 0, 11520, 128, 1024, 128, 0, 0, 2},     // dense lines (SBO = 128)
      {64, 0, 0, 11520, 256, 1024, 128, 0, 0, 2},
      {64, 0, 0, 128, 256, 128, 256, 0, 0, 2},        // "textbook" interleaved tile: LBO 128, SBO 256
      {64, 0, 0, 11520, 160, 1024 + 16, 128, 0, 0, 2},
      {64, 1, 1, 128, 4096, 160, 5760, 0, 0, 2},      // wgrad-style MN-major
      {32, 1, 1, 128, 4096, 160, 5760, 0, 0, 2},
      {32, 0, 0, 11520, 160, 512, 128, 0, 0, 2},
      {16, 0, 0, 11520, 160, 256, 128, 0, 0, 2},
      {128, 0, 0, 11520, 160, 2048, 128, 0, 0, 2},
      {256, 0, 0, 11520, 160, 4096, 128, 0, 0, 2},
      {64, 0, 0, 0, 1024, 0, 1024, 2, 2, 2},          // SWIZZLE_128B K-major (standard GEMM layout)
      {128, 0, 0, 0, 1024, 0, 1024, 2, 2, 2},
      {256, 0, 0, 0, 1024, 0, 1024, 2, 2, 2},
      {64, 0, 0, 0, 256, 0, 256, 6, 6, 2},            // SWIZZLE_32B K-major (32-byte rows)
      {64, 0, 0, 0, 512, 0, 512, 4, 4, 2},            // SWIZZLE_64B K-major
      {64, 0, 0, 11520, 160, 1024, 128, 0, 0, 1},     // single accumulator (dependent MMAs)
      {64, 0, 0, 11520, 160, 1024, 128, 0, 0, 4},
  };
  for (const Case& c : cases) {
    for (int grid : {1, 148}) {
      bench<<<grid, 128, 180 * 1024>>>(c, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("N=%3d a:%s b:%s aLBO=%5d aSBO=%4d bLBO=%4d bSBO=%4d layout=%d/%d nacc=%d grid=%3d : %.1f cycles/MMA\n", c.N,
             c.a_mn ? "MN" : "K ", c.b_mn ? "MN" : "K ", c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, c.a_layout, c.b_layout, c.nacc,
             grid, (double)mx / (iters * 8.0));
    }
  }
  return 0;
}
