// Microbenchmark: does the A-operand COLLECTOR of tcgen05.mma lift the ~45-cycle floor of an M=128, K=16 MMA with
// N <= 64?  Groups of G consecutive MMAs share one A descriptor (different B, different accumulators) and are issued
// (a) plainly, (b) with .collector::a::fill / ::use / ::lastuse.  Reports cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_collector_bench tools/mma_collector_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
#define MMA(COLL)                                                                                                     \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                      \
               "tcgen05.mma.cta_group::1.kind::f16" COLL " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b),  \
               "r"(idesc), "r"(acc) : "memory")
__device__ __forceinline__ void mma_plain(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { MMA(""); }
__device__ __forceinline__ void mma_fill(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { MMA(".collector::a::fill"); }
__device__ __forceinline__ void mma_use(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { MMA(".collector::a::use"); }
__device__ __forceinline__ void mma_last(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { MMA(".collector::a::lastuse"); }

template <int G, int MODE>
__global__ void __launch_bounds__(128) bench(int N, int iters, long long* out, float* check) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A region: bf16 values that differ per 16-byte unit so that a stale collector would change the result
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    const uint32_t v = 0x3c00u + ((i >> 2) & 7);
    ((uint32_t*)smem)[i] = v | (v << 16);
  }
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem), 11520, 160);
    const uint64_t b0 = make_desc(smem_u32(smem) + 96 * 1024, (uint32_t)N * 16, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) {
        const uint64_t a = a0 + (uint64_t)(grp * 10 + (i & 3));      // a different A per group
#pragma unroll
        for (int j = 0; j < G; ++j) {
          const uint32_t d = tmem + (uint32_t)(((grp & 1) * G + j) * N);   // G accumulators per group, two sets
          const uint64_t b = b0 + (uint64_t)(j * 2 * N);
          const uint32_t acc = (i | (grp >> 1)) ? 1u : 0u;
          if (MODE == 0 || G == 1) mma_plain(d, a, b, idesc, acc);
          else if (j == 0) mma_fill(d, a, b, idesc, acc);
          else if (j == G - 1) mma_last(d, a, b, idesc, acc);
          else mma_use(d, a, b, idesc, acc);
        }
      }
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
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) {
    // checksum of accumulator column 0 and N (first two accumulators) of lane `lane`: must not depend on MODE
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tmem));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r1) : "r"(tmem + (uint32_t)N));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) { check[lane] = __uint_as_float(r0); check[32 + lane] = __uint_as_float(r1); }
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int G, int MODE>
void run(int N, long long* out, float* check) {
  const int iters = 500;
  cudaFuncSetAttribute(bench<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int grid : {1, 148}) {
    bench<G, MODE><<<grid, 128, 180 * 1024>>>(N, iters, out, check);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    long long h[148];
    float c[64];
    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpy(c, check, sizeof(c), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    double cs = 0;
    for (int i = 0; i < 64; ++i) cs += c[i];
    printf("N=%3d group=%d %s grid=%3d : %6.1f cycles/MMA  (checksum %.6e)\n", N, G,
           MODE ? "collector fill/use/lastuse" : "plain                     ", grid, (double)mx / (iters * 4.0 * G), cs);
  }
}

int main() {
  long long* out;
  float* check;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&check, 64 * sizeof(float));
  for (int N : {16, 32, 64, 128}) {
    run<1, 0>(N, out, check);
    if (N <= 64) { run<3, 0>(N, out, check); run<3, 1>(N, out, check); }
    if (N <= 128) { run<2, 0>(N, out, check); run<2, 1>(N, out, check); }
  }
  return 0;
}
