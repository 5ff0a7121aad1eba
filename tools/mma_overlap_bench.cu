// Microbenchmark: cost of tcgen05.mma (M=128, K=16, N=96 / 192) when consecutive MMAs accumulate into (a) the same
// TMEM columns, (b) column ranges SHIFTED by one third of N (the in-place kd stacking of conv_tc_res.cuh: plane q
// covers accumulators q-1..q+1, plane q+1 covers q..q+2), (c) disjoint column ranges.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_overlap_bench tools/mma_overlap_bench.cu
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
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// PATTERN 0: same columns; 1: shifted by N/3 per MMA over 6 positions (wraps); 2: two disjoint ranges alternating;
//         3: shifted, but 9 MMAs on each position before moving on (plane-outer order)
template <int PATTERN>
__global__ void __launch_bounds__(128) bench(int N, int iters, long long* out) {
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem), 11520, 160);
    const uint64_t b0 = make_desc(smem_u32(smem) + 96 * 1024, (uint32_t)N * 16, 128);
    const int third = N / 3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        uint32_t col;
        if (PATTERN == 0) col = 0;
        else if (PATTERN == 1) col = (uint32_t)((j % 6) * third);
        else if (PATTERN == 2) col = (uint32_t)((j & 1) * N);
        else col = (uint32_t)((j / 9) * third + (i & 1) * 2 * third);
        mma(tmem + col, a0 + (uint64_t)(j * 2), b0 + (uint64_t)(j & 1), idesc, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int PATTERN>
void run(int N, long long* out, const char* what) {
  const int iters = 300;
  cudaFuncSetAttribute(bench<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<PATTERN><<<148, 128, 180 * 1024>>>(N, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d %-58s: %6.1f cycles/MMA\n", N, what, (double)mx / (iters * 18.0));
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  for (int N : {96, 192}) {
    run<0>(N, out, "same accumulator columns every MMA");
    run<1>(N, out, "columns shifted by N/3 every MMA (overlapping)");
    run<2>(N, out, "two disjoint ranges alternating");
    run<3>(N, out, "9 MMAs per position, then shifted by N/3 (plane-outer)");
  }
  return 0;
}
