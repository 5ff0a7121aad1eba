// Microbenchmark: tcgen05.mma (M=128, K=16, kind::f16) with the A operand read from SHARED memory (descriptor) against
// the A operand read from TENSOR memory ([taddr]), B always a shared-memory descriptor, for N = 32 ... 256.
// Question it answers: is the ~46-cycle floor / the 56 cycles at N = 96 the shared-memory read of
// A (4 KB) + B (N*32 B), i.e. does an MMA whose A comes from TMEM run at max(N/2, B bytes / 128)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tmem_a_bench tools/mma_tmem_a_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem),
               "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <bool A_TMEM>
__global__ void __launch_bounds__(128) bench(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // bf16 1.0
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
  // the A operand in tensor memory: 128 lanes x 8 columns of 32 bits (16 bf16 of K) per K step, 16 K steps side by side
  // at columns 256..383; every warp fills its 32 lanes with bf16 ones
  {
    const uint32_t one2 = 0x3c003c00u;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
    for (int c = 0; c < 128; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + (uint32_t)c), "r"(one2) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a0 = make_desc(smem_u32(smem), 11520, 160);
    const uint64_t b0 = make_desc(smem_u32(smem) + 96 * 1024, (uint32_t)N * 16, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (A_TMEM) mma_ts(tmem, tmem + 256u + (uint32_t)(8 * j), b0 + (uint64_t)(j & 1), idesc, 1);
        else mma_ss(tmem, a0 + (uint64_t)(j * 2), b0 + (uint64_t)(j & 1), idesc, 1);
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

template <bool A_TMEM>
double run(int N, long long* out) {
  const int iters = 300;
  cudaFuncSetAttribute(bench<A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<A_TMEM><<<148, 128, 180 * 1024>>>(N, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx / (iters * 16.0);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  printf("tcgen05.mma M=128 K=16 kind::f16, cycles per MMA (148 CTAs, one per SM)\n");
  printf("  N    A from smem   A from TMEM   N/2 (tensor-bound)   (4096 + 32 N)/128   32 N/128\n");
  for (int N : {32, 64, 96, 128, 192, 256}) {
    const double ss = run<false>(N, out), ts = run<true>(N, out);
    printf("%4d   %10.1f   %11.1f   %18.1f   %17.1f   %8.1f\n", N, ss, ts, N / 2.0, (4096 + 32.0 * N) / 128, 32.0 * N / 128);
  }
  return 0;
}
