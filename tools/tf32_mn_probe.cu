// Probe: one tcgen05.mma (M=128, N=32) with MN-major, un-swizzled operands for kind::f16 (bf16, K=16) and kind::tf32
// (K=8) on known integer data; prints the mismatch against the expected product under the layout
//   element (mn, k) at (mn / T) * SBO + (k / 8) * LBO + (k % 8) * 16 + (mn % T) * elt     (T = 16 / elt bytes)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tf32_mn_probe tools/tf32_mn_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// MODE 0: bf16 MN-major, 1: tf32 MN-major, 2: tf32 K-major
template <int MODE>
__global__ void __launch_bounds__(128) probe(float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  constexpr int N = 32, K = MODE == 0 ? 16 : 8;
  constexpr int ELT = MODE == 0 ? 2 : 4, T = 16 / ELT;
  const int warp = threadIdx.x >> 5;
  uint8_t* A = smem;
  uint8_t* B = smem + 32 * 1024;
  // A[m][k] = (m % 7) - 3 + k,  B[n][k] = (n % 5) - 2 + 2 * k   (small integers: exact in bf16 / tf32)
  constexpr uint32_t SBO_A = 256, SBO_B = 256, LBO = 128;   // slots (T channels x 8 k) of 128 B, two per 256 B (second = next k group)
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 128 * K; idx += blockDim.x) {
    const int m = idx / K, k = idx % K;
    const float v = (float)((m % 7) - 3 + k);
    if (MODE == 2) {   // K-major: element (m, k) at (m / 8) * SBO + (k / T) * LBO + (m % 8) * 16 + (k % T) * ELT
      *(float*)(A + (m / 8) * 256 + (k / T) * 128 + (m % 8) * 16 + (k % T) * 4) = v;
    } else {
      uint8_t* p = A + (m / T) * SBO_A + (k / 8) * LBO + (k % 8) * 16 + (m % T) * ELT;
      if (MODE == 0) *(__nv_bfloat16*)p = __float2bfloat16(v); else *(float*)p = v;
    }
  }
  for (int idx = threadIdx.x; idx < N * K; idx += blockDim.x) {
    const int n = idx / K, k = idx % K;
    const float v = (float)((n % 5) - 2 + 2 * k);
    if (MODE == 2) {
      *(float*)(B + (n / 8) * 256 + (k / T) * 128 + (n % 8) * 16 + (k % T) * 4) = v;
    } else {
      uint8_t* p = B + (n / T) * SBO_B + (k / 8) * LBO + (k % 8) * 16 + (n % T) * ELT;
      if (MODE == 0) *(__nv_bfloat16*)p = __float2bfloat16(v); else *(float*)p = v;
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 32) {
    const uint32_t fmt = MODE == 0 ? 1u : 2u;
    const uint32_t mn = MODE == 2 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (mn << 15) | (mn << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a = make_desc(smem_u32(A), 128, 256), b = make_desc(smem_u32(B), 128, 256);
    if (MODE == 0)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(0u) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(a), "l"(b), "r"(idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[threadIdx.x * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

template <int MODE>
void run(const char* name, float* d_out) {
  constexpr int K = MODE == 0 ? 16 : 8;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaMemset(d_out, 0, 128 * 32 * 4);
  probe<MODE><<<1, 128, 64 * 1024>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  static float h[128 * 32];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double worst = 0; int bad = 0, zeros = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      double want = 0;
      for (int k = 0; k < K; ++k) want += ((m % 7) - 3 + k) * (double)((n % 5) - 2 + 2 * k);
      const double d = fabs(h[m * 32 + n] - want);
      if (d > worst) worst = d;
      bad += d > 1e-3;
      zeros += h[m * 32 + n] == 0.f;
    }
  printf("%-18s: max |got - want| = %g, %d / 4096 wrong, %d zeros; row 0: %g %g %g %g  row 5: %g %g %g %g  row 100: %g %g\n", name, worst, bad,
         zeros, h[0], h[1], h[2], h[3], h[160], h[161], h[162], h[163], h[3200], h[3201]);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  run<0>("bf16 MN-major", d);
  run<2>("tf32 K-major", d);
  run<1>("tf32 MN-major", d);
  return 0;
}
