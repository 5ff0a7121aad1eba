// Shared device/host helpers for libsaragan_b200 (sm_100a only).
//
// Activation layout ("act"): T act[N][CC][D][H][W][8]  -- channel-blocked, 8 channels per
// 16-byte (bf16) / 32-byte (fp32) vector, CC = 2*ceil(C/16) chunks, pad channels are zero.
// It is the layout the tcgen05 implicit-GEMM reads with plain (un-swizzled) UMMA
// descriptors straight out of a TMA-loaded halo tile, see conv_tc.cu.
// Image layout ("img"):     float img[N][D][H][W]      (the networks' C == 1 boundary)
// Plain layout ("plain"):   float x[N][C][D][H][W]     (torch NCDHW, boundary/debug only)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SG_DTYPE_BF16 0
#define SG_DTYPE_F32 1
#define SG_DTYPE_TF32 2   /* packed-weight flavour only: fp32 [tap][K/4][RP][4], values rounded to tf32 */

// error plumbing ------------------------------------------------------------------------
void sg_set_error(const char* fmt, ...);
int sg_check_launch(const char* what);

#define SG_REQUIRE(cond, ...)             \
  do {                                    \
    if (!(cond)) {                        \
      sg_set_error(__VA_ARGS__);          \
      return -1;                          \
    }                                     \
  } while (0)

static inline int sg_chunks(int C) { return 2 * ((C + 15) / 16); }

static inline int sg_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// grid for a grid-stride kernel over `work` items with `threads` per block: a whole number
// of waves of the SM count, at most 8 blocks per SM.
static inline unsigned sg_grid(int64_t work, int threads) {
  int64_t blocks = (work + threads - 1) / threads;
  int64_t cap = (int64_t)sg_num_sms() * 8;
  if (blocks < 1) blocks = 1;
  if (blocks > cap) blocks = cap;
  return (unsigned)blocks;
}

// 8-wide vector access ---------------------------------------------------------------------
struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ F8 ld8(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }

// Negative slope of the fused LeakyReLU and of its backward masks.  One copy of the constant per translation
// unit (the library is built without relocatable device code); sg_set_leaky_slope() writes all of them.
// Default 0.2 = pgan_pytorch/network.py; network_dict.py uses LEAKINESS = 0.3 (network_dict.py:18) or ReLU (0).
static __constant__ float c_sg_leak = 0.2f;
#define SG_DEFINE_LEAK_SETTER(name) \
  int name(float slope) { return (int)cudaMemcpyToSymbol(c_sg_leak, &slope, sizeof(float)); }
int sg_set_leak_elementwise(float slope);
int sg_set_leak_conv_direct(float slope);
int sg_set_leak_conv_tc(float slope);
int sg_set_leak_conv_tc_wgrad(float slope);
__device__ __forceinline__ float lrelu02(float x) { return x > 0.f ? x : c_sg_leak * x; }
__device__ __forceinline__ float lmask02(float ref) { return ref > 0.f ? 1.f : c_sg_leak; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dtype dispatch: FN is a generic lambda taking a null pointer of the element type
#define SG_DISPATCH(dtype, ...)                                           \
  do {                                                                    \
    if ((dtype) == SG_DTYPE_BF16) {                                       \
      typedef __nv_bfloat16 T;                                            \
      __VA_ARGS__                                                         \
    } else if ((dtype) == SG_DTYPE_F32) {                                 \
      typedef float T;                                                    \
      __VA_ARGS__                                                         \
    } else {                                                              \
      sg_set_error("unsupported dtype %d", (int)(dtype));                 \
      return -1;                                                          \
    }                                                                     \
  } while (0)

// ------------------------------------------------------------------- programmatic dependent launch
// OPTIONAL (sg_set_pdl(1), off by default).  Every kernel of the library goes through sg_launch and starts with
// sg_pdl_enter(): with the programmatic-stream-serialization attribute set, `launch_dependents` lets the
// NEXT kernel of the stream be scheduled as soon as all CTAs of this one have started, `wait` blocks until
// the PREVIOUS kernel has completed and flushed -- so no kernel touches global memory before its producer is
// done, and completion stays transitive along the stream.  (The tcgen05 kernels run barrier init, TMEM
// allocation and tensor-map prefetch before the wait.)  Without the attribute both instructions are no-ops.
// Measured on the cfg3 step (CUDA-graph replay, ~800 kernels): 17.57 ms with the attribute against 17.22 ms
// without -- early-resident dependents cost more than the launch gaps they hide -- hence off by default.
__device__ __forceinline__ void sg_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void sg_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void sg_pdl_enter() {
  sg_pdl_trigger();
  sg_pdl_wait();
}
extern int g_sg_pdl;   // 1 = launch with the attribute, 0 (default) = plain stream order (sg_set_pdl)
template <typename... KArgs, typename... Args>
inline void sg_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_sg_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in sg_check_launch
}
