// Memory-bound kernels of the PGAN step: layout conversion, resampling, fade-in blend,
// LeakyReLU mask, pixel-norm, 1x1x1 to/from-RGB, gradient-penalty reductions, tiny linears.
// All are HBM-bound: 16/32-byte vector accesses, consecutive threads on consecutive vectors,
// grid-stride loops sized in whole waves of the SM count, warp-shuffle reductions.
#include <stdarg.h>
#include <string.h>

#include "../../include/saragan_b200.h"
#include "common.cuh"

// ------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

void sg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static unsigned long long g_launches = 0;
extern "C" int64_t sg_launch_count(int reset) {
  unsigned long long v = g_launches;
  if (reset) g_launches = 0;
  return (int64_t)v;
}
int sg_check_launch(const char* what) {
  ++g_launches;   // called exactly once after every kernel launch of this library
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sg_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
int g_sg_pdl = 0;   // measured on cfg3 (graph replay): 17.57 ms/step with the attribute, 17.22 ms without
extern "C" void sg_set_pdl(int on) { g_sg_pdl = on; }
extern "C" const char* sg_last_error(void) { return g_err; }
SG_DEFINE_LEAK_SETTER(sg_set_leak_elementwise)
// host copy of the slope, PER DEVICE (the __constant__ copies live per device; one process per GPU is the intended use,
// but a process that touches a second device must not believe it carries the first one's slope)
static float* sg_leak_slot() {
  static float slots[64];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 64; ++i) slots[i] = 0.2f;
    init = true;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  return &slots[dev & 63];
}
extern "C" float sg_get_leaky_slope(void) { return *sg_leak_slot(); }
extern "C" int sg_set_leaky_slope(float slope) {
  SG_REQUIRE(slope >= 0.f && slope <= 1.f, "leaky slope %g outside [0, 1]", (double)slope);
  float* slot = sg_leak_slot();
  if (slope == *slot) return 0;
  // configuration call, not on the hot path: everything already enqueued keeps the old slope, everything
  // enqueued afterwards (including replays of captured graphs) sees the new one
  cudaError_t e = cudaDeviceSynchronize();
  int rc = (int)e;
  if (!rc) rc = sg_set_leak_elementwise(slope);
  if (!rc) rc = sg_set_leak_conv_direct(slope);
  if (!rc) rc = sg_set_leak_conv_tc(slope);
  if (!rc) rc = sg_set_leak_conv_tc_wgrad(slope);
  if (!rc) rc = (int)cudaDeviceSynchronize();
  if (rc) {
    sg_set_error("sg_set_leaky_slope: %s", cudaGetErrorString((cudaError_t)rc));
    return rc;
  }
  *slot = slope;
  return 0;
}
extern "C" int sg_version(void) { return 200; }

// -------------------------------------------------------------------- layout conversion
// plain [N][C][V] fp32 -> act [N][CC][V][8] (pad channels zero)
template <typename T>
__global__ void k_plain_to_act(const float* __restrict__ src, T* __restrict__ dst, int N, int C,
                               int CC, int64_t V) {
  sg_pdl_enter();
  int64_t total = (int64_t)N * CC * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i % V;
    int64_t t = i / V;
    int cc = (int)(t % CC);
    int n = (int)(t / CC);
    F8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = cc * 8 + j;
      r.v[j] = c < C ? src[((int64_t)n * C + c) * V + v] : 0.f;
    }
    st8(dst + i * 8, r);
  }
}
template <typename T>
__global__ void k_act_to_plain(const T* __restrict__ src, float* __restrict__ dst, int N, int C,
                               int CC, int64_t V) {
  sg_pdl_enter();
  int64_t total = (int64_t)N * CC * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i % V;
    int64_t t = i / V;
    int cc = (int)(t % CC);
    int n = (int)(t / CC);
    F8 r = ld8(src + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = cc * 8 + j;
      if (c < C) dst[((int64_t)n * C + c) * V + v] = r.v[j];
    }
  }
}

extern "C" int sg_plain_to_act(const float* plain, void* act, int dtype, int N, int C, int64_t V,
                               cudaStream_t s) {
  int CC = sg_chunks(C);
  int64_t total = (int64_t)N * CC * V;
  if (total == 0) return 0;
  SG_DISPATCH(dtype, sg_launch((k_plain_to_act<T>), sg_grid(total, 256), 256, 0, s, plain, (T*)act, N, C, CC, V););
  return sg_check_launch("sg_plain_to_act");
}
extern "C" int sg_act_to_plain(const void* act, float* plain, int dtype, int N, int C, int64_t V,
                               cudaStream_t s) {
  int CC = sg_chunks(C);
  int64_t total = (int64_t)N * CC * V;
  if (total == 0) return 0;
  SG_DISPATCH(dtype, sg_launch((k_act_to_plain<T>), sg_grid(total, 256), 256, 0, s, (const T*)act, plain, N, C, CC, V););
  return sg_check_launch("sg_act_to_plain");
}

// ----------------------------------------------------------------------- elementwise
// y = alpha*a + beta*b   (b may be null).  Fade-in blend (network.py:185,281) and its
// backward, instance noise (train.py:144), gradient scaling.
// coef (nullable): alpha = coef[0], beta = coef[1] read from device memory -- the fade-in alpha of a captured CUDA
// graph follows the schedule (train.py:33,63,82) without a re-capture and a tensor alpha costs no host sync.
template <typename T>
__global__ void k_lincomb(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y,
                          int64_t nvec, int64_t n, float alpha, float beta, const float* __restrict__ coef) {
  sg_pdl_enter();
  if (coef) {
    alpha = __ldg(coef);
    beta = __ldg(coef + 1);
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    F8 x = ld8(a + i * 8);
    if (b) {
      F8 z = ld8(b + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) x.v[j] = alpha * x.v[j] + beta * z.v[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x.v[j] = alpha * x.v[j];
    }
    st8(y + i * 8, x);
  }
  // scalar tail (n not a multiple of 8): handled by the first threads of block 0
  int64_t tail0 = nvec * 8;
  if (blockIdx.x == 0) {
    for (int64_t i = tail0 + threadIdx.x; i < n; i += blockDim.x) {
      float x = alpha * ld1(a + i) + (b ? beta * ld1(b + i) : 0.f);
      st1(y + i, x);
    }
  }
}
extern "C" int sg_lincomb(const void* a, const void* b, void* y, int dtype, int64_t n, float alpha,
                          float beta, cudaStream_t s) {
  if (n == 0) return 0;
  int64_t nvec = n / 8;
  SG_DISPATCH(dtype, sg_launch((k_lincomb<T>), sg_grid(nvec > 0 ? nvec : 1, 256), 256, 0, s,
                         (const T*)a, (const T*)b, (T*)y, nvec, n, alpha, beta, (const float*)nullptr););
  return sg_check_launch("sg_lincomb");
}
extern "C" int sg_lincomb_dev(const void* a, const void* b, void* y, int dtype, int64_t n, const float* coef,
                              cudaStream_t s) {
  SG_REQUIRE(coef != nullptr, "sg_lincomb_dev: coef must point at {alpha, beta} in device memory");
  if (n == 0) return 0;
  int64_t nvec = n / 8;
  SG_DISPATCH(dtype, sg_launch((k_lincomb<T>), sg_grid(nvec > 0 ? nvec : 1, 256), 256, 0, s,
                         (const T*)a, (const T*)b, (T*)y, nvec, n, 0.f, 0.f, coef););
  return sg_check_launch("sg_lincomb_dev");
}

// mode 0: y = lrelu(x)            (network.py:89 etc.)
// mode 1: y = x * m(ref), m = 1 if ref > 0 else 0.2   (LeakyReLU backward and its double
//         backward, SURVEY Appendix B: the mask is recovered from the layer OUTPUT's sign)
template <typename T>
__global__ void k_lrelu(const T* __restrict__ x, const T* __restrict__ ref, T* __restrict__ y,
                        int64_t nvec, int mode) {
  sg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    F8 a = ld8(x + i * 8);
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] = lrelu02(a.v[j]);
    } else {
      F8 r = ld8(ref + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) a.v[j] *= lmask02(r.v[j]);
    }
    st8(y + i * 8, a);
  }
}
extern "C" int sg_lrelu_fwd(const void* x, void* y, int dtype, int64_t n, cudaStream_t s) {
  SG_REQUIRE(n % 8 == 0, "sg_lrelu_fwd: n must be a multiple of 8");
  if (n == 0) return 0;
  SG_DISPATCH(dtype, sg_launch((k_lrelu<T>), sg_grid(n / 8, 256), 256, 0, s, (const T*)x, (const T*)nullptr, (T*)y, n / 8, 0););
  return sg_check_launch("sg_lrelu_fwd");
}
extern "C" int sg_mask_mul(const void* g, const void* ref, void* y, int dtype, int64_t n,
                           cudaStream_t s) {
  SG_REQUIRE(n % 8 == 0, "sg_mask_mul: n must be a multiple of 8");
  if (n == 0) return 0;
  SG_DISPATCH(dtype, sg_launch((k_lrelu<T>), sg_grid(n / 8, 256), 256, 0, s, (const T*)g, (const T*)ref, (T*)y, n / 8, 1););
  return sg_check_launch("sg_mask_mul");
}

// ----------------------------------------------------------------------- resampling
// x: [P][D][H][W][VEC] -> y: [P][D/2][H/2][W/2][VEC], y = scale * sum over the 2x2x2 block.
// avg-pool (network.py:90,154) = scale 1/8; backward of nearest-upsample = scale 1.
template <typename T, typename TO, int VEC>
__global__ void k_down2(const T* __restrict__ x, TO* __restrict__ y, int64_t P, int D, int H, int W,
                        float scale) {
  sg_pdl_enter();
  int Do = D / 2, Ho = H / 2, Wo = W / 2;
  int64_t total = P * Do * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int wo = (int)(i % Wo);
    int64_t t = i / Wo;
    int ho = (int)(t % Ho);
    t /= Ho;
    int d_o = (int)(t % Do);
    int64_t p = t / Do;
    const T* base = x + ((((p * D + 2 * d_o) * H + 2 * ho) * (int64_t)W) + 2 * wo) * VEC;
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const T* q = base + (((int64_t)dz * H + dy) * W + dx) * VEC;
          if (VEC == 8) {
            F8 r = ld8(q);
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[j] += r.v[j];
          } else {
            acc[0] += ld1(q);
          }
        }
    if (VEC == 8) {
      F8 r;
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] = acc[j % VEC] * scale;
      st8(y + i * 8, r);
    } else {
      st1(y + i, acc[0] * scale);
    }
  }
}
// x: [P][D][H][W][VEC] -> y: [P][2D][2H][2W][VEC], y[child] = scale * x[parent].
// nearest upsample (network.py:203,265) = scale 1; backward of avg-pool = scale 1/8.
// mask_ref (nullable, shaped like y): y *= 1 / 0.2 by its sign -- the LeakyReLU backward of the
// conv that fed an avg-pool, fused into the pool's backward.
template <typename T, typename TO, int VEC>
__global__ void k_up2(const T* __restrict__ x, TO* __restrict__ y, const TO* __restrict__ mask_ref, int64_t P,
                      int D, int H, int W, float scale) {
  sg_pdl_enter();
  int64_t total = P * D * H * W;
  int H2 = 2 * H, W2 = 2 * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (the entry point checks total < 2^31): 64-bit div/mod cost more than the stores
    const unsigned ii = (unsigned)i;
    const unsigned w = ii % (unsigned)W;
    unsigned t = ii / (unsigned)W;
    const unsigned h = t % (unsigned)H;
    t /= (unsigned)H;
    const unsigned d = t % (unsigned)D;
    const int64_t p = t / (unsigned)D;
    TO* base = y + ((((p * 2 * D + 2 * d) * H2 + 2 * h) * (int64_t)W2) + 2 * w) * VEC;
    if (VEC == 8) {
      F8 r = ld8(x + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) r.v[j] *= scale;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int64_t off = (((int64_t)dz * H2 + dy) * W2 + dx) * 8;
            if (mask_ref) {
              const F8 m = ld8(mask_ref + (base - y) + off);
              F8 o;
#pragma unroll
              for (int j = 0; j < 8; ++j) o.v[j] = r.v[j] * lmask02(m.v[j]);
              st8(base + off, o);
            } else {
              st8(base + off, r);
            }
          }
    } else {
      float r = ld1(x + i) * scale;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) st1(base + (((int64_t)dz * H2 + dy) * W2 + dx), r);
    }
  }
}
// in/out element types may differ (the 1x4x4 base level is kept in fp32, config.py)
#define SG_DISPATCH2(din, dout, ...)                                         \
  do {                                                                       \
    if ((din) == SG_DTYPE_BF16 && (dout) == SG_DTYPE_BF16) {                 \
      typedef __nv_bfloat16 T; typedef __nv_bfloat16 TO; __VA_ARGS__         \
    } else if ((din) == SG_DTYPE_BF16 && (dout) == SG_DTYPE_F32) {           \
      typedef __nv_bfloat16 T; typedef float TO; __VA_ARGS__                 \
    } else if ((din) == SG_DTYPE_F32 && (dout) == SG_DTYPE_BF16) {           \
      typedef float T; typedef __nv_bfloat16 TO; __VA_ARGS__                 \
    } else if ((din) == SG_DTYPE_F32 && (dout) == SG_DTYPE_F32) {            \
      typedef float T; typedef float TO; __VA_ARGS__                         \
    } else {                                                                 \
      sg_set_error("unsupported dtype pair %d/%d", (int)(din), (int)(dout)); \
      return -1;                                                             \
    }                                                                        \
  } while (0)

extern "C" int sg_down2(const void* x, void* y, int dtype_in, int dtype_out, int vec, int64_t P,
                        int D, int H, int W, float scale, cudaStream_t s) {
  SG_REQUIRE(D % 2 == 0 && H % 2 == 0 && W % 2 == 0, "sg_down2: odd extent %dx%dx%d", D, H, W);
  SG_REQUIRE(vec == 8 || vec == 1, "sg_down2: vec must be 1 or 8");
  int64_t total = P * (D / 2) * (H / 2) * (W / 2);
  if (total == 0) return 0;
  unsigned g = sg_grid(total, 256);
  if (vec == 8) {
    SG_DISPATCH2(dtype_in, dtype_out, sg_launch((k_down2<T, TO, 8>), g, 256, 0, s, (const T*)x, (TO*)y, P, D, H, W, scale););
  } else {
    SG_DISPATCH2(dtype_in, dtype_out, sg_launch((k_down2<T, TO, 1>), g, 256, 0, s, (const T*)x, (TO*)y, P, D, H, W, scale););
  }
  return sg_check_launch("sg_down2");
}
extern "C" int sg_up2(const void* x, void* y, const void* mask_ref, int dtype_in, int dtype_out, int vec,
                      int64_t P, int D, int H, int W, float scale, cudaStream_t s) {
  SG_REQUIRE(vec == 8 || vec == 1, "sg_up2: vec must be 1 or 8");
  SG_REQUIRE(mask_ref == nullptr || vec == 8, "sg_up2: mask_ref needs an activation tensor (vec 8)");
  int64_t total = P * D * H * W;
  if (total == 0) return 0;
  SG_REQUIRE(total < (1ll << 31), "sg_up2: %lld input vectors exceed the 32-bit index range", (long long)total);
  unsigned g = sg_grid(total, 256);
  if (vec == 8) {
    SG_DISPATCH2(dtype_in, dtype_out, sg_launch((k_up2<T, TO, 8>), g, 256, 0, s, (const T*)x, (TO*)y, (const TO*)mask_ref, P, D, H, W, scale););
  } else {
    SG_DISPATCH2(dtype_in, dtype_out, sg_launch((k_up2<T, TO, 1>), g, 256, 0, s, (const T*)x, (TO*)y, (const TO*)nullptr, P, D, H, W, scale););
  }
  return sg_check_launch("sg_up2");
}

// ------------------------------------------------------------------------ pixel-norm
// y = x * rsqrt(mean_c(x^2) + eps)   (network.py:196-197), optionally followed by LeakyReLU
// (GeneratorBlock's second conv: conv -> pixel-norm -> lrelu, network.py:214-216).
// LPV = lanes per voxel: 1 (a thread walks all channel chunks of its voxel) at the high-resolution levels;
// 32 (a warp shares the chunks, shuffle reduction) at the low-resolution levels, where N*V is a few hundred
// and C = 512 -- a thread per voxel would leave 64 dependent 16-byte loads on one thread of one block.
template <typename T, int LPV>
__global__ void k_pixelnorm_fwd(const T* __restrict__ x, T* __restrict__ y, int N, int C, int CC,
                                int64_t V, float eps, int lrelu_after) {
  sg_pdl_enter();
  int64_t total = (int64_t)N * V;
  float invC = 1.f / (float)C;
  const int lane = LPV == 1 ? 0 : (int)(threadIdx.x & (LPV - 1));
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPV; i < total;
       i += (int64_t)gridDim.x * blockDim.x / LPV) {
    int64_t v = i % V;
    int64_t n = i / V;
    const T* px = x + (n * CC * V + v) * 8;
    T* py = y + (n * CC * V + v) * 8;
    float ss = 0.f;
    for (int cc = lane; cc < CC; cc += LPV) {
      F8 r = ld8(px + (int64_t)cc * V * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += r.v[j] * r.v[j];
    }
    if (LPV > 1) ss = warp_sum(ss);
    float r_ = rsqrtf(ss * invC + eps);
    for (int cc = lane; cc < CC; cc += LPV) {
      F8 r = ld8(px + (int64_t)cc * V * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float o = r.v[j] * r_;
        r.v[j] = lrelu_after ? lrelu02(o) : o;
      }
      st8(py + (int64_t)cc * V * 8, r);
    }
  }
}
// g' = gy * m(x) if lrelu_after;  gx = r*g' - x * r^3 * mean_c(x*g')   
template <typename T, int LPV>
__global__ void k_pixelnorm_bwd(const T* __restrict__ x, const T* __restrict__ gy,
                                T* __restrict__ gx, int N, int C, int CC, int64_t V, float eps,
                                int lrelu_after, int mask_input) {
  sg_pdl_enter();
  int64_t total = (int64_t)N * V;
  float invC = 1.f / (float)C;
  const int lane = LPV == 1 ? 0 : (int)(threadIdx.x & (LPV - 1));
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPV; i < total;
       i += (int64_t)gridDim.x * blockDim.x / LPV) {
    int64_t v = i % V;
    int64_t n = i / V;
    int64_t off = (n * CC * V + v) * 8;
    float ss = 0.f, xg = 0.f;
    for (int cc = lane; cc < CC; cc += LPV) {
      F8 a = ld8(x + off + (int64_t)cc * V * 8);
      F8 g = ld8(gy + off + (int64_t)cc * V * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = lrelu_after ? g.v[j] * lmask02(a.v[j]) : g.v[j];
        ss += a.v[j] * a.v[j];
        xg += a.v[j] * gj;
      }
    }
    if (LPV > 1) {
      ss = warp_sum(ss);
      xg = warp_sum(xg);
    }
    float r_ = rsqrtf(ss * invC + eps);
    float k = r_ * r_ * r_ * xg * invC;
    for (int cc = lane; cc < CC; cc += LPV) {
      F8 a = ld8(x + off + (int64_t)cc * V * 8);
      F8 g = ld8(gy + off + (int64_t)cc * V * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = lrelu_after ? g.v[j] * lmask02(a.v[j]) : g.v[j];
        float o = r_ * gj - a.v[j] * k;
        // mask_input: x is the LeakyReLU output of the producing conv; fold that conv's mask in
        a.v[j] = mask_input ? o * lmask02(a.v[j]) : o;
      }
      st8(gx + off + (int64_t)cc * V * 8, a);
    }
  }
}
// a warp per voxel pays off when the voxels alone cannot fill the GPU and there are chunks to share
static bool sg_warp_per_voxel(int64_t total, int CC) { return total <= 16384 && CC >= 8; }
extern "C" int sg_pixelnorm_fwd(const void* x, void* y, int dtype, int N, int C, int64_t V,
                                float eps, int lrelu_after, cudaStream_t s) {
  int64_t total = (int64_t)N * V;
  if (total == 0) return 0;
  int CC = sg_chunks(C);
  if (sg_warp_per_voxel(total, CC)) {
    SG_DISPATCH(dtype, sg_launch((k_pixelnorm_fwd<T, 32>), sg_grid(total * 32, 256), 256, 0, s, (const T*)x, (T*)y, N, C, CC, V, eps, lrelu_after););
  } else {
    SG_DISPATCH(dtype, sg_launch((k_pixelnorm_fwd<T, 1>), sg_grid(total, 256), 256, 0, s, (const T*)x, (T*)y, N, C, CC, V, eps, lrelu_after););
  }
  return sg_check_launch("sg_pixelnorm_fwd");
}
extern "C" int sg_pixelnorm_bwd(const void* x, const void* gy, void* gx, int dtype, int N, int C,
                                int64_t V, float eps, int lrelu_after, int mask_input, cudaStream_t s) {
  int64_t total = (int64_t)N * V;
  if (total == 0) return 0;
  int CC = sg_chunks(C);
  if (sg_warp_per_voxel(total, CC)) {
    SG_DISPATCH(dtype, sg_launch((k_pixelnorm_bwd<T, 32>), sg_grid(total * 32, 256), 256, 0, s, (const T*)x, (const T*)gy, (T*)gx, N, C, CC, V, eps, lrelu_after, mask_input););
  } else {
    SG_DISPATCH(dtype, sg_launch((k_pixelnorm_bwd<T, 1>), sg_grid(total, 256), 256, 0, s, (const T*)x, (const T*)gy, (T*)gx, N, C, CC, V, eps, lrelu_after, mask_input););
  }
  return sg_check_launch("sg_pixelnorm_bwd");
}

// ------------------------------------------------------------------ 1x1x1 to/from RGB
// FromRGB (network.py:101-110): y[n][c][v] = act(scale*w[c]*img[n][v] + bias[c])
// grid.y = (n, chunk): the 8 weights / biases of the chunk live in registers.  Consecutive LANES own consecutive
// voxels, so every store instruction of a warp writes one contiguous run (512 B of bf16); four voxels per thread
// and iteration, 32 voxels apart, keep four stores in flight.  (v2 gave a thread four CONSECUTIVE voxels: each
// store instruction then touched 32 different 64-byte segments with 16 bytes each -- 1.5 TB/s; v1 spent 16 __ldg
// and two 64-bit divisions per 16 bytes written.)
// mask_ref (nullable, shaped like y): y *= (mask_ref > 0 ? 1 : slope) -- the LeakyReLU mask of the gradient
// penalty's double backward riding in the producer of the masked tensor.
template <typename T>
__global__ void __launch_bounds__(256)
k_pw_expand(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
            const T* __restrict__ mask_ref, T* __restrict__ y, int N, int C, int CC, int64_t V, float scale,
            int lrelu) {
  sg_pdl_enter();
  const int row = blockIdx.y, cc = row % CC, n = row / CC;
  float ws[8], bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cc * 8 + j;
    ws[j] = c < C ? scale * __ldg(w + c) : 0.f;
    bs[j] = (c < C && bias) ? __ldg(bias + c) : 0.f;
  }
  const float* pi = img + (int64_t)n * V;
  T* py = y + (int64_t)row * V * 8;
  const T* pm = mask_ref ? mask_ref + (int64_t)row * V * 8 : nullptr;
  const int64_t step = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; v0 < V; v0 += step) {
    float p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t v = v0 + (int64_t)q * blockDim.x;
      p[q] = v < V ? __ldg(pi + v) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t v = v0 + (int64_t)q * blockDim.x;
      if (v < V) {
        F8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o = fmaf(ws[j], p[q], bs[j]);
          r.v[j] = lrelu ? lrelu02(o) : o;
        }
        if (pm) {
          const F8 m = ld8(pm + v * 8);
#pragma unroll
          for (int j = 0; j < 8; ++j) r.v[j] *= lmask02(m.v[j]);
        }
        st8(py + v * 8, r);
      }
    }
  }
}
// ToRGB (network.py:219-225): img[n][v] = scale*sum_c w[c]*x[n][c][v] + bias[0]
template <typename T, int LPV>
__global__ void k_pw_reduce(const T* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ img, int N, int C,
                            int CC, int64_t V, float scale) {
  sg_pdl_enter();
  int64_t total = (int64_t)N * V;
  const int lane = LPV == 1 ? 0 : (int)(threadIdx.x & (LPV - 1));
  // the C weights once per block in shared memory (zero for pad channels): broadcast reads instead of 8 global
  // loads per 16 bytes of activations
  extern __shared__ float w_s[];
  for (int c = threadIdx.x; c < CC * 8; c += blockDim.x) w_s[c] = c < C ? w[c] : 0.f;
  __syncthreads();
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPV; i < total;
       i += (int64_t)gridDim.x * blockDim.x / LPV) {
    int64_t v = i % V;
    int64_t n = i / V;
    const T* px = x + (n * CC * V + v) * 8;
    float acc = 0.f;
    auto dot8 = [&](const F8& r, int cc) {
      const float4 w0 = *reinterpret_cast<const float4*>(w_s + cc * 8), w1 = *reinterpret_cast<const float4*>(w_s + cc * 8 + 4);
      return w0.x * r.v[0] + w0.y * r.v[1] + w0.z * r.v[2] + w0.w * r.v[3] + w1.x * r.v[4] + w1.y * r.v[5] +
             w1.z * r.v[6] + w1.w * r.v[7];
    };
    int cc = lane;
    // four chunk loads in flight per thread (the chunks of a voxel are V*16 bytes apart: independent requests)
    for (; cc + 3 * LPV < CC; cc += 4 * LPV) {
      const F8 r0 = ld8(px + (int64_t)cc * V * 8), r1 = ld8(px + (int64_t)(cc + LPV) * V * 8),
               r2 = ld8(px + (int64_t)(cc + 2 * LPV) * V * 8), r3 = ld8(px + (int64_t)(cc + 3 * LPV) * V * 8);
      acc += dot8(r0, cc) + dot8(r1, cc + LPV) + dot8(r2, cc + 2 * LPV) + dot8(r3, cc + 3 * LPV);
    }
    for (; cc < CC; cc += LPV) acc += dot8(ld8(px + (int64_t)cc * V * 8), cc);
    if (LPV > 1) acc = warp_sum(acc);
    if (lane == 0) img[i] = scale * acc + (bias ? __ldg(bias) : 0.f);
  }
}
// gw[c] = scale * sum_{n,v} g[n][c][v]*img[n][v]  (img == null: plain channel sum),
// gb[c] = sum_{n,v} g[n][c][v].  grid = (voxel tiles, chunk, sample): a block walks tiles of 1024 voxels of its
// (chunk, sample) plane with the stride of the grid, so the blocks resident at any moment read ONE contiguous
// window per plane (private contiguous slabs per block -- 592 separate streams -- ran at 2.1 TB/s); four 16-byte
// loads in flight per thread, no bounds checks in the main loop, warp-shuffle + one atomic per channel per block.
// Outputs must be zeroed by the caller (the entry point does).
template <typename T>
__global__ void __launch_bounds__(256)
k_pw_wgrad(const T* __restrict__ g, const float* __restrict__ img, float* __restrict__ gw,
           float* __restrict__ gb, int N, int C, int CC, int64_t V, float scale) {
  sg_pdl_enter();
  const int cc = blockIdx.y, n = blockIdx.z;
  const T* pg = g + ((int64_t)n * CC + cc) * V * 8;
  const float* pi = img ? img + (int64_t)n * V : nullptr;
  float aw[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) aw[j] = ab[j] = 0.f;
  const int64_t n_full = V / 1024;
  for (int64_t tile = blockIdx.x; tile < n_full; tile += gridDim.x) {
    const int64_t v0 = tile * 1024 + threadIdx.x;
    F8 r[4];
    float p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      r[q] = ld8(pg + (v0 + q * 256) * 8);
      p[q] = pi ? __ldg(pi + v0 + q * 256) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        aw[j] = fmaf(r[q].v[j], p[q], aw[j]);
        ab[j] += r[q].v[j];
      }
  }
  if (blockIdx.x == 0) {   // ragged tail of the plane (V not a multiple of 1024)
    for (int64_t v = n_full * 1024 + threadIdx.x; v < V; v += 256) {
      const F8 r = ld8(pg + v * 8);
      const float p = pi ? pi[v] : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        aw[j] = fmaf(r.v[j], p, aw[j]);
        ab[j] += r.v[j];
      }
    }
  }
  __shared__ float sm[2][8][8];  // [w|b][warp][j]
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = warp_sum(aw[j]);
    float b = warp_sum(ab[j]);
    if (lane == 0) {
      sm[0][warp][j] = a;
      sm[1][warp][j] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    int which = threadIdx.x >> 3, j = threadIdx.x & 7;
    float tot = 0.f;
    for (int wp = 0; wp < (int)(blockDim.x >> 5); ++wp) tot += sm[which][wp][j];
    int c = cc * 8 + j;
    if (c < C) {
      if (which == 0 && gw) atomicAdd(gw + c, tot * scale);
      if (which == 1 && gb) atomicAdd(gb + c, tot);
    }
  }
}
static int pw_expand_launch(const float* img, const float* w, const float* bias, const void* mask_ref, void* y,
                            int dtype, int N, int C, int64_t V, float scale, int lrelu, cudaStream_t s) {
  int CC = sg_chunks(C);
  int64_t total = (int64_t)N * CC * V;
  if (total == 0) return 0;
  SG_REQUIRE((int64_t)N * CC <= 65535, "sg_pw_expand: N * channel chunks = %lld exceeds the grid", (long long)N * CC);
  const int64_t per_row = (V + 1023) / 1024;       // blocks of 256 threads x 4 voxels
  int64_t gx = per_row;
  const int64_t cap = ((int64_t)sg_num_sms() * 8 + (int64_t)N * CC - 1) / ((int64_t)N * CC);   // ~8 blocks per SM in total
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  SG_DISPATCH(dtype, sg_launch((k_pw_expand<T>), dim3((unsigned)gx, (unsigned)(N * CC)), 256, 0, s, img, w, bias,
                               (const T*)mask_ref, (T*)y, N, C, CC, V, scale, lrelu););
  return sg_check_launch("sg_pw_expand");
}
extern "C" int sg_pw_expand(const float* img, const float* w, const float* bias, void* y, int dtype,
                            int N, int C, int64_t V, float scale, int lrelu, cudaStream_t s) {
  return pw_expand_launch(img, w, bias, nullptr, y, dtype, N, C, V, scale, lrelu, s);
}
extern "C" int sg_pw_expand_masked(const float* img, const float* w, const float* bias, const void* mask_ref, void* y,
                                   int dtype, int N, int C, int64_t V, float scale, int lrelu, cudaStream_t s) {
  return pw_expand_launch(img, w, bias, mask_ref, y, dtype, N, C, V, scale, lrelu, s);
}
extern "C" int sg_pw_reduce(const void* x, const float* w, const float* bias, float* img, int dtype,
                            int N, int C, int64_t V, float scale, cudaStream_t s) {
  int CC = sg_chunks(C);
  int64_t total = (int64_t)N * V;
  if (total == 0) return 0;
  if (sg_warp_per_voxel(total, CC)) {
    SG_DISPATCH(dtype, sg_launch((k_pw_reduce<T, 32>), sg_grid(total * 32, 256), 256, CC * 8 * sizeof(float), s, (const T*)x, w, bias, img, N, C, CC, V, scale););
  } else {
    SG_DISPATCH(dtype, sg_launch((k_pw_reduce<T, 1>), sg_grid(total, 256), 256, CC * 8 * sizeof(float), s, (const T*)x, w, bias, img, N, C, CC, V, scale););
  }
  return sg_check_launch("sg_pw_reduce");
}
extern "C" int sg_pw_wgrad(const void* g, const float* img, float* gw, float* gb, int dtype, int N,
                           int C, int64_t V, float scale, cudaStream_t s) {
  int CC = sg_chunks(C);
  int64_t total = (int64_t)N * V;
  if (gw) cudaMemsetAsync(gw, 0, sizeof(float) * C, s);
  if (gb) cudaMemsetAsync(gb, 0, sizeof(float) * C, s);
  if (total == 0) return 0;
  // ~8 resident blocks per SM in total, never more blocks per plane than it has 1024-voxel tiles
  SG_REQUIRE(N <= 65535, "sg_pw_wgrad: batch %d exceeds the grid", N);
  int64_t gx = ((int64_t)sg_num_sms() * 8 + (int64_t)CC * N - 1) / ((int64_t)CC * N);
  const int64_t tiles = V / 1024 > 0 ? V / 1024 : 1;
  if (gx > tiles) gx = tiles;
  dim3 grid((unsigned)gx, (unsigned)CC, (unsigned)N);
  SG_DISPATCH(dtype, sg_launch((k_pw_wgrad<T>), grid, 256, 0, s, (const T*)g, img, gw, gb, N, C, CC, V, scale););
  return sg_check_launch("sg_pw_wgrad");
}

// ------------------------------------------------------------ gradient-penalty helpers
// out[n] = sum_v x[n][v]^2   (loss.py:25-26: per-sample squared L2 norm of the input gradient)
__global__ void k_sumsq_rows(const float* __restrict__ x, float* __restrict__ out, int64_t V,
                             int64_t per_slab) {
  sg_pdl_enter();
  int n = blockIdx.y;
  int64_t lo = blockIdx.x * per_slab;
  int64_t hi = lo + per_slab < V ? lo + per_slab : V;
  const float* p = x + (int64_t)n * V;
  float acc = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float a = p[i];
    acc += a * a;
  }
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(out + n, t);
  }
}
extern "C" int sg_sumsq_rows(const float* x, float* out, int N, int64_t V, cudaStream_t s) {
  if (N == 0) return 0;
  cudaMemsetAsync(out, 0, sizeof(float) * N, s);
  if (V == 0) return 0;
  int64_t slabs = (V + 8191) / 8192;
  int64_t want = ((int64_t)sg_num_sms() * 4 + N - 1) / N;
  if (slabs > want) slabs = want;
  int64_t per = (V + slabs - 1) / slabs;
  slabs = (V + per - 1) / per;
  sg_launch((k_sumsq_rows), dim3((unsigned)slabs, (unsigned)N), 256, 0, s, x, out, V, per);
  return sg_check_launch("sg_sumsq_rows");
}
// y[n][v] = s[n] * x[n][v]
__global__ void k_rowscale(const float* __restrict__ x, const float* __restrict__ sc,
                           float* __restrict__ y, int64_t V, int64_t total) {
  sg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * __ldg(sc + i / V);
}
extern "C" int sg_rowscale(const float* x, const float* sc, float* y, int N, int64_t V,
                           cudaStream_t s) {
  int64_t total = (int64_t)N * V;
  if (total == 0) return 0;
  sg_launch((k_rowscale), sg_grid(total, 256), 256, 0, s, x, sc, y, V, total);
  return sg_check_launch("sg_rowscale");
}
// out = eps[n]*real + (1-eps[n])*fake   (loss.py:13)
__global__ void k_interp(const float* __restrict__ real, const float* __restrict__ fake,
                         const float* __restrict__ eps, float* __restrict__ out, int64_t V,
                         int64_t total) {
  sg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float e = __ldg(eps + i / V);
    out[i] = e * real[i] + (1.f - e) * fake[i];
  }
}
extern "C" int sg_interp(const float* real, const float* fake, const float* eps, float* out, int N,
                         int64_t V, cudaStream_t s) {
  int64_t total = (int64_t)N * V;
  if (total == 0) return 0;
  sg_launch((k_interp), sg_grid(total, 256), 256, 0, s, real, fake, eps, out, V, total);
  return sg_check_launch("sg_interp");
}

// -------------------------------------------------------------- tiny-batch linears (fp32)
// EqualizedLinear (network.py:59-77) with batch <= a few dozen rows: weight-bandwidth bound.
// y[b][o] = act(scale * sum_i x[b][i]*w[o][i] + bias[o]);  one warp per output feature.
#define LIN_BT 8
// one 256-thread block per output feature streams that feature's weight row with 16-byte loads, four in
// flight per thread (the first version's scalar loads left the 16.8 MB of D's 8192 -> 512 linear at
// 0.36 TB/s); scalar tail / fallback when In is not a multiple of 4
__global__ void __launch_bounds__(256)
k_linear_fwd(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
             float* __restrict__ y, int B, int In, int Out, float scale, int lrelu) {
  sg_pdl_enter();
  const int o = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* wr = w + (int64_t)o * In;
  const bool vec = (In & 3) == 0;
  const int In4 = vec ? In >> 2 : 0;
  __shared__ float red[8][LIN_BT];
  for (int b0 = 0; b0 < B; b0 += LIN_BT) {
    float acc[LIN_BT];
#pragma unroll
    for (int k = 0; k < LIN_BT; ++k) acc[k] = 0.f;
    const float4* wr4 = reinterpret_cast<const float4*>(wr);
#pragma unroll 4
    for (int i = threadIdx.x; i < In4; i += 256) {
      const float4 wv = __ldcs(wr4 + i);
#pragma unroll
      for (int k = 0; k < LIN_BT; ++k)
        if (b0 + k < B) {
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(b0 + k) * In) + i);
          acc[k] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[k]))));
        }
    }
    for (int i = 4 * In4 + threadIdx.x; i < In; i += 256) {
      const float wv = wr[i];
#pragma unroll
      for (int k = 0; k < LIN_BT; ++k)
        if (b0 + k < B) acc[k] = fmaf(wv, __ldg(x + (int64_t)(b0 + k) * In + i), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < LIN_BT; ++k) {
      const float t = warp_sum(acc[k]);
      if (lane == 0) red[warp][k] = t;
    }
    __syncthreads();
    if (threadIdx.x < LIN_BT && b0 + threadIdx.x < B) {
      const int k = threadIdx.x;
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) t += red[wq][k];
      const float r = scale * t + (bias ? bias[o] : 0.f);
      y[(int64_t)(b0 + k) * Out + o] = lrelu ? lrelu02(r) : r;
    }
    __syncthreads();
  }
}
// gx[b][i] = scale * sum_o g[b][o]*w[o][i];  a thread owns four consecutive input features (16-byte weight
// loads, coalesced across the warp), the outputs are split over gridDim.y slices with atomics (gx zeroed by
// the entry point).  V = 4 needs In % 4 == 0; V = 1 is the general form.
template <int V>
__global__ void k_linear_dgrad(const float* __restrict__ g, const float* __restrict__ w,
                               float* __restrict__ gx, int B, int In, int Out, float scale,
                               int o_per) {
  sg_pdl_enter();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * V;
  if (i >= In) return;
  const int o_lo = blockIdx.y * o_per;
  const int o_hi = o_lo + o_per < Out ? o_lo + o_per : Out;
  for (int b0 = 0; b0 < B; b0 += LIN_BT) {
    float acc[LIN_BT][V];
#pragma unroll
    for (int k = 0; k < LIN_BT; ++k)
#pragma unroll
      for (int q = 0; q < V; ++q) acc[k][q] = 0.f;
#pragma unroll 4
    for (int o = o_lo; o < o_hi; ++o) {
      float wv[V];
      if (V == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(w + (int64_t)o * In + i));
        wv[0] = t.x; wv[1 % V] = t.y; wv[2 % V] = t.z; wv[3 % V] = t.w;
      } else {
        wv[0] = w[(int64_t)o * In + i];
      }
#pragma unroll
      for (int k = 0; k < LIN_BT; ++k)
        if (b0 + k < B) {
          const float gv = __ldg(g + (int64_t)(b0 + k) * Out + o);
#pragma unroll
          for (int q = 0; q < V; ++q) acc[k][q] = fmaf(wv[q], gv, acc[k][q]);
        }
    }
#pragma unroll
    for (int k = 0; k < LIN_BT; ++k)
      if (b0 + k < B) {
#pragma unroll
        for (int q = 0; q < V; ++q) atomicAdd(gx + (int64_t)(b0 + k) * In + i + q, scale * acc[k][q]);
      }
  }
}
// gw[o][i] = scale * sum_b g[b][o]*x[b][i];  gb[o] = sum_b g[b][o]
__global__ void k_linear_wgrad(const float* __restrict__ g, const float* __restrict__ x,
                               float* __restrict__ gw, float* __restrict__ gb, int B, int In,
                               int Out, float scale) {
  sg_pdl_enter();
  int64_t total = (int64_t)Out * In;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(e % In);
    int o = (int)(e / In);
    float acc = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
      float gv = __ldg(g + (int64_t)b * Out + o);
      acc += gv * __ldg(x + (int64_t)b * In + i);
      sb += gv;
    }
    gw[e] = scale * acc;
    if (gb && i == 0) gb[o] = sb;
  }
}
extern "C" int sg_linear_fwd(const float* x, const float* w, const float* bias, float* y, int B,
                             int In, int Out, float scale, int lrelu, cudaStream_t s) {
  if (B == 0 || Out == 0) return 0;
  sg_launch((k_linear_fwd), (unsigned)Out, 256, 0, s, x, w, bias, y, B, In, Out, scale, lrelu);
  return sg_check_launch("sg_linear_fwd");
}
extern "C" int sg_linear_dgrad(const float* g, const float* w, float* gx, int B, int In, int Out,
                               float scale, cudaStream_t s) {
  if (B == 0 || In == 0) return 0;
  cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)B * In, s);
  const int V = (In & 3) == 0 ? 4 : 1;
  int bx = (In / V + 127) / 128;
  int splits = (sg_num_sms() * 4 + bx - 1) / bx;
  if (splits > Out) splits = Out;
  if (splits < 1) splits = 1;
  int o_per = (Out + splits - 1) / splits;
  splits = (Out + o_per - 1) / o_per;
  if (V == 4)
    sg_launch((k_linear_dgrad<4>), dim3(bx, splits), 128, 0, s, g, w, gx, B, In, Out, scale, o_per);
  else
    sg_launch((k_linear_dgrad<1>), dim3(bx, splits), 128, 0, s, g, w, gx, B, In, Out, scale, o_per);
  return sg_check_launch("sg_linear_dgrad");
}
extern "C" int sg_linear_wgrad(const float* g, const float* x, float* gw, float* gb, int B, int In,
                               int Out, float scale, cudaStream_t s) {
  int64_t total = (int64_t)Out * In;
  if (total == 0) return 0;
  sg_launch((k_linear_wgrad), sg_grid(total, 256), 256, 0, s, g, x, gw, gb, B, In, Out, scale);
  return sg_check_launch("sg_linear_wgrad");
}

// ------------------------------------------------------ fused multi-tensor Adam (+ EMA)
// main.py:141-142 `torch.optim.Adam(params, lr, betas=(0, 0.99))` for every active parameter of a
// network in ONE launch (memory-bound: 16-20 B/parameter), optionally followed
// by the generator-weight EMA of the TF path (SURFGAN_3D/ExtendedEMA.py).  Same update as
// torch.optim.Adam (no amsgrad, no weight decay):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// `step` is a device counter (t-1 before the call) so the launch can live in a CUDA graph;
// sg_adam_advance increments it.
struct SgAdamTensor {
  float* p;
  const float* g;
  float* m;       // may be null when beta1 == 0
  float* v;
  float* ema;     // may be null
  int64_t n;
};
// lr_dev (nullable): learning rate read from device memory instead of the `lr` argument -- a captured CUDA graph
// then follows a LambdaLR schedule (main.py:145) without being re-captured.
__global__ void __launch_bounds__(256)
k_adam_multi(const SgAdamTensor* __restrict__ tensors, const int* __restrict__ block_tensor,
             const int64_t* __restrict__ block_offset, const int* __restrict__ step, float lr,
             const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float ema_beta) {
  sg_pdl_enter();
  const SgAdamTensor t = tensors[block_tensor[blockIdx.x]];
  const int64_t base = block_offset[blockIdx.x];
  const float tt = (float)(*step + 1);
  const float bc1 = beta1 > 0.f ? 1.f - powf(beta1, tt) : 1.f;
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, tt));
  if (lr_dev) lr = *lr_dev;
  const float step_size = lr / bc1;
  auto update = [&](float g, float& m, float& v, float& p, float& e) {
    m = beta1 > 0.f ? beta1 * m + (1.f - beta1) * g : g;
    v = beta2 * v + (1.f - beta2) * g * g;
    p = p - step_size * m / (sqrtf(v) / bc2_sqrt + eps);
    e = ema_beta * e + (1.f - ema_beta) * p;
  };
  // a block owns 1024 consecutive elements; with 16-byte aligned tensors a thread updates four consecutive ones
  // through float4 accesses (20 B/parameter at beta1 = 0: three 16-byte loads and two stores per thread)
  const bool vec = ((((uintptr_t)t.p) | ((uintptr_t)t.g) | ((uintptr_t)t.v) | ((uintptr_t)t.m) | ((uintptr_t)t.ema)) & 15) == 0;
  const int64_t i = base + 4 * (int64_t)threadIdx.x;
  if (vec && i + 3 < t.n) {
    const float4 g4 = *reinterpret_cast<const float4*>(t.g + i);
    float4 v4 = *reinterpret_cast<const float4*>(t.v + i);
    float4 p4 = *reinterpret_cast<const float4*>(t.p + i);
    float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), e4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beta1 > 0.f) m4 = *reinterpret_cast<const float4*>(t.m + i);
    if (t.ema) e4 = *reinterpret_cast<const float4*>(t.ema + i);
    update(g4.x, m4.x, v4.x, p4.x, e4.x);
    update(g4.y, m4.y, v4.y, p4.y, e4.y);
    update(g4.z, m4.z, v4.z, p4.z, e4.z);
    update(g4.w, m4.w, v4.w, p4.w, e4.w);
    if (beta1 > 0.f) *reinterpret_cast<float4*>(t.m + i) = m4;
    *reinterpret_cast<float4*>(t.v + i) = v4;
    *reinterpret_cast<float4*>(t.p + i) = p4;
    if (t.ema) *reinterpret_cast<float4*>(t.ema + i) = e4;
  } else {
    for (int k = 0; k < 4; ++k) {
      const int64_t e_i = i + k;
      if (e_i >= t.n) break;
      float m = beta1 > 0.f ? t.m[e_i] : 0.f, v = t.v[e_i], p = t.p[e_i], e = t.ema ? t.ema[e_i] : 0.f;
      update(t.g[e_i], m, v, p, e);
      if (beta1 > 0.f) t.m[e_i] = m;
      t.v[e_i] = v;
      t.p[e_i] = p;
      if (t.ema) t.ema[e_i] = e;
    }
  }
}
__global__ void k_adam_advance(int* step) {
  sg_pdl_enter(); *step += 1; }

extern "C" int sg_adam_step(const void* tensors, const int* block_tensor, const int64_t* block_offset, int n_blocks,
                            const int* step, float lr, const float* lr_dev, float beta1, float beta2, float eps,
                            float ema_beta, cudaStream_t s) {
  if (n_blocks == 0) return 0;
  sg_launch((k_adam_multi), (unsigned)n_blocks, 256, 0, s, (const SgAdamTensor*)tensors, block_tensor, block_offset, step, lr,
            lr_dev, beta1, beta2, eps, ema_beta);
  return sg_check_launch("sg_adam_step");
}
extern "C" int sg_adam_advance(int* step, cudaStream_t s) {
  sg_launch((k_adam_advance), 1, 1, 0, s, step);
  return sg_check_launch("sg_adam_advance");
}

// ------------------------------------------------- gradient arena packing (data-parallel exchange)
// dst[i] = scale * src[i] for every (src, dst, n) row of a device table in ONE launch: the gradients of a network's
// active parameters gathered into the contiguous arena that ncclAllReduce sums (scale = 1 / world) and that the fused
// Adam then reads directly -- replaces torch.cat + mul_ + _foreach_copy_ around the all-reduce (main.py:147-160's
// hvd.DistributedOptimizer).  block b copies elements [block_offset[b], +1024) of row block_row[b].
struct SgCopyRow {
  const float* src;
  float* dst;
  int64_t n;
};
__global__ void __launch_bounds__(256)
k_multi_copy_scale(const SgCopyRow* __restrict__ rows, const int* __restrict__ block_row,
                   const int64_t* __restrict__ block_offset, float scale) {
  sg_pdl_enter();
  const SgCopyRow r = rows[block_row[blockIdx.x]];
  const int64_t i = block_offset[blockIdx.x] + 4 * (int64_t)threadIdx.x;
  if (((((uintptr_t)r.src) | ((uintptr_t)r.dst)) & 15) == 0 && i + 3 < r.n) {
    float4 v = *reinterpret_cast<const float4*>(r.src + i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    *reinterpret_cast<float4*>(r.dst + i) = v;
  } else {
    for (int k = 0; k < 4 && i + k < r.n; ++k) r.dst[i + k] = scale * r.src[i + k];
  }
}
extern "C" int sg_multi_copy_scale(const void* rows, const int* block_row, const int64_t* block_offset, int n_blocks,
                                   float scale, cudaStream_t s) {
  if (n_blocks == 0) return 0;
  sg_launch((k_multi_copy_scale), (unsigned)n_blocks, 256, 0, s, (const SgCopyRow*)rows, block_row, block_offset, scale);
  return sg_check_launch("sg_multi_copy_scale");
}

// --------------------------------------------------------------- input preparation
// main.py:85-87 loader (`np.load -> float32 -> [None] / 1024`) + train.py:144 instance noise in one
// pass over the raw uint16 voxels: out = raw * scale + sigma * noise.
__global__ void k_prepare_real(const uint16_t* __restrict__ raw, const float* __restrict__ noise,
                               float* __restrict__ out, int64_t n, float scale, float sigma) {
  sg_pdl_enter();
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; i < n; i += (int64_t)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const ushort4 r = *reinterpret_cast<const ushort4*>(raw + i);
      float4 z = noise ? *reinterpret_cast<const float4*>(noise + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(out + i) = make_float4(r.x * scale + sigma * z.x, r.y * scale + sigma * z.y,
                                                        r.z * scale + sigma * z.z, r.w * scale + sigma * z.w);
    } else {
      for (int64_t j = i; j < n; ++j) out[j] = raw[j] * scale + (noise ? sigma * noise[j] : 0.f);
    }
  }
}
extern "C" int sg_prepare_real(const void* raw_u16, const float* noise, float* out, int64_t n, float scale,
                               float sigma, cudaStream_t s) {
  if (n == 0) return 0;
  sg_launch((k_prepare_real), sg_grid((n + 3) / 4, 256), 256, 0, s, (const uint16_t*)raw_u16, noise, out, n, scale, sigma);
  return sg_check_launch("sg_prepare_real");
}

// ------------------------------------------------------------- minibatch standard deviation
// network.py:113-133 on the plain fp32 base-level tensor x[B][C][V] viewed as [G][M][F], B = G*M,
// F = C*V (n = g*M + m):  xc = x - mean_g x;  s = sqrt(mean_g xc^2 + eps);  t[m] = mean_f s;
// out[B][C+1][V] = cat(xc, t[n % M] broadcast).  S independent minibatches of G*M samples each may share
// a launch (sample = sb*G*M + g*M + m; s, t, gt are [S*M]...) so D(real) and D(fake) can run as one
// batch-2B forward without mixing their statistics.  It is the only non-piecewise-linear op of D, so the
// gradient penalty needs its SECOND derivative: the backward is its own kernel and has a backward.
//   bwd   : gx = gxc - mean_g gxc + gt[m] * xc / (F*G*s),   gt[m] = sum of the stat channel's gradient
//   bwdbwd: given u = d/d(gx):  d/d(gxc) = u - mean_g u;  d/d(gt[m]) = sum_{g,f} u*xc/(F*G*s);
//           d/dx = C( gt[m]/(F*G) * ( u/s - xc * (sum_g u*xc) / (G*s^3) ) ),  C = group centring
// One thread per (m, f); G <= 8 values per thread in registers; block + atomic reductions over f.
#define MB_MAXG 8
__global__ void k_mbstd_fwd(const float* __restrict__ x, float* __restrict__ out, float* __restrict__ s_out,
                            float* __restrict__ t, int G, int M, int C, int V, float eps) {
  sg_pdl_enter();
  const int F = C * V;
  const int mm = blockIdx.y;              // (sub-batch, m) pair: independent minibatches share a launch
  const int sb = mm / M, m = mm % M;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) {
    const int c = f / V, v = f % V;
    float xv[MB_MAXG], mean = 0.f;
#pragma unroll
    for (int g = 0; g < MB_MAXG; ++g)
      if (g < G) { xv[g] = x[((int64_t)(sb * G * M + g * M + m) * C + c) * V + v]; mean += xv[g]; }
    mean /= (float)G;
    float var = 0.f;
#pragma unroll
    for (int g = 0; g < MB_MAXG; ++g)
      if (g < G) { xv[g] -= mean; var += xv[g] * xv[g]; out[((int64_t)(sb * G * M + g * M + m) * (C + 1) + c) * V + v] = xv[g]; }
    s_out[(int64_t)mm * F + f] = sqrtf(var / (float)G + eps);
  }
}
// t[mm] = mean_f s[mm][f]: one block per (sub-batch, m), fixed summation order (the stat channel feeds the
// last conv of D: with atomics its last bits -- and now and then a LeakyReLU mask behind it -- changed from
// run to run)
__global__ void __launch_bounds__(256) k_mbstd_tmean(const float* __restrict__ s, float* __restrict__ t, int F) {
  sg_pdl_enter();
  const int mm = blockIdx.x;
  float acc = 0.f;
  for (int f = threadIdx.x; f < F; f += 256) acc += s[(int64_t)mm * F + f];
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += sm[w];
    t[mm] = tot / (float)F;
  }
}
// writes the stat channel out[n][C][v] = t[n % M]
__global__ void k_mbstd_stat(float* __restrict__ out, const float* __restrict__ t, int B, int G, int M, int C,
                             int V) {
  sg_pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * V) return;
  const int n = i / V, v = i % V;
  const int sb = n / (G * M), m = (n % (G * M)) % M;
  out[((int64_t)n * (C + 1) + C) * V + v] = t[sb * M + m];
}
// gt[m] = sum over g, v of gout[(g*M+m)][C][v]
__global__ void k_mbstd_gt(const float* __restrict__ gout, float* __restrict__ gt, int G, int M, int C, int V) {
  sg_pdl_enter();
  const int mm = blockIdx.x;
  const int sb = mm / M, m = mm % M;
  float acc = 0.f;
  for (int i = threadIdx.x; i < G * V; i += blockDim.x) {
    const int g = i / V, v = i % V;
    acc += gout[((int64_t)(sb * G * M + g * M + m) * (C + 1) + C) * V + v];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(gt + mm, acc);
}
__global__ void k_mbstd_bwd(const float* __restrict__ gout, const float* __restrict__ gt, const float* __restrict__ out,
                            const float* __restrict__ s, float* __restrict__ gx, int G, int M, int C, int V) {
  sg_pdl_enter();
  const int F = C * V;
  const int mm = blockIdx.y;              // (sub-batch, m) pair: independent minibatches share a launch
  const int sb = mm / M, m = mm % M;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int c = f / V, v = f % V;
  float gv[MB_MAXG], mean = 0.f;
#pragma unroll
  for (int g = 0; g < MB_MAXG; ++g)
    if (g < G) { gv[g] = gout[((int64_t)(sb * G * M + g * M + m) * (C + 1) + c) * V + v]; mean += gv[g]; }
  mean /= (float)G;
  const float k = gt[mm] / ((float)F * (float)G * s[(int64_t)mm * F + f]);
#pragma unroll
  for (int g = 0; g < MB_MAXG; ++g)
    if (g < G) {
      const float xc = out[((int64_t)(sb * G * M + g * M + m) * (C + 1) + c) * V + v];
      gx[((int64_t)(sb * G * M + g * M + m) * C + c) * V + v] = gv[g] - mean + k * xc;
    }
}
// u = d/d(gx) [B][C][V];  outputs: d_gout [B][C+1][V] (feature part here, stat part by k_mbstd_stat
// from d_gt), d_gt[m] (atomics, zeroed by the caller), d_x [B][C][V]
__global__ void k_mbstd_bwdbwd(const float* __restrict__ u, const float* __restrict__ gt, const float* __restrict__ out,
                               const float* __restrict__ s, float* __restrict__ d_gout, float* __restrict__ d_gt,
                               float* __restrict__ d_x, int G, int M, int C, int V) {
  sg_pdl_enter();
  const int F = C * V;
  const int mm = blockIdx.y;              // (sub-batch, m) pair: independent minibatches share a launch
  const int sb = mm / M, m = mm % M;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float part = 0.f;
  if (f < F) {
    const int c = f / V, v = f % V;
    const float sv = s[(int64_t)mm * F + f];
    float uv[MB_MAXG], xc[MB_MAXG], umean = 0.f, uxc = 0.f;
#pragma unroll
    for (int g = 0; g < MB_MAXG; ++g)
      if (g < G) {
        uv[g] = u[((int64_t)(sb * G * M + g * M + m) * C + c) * V + v];
        xc[g] = out[((int64_t)(sb * G * M + g * M + m) * (C + 1) + c) * V + v];
        umean += uv[g];
        uxc += uv[g] * xc[g];
      }
    umean /= (float)G;
    const float fg = (float)F * (float)G;
    part = uxc / (fg * sv);                         // contribution to d_gt[m]
    const float a = gt[mm] / fg;
    float dxc[MB_MAXG], dmean = 0.f;
#pragma unroll
    for (int g = 0; g < MB_MAXG; ++g)
      if (g < G) {
        dxc[g] = a * (uv[g] / sv - xc[g] * uxc / ((float)G * sv * sv * sv));
        dmean += dxc[g];
      }
    dmean /= (float)G;
#pragma unroll
    for (int g = 0; g < MB_MAXG; ++g)
      if (g < G) {
        d_gout[((int64_t)(sb * G * M + g * M + m) * (C + 1) + c) * V + v] = uv[g] - umean;
        d_x[((int64_t)(sb * G * M + g * M + m) * C + c) * V + v] = dxc[g] - dmean;
      }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(d_gt + mm, part);
}

extern "C" int sg_mbstd_fwd(const float* x, float* out, float* s, float* t, int S, int G, int M, int C, int V,
                            float eps, cudaStream_t st) {
  SG_REQUIRE(G >= 1 && G <= MB_MAXG, "sg_mbstd_fwd: group size %d not in [1, %d]", G, MB_MAXG);
  const int F = C * V;
  sg_launch((k_mbstd_fwd), dim3((F + 255) / 256, S * M), 256, 0, st, x, out, s, t, G, M, C, V, eps);
  int rc = sg_check_launch("sg_mbstd_fwd");
  if (rc) return rc;
  sg_launch((k_mbstd_tmean), S * M, 256, 0, st, (const float*)s, t, F);
  rc = sg_check_launch("sg_mbstd_fwd(mean)");
  if (rc) return rc;
  sg_launch((k_mbstd_stat), (S * G * M * V + 255) / 256, 256, 0, st, out, t, S * G * M, G, M, C, V);
  return sg_check_launch("sg_mbstd_fwd(stat)");
}
extern "C" int sg_mbstd_bwd(const float* gout, const float* out, const float* s, float* gt, float* gx, int S, int G,
                            int M, int C, int V, cudaStream_t st) {
  SG_REQUIRE(G >= 1 && G <= MB_MAXG, "sg_mbstd_bwd: group size %d not in [1, %d]", G, MB_MAXG);
  const int F = C * V;
  cudaMemsetAsync(gt, 0, sizeof(float) * S * M, st);
  sg_launch((k_mbstd_gt), S * M, 128, 0, st, gout, gt, G, M, C, V);
  int rc = sg_check_launch("sg_mbstd_bwd(gt)");
  if (rc) return rc;
  sg_launch((k_mbstd_bwd), dim3((F + 255) / 256, S * M), 256, 0, st, gout, gt, out, s, gx, G, M, C, V);
  return sg_check_launch("sg_mbstd_bwd");
}
extern "C" int sg_mbstd_bwdbwd(const float* u, const float* gt, const float* out, const float* s, float* d_gout,
                               float* d_gt, float* d_x, int S, int G, int M, int C, int V, cudaStream_t st) {
  SG_REQUIRE(G >= 1 && G <= MB_MAXG, "sg_mbstd_bwdbwd: group size %d not in [1, %d]", G, MB_MAXG);
  const int F = C * V;
  cudaMemsetAsync(d_gt, 0, sizeof(float) * S * M, st);
  sg_launch((k_mbstd_bwdbwd), dim3((F + 255) / 256, S * M), 256, 0, st, u, gt, out, s, d_gout, d_gt, d_x, G, M, C, V);
  int rc = sg_check_launch("sg_mbstd_bwdbwd");
  if (rc) return rc;
  sg_launch((k_mbstd_stat), (S * G * M * V + 255) / 256, 256, 0, st, d_gout, d_gt, S * G * M, G, M, C, V);
  return sg_check_launch("sg_mbstd_bwdbwd(stat)");
}
