// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core convolution kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

#ifndef SG_TC_WATCHDOG
#define SG_TC_WATCHDOG 1
#endif

namespace {

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// A protocol bug must fault within ~2 s instead of hanging the GPU.  Measured on the cfg3 step (round 2, same
// commit): this inline form 16.3 ms/step = the build without the watchdog (16.3 ms); moving the time-out into a
// __noinline__ slow path cost 4 % (17.0 ms: the call forces the issuing warp's descriptors out of uniform registers).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if SG_TC_WATCHDOG
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("conv_tc: mbarrier timeout (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// The MMA warp runs its loops with all 32 lanes (warp-uniform control flow and operands, so ptxas
// keeps descriptors in uniform registers) and predicates only the tcgen05 instructions on the
// elected lane.  Issuing from inside an `if (lane == 0)` region instead costs ~10 extra SASS
// instructions per MMA (R2UR moves and an ELECT/BRA.U.ANY loop around every UTCHMMA).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tc_commit(uint32_t bar, uint32_t leader = 1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
      "r"(leader)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, M = 128, N from idesc, K = 16
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate, uint32_t leader = 1) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
// The same for fp32 operands read as TF32 (kind::tf32, K = 8 per instruction: with 16-byte core-matrix rows of
// FOUR channels the shared-memory operand layout is byte-for-byte the bf16 one -- DESIGN.md "TF32").
template <bool TF32>
__device__ __forceinline__ void tc_mma_t(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate, uint32_t leader = 1) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
  } else {
    tc_mma(d_tmem, adesc, bdesc, idesc, accumulate, leader);
  }
}
// instruction descriptor without the N / M / major-ness fields: D = f32, A = B = bf16 (1) or tf32 (2)
__host__ __device__ constexpr uint32_t idesc_formats(bool tf32) {
  return (1u << 4) | ((tf32 ? 2u : 1u) << 7) | ((tf32 ? 2u : 1u) << 10);
}
// round-to-nearest fp32 -> tf32 (the tensor core itself TRUNCATES the 13 low mantissa bits: a systematic -2^-11
// relative bias per operand that would accumulate over the layers)
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// In-place tf32 rounding of `bytes` (multiple of 16) of a TMA-written shared-memory tile by the `nthreads` calling
// threads (thread index `tid`), followed by the proxy fence that makes the generic-proxy writes visible to the
// tensor core's async-proxy reads.  The caller then arrives on the barrier the MMA issuer waits on.
__device__ __forceinline__ void round_tile_tf32(uint8_t* base, int bytes, int tid, int nthreads) {
  float4* p = reinterpret_cast<float4*>(base);
  for (int i = tid; i < bytes / 16; i += nthreads) {
    float4 v = p[i];
    v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
    p[i] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// K-major, no swizzle: 8 rows x 16 B core matrices; LBO = byte distance between the two core
// matrices along K, SBO = byte distance between 8-row groups along M/N (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version 1 (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// CB (16, 32 or 64) consecutive accumulator columns of this thread's row: all loads issued back to back, ONE
// wait (a wait per 16 columns cost the epilogue ~300 cycles per block).  The empty asm statements after the
// wait pin every consumer of the registers behind it (volatile asms keep their order).
template <int CB>
__device__ __forceinline__ void tmem_ld_block(uint32_t taddr, float (&v)[CB]) {
  uint32_t r[CB];
#pragma unroll
  for (int j = 0; j < CB / 16; ++j) {
    uint32_t* q = r + 16 * j;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(taddr + 16u * j));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < CB; ++i) {
    asm volatile("" : "+r"(r[i]));
    v[i] = __uint_as_float(r[i]);
  }
}

// three 16-column blocks of this thread's row (three accumulators of the z-stacked kernel), one wait
__device__ __forceinline__ void tmem_ld3x16(uint32_t t0, uint32_t t1, uint32_t t2, float (&v)[48]) {
  uint32_t r[48];
  const uint32_t ta[3] = {t0, t1, t2};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    uint32_t* q = r + 16 * j;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
        "%15}, [%16];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
        : "r"(ta[j]));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 48; ++i) {
    asm volatile("" : "+r"(r[i]));
    v[i] = __uint_as_float(r[i]);
  }
}

constexpr int kMaxSub = 8;   // MMA tiles (128 rows each) per CTA tile

__device__ __forceinline__ F8 unpack8(const uint4& u) {
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

// Epilogue of 16 accumulator columns (= two 8-channel chunks) of one output voxel:
// y = [mask] [lrelu] (acc*scale + bias) -> two 16-byte stores.  `sb` points at the 16 biases in
// SHARED memory (broadcast float4 reads; the first version fetched them with per-element __ldg
// and the four epilogue warps became the kernel's bottleneck).  Pad output channels need no
// special case: their packed weight rows and biases are zero, so they come out as exact zeros.
template <typename T>
__device__ __forceinline__ void epilogue16(const float* v, const float* sb, float scale, int lrelu,
                                           const T* mask, T* y, int64_t chunk_stride) {
  float r[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    r[4 * q + 0] = fmaf(v[4 * q + 0], scale, b.x);
    r[4 * q + 1] = fmaf(v[4 * q + 1], scale, b.y);
    r[4 * q + 2] = fmaf(v[4 * q + 2], scale, b.z);
    r[4 * q + 3] = fmaf(v[4 * q + 3], scale, b.w);
  }
  if (lrelu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = lrelu02(r[j]);
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = r[half * 8 + j];
    if (mask) {
      const F8 m = ld8(mask + half * chunk_stride);
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] *= lmask02(m.v[j]);
    }
    st8(y + half * chunk_stride, o);
  }
}

// One tap of one K block: (kpairs <= 2) x (n_sub <= kMaxSub) MMAs.  The issuing thread is a
// single lane, so the per-MMA instruction count is what bounds the tensor pipe for small N
// (an MMA with N = 64 lasts ~32-48 cycles): descriptors are formed with one 64-bit add each
// from bases and offsets hoisted out of the loops (all offsets in 16-byte units, i.e. added to
// the descriptors' start-address field).
template <int NT, bool TF32 = false>
__device__ __forceinline__ void issue_tap(uint32_t tmem_acc, uint64_t a_desc, uint64_t b_desc,
                                          const uint32_t (&sub_off)[kMaxSub], int n_sub, int kpairs, uint32_t kk_a,
                                          uint32_t idesc, uint32_t acc_first, uint32_t leader) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    if (kk < kpairs) {
      const uint64_t b = b_desc + (uint64_t)(kk * 2 * NT);
      const uint64_t a_k = a_desc + (uint64_t)(kk * kk_a);
      const uint32_t acc = kk == 0 ? acc_first : 1u;
#pragma unroll
      for (int sub = 0; sub < kMaxSub; ++sub)
        if (sub < n_sub) tc_mma_t<TF32>(tmem_acc + sub * NT, a_k + sub_off[sub], b, idesc, acc, leader);
    }
  }
}

// fp32 vector reduction into global memory (sm_90+): one 16-byte L2 transaction for four adds
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Same as epilogue16 with the two mask vectors already in registers (prefetched while the MMAs of
// the tile were still running, so their global-load latency is off the epilogue's critical path).
__device__ __forceinline__ void epilogue16_regmask(const float* v, const float* sb, float scale, int lrelu,
                                                   bool has_mask, const uint4& m0, const uint4& m1, __nv_bfloat16* y,
                                                   int64_t chunk_stride) {
  float r[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    r[4 * q + 0] = fmaf(v[4 * q + 0], scale, b.x);
    r[4 * q + 1] = fmaf(v[4 * q + 1], scale, b.y);
    r[4 * q + 2] = fmaf(v[4 * q + 2], scale, b.z);
    r[4 * q + 3] = fmaf(v[4 * q + 3], scale, b.w);
  }
  if (lrelu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = lrelu02(r[j]);
  }
  if (has_mask) {
    const F8 a = unpack8(m0), b = unpack8(m1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r[j] *= lmask02(a.v[j]);
      r[8 + j] *= lmask02(b.v[j]);
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = r[half * 8 + j];
    st8(y + half * chunk_stride, o);
  }
}

// MN-major, no swizzle (operand element (mn, k) at (mn/8)*SBO + (k/8)*LBO + (k%8)*16 + (mn%8)*2):
// same field encoding as make_desc; the major-ness is selected in the instruction descriptor.

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace
