// 3x3x3 / stride 1 / pad 1 conv3d on the channel-blocked layout: weight packing, the
// CUDA-core ("direct") fprop / wgrad kernels and the public conv entry points.
//
// The direct kernels serve (a) the fp32 precision mode, (b) shapes the tcgen05 kernels in
// conv_tc.cu do not cover, (c) the on-device cross-check of the tcgen05 kernels in tests/.
// dgrad is fprop with the flipped/transposed packing: no third kernel.
#include "../../include/saragan_b200.h"
#include "common.cuh"
SG_DEFINE_LEAK_SETTER(sg_set_leak_conv_direct)

// tcgen05 paths (conv_tc.cu); return 1 when the shape is not covered so the caller can
// fall through to the direct kernel, 0 on success, <0 / cudaError on failure.
int sg_tc_fprop(const void* x, const void* wp, const float* bias, const void* mask_src, void* y,
                int N, int Cin, int Cout, int D, int H, int W, float scale, int lrelu,
                void* workspace, int64_t workspace_bytes, cudaStream_t s, int tf32, void* pn_y = nullptr,
                float pn_eps = 0.f, int pn_lrelu_after = 0, void* pool_y = nullptr, float pool_scale = 0.f);
int sg_tc_wgrad(const void* x, const void* gy, float* gw, float* gb, int N, int Cin, int Cout,
                int D, int H, int W, float scale, void* workspace, int64_t workspace_bytes,
                cudaStream_t s, int tf32);
int64_t sg_tc_workspace_bytes(int kind, int N, int Cin, int Cout, int D, int H, int W, int tf32);
int sg_wgrad_finish(const float* ws, float* gw, int Cout, int Cin, int CinP, float scale, cudaStream_t s);

// bf16 convolutions that the tcgen05 planners declined and the CUDA-core kernels ran instead
// (bench.py reports the count: a non-zero value on a large layer is a performance bug)
static unsigned long long g_cuda_core_fallbacks = 0;
extern "C" int64_t sg_cuda_core_fallbacks(int reset) {
  unsigned long long v = g_cuda_core_fallbacks;
  if (reset) g_cuda_core_fallbacks = 0;
  return (int64_t)v;
}

// ---------------------------------------------------------------------- weight packing
// src: fp32 [Cout][Cin][27] (torch (Cout,Cin,3,3,3) contiguous).
// fwd packing  (transpose_flip = 0): dst[tap][CCin ][CoutP][8]: elem(tap, ci, co) = w[co][ci][tap]
// bwd packing  (transpose_flip = 1): dst[tap][CCout][CinP ][8]: elem(tap, co, ci) = w[co][ci][26-tap]
// i.e. in both cases dst[tap][kchunk][row][k%8] with (k = contraction channel, row = output
// channel) of the convolution the packing will be used for.  Pad entries are zero.
// One block packs a tile of 32 output rows x one 8-channel K chunk x 27 taps through shared memory, so that both
// the fp32 reads (runs of 216 / 864 consecutive floats of the parameter) and the packed writes (runs of 256
// consecutive elements per tap) are coalesced.  (The first version gathered one 4-byte element per thread, 108
// bytes apart: 8x read amplification, 0.5 ms per step for the 44 packings of the cfg3 networks.)
template <typename T, bool TF32 = false>
__device__ __forceinline__ void pack_tile(const float* __restrict__ w, T* __restrict__ dst, int Cout, int Cin,
                                          int transpose_flip, int kc, int r0, float* s /* [32][8][27] */) {
  const int K = transpose_flip ? Cout : Cin;     // contraction channels
  const int R = transpose_flip ? Cin : Cout;     // output rows
  const int KC = 2 * ((K + 15) / 16);
  const int RP = 16 * ((R + 15) / 16);
  if (!transpose_flip) {
    // row = co, k = ci: for a fixed co the (ci chunk, tap) block is 216 consecutive floats
    for (int idx = threadIdx.x; idx < 32 * 216; idx += blockDim.x) {
      const int r = idx / 216, rem = idx - r * 216;
      const int j = rem / 27;
      const int co = r0 + r, ci = kc * 8 + j;
      s[idx] = (co < Cout && ci < Cin) ? w[((int64_t)co * Cin + kc * 8) * 27 + rem] : 0.f;
    }
  } else {
    // row = ci, k = co, flipped taps: for a fixed co the (ci tile, tap) block is 32*27 consecutive floats
    for (int idx = threadIdx.x; idx < 8 * 864; idx += blockDim.x) {
      const int j = idx / 864, rem = idx - j * 864;
      const int r = rem / 27, tp = rem - r * 27;
      const int co = kc * 8 + j, ci = r0 + r;
      s[(r * 8 + j) * 27 + (26 - tp)] = (co < Cout && ci < Cin) ? w[((int64_t)co * Cin + r0) * 27 + rem] : 0.f;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 27 * 256; idx += blockDim.x) {
    const int tap = idx >> 8, e = idx & 255;
    if (TF32) {
      // SG_TF32 packing [tap][K/4][RP][4] fp32, rounded to tf32: the 8-channel chunk is two 16-byte K chunks
      const int h = e >> 7, r = (e >> 2) & 31, j4 = e & 3;
      if (r0 + r < RP) {
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(s[(r * 8 + h * 4 + j4) * 27 + tap]));
        st1(dst + (((int64_t)tap * 2 * KC + 2 * kc + h) * RP + r0 + r) * 4 + j4, __uint_as_float(t));
      }
    } else {
      const int r = e >> 3, j = e & 7;
      if (r0 + r < RP) st1(dst + (((int64_t)tap * KC + kc) * RP + r0 + r) * 8 + j, s[(r * 8 + j) * 27 + tap]);
    }
  }
}
template <typename T, bool TF32>
__global__ void __launch_bounds__(256)
k_pack_conv_weight(const float* __restrict__ w, T* __restrict__ dst, int Cout, int Cin, int transpose_flip) {
  sg_pdl_enter();
  __shared__ float s[32 * 8 * 27];
  pack_tile<T, TF32>(w, dst, Cout, Cin, transpose_flip, blockIdx.y, blockIdx.x * 32, s);
}
// every stale packing of a network in ONE launch: table row = one (weight, packing), block -> (row, kc, r tile)
struct SgPackJob {
  const float* w;
  void* dst;
  int Cout, Cin, flip, dtype;
};
__global__ void __launch_bounds__(256)
k_pack_conv_weights_multi(const SgPackJob* __restrict__ jobs, const int* __restrict__ block_job,
                          const int* __restrict__ block_kc, const int* __restrict__ block_r0) {
  sg_pdl_enter();
  __shared__ float s[32 * 8 * 27];
  const SgPackJob jb = jobs[block_job[blockIdx.x]];
  if (jb.dtype == SG_DTYPE_BF16)
    pack_tile<__nv_bfloat16>(jb.w, (__nv_bfloat16*)jb.dst, jb.Cout, jb.Cin, jb.flip, block_kc[blockIdx.x], block_r0[blockIdx.x], s);
  else if (jb.dtype == SG_DTYPE_TF32)
    pack_tile<float, true>(jb.w, (float*)jb.dst, jb.Cout, jb.Cin, jb.flip, block_kc[blockIdx.x], block_r0[blockIdx.x], s);
  else
    pack_tile<float>(jb.w, (float*)jb.dst, jb.Cout, jb.Cin, jb.flip, block_kc[blockIdx.x], block_r0[blockIdx.x], s);
}
extern "C" int64_t sg_packed_weight_elems(int Cout, int Cin, int transpose_flip) {
  int K = transpose_flip ? Cout : Cin;
  int R = transpose_flip ? Cin : Cout;
  return (int64_t)27 * (2 * ((K + 15) / 16)) * (16 * ((R + 15) / 16)) * 8;
}
extern "C" int sg_pack_conv_weight(const float* w, void* dst, int dtype, int Cout, int Cin,
                                   int transpose_flip, cudaStream_t s) {
  const int K = transpose_flip ? Cout : Cin, R = transpose_flip ? Cin : Cout;
  dim3 grid((unsigned)((16 * ((R + 15) / 16) + 31) / 32), (unsigned)(2 * ((K + 15) / 16)));
  if (dtype == SG_DTYPE_TF32) {
    sg_launch((k_pack_conv_weight<float, true>), grid, 256, 0, s, w, (float*)dst, Cout, Cin, transpose_flip);
  } else {
    SG_DISPATCH(dtype, sg_launch((k_pack_conv_weight<T, false>), grid, 256, 0, s, w, (T*)dst, Cout, Cin, transpose_flip););
  }
  return sg_check_launch("sg_pack_conv_weight");
}
extern "C" int sg_pack_conv_weights_multi(const void* jobs, const int* block_job, const int* block_kc,
                                          const int* block_r0, int n_blocks, cudaStream_t s) {
  if (n_blocks == 0) return 0;
  sg_launch((k_pack_conv_weights_multi), (unsigned)n_blocks, 256, 0, s, (const SgPackJob*)jobs, block_job, block_kc, block_r0);
  return sg_check_launch("sg_pack_conv_weights_multi");
}

// ------------------------------------------------------------------------ direct fprop
// thread = (voxel, output chunk of 8).  Per (tap, input chunk): one 8-wide x load and eight
// 8-wide (warp-uniform, L1-broadcast) weight loads feed 64 FMAs.
template <typename T>
__global__ void __launch_bounds__(128)
k_conv_direct(const T* __restrict__ x, const T* __restrict__ wp, const float* __restrict__ bias,
              const T* __restrict__ mask_src, T* __restrict__ y, int Cout, int CCin, int CCout,
              int CoutP, int D, int H, int W, float scale, int lrelu) {
  sg_pdl_enter();
  int64_t V = (int64_t)D * H * W;
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int cco = blockIdx.y;
  int n = blockIdx.z;
  if (v >= V) return;
  int w0 = (int)(v % W);
  int h0 = (int)((v / W) % H);
  int d0 = (int)(v / ((int64_t)W * H));
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  const T* xn = x + (int64_t)n * CCin * V * 8;
  for (int kd = 0; kd < 3; ++kd) {
    int d = d0 + kd - 1;
    if (d < 0 || d >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      int h = h0 + kh - 1;
      if (h < 0 || h >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        int w_ = w0 + kw - 1;
        if (w_ < 0 || w_ >= W) continue;
        int tap = (kd * 3 + kh) * 3 + kw;
        int64_t vin = ((int64_t)d * H + h) * W + w_;
        const T* wt = wp + (((int64_t)tap * CCin) * CoutP + cco * 8) * 8;
        for (int cci = 0; cci < CCin; ++cci) {
          F8 xv = ld8(xn + ((int64_t)cci * V + vin) * 8);
          const T* wr = wt + (int64_t)cci * CoutP * 8;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            F8 wv = ld8(wr + r * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[r] = fmaf(xv.v[j], wv.v[j], acc[r]);
          }
        }
      }
    }
  }
  int64_t oidx = (((int64_t)n * CCout + cco) * V + v) * 8;
  F8 o;
  F8 m;
  if (mask_src) m = ld8(mask_src + oidx);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int co = cco * 8 + r;
    float t = 0.f;
    if (co < Cout) {
      t = acc[r] * scale + (bias ? __ldg(bias + co) : 0.f);
      if (lrelu) t = lrelu02(t);
      if (mask_src) t *= lmask02(m.v[r]);
    }
    o.v[r] = t;
  }
  st8(y + oidx, o);
}


// ------------------------------------------------- small-volume fp32 fprop (base level)
// The 1x4x4 base level of both networks runs in fp32 (config.py): tiny M = N*V (64 rows at
// B=4) against K = 27*Cin up to 13851 and Cout = 512 -- a skinny GEMM bound by streaming the
// fp32 weights.  Implicit-im2col SGEMM: block = 64 rows x 64 output channels x one LIVE tap x one
// slice of the input-channel chunks (taps that only ever see padding -- kd != 1 when D == 1 -- are
// skipped; split-K over taps and chunk slices gives ~600 blocks so the weight stream has enough loads
// in flight), next k-tile prefetched into registers during the FMAs, 4x4 register tile per thread,
// partial sums into a [slices][M][CoutP] workspace, then a finishing kernel adds the slices, applies scale / bias /
// LeakyReLU / mask and writes the blocked layout.  (Partial sums per (tap, slice) in a workspace, added in a
// fixed order by the finishing kernel: deterministic.)
struct SmallTaps {
  int kd_lo, nkd, kh_lo, nkh, kw_lo, nkw;   // live tap ranges
  int ksplit, cc_per;                       // chunk slices per tap, chunks per slice (even)
};
static SmallTaps small_taps(int CCin, int D, int H, int W, int64_t blocks_per_tap_slice) {
  SmallTaps t;
  t.kd_lo = D == 1 ? 1 : 0; t.nkd = D == 1 ? 1 : 3;
  t.kh_lo = H == 1 ? 1 : 0; t.nkh = H == 1 ? 1 : 3;
  t.kw_lo = W == 1 ? 1 : 0; t.nkw = W == 1 ? 1 : 3;
  int64_t blocks = blocks_per_tap_slice * t.nkd * t.nkh * t.nkw;
  int ks = (int)((4 * (int64_t)sg_num_sms() + blocks - 1) / blocks);
  int max_ks = CCin / 8 > 1 ? CCin / 8 : 1;      // at least four k-tiles (of two chunks) per block
  if (ks > max_ks) ks = max_ks;
  if (ks < 1) ks = 1;
  t.cc_per = 2 * ((CCin / 2 + ks - 1) / ks);
  t.ksplit = (CCin + t.cc_per - 1) / t.cc_per;
  return t;
}

__global__ void __launch_bounds__(256)
k_conv_small_f32(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ acc,
                 int N, int CCin, int CoutP, int D, int H, int W, SmallTaps st) {
  sg_pdl_enter();
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  int z = blockIdx.z;
  const int ks = z % st.ksplit; z /= st.ksplit;
  const int kw = st.kw_lo + z % st.nkw; z /= st.nkw;
  const int kh = st.kh_lo + z % st.nkh; z /= st.nkh;
  const int kd = st.kd_lo + z;
  const int tap = (kd * 3 + kh) * 3 + kw;
  int V = D * H * W;
  int M = N * V;
  int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  int t = threadIdx.x;
  int tx = t & 15, ty = t >> 4;
  // loader mapping: row/col = t/4, quarter = t%4 (4 consecutive k of the 16-wide k tile)
  int lrow = t >> 2, lq = t & 3;
  int m = m0 + lrow;
  int64_t a_off = -1;   // element offset of (n, chunk 0, vin, 0) or -1 if the tap falls outside
  if (m < M) {
    int n = m / V, v = m % V;
    int w0 = v % W, h0 = (v / W) % H, d0 = v / (W * H);
    int d = d0 + kd - 1, h = h0 + kh - 1, w_ = w0 + kw - 1;
    if (d >= 0 && d < D && h >= 0 && h < H && w_ >= 0 && w_ < W)
      a_off = ((int64_t)n * CCin * V + ((int64_t)d * H + h) * W + w_) * 8;
  }
  int co = n0 + lrow;
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  const int cc_begin = ks * st.cc_per;
  const int cc_end = cc_begin + st.cc_per < CCin ? cc_begin + st.cc_per : CCin;
  const int sub = (lq & 1) * 4;
  auto load = [&](int cc0, float4& av, float4& bv) {
    const int chunk = cc0 + (lq >> 1);
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (chunk < cc_end) {
      if (a_off >= 0) av = *reinterpret_cast<const float4*>(x + a_off + (int64_t)chunk * V * 8 + sub);
      if (co < CoutP)
        bv = __ldcs(reinterpret_cast<const float4*>(wp + (((int64_t)tap * CCin + chunk) * CoutP + co) * 8 + sub));
    }
  };
  float4 av, bv;
  load(cc_begin, av, bv);
  for (int cc0 = cc_begin; cc0 < cc_end; cc0 += 2) {
    int k0 = lq * 4;
    As[k0 + 0][lrow] = av.x; As[k0 + 1][lrow] = av.y; As[k0 + 2][lrow] = av.z; As[k0 + 3][lrow] = av.w;
    Bs[k0 + 0][lrow] = bv.x; Bs[k0 + 1][lrow] = bv.y; Bs[k0 + 2][lrow] = bv.z; Bs[k0 + 3][lrow] = bv.w;
    __syncthreads();
    if (cc0 + 2 < cc_end) load(cc0 + 2, av, bv);   // in flight during the FMAs below
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fmaf(a[i], b[j], c[i][j]);
    }
    __syncthreads();
  }
  // this block's partial sums go to its own slice [blockIdx.z][M][CoutP] (plain 16-byte stores): the finishing
  // kernel adds the slices in a fixed order, so the forward value -- and with it every LeakyReLU mask
  // downstream -- does not depend on the order atomics would have landed in
  float* slice = acc + (int64_t)blockIdx.z * M * CoutP;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int mm = m0 + ty * 4 + i;
    int nn = n0 + tx * 4;
    if (mm < M && nn < CoutP)
      *reinterpret_cast<float4*>(slice + (int64_t)mm * CoutP + nn) = make_float4(c[i][0], c[i][1], c[i][2], c[i][3]);
  }
}
// acc [slices][M][CoutP] fp32 -> y blocked (T), y = [mask][lrelu](scale * sum_z acc[z] + bias).
// LPV lanes share one output vector (8 channels of one voxel): lane l adds slices l, l + LPV, ... and the partial sums
// meet in a shuffle tree -- a FIXED summation order (the forward value, and with it every LeakyReLU mask
// downstream, does not depend on the run), but `slices` / LPV dependent loads per thread instead of `slices`
// (the first version walked up to 72 slices serially on a 16-block grid: 45 us per launch, 11 launches per step).
template <typename T, int LPV>
__global__ void k_conv_finish(const float* __restrict__ acc, const float* __restrict__ bias,
                              const T* __restrict__ mask_src, T* __restrict__ y, int N, int Cout,
                              int CCout, int CoutP, int64_t V, float scale, int lrelu, int slices) {
  sg_pdl_enter();
  const int64_t total = (int64_t)N * CCout * V;
  const int lane = LPV == 1 ? 0 : (int)(threadIdx.x & (LPV - 1));
  const int64_t slice_stride = (int64_t)N * V * CoutP;
  // the loop bound is uniform per warp (whole groups of 32 / LPV vectors), so the full-mask shuffles below are safe
  // when `total` is not a multiple of the vectors per warp; lanes past the end just skip their loads and the store
  constexpr int VPW = 32 / (LPV == 1 ? 32 : LPV);   // vectors per warp step (LPV = 1: no shuffles, per-thread loop)
  const int64_t i0 = LPV == 1 ? blockIdx.x * (int64_t)blockDim.x + threadIdx.x
                              : ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32) * VPW;
  const int64_t istep = LPV == 1 ? (int64_t)gridDim.x * blockDim.x : ((int64_t)gridDim.x * blockDim.x / 32) * VPW;
  for (int64_t ib = i0; ib < total; ib += istep) {
    const int64_t i = LPV == 1 ? ib : ib + (threadIdx.x & 31) / LPV;
    const bool live = i < total;
    int64_t v = i % V;
    int64_t t = i / V;
    int cc = (int)(t % CCout);
    int64_t n = t / CCout;
    const float* a = acc + (n * V + v) * CoutP + cc * 8;
    float sum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = 0.f;
#pragma unroll 4
    for (int z = lane; z < (live ? slices : 0); z += LPV) {
      const float4 p0 = *reinterpret_cast<const float4*>(a + z * slice_stride);
      const float4 p1 = *reinterpret_cast<const float4*>(a + z * slice_stride + 4);
      sum[0] += p0.x; sum[1] += p0.y; sum[2] += p0.z; sum[3] += p0.w;
      sum[4] += p1.x; sum[5] += p1.y; sum[6] += p1.z; sum[7] += p1.w;
    }
    if (LPV > 1) {
#pragma unroll
      for (int o = LPV / 2; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] += __shfl_xor_sync(0xffffffffu, sum[j], o);
    }
    if (lane == 0 && live) {
      F8 o, m;
      if (mask_src) m = ld8(mask_src + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int co = cc * 8 + j;
        float r = 0.f;
        if (co < Cout) {
          r = sum[j] * scale + (bias ? __ldg(bias + co) : 0.f);
          if (lrelu) r = lrelu02(r);
          if (mask_src) r *= lmask02(m.v[j]);
        }
        o.v[j] = r;
      }
      st8(y + i * 8, o);
    }
  }
}
// `slices` partial sums [slices][N*V][CoutP] -> the activation tensor (bias, scale, LeakyReLU / mask applied once)
template <typename T>
static int conv_finish_launch(const float* acc, const float* bias, const void* mask_src, void* y, int N, int Cout,
                              int64_t V, float scale, int lrelu, int slices, cudaStream_t s) {
  int CCout = sg_chunks(Cout), CoutP = 16 * ((Cout + 15) / 16);
  int64_t total = (int64_t)N * CCout * V;
  if (slices >= 8)
    sg_launch((k_conv_finish<T, 8>), sg_grid(total * 8, 256), 256, 0, s, acc, bias, (const T*)mask_src, (T*)y, N, Cout,
              CCout, CoutP, V, scale, lrelu, slices);
  else
    sg_launch((k_conv_finish<T, 1>), sg_grid(total, 256), 256, 0, s, acc, bias, (const T*)mask_src, (T*)y, N, Cout, CCout,
              CoutP, V, scale, lrelu, slices);
  return sg_check_launch("sg_conv_finish");
}
int sg_conv_finish_bf16(const float* acc, const float* bias, const void* mask_src, void* y, int N,
                        int Cout, int64_t V, float scale, int lrelu, int slices, cudaStream_t s) {
  return conv_finish_launch<__nv_bfloat16>(acc, bias, mask_src, y, N, Cout, V, scale, lrelu, slices, s);
}
int sg_conv_finish_f32(const float* acc, const float* bias, const void* mask_src, void* y, int N,
                       int Cout, int64_t V, float scale, int lrelu, int slices, cudaStream_t s) {
  return conv_finish_launch<float>(acc, bias, mask_src, y, N, Cout, V, scale, lrelu, slices, s);
}

static bool small_f32_applies(int dtype, int N, int D, int H, int W) {
  return dtype == SG_DTYPE_F32 && (int64_t)D * H * W <= 128 && (int64_t)N * D * H * W <= 8192;
}
static int launch_small_f32(const void* x, const void* wp, const float* bias, const void* mask_src,
                            void* y, int N, int Cin, int Cout, int D, int H, int W, float scale,
                            int lrelu, void* ws, int64_t ws_bytes, cudaStream_t s) {
  int CCin = sg_chunks(Cin), CCout = sg_chunks(Cout), CoutP = 16 * ((Cout + 15) / 16);
  int64_t V = (int64_t)D * H * W, M = (int64_t)N * V;
  const int64_t mt = (M + 63) / 64, nt = (CoutP + 63) / 64;
  const SmallTaps st = small_taps(CCin, D, H, W, mt * nt);
  const int slices = st.nkd * st.nkh * st.nkw * st.ksplit;
  int64_t need = (int64_t)slices * M * CoutP * (int64_t)sizeof(float);
  SG_REQUIRE(ws != nullptr && ws_bytes >= need, "sg_conv3d_fprop: workspace too small (%lld < %lld)",
             (long long)ws_bytes, (long long)need);
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)slices);
  sg_launch((k_conv_small_f32), grid, 256, 0, s, (const float*)x, (const float*)wp, (float*)ws, N, CCin, CoutP, D, H, W, st);
  int rc = sg_check_launch("sg_conv3d_fprop(small f32)");
  if (rc) return rc;
  int64_t total = (int64_t)N * CCout * V;
  if (slices >= 8)
    sg_launch((k_conv_finish<float, 8>), sg_grid(total * 8, 256), 256, 0, s, (const float*)ws, bias, (const float*)mask_src,
              (float*)y, N, Cout, CCout, CoutP, V, scale, lrelu, slices);
  else
    sg_launch((k_conv_finish<float, 1>), sg_grid(total, 256), 256, 0, s, (const float*)ws, bias, (const float*)mask_src,
              (float*)y, N, Cout, CCout, CoutP, V, scale, lrelu, slices);
  return sg_check_launch("sg_conv3d_fprop(small f32 finish)");
}

// bytes of a bf16 copy of an fp32 act (256-byte aligned): impl = SG_IMPL_F32_AS_BF16 casts both wgrad operands once and
// runs the bf16 kernel on them (the fp32 tensors' half-chunk tensor maps move 16 bytes per TMA row: a wgrad, which
// reads every voxel once per launch, is bound by that; a 3x3x3 fprop re-uses its halo tile 27 times and is not)
static int64_t bf16_copy_bytes(int N, int C, int D, int H, int W) {
  return ((int64_t)N * sg_chunks(C) * D * H * W * 8 * 2 + 255) / 256 * 256;
}
__global__ void k_cast_f32_bf16(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t nvec) {
  sg_pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
    st8(dst + i * 8, ld8(src + i * 8));
}

extern "C" int64_t sg_conv3d_workspace_bytes(int kind, int dtype, int N, int Cin, int Cout, int D,
                                             int H, int W) {
  int64_t need = 0;
  if (kind == 0 && small_f32_applies(dtype, N, D, H, W)) {
    const int64_t M = (int64_t)N * D * H * W, CoutP = 16 * ((Cout + 15) / 16);
    const SmallTaps st = small_taps(sg_chunks(Cin), D, H, W, ((M + 63) / 64) * ((CoutP + 63) / 64));
    need = (int64_t)(st.nkd * st.nkh * st.nkw * st.ksplit) * M * CoutP * (int64_t)sizeof(float);
  }
  if (kind == 1 && small_f32_applies(dtype, N, D, H, W))
    need = (int64_t)27 * Cout * (16 * ((Cin + 15) / 16)) * (int64_t)sizeof(float);
  // fp32 tensors may take the tensor-core paths (impl = SG_IMPL_TF32 / SG_IMPL_F32_AS_BF16): size for whichever needs
  // more; the wgrad of impl 4 also keeps bf16 copies of both operands in the workspace
  int64_t t = sg_tc_workspace_bytes(kind, N, Cin, Cout, D, H, W, dtype == SG_DTYPE_F32);
  if (kind == 1 && dtype == SG_DTYPE_F32) {
    const int64_t tb = sg_tc_workspace_bytes(1, N, Cin, Cout, D, H, W, 0);      // the bf16 kernel behind impl 4
    if (tb > 0) {
      if (tb > t) t = tb;
      t += bf16_copy_bytes(N, Cin, D, H, W) + bf16_copy_bytes(N, Cout, D, H, W);
    }
  }
  if (t > need) need = t;
  return need;
}

template <typename T>
static int launch_direct_fprop(const void* x, const void* wp, const float* bias,
                               const void* mask_src, void* y, int N, int Cin, int Cout, int D,
                               int H, int W, float scale, int lrelu, cudaStream_t s) {
  int CCin = sg_chunks(Cin), CCout = sg_chunks(Cout);
  int CoutP = 16 * ((Cout + 15) / 16);
  int64_t V = (int64_t)D * H * W;
  dim3 grid((unsigned)((V + 127) / 128), (unsigned)CCout, (unsigned)N);
  sg_launch((k_conv_direct<T>), grid, 128, 0, s, (const T*)x, (const T*)wp, bias, (const T*)mask_src, (T*)y,
                                        Cout, CCin, CCout, CoutP, D, H, W, scale, lrelu);
  return sg_check_launch("sg_conv3d_fprop(direct)");
}

extern "C" int sg_conv3d_fprop(const void* x, const void* wp, const float* bias,
                               const void* mask_src, void* y, int dtype, int N, int Cin, int Cout,
                               int D, int H, int W, float scale, int lrelu, int impl, void* ws,
                               int64_t ws_bytes, cudaStream_t s) {
  SG_REQUIRE(N >= 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "sg_conv3d_fprop: bad shape");
  SG_REQUIRE(impl >= 0 && impl <= 3, "sg_conv3d_fprop: impl must be 0 (auto), 1 (direct), 2 (tcgen05), 3 (tcgen05 tf32)");
  if (N == 0) return 0;
  if (impl == SG_IMPL_TF32) {
    SG_REQUIRE(dtype == SG_DTYPE_F32, "sg_conv3d_fprop: the TF32 path takes fp32 activations (and the SG_TF32 packing)");
    int rc = sg_tc_fprop(x, wp, bias, mask_src, y, N, Cin, Cout, D, H, W, scale, lrelu, ws, ws_bytes, s, 1);
    if (rc == 1) sg_set_error("sg_conv3d_fprop: shape not covered by the tcgen05 tf32 kernel");
    return rc == 1 ? -5 : rc;     // no fallback: the packing differs from the CUDA-core kernels' (caller decides)
  }
  if (impl != SG_IMPL_DIRECT && dtype == SG_DTYPE_BF16) {
    int rc = sg_tc_fprop(x, wp, bias, mask_src, y, N, Cin, Cout, D, H, W, scale, lrelu, ws, ws_bytes, s, 0);
    if (rc != 1) return rc;
    SG_REQUIRE(impl != SG_IMPL_TCGEN05, "sg_conv3d_fprop: shape not covered by the tcgen05 kernel");
    if (impl == SG_IMPL_AUTO) ++g_cuda_core_fallbacks;
  } else {
    SG_REQUIRE(impl != SG_IMPL_TCGEN05, "sg_conv3d_fprop: tcgen05 path needs bf16 activations");
  }
  if (impl == SG_IMPL_AUTO && small_f32_applies(dtype, N, D, H, W))
    return launch_small_f32(x, wp, bias, mask_src, y, N, Cin, Cout, D, H, W, scale, lrelu, ws, ws_bytes, s);
  SG_DISPATCH(dtype, return launch_direct_fprop<T>(x, wp, bias, mask_src, y, N, Cin, Cout, D, H, W, scale, lrelu, s););
}

// conv3d + ChannelNormalization in one kernel (generator blocks, network.py:204-216): y = [lrelu](scale*conv + bias) is
// written for the backward, y_norm = [lrelu_after](y * rsqrt(mean_c y^2 + eps)) is what the next layer reads.  bf16
// only, shapes for which sg_conv3d_pixelnorm_supported() is 1 (weight-resident tcgen05 kernel, one N tile).
extern "C" int sg_conv3d_fprop_pixelnorm(const void* x, const void* wp, const float* bias, void* y, void* y_norm, int dtype,
                                         int N, int Cin, int Cout, int D, int H, int W, float scale, int lrelu,
                                         int lrelu_after, float eps, cudaStream_t s) {
  SG_REQUIRE(dtype == SG_DTYPE_BF16, "sg_conv3d_fprop_pixelnorm: bf16 activations only");
  SG_REQUIRE(y_norm != nullptr, "sg_conv3d_fprop_pixelnorm: y_norm is null");
  if (N == 0) return 0;
  int rc = sg_tc_fprop(x, wp, bias, nullptr, y, N, Cin, Cout, D, H, W, scale, lrelu, nullptr, 0, s, 0, y_norm, eps, lrelu_after);
  if (rc == 1) {
    sg_set_error("sg_conv3d_fprop_pixelnorm: shape not covered (ask sg_conv3d_pixelnorm_supported first)");
    return -6;
  }
  return rc;
}

// conv3d + AvgPool3d(2) in one kernel (discriminator blocks, network.py:88-90): y = [lrelu](scale*conv + bias) is still
// written (its sign is the LeakyReLU mask of the backward pass), y_pool = pool_scale * (2x2x2 block sums of y) is what
// the next level reads.  bf16 only, shapes for which sg_conv3d_pool_supported() is 1.
extern "C" int sg_conv3d_fprop_pool(const void* x, const void* wp, const float* bias, void* y, void* y_pool, int dtype, int N,
                                    int Cin, int Cout, int D, int H, int W, float scale, int lrelu, float pool_scale,
                                    cudaStream_t s) {
  SG_REQUIRE(dtype == SG_DTYPE_BF16, "sg_conv3d_fprop_pool: bf16 activations only");
  SG_REQUIRE(y_pool != nullptr, "sg_conv3d_fprop_pool: y_pool is null");
  if (N == 0) return 0;
  int rc = sg_tc_fprop(x, wp, bias, nullptr, y, N, Cin, Cout, D, H, W, scale, lrelu, nullptr, 0, s, 0, nullptr, 0.f, 0, y_pool,
                       pool_scale);
  if (rc == 1) {
    sg_set_error("sg_conv3d_fprop_pool: shape not covered (ask sg_conv3d_pool_supported first)");
    return -6;
  }
  return rc;
}

// ------------------------------------------------------------------------ direct wgrad
// gw[co][ci][tap] = scale * sum_{n,v} gy[n][co][v] * x[n][ci][v + tap - 1]
// block = (voxel slab, (cco, cci) chunk pair, tap); thread accumulates an 8x8 register tile
// over its voxels, warp-shuffle reduce, one atomic per output per warp.
template <typename T>
__global__ void __launch_bounds__(128)
k_wgrad_direct(const T* __restrict__ x, const T* __restrict__ gy, float* __restrict__ gw, int N,
               int Cin, int Cout, int CCin, int CCout, int D, int H, int W, float scale,
               int64_t per_slab) {
  sg_pdl_enter();
  int tap = blockIdx.z;
  int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  int cco = blockIdx.y / CCin;
  int cci = blockIdx.y % CCin;
  int64_t V = (int64_t)D * H * W;
  int64_t total = (int64_t)N * V;
  int64_t lo = blockIdx.x * per_slab;
  int64_t hi = lo + per_slab < total ? lo + per_slab : total;
  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    int64_t v = i % V;
    int64_t n = i / V;
    int w0 = (int)(v % W);
    int h0 = (int)((v / W) % H);
    int d0 = (int)(v / ((int64_t)W * H));
    int d = d0 + kd - 1, h = h0 + kh - 1, w_ = w0 + kw - 1;
    if (d < 0 || d >= D || h < 0 || h >= H || w_ < 0 || w_ >= W) continue;
    int64_t vin = ((int64_t)d * H + h) * W + w_;
    F8 g = ld8(gy + ((n * CCout + cco) * V + v) * 8);
    F8 a = ld8(x + ((n * CCin + cci) * V + vin) * 8);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(g.v[r], a.v[j], acc[r][j]);
  }
  int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = warp_sum(acc[r][j]);
      int co = cco * 8 + r, ci = cci * 8 + j;
      if (lane == 0 && co < Cout && ci < Cin && t != 0.f)
        atomicAdd(gw + ((int64_t)co * Cin + ci) * 27 + tap, t * scale);
    }
}


// ------------------------------------------------- small-volume fp32 wgrad (base level)
// gw[co][ci][tap] = scale * sum_m gy[m][co] * x[m + tap][ci] with only M = N*V rows (64 at B=4)
// but Cout*Cin*27 = 7 M outputs: an SGEMM with a tiny K.  block = 64 co x 64 ci x one tap,
// 4x4 register tile per thread, K = M streamed through shared memory in steps of 16; every
// output is owned by exactly one thread (no atomics, deterministic).  Output: tap-major workspace.
__global__ void __launch_bounds__(256)
k_wgrad_small_f32(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ ws,
                  int N, int Cin, int Cout, int CCin, int CCout, int D, int H, int W) {
  sg_pdl_enter();
  __shared__ float As[16][64 + 4];   // [m][co]
  __shared__ float Bs[16][64 + 4];   // [m][ci]
  const int tap = blockIdx.z;
  const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
  const int V = D * H * W, M = N * V;
  const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lm = t >> 4, l4 = (t & 15) * 4;     // loader: row m (0..15), 4 consecutive channels
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  // a tap that only ever sees padding (kd != 1 when D == 1, ...) has a zero gradient: skip the loop
  const bool dead = (D == 1 && kd != 1) || (H == 1 && kh != 1) || (W == 1 && kw != 1);
  for (int m0 = 0; m0 < (dead ? 0 : M); m0 += 16) {
    const int m = m0 + lm;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < M) {
      const int n = m / V, v = m % V;
      const int co = co0 + l4, ci = ci0 + l4;
      if (co < CCout * 8)
        av = *reinterpret_cast<const float4*>(gy + (((int64_t)n * CCout + (co >> 3)) * V + v) * 8 + (co & 7));
      const int w0 = v % W, h0 = (v / W) % H, d0 = v / (W * H);
      const int d = d0 + kd - 1, h = h0 + kh - 1, w_ = w0 + kw - 1;
      if (ci < CCin * 8 && d >= 0 && d < D && h >= 0 && h < H && w_ >= 0 && w_ < W)
        bv = *reinterpret_cast<const float4*>(
            x + (((int64_t)n * CCin + (ci >> 3)) * V + ((int64_t)d * H + h) * W + w_) * 8 + (ci & 7));
    }
    As[lm][l4 + 0] = av.x; As[lm][l4 + 1] = av.y; As[lm][l4 + 2] = av.z; As[lm][l4 + 3] = av.w;
    Bs[lm][l4 + 0] = bv.x; Bs[lm][l4 + 1] = bv.y; Bs[lm][l4 + 2] = bv.z; Bs[lm][l4 + 3] = bv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fmaf(a[i], b[j], c[i][j]);
    }
    __syncthreads();
  }
  // tap-major workspace [27][Cout][CinP] (16-byte coalesced stores); k_wgrad_finish transposes it into the
  // parameter layout [Cout][Cin][27] -- 4-byte stores 108 bytes apart straight into that layout cost 3x more
  const int CinP = CCin * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    const int ci = ci0 + tx * 4;
    if (co < Cout && ci < CinP)
      *reinterpret_cast<float4*>(ws + ((int64_t)tap * Cout + co) * CinP + ci) = make_float4(c[i][0], c[i][1], c[i][2], c[i][3]);
  }
}

template <typename T>
static int launch_direct_wgrad(const void* x, const void* gy, float* gw, int N, int Cin, int Cout,
                               int D, int H, int W, float scale, cudaStream_t s) {
  int CCin = sg_chunks(Cin), CCout = sg_chunks(Cout);
  int64_t total = (int64_t)N * D * H * W;
  int64_t pairs = (int64_t)CCin * CCout * 27;
  int64_t want = ((int64_t)sg_num_sms() * 16 + pairs - 1) / pairs;
  int64_t slabs = (total + 1023) / 1024;
  if (slabs > want) slabs = want;
  if (slabs < 1) slabs = 1;
  int64_t per = (total + slabs - 1) / slabs;
  slabs = (total + per - 1) / per;
  SG_REQUIRE((int64_t)CCin * CCout <= 65535, "sg_conv3d_wgrad(direct): too many channel chunks");
  dim3 grid((unsigned)slabs, (unsigned)(CCin * CCout), 27);
  sg_launch((k_wgrad_direct<T>), grid, 128, 0, s, (const T*)x, (const T*)gy, gw, N, Cin, Cout, CCin, CCout,
                                         D, H, W, scale, per);
  return sg_check_launch("sg_conv3d_wgrad(direct)");
}

// gb[co] = sum_{n,v} gy[n][co][v]
extern "C" int sg_conv3d_wgrad(const void* x, const void* gy, float* gw, float* gb, int dtype,
                               int N, int Cin, int Cout, int D, int H, int W, float scale, int impl,
                               void* ws, int64_t ws_bytes, cudaStream_t s) {
  SG_REQUIRE(N >= 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, "sg_conv3d_wgrad: bad shape");
  SG_REQUIRE(impl >= 0 && impl <= 4, "sg_conv3d_wgrad: impl must be 0 (auto), 1 (direct), 2 (tcgen05), 3 (fp32 tensors, "
             "split-bf16 tensor cores), 4 (fp32 tensors, bf16 operands)");
  if (impl == SG_IMPL_F32_AS_BF16 && N > 0 && sg_tc_workspace_bytes(1, N, Cin, Cout, D, H, W, 0) > 0) {
    // fp32 tensors, bf16 operands: cast both once into the workspace, then the bf16 tensor-core kernel
    SG_REQUIRE(dtype == SG_DTYPE_F32, "sg_conv3d_wgrad: impl 4 takes fp32 activations");
    const int64_t xb = bf16_copy_bytes(N, Cin, D, H, W), gyb = bf16_copy_bytes(N, Cout, D, H, W);
    SG_REQUIRE(ws != nullptr && ws_bytes >= xb + gyb, "sg_conv3d_wgrad(impl 4): workspace too small");
    const int64_t nx = (int64_t)N * sg_chunks(Cin) * D * H * W, ng = (int64_t)N * sg_chunks(Cout) * D * H * W;
    sg_launch((k_cast_f32_bf16), sg_grid(nx, 256), 256, 0, s, (const float*)x, (__nv_bfloat16*)ws, nx);
    int rc = sg_check_launch("sg_conv3d_wgrad(cast x)");
    if (rc) return rc;
    sg_launch((k_cast_f32_bf16), sg_grid(ng, 256), 256, 0, s, (const float*)gy, (__nv_bfloat16*)((char*)ws + xb), ng);
    rc = sg_check_launch("sg_conv3d_wgrad(cast gy)");
    if (rc) return rc;
    rc = sg_tc_wgrad(ws, (char*)ws + xb, gw, gb, N, Cin, Cout, D, H, W, scale, (char*)ws + xb + gyb, ws_bytes - xb - gyb, s, 0);
    if (rc != 1) return rc;
    ++g_cuda_core_fallbacks;
    impl = SG_IMPL_AUTO;
  }
  if ((impl == SG_IMPL_TF32 || impl == SG_IMPL_F32_AS_BF16) && N > 0) {
    SG_REQUIRE(dtype == SG_DTYPE_F32, "sg_conv3d_wgrad: impl 3 / 4 take fp32 activations");
    int rc = sg_tc_wgrad(x, gy, gw, gb, N, Cin, Cout, D, H, W, scale, ws, ws_bytes, s, impl == SG_IMPL_TF32 ? 1 : 2);
    if (rc != 1) return rc;
    ++g_cuda_core_fallbacks;      // shape not covered: the fp32 CUDA-core kernels below take the same operands
    impl = SG_IMPL_AUTO;
  }
  if (impl != SG_IMPL_DIRECT && dtype == SG_DTYPE_BF16 && N > 0) {
    int rc = sg_tc_wgrad(x, gy, gw, gb, N, Cin, Cout, D, H, W, scale, ws, ws_bytes, s, 0);
    if (rc != 1) return rc;
    SG_REQUIRE(impl != SG_IMPL_TCGEN05, "sg_conv3d_wgrad: shape not covered by the tcgen05 kernel");
    if (impl == SG_IMPL_AUTO) ++g_cuda_core_fallbacks;
  } else {
    SG_REQUIRE(impl != SG_IMPL_TCGEN05 || N == 0, "sg_conv3d_wgrad: tcgen05 path needs bf16 activations");
  }
  if (gb) {
    int rc = sg_pw_wgrad(gy, nullptr, nullptr, gb, dtype, N, Cout, (int64_t)D * H * W, 1.f, s);
    if (rc) return rc;
  }
  if (impl == SG_IMPL_AUTO && N > 0 && small_f32_applies(dtype, N, D, H, W)) {
    int CCin = sg_chunks(Cin), CCout = sg_chunks(Cout);
    const int64_t need = (int64_t)27 * Cout * CCin * 8 * (int64_t)sizeof(float);
    SG_REQUIRE(ws != nullptr && ws_bytes >= need, "sg_conv3d_wgrad(small f32): workspace too small (%lld < %lld)",
               (long long)ws_bytes, (long long)need);
    dim3 grid((unsigned)((Cout + 63) / 64), (unsigned)((CCin * 8 + 63) / 64), 27);
    sg_launch((k_wgrad_small_f32), grid, 256, 0, s, (const float*)x, (const float*)gy, (float*)ws, N, Cin, Cout, CCin, CCout,
              D, H, W);
    int rc = sg_check_launch("sg_conv3d_wgrad(small f32)");
    if (rc) return rc;
    return sg_wgrad_finish((const float*)ws, gw, Cout, Cin, CCin * 8, scale, s);
  }
  cudaMemsetAsync(gw, 0, sizeof(float) * (size_t)Cout * Cin * 27, s);
  if (N == 0) return 0;
  SG_DISPATCH(dtype, return launch_direct_wgrad<T>(x, gy, gw, N, Cin, Cout, D, H, W, scale, s););
}
