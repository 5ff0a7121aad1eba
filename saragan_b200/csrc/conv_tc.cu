// tcgen05 / TMEM / TMA implicit-GEMM 3x3x3 convolution (fprop; dgrad through the flipped
// packing) for sm_100a, bf16 operands, fp32 accumulation in tensor memory.
//
// Design (DESIGN.md "conv3d"):
//  * activations live in the channel-blocked layout [N][C/8][D][H][W][8]; one TMA tiled load
//    per 8-channel chunk brings a HALO block (tn x (td+2) x (th+2) x (tw+2) voxels, out-of-
//    bounds = zero = the conv padding) into shared memory as [line][w][8ch], i.e. 16 bytes per
//    voxel and one "line" per (n,d,h);
//  * that block is exactly a K-major, un-swizzled UMMA operand: 8 consecutive voxels of a line
//    are one 8x16B core matrix, lines are SBO apart, channel chunks LBO apart.  The A operand
//    of tap (kd,kh,kw) is the SAME block with the descriptor start address shifted by
//    ((kd-1)*(th+2) + (kh-1)) lines + kw voxels -- the 27 taps re-use one halo load instead of
//    27 im2col loads (L2->SMEM traffic /9);
//  * an MMA covers 16 consecutive lines x 8 voxels = 128 rows (M=128) and N = NT output
//    channels; rows that fall on halo lines/columns are computed and discarded;
//  * weights [tap][C/8][CoutP][8] stream through a ring of TMA stages of `tps` taps each: ONE tiled load
//    (box = [tps taps][K-block chunks][NT channels]) per stage -- issuing a copy costs the producer thread
//    ~90 cycles plus ~360 for the barrier bookkeeping (tools/bulk_bench.cu), so per-tap stages of four
//    2 KB bulk copies (the first version) capped the whole kernel at ~1300 cycles per tap;
//  * warp roles: warp 0 = TMA/bulk producer, warp 1 = TMEM allocator + single-thread MMA
//    issuer, warps 2-5 = epilogue (tcgen05.ld -> scale, bias, LeakyReLU, mask -> 16-byte stores,
//    or plain fp32 stores into this K slice's slab of the split-K workspace, summed in a fixed order by k_conv_finish).
#include <cuda.h>

#include <array>
#include <map>
#include <mutex>
#include <type_traits>

#include "../../include/saragan_b200.h"
#include "common.cuh"

int sg_conv_finish_bf16(const float* acc, const float* bias, const void* mask_src, void* y, int N,
                        int Cout, int64_t V, float scale, int lrelu, int slices, cudaStream_t s);

#include "tc_common.cuh"
SG_DEFINE_LEAK_SETTER(sg_set_leak_conv_tc)

namespace {

constexpr int kThreads = 192;

struct TcParams {
  const void* wp;            // packed weights [27][CCin][CoutP][16 bytes]
  const float* bias;         // nullable
  const void* mask;          // nullable, act of y's type
  void* y;                   // output act (splits == 1): bf16, or fp32 for the TF32 kernel
  float* ws;                 // fp32 [N*V][CoutP] (splits > 1)
  int N, D, H, W;
  int CCin;                  // 16-byte K chunks of the input: 8 bf16 channels, or 4 fp32 channels (TF32)
  int Cout, CoutP, CCout;    // CCout: 8-channel chunks of the output tensor (both types)
  int td, th, tn;            // tile extents in output voxels (tw == 8)
  int tiles_w, tiles_h, tiles_d, tiles_n;
  int halo_w, halo_h, halo_d;
  int chunk_bytes;           // stride of one 8-channel halo block (128-byte aligned for the TMA)
  int chunk_tx_bytes;        // bytes the TMA actually writes per block
  int n_sub;
  int sub_line[kMaxSub];     // first halo line (centre coordinates) of each 16-line MMA tile
  int kb_chunks;             // 8-channel chunks per K block (2 or 4)
  int kblocks_per_split;
  int splits;
  int sw;                    // weight ring stages
  int a_bytes, w_stage_bytes;
  int tps;                   // taps per weight stage (9, 3 or 1)
  int w_tap_bytes;           // kb_chunks * NT * 16
  int a_stages;              // 1 or 2 halo-block buffers
  int tmem_cols;
  float scale;
  int lrelu;
};

// greedy cover of a tile's valid output lines with 16-line MMA tiles (same rule as cover_lines on the host),
// evaluated at compile time for the specialised kernels
struct CoverCE {
  int n;
  int line[kMaxSub + 1];
};
constexpr CoverCE cover_ce(int tn, int td, int th) {
  CoverCE c{};
  const int halo_d = td + 2, halo_h = th + 2;
  int covered_to = -1;
  for (int nl = 0; nl < tn; ++nl)
    for (int dl = 0; dl < td; ++dl)
      for (int hl = 0; hl < th; ++hl) {
        const int line = (nl * halo_d + dl + 1) * halo_h + hl + 1;
        if (line <= covered_to) continue;
        if (c.n < kMaxSub) c.line[c.n] = line;
        ++c.n;
        covered_to = line + 15;
      }
  return c;
}

// --------------------------------------------------------------------------------- kernel
// TD = 0: generic geometry (tile extents, K block, taps per stage from TcParams at run time).
// TD > 0: geometry fixed at compile time -- tile TN x TD x TH x 8 voxels, K block of 2 chunks, 9 taps per
// weight stage -- so that every descriptor offset of the 27 x n_sub MMAs of a K block is a constant and
// the fully unrolled issue loop costs one uniform add per MMA (the generic loop costs the single issuing
// warp ~200 cycles per tap plus ~100 per MMA: more than the MMAs themselves).
// TF32: fp32 activations (the 8-blocked fp32 layout, read through TWO tensor maps -- one per 16-byte half of a
// voxel's 32-byte chunk -- so that a shared-memory voxel is again one 16-byte core-matrix row, now of 4 channels),
// fp32 packed weights [27][C/4][CoutP][4] pre-rounded to tf32, kind::tf32 MMAs (K = 8 = two 16-byte chunks, exactly
// the bf16 kernel's byte geometry), fp32 output.  The tensor core would TRUNCATE the fp32 activations to tf32: the
// four epilogue warps, idle during the main loop, round every landed halo block in place (tf32_rna) and hand it to
// the MMA issuer through the ROUND_A barriers.
template <int NT, int TD, int TH, int TN, bool TF32>
__global__ void __launch_bounds__(kThreads)
k_conv_tc(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap1,
          const __grid_constant__ CUtensorMap wmap, const __grid_constant__ TcParams p) {
  using T = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  extern __shared__ __align__(128) uint8_t smem[];
  sg_pdl_trigger();
  // carve-up: [A halo block][weight ring][barriers][tmem base]
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + p.a_stages * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.sw * p.w_stage_bytes);
  // bars: [0,1] full_a, [2,3] empty_a, [4] acc_full, [5 .. 5+sw) full_w, [5+sw .. 5+2sw) empty_w   (sw <= 8),
  //       [22,23] round_a (TF32)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int FULL_A = 0, EMPTY_A = 2, ACC_FULL = 4, FULL_W = 5, EMPTY_W = 5 + p.sw, ROUND_A = 22;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tile_w = t % p.tiles_w; t /= p.tiles_w;
  const int tile_h = t % p.tiles_h; t /= p.tiles_h;
  const int tile_d = t % p.tiles_d; t /= p.tiles_d;
  const int tile_n = t;
  const int w0 = tile_w * 8, h0 = tile_h * p.th, d0 = tile_d * p.td, n0 = tile_n * p.tn;
  const int co0 = blockIdx.y * NT;
  const int kb0 = blockIdx.z * p.kblocks_per_split;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    if constexpr (TF32) asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(FULL_A + i), 1);
      mbar_init(BAR(EMPTY_A + i), 1);
      mbar_init(BAR(ROUND_A + i), 128);   // every thread of the four epilogue warps
    }
    mbar_init(BAR(ACC_FULL), 1);
    for (int i = 0; i < p.sw; ++i) {
      mbar_init(BAR(FULL_W + i), 1);
      mbar_init(BAR(EMPTY_W + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  sg_pdl_wait();   // everything above touched only shared / tensor memory and kernel parameters

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const uint32_t a_addr = smem_u32(a_smem);
      const uint32_t w_addr = smem_u32(w_smem);
      const uint32_t a_tx = (uint32_t)p.kb_chunks * (uint32_t)p.chunk_tx_bytes;
      const uint32_t w_tx = (uint32_t)p.w_stage_bytes;
      const int groups = 27 / p.tps;
      int it = 0;
      for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
        // the halo block of K block kb + 1 is loaded (second A stage) while the MMAs of kb still run
        const int sa = kb % p.a_stages;
        mbar_wait(BAR(EMPTY_A + sa), ((kb / p.a_stages) & 1) ^ 1);
        const int chunk0 = (kb0 + kb) * p.kb_chunks;
        mbar_expect_tx(BAR(FULL_A + sa), a_tx);
        for (int c = 0; c < p.kb_chunks; ++c) {
          if constexpr (TF32) {
            // fp32 tensor as [4 | W | H | D | N*CC8] (voxel pitch 32 bytes), one map per 16-byte half: tn == 1
            const int q = chunk0 + c;
            tma_load_5d(a_addr + sa * p.a_bytes + c * p.chunk_bytes, (q & 1) ? &xmap1 : &xmap, BAR(FULL_A + sa), 0, w0 - 1,
                        h0 - 1, d0 - 1, n0 * (p.CCin >> 1) + (q >> 1));
          } else {
            tma_load_5d(a_addr + sa * p.a_bytes + c * p.chunk_bytes, &xmap, BAR(FULL_A + sa), (w0 - 1) * 8, h0 - 1, d0 - 1,
                        chunk0 + c, n0);
          }
        }
        for (int g = 0; g < groups; ++g, ++it) {
          const int s = it % p.sw;
          mbar_wait(BAR(EMPTY_W + s), ((it / p.sw) & 1) ^ 1);
          mbar_expect_tx(BAR(FULL_W + s), w_tx);
          // weights as [64 = 8 co x 8 ci | CoutP/8 | CCin | 27]: box [tps][kb_chunks][NT/8][64] lands as [tap][chunk][co][8]
          tma_load_4d(w_addr + s * p.w_stage_bytes, &wmap, BAR(FULL_W + s), 0, co0 / 8, chunk0, g * p.tps);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t leader = elect_one();   // all lanes run the loops; one issues
      // instruction descriptor: D=f32, A=B=bf16 (tf32), both K-major, N = NT, M = 128
      const uint32_t idesc = idesc_formats(TF32) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const int A_READY = TF32 ? ROUND_A : FULL_A;   // TF32: the halo block is usable once it has been rounded
      const uint32_t line_pitch = (uint32_t)p.halo_w * 16u;
      // hoisted descriptor pieces (16-byte units): the first valid line sits (halo_h + 1) lines into
      // the block, which is also the most negative tap offset, so both offset families are >= 0
      const int bias_vox = (p.halo_h + 1) * p.halo_w;
      uint32_t sub_off[kMaxSub];
#pragma unroll
      for (int i = 0; i < kMaxSub; ++i) sub_off[i] = i < p.n_sub ? (uint32_t)(p.sub_line[i] * p.halo_w - bias_vox) : 0u;
      const uint64_t a_desc0 = make_desc(smem_u32(a_smem), (uint32_t)p.chunk_bytes, line_pitch);
      const uint64_t w_desc0 = make_desc(smem_u32(w_smem), NT * 16u, 128u);
      const uint32_t kk_a = (uint32_t)(2 * p.chunk_bytes) >> 4;
      const uint32_t w_stage16 = (uint32_t)p.w_stage_bytes >> 4, w_tap16 = (uint32_t)p.w_tap_bytes >> 4;
      const int tps = p.tps;
      int tl = 0;   // tap within the current weight stage
      const int kpairs = p.kb_chunks / 2, n_sub = p.n_sub, halo_w = p.halo_w, halo_h = p.halo_h, sw = p.sw;
      int s = 0, ph = 0;
      if constexpr (TD > 0) {
        constexpr CoverCE cov = cover_ce(TN, TD, TH);
        constexpr int HALO_H = TH + 2, HALO_W = 10;
        constexpr int BIAS_VOX = (HALO_H + 1) * HALO_W;
        constexpr uint32_t W_TAP16 = 2u * NT;   // one tap of a stage: 2 chunks x NT rows x 16 B
        for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
          const int sa = kb % p.a_stages;
          mbar_wait(BAR(A_READY + sa), (kb / p.a_stages) & 1);
          const uint64_t a_kb = a_desc0 + (uint64_t)((uint32_t)(sa * p.a_bytes) >> 4);
          const uint32_t acc0 = kb != 0;
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            mbar_wait(BAR(FULL_W + s), ph);
            tc_fence_after();
            const uint64_t b_stage = w_desc0 + (uint64_t)(s * w_stage16);
#pragma unroll
            for (int t9 = 0; t9 < 9; ++t9) {
              constexpr int dummy = 0;
              (void)dummy;
              const int tap = g * 9 + t9;
              // tap (kd,kh,kw) reads halo voxel (line + (kd-1)*HALO_H + (kh-1), w + kw): all offsets >= 0 after
              // folding the first valid line's position (BIAS_VOX) into the sub-tile offsets
              const uint32_t tap_off = (uint32_t)((((tap / 9) - 1) * HALO_H + ((tap / 3) % 3 - 1)) * HALO_W + BIAS_VOX + tap % 3);
#pragma unroll
              for (int sub = 0; sub < cov.n; ++sub)
                tc_mma_t<TF32>(tmem_base + sub * NT, a_kb + (uint64_t)(tap_off + (uint32_t)(cov.line[sub] * HALO_W - BIAS_VOX)),
                               b_stage + (uint64_t)(t9 * W_TAP16), idesc, tap ? 1u : acc0, leader);
            }
            tc_commit(BAR(EMPTY_W + s), leader);
            if (++s == sw) { s = 0; ph ^= 1; }
          }
          tc_commit(BAR(EMPTY_A + sa), leader);
        }
      } else {
      for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
        const int sa = kb % p.a_stages;
        mbar_wait(BAR(A_READY + sa), (kb / p.a_stages) & 1);
        const uint64_t a_desc_kb = a_desc0 + (uint64_t)((uint32_t)(sa * p.a_bytes) >> 4);
        int tap = 0;
        for (int kd = 0; kd < 3; ++kd)
          for (int kh = 0; kh < 3; ++kh) {
            const int row_off = ((kd - 1) * halo_h + (kh - 1)) * halo_w + bias_vox;
            for (int kw = 0; kw < 3; ++kw, ++tap) {
              if (tl == 0) {
                mbar_wait(BAR(FULL_W + s), ph);
                tc_fence_after();
              }
              issue_tap<NT, TF32>(tmem_base, a_desc_kb + (uint64_t)(uint32_t)(row_off + kw),
                            w_desc0 + (uint64_t)(s * w_stage16 + tl * w_tap16), sub_off, n_sub, kpairs, kk_a, idesc,
                            (kb | tap) != 0, leader);
              if (++tl == tps) {
                tl = 0;
                tc_commit(BAR(EMPTY_W + s), leader);
                if (++s == sw) { s = 0; ph ^= 1; }
              }
            }
          }
        tc_commit(BAR(EMPTY_A + sa), leader);
      }
      }
      tc_commit(BAR(ACC_FULL), leader);
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;       // accumulator row == TMEM lane
    const float scale = p.scale;
    const int lrelu = p.lrelu;
    const T* mask = reinterpret_cast<const T*>(p.mask);
    T* yout = reinterpret_cast<T*>(p.y);
    float* s_bias = reinterpret_cast<float*>(bars + 26);   // NT floats, 16-byte aligned, after the barriers
    for (int i = row; i < NT; i += 128) s_bias[i] = (p.bias && co0 + i < p.Cout) ? p.bias[co0 + i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
    if constexpr (TF32) {
      // main loop duty of these warps: round every landed halo block to tf32 in place, then release it to the issuer
      for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
        const int sa = kb % p.a_stages;
        mbar_wait(BAR(FULL_A + sa), (kb / p.a_stages) & 1);
        round_tile_tf32(a_smem + sa * p.a_bytes, p.kb_chunks * p.chunk_bytes, row, 128);
        mbar_arrive(BAR(ROUND_A + sa));
      }
    }
    mbar_wait(BAR(ACC_FULL), 0);
    tc_fence_after();
    const int64_t V = (int64_t)p.D * p.H * p.W;
    const int plane_lines = p.halo_d * p.halo_h;
    for (int sub = 0; sub < p.n_sub; ++sub) {
      const int line = p.sub_line[sub] + (row >> 3);
      const int wl = row & 7;
      const int nl = line / plane_lines;
      const int rem = line - nl * plane_lines;
      const int dh = rem / p.halo_h, hh = rem - dh * p.halo_h;
      const int n = n0 + nl, d = d0 + dh - 1, h = h0 + hh - 1, w = w0 + wl;
      const bool valid = nl < p.tn && dh >= 1 && dh <= p.td && hh >= 1 && hh <= p.th && n < p.N && d < p.D &&
                         h < p.H && w < p.W;
      const int64_t vox = ((int64_t)d * p.H + h) * p.W + w;
      constexpr int CB = NT < 64 ? NT : 64;   // accumulator columns per TMEM round trip
      const bool direct = valid && p.splits == 1;
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += CB) {
        const int64_t o = (((int64_t)n * p.CCout + (co0 + c0) / 8) * V + vox) * 8;
        // the LeakyReLU-mask vectors of these columns are fetched before the accumulator round trip
        uint4 mk[CB / 8];
        if constexpr (!TF32) {
          if (direct && mask) {
#pragma unroll
            for (int q = 0; q < CB / 8; ++q) mk[q] = __ldg(reinterpret_cast<const uint4*>(mask + o + (int64_t)q * V * 8));
          }
        }
        float v[CB];
        __syncwarp();   // tcgen05.ld is .sync.aligned: the warp must be converged here
        tmem_ld_block<CB>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * NT + c0), v);
        if (!valid) {
          // row lies on a halo line / column or outside the volume: computed, discarded
        } else if (p.splits > 1) {
          // split-K: every K slice owns a slab of the workspace (plain 16-byte stores, no zero-fill, and the finishing
          // kernel adds the slabs in a fixed order: the result does not depend on the run, unlike fp32 atomics)
          float* dst = p.ws + (((int64_t)blockIdx.z * p.N + n) * V + vox) * p.CoutP + co0 + c0;
#pragma unroll
          for (int q = 0; q < CB / 4; ++q)
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < CB / 16; ++j) {
            if constexpr (TF32)
              epilogue16<float>(v + 16 * j, s_bias + c0 + 16 * j, scale, lrelu, mask ? mask + o + (int64_t)(2 * j) * V * 8 : nullptr,
                                yout + o + (int64_t)(2 * j) * V * 8, V * 8);
            else
              epilogue16_regmask(v + 16 * j, s_bias + c0 + 16 * j, scale, lrelu, mask != nullptr, mk[2 * j], mk[2 * j + 1],
                                 yout + o + (int64_t)(2 * j) * V * 8, V * 8);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
struct Plan {
  bool ok = false;
  int NT = 0;
  TcParams p{};
  size_t smem = 0;
  dim3 grid;
  double cost = 0;   // estimated cycles (make_plan_cfg)
  bool big = false;
  bool spec = false; // compile-time-geometry kernel (k_conv_tc<NT, TD, TH, TN>)
};

// cost-model constants (cycles at ~1.9 GHz)
constexpr double kL2BytesPerCycle = 4500.0;   // sustained L2 -> SM bytes per cycle, whole chip
constexpr double kCtaFixedCycles = 6000.0;    // launch, tensor-map fetch, first loads, epilogue
constexpr double kSplitFixedCycles = 12000.0; // finishing kernel of a split-K launch
constexpr double kLoadLatencyCycles = 4000.0; // TMA round trip seen by a stage whose data is not in L2 yet
constexpr double kStageIssueCycles = 500.0;   // floor per weight stage: TMA issue + barrier round (tools/bulk_bench.cu)

int round_pow2_cols(int c) {
  int r = 32;
  while (r < c) r <<= 1;
  return r;
}

// greedy cover of the tile's valid output lines with 16-line MMA tiles; returns the count
int cover_lines(int tn, int td, int th, int halo_d, int halo_h, int* out) {
  int n_sub = 0, covered_to = -1;
  for (int nl = 0; nl < tn; ++nl)
    for (int dl = 0; dl < td; ++dl)
      for (int hl = 0; hl < th; ++hl) {
        int line = (nl * halo_d + dl + 1) * halo_h + hl + 1;
        if (line <= covered_to) continue;
        if (n_sub == kMaxSub) return kMaxSub + 1;
        out[n_sub++] = line;
        covered_to = line + 15;
      }
  return n_sub;
}

// One candidate tiling of the streaming kernel.  `big` = sized for ONE CTA per SM (<= 512 TMEM columns,
// ~200 KB of shared memory) instead of two (<= 256 columns, ~110 KB each); NT = output channels per CTA;
// td_max caps the planes per tile; splits = split-K factor (must divide the K-block count).
Plan make_plan_cfg(int N, int Cin, int Cout, int D, int H, int W, int NT, bool big, int td_max, int splits,
                   bool tf32 = false) {
  Plan pl;
  TcParams& p = pl.p;
  if (W % 8 != 0 || H < 8) return pl;
  // 16-byte K chunks: 8 bf16 channels, or 4 fp32 channels (TF32) -- from here on the geometry is in bytes and equal
  const int CCin = sg_chunks(Cin) * (tf32 ? 2 : 1), CoutP = 16 * ((Cout + 15) / 16);
  if (CoutP % NT != 0) return pl;
  const int tmem_budget = big ? 512 : 256;
  const int max_sub = tmem_budget / NT < kMaxSub ? tmem_budget / NT : kMaxSub;
  int th = H < 16 ? H : 16;
  if (H % th != 0) return pl;
  const int a_cap = (big ? 96 : 72) * 1024;
  // largest (tn, td) whose MMA tiles fit the TMEM budget
  int best_tn = 0, best_td = 0, best_sub = 0, best_lines[kMaxSub];
  for (int td = 1; td <= D && td <= 8 && td <= td_max; ++td) {
    if (D % td) continue;
    for (int tn = 1; tn <= N && tn <= 8; ++tn) {
      if (td < D && tn > 1) continue;   // span samples only when a tile already holds a whole volume
      if (tf32 && tn > 1) continue;     // the fp32 tensor maps merge (sample, chunk) into one dimension
      if (tn > 1 && tn * td > td_max) continue;
      int lines[kMaxSub];
      int ns = cover_lines(tn, td, th, td + 2, th + 2, lines);
      if (ns > max_sub) continue;
      if (tn * (td + 2) * (th + 2) * 10 * 16 * 2 > a_cap) continue;   // >= 2 chunks per K block must fit
      if (tn * td > best_tn * best_td) {
        best_tn = tn; best_td = td; best_sub = ns;
        for (int i = 0; i < ns; ++i) best_lines[i] = lines[i];
      }
    }
  }
  if (best_sub == 0) return pl;
  p.td = best_td; p.th = th; p.tn = best_tn;
  p.n_sub = best_sub;
  for (int i = 0; i < best_sub; ++i) p.sub_line[i] = best_lines[i];
  p.halo_w = 10; p.halo_h = th + 2; p.halo_d = p.td + 2;
  p.chunk_tx_bytes = p.tn * p.halo_d * p.halo_h * p.halo_w * 16;
  p.chunk_bytes = (p.chunk_tx_bytes + 127) / 128 * 128;
  // K block (4 or 2 chunks), taps per weight stage (9, 3, 1) and ring depth: the largest stage that still
  // leaves a ring of >= 3 (else >= 2) stages; two halo buffers when more than one K block follows
  const int total_budget = (big ? 200 : 110) * 1024 - 256;
  static const int cand[6][2] = {{4, 9}, {2, 9}, {4, 3}, {2, 3}, {4, 1}, {2, 1}};
  // geometries with a specialised (fully unrolled) kernel: those always use K blocks of 2 chunks, 9 taps per stage
  const bool spec_geom = (NT == 128 || NT == 64) &&
                         ((th == 16 && p.tn == 1 && (p.td == 1 || p.td == 2 || p.td == 4)) ||
                          (th == 8 && p.td == 2 && (p.tn == 1 || p.tn == 2)));
  int pick = -1, pick_sw = 0, pick_as = 1;
  for (int want_sw = 3; want_sw >= 2 && pick < 0; --want_sw)
    for (int i = 0; i < 6 && pick < 0; ++i) {
      const int kb = cand[i][0], tps = cand[i][1];
      if (spec_geom && i != 1) continue;
      if (CCin % kb || kb * p.chunk_bytes > a_cap) continue;
      const int n_kb = CCin / kb;
      if (splits < 1 || n_kb % splits) continue;
      const int a_b = (kb * p.chunk_bytes + 127) / 128 * 128, stage = tps * kb * NT * 16;
      int as = n_kb / splits > 1 ? 2 : 1;
      int sw = (total_budget - as * a_b) / stage;
      if (sw < want_sw && as == 2) { as = 1; sw = (total_budget - a_b) / stage; }
      if (sw < want_sw) continue;
      pick = i; pick_sw = sw > 8 ? 8 : sw; pick_as = as;
    }
  pl.spec = spec_geom && pick == 1;
  if (pick < 0) return pl;
  p.kb_chunks = cand[pick][0];
  p.tps = cand[pick][1];
  p.a_bytes = (p.kb_chunks * p.chunk_bytes + 127) / 128 * 128;
  p.w_tap_bytes = p.kb_chunks * NT * 16;
  p.w_stage_bytes = p.tps * p.w_tap_bytes;
  p.a_stages = pick_as;
  const int n_kblocks = CCin / p.kb_chunks;
  p.splits = splits;
  p.kblocks_per_split = n_kblocks / splits;
  const int sw = pick_sw;
  // garbage rows of the last MMA tile may read past the block: keep those reads inside the allocation
  int last_line = p.sub_line[p.n_sub - 1] + 15 + p.halo_h + 1;
  int over = (last_line + 1) * p.halo_w * 16 + 8 * 16 + (p.kb_chunks - 1) * p.chunk_bytes - p.a_bytes;
  if (over > sw * p.w_stage_bytes) return pl;
  p.sw = sw;
  p.tiles_w = W / 8; p.tiles_h = H / th; p.tiles_d = D / p.td; p.tiles_n = (N + p.tn - 1) / p.tn;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.CCin = CCin; p.Cout = Cout; p.CoutP = CoutP; p.CCout = sg_chunks(Cout);
  p.tmem_cols = round_pow2_cols(p.n_sub * NT);
  const int64_t ctas = (int64_t)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * (CoutP / NT);
  pl.NT = NT;
  pl.big = big;
  pl.smem = (size_t)p.a_stages * p.a_bytes + (size_t)p.sw * p.w_stage_bytes + 8 * 26 + 4 * 128 + 16;
  pl.grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n), (unsigned)(CoutP / NT), (unsigned)splits);
  pl.ok = true;
  // ---- estimated cycles (constants fitted to tools/plan_sweep.py measurements)
  const double mma_cyc = (NT == 128 ? 64.0 : NT == 64 ? 48.0 : 45.5) * (tf32 ? 2.0 : 1.0);
  const int co_res = (big || pl.smem > 112 * 1024 || p.tmem_cols > 256) ? 1 : 2;
  const double n_cta = (double)ctas * splits;
  const double slots = (double)sg_num_sms() * co_res;
  const double waves = (double)(int64_t)((n_cta + slots - 1) / slots);
  const double sharing = n_cta > sg_num_sms() ? co_res : 1.0;   // co-resident CTAs share one tensor pipe
  // the generic kernel's issuing warp spends ~200 cycles per tap and ~100 per MMA; the specialised one keeps up
  const double mmas_per_tap = (double)p.n_sub * (p.kb_chunks / 2);
  double tap_cyc = mmas_per_tap * mma_cyc;
  if (!pl.spec) {
    const double issue = 200.0 + 100.0 * mmas_per_tap;
    if (issue > tap_cyc) tap_cyc = issue;
  }
  // a weight stage: its MMAs (shared with the co-resident CTA), >= ~500 cycles of issue + barrier round
  // (tools/bulk_bench.cu), and the load latency a shallow ring exposes (hidden in part by a co-resident CTA)
  double stage = (double)p.tps * tap_cyc * sharing;
  if (stage < kStageIssueCycles) stage = kStageIssueCycles;
  const double exposed = kLoadLatencyCycles / ((p.sw - 1) * sharing);
  if (stage < exposed) stage = exposed;
  double body = (double)p.kblocks_per_split * (27 / p.tps) * stage;
  if (p.a_stages == 1) body += (double)p.kblocks_per_split * kLoadLatencyCycles / sharing;   // single halo buffer
  const double per_cta_bytes = (double)p.kblocks_per_split * (27.0 * p.w_tap_bytes + (double)p.kb_chunks * p.chunk_tx_bytes);
  const double active = n_cta < slots ? n_cta : slots;
  const double stream = active * per_cta_bytes / kL2BytesPerCycle;
  if (stream > body) body = stream;
  // epilogue: ~330 cycles per 16 accumulator columns per MMA tile; overlapped by the co-resident CTA's MMAs
  const double epi = (double)p.n_sub * (NT / 16) * 330.0 * (sharing > 1.0 ? 0.5 : 1.0);
  const double ws_bytes = (double)N * D * H * W * CoutP * 4.0;
  pl.cost = waves * (body + kCtaFixedCycles + epi) +
            (splits > 1 ? kSplitFixedCycles + ws_bytes * (2.5 + splits) / kL2BytesPerCycle : 0.0);
  return pl;
}

// test / tuning hook: NT, big, td_max, splits (0 = choose by estimated cost)
int g_force_plan[4] = {0, 0, 0, 0};

Plan make_plan(int N, int Cin, int Cout, int D, int H, int W, bool tf32 = false) {
  if (g_force_plan[0] > 0)
    return make_plan_cfg(N, Cin, Cout, D, H, W, g_force_plan[0], g_force_plan[1] != 0, g_force_plan[2] > 0 ? g_force_plan[2] : 8,
                         g_force_plan[3] > 0 ? g_force_plan[3] : 1, tf32);
  // the search walks ~1000 candidates: remember the winner per shape (launch-time cost matters in eager mode)
  static std::mutex mu;
  static std::map<std::array<int, 7>, Plan> cache;
  const std::array<int, 7> key = {N, Cin, Cout, D, H, W, tf32 ? 1 : 0};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  Plan best;
  const int CCin = sg_chunks(Cin) * (tf32 ? 2 : 1);
  static const int nts[4] = {128, 64, 32, 16};
  for (int ni = 0; ni < 4; ++ni)
    for (int big = 0; big < 2; ++big)
      for (int td_max = 8; td_max >= 1; td_max /= 2)
        for (int splits = 1; splits <= CCin / 2; ++splits) {
          if ((CCin / 2) % splits && (CCin / 4 == 0 || (CCin / 4) % splits)) continue;
          Plan c = make_plan_cfg(N, Cin, Cout, D, H, W, nts[ni], big != 0, td_max, splits, tf32);
          if (c.ok && (!best.ok || c.cost < best.cost)) best = c;
        }
  {
    std::lock_guard<std::mutex> lock(mu);
    cache[key] = best;
  }
  return best;
}

template <int NT, int TD, int TH, int TN, bool TF32 = false>
int launch(const Plan& pl, const CUtensorMap& map, const CUtensorMap& map1, const CUtensorMap& wmap, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc<NT, TD, TH, TN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    if (e != cudaSuccess) {
      sg_set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  sg_launch((k_conv_tc<NT, TD, TH, TN, TF32>), pl.grid, kThreads, pl.smem, s, map, map1, wmap, pl.p);
  return sg_check_launch(TF32 ? "sg_conv3d_fprop(tcgen05 tf32)" : "sg_conv3d_fprop(tcgen05)");
}

template <int NT, bool TF32 = false>
int launch_spec(const Plan& pl, const CUtensorMap& map, const CUtensorMap& map1, const CUtensorMap& wmap, cudaStream_t s) {
  const TcParams& p = pl.p;
  if (p.th == 16 && p.tn == 1) {
    if (p.td == 1) return launch<NT, 1, 16, 1, TF32>(pl, map, map1, wmap, s);
    if (p.td == 2) return launch<NT, 2, 16, 1, TF32>(pl, map, map1, wmap, s);
    if (p.td == 4) return launch<NT, 4, 16, 1, TF32>(pl, map, map1, wmap, s);
  } else if (p.th == 8 && p.td == 2) {
    if (p.tn == 1) return launch<NT, 2, 8, 1, TF32>(pl, map, map1, wmap, s);
    if constexpr (!TF32)
      if (p.tn == 2) return launch<NT, 2, 8, 2, false>(pl, map, map1, wmap, s);
  }
  sg_set_error("conv_tc: no specialised kernel for tile %dx%dx%d", p.tn, p.td, p.th);
  return -4;
}

}  // namespace

#include "conv_tc_res.cuh"
#include "conv_tc_roll.cuh"

namespace {
// tests: 0 = auto, 1 = always the streaming kernel, 2 = the weight-resident kernel whenever the
// geometry allows (ignoring the "enough tiles to amortise the weight load" heuristic)
int g_force_streaming = 0;

int encode_halo_map(CUtensorMap* map, const void* x, int N, int CC, int D, int H, int W, int halo_w, int halo_h,
                    int halo_d, int tn, int box_c = 1) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  // activations as a 5-D tensor [W*8 | H | D | CC | N] of bf16; box = one 8-channel halo block
  cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)CC, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                           (cuuint64_t)CC * D * H * W * 16};
  cuuint32_t box[5] = {(cuuint32_t)halo_w * 8, (cuuint32_t)halo_h, (cuuint32_t)halo_d, (cuuint32_t)box_c, (cuuint32_t)tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  return 0;
}
}  // namespace

extern "C" void sg_tc_force_streaming(int on) { g_force_streaming = on; }
extern "C" void sg_tc_res_zs_mode(int mode) { g_res_zs_mode = mode; }
extern "C" void sg_tc_res_force(int td, int kb_chunks) { g_res_force_td = td; g_res_force_kb = kb_chunks; }
extern "C" void sg_tc_force_plan(int nt, int big, int td_max, int splits) {
  g_force_plan[0] = nt; g_force_plan[1] = big; g_force_plan[2] = td_max; g_force_plan[3] = splits;
}

int64_t sg_tc_wgrad_workspace_bytes(int N, int Cin, int Cout, int D, int H, int W, int tf32);

extern "C" int sg_conv3d_pixelnorm_supported(int N, int Cin, int Cout, int D, int H, int W);
extern "C" int sg_conv3d_pool_supported(int N, int Cin, int Cout, int D, int H, int W);
int sg_conv_finish_f32(const float* acc, const float* bias, const void* mask_src, void* y, int N,
                       int Cout, int64_t V, float scale, int lrelu, int slices, cudaStream_t s);

int64_t sg_tc_workspace_bytes(int kind, int N, int Cin, int Cout, int D, int H, int W, int tf32) {
  if (kind == 1) return sg_tc_wgrad_workspace_bytes(N, Cin, Cout, D, H, W, tf32);
  if (kind != 0) return 0;
  if (!tf32 && g_force_streaming != 1 && make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2).ok) return 0;
  Plan pl = make_plan(N, Cin, Cout, D, H, W, tf32 != 0);
  if (!pl.ok || pl.p.splits == 1) return 0;
  return (int64_t)pl.p.splits * N * D * H * W * pl.p.CoutP * (int64_t)sizeof(float);
}

namespace {
// fp32 activations [N][CC8][D][H][W][8] as the 5-D tensor [4 | W | H | D | N*CC8] (voxel pitch 32 bytes) starting at the
// `half`-th 16 bytes of the voxels: box = one 4-channel halo block, which lands as [d][h][w][16 bytes]
int encode_f32_half_map(CUtensorMap* map, const void* x, int half, int N, int CC8, int D, int H, int W, int box_w,
                        int box_h, int box_d, int box_c) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  cuuint64_t dims[5] = {4, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N * CC8};
  cuuint64_t strides[4] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32, (cuuint64_t)D * H * W * 32};
  cuuint32_t box[5] = {4, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_d, (cuuint32_t)box_c};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)((const char*)x + 16 * half), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled (fp32 half map) failed (%d)", (int)r);
    return -3;
  }
  return 0;
}
}  // namespace
int sg_encode_f32_half_map(CUtensorMap* map, const void* x, int half, int N, int CC8, int D, int H, int W, int box_w,
                           int box_h, int box_d, int box_c) {
  return encode_f32_half_map(map, x, half, N, CC8, D, H, W, box_w, box_h, box_d, box_c);
}

// 1 when the fused conv + pixel-norm epilogue covers the shape: bf16, the weight-resident kernel with ONE N tile
extern "C" int sg_conv3d_pixelnorm_supported(int N, int Cin, int Cout, int D, int H, int W) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || D <= 0 || H <= 0 || W <= 0 || g_force_streaming == 1) return 0;
  ResPlan rp = make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2);
  return rp.ok && rp.NT == rp.p.CoutP ? 1 : 0;
}

// 1 when the fused conv + 2x2x2 pooling epilogue covers the shape: bf16, the weight-resident kernel with an even number
// of planes per tile
extern "C" int sg_conv3d_pool_supported(int N, int Cin, int Cout, int D, int H, int W) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || D <= 0 || H <= 0 || W <= 0 || g_force_streaming == 1 || D % 2) return 0;
  ResPlan rp = make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2);
  return rp.ok && rp.p.td % 2 == 0 ? 1 : 0;
}

// tf32 != 0: x, y, mask_src are fp32 acts, wp is the SG_TF32 packing; the streaming kernel with kind::tf32.
// pn_y != null: fused pixel-norm second output; pool_y != null: fused 2x2x2 pooling second output (bf16, resident kernel
// only: returns 1 otherwise).
int sg_tc_fprop(const void* x, const void* wp, const float* bias, const void* mask_src, void* y, int N, int Cin,
                int Cout, int D, int H, int W, float scale, int lrelu, void* ws, int64_t ws_bytes, cudaStream_t s,
                int tf32, void* pn_y = nullptr, float pn_eps = 0.f, int pn_lrelu_after = 0, void* pool_y = nullptr,
                float pool_scale = 0.f) {
  if (pn_y != nullptr && (tf32 || !sg_conv3d_pixelnorm_supported(N, Cin, Cout, D, H, W) || mask_src != nullptr)) return 1;
  if (pool_y != nullptr && (tf32 || !sg_conv3d_pool_supported(N, Cin, Cout, D, H, W) || mask_src != nullptr || pn_y)) return 1;
  if (!tf32 && g_force_streaming != 1 && roll_mode()) {
    // the rolling form of the resident kernel (conv_tc_roll.cuh): columns through the depth, per-plane pipelining
    RollPlan rl = make_roll_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2);
    if (roll_wanted(rl) && (pn_y == nullptr || rl.NT == rl.p.CoutP)) {
      RollParams& q = rl.p;
      q.wp = (const __nv_bfloat16*)wp;
      q.bias = bias;
      q.mask = (const __nv_bfloat16*)mask_src;
      q.y = (__nv_bfloat16*)y;
      q.scale = scale;
      q.lrelu = lrelu;
      q.pn_y = (__nv_bfloat16*)pn_y;
      q.pn_eps = pn_eps;
      q.pn_inv_c = 1.f / (float)Cout;
      q.pn_lrelu_after = pn_lrelu_after;
      q.pool_y = (__nv_bfloat16*)pool_y;
      q.pool_scale = pool_scale;
      CUtensorMap rmap;
      int rc = encode_halo_map(&rmap, x, N, rl.CCin, D, H, W, 10, 18, 2, 1, rl.CCin);
      if (rc) return rc;
      switch (rl.NT) {
        case 16: return launch_roll<16>(rl, rmap, s);
        case 32: return launch_roll<32>(rl, rmap, s);
        default: return launch_roll<64>(rl, rmap, s);
      }
    }
  }
  if (!tf32 && g_force_streaming != 1) {
    ResPlan rp = make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2);
    if (rp.ok) {
      ResParams& q = rp.p;
      q.wp = (const __nv_bfloat16*)wp;
      q.bias = bias;
      q.mask = (const __nv_bfloat16*)mask_src;
      q.y = (__nv_bfloat16*)y;
      q.scale = scale;
      q.lrelu = lrelu;
      q.pn_y = (__nv_bfloat16*)pn_y;
      q.pn_eps = pn_eps;
      q.pn_inv_c = 1.f / (float)Cout;
      q.pn_lrelu_after = pn_lrelu_after;
      q.pool_y = (__nv_bfloat16*)pool_y;
      q.pool_scale = pool_scale;
      CUtensorMap rmap;
      int rc = encode_halo_map(&rmap, x, N, q.CCin, D, H, W, q.halo_w, q.halo_h, q.halo_d, 1);
      if (rc) return rc;
      switch (rp.NT) {
        case 16: return launch_res<16>(rp, rmap, s);
        case 32: return launch_res<32>(rp, rmap, s);
        default: return launch_res<64>(rp, rmap, s);
      }
    }
  }
  Plan pl = make_plan(N, Cin, Cout, D, H, W, tf32 != 0);
  if (!pl.ok) return 1;
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  TcParams& p = pl.p;
  p.wp = wp;
  p.bias = bias;
  p.mask = mask_src;
  p.y = y;
  p.ws = (float*)ws;
  p.scale = scale;
  p.lrelu = lrelu;
  if (p.splits > 1) {
    int64_t need = (int64_t)p.splits * N * D * H * W * p.CoutP * (int64_t)sizeof(float);
    SG_REQUIRE(ws != nullptr && ws_bytes >= need, "sg_conv3d_fprop(tcgen05): workspace too small (%lld < %lld)",
               (long long)ws_bytes, (long long)need);
  }
  CUtensorMap map, map1;
  CUresult r;
  if (tf32) {
    int rc = encode_f32_half_map(&map, x, 0, N, p.CCin / 2, D, H, W, p.halo_w, p.halo_h, p.halo_d, 1);
    if (!rc) rc = encode_f32_half_map(&map1, x, 1, N, p.CCin / 2, D, H, W, p.halo_w, p.halo_h, p.halo_d, 1);
    if (rc) return rc;
  } else {
    // activations as a 5-D tensor [W*8 | H | D | CC | N] of bf16; box = one 8-channel halo block
    cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)p.CCin, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                             (cuuint64_t)p.CCin * D * H * W * 16};
    cuuint32_t box[5] = {(cuuint32_t)p.halo_w * 8, (cuuint32_t)p.halo_h, (cuuint32_t)p.halo_d, 1, (cuuint32_t)p.tn};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      sg_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
      return -3;
    }
    map1 = map;
  }
  // packed weights [27][CCin][CoutP][16 B] as a 4-D tensor [128 B = 8 co x 16 B | CoutP/8 | CCin | 27]; box = one stage
  CUtensorMap wmap;
  {
    const cuuint64_t inner = tf32 ? 32 : 64;   // elements per 128 bytes
    cuuint64_t wd[4] = {inner, (cuuint64_t)p.CoutP / 8, (cuuint64_t)p.CCin, 27};
    cuuint64_t wst[3] = {128, (cuuint64_t)p.CoutP * 16, (cuuint64_t)p.CCin * p.CoutP * 16};
    cuuint32_t wbox[4] = {(cuuint32_t)inner, (cuuint32_t)pl.NT / 8, (cuuint32_t)p.kb_chunks, (cuuint32_t)p.tps};
    cuuint32_t we[4] = {1, 1, 1, 1};
    r = encode(&wmap, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(wp), wd,
               wst, wbox, we, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      sg_set_error("conv_tc: cuTensorMapEncodeTiled (weights) failed (%d)", (int)r);
      return -3;
    }
  }
  int rc;
  if (tf32) {
    if (pl.spec) {
      rc = pl.NT == 128 ? launch_spec<128, true>(pl, map, map1, wmap, s) : launch_spec<64, true>(pl, map, map1, wmap, s);
    } else {
      switch (pl.NT) {
        case 16: rc = launch<16, 0, 0, 0, true>(pl, map, map1, wmap, s); break;
        case 32: rc = launch<32, 0, 0, 0, true>(pl, map, map1, wmap, s); break;
        case 64: rc = launch<64, 0, 0, 0, true>(pl, map, map1, wmap, s); break;
        default: rc = launch<128, 0, 0, 0, true>(pl, map, map1, wmap, s); break;
      }
    }
  } else if (pl.spec) {
    rc = pl.NT == 128 ? launch_spec<128>(pl, map, map1, wmap, s) : launch_spec<64>(pl, map, map1, wmap, s);
  } else {
    switch (pl.NT) {
      case 16: rc = launch<16, 0, 0, 0>(pl, map, map1, wmap, s); break;
      case 32: rc = launch<32, 0, 0, 0>(pl, map, map1, wmap, s); break;
      case 64: rc = launch<64, 0, 0, 0>(pl, map, map1, wmap, s); break;
      default: rc = launch<128, 0, 0, 0>(pl, map, map1, wmap, s); break;
    }
  }
  if (rc) return rc;
  if (p.splits > 1)
    return tf32 ? sg_conv_finish_f32((const float*)ws, bias, mask_src, y, N, Cout, (int64_t)D * H * W, scale, lrelu, p.splits, s)
                : sg_conv_finish_bf16((const float*)ws, bias, mask_src, y, N, Cout, (int64_t)D * H * W, scale, lrelu, p.splits, s);
  return 0;
}

// 1 when sg_conv3d_fprop(..., impl = SG_IMPL_TF32) covers the shape (fp32 activations, kind::tf32)
extern "C" int sg_conv3d_tf32_supported(int N, int Cin, int Cout, int D, int H, int W) {
  if (N <= 0 || Cin <= 0 || Cout <= 0 || D <= 0 || H <= 0 || W <= 0) return 0;
  return make_plan(N, Cin, Cout, D, H, W, true).ok ? 1 : 0;
}

#ifdef SG_RES_TIMING
// diagnostic build only (tools/res_timing.py): the per-CTA wait counters of the last resident-kernel launch
extern "C" int sg_tc_res_timing(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_res_timing, sizeof(long long) * 148 * 8);
}
#endif

// Introspection for tests / DESIGN.md: the tiling the tcgen05 path would use for a shape.
// out[0..15] = ok, NT, tn, td, th, n_sub, kb_chunks, sw, splits, kblocks_per_split, grid.x, grid.y,
//              grid.z, smem bytes, tmem columns, a_bytes
extern "C" int sg_tc_plan_debug(int N, int Cin, int Cout, int D, int H, int W, int* out) {
  Plan pl = make_plan(N, Cin, Cout, D, H, W);
  const TcParams& p = pl.p;
  int v[16] = {pl.ok, pl.NT, p.tn, p.td, p.th, p.n_sub, p.kb_chunks + 100 * p.tps + (pl.spec ? 10000 : 0), p.sw, p.splits, p.kblocks_per_split,
               (int)pl.grid.x, (int)pl.grid.y, (int)pl.grid.z, (int)pl.smem, p.tmem_cols, p.a_bytes};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return 0;
}
