// Persistent, weight-resident variant of the tcgen05 implicit-GEMM convolution (included by
// conv_tc.cu).  For the high-resolution, small-channel layers that hold ~85 % of the step's
// FLOPs the whole packed weight slice of a CTA's N tile (27 taps x Cin x NT <= ~110 KB) fits in
// shared memory next to a double-buffered halo tile:
//   * one CTA per SM, weights loaded ONCE per launch (the streaming kernel re-fetched them per
//     tile: 60 % of its L2->SMEM bytes, and its MMA issuer starved on the 1 KB bulk copies);
//   * static tile schedule tile = blockIdx.x + i*gridDim.x; the producer runs ahead across
//     tile boundaries through a ring of halo stages;
//   * two TMEM accumulator sets: the epilogue of tile i overlaps the MMAs of tile i+1.
// Geometry, operand descriptors and epilogue are those of k_conv_tc.
//
// ZS ("z-stacked") form.  An M = 128, K = 16 MMA costs max(N / 2, ~46) cycles (tools/mma_bench.cu; the floor is per
// instruction -- the A-operand collector does not lift it, tools/mma_collector_bench.cu), so an MMA with N = 32 or 64
// runs the tensor pipe at 35 % / 67 %.  Instead of y[p] += x[p + kd - 1] * W[kd] (three MMAs with three different A
// operands per output plane), the ZS loop walks the INPUT planes q of the halo tile and issues ONE MMA per plane with
// the kd taps stacked along N in DESCENDING order, B = [W[kd=2]; W[kd=1]; W[kd=0]] (3*NT rows), whose accumulator is
// the block of THREE ADJACENT output-plane accumulators starting at plane q - 1:
//     columns [acc[q-1] | acc[q] | acc[q+1]]  +=  x[q] * [W[2] | W[1] | W[0]]
// -- every term lands where it belongs, the tensor core's own accumulate does the sum over kd, and TMEM holds TD * NT
// columns per set like the plain form.  The two planes at either end of the halo use the sub-ranges of B that fall
// inside the tile (N = NT, 2*NT).  Per (kh, kw, K step) that is TD + 2 MMAs of N up to 3*NT instead of 3 * TD of N = NT:
// NT = 64, TD = 4: 416 cycles for 384 cycles of work (92 % against 67 %); NT = 32, TD = 4: 64 % against 35 %.
// One instruction has one accumulate flag for all its columns, so the very first (kh, kw, K step) of a tile is issued
// tap by tap (each accumulator is first touched by its kd = 0 tap, with the flag off); everything after it is stacked.
// (Round 1 stacked kd into SEPARATE accumulators Z[q][kd] and summed three of them per output in the epilogue: three
// times the TMEM columns, which capped TD at 2 for NT = 32.)
#pragma once

namespace {

constexpr int kResStages = 4;   // upper bound of halo stages

struct ResParams {
  const __nv_bfloat16* wp;
  const float* bias;
  const __nv_bfloat16* mask;
  __nv_bfloat16* y;
  int N, D, H, W;
  int CCin, Cout, CoutP, CCout;
  int td, th;
  int tiles_w, tiles_h, tiles_d, n_tiles;
  int halo_w, halo_h, halo_d;
  int chunk_bytes, chunk_tx_bytes;
  int n_sub;
  int sub_line[kMaxSub];
  int kb_chunks, n_kblocks;
  int stages, stage_bytes;
  int w_bytes;               // resident weights: 27 * CCin * NT * 16
  int tmem_cols;
  float scale;
  int lrelu;
  // fused ChannelNormalization (network.py:192-197) of the generator blocks: when pn_y is set the CTA's N tile holds
  // ALL output channels of a voxel, the epilogue thread of that voxel normalises them and writes y (the conv / LeakyReLU
  // output the backward needs) AND pn_y = [lrelu](y * rsqrt(mean_c y^2 + eps))
  __nv_bfloat16* pn_y;
  float pn_eps, pn_inv_c;
  int pn_lrelu_after;
  // fused AvgPool3d(2) of the discriminator blocks (network.py:88-90 conv2 -> lrelu -> avg-pool): when pool_y is set the
  // epilogue also writes pool_y[N][CCout][D/2][H/2][W/2][8] = pool_scale * (2x2x2 block sums of y).  A CTA tile holds
  // whole blocks (TD even, 16 lines, 8 voxels): the two planes are the same thread's, the h / w neighbours sit 8 / 1
  // lanes away (two shuffles per value).  y itself is still written: its sign is the mask of the backward pass.
  __nv_bfloat16* pool_y;
  float pool_scale;
};

// NT = output-channel tile, TD = depth planes per tile (= MMA tiles per accumulator set),
// KBC = 8-channel chunks per K block.  With th = 16 and tw = 8 fixed, every descriptor offset of
// the 27 x KBC/2 x TD MMAs of a K block is a compile-time constant: the fully unrolled issue
// loop costs one uniform-datapath add per MMA.
// The MMAs of one K block of the z-stacked form.  B rows [kd = 2, 1, 0][co] of a (kh, kw, chunk): chunk stride (LBO) =
// 3*NT rows.  PLANE-OUTER order: the 9 * KBC/2 MMAs of an input plane go to the same accumulator columns back to back.
// FIRST (the tile's first K block): the first MMA of every plane is issued in two pieces, the kd = 0 tap (first touch
// of output plane q + 1) with the accumulate flag off -- one instruction has one flag for all its columns.  Every
// descriptor offset is a compile-time constant or a running uniform add.
template <int NT, int TD, int KBC, bool FIRST>
__device__ __forceinline__ void zs_issue_kblock(uint32_t d_tmem, uint64_t a_stage, uint64_t b_kb, uint32_t wz_khw16,
                                                uint32_t leader) {
  constexpr int HALO_H = 18, HALO_W = 10;
  constexpr int CHUNK_BYTES = ((TD + 2) * HALO_H * HALO_W * 16 + 127) / 128 * 128;
  constexpr uint32_t KK_A = (2u * CHUNK_BYTES) >> 4;
  constexpr uint32_t PLANE = HALO_H * HALO_W;
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
#pragma unroll
  for (int j = 0; j < TD + 2; ++j) {
    // input plane q = j - 1 feeds output planes q + 1 - kd, kd in [kd_lo, kd_hi]
    const int q = j - 1;
    const int kd_hi = q + 1 < 2 ? q + 1 : 2, kd_lo = q + 2 - TD > 0 ? q + 2 - TD : 0;
    const uint32_t idz = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(((kd_hi - kd_lo + 1) * NT) >> 3) << 17) |
                         ((128u >> 4) << 24);
    const uint64_t a_plane = a_stage + (uint64_t)(j * PLANE);
    uint64_t b_run = b_kb + (uint64_t)((2 - kd_hi) * NT);
#pragma unroll
    for (int khw = 0; khw < 9; ++khw) {
      const uint32_t a_off = (uint32_t)((khw / 3) * HALO_W + khw % 3);
#pragma unroll
      for (int kk = 0; kk < KBC / 2; ++kk) {
        const uint64_t a = a_plane + (uint64_t)(a_off + kk * KK_A);
        if (FIRST && (khw | kk) == 0 && kd_lo == 0) {
          // first touch of output plane q + 1 (tap kd = 0): overwrite; the taps kd_hi..1 before it in the same B rows
          // accumulate into planes that earlier input planes have already touched -- one MMA for both of them
          if (kd_hi >= 1) {
            const uint32_t idacc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((kd_hi * NT) >> 3) << 17) | ((128u >> 4) << 24);
            tc_mma(d_tmem + (q + 1 - kd_hi) * NT, a, b_run, idacc, 1u, leader);
          }
          tc_mma(d_tmem + (q + 1) * NT, a, b_run + (uint64_t)(kd_hi * NT), idesc, 0u, leader);
        } else {
          tc_mma(d_tmem + (q + 1 - kd_hi) * NT, a, b_run + (uint64_t)(kk * 2 * 3 * NT), idz, 1u, leader);
        }
      }
      b_run += wz_khw16;
    }
  }
}

// -DSG_RES_TIMING (tools/res_timing.py, never in the release build): per-CTA clock64 stamps of where the three roles wait.
//   [0] issuer: total, [1] issuer: waiting for A_FULL, [2] issuer: waiting for ACC_EMPTY, [3] producer: waiting for
//   A_EMPTY, [4] epilogue warp 2: total, [5] epilogue: waiting for ACC_FULL, [6] tiles, [7] producer total
#ifdef SG_RES_TIMING
__device__ long long g_res_timing[148 * 8];
#define RT(...) __VA_ARGS__
#else
#define RT(...)
#endif

template <int NT, int TD, int KBC, bool ZS>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc_res(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ ResParams p) {
  constexpr int HALO_H = 18, HALO_W = 10;
  constexpr int CHUNK_BYTES = ((TD + 2) * HALO_H * HALO_W * 16 + 127) / 128 * 128;
  constexpr uint32_t KK_A = (2u * CHUNK_BYTES) >> 4;
  constexpr uint32_t PLANE = HALO_H * HALO_W;   // voxels (16-byte units) per halo plane
  extern __shared__ __align__(128) uint8_t smem[];
  sg_pdl_trigger();
  // carve-up: [weights][halo stages][slack for garbage-row reads][barriers][tmem slot]
  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + p.stages * p.stage_bytes + 4096);
  // bars: [0] w_full, [1..1+S) a_full, [1+S..1+2S) a_empty, then acc_full[2], acc_empty[2]
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int W_FULL = 0, A_FULL = 1, A_EMPTY = 1 + p.stages, ACC_FULL = 1 + 2 * p.stages, ACC_EMPTY = ACC_FULL + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kResStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * NT;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    mbar_init(BAR(W_FULL), 1);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(BAR(A_FULL + i), 1);
      mbar_init(BAR(A_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(ACC_FULL + i), 1);
      mbar_init(BAR(ACC_EMPTY + i), 4);   // one arrival per epilogue warp
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
  const int acc_cols = p.n_sub * NT;                     // columns of one accumulator set
  float* s_bias = reinterpret_cast<float*>(bars + 16);   // NT floats, 16-byte aligned, after the barriers

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      // resident weights: 27*CCin pieces of NT rows x 16 B, contiguous in the packed tensor
      const uint32_t w_addr = smem_u32(w_smem);
      mbar_expect_tx(BAR(W_FULL), (uint32_t)p.w_bytes);
      for (int i = 0; i < 27 * p.CCin; ++i) {
        uint32_t dst = w_addr + i * NT * 16;
        if (ZS) {   // [kh][kw][chunk][kd = 2, 1, 0][co][8]: the three kd taps of a (kh, kw, chunk) are consecutive N rows
          const int tap = i / p.CCin, chunk = i - tap * p.CCin;
          dst = w_addr + ((((tap % 9) * p.CCin + chunk) * 3 + (2 - tap / 9)) * NT) * 16;
        }
        bulk_load(dst, p.wp + ((int64_t)i * p.CoutP + co0) * 8, NT * 16u, BAR(W_FULL));
      }
      const uint32_t a_addr = smem_u32(a_smem);
      const uint32_t a_tx = (uint32_t)p.kb_chunks * (uint32_t)p.chunk_tx_bytes;
      int it = 0;
      RT(long long rt_w = 0; const long long rt_p0 = clock64();)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int t = tile;
        const int tile_w = t % p.tiles_w; t /= p.tiles_w;
        const int tile_h = t % p.tiles_h; t /= p.tiles_h;
        const int tile_d = t % p.tiles_d; t /= p.tiles_d;
        const int n = t;
        const int w0 = tile_w * 8, h0 = tile_h * p.th, d0 = tile_d * p.td;
        for (int kb = 0; kb < p.n_kblocks; ++kb, ++it) {
          const int s = it % p.stages;
          RT(const long long rt_t = clock64();)
          mbar_wait(BAR(A_EMPTY + s), ((it / p.stages) & 1) ^ 1);
          RT(rt_w += clock64() - rt_t;)
          mbar_expect_tx(BAR(A_FULL + s), a_tx);
          for (int c = 0; c < p.kb_chunks; ++c)
            tma_load_5d(a_addr + s * p.stage_bytes + c * p.chunk_bytes, &xmap, BAR(A_FULL + s), (w0 - 1) * 8, h0 - 1,
                        d0 - 1, kb * p.kb_chunks + c, n);
        }
      }
      RT(g_res_timing[blockIdx.x * 8 + 3] = rt_w; g_res_timing[blockIdx.x * 8 + 7] = clock64() - rt_p0;)
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t leader = elect_one();   // all lanes run the loops; one issues
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t a_desc0 = make_desc(smem_u32(a_smem), (uint32_t)CHUNK_BYTES, HALO_W * 16u);
      const uint64_t w_desc0 = make_desc(smem_u32(w_smem), NT * 16u, 128u);
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
      const int stages = p.stages, n_kblocks = p.n_kblocks;
      const uint32_t w_tap16 = (uint32_t)(p.CCin * NT);
      const uint64_t wz_desc0 = make_desc(smem_u32(w_smem), 3 * NT * 16u, 128u);
      const uint32_t wz_khw16 = (uint32_t)(p.CCin * 3 * NT);
      mbar_wait(BAR(W_FULL), 0);
      int s = 0, ph = 0, ti = 0;
      RT(long long rt_full = 0, rt_acc = 0; const long long rt_i0 = clock64();)
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
        const int buf = ti & 1;
        RT(const long long rt_t = clock64();)
        mbar_wait(BAR(ACC_EMPTY + buf), ((ti >> 1) & 1) ^ 1);   // epilogue has drained this set
        RT(rt_acc += clock64() - rt_t;)
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * acc_cols;
        for (int kb = 0; kb < n_kblocks; ++kb) {
          RT(const long long rt_u = clock64();)
          mbar_wait(BAR(A_FULL + s), ph);
          RT(rt_full += clock64() - rt_u;)
          tc_fence_after();
          const uint64_t a_stage = a_desc0 + (uint64_t)(s * stage16);
          const uint32_t acc0 = kb != 0;
          if constexpr (ZS) {
            const uint64_t b_kb = wz_desc0 + (uint64_t)(kb * (KBC * 3 * NT));
            if (kb == 0) zs_issue_kblock<NT, TD, KBC, true>(d_tmem, a_stage, b_kb, wz_khw16, leader);
            else zs_issue_kblock<NT, TD, KBC, false>(d_tmem, a_stage, b_kb, wz_khw16, leader);
          } else {
          uint64_t b_tap = w_desc0 + (uint64_t)(kb * (KBC * NT));
#pragma unroll
          for (int tap = 0; tap < 27; ++tap) {
            // output voxel (d,h,w) of MMA tile `sub` reads halo voxel (sub+kd, h+kh, w+kw)
            const uint32_t a_off = (uint32_t)(((tap / 9) * HALO_H + (tap / 3) % 3) * HALO_W + tap % 3);
#pragma unroll
            for (int kk = 0; kk < KBC / 2; ++kk)
#pragma unroll
              for (int sub = 0; sub < TD; ++sub)
                tc_mma(d_tmem + sub * NT, a_stage + (uint64_t)(a_off + sub * PLANE + kk * KK_A),
                       b_tap + (uint64_t)(kk * 2 * NT), idesc, (tap | kk) ? 1u : acc0, leader);
            b_tap += w_tap16;
          }
          }
          tc_commit(BAR(A_EMPTY + s), leader);
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        tc_commit(BAR(ACC_FULL + buf), leader);
      }
      RT(if (lane == 0) {
        g_res_timing[blockIdx.x * 8 + 0] = clock64() - rt_i0;
        g_res_timing[blockIdx.x * 8 + 1] = rt_full;
        g_res_timing[blockIdx.x * 8 + 2] = rt_acc;
        g_res_timing[blockIdx.x * 8 + 6] = ti;
      })
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int64_t V = (int64_t)p.D * p.H * p.W;
    const float scale = p.scale;
    const int lrelu = p.lrelu;
    const __nv_bfloat16* mask = p.mask;
    __nv_bfloat16* yout = p.y;
    for (int i = row; i < NT; i += 128) s_bias[i] = (p.bias && co0 + i < p.Cout) ? p.bias[co0 + i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
    int ti = 0;
    RT(long long rt_e = 0; const long long rt_e0 = clock64();)
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++ti) {
      int t = tile;
      const int tile_w = t % p.tiles_w; t /= p.tiles_w;
      const int tile_h = t % p.tiles_h; t /= p.tiles_h;
      const int tile_d = t % p.tiles_d; t /= p.tiles_d;
      const int n = t;
      const int w0 = tile_w * 8, h0 = tile_h * p.th, d0 = tile_d * p.td;
      const int buf = ti & 1;
      // resident geometry: tiles divide the volume exactly, every accumulator row is a real voxel
      const int64_t plane8 = (int64_t)p.H * p.W * 8;
      const int64_t obase0 =
          (((int64_t)n * p.CCout + co0 / 8) * V + ((int64_t)d0 * p.H + h0 + (row >> 3)) * p.W + w0 + (row & 7)) * 8;
      // The LeakyReLU-mask vectors are prefetched before the accumulators are waited for: those of the whole tile when
      // they fit 16 registers x 4 (TD * NT <= 128), else those of one plane at a time (a whole 4 x 64 tile is 128
      // registers: the NT = 64 kernel with a mask ran 227 us against 171 us without).
      constexpr bool PREFETCH_TILE = TD * NT <= 128;
      constexpr int MK_SUB = PREFETCH_TILE ? TD : 1;
      uint4 mk[MK_SUB][NT / 8];
      if (mask) {
#pragma unroll
        for (int sub = 0; sub < MK_SUB; ++sub)
#pragma unroll
          for (int c = 0; c < NT / 8; ++c)
            mk[sub][c] = __ldg(reinterpret_cast<const uint4*>(mask + obase0 + sub * plane8 + (int64_t)c * V * 8));
        if constexpr (!PREFETCH_TILE) {
          // the other planes' masks are loaded one plane ahead while the epilogue drains the tile in a burst (~0.25 us
          // per plane, a fraction of the DRAM latency): pull them into L2 now, while the tile's MMAs still run -- one
          // 128-byte line per 8 lanes (32 -> 32 with a mask: 141 -> 129 us; 115 us without a mask).  Prefetching the
          // NEXT tile's masks as well measured worse (152 us).
          if ((lane & 7) == 0) {
#pragma unroll
            for (int sub = 1; sub < TD; ++sub)
#pragma unroll
              for (int c = 0; c < NT / 8; ++c)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(mask + obase0 + sub * plane8 + (int64_t)c * V * 8));
          }
        }
      }
      RT(const long long rt_t = clock64();)
      mbar_wait(BAR(ACC_FULL + buf), (ti >> 1) & 1);
      RT(rt_e += clock64() - rt_t;)
      tc_fence_after();
      if (p.pn_y != nullptr) {
        // conv -> [lrelu] -> pixel-norm [-> lrelu]: this thread holds all NT channels of its voxel
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols);
#pragma unroll 1
        for (int sub = 0; sub < TD; ++sub) {
          float t[NT];
          __syncwarp();
          tmem_ld_block<NT>(trow + (uint32_t)(sub * NT), t);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < NT; ++i) {
            float a = fmaf(t[i], scale, s_bias[i]);
            if (lrelu) a = lrelu02(a);
            t[i] = a;
            ss = fmaf(a, a, ss);
          }
          const float r = rsqrtf(ss * p.pn_inv_c + p.pn_eps);
          __nv_bfloat16* y0 = yout + obase0 + sub * plane8;
          __nv_bfloat16* y1 = p.pn_y + obase0 + sub * plane8;
#pragma unroll
          for (int c = 0; c < NT / 8; ++c) {
            F8 a, b;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a.v[j] = t[8 * c + j];
              const float u = t[8 * c + j] * r;
              b.v[j] = p.pn_lrelu_after ? lrelu02(u) : u;
            }
            st8(y0 + (int64_t)c * V * 8, a);
            st8(y1 + (int64_t)c * V * 8, b);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
        continue;
      }
      if constexpr (TD % 2 == 0) {
        if (p.pool_y != nullptr) {
          const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols);
          const int Hp = p.H >> 1, Wp = p.W >> 1;
          const int64_t Vp = (int64_t)(p.D >> 1) * Hp * Wp;
          const bool writer = (lane & 9) == 0;                       // even w (lane bit 0) and even h (lane bit 3)
          const int64_t pbase0 = (((int64_t)n * p.CCout + co0 / 8) * Vp +
                                  ((int64_t)(d0 >> 1) * Hp + ((h0 + (row >> 3)) >> 1)) * Wp + ((w0 + (row & 7)) >> 1)) * 8;
#pragma unroll 1
          for (int sub = 0; sub < TD; sub += 2) {
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 16) {
              float v0[16], v1[16];
              __syncwarp();
              tmem_ld16(trow + (uint32_t)(sub * NT + c0), v0);
              tmem_ld16(trow + (uint32_t)((sub + 1) * NT + c0), v1);
              float sum[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a = fmaf(v0[i], scale, s_bias[c0 + i]), b = fmaf(v1[i], scale, s_bias[c0 + i]);
                if (lrelu) {
                  a = lrelu02(a);
                  b = lrelu02(b);
                }
                v0[i] = a;
                v1[i] = b;
                float t = a + b;
                t += __shfl_xor_sync(0xffffffffu, t, 1);
                t += __shfl_xor_sync(0xffffffffu, t, 8);
                sum[i] = t * p.pool_scale;
              }
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                F8 o0, o1, op;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  o0.v[j] = v0[half * 8 + j];
                  o1.v[j] = v1[half * 8 + j];
                  op.v[j] = sum[half * 8 + j];
                }
                const int64_t ch = (int64_t)(c0 / 8 + half);
                st8(yout + obase0 + sub * plane8 + ch * V * 8, o0);
                st8(yout + obase0 + (sub + 1) * plane8 + ch * V * 8, o1);
                if (writer) st8(p.pool_y + pbase0 + (int64_t)(sub >> 1) * Hp * Wp * 8 + ch * Vp * 8, op);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
          continue;
        }
      }
#pragma unroll
      for (int sub = 0; sub < TD; ++sub) {
        uint4 mn[NT / 8];      // !PREFETCH_TILE: the next plane's masks, in flight during this plane's epilogue
        if (!PREFETCH_TILE && mask && sub + 1 < TD) {
#pragma unroll
          for (int c = 0; c < NT / 8; ++c)
            mn[c] = __ldg(reinterpret_cast<const uint4*>(mask + obase0 + (sub + 1) * plane8 + (int64_t)c * V * 8));
        }
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 16) {
          float v[16];
          __syncwarp();
          const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols);
          tmem_ld16(trow + (uint32_t)(sub * NT + c0), v);
          const int ms = PREFETCH_TILE ? sub : 0;
          epilogue16_regmask(v, s_bias + c0, scale, lrelu, mask != nullptr, mk[ms][c0 / 8], mk[ms][c0 / 8 + 1],
                             yout + obase0 + sub * plane8 + (int64_t)(c0 / 8) * V * 8, V * 8);
        }
        if (!PREFETCH_TILE && mask && sub + 1 < TD) {
#pragma unroll
          for (int c = 0; c < NT / 8; ++c) mk[0][c] = mn[c];
        }
      }
      // this warp is done reading the accumulator set: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(ACC_EMPTY + buf));
    }
    RT(if (warp == 2 && lane == 0) {
      g_res_timing[blockIdx.x * 8 + 4] = clock64() - rt_e0;
      g_res_timing[blockIdx.x * 8 + 5] = rt_e;
    })
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

struct ResPlan {
  bool ok = false;
  int NT = 0;
  bool zs = false;   // kd taps stacked along N (NT <= 32)
  ResParams p{};
  size_t smem = 0;
  dim3 grid;
};

int g_res_zs_mode = 0;   // test hook: 0 = auto, 1 = never z-stack, 2 = z-stack whenever possible
int g_res_force_td = 0, g_res_force_kb = 0;   // tuning hook (sg_tc_res_force): planes per tile / chunks per K block, 0 = auto

// Applies when H >= 16, W % 8 == 0 and the weight slice of one N tile fits next to >= 2 halo stages.
ResPlan make_res_plan(int N, int Cin, int Cout, int D, int H, int W, bool for_test = false) {
  ResPlan pl;
  ResParams& p = pl.p;
  if (W % 8 != 0 || H % 16 != 0) return pl;
  const int CCin = sg_chunks(Cin), CoutP = 16 * ((Cout + 15) / 16);
  const int budget = 222 * 1024;
  // N tile: the largest of {64, 32, 16} dividing CoutP whose resident weights leave room for the halo ring
  int NT = 0;
  for (int cand : {64, 32, 16}) {
    if (CoutP % cand) continue;
    if (27 * CCin * cand * 16 <= 120 * 1024) { NT = cand; break; }
  }
  if (NT == 0) return pl;
  if (CoutP / NT > 1) return pl;           // a second N tile re-reads every halo tile: the streaming kernel wins
                                           // (64->64 @16x64x64: 69.5 us streamed, 79.9 us resident with two N tiles)
  p.w_bytes = 27 * CCin * NT * 16;
  p.th = 16;
  p.halo_w = 10; p.halo_h = 18;
  // (td, K block): two accumulator sets of td MMA tiles must fit TMEM, >= 2 halo stages must fit shared memory.
  // The z-stacked form (td >= 2) gains with every extra plane per tile ((td + 2) wide MMAs instead of 3 * td narrow
  // ones), so the largest td wins; a K block of 4 chunks halves the barrier rounds when it still leaves 2 stages.
  int best_td = 0, best_kb = 0;
  bool zs = false;
  for (int td = 8; td >= 1 && best_td == 0; td /= 2) {
    if (td > D || D % td) continue;
    if (2 * td * NT > 512) continue;
    if (g_res_force_td && td != g_res_force_td) continue;
    // measured (tools/res_sweep.py): 16 -> 32 @64x256x256 runs 236 us at td = 4 against 277 us at td = 8 (one K block
    // of two chunks per tile: the epilogue of 8 planes outlasts the 90 MMAs that produce them)
    if (!g_res_force_td && td == 8 && NT == 32 && CCin == 2 && D % 4 == 0) continue;
    const bool this_zs = td >= 2 && g_res_zs_mode != 1;
    if (!this_zs && td > 4) continue;                 // the plain form is instantiated up to td = 4
    // every CTA should walk >= 2 tiles (weight load amortised, both accumulator sets in use): with too few tiles at
    // this td try a smaller one before giving the layer to the streaming kernel (32 -> 32 @16x64x64, B = 2: 128 tiles
    // at td = 8, 512 at td = 2 -- the generic streaming kernel it used to fall to ran 28 us)
    if (!for_test && !g_res_force_td) {
      const int tiles = (W / 8) * (H / 16) * (D / td) * N;
      const int ctas_td = tiles < sg_num_sms() ? tiles : sg_num_sms();
      if (tiles < 2 * ctas_td) continue;
    }
    for (int kb : {4, 2}) {
      if (CCin % kb) continue;
      if (g_res_force_kb && kb != g_res_force_kb) continue;
      const int chunk = ((td + 2) * 18 * 10 * 16 + 127) / 128 * 128;
      if (p.w_bytes + 2 * kb * chunk + 4096 + 512 > budget) continue;
      best_td = td; best_kb = kb; zs = this_zs;
      break;
    }
  }
  if (best_td == 0) return pl;
  p.kb_chunks = best_kb;
  p.n_kblocks = CCin / p.kb_chunks;
  p.td = best_td;
  p.halo_d = p.td + 2;
  p.chunk_tx_bytes = p.halo_d * p.halo_h * p.halo_w * 16;
  p.chunk_bytes = (p.chunk_tx_bytes + 127) / 128 * 128;
  p.stage_bytes = p.kb_chunks * p.chunk_bytes;
  int stages = (budget - p.w_bytes - 4096 - 512) / p.stage_bytes;
  if (stages > kResStages) stages = kResStages;
  p.stages = stages;
  p.n_sub = p.td;
  for (int i = 0; i < p.td; ++i) p.sub_line[i] = (i + 1) * p.halo_h + 1;
  p.tiles_w = W / 8; p.tiles_h = H / 16; p.tiles_d = D / p.td;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_d * N;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.CCin = CCin; p.Cout = Cout; p.CoutP = CoutP; p.CCout = sg_chunks(Cout);
  p.tmem_cols = round_pow2_cols(2 * p.n_sub * NT);
  pl.zs = zs;
  const int n_ntiles = CoutP / NT;
  int ctas = sg_num_sms() / n_ntiles;
  if (ctas > p.n_tiles) ctas = p.n_tiles;
  if (ctas < 1) return pl;
  if (for_test) {
    ctas = p.n_tiles >= 3 ? p.n_tiles / 3 : 1;   // tests: every CTA walks ~3 tiles (both accumulator sets, ring wrap)
    if (ctas > sg_num_sms() / n_ntiles) ctas = sg_num_sms() / n_ntiles;
  } else if (p.n_tiles < 2 * ctas) {
    return pl;                                   // too few tiles to amortise the weight load: stream instead
  }
  pl.grid = dim3((unsigned)ctas, (unsigned)n_ntiles, 1);
  pl.smem = (size_t)p.w_bytes + (size_t)p.stages * p.stage_bytes + 4096 + 8 * 16 + 4 * 64 + 16;
  pl.NT = NT;
  pl.ok = true;
  return pl;
}

template <int NT, int TD, int KBC, bool ZS>
int launch_res_inst(const ResPlan& pl, const CUtensorMap& map, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(k_conv_tc_res<NT, TD, KBC, ZS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      sg_set_error("conv_tc_res: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  sg_launch((k_conv_tc_res<NT, TD, KBC, ZS>), pl.grid, kThreads, pl.smem, s, map, pl.p);
  return sg_check_launch("sg_conv3d_fprop(tcgen05 resident)");
}

template <int NT>
int launch_res(const ResPlan& pl, const CUtensorMap& map, cudaStream_t s) {
  const int td = pl.p.td, kbc = pl.p.kb_chunks;
  if (pl.zs) {
    if (kbc == 4) {
      if (td == 2) return launch_res_inst<NT, 2, 4, true>(pl, map, s);
      if (td == 4) return launch_res_inst<NT, 4, 4, true>(pl, map, s);
      if constexpr (NT <= 32)
        if (td == 8) return launch_res_inst<NT, 8, 4, true>(pl, map, s);
    } else {
      if (td == 2) return launch_res_inst<NT, 2, 2, true>(pl, map, s);
      if (td == 4) return launch_res_inst<NT, 4, 2, true>(pl, map, s);
      if constexpr (NT <= 32)
        if (td == 8) return launch_res_inst<NT, 8, 2, true>(pl, map, s);
    }
    sg_set_error("conv_tc_res: no z-stacked instantiation for td=%d kb_chunks=%d", td, kbc);
    return -4;
  }
  if (kbc == 4) {
    if (td == 1) return launch_res_inst<NT, 1, 4, false>(pl, map, s);
    if (td == 2) return launch_res_inst<NT, 2, 4, false>(pl, map, s);
    if (td == 4) return launch_res_inst<NT, 4, 4, false>(pl, map, s);
  } else {
    if (td == 1) return launch_res_inst<NT, 1, 2, false>(pl, map, s);
    if (td == 2) return launch_res_inst<NT, 2, 2, false>(pl, map, s);
    if (td == 4) return launch_res_inst<NT, 4, 2, false>(pl, map, s);
  }
  sg_set_error("conv_tc_res: no instantiation for td=%d kb_chunks=%d", td, kbc);
  return -4;
}

}  // namespace
