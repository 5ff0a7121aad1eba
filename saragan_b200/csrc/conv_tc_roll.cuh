// "Rolling" form of the weight-resident tcgen05 convolution (included by conv_tc.cu after conv_tc_res.cuh).
//
// k_conv_tc_res cuts the volume into tiles of TD <= 8 planes: every tile pays for its two halo planes with narrow MMAs
// (N = NT and 2*NT cost 46 / 48 cycles whatever they compute, tools/mma_tmem_a_bench.cu) and TMEM holds two accumulator
// sets of TD planes.  Here a work item is a COLUMN -- 16 lines x 8 voxels through a whole depth segment of SEG planes --
// and everything is pipelined per plane:
//   * the producer streams the input planes of the column one by one (ONE TMA box [80 | 18 | 1 | all chunks] per plane)
//     through a ring of plane buffers.  With the in-place kd stacking of conv_tc_res.cuh every input plane is used
//     exactly once -- multiplied by [W[kd=2]; W[kd=1]; W[kd=0]] into the three adjacent accumulators of the output planes
//     q-1, q, q+1 -- so nothing but the plane in flight has to stay in shared memory;
//   * the accumulators are a RING of R = 512 / NT output planes in tensor memory: output plane r is complete as soon as
//     input plane r+1 has been issued (one tcgen05.commit per plane), the epilogue drains it while the MMAs of the next
//     planes run, and its columns are handed back R planes later;
//   * input planes outside the volume are pure zero padding and are skipped altogether (a tile of the other kernel
//     multiplies them); only the two ends of a segment inside the volume see narrow MMAs.
// Per output plane that is 9 * Cin/16 MMAs of N = 3*NT (+ segment ends) against 9 * Cin/16 * (TD + 2) / TD with
// narrower ends: NT = 32, SEG = 16: 58 cycles per (kh, kw, K step, plane) against 65.5 at TD = 8.
#pragma once

namespace {

// non-blocking barrier test (try_wait may suspend the thread for a while when the phase is not complete)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

constexpr int kRollStagesMax = 8;    // plane buffers
constexpr int kRollAccMax = 16;      // accumulator ring slots

struct RollParams {
  const __nv_bfloat16* wp;
  const float* bias;
  const __nv_bfloat16* mask;
  __nv_bfloat16* y;
  int N, D, H, W;
  int Cout, CoutP, CCout;
  int seg;                   // output planes per work item (divides D, even)
  int n_seg, tiles_w, tiles_h, n_items;
  int stages, stages_log2;   // plane buffers in the ring (power of two)
  int acc_slots, acc_log2;   // accumulator ring slots R (power of two)
  int w_bytes;               // resident weights: 27 * CCIN * NT * 16
  int tmem_cols;
  float scale;
  int lrelu;
  __nv_bfloat16* pn_y;       // fused pixel-norm second output (see ResParams)
  float pn_eps, pn_inv_c;
  int pn_lrelu_after;
  __nv_bfloat16* pool_y;     // fused 2x2x2 average pooling second output (see ResParams)
  float pool_scale;
};

// The MMAs of one INTERIOR input plane (all three kd taps land inside the segment, the three accumulator slots are
// contiguous): every operand offset is a compile-time constant relative to (d1, a_plane, wz_desc0).  The issuing thread
// executes its scalar code in order with the MMA issue and the MMA queue is shallow, so whatever it computes per plane is
// exposed: the general path below costs ~600 cycles per plane, this one a handful of uniform adds.
template <int NT, int CCIN>
__device__ __forceinline__ void roll_issue_interior(uint32_t d1, uint64_t a_plane, uint64_t wz_desc0, uint32_t leader) {
  constexpr int HALO_W = 10;
  constexpr uint32_t KK_A = (2u * 2u * 18 * HALO_W * 16) >> 4;
  constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
  constexpr uint32_t id3 = idesc0 | ((uint32_t)((3 * NT) >> 3) << 17), id2 = idesc0 | ((uint32_t)((2 * NT) >> 3) << 17),
                     id1 = idesc0 | ((uint32_t)(NT >> 3) << 17);
#pragma unroll
  for (int khw = 0; khw < 9; ++khw) {
    const uint32_t a_off = (uint32_t)((khw / 3) * HALO_W + khw % 3);
#pragma unroll
    for (int kk = 0; kk < CCIN / 2; ++kk) {
      const uint64_t a = a_plane + (uint64_t)(a_off + kk * KK_A);
      const uint32_t b_off = (uint32_t)((khw * CCIN + 2 * kk) * 3 * NT);
      if (khw == 0 && kk == 0) {
        tc_mma(d1, a, wz_desc0, id2, 1u, leader);                                  // kd = 2, 1: planes already open
        tc_mma(d1 + 2 * NT, a, wz_desc0 + (uint64_t)(2 * NT), id1, 0u, leader);    // kd = 0: first touch of plane j + 1
      } else {
        tc_mma(d1, a, wz_desc0 + (uint64_t)b_off, id3, 1u, leader);
      }
    }
  }
}

template <int NT, int CCIN>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_tc_roll(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ RollParams p) {
  constexpr int HALO_H = 18, HALO_W = 10;
  constexpr int CHUNK_BYTES = HALO_H * HALO_W * 16;       // one 8-channel chunk of one halo plane: 2880
  constexpr int STAGE_BYTES = CCIN * 2 * CHUNK_BYTES;     // a stage holds TWO consecutive planes: [chunk][plane][line][w]
  constexpr uint32_t KK_A = (2u * 2u * CHUNK_BYTES) >> 4; // descriptor units per K step (two chunks of two planes)
  constexpr int KS = CCIN / 2;                            // K steps of 16 channels
  extern __shared__ __align__(128) uint8_t smem[];
  sg_pdl_trigger();
  // carve-up: [weights][plane ring][barriers][tmem slot][bias]
  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + p.stages * STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  // bars: [0] w_full, then p_full[8], p_empty[8], acc_full[16], acc_empty[16]
  constexpr int W_FULL = 0, P_FULL = 1, P_EMPTY = P_FULL + kRollStagesMax, A_FULL = P_EMPTY + kRollStagesMax,
                A_EMPTY = A_FULL + kRollAccMax, N_BARS = A_EMPTY + kRollAccMax;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  float* s_bias = reinterpret_cast<float*>(bars + N_BARS + 3);   // 16-byte aligned (N_BARS + 3 is even... 52 * 8 = 416)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * NT;
  const int SEG = p.seg;
  // both rings have power-of-two sizes: slot = index & mask, phase = (index >> log2) & 1
  const uint32_t s_mask = (uint32_t)p.stages - 1u, s_log = (uint32_t)p.stages_log2, r_mask = (uint32_t)p.acc_slots - 1u,
                 r_log = (uint32_t)p.acc_log2, R = (uint32_t)p.acc_slots;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    mbar_init(BAR(W_FULL), 1);
    for (int i = 0; i < kRollStagesMax; ++i) {
      mbar_init(BAR(P_FULL + i), 1);
      mbar_init(BAR(P_EMPTY + i), 1);
    }
    for (int i = 0; i < kRollAccMax; ++i) {
      mbar_init(BAR(A_FULL + i), 1);
      mbar_init(BAR(A_EMPTY + i), 4);   // one arrival per epilogue warp
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
  sg_pdl_wait();

  // work item -> (n, segment, tile_h, tile_w); tile_w fastest so that neighbouring CTAs share halo columns in L2
  auto item_coords = [&](int item, int& n, int& d0, int& h0, int& w0) {
    int t = item;
    const int tile_w = t % p.tiles_w; t /= p.tiles_w;
    const int tile_h = t % p.tiles_h; t /= p.tiles_h;
    const int sg = t % p.n_seg; t /= p.n_seg;
    n = t; d0 = sg * SEG; h0 = tile_h * 16; w0 = tile_w * 8;
  };

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const uint32_t w_addr = smem_u32(w_smem);
      mbar_expect_tx(BAR(W_FULL), (uint32_t)p.w_bytes);
      for (int i = 0; i < 27 * CCIN; ++i) {
        // [kh][kw][chunk][kd = 2, 1, 0][co][8]: the three kd taps of a (kh, kw, chunk) are consecutive N rows
        const int tap = i / CCIN, chunk = i - tap * CCIN;
        const uint32_t dst = w_addr + ((((tap % 9) * CCIN + chunk) * 3 + (2 - tap / 9)) * NT) * 16;
        bulk_load(dst, p.wp + ((int64_t)i * p.CoutP + co0) * 8, NT * 16u, BAR(W_FULL));
      }
      const uint32_t a_addr = smem_u32(a_smem);
      uint32_t gp = 0;   // planes loaded so far
      RT(long long rt_w = 0; const long long rt_p0 = clock64();)
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int n, d0, h0, w0;
        item_coords(item, n, d0, h0, w0);
        // planes d0 + j, j = j_first .. j_last (those inside the volume), two per stage; the second plane of the last
        // pair may lie past the segment's last plane or past the volume (zero-filled by the TMA): loaded, never multiplied
        const int j_first = d0 == 0 ? 0 : -1;
        const int j_last = d0 + SEG >= p.D ? SEG - 1 : SEG;
        for (int jp = j_first; jp <= j_last; jp += 2) {
          const uint32_t s = gp & s_mask;
          RT(const long long rt_t = clock64();)
          mbar_wait(BAR(P_EMPTY + s), ((gp >> s_log) & 1u) ^ 1u);
          RT(rt_w += clock64() - rt_t;)
          mbar_expect_tx(BAR(P_FULL + s), (uint32_t)STAGE_BYTES);
          tma_load_5d(a_addr + s * STAGE_BYTES, &xmap, BAR(P_FULL + s), (w0 - 1) * 8, h0 - 1, d0 + jp, 0, n);
          ++gp;
        }
      }
      RT(g_res_timing[blockIdx.x * 8 + 3] = rt_w; g_res_timing[blockIdx.x * 8 + 7] = clock64() - rt_p0;)
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t leader = elect_one();   // all lanes run the loops; one issues
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint64_t a_desc0 = make_desc(smem_u32(a_smem), 2u * CHUNK_BYTES, HALO_W * 16u);   // chunk stride = two planes
    const uint64_t wz_desc0 = make_desc(smem_u32(w_smem), 3 * NT * 16u, 128u);
    constexpr uint32_t stage16 = (uint32_t)STAGE_BYTES >> 4;
    mbar_wait(BAR(W_FULL), 0);
    uint32_t gp = 0, go = 0;   // planes consumed / output planes opened by earlier items
    uint32_t safe_upto = 0;    // output planes below this index may be opened without looking at the epilogue's barriers
    bool next_ready = false;   // the stage about to be consumed has already been seen full
    RT(long long rt_full = 0, rt_acc = 0; int rt_items = 0; const long long rt_i0 = clock64();)
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int n, d0, h0, w0;
      item_coords(item, n, d0, h0, w0);
      const int j_first = d0 == 0 ? 0 : -1;
      const int j_last = d0 + SEG >= p.D ? SEG - 1 : SEG;
      int r_next = 0;   // next output plane of this item to be reported complete
      for (int jp = j_first; jp <= j_last; jp += 2) {
        const uint32_t s = gp & s_mask;
        if (!next_ready) {
          RT(const long long rt_u = clock64();)
          mbar_wait(BAR(P_FULL + s), (gp >> s_log) & 1u);
          RT(rt_full += clock64() - rt_u;)
        }
        tc_fence_after();
        // the producer runs stages ahead: look at the NEXT stage's barrier now, without blocking -- the latency of the
        // test hides behind this pair's MMA issue, and the (normally true) answer saves the blocking wait next time round
        next_ready = mbar_test_wait(BAR(P_FULL + ((gp + 1u) & s_mask)), ((gp + 1u) >> s_log) & 1u);
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          const int j = jp + q;
          if (j > j_last) break;
          const uint64_t a_plane = a_desc0 + (uint64_t)(s * stage16 + (uint32_t)q * (CHUNK_BYTES >> 4));
          {
            // interior plane with contiguous slots: the fast path
            const uint32_t g_hi = go + (uint32_t)(j - 1);
            const uint32_t slot = g_hi & r_mask;
            if (j >= 1 && j <= SEG - 2 && slot + 2u < R) {
              const uint32_t g = g_hi + 2u;                 // the plane this input plane opens
              if (g >= safe_upto) {
                const uint32_t ahead = (R >> 1) - 1u;
                if (g + ahead >= R) {
                  const uint32_t h = g + ahead - R;
                  RT(const long long rt_t = clock64();)
                  mbar_wait(BAR(A_EMPTY + (h & r_mask)), (h >> r_log) & 1u);
                  RT(rt_acc += clock64() - rt_t;)
                  tc_fence_after();
                }
                safe_upto = g + ahead + 1u;
              }
              roll_issue_interior<NT, CCIN>(tmem_base + slot * NT, a_plane, wz_desc0, leader);
              continue;
            }
          }
          // input plane j feeds the output planes r = j + 1 - kd, kd in [kd_lo, kd_hi], of this segment
          const int kd_hi = j + 1 < 2 ? j + 1 : 2, kd_lo = j + 2 - SEG > 0 ? j + 2 - SEG : 0;
          // taps <= fresh_max touch their output plane for the first time: kd = 0 always does (plane j + 1 opens); when
          // the plane before the volume was skipped, the kd = 1 tap of the first plane opens output plane 0 as well
          const int fresh_max = kd_lo > 0 ? -1 : (j == 0 && j_first == 0 ? 1 : 0);
          // Opening output plane g re-uses the slot of plane g - R, which the epilogue must have drained.  The epilogue
          // drains in order, so ONE wait per R/2 planes -- on the drain of plane g - R + R/2 - 1, complete since input
          // plane g - R/2 was issued -- covers the next R/2 openings (a barrier wait per plane cost ~150 cycles of the
          // issuing thread, which the MMA queue does not hide).
          if (kd_lo <= fresh_max) {
            const uint32_t g = go + (uint32_t)(j + 1 - kd_lo);      // the newest plane opened by this input plane
            if (g >= safe_upto) {
              const uint32_t ahead = (R >> 1) - 1u;
              if (g + ahead >= R) {
                const uint32_t h = g + ahead - R;
                RT(const long long rt_t = clock64();)
                mbar_wait(BAR(A_EMPTY + (h & r_mask)), (h >> r_log) & 1u);
                RT(rt_acc += clock64() - rt_t;)
                tc_fence_after();
              }
              safe_upto = g + ahead + 1u;
            }
          }
          // accumulator slots ascend as kd descends; a run ends where the ring wraps
          const uint32_t slot_hi = (go + (uint32_t)(j + 1 - kd_hi)) & r_mask;
          const int n_kd = kd_hi - kd_lo + 1;
          const int len1 = n_kd < (int)(R - slot_hi) ? n_kd : (int)(R - slot_hi);
          const int len2 = n_kd - len1;
          const uint32_t d1 = tmem_base + slot_hi * NT, d2 = tmem_base;
          const uint64_t b1 = wz_desc0 + (uint64_t)((2 - kd_hi) * NT), b2 = b1 + (uint64_t)(len1 * NT);
          const uint32_t id1 = idesc0 | ((uint32_t)((len1 * NT) >> 3) << 17), id2 = idesc0 | ((uint32_t)((len2 * NT) >> 3) << 17);
#pragma unroll
          for (int khw = 0; khw < 9; ++khw) {
            const uint32_t a_off = (uint32_t)((khw / 3) * HALO_W + khw % 3);
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              const uint64_t a = a_plane + (uint64_t)(a_off + kk * KK_A);
              const uint32_t b_off = (uint32_t)((khw * CCIN + 2 * kk) * 3 * NT);
              if (khw == 0 && kk == 0) {
                // first MMA of the plane: one instruction has one accumulate flag, so the taps are grouped by (flag,
                // contiguous slots)
                int kd = kd_hi;
                while (kd >= kd_lo) {
                  const uint32_t slot = (go + (uint32_t)(j + 1 - kd)) & r_mask;
                  const bool acc = kd > fresh_max;
                  int len = 1;
                  while (kd - len >= kd_lo && ((kd - len) > fresh_max) == acc && slot + (uint32_t)len < R) ++len;
                  tc_mma(tmem_base + slot * NT, a, wz_desc0 + (uint64_t)((2 - kd) * NT),
                         idesc0 | ((uint32_t)((len * NT) >> 3) << 17), acc ? 1u : 0u, leader);
                  kd -= len;
                }
              } else {
                tc_mma(d1, a, b1 + (uint64_t)b_off, id1, 1u, leader);
                if (len2 > 0) tc_mma(d2, a, b2 + (uint64_t)b_off, id2, 1u, leader);
              }
            }
          }
        }
#ifdef SG_ROLL_EXTRA_COMMIT
        tc_commit(BAR(W_FULL), leader);           // probe: what one more commit costs (nobody waits on it): nothing
#endif
        tc_commit(BAR(P_EMPTY + s), leader);     // the stage is free once these MMAs have read it
        ++gp;
        // output planes up to j_end - 1 have received their last tap (kd = 2 from plane r + 1); with the plane behind the
        // volume skipped, the segment's last output plane is complete with its own input plane
        const int j_end = jp + 1 < j_last ? jp + 1 : j_last;
        const int r_done = (j_end == j_last && j_last == SEG - 1) ? SEG - 1 : j_end - 1;
        for (; r_next <= r_done; ++r_next) tc_commit(BAR(A_FULL + ((go + (uint32_t)r_next) & r_mask)), leader);
      }
      go += (uint32_t)SEG;
      RT(++rt_items;)
    }
    RT(if (lane == 0) {
      g_res_timing[blockIdx.x * 8 + 0] = clock64() - rt_i0;
      g_res_timing[blockIdx.x * 8 + 1] = rt_full;
      g_res_timing[blockIdx.x * 8 + 2] = rt_acc;
      g_res_timing[blockIdx.x * 8 + 6] = rt_items;
    })
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
    const uint32_t trow0 = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int64_t plane8 = (int64_t)p.H * p.W * 8;
    uint32_t go = 0;
    RT(long long rt_e = 0; const long long rt_e0 = clock64();)
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int n, d0, h0, w0;
      item_coords(item, n, d0, h0, w0);
      const int64_t obase0 =
          (((int64_t)n * p.CCout + co0 / 8) * V + ((int64_t)d0 * p.H + h0 + (row >> 3)) * p.W + w0 + (row & 7)) * 8;
      if (p.pool_y != nullptr) {
        // conv -> [lrelu] -> 2x2x2 average: two planes at a time (SEG and d0 are even)
        const int Hp = p.H >> 1, Wp = p.W >> 1;
        const int64_t Vp = (int64_t)(p.D >> 1) * Hp * Wp;
        const bool writer = (lane & 9) == 0;                       // even w (lane bit 0) and even h (lane bit 3)
        const int64_t pbase0 = (((int64_t)n * p.CCout + co0 / 8) * Vp +
                                ((int64_t)(d0 >> 1) * Hp + ((h0 + (row >> 3)) >> 1)) * Wp + ((w0 + (row & 7)) >> 1)) * 8;
#pragma unroll 1
        for (int r = 0; r < SEG; r += 2) {
          const uint32_t g0 = go + (uint32_t)r, g1 = g0 + 1;
          const uint32_t s0 = g0 & r_mask, s1 = g1 & r_mask;
          mbar_wait(BAR(A_FULL + s0), (g0 >> r_log) & 1u);
          mbar_wait(BAR(A_FULL + s1), (g1 >> r_log) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < NT; c0 += 16) {
            float v0[16], v1[16];
            __syncwarp();
            tmem_ld16(trow0 + s0 * NT + (uint32_t)c0, v0);
            tmem_ld16(trow0 + s1 * NT + (uint32_t)c0, v1);
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
              for (int q = 0; q < 8; ++q) {
                o0.v[q] = v0[half * 8 + q];
                o1.v[q] = v1[half * 8 + q];
                op.v[q] = sum[half * 8 + q];
              }
              const int64_t ch = (int64_t)(c0 / 8 + half);
              st8(yout + obase0 + r * plane8 + ch * V * 8, o0);
              st8(yout + obase0 + (r + 1) * plane8 + ch * V * 8, o1);
              if (writer) st8(p.pool_y + pbase0 + (int64_t)(r >> 1) * Hp * Wp * 8 + ch * Vp * 8, op);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(BAR(A_EMPTY + s0));
            mbar_arrive(BAR(A_EMPTY + s1));
          }
        }
        go += (uint32_t)SEG;
        continue;
      }
      uint4 mk[NT / 8];
      if (mask) {
#pragma unroll
        for (int c = 0; c < NT / 8; ++c) mk[c] = __ldg(reinterpret_cast<const uint4*>(mask + obase0 + (int64_t)c * V * 8));
      }
#pragma unroll 1
      for (int r = 0; r < SEG; ++r) {
        const uint32_t g = go + (uint32_t)r;
        const uint32_t slot = g & r_mask;
        const uint32_t trow = trow0 + slot * NT;
        uint4 mn[NT / 8];      // the next plane's masks, in flight while this plane is drained
        if (mask && r + 1 < SEG) {
#pragma unroll
          for (int c = 0; c < NT / 8; ++c)
            mn[c] = __ldg(reinterpret_cast<const uint4*>(mask + obase0 + (r + 1) * plane8 + (int64_t)c * V * 8));
        }
        RT(const long long rt_t = clock64();)
        mbar_wait(BAR(A_FULL + slot), (g >> r_log) & 1u);
        RT(rt_e += clock64() - rt_t;)
        tc_fence_after();
        if (p.pn_y != nullptr) {
          // conv -> [lrelu] -> pixel-norm [-> lrelu]: this thread holds all NT channels of its voxel
          float t[NT];
          __syncwarp();
          tmem_ld_block<NT>(trow, t);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < NT; ++i) {
            float a = fmaf(t[i], scale, s_bias[i]);
            if (lrelu) a = lrelu02(a);
            t[i] = a;
            ss = fmaf(a, a, ss);
          }
          const float rn = rsqrtf(ss * p.pn_inv_c + p.pn_eps);
          __nv_bfloat16* y0 = yout + obase0 + r * plane8;
          __nv_bfloat16* y1 = p.pn_y + obase0 + r * plane8;
#pragma unroll
          for (int c = 0; c < NT / 8; ++c) {
            F8 a, b;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              a.v[q] = t[8 * c + q];
              const float u = t[8 * c + q] * rn;
              b.v[q] = p.pn_lrelu_after ? lrelu02(u) : u;
            }
            st8(y0 + (int64_t)c * V * 8, a);
            st8(y1 + (int64_t)c * V * 8, b);
          }
        } else {
#pragma unroll
          for (int c0 = 0; c0 < NT; c0 += 16) {
            float v[16];
            __syncwarp();
            tmem_ld16(trow + (uint32_t)c0, v);
            epilogue16_regmask(v, s_bias + c0, scale, lrelu, mask != nullptr, mk[c0 / 8], mk[c0 / 8 + 1],
                               yout + obase0 + r * plane8 + (int64_t)(c0 / 8) * V * 8, V * 8);
          }
        }
        if (mask && r + 1 < SEG) {
#pragma unroll
          for (int c = 0; c < NT / 8; ++c) mk[c] = mn[c];
        }
        // this warp is done reading the slot: hand it back to the MMA issuer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(A_EMPTY + slot));
      }
      go += (uint32_t)SEG;
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

struct RollPlan {
  bool ok = false;
  int NT = 0, CCin = 0;
  RollParams p{};
  size_t smem = 0;
  dim3 grid;
};

// SG_TC_ROLL in the environment (read once): 0 = never, 1 = wherever the plan fits, 2 (default) = where it measured
// faster than the tiled kernel (tools/res_sweep.py, B200, 32x128x128): Cin = 64 (36 MMAs per plane: 64 -> 32 191 us
// against 223) and the NT = 64 layers (32 -> 64 162 us against 171-182, 207 against 240 with a mask); with Cin <= 32 and
// NT <= 32 a plane is 9-18 MMAs and the per-plane bookkeeping of the issuing thread, which the shallow MMA queue does not
// hide, eats the gain (32 -> 32 114 us either way, 16 -> 16 75 us against 55).
inline int roll_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SG_TC_ROLL");
    mode = e ? atoi(e) : 2;
  }
  return mode;
}
inline bool roll_wanted(const RollPlan& pl);

RollPlan make_roll_plan(int N, int Cin, int Cout, int D, int H, int W, bool for_test = false) {
  RollPlan pl;
  RollParams& p = pl.p;
  if (W % 8 != 0 || H % 16 != 0 || D < 4) return pl;
  const int CCin = sg_chunks(Cin), CoutP = 16 * ((Cout + 15) / 16);
  if (CCin != 2 && CCin != 4 && CCin != 8) return pl;
  int NT = 0;
  for (int cand : {64, 32, 16}) {
    if (CoutP % cand) continue;
    if (27 * CCin * cand * 16 <= 120 * 1024) { NT = cand; break; }
  }
  if (NT == 0 || CoutP / NT > 1) return pl;
  p.w_bytes = 27 * CCin * NT * 16;
  const int plane_bytes = 2 * CCin * 18 * 10 * 16;     // one stage = two planes
  const int budget = 222 * 1024 - 1024;     // barriers, tmem slot, bias
  int stages = (budget - p.w_bytes) / plane_bytes;
  if (stages < 2) return pl;
  stages = stages >= 8 ? 8 : stages >= 4 ? 4 : 2;      // power of two: ring indices are masks
  int R = 512 / NT;
  if (R > kRollAccMax) R = kRollAccMax;
  // depth segments: the fewest planes per item that still give every CTA a balanced share (cost in plane equivalents:
  // SEG planes + ~0.8 per interior segment end for its narrow MMAs)
  const int cols = N * (H / 16) * (W / 8);
  const int sms = sg_num_sms();
  double best = 1e30;
  int best_seg = 0;
  for (int seg = D; seg >= 4; seg /= 2) {
    if (D % seg || seg % 2) break;
    const int n_seg = D / seg;
    const int64_t items = (int64_t)cols * n_seg;
    if (!for_test && items < 2 * (int64_t)(items < sms ? items : sms)) continue;
    const double rounds = (double)((items + sms - 1) / sms);
    const double interior_ends = n_seg > 1 ? 2.0 * (n_seg - 1) / n_seg : 0.0;
    const double cost = rounds * (seg + 0.8 * interior_ends + 0.5);
    if (cost < best) { best = cost; best_seg = seg; }
  }
  if (best_seg == 0) return pl;
  p.seg = best_seg;
  p.n_seg = D / best_seg;
  p.tiles_w = W / 8; p.tiles_h = H / 16;
  p.n_items = cols * p.n_seg;
  p.stages = stages;
  p.stages_log2 = stages == 8 ? 3 : stages == 4 ? 2 : 1;
  p.acc_slots = R;
  p.acc_log2 = R == 16 ? 4 : R == 8 ? 3 : 2;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.Cout = Cout; p.CoutP = CoutP; p.CCout = sg_chunks(Cout);
  p.tmem_cols = round_pow2_cols(R * NT);
  int ctas = sms < p.n_items ? sms : p.n_items;
  if (for_test) {
    ctas = p.n_items >= 3 ? p.n_items / 3 : 1;   // tests: every CTA walks ~3 items (ring wrap, item boundaries)
    if (ctas > sms) ctas = sms;
  }
  pl.grid = dim3((unsigned)ctas, 1, 1);
  pl.smem = (size_t)p.w_bytes + (size_t)stages * plane_bytes + 8 * (1 + 2 * kRollStagesMax + 2 * kRollAccMax) + 16 + 4 * 64 + 16;
  pl.NT = NT;
  pl.CCin = CCin;
  pl.ok = true;
  return pl;
}

inline bool roll_wanted(const RollPlan& pl) {
  if (!pl.ok || roll_mode() == 0) return false;
  return roll_mode() == 1 || pl.CCin == 8 || pl.NT == 64;
}

template <int NT, int CCIN>
int launch_roll_inst(const RollPlan& pl, const CUtensorMap& map, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc_roll<NT, CCIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      sg_set_error("conv_tc_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  sg_launch((k_conv_tc_roll<NT, CCIN>), pl.grid, kThreads, pl.smem, s, map, pl.p);
  return sg_check_launch("sg_conv3d_fprop(tcgen05 rolling)");
}

template <int NT>
int launch_roll(const RollPlan& pl, const CUtensorMap& map, cudaStream_t s) {
  switch (pl.CCin) {
    case 2: return launch_roll_inst<NT, 2>(pl, map, s);
    case 4: return launch_roll_inst<NT, 4>(pl, map, s);
    default: return launch_roll_inst<NT, 8>(pl, map, s);
  }
}

}  // namespace
