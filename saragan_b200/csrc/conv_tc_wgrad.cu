// tcgen05 / TMEM / TMA weight-gradient of the 3x3x3 convolution for sm_100a:
//   gw[co][ci][tap] = scale * sum_{n,p} gy[n][co][p] * x[n][ci][p + tap - 1]     (fp32 out)
//
// GEMM view per tap: D[M = co][N = ci] += A[co][K = voxel] * B[ci][K = voxel]^T with K running
// over the voxels of a tile.  In the channel-blocked layout a voxel is a 16-byte vector of 8
// channels, i.e. both operands are "MN-major" UMMA operands straight out of the TMA tiles:
//   element (channel c, voxel k) at (c/8)*SBO + (k/8)*LBO + (k%8)*16 + (c%8)*2
// with LBO = the line pitch (one K=16 MMA step = two consecutive lines x 8 voxels) and SBO = the
// stride between 8-channel chunks.
//
// A tap's output is only [Cout x NT] -- far too small for the tensor core, whose M=128 MMA costs
// ~45 cycles whatever N <= 64 is (tools/mma_bench.cu).  So taps are STACKED along both MMA axes:
//  * N: the three kw taps.  The x tile is loaded three times, shifted by one voxel in w each (3 TMA
//    boxes of 8 w-voxels instead of one of 10), into consecutive chunk slots; one B descriptor then
//    spans 3 x NT/8 slots = the N = 3*NT columns [kw][ci] and one MMA does three taps.
//  * M: the kd taps, when Cout < 128 leaves M = 128 rows to spare.  The gy tile sits in shared memory
//    as [plane][chunk][line] (one TMA box per plane), so with SBO = one plane of one chunk the 16 chunk
//    strides of the A operand walk through chunk 0..g-1 of plane d, then chunk 0..g-1 of plane d+1, ...:
//    row block s of the accumulator pairs x plane d with gy plane d + s + first_shift, i.e. it IS the tap
//    kd = 1 - (s + first_shift).  Cout <= 32: all three kd in one MMA (gy tile with a +-1 plane halo);
//    Cout <= 64: two (kd = 1, 0), a second CTA group does kd = 2; Cout > 64: one kd per CTA group.
//  * the three kh taps shift the B descriptor start by one line (3 MMAs per K step).
// The BIAS gradient sum_p gy[co,p] rides along for free: two extra chunk slots behind the x copies are
// filled with ones once per CTA, and the kh = 1 MMAs of one CTA group run with N = 3*NT + 16, so 16 more
// accumulator columns hold gy (x) 1 (the rows of the unshifted block are the bias gradient).
// Each CTA keeps its 3 accumulators [128 x 3*NT] in tensor memory while it streams through its
// share of the voxel tiles (persistent, multi-stage TMA pipeline), then adds them -- 16-byte vector
// reductions, or plain stores when it is the only CTA of its group -- into a tap-major fp32 workspace
// [27][Cout][CinP]; k_wgrad_finish transposes that into the parameter layout [Cout][Cin][27] (scattered
// 4-byte atomics straight into that layout cost more than the MMAs at the low-resolution levels).
//
// SPLIT variant -- the weight gradient of the fp32 / TF32 levels.  tcgen05 kind::tf32 does not take MN-major operands
// (tools/tf32_mn_probe.cu: an MN-major tf32 MMA returns zeros on B200, whatever the header comments of CUTLASS say), and
// the voxel-contraction of a wgrad on this layout IS MN-major.  So the fp32 tensors are split into bf16 halves in
// shared memory, x = hi + lo with hi = bf16(x), lo = bf16(x - hi), and the product is formed from three bf16 MMAs,
//   gy (x) x  ~  g_hi (x) x_hi + g_lo (x) x_hi + g_hi (x) x_lo        (lo (x) lo ~ 2^-18 relative is dropped),
// i.e. 16 mantissa bits per operand: more accurate than TF32's 10.  The fp32 tensors are read through two tensor maps
// (one per 16-byte half of a voxel's 32-byte chunk, see conv_tc.cu) into two regions H0 / H1 of identical geometry;
// the four epilogue warps, idle during the main loop, turn {H0, H1} = {channels 0-3, channels 4-7} of every voxel in
// place into {H0, H1} = {hi, lo} of its 8 channels -- each then IS a bf16 tile of the layout described above.
#include <type_traits>

#include "../../include/saragan_b200.h"
#include "tc_common.cuh"
SG_DEFINE_LEAK_SETTER(sg_set_leak_conv_tc_wgrad)

int sg_encode_f32_half_map(CUtensorMap* map, const void* x, int half, int N, int CC8, int D, int H, int W, int box_w,
                           int box_h, int box_d, int box_c);

namespace {

constexpr int kThreadsW = 192;

struct WgParams {
  float* ws;               // [27][Cout][CinP] tap-major sums
  float* gb;               // nullable [Cout]: bias gradient = sum of gy over batch and voxels
  int N, D, H, W;
  int Cin, Cout, CCin, CCout, CinP;
  int td, th;              // tile = td x th x 8 voxels
  int tiles_w, tiles_h, tiles_d;
  int n_tiles;             // tiles_w * tiles_h * tiles_d * N
  int g_chunks;            // 8-channel chunks of gy per plane box (<= 16)
  int shifts;              // kd taps stacked along M: 3, 2 or 1
  int g_planes;            // gy planes loaded per tile: td + shifts - 1
  int g_plane_bytes;       // g_chunks * th * 8 * 16
  int x_chunk_bytes;       // td*(th+2)*8*16: one (kw copy, 8-channel chunk) slot of the x tile
  int stage_bytes;
  int stages;
  int slack_bytes;         // room after the last stage for the garbage-row reads of the M = 128 operand
  int ci_tiles;            // CinP / NT
  int tmem_cols;
  int direct;              // one CTA per group: plain stores, no zero-fill of ws needed
};

// in place: {h0[u], h1[u]} = fp32 channels {0-3, 4-7} of voxel u  ->  {bf16 hi, bf16 lo} of its 8 channels
// (WANT_LO = false: only the hi half is formed -- plain bf16 operands, SPLIT == 2)
template <bool WANT_LO>
__device__ __forceinline__ void split_tile_bf16(uint8_t* h0, uint8_t* h1, int bytes, int tid, int nthreads) {
  for (int u = tid; u < bytes / 16; u += nthreads) {
    const float4 a = reinterpret_cast<const float4*>(h0)[u], b = reinterpret_cast<const float4*>(h1)[u];
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, lo;
    __nv_bfloat162* ph = reinterpret_cast<__nv_bfloat162*>(&hi);
    __nv_bfloat162* pl = reinterpret_cast<__nv_bfloat162*>(&lo);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ph[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      const float2 f = __bfloat1622float2(ph[i]);
      pl[i] = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    }
    reinterpret_cast<uint4*>(h0)[u] = hi;
    if (WANT_LO) reinterpret_cast<uint4*>(h1)[u] = lo;
  }
}

// SPLIT: 0 = bf16 tensors; 1 = fp32 tensors, three bf16 MMAs per product (hi/lo split, ~16 mantissa bits);
//        2 = fp32 tensors, bf16 operands (hi halves only) -- the weight gradient of the fp32-storage levels under the
//            bf16 policy: rounding its operands to bf16 adds 3e-3 of noise to a weight gradient and flips no mask
template <int NT, int SPLIT>
__global__ void __launch_bounds__(kThreadsW)
k_wgrad_tc(const __grid_constant__ CUtensorMap gmap, const __grid_constant__ CUtensorMap gmap1,
           const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap xmap1,
           const __grid_constant__ WgParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  sg_pdl_trigger();
  // carve-up: stages x [gy planes | 3 kw copies of the x halo tile | ones], then barriers
  // (SPLIT: stages x [gy H0 | gy H1 | x H0 | ones | x H1])
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes + p.slack_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * 8 + 1);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int FULL = 0, EMPTY = p.stages, ACC_FULL = 2 * p.stages, READY = 2 * p.stages + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.y;
  const int co_tile = blockIdx.z / p.ci_tiles, ci_tile = blockIdx.z % p.ci_tiles;
  const int x_chunks = NT / 8;
  const int halo_h = p.th + 2;
  // CTA group -> which planes it pairs.  Row block s of the accumulator holds the tap kd_hi - s.
  //   shifts 3: gy planes from d0 - 1, x plane d0      -> kd = 2, 1, 0
  //   shifts 2: group 0: gy from d0, x plane d0        -> kd = 1, 0;  group 1: x plane d0 + 1 -> kd = 2 (s = 0 only)
  //   shifts 1: group g: gy from d0, x plane d0 + g - 1 -> kd = g
  const int g_plane0 = p.shifts == 3 ? -1 : 0;
  const int x_plane0 = p.shifts == 3 ? 0 : p.shifts == 2 ? grp : grp - 1;
  const int kd_hi = p.shifts == 3 ? 2 : p.shifts == 2 ? 1 + grp : grp;
  const int live_shifts = (p.shifts == 2 && grp == 1) ? 1 : p.shifts;
  const int rows_per_shift = 8 * p.g_chunks;
  // the row block whose gy plane is unshifted against its own voxels (kd = 1) carries the bias gradient
  const int bias_row0 = (kd_hi - 1) * rows_per_shift;
  const bool do_bias = p.gb != nullptr && ci_tile == 0 && kd_hi >= 1 && kd_hi - 1 < live_shifts;
  const int g_bytes = p.g_planes * p.g_plane_bytes;                  // one (bf16-sized) set of gy planes
  const int xh_bytes = 3 * x_chunks * p.x_chunk_bytes;               // one set of the three x copies
  const int x_off = SPLIT ? 2 * g_bytes : g_bytes;                   // x (H0) region of a stage, the ones slots behind it
  const int xlo_off = x_off + xh_bytes + 2 * p.x_chunk_bytes;        // SPLIT: x H1 region
  if (do_bias) {
    // ones slots (bf16 1.0 = 0x3F80) behind the 3*x_chunks x slots of every stage; the TMA never writes them
    for (int st = 0; st < p.stages; ++st) {
      uint32_t* ones = reinterpret_cast<uint32_t*>(smem + (size_t)st * p.stage_bytes + x_off + xh_bytes);
      for (int i = threadIdx.x; i < 2 * p.x_chunk_bytes / 4; i += blockDim.x) ones[i] = 0x3F803F80u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
    if constexpr (SPLIT) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap1) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap1) : "memory");
    }
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(BAR(FULL + i), 1);
      mbar_init(BAR(EMPTY + i), 1);
      mbar_init(BAR(READY + i), 128);   // SPLIT: every thread of the four epilogue warps after the hi/lo pass
    }
    mbar_init(BAR(ACC_FULL), 1);
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
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(g_bytes + xh_bytes) * (SPLIT ? 2u : 1u);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        int t = tile;
        const int tw_ = t % p.tiles_w; t /= p.tiles_w;
        const int th_ = t % p.tiles_h; t /= p.tiles_h;
        const int td_ = t % p.tiles_d; t /= p.tiles_d;
        const int n = t;
        const int w0 = tw_ * 8, h0 = th_ * p.th, d0 = td_ * p.td;
        const int s = it % p.stages;
        mbar_wait(BAR(EMPTY + s), ((it / p.stages) & 1) ^ 1);
        mbar_expect_tx(BAR(FULL + s), tx);
        const uint32_t g_dst = smem_base + s * p.stage_bytes;
        const uint32_t x_dst = g_dst + x_off;
        if constexpr (SPLIT) {
          // fp32 tensors as [4 | W | H | D | N*CC8], one map per 16-byte half of the voxels; H0 / H1 regions alike
          for (int hf = 0; hf < 2; ++hf) {
            for (int pl = 0; pl < p.g_planes; ++pl)
              tma_load_5d(g_dst + hf * g_bytes + pl * p.g_plane_bytes, hf ? &gmap1 : &gmap, BAR(FULL + s), 0, w0, h0,
                          d0 + g_plane0 + pl, n * p.CCout + co_tile * 16);
            for (int k = 0; k < 3; ++k)
              for (int c = 0; c < x_chunks; ++c)
                tma_load_5d(g_dst + (hf ? xlo_off : x_off) + (k * x_chunks + c) * p.x_chunk_bytes, hf ? &xmap1 : &xmap,
                            BAR(FULL + s), 0, w0 - 1 + k, h0 - 1, d0 + x_plane0, n * p.CCin + ci_tile * x_chunks + c);
          }
        } else {
          for (int pl = 0; pl < p.g_planes; ++pl)   // one box = all chunks of one plane (planes outside D: zeros)
            tma_load_5d(g_dst + pl * p.g_plane_bytes, &gmap, BAR(FULL + s), w0 * 8, h0, d0 + g_plane0 + pl, co_tile * 16, n);
          for (int k = 0; k < 3; ++k)        // kw copy k = the tile shifted by k - 1 voxels in w
            for (int c = 0; c < x_chunks; ++c)
              tma_load_5d(x_dst + (k * x_chunks + c) * p.x_chunk_bytes, &xmap, BAR(FULL + s), (w0 - 1 + k) * 8, h0 - 1,
                          d0 + x_plane0, ci_tile * x_chunks + c, n);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t leader = elect_one();   // all lanes run the loops; one issues
      // D=f32, A=B=bf16, both MN-major (bits 15,16), N = 3*NT ([kw][ci]), M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)((3 * NT) >> 3) << 17) | ((128u >> 4) << 24);
      // the kh = 1 MMA of a bias-computing CTA also spans the two ones slots: N = 3*NT + 16
      const uint32_t idesc_mid = do_bias ? ((1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                            ((uint32_t)((3 * NT + 16) >> 3) << 17) | ((128u >> 4) << 24))
                                         : idesc;
      // descriptors: bases hoisted, per-MMA cost = one 64-bit add (offsets in 16-byte units = voxels)
      const uint64_t g_desc0 = make_desc(smem_base, 128u, (uint32_t)(p.th * 128));
      const uint64_t x_desc0 = make_desc(smem_base + x_off, 128u, (uint32_t)p.x_chunk_bytes);
      const uint32_t g_lo16 = (uint32_t)g_bytes >> 4, x_lo16 = (uint32_t)(xlo_off - x_off) >> 4;   // SPLIT: H1 regions
      const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4, plane16 = (uint32_t)p.g_plane_bytes >> 4;
      const int ksteps_per_plane = p.th / 2, td = p.td, stages = p.stages;
      int s = 0, ph = 0;
      uint32_t acc = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        mbar_wait(BAR((SPLIT ? READY : FULL) + s), ph);
        tc_fence_after();
        const uint64_t g_stage = g_desc0 + (uint64_t)(s * stage16);
        const uint64_t x_stage = x_desc0 + (uint64_t)(s * stage16);
        for (int dl = 0; dl < td; ++dl) {
          uint64_t a_k = g_stage + (uint64_t)(dl * plane16);
          uint64_t b_k = x_stage + (uint64_t)(dl * halo_h * 8);
          for (int j = 0; j < ksteps_per_plane; ++j) {
            // accumulator columns: kh=0 at 0, kh=1 at 3*NT (3*NT + 16 wide), kh=2 at 6*NT + 16
            tc_mma(tmem_base, a_k, b_k, idesc, acc, leader);
            tc_mma(tmem_base + 3 * NT, a_k, b_k + 8, idesc_mid, acc, leader);
            tc_mma(tmem_base + 6 * NT + 16, a_k, b_k + 16, idesc, acc, leader);
            acc = 1;
            if constexpr (SPLIT == 1) {
              // + g_lo (x) x_hi (the ones columns take g_lo too: bias = sum of hi + lo) + g_hi (x) x_lo
              const uint64_t a_lo = a_k + g_lo16, b_lo = b_k + x_lo16;
              tc_mma(tmem_base, a_lo, b_k, idesc, 1u, leader);
              tc_mma(tmem_base + 3 * NT, a_lo, b_k + 8, idesc_mid, 1u, leader);
              tc_mma(tmem_base + 6 * NT + 16, a_lo, b_k + 16, idesc, 1u, leader);
              tc_mma(tmem_base, a_k, b_lo, idesc, 1u, leader);
              tc_mma(tmem_base + 3 * NT, a_k, b_lo + 8, idesc, 1u, leader);
              tc_mma(tmem_base + 6 * NT + 16, a_k, b_lo + 16, idesc, 1u, leader);
            }
            a_k += 16;   // two lines of 8 voxels (gy plane and x copies alike)
            b_k += 16;
          }
        }
        tc_commit(BAR(EMPTY + s), leader);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      tc_commit(BAR(ACC_FULL), leader);
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int shift = row / rows_per_shift;
    const int co = co_tile * 128 + row - shift * rows_per_shift;
    const int kd = kd_hi - shift;
    if constexpr (SPLIT) {
      // main loop duty of these warps: the in-place fp32 -> bf16 hi/lo pass over every landed stage
      int s2 = 0;
      for (int tile = blockIdx.x, it = 0; tile < p.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(BAR(FULL + s2), (it / p.stages) & 1);
        uint8_t* st = smem + (size_t)s2 * p.stage_bytes;
        split_tile_bf16<SPLIT == 1>(st, st + g_bytes, g_bytes, row, 128);
        split_tile_bf16<SPLIT == 1>(st + x_off, st + xlo_off, xh_bytes, row, 128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(BAR(READY + s2));
        if (++s2 == p.stages) s2 = 0;
      }
    }
    mbar_wait(BAR(ACC_FULL), 0);
    tc_fence_after();
    const bool mine = blockIdx.x < p.n_tiles && shift < live_shifts && co < p.Cout;
    float* dst_row = p.ws + ((int64_t)(kd * 9) * p.Cout + co) * p.CinP + ci_tile * NT;
    for (int t9 = 0; t9 < 9; ++t9) {
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        __syncwarp();
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t9 * NT + (t9 >= 6 ? 16 : 0) + c0), v);
        if (mine) {
          float* dst = dst_row + (int64_t)t9 * p.Cout * p.CinP + c0;
          if (p.direct) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
      }
    }
    if (do_bias) {
      // 16 identical columns gy (x) 1 behind the kh = 1 accumulator: column 0 is the bias gradient
      float v[16];
      __syncwarp();
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(6 * NT), v);
      const int cb = co_tile * 128 + row - bias_row0;
      if (blockIdx.x < p.n_tiles && row >= bias_row0 && row < bias_row0 + rows_per_shift && cb < p.Cout)
        atomicAdd(p.gb + cb, v[0]);
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

// ws [27][Cout][CinP] -> gw [Cout][Cin][27] * scale.  block = (64 input channels, one output channel)
__global__ void __launch_bounds__(256)
k_wgrad_finish(const float* __restrict__ ws, float* __restrict__ gw, int Cout, int Cin, int CinP, float scale) {
  sg_pdl_enter();
  __shared__ float s[64 * 27];
  const int co = blockIdx.y, ci0 = blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 27 * 64; idx += 256) {
    const int tap = idx >> 6, cl = idx & 63;
    float v = 0.f;
    if (ci0 + cl < Cin) v = ws[((int64_t)tap * Cout + co) * CinP + ci0 + cl] * scale;
    s[cl * 27 + tap] = v;
  }
  __syncthreads();
  const int n = (Cin - ci0 < 64 ? Cin - ci0 : 64) * 27;
  float* dst = gw + ((int64_t)co * Cin + ci0) * 27;
  for (int idx = threadIdx.x; idx < n; idx += 256) dst[idx] = s[idx];
}

struct WgPlan {
  bool ok = false;
  int NT = 0;
  WgParams p{};
  size_t smem = 0;
  dim3 grid;
  int64_t ws_bytes = 0;
};

WgPlan make_wgrad_plan(int N, int Cin, int Cout, int D, int H, int W, bool split = false) {
  WgPlan pl;
  WgParams& p = pl.p;
  if (W % 8 != 0 || H % 2 != 0 || H < 2) return pl;
  const int CinP = 16 * ((Cin + 15) / 16), CoutP = 16 * ((Cout + 15) / 16);
  p.g_chunks = CoutP / 8 < 16 ? CoutP / 8 : 16;
  const int fit = 16 / p.g_chunks;            // row blocks of 8*g_chunks rows in M = 128
  p.shifts = fit >= 3 ? 3 : fit;
  // Tile = td x th x 8 voxels, NT input channels per CTA: as many planes per tile (4, 2, 1) as leave room for >= 2
  // pipeline stages.  The M = 128 operand reads 16 slot strides from its plane on: past the gy planes these are
  // garbage rows (discarded) read from the following bytes -- `slack` keeps those reads of the LAST stage inside the
  // allocation.  The SPLIT stages hold every tile twice (H0 / H1): they may fall back to th = 8 and NT = 16.
  int td = 0, stages = 0, NT = 0, th = 0;
  const int th_max = H < 16 ? H : 16;
  if (H % th_max) return pl;
  const int th_cands[2] = {th_max, (split && th_max > 8) ? 8 : 0};
  const int nt_cands[2] = {CinP % 32 == 0 ? 32 : 16, (split && CinP % 32 == 0) ? 16 : 0};
  for (int ti = 0; ti < 2 && td == 0; ++ti)
    for (int ni = 0; ni < 2 && td == 0; ++ni)
      for (int cand = 4; cand >= 1 && td == 0; cand /= 2) {
        const int th_c = th_cands[ti], nt_c = nt_cands[ni];
        if (th_c == 0 || nt_c == 0 || H % th_c || cand > D || D % cand) continue;
        const int plane = p.g_chunks * th_c * 128, xb = cand * (th_c + 2) * 128;
        const int g_planes = cand + p.shifts - 1;
        // [gy | x copies | two ones slots]  /  SPLIT: [gy H0 | gy H1 | x H0 | ones | x H1]
        const int stage = (split ? 2 : 1) * g_planes * plane + ((split ? 6 : 3) * (nt_c / 8) + 2) * xb;
        const int over = (split ? g_planes * plane : 0) + (cand - 1) * plane + 16 * th_c * 128 - stage;
        const int slack = over > 0 ? (over + 255) / 128 * 128 : 128;
        int st = (200 * 1024 - slack) / stage;
        if (st > 8) st = 8;
        if (st >= 2) {
          td = cand; stages = st; NT = nt_c; th = th_c;
          p.g_planes = g_planes; p.g_plane_bytes = plane; p.x_chunk_bytes = xb; p.stage_bytes = stage;
          p.slack_bytes = slack;
        }
      }
  p.th = th;
  if (td == 0) return pl;
  p.td = td;
  size_t total = (size_t)stages * p.stage_bytes;
  p.stages = stages;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.Cin = Cin; p.Cout = Cout; p.CCin = sg_chunks(Cin); p.CCout = sg_chunks(Cout); p.CinP = CinP;
  p.tiles_w = W / 8; p.tiles_h = H / th; p.tiles_d = D / td;
  p.n_tiles = p.tiles_w * p.tiles_h * p.tiles_d * N;
  p.ci_tiles = CinP / NT;
  p.tmem_cols = 9 * NT + 16 <= 256 ? 256 : 512;
  const int co_tiles = (CoutP + 127) / 128;
  const int kd_groups = p.shifts == 3 ? 1 : p.shifts == 2 ? 2 : 3;
  const int groups = kd_groups * co_tiles * p.ci_tiles;
  int per_group = sg_num_sms() / groups;   // one CTA per SM (smem-limited): never more than one wave
  if (per_group > p.n_tiles) per_group = p.n_tiles;
  if (per_group < 1) per_group = 1;
  p.direct = per_group == 1;
  pl.grid = dim3((unsigned)per_group, (unsigned)kd_groups, (unsigned)(co_tiles * p.ci_tiles));
  pl.smem = total + p.slack_bytes + 8 * (3 * 8 + 1) + 16;
  pl.ws_bytes = (int64_t)27 * Cout * CinP * (int64_t)sizeof(float);
  pl.NT = NT;
  pl.ok = true;
  return pl;
}

template <int NT, int SPLIT = 0>
int launch_wgrad(const WgPlan& pl, const CUtensorMap& gmap, const CUtensorMap& gmap1, const CUtensorMap& xmap,
                 const CUtensorMap& xmap1, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc<NT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
    if (e != cudaSuccess) {
      sg_set_error("conv_tc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  sg_launch((k_wgrad_tc<NT, SPLIT>), pl.grid, kThreadsW, pl.smem, s, gmap, gmap1, xmap, xmap1, pl.p);
  return sg_check_launch(SPLIT ? "sg_conv3d_wgrad(tcgen05 split-bf16)" : "sg_conv3d_wgrad(tcgen05)");
}

int encode_act_map(CUtensorMap* map, const void* base, int N, int CC, int D, int H, int W, int box_w_vox, int box_h,
                   int box_d, int box_c) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)CC, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                           (cuuint64_t)CC * D * H * W * 16};
  cuuint32_t box[5] = {(cuuint32_t)box_w_vox * 8, (cuuint32_t)box_h, (cuuint32_t)box_d, (cuuint32_t)box_c, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  return 0;
}

}  // namespace

// ws [27][Cout][CinP] -> gw [Cout][Cin][27] * scale (shared with the small-volume fp32 wgrad of conv_direct.cu)
int sg_wgrad_finish(const float* ws, float* gw, int Cout, int Cin, int CinP, float scale, cudaStream_t s) {
  sg_launch((k_wgrad_finish), dim3((unsigned)((Cin + 63) / 64), (unsigned)Cout), 256, 0, s, ws, gw, Cout, Cin, CinP, scale);
  return sg_check_launch("sg_conv3d_wgrad(finish)");
}

int64_t sg_tc_wgrad_workspace_bytes(int N, int Cin, int Cout, int D, int H, int W, int f32) {
  WgPlan pl = make_wgrad_plan(N, Cin, Cout, D, H, W, f32 != 0);
  return pl.ok ? pl.ws_bytes : 0;
}

// returns 1 if the shape is not covered (caller falls through to the direct kernel).  f32 != 0: x and gy are fp32
// acts: 1 = the SPLIT kernel (bf16 hi/lo halves formed in shared memory, three bf16 MMAs per product), 2 = bf16
// operands (hi halves only)
int sg_tc_wgrad(const void* x, const void* gy, float* gw, float* gb, int N, int Cin, int Cout, int D, int H, int W,
                float scale, void* ws, int64_t ws_bytes, cudaStream_t s, int f32) {
  WgPlan pl = make_wgrad_plan(N, Cin, Cout, D, H, W, f32 != 0);
  if (!pl.ok) return 1;
  WgParams& p = pl.p;
  SG_REQUIRE(ws != nullptr && ws_bytes >= pl.ws_bytes, "sg_conv3d_wgrad(tcgen05): workspace too small (%lld < %lld)",
             (long long)ws_bytes, (long long)pl.ws_bytes);
  p.ws = (float*)ws;
  p.gb = gb;
  CUtensorMap gmap, gmap1, xmap, xmap1;
  int rc;
  if (f32) {
    rc = sg_encode_f32_half_map(&gmap, gy, 0, N, p.CCout, D, H, W, 8, p.th, 1, p.g_chunks);
    if (!rc) rc = sg_encode_f32_half_map(&gmap1, gy, 1, N, p.CCout, D, H, W, 8, p.th, 1, p.g_chunks);
    if (!rc) rc = sg_encode_f32_half_map(&xmap, x, 0, N, p.CCin, D, H, W, 8, p.th + 2, p.td, 1);
    if (!rc) rc = sg_encode_f32_half_map(&xmap1, x, 1, N, p.CCin, D, H, W, 8, p.th + 2, p.td, 1);
    if (rc) return rc;
  } else {
    rc = encode_act_map(&gmap, gy, N, p.CCout, D, H, W, 8, p.th, 1, p.g_chunks);
    if (rc) return rc;
    rc = encode_act_map(&xmap, x, N, p.CCin, D, H, W, 8, p.th + 2, p.td, 1);
    if (rc) return rc;
    gmap1 = gmap;
    xmap1 = xmap;
  }
  if (!p.direct) cudaMemsetAsync(ws, 0, (size_t)pl.ws_bytes, s);
  if (gb) cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)Cout, s);
  if (f32 == 2)
    rc = pl.NT == 32 ? launch_wgrad<32, 2>(pl, gmap, gmap1, xmap, xmap1, s) : launch_wgrad<16, 2>(pl, gmap, gmap1, xmap, xmap1, s);
  else if (f32)
    rc = pl.NT == 32 ? launch_wgrad<32, 1>(pl, gmap, gmap1, xmap, xmap1, s) : launch_wgrad<16, 1>(pl, gmap, gmap1, xmap, xmap1, s);
  else
    rc = pl.NT == 32 ? launch_wgrad<32>(pl, gmap, gmap1, xmap, xmap1, s) : launch_wgrad<16>(pl, gmap, gmap1, xmap, xmap1, s);
  if (rc) return rc;
  return sg_wgrad_finish((const float*)ws, gw, Cout, Cin, p.CinP, scale, s);
}
