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
//  * weights [tap][C/8][CoutP][8] stream through a ring of bulk-copy stages;
//  * warp roles: warp 0 = TMA/bulk producer, warp 1 = TMEM allocator + single-thread MMA
//    issuer, warps 2-5 = epilogue (tcgen05.ld -> scale, bias, LeakyReLU, mask -> 16-byte stores,
//    or fp32 atomics into the split-K workspace).
#include <cuda.h>

#include "../../include/saragan_b200.h"
#include "common.cuh"

int sg_conv_finish_bf16(const float* acc, const float* bias, const void* mask_src, void* y, int N,
                        int Cout, int64_t V, float scale, int lrelu, cudaStream_t s);

#include "tc_common.cuh"

namespace {

constexpr int kThreads = 192;

struct TcParams {
  const __nv_bfloat16* wp;   // packed weights [27][CCin][CoutP][8]
  const float* bias;         // nullable
  const __nv_bfloat16* mask; // nullable
  __nv_bfloat16* y;          // output act (splits == 1)
  float* ws;                 // fp32 [N*V][CoutP] (splits > 1)
  int N, D, H, W;
  int CCin, Cout, CoutP, CCout;
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
  int tmem_cols;
  float scale;
  int lrelu;
};

// --------------------------------------------------------------------------------- kernel
template <int NT>
__global__ void __launch_bounds__(kThreads)
k_conv_tc(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // carve-up: [A halo block][weight ring][barriers][tmem base]
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.sw * p.w_stage_bytes);
  // bars: [0] full_a, [1] empty_a, [2] acc_full, [3 .. 3+sw) full_w, [3+sw .. 3+2sw) empty_w
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * kMaxSub + 2);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int FULL_A = 0, EMPTY_A = 1, ACC_FULL = 2, FULL_W = 3, EMPTY_W = 3 + p.sw;

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
    mbar_init(BAR(FULL_A), 1);
    mbar_init(BAR(EMPTY_A), 1);
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

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const uint32_t a_addr = smem_u32(a_smem);
      const uint32_t w_addr = smem_u32(w_smem);
      const uint32_t a_tx = (uint32_t)p.kb_chunks * (uint32_t)p.chunk_tx_bytes;
      const uint32_t w_tx = (uint32_t)p.kb_chunks * NT * 16u;
      int it = 0;
      for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
        mbar_wait(BAR(EMPTY_A), (kb & 1) ^ 1);
        mbar_expect_tx(BAR(FULL_A), a_tx);
        const int chunk0 = (kb0 + kb) * p.kb_chunks;
        for (int c = 0; c < p.kb_chunks; ++c)
          tma_load_5d(a_addr + c * p.chunk_bytes, &xmap, BAR(FULL_A), (w0 - 1) * 8, h0 - 1, d0 - 1, chunk0 + c, n0);
        for (int tap = 0; tap < 27; ++tap, ++it) {
          const int s = it % p.sw;
          mbar_wait(BAR(EMPTY_W + s), ((it / p.sw) & 1) ^ 1);
          mbar_expect_tx(BAR(FULL_W + s), w_tx);
          for (int c = 0; c < p.kb_chunks; ++c) {
            const __nv_bfloat16* src = p.wp + (((int64_t)tap * p.CCin + chunk0 + c) * p.CoutP + co0) * 8;
            bulk_load(w_addr + s * p.w_stage_bytes + c * NT * 16, src, NT * 16u, BAR(FULL_W + s));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t leader = elect_one();   // all lanes run the loops; one issues
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = NT, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);
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
      const uint32_t w_stage16 = (uint32_t)p.w_stage_bytes >> 4;
      const int kpairs = p.kb_chunks / 2, n_sub = p.n_sub, halo_w = p.halo_w, halo_h = p.halo_h, sw = p.sw;
      int s = 0, ph = 0;
      for (int kb = 0; kb < p.kblocks_per_split; ++kb) {
        mbar_wait(BAR(FULL_A), kb & 1);
        int tap = 0;
        for (int kd = 0; kd < 3; ++kd)
          for (int kh = 0; kh < 3; ++kh) {
            const int row_off = ((kd - 1) * halo_h + (kh - 1)) * halo_w + bias_vox;
            for (int kw = 0; kw < 3; ++kw, ++tap) {
              mbar_wait(BAR(FULL_W + s), ph);
              tc_fence_after();
              issue_tap<NT>(tmem_base, a_desc0 + (uint64_t)(uint32_t)(row_off + kw), w_desc0 + (uint64_t)(s * w_stage16),
                            sub_off, n_sub, kpairs, kk_a, idesc, (kb | tap) != 0, leader);
              tc_commit(BAR(EMPTY_W + s), leader);
              if (++s == sw) { s = 0; ph ^= 1; }
            }
          }
        tc_commit(BAR(EMPTY_A), leader);
      }
      tc_commit(BAR(ACC_FULL), leader);
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;       // accumulator row == TMEM lane
    const float scale = p.scale;
    const int lrelu = p.lrelu;
    const __nv_bfloat16* mask = p.mask;
    __nv_bfloat16* yout = p.y;
    float* s_bias = reinterpret_cast<float*>(bars + 24);   // NT floats, 16-byte aligned, after the barriers
    for (int i = row; i < NT; i += 128) s_bias[i] = (p.bias && co0 + i < p.Cout) ? p.bias[co0 + i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
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
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float v[16];
        __syncwarp();   // tcgen05.ld is .sync.aligned: the warp must be converged here
        tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * NT + c0), v);
        if (!valid) {
          // row lies on a halo line / column or outside the volume: computed, discarded
        } else if (p.splits > 1) {
          float* dst = p.ws + ((int64_t)n * V + vox) * p.CoutP + co0 + c0;
          // 16 consecutive fp32 of one workspace row: four 16-byte vector reductions instead of 16 scalar ones
#pragma unroll
          for (int q = 0; q < 4; ++q) red_add_v4(dst + 4 * q, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
          const int64_t o = (((int64_t)n * p.CCout + (co0 + c0) / 8) * V + vox) * 8;
          epilogue16(v, s_bias + c0, scale, lrelu, mask ? mask + o : nullptr, yout + o, V * 8);
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
};

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

Plan make_plan(int N, int Cin, int Cout, int D, int H, int W) {
  Plan pl;
  TcParams& p = pl.p;
  if (W % 8 != 0 || H < 8) return pl;
  const int CCin = sg_chunks(Cin), CoutP = 16 * ((Cout + 15) / 16);
  int NT = CoutP % 128 == 0 ? 128 : CoutP % 64 == 0 ? 64 : CoutP % 32 == 0 ? 32 : 16;
  const int max_sub = 256 / NT < kMaxSub ? 256 / NT : kMaxSub;
  int th = H < 16 ? H : 16;
  if (H % th != 0) return pl;
  // largest (tn, td) whose MMA tiles fit the TMEM budget of two co-resident CTAs
  int best_tn = 0, best_td = 0, best_sub = 0, best_lines[kMaxSub];
  for (int td = 1; td <= D && td <= 8; ++td) {
    if (D % td) continue;
    for (int tn = 1; tn <= N && tn <= 8; ++tn) {
      if (td < D && tn > 1) continue;   // span samples only when a tile already holds a whole volume
      int lines[kMaxSub];
      int ns = cover_lines(tn, td, th, td + 2, th + 2, lines);
      if (ns > max_sub) continue;
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
  p.kb_chunks = (CCin % 4 == 0 && 4 * p.chunk_bytes <= 72 * 1024) ? 4 : 2;
  p.a_bytes = p.kb_chunks * p.chunk_bytes;
  // garbage rows of the last MMA tile may read past the block: keep those reads inside the allocation
  int last_line = p.sub_line[p.n_sub - 1] + 15 + p.halo_h + 1;
  int over = (last_line + 1) * p.halo_w * 16 + 8 * 16 + (p.kb_chunks - 1) * p.chunk_bytes - p.a_bytes;
  p.w_stage_bytes = p.kb_chunks * NT * 16;
  p.a_bytes = (p.a_bytes + 127) / 128 * 128;
  int budget = 110 * 1024 - p.a_bytes - 256;
  int sw = budget / p.w_stage_bytes;
  if (sw > 8) sw = 8;
  if (sw < 2) return pl;
  if (over > sw * p.w_stage_bytes) return pl;
  p.sw = sw;
  p.tiles_w = W / 8; p.tiles_h = H / th; p.tiles_d = D / p.td; p.tiles_n = (N + p.tn - 1) / p.tn;
  p.N = N; p.D = D; p.H = H; p.W = W;
  p.CCin = CCin; p.Cout = Cout; p.CoutP = CoutP; p.CCout = sg_chunks(Cout);
  p.tmem_cols = round_pow2_cols(p.n_sub * NT);
  const int n_kblocks = CCin / p.kb_chunks;
  const int64_t ctas = (int64_t)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * (CoutP / NT);
  int splits = 1;
  if (ctas < sg_num_sms()) {
    int want = (int)((2 * sg_num_sms() + ctas - 1) / ctas);
    for (int s = 1; s <= n_kblocks; ++s)
      if (n_kblocks % s == 0 && s <= want) splits = s;
  }
  p.splits = splits;
  p.kblocks_per_split = n_kblocks / splits;
  pl.NT = NT;
  pl.smem = (size_t)p.a_bytes + (size_t)p.sw * p.w_stage_bytes + 8 * 24 + 4 * 128 + 16;
  pl.grid = dim3((unsigned)(p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n), (unsigned)(CoutP / NT), (unsigned)splits);
  pl.ok = true;
  return pl;
}

template <int NT>
int launch(const Plan& pl, const CUtensorMap& map, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_conv_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      sg_set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  k_conv_tc<NT><<<pl.grid, kThreads, pl.smem, s>>>(map, pl.p);
  return sg_check_launch("sg_conv3d_fprop(tcgen05)");
}

}  // namespace

#include "conv_tc_res.cuh"

namespace {
// tests: 0 = auto, 1 = always the streaming kernel, 2 = the weight-resident kernel whenever the
// geometry allows (ignoring the "enough tiles to amortise the weight load" heuristic)
int g_force_streaming = 0;

int encode_halo_map(CUtensorMap* map, const void* x, int N, int CC, int D, int H, int W, int halo_w, int halo_h,
                    int halo_d, int tn) {
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  // activations as a 5-D tensor [W*8 | H | D | CC | N] of bf16; box = one 8-channel halo block
  cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)CC, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                           (cuuint64_t)CC * D * H * W * 16};
  cuuint32_t box[5] = {(cuuint32_t)halo_w * 8, (cuuint32_t)halo_h, (cuuint32_t)halo_d, 1, (cuuint32_t)tn};
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

int64_t sg_tc_wgrad_workspace_bytes(int N, int Cin, int Cout, int D, int H, int W);

int64_t sg_tc_workspace_bytes(int kind, int N, int Cin, int Cout, int D, int H, int W) {
  if (kind == 1) return sg_tc_wgrad_workspace_bytes(N, Cin, Cout, D, H, W);
  if (kind != 0) return 0;
  if (g_force_streaming != 1 && make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2).ok) return 0;
  Plan pl = make_plan(N, Cin, Cout, D, H, W);
  if (!pl.ok || pl.p.splits == 1) return 0;
  return (int64_t)N * D * H * W * pl.p.CoutP * (int64_t)sizeof(float);
}

int sg_tc_fprop(const void* x, const void* wp, const float* bias, const void* mask_src, void* y, int N, int Cin,
                int Cout, int D, int H, int W, float scale, int lrelu, void* ws, int64_t ws_bytes, cudaStream_t s) {
  if (g_force_streaming != 1) {
    ResPlan rp = make_res_plan(N, Cin, Cout, D, H, W, g_force_streaming == 2);
    if (rp.ok) {
      ResParams& q = rp.p;
      q.wp = (const __nv_bfloat16*)wp;
      q.bias = bias;
      q.mask = (const __nv_bfloat16*)mask_src;
      q.y = (__nv_bfloat16*)y;
      q.scale = scale;
      q.lrelu = lrelu;
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
  Plan pl = make_plan(N, Cin, Cout, D, H, W);
  if (!pl.ok) return 1;
  EncodeTiledFn encode = get_encode();
  if (!encode) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled not available");
    return -2;
  }
  TcParams& p = pl.p;
  p.wp = (const __nv_bfloat16*)wp;
  p.bias = bias;
  p.mask = (const __nv_bfloat16*)mask_src;
  p.y = (__nv_bfloat16*)y;
  p.ws = (float*)ws;
  p.scale = scale;
  p.lrelu = lrelu;
  if (p.splits > 1) {
    int64_t need = (int64_t)N * D * H * W * p.CoutP * (int64_t)sizeof(float);
    SG_REQUIRE(ws != nullptr && ws_bytes >= need, "sg_conv3d_fprop(tcgen05): workspace too small (%lld < %lld)",
               (long long)ws_bytes, (long long)need);
    cudaMemsetAsync(ws, 0, (size_t)need, s);
  }
  // activations as a 5-D tensor [W*8 | H | D | CC | N] of bf16; box = one 8-channel halo block
  CUtensorMap map;
  cuuint64_t dims[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)p.CCin, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                           (cuuint64_t)p.CCin * D * H * W * 16};
  cuuint32_t box[5] = {(cuuint32_t)p.halo_w * 8, (cuuint32_t)p.halo_h, (cuuint32_t)p.halo_d, 1, (cuuint32_t)p.tn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    sg_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return -3;
  }
  int rc;
  switch (pl.NT) {
    case 16: rc = launch<16>(pl, map, s); break;
    case 32: rc = launch<32>(pl, map, s); break;
    case 64: rc = launch<64>(pl, map, s); break;
    default: rc = launch<128>(pl, map, s); break;
  }
  if (rc) return rc;
  if (p.splits > 1)
    return sg_conv_finish_bf16((const float*)ws, bias, mask_src, y, N, Cout, (int64_t)D * H * W, scale, lrelu, s);
  return 0;
}

// Introspection for tests / DESIGN.md: the tiling the tcgen05 path would use for a shape.
// out[0..15] = ok, NT, tn, td, th, n_sub, kb_chunks, sw, splits, kblocks_per_split, grid.x, grid.y,
//              grid.z, smem bytes, tmem columns, a_bytes
extern "C" int sg_tc_plan_debug(int N, int Cin, int Cout, int D, int H, int W, int* out) {
  Plan pl = make_plan(N, Cin, Cout, D, H, W);
  const TcParams& p = pl.p;
  int v[16] = {pl.ok, pl.NT, p.tn, p.td, p.th, p.n_sub, p.kb_chunks, p.sw, p.splits, p.kblocks_per_split,
               (int)pl.grid.x, (int)pl.grid.y, (int)pl.grid.z, (int)pl.smem, p.tmem_cols, p.a_bytes};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return 0;
}
