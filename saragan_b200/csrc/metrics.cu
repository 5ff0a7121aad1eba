// Evaluation metrics of the training loop on the GPU: the 3-D sliced Wasserstein distance over a
// Laplacian pyramid (pgan_pytorch/metrics/swd.py) and the Kolmogorov-Smirnov histogram distance
// (pgan_pytorch/metrics/kms.py), which train.py:12-27,76,99,120 evaluates on the host with numpy/scipy after every
// stabilising epoch.  All of it is HBM-bound gather / stencil / histogram work plus one skinny projection:
//
//   k_pyr_down / k_pyr_up_sub   5x5x5 binomial stencil with scipy's 'mirror' boundary (swd.py:55-78); sums in
//                               fp64 like scipy.ndimage, rounded to fp32 once
//   k_swd_descriptors           gather of the 3x9x9 neighbourhoods + per-position standardisation (swd.py:8-32)
//   k_swd_project               P[r][c] = sum_k A[r][k] * dirs[k][c]  (swd.py:43-44), rows = real and fake images
//   k_swd_finish                per direction: sort the two arms, mean |difference| (swd.py:44-47)
//   k_value_hist                integer-value histogram of (x*i + i).astype(int).clip(lo, hi) (kms.py:6-10)
//
// What the reference's broadcasting really computes (oracle/metrics_oracle.py header, pinned against the
// reference): the descriptor array has shape (N, N, 3, 9, 9) with N = 128*batch; axis 0 only selects the IMAGE, axis
// 1 runs over the N random positions.  A row of the projected matrix therefore depends on its image only, so the
// projection needs batch (not N) rows per arm, and the sorted columns are `batch` values each repeated 128 times:
// the mean absolute difference over the N rows equals the one over the `batch` sorted values.  Component k of a
// row is ((j*3 + dz)*9 + dx)*9 + dy  (position j; dx walks the LAST array axis, dy the one before).
#include "../../include/saragan_b200.h"
#include "common.cuh"

namespace {

constexpr int kDirs = 128;          // swd.py:108 dirs_per_repeat
constexpr int kPatch = 3 * 9 * 9;   // swd.py:99 nhood_size (1, 2, 8, 8) -> (2*1+1) x (2*4+1) x (2*4+1)

// scipy.ndimage mode='mirror': reflect about the centre of the edge sample (d c b | a b c d | c b a)
__device__ __forceinline__ int mirror(int i, int n) {
  if ((unsigned)i < (unsigned)n) return i;       // interior taps: no integer division
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

__constant__ float c_binom[5] = {1.f / 16, 4.f / 16, 6.f / 16, 4.f / 16, 1.f / 16};

// swd.py:61-63: y = convolve(x, G, 'mirror')[::2, ::2, ::2]; one thread per output voxel
__global__ void k_pyr_down(const float* __restrict__ x, float* __restrict__ y, int64_t P, int D, int H, int W, int oD,
                           int oH, int oW) {
  sg_pdl_enter();
  const int64_t total = P * oD * oH * oW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % oW), oh = (int)((i / oW) % oH), od = (int)((i / ((int64_t)oW * oH)) % oD);
    const int64_t p = i / ((int64_t)oW * oH * oD);
    const float* xp = x + p * (int64_t)D * H * W;
    int zw[5], zh[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      zw[t] = mirror(2 * ow + t - 2, W);
      zh[t] = mirror(2 * oh + t - 2, H);
    }
    double acc = 0.0;
    for (int a = 0; a < 5; ++a) {
      const float* xd = xp + (int64_t)mirror(2 * od + a - 2, D) * H * W;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const float wab = c_binom[a] * c_binom[b];
#pragma unroll
        for (int c = 0; c < 5; ++c) acc += (double)(wab * c_binom[c]) * (double)xd[(int64_t)zh[b] * W + zw[c]];
      }
    }
    y[i] = (float)acc;
  }
}

// swd.py:65-78: lap = fine - convolve(zero_insert(coarse), 4*G, 'mirror'); fine is (2cD, 2cH, 2cW)
__global__ void k_pyr_up_sub(const float* __restrict__ fine, const float* __restrict__ coarse, float* __restrict__ lap,
                             int64_t P, int cD, int cH, int cW) {
  sg_pdl_enter();
  const int D = 2 * cD, H = 2 * cH, W = 2 * cW;
  const int64_t total = P * D * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / ((int64_t)W * H)) % D);
    const int64_t p = i / ((int64_t)W * H * D);
    const float* cp = coarse + p * (int64_t)cD * cH * cW;
    int zd[5], zh[5], zw[5];                       // -1: the tap falls on a zero-inserted sample
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int md = mirror(d + t - 2, D), mh = mirror(h + t - 2, H), mw = mirror(w + t - 2, W);
      zd[t] = (md & 1) ? -1 : md >> 1;
      zh[t] = (mh & 1) ? -1 : mh >> 1;
      zw[t] = (mw & 1) ? -1 : mw >> 1;
    }
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      if (zd[a] < 0) continue;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if (zh[b] < 0) continue;
        const float wab = c_binom[a] * c_binom[b] * 4.0f;
        const float* row = cp + ((int64_t)zd[a] * cH + zh[b]) * cW;
#pragma unroll
        for (int c = 0; c < 5; ++c)
          if (zw[c] >= 0) acc += (double)(wab * c_binom[c]) * (double)row[zw[c]];
      }
    }
    lap[i] = fine[i] - (float)acc;
  }
}

// swd.py:8-32 for one position j per block: gather the B x 3x9x9 values, centre and scale them by their
// mean / standard deviation over (images, patch), write row-major a[b][j*243 + e]
__global__ void k_swd_descriptors(const float* __restrict__ level, const int* __restrict__ pz, const int* __restrict__ py,
                                  const int* __restrict__ px, float* __restrict__ a, int B, int D, int H, int W, int N,
                                  int64_t row_stride) {
  sg_pdl_enter();
  extern __shared__ float sv[];            // B * 243 gathered values
  __shared__ double red[32];
  __shared__ double bc;
  const int j = blockIdx.x;
  const int z0 = pz[j], y0 = py[j], x0 = px[j];
  const int n = B * kPatch;
  auto block_sum = [&](double v) -> double {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
      bc = s;
    }
    __syncthreads();
    return bc;
  };
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / kPatch, e = i % kPatch;
    const int dz = e / 81, dx = (e / 9) % 9, dy = e % 9;     // 4th descriptor axis walks the LAST array axis
    const float v = level[(((int64_t)b * D + (z0 + dz - 1)) * H + (y0 + dy - 4)) * W + (x0 + dx - 4)];
    sv[i] = v;
    s += v;
  }
  const float mean = (float)(block_sum(s) / n);
  s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float c = sv[i] - mean;
    sv[i] = c;
    s += c;
  }
  const double m2 = block_sum(s) / n;        // np.std re-centres the (already centred) values
  s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double c = (double)sv[i] - m2;
    s += c * c;
  }
  const float sd = (float)sqrt(block_sum(s) / n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / kPatch, e = i % kPatch;
    a[(int64_t)b * row_stride + (int64_t)j * kPatch + e] = sv[i] / sd;    // 0/0 = NaN for a constant patch, as numpy
  }
}

// P[r][c] += sum_{k in slice} A[r][k] * dirs[k][c];  colsq[c] += sum_k dirs[k][c]^2.
// Block = 128 columns x KY k-lanes over one slice of K.  dirs dominates the traffic (K x 512 bytes against R x K x 4)
// and is read straight from global memory, each element once, 512 contiguous bytes per row; the A tile [R][KT] goes
// through shared memory and is broadcast to the 128 columns.
constexpr int kProjKY = 4, kProjKT = 64;
template <int ROWS>      // ROWS in {4, 8, 16}: rows of A per pass, a multiple of 4 (float4 shared-memory reads)
__global__ void __launch_bounds__(kDirs* kProjKY)
    k_swd_project(const float* __restrict__ a, const float* __restrict__ dirs, float* __restrict__ p,
                  float* __restrict__ colsq, int R, int64_t K, int64_t row_stride) {
  sg_pdl_enter();
  // A tile transposed to [k][row]: a thread reads the ROWS values of one k as ROWS/4 broadcast float4 loads (the
  // first version read them as 16 scalar loads per k and was bound by shared-memory wavefronts: ncu 8.3 M per
  // launch, l1tex 77 % busy at 19 % of the HBM peak).  The grid is ONE wave of resident blocks striding over the
  // 64-row tiles of dirs (v2 launched 1.09 waves), and a thread has 8 independent loads of dirs in flight.
  __shared__ __align__(16) float as[kProjKT][ROWS];
  __shared__ float part[kProjKY][ROWS + 1][kDirs];
  const int c = threadIdx.x, ky = threadIdx.y;
  const int tid = ky * kDirs + c;
  const int64_t n_tiles = (K + kProjKT - 1) / kProjKT;
  float acc[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
  float sq = 0.f;
  auto accumulate = [&](int k, float d) {
    sq = fmaf(d, d, sq);
#pragma unroll
    for (int q = 0; q < ROWS / 4; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(&as[k][4 * q]);
      acc[4 * q + 0] = fmaf(v.x, d, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(v.y, d, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(v.z, d, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(v.w, d, acc[4 * q + 3]);
    }
  };
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int64_t kt = t * kProjKT;
    const int len = (int)min((int64_t)kProjKT, K - kt);
    const float* dp = dirs + kt * kDirs + c;
    float d0[8];
    if (len == kProjKT) {                      // issue the first batch of loads before waiting on the A tile
#pragma unroll
      for (int j = 0; j < 8; ++j) d0[j] = dp[(int64_t)(ky + kProjKY * j) * kDirs];
    }
    __syncthreads();
    for (int i = tid; i < ROWS * kProjKT; i += kDirs * kProjKY) {
      const int r = i / kProjKT, k = i % kProjKT;          // consecutive threads read consecutive k of one row
      as[k][r] = (r < R && k < len) ? a[(int64_t)r * row_stride + kt + k] : 0.f;
    }
    __syncthreads();
    if (len == kProjKT) {
      float d1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) d1[j] = dp[(int64_t)(ky + kProjKY * (8 + j)) * kDirs];
#pragma unroll
      for (int j = 0; j < 8; ++j) accumulate(ky + kProjKY * j, d0[j]);
#pragma unroll
      for (int j = 0; j < 8; ++j) accumulate(ky + kProjKY * (8 + j), d1[j]);
    } else {
      for (int k = ky; k < len; k += kProjKY) accumulate(k, dp[(int64_t)k * kDirs]);
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) part[ky][r][c] = acc[r];
  part[ky][ROWS][c] = sq;
  __syncthreads();
  if (ky == 0) {
    for (int r = 0; r <= ROWS; ++r) {
      if (r < R || r == ROWS) {
        float v = 0.f;
#pragma unroll
        for (int y = 0; y < kProjKY; ++y) v += part[y][r][c];
        if (r < ROWS) atomicAdd(p + (int64_t)r * kDirs + c, v);
        else if (colsq) atomicAdd(colsq + c, v);
      }
    }
  }
}

template <int ROWS>
static unsigned project_grid(int64_t K) {
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_swd_project<ROWS>, kDirs * kProjKY, 0);
    if (per_sm < 1) per_sm = 1;
  }
  const int64_t tiles = (K + kProjKT - 1) / kProjKT;
  const int64_t wave = (int64_t)sg_num_sms() * per_sm;
  return (unsigned)(tiles < wave ? tiles : wave);
}

// swd.py:44-47 for one repeat: per direction sort the B projections of each arm, sum |real - fake|;
// out[0] = mean over (B, 128).  p is [2B][128] (real rows first); colsq (nullable) holds the squared column norms of
// un-normalised directions (swd.py:42 divides the directions by them).
constexpr int kMaxBatch = 64;
__global__ void k_swd_finish(const float* __restrict__ p, const float* __restrict__ colsq, float* __restrict__ out, int B) {
  sg_pdl_enter();
  __shared__ float red[kDirs / 32];
  const int c = threadIdx.x;
  const float inv = colsq ? rsqrtf(colsq[c]) : 1.f;
  float ra[kMaxBatch], rb[kMaxBatch];
  for (int b = 0; b < B; ++b) {
    ra[b] = p[(int64_t)b * kDirs + c] * inv;
    rb[b] = p[(int64_t)(B + b) * kDirs + c] * inv;
  }
  for (int i = 1; i < B; ++i) {              // insertion sorts (B is the per-GPU batch)
    float va = ra[i], vb = rb[i];
    int ja = i - 1, jb = i - 1;
    while (ja >= 0 && ra[ja] > va) { ra[ja + 1] = ra[ja]; --ja; }
    ra[ja + 1] = va;
    while (jb >= 0 && rb[jb] > vb) { rb[jb + 1] = rb[jb]; --jb; }
    rb[jb + 1] = vb;
  }
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += fabsf(ra[b] - rb[b]);
  s = warp_sum(s);
  if ((c & 31) == 0) red[c >> 5] = s;
  __syncthreads();
  if (c == 0) {
    float t = 0.f;
    for (int i = 0; i < kDirs / 32; ++i) t += red[i];
    out[0] = t / (float)(B * kDirs);
  }
}

// kms.py:6-10: q = clip(int((x * i) + i), lo, hi) (two fp32 roundings, truncation towards zero), counted per
// image in hist[n][hi - lo + 1]
__global__ void k_value_hist(const float* __restrict__ x, int* __restrict__ hist, int64_t V, float intercept, int lo, int hi) {
  sg_pdl_enter();
  extern __shared__ int sh[];
  const int bins = hi - lo + 1;
  for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float* xi = x + (int64_t)blockIdx.y * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __fadd_rn(__fmul_rn(xi[i], intercept), intercept);
    long long q = (long long)v;
    q = q < lo ? lo : (q > hi ? hi : q);
    atomicAdd(&sh[(int)q - lo], 1);
  }
  __syncthreads();
  int* h = hist + (int64_t)blockIdx.y * bins;
  for (int i = threadIdx.x; i < bins; i += blockDim.x)
    if (sh[i]) atomicAdd(h + i, sh[i]);
}

}  // namespace

extern "C" int sg_pyr_down(const float* x, float* y, int64_t P, int D, int H, int W, cudaStream_t s) {
  SG_REQUIRE(P >= 0 && D > 0 && H > 0 && W > 0, "sg_pyr_down: bad shape");
  const int oD = (D + 1) / 2, oH = (H + 1) / 2, oW = (W + 1) / 2;
  const int64_t total = P * oD * oH * oW;
  if (total == 0) return 0;
  sg_launch((k_pyr_down), sg_grid(total, 256), 256, 0, s, x, y, P, D, H, W, oD, oH, oW);
  return sg_check_launch("sg_pyr_down");
}

extern "C" int sg_pyr_up_sub(const float* fine, const float* coarse, float* lap, int64_t P, int cD, int cH, int cW,
                             cudaStream_t s) {
  SG_REQUIRE(P >= 0 && cD > 0 && cH > 0 && cW > 0, "sg_pyr_up_sub: bad shape");
  const int64_t total = P * 8 * cD * cH * cW;
  if (total == 0) return 0;
  sg_launch((k_pyr_up_sub), sg_grid(total, 256), 256, 0, s, fine, coarse, lap, P, cD, cH, cW);
  return sg_check_launch("sg_pyr_up_sub");
}

extern "C" int sg_swd_descriptors(const float* level, const int* pos_z, const int* pos_y, const int* pos_x, float* a,
                                  int B, int D, int H, int W, int N, int64_t row_stride, cudaStream_t s) {
  SG_REQUIRE(B > 0 && N > 0, "sg_swd_descriptors: empty batch");
  SG_REQUIRE(D >= 3 && H >= 9 && W >= 9, "sg_swd_descriptors: volume %dx%dx%d smaller than the 3x9x9 neighbourhood", D, H, W);
  SG_REQUIRE(row_stride >= (int64_t)N * kPatch, "sg_swd_descriptors: row stride too small");
  const size_t smem = (size_t)B * kPatch * sizeof(float);
  SG_REQUIRE(smem <= 200 * 1024, "sg_swd_descriptors: batch %d too large", B);
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaFuncSetAttribute(k_swd_descriptors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = smem;
  }
  sg_launch((k_swd_descriptors), (unsigned)N, 256, smem, s, level, pos_z, pos_y, pos_x, a, B, D, H, W, N, row_stride);
  return sg_check_launch("sg_swd_descriptors");
}

extern "C" int sg_swd_project(const float* a, const float* dirs, float* p, float* colsq, int R, int64_t K,
                              int64_t row_stride, cudaStream_t s) {
  SG_REQUIRE(R > 0 && K > 0 && row_stride >= K, "sg_swd_project: bad shape");
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)R * kDirs * sizeof(float), s);
  if (e == cudaSuccess && colsq) e = cudaMemsetAsync(colsq, 0, kDirs * sizeof(float), s);
  if (e != cudaSuccess) {
    sg_set_error("sg_swd_project: %s", cudaGetErrorString(e));
    return (int)e;
  }
  for (int r0 = 0; r0 < R;) {     // up to 16 rows per pass; only the first pass accumulates the column norms
    const int left = R - r0;
    const int rows = left > 8 ? (left < 16 ? left : 16) : left;
    const float* ap = a + (int64_t)r0 * row_stride;
    float* pp = p + (int64_t)r0 * kDirs;
    float* cs = r0 == 0 ? colsq : (float*)nullptr;
    const dim3 block(kDirs, kProjKY);
    if (rows > 8) sg_launch((k_swd_project<16>), project_grid<16>(K), block, 0, s, ap, dirs, pp, cs, rows, K, row_stride);
    else if (rows > 4) sg_launch((k_swd_project<8>), project_grid<8>(K), block, 0, s, ap, dirs, pp, cs, rows, K, row_stride);
    else sg_launch((k_swd_project<4>), project_grid<4>(K), block, 0, s, ap, dirs, pp, cs, rows, K, row_stride);
    int rc = sg_check_launch("sg_swd_project");
    if (rc) return rc;
    r0 += rows;
  }
  return 0;
}

extern "C" int sg_swd_finish(const float* p, const float* colsq, float* out, int B, cudaStream_t s) {
  SG_REQUIRE(B > 0 && B <= kMaxBatch, "sg_swd_finish: batch %d outside 1..%d", B, kMaxBatch);
  sg_launch((k_swd_finish), 1, kDirs, 0, s, p, colsq, out, B);
  return sg_check_launch("sg_swd_finish");
}

extern "C" int sg_value_hist(const float* x, int* hist, int N, int64_t V, float intercept, int lo, int hi, cudaStream_t s) {
  SG_REQUIRE(hi >= lo && hi - lo < 12000, "sg_value_hist: %d value bins do not fit shared memory", hi - lo + 1);
  if (N == 0 || V == 0) return 0;
  const int bins = hi - lo + 1;
  cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)N * bins * sizeof(int), s);
  if (e != cudaSuccess) {
    sg_set_error("sg_value_hist: %s", cudaGetErrorString(e));
    return (int)e;
  }
  unsigned bx = sg_grid(V, 256 * 8);
  const unsigned cap = (unsigned)((sg_num_sms() * 4 + N - 1) / N);
  if (bx > cap) bx = cap;
  sg_launch((k_value_hist), dim3(bx, (unsigned)N), 256, (size_t)bins * sizeof(int), s, x, hist, V, intercept, lo, hi);
  return sg_check_launch("sg_value_hist");
}
