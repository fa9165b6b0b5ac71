// K4: feature-space nearest neighbours in both directions from ONE pass over the (never materialised)
// N x M distance matrix; exact fp32 direct-difference form sum_c (a_c - b_c)^2 like lib/metrics.py:22-29.
// A CTA owns a 128 x 128 tile (8 x 8 per thread); row minima are reduced with shuffles, column minima through
// shared memory, and merged across tiles with 64-bit atomicMin on (distance bits << 32 | index) so ties go to
// the smallest index (torch.min's first-occurrence rule on CPU).
#include "common.cuh"

namespace gclb {

constexpr int NT = 128;       // tile edge
constexpr int NPAD = NT + 4;
constexpr unsigned long long kNoBest = ~0ull;

__device__ __forceinline__ unsigned long long pack_best(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m);
  unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

template <bool VEC>
__device__ __forceinline__ void load_tile_T(float* S, const float* __restrict__ X, const int64_t* __restrict__ rows,
                                            int64_t row0, int64_t nrows, int C, int cs, int tid) {
  // S[c][r] <- X[row(row0 + r)][cs + c], 128 rows x 32 channels, zero filled outside; row(i) = rows ? rows[i] : i
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    int e = pass * 256 + tid;
    int r = e / 8, cv = (e % 8) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows) {
      const int64_t gr = rows ? __ldg(rows + row0 + r) : row0 + r;
      const float* src = X + (size_t)gr * C + cs + cv;
      if (VEC) {
        if (cs + cv < C) v = __ldg(reinterpret_cast<const float4*>(src));
      } else {
        if (cs + cv + 0 < C) v.x = __ldg(src + 0);
        if (cs + cv + 1 < C) v.y = __ldg(src + 1);
        if (cs + cv + 2 < C) v.z = __ldg(src + 2);
        if (cs + cv + 3 < C) v.w = __ldg(src + 3);
      }
    }
    S[(cv + 0) * NPAD + r] = v.x;
    S[(cv + 1) * NPAD + r] = v.y;
    S[(cv + 2) * NPAD + r] = v.z;
    S[(cv + 3) * NPAD + r] = v.w;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256) nn_tile_kernel(const float* __restrict__ A, const float* __restrict__ B, int C,
                                                      const int64_t* __restrict__ a_ptr,
                                                      const int64_t* __restrict__ b_ptr,
                                                      const int64_t* __restrict__ a_rows,
                                                      const int64_t* __restrict__ b_rows,
                                                      unsigned long long* __restrict__ rowbest,
                                                      unsigned long long* __restrict__ colbest) {
  __shared__ __align__(16) float smem[2 * 32 * NPAD];
  float* As = smem;
  float* Bs = smem + 32 * NPAD;
  const int pair = blockIdx.z;
  const int64_t a0 = a_ptr[pair], b0 = b_ptr[pair];
  const int64_t N = a_ptr[pair + 1] - a0, M = b_ptr[pair + 1] - b0;
  const int64_t tr = (int64_t)blockIdx.y * NT, tc = (int64_t)blockIdx.x * NT;
  if (tr >= N || tc >= M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t nrows = min((int64_t)NT, N - tr), ncols = min((int64_t)NT, M - tc);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int cs = 0; cs < C; cs += 32) {
    load_tile_T<VEC>(As, A, a_rows, a0 + tr, nrows, C, cs, tid);
    load_tile_T<VEC>(Bs, B, b_rows, b0 + tc, ncols, C, cs, tid);
    __syncthreads();
    const int kmax = min(32, C - cs);
#pragma unroll 4
    for (int kk = 0; kk < kmax; ++kk) {
      float4 a0v = *reinterpret_cast<const float4*>(&As[kk * NPAD + ty * 4]);
      float4 a1v = *reinterpret_cast<const float4*>(&As[kk * NPAD + 64 + ty * 4]);
      float4 b0v = *reinterpret_cast<const float4*>(&Bs[kk * NPAD + tx * 4]);
      float4 b1v = *reinterpret_cast<const float4*>(&Bs[kk * NPAD + 64 + tx * 4]);
      float av[8] = {a0v.x, a0v.y, a0v.z, a0v.w, a1v.x, a1v.y, a1v.z, a1v.w};
      float bv[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d = av[i] - bv[j];
          acc[i][j] = fmaf(d, d, acc[i][j]);
        }
    }
    __syncthreads();
  }

  // local row r(i) = (i<4 ? ty*4+i : 64+ty*4+i-4); local col c(j) likewise with tx
  unsigned long long rbest[8], cbest[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rbest[i] = kNoBest;
#pragma unroll
  for (int j = 0; j < 8; ++j) cbest[j] = kNoBest;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
      if (r < nrows && c < ncols) {
        unsigned long long pr = pack_best(acc[i][j], (int)(tc + c));
        unsigned long long pc = pack_best(acc[i][j], (int)(tr + r));
        rbest[i] = min(rbest[i], pr);
        cbest[j] = min(cbest[j], pc);
      }
    }
  }
  // rows: reduce over the 16 tx lanes (a half warp)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long v = rbest[i];
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) v = min(v, shfl_xor_u64(v, m));
    int r = (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
    if (tx == 0 && r < nrows && v != kNoBest) atomicMin(&rowbest[a0 + tr + r], v);
  }
  // columns: reduce over the 16 ty groups through shared memory (aliases the operand tiles; all reads are done)
  if (colbest) {
    unsigned long long* red = reinterpret_cast<unsigned long long*>(smem);  // [16][128]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
      red[ty * NT + c] = cbest[j];
    }
    __syncthreads();
    if (tid < NT && tid < ncols) {
      unsigned long long v = red[tid];
#pragma unroll
      for (int t = 1; t < 16; ++t) v = min(v, red[t * NT + tid]);
      if (v != kNoBest) atomicMin(&colbest[b0 + tc + tid], v);
    }
  }
}

__global__ void __launch_bounds__(256) nn_unpack_kernel(const unsigned long long* __restrict__ best, int64_t n,
                                                        int64_t* __restrict__ idx, float* __restrict__ dist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long v = best[i];
  if (v == kNoBest) {
    idx[i] = -1;
    if (dist) dist[i] = __int_as_float(0x7f800000);
  } else {
    idx[i] = (int64_t)(unsigned)(v & 0xffffffffu);
    if (dist) dist[i] = __uint_as_float((unsigned)(v >> 32));
  }
}

__device__ __forceinline__ int find_segment(const int64_t* ptr, int n_seg, int64_t i) {
  int lo = 0, hi = n_seg;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(&ptr[mid]) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ bool is_mutual(const int64_t* idx01, const int64_t* idx10, const int64_t* a_ptr,
                                          const int64_t* b_ptr, int n_pairs, int64_t i, int64_t* j_out) {
  int s = find_segment(a_ptr, n_pairs, i);
  int64_t j = idx01[i];
  *j_out = j;
  if (j < 0) return false;
  return idx10[b_ptr[s] + j] == i - a_ptr[s];
}

__global__ void __launch_bounds__(kCompactBlock) mutual_count_kernel(const int64_t* idx01, const int64_t* idx10,
                                                                     const int64_t* a_ptr, const int64_t* b_ptr,
                                                                     int n_pairs, int64_t n, int32_t* counts) {
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int64_t j;
  int f = (i < n) && is_mutual(idx01, idx10, a_ptr, b_ptr, n_pairs, i, &j);
  int c = __syncthreads_count(f);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCompactBlock) mutual_scatter_kernel(const int64_t* idx01, const int64_t* idx10,
                                                                       const int64_t* a_ptr, const int64_t* b_ptr,
                                                                       int n_pairs, int64_t n, const int32_t* counts,
                                                                       int64_t* pairs_out, int64_t* pair_ptr) {
  __shared__ int total;
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int64_t j = -1;
  int f = (i < n) && is_mutual(idx01, idx10, a_ptr, b_ptr, n_pairs, i, &j);
  int pos = counts[blockIdx.x] + block_exclusive_scan(f, &total);
  if (i < n) {
    int s = find_segment(a_ptr, n_pairs, i);
    if (f) {
      pairs_out[2 * (int64_t)pos] = i - a_ptr[s];
      pairs_out[2 * (int64_t)pos + 1] = j;
    }
    // segment starts: every (possibly empty) segment beginning at row i gets this position
    if (i == a_ptr[s]) {
      pair_ptr[s] = pos;
      for (int e = s - 1; e >= 0 && a_ptr[e] == i; --e) pair_ptr[e] = pos;   // empty segments before s
    }
    if (i == n - 1) {
      for (int e = s + 1; e <= n_pairs; ++e) pair_ptr[e] = pos + f;          // end + trailing empty segments
    }
  }
}

// ---- per-cloud random subsample without replacement (scripts/test_kitti.py:29-43 `np.random.choice(N, 5000, False)`) ----
// rows of a batched coordinate map are grouped by cloud in order: cloud c owns rows [lower_bound(batch >= c), ...).
__global__ void cloud_ranges_kernel(const int32_t* __restrict__ coords4, const int64_t* __restrict__ n_rows_dev,
                                    int64_t n_rows, int n_clouds, int64_t S, int groups,
                                    int64_t* __restrict__ cloud_ptr, int64_t* __restrict__ sel_ptr) {
  __shared__ int64_t start[1025];
  const int64_t n = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
  for (int c = threadIdx.x; c <= n_clouds; c += blockDim.x) {
    int64_t lo = 0, hi = n;                       // first row with batch >= c
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (__ldg(&coords4[4 * mid]) < c) lo = mid + 1; else hi = mid;
    }
    start[c] = lo;
    cloud_ptr[c] = lo;
  }
  __syncthreads();
  if (threadIdx.x < groups) {   // group g = clouds g, g+groups, ...: its own CSR over the selected rows
    const int g = threadIdx.x, n_seg = n_clouds / groups;
    int64_t* sp = sel_ptr + (int64_t)g * (n_seg + 1);
    int64_t acc = 0;
    sp[0] = 0;
    for (int sgm = 0; sgm < n_seg; ++sgm) {
      int c = sgm * groups + g;
      int64_t v = start[c + 1] - start[c];
      acc += (S > 0 && v > S) ? S : v;
      sp[sgm + 1] = acc;
    }
  }
}

// bijective mixer on `bits` bits (xorshift / odd multiply / add rounds), cycle-walked into [0, v): a pseudo-random
// permutation of the cloud's rows; its first S images are a uniform-looking sample without replacement.
__device__ __forceinline__ uint32_t mix_bits(uint32_t x, int bits, uint32_t k0, uint32_t k1) {
  const uint32_t mask = (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
  const int sh = max(1, bits / 2);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    x = (x + k0) & mask;
    x = (x * 0x9E3779B1u) & mask;
    x ^= x >> sh;
    x = (x * (k1 | 1u)) & mask;
    x ^= x >> sh;
    k0 = k0 * 0x85EBCA6Bu + 0xC2B2AE35u;
    k1 = k1 * 0x27D4EB2Fu + 0x165667B1u;
  }
  return x & mask;
}

__global__ void __launch_bounds__(256) subsample_kernel(const int64_t* __restrict__ cloud_ptr,
                                                        const int64_t* __restrict__ sel_ptr_all, int n_clouds, int64_t S,
                                                        int groups, int64_t group_capacity, uint64_t seed,
                                                        int64_t* __restrict__ sel_all) {
  const int c = blockIdx.y;
  const int g = c % groups, sgm = c / groups, n_seg = n_clouds / groups;
  const int64_t* sel_ptr = sel_ptr_all + (int64_t)g * (n_seg + 1) + sgm;
  int64_t* sel = sel_all + (int64_t)g * group_capacity;
  const int64_t begin = cloud_ptr[c], v = cloud_ptr[c + 1] - begin;
  const int64_t m = sel_ptr[1] - sel_ptr[0];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int64_t pick = i;
  if (m < v) {
    int bits = 1;
    while ((1ll << bits) < v) ++bits;
    uint32_t k0 = (uint32_t)(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(c + 1)));
    uint32_t k1 = (uint32_t)((seed >> 32) + 0x7F4A7C15u * (uint32_t)(c + 1));
    uint32_t x = (uint32_t)i;
    do { x = mix_bits(x, bits, k0, k1); } while ((int64_t)x >= v);
    pick = (int64_t)x;
  }
  sel[sel_ptr[0] + i] = begin + pick;
}

int nn_tc(const float* A, const float* B, int C, const int64_t* a_ptr, const int64_t* b_ptr, int n_pairs,
          const int64_t* a_rows, const int64_t* b_rows, int64_t max_n, int64_t max_m, int64_t* idx01, float* d01,
          int64_t* idx10, float* d10, void* workspace, cudaStream_t st);   // nn_tc.cu
bool nn_tc_supported(int C);
size_t nn_tc_workspace_bytes(int n_pairs, int64_t max_n, int64_t max_m);

}  // namespace gclb

using namespace gclb;

extern "C" {

size_t gclb_nn_workspace_bytes(int64_t n_total, int64_t m_total, int32_t n_pairs, int64_t max_n, int64_t max_m) {
  size_t a = (size_t)(n_total + m_total + 2) * 8;
  size_t b = gclb_compact_workspace_bytes(n_total);
  size_t c = nn_tc_workspace_bytes(n_pairs, max_n, max_m);
  return (a + b > c ? a + b : c) + 64;
}

int gclb_nn(const float* A, const float* B, int32_t C, const int64_t* a_ptr, const int64_t* b_ptr, int32_t n_pairs,
            const int64_t* a_rows, const int64_t* b_rows, int64_t n_total, int64_t m_total, int64_t max_n,
            int64_t max_m, int64_t* idx01, float* d01, int64_t* idx10, float* d10, int32_t algo, void* workspace,
            void* stream) {
  GCLB_CHECK_ARG(a_ptr && b_ptr && workspace && n_pairs >= 1 && C >= 1, "bad arguments");
  GCLB_CHECK_ARG(n_total == 0 || (A && idx01), "null pointer");
  GCLB_CHECK_ARG(max_n <= n_total && max_m <= m_total, "max_n / max_m exceed totals");
  GCLB_CHECK_ARG(n_pairs <= 65535, "too many segments");
  GCLB_CHECK_ARG(algo >= 0 && algo <= 2, "algo must be 0, 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  const bool use_tc = (algo == 2) || (algo == 0 && nn_tc_supported(C));
  if (use_tc) {
    if (!nn_tc_supported(C)) {
      set_error("gclb_nn: C=%d is not covered by the tcgen05 kernel (needs C == 32)", C);
      return GCLB_ERR_UNSUPPORTED;
    }
    if (n_total > 0) cudaMemsetAsync(idx01, 0xff, (size_t)n_total * 8, st);          // rows nobody owns read as -1
    if (idx10 && m_total > 0) cudaMemsetAsync(idx10, 0xff, (size_t)m_total * 8, st);
    if (max_n > 0 || max_m > 0) {
      GCLB_CHECK_ARG(A && B, "null pointer");
      return nn_tc(A, B, C, a_ptr, b_ptr, n_pairs, a_rows, b_rows, max_n, max_m, idx01, d01, idx10, d10, workspace, st);
    }
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  unsigned long long* rowbest = (unsigned long long*)workspace;
  unsigned long long* colbest = idx10 ? rowbest + n_total : nullptr;
  cudaMemsetAsync(rowbest, 0xff, (size_t)(n_total + (idx10 ? m_total : 0)) * 8, st);
  if (max_n > 0 && max_m > 0) {
    GCLB_CHECK_ARG(B != nullptr, "null pointer");
    dim3 grid((unsigned)((max_m + NT - 1) / NT), (unsigned)((max_n + NT - 1) / NT), (unsigned)n_pairs);
    GCLB_CHECK_ARG(grid.y <= 65535, "too many row tiles");
    if (C % 4 == 0) nn_tile_kernel<true><<<grid, 256, 0, st>>>(A, B, C, a_ptr, b_ptr, a_rows, b_rows, rowbest, colbest);
    else nn_tile_kernel<false><<<grid, 256, 0, st>>>(A, B, C, a_ptr, b_ptr, a_rows, b_rows, rowbest, colbest);
    count_launches(1);
  }
  count_launches((n_total > 0) + (idx10 && m_total > 0));
  if (n_total > 0) nn_unpack_kernel<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(rowbest, n_total, idx01, d01);
  if (idx10 && m_total > 0)
    nn_unpack_kernel<<<(unsigned)((m_total + 255) / 256), 256, 0, st>>>(colbest, m_total, idx10, d10);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_subsample(const int32_t* coords4, const int64_t* n_rows_dev, int64_t n_rows, int32_t n_clouds, int64_t S,
                   int32_t groups, uint64_t seed, int64_t* cloud_ptr_out, int64_t* sel_ptr_out, int64_t* sel_out,
                   void* stream) {
  GCLB_CHECK_ARG(cloud_ptr_out && sel_ptr_out && n_clouds >= 1 && n_clouds <= 1024, "bad arguments");
  GCLB_CHECK_ARG(groups >= 1 && groups <= 8 && n_clouds % groups == 0, "n_clouds must be a multiple of groups");
  GCLB_CHECK_ARG(n_rows == 0 || (coords4 && sel_out), "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cloud_ranges_kernel<<<1, 256, 0, st>>>(coords4, n_rows_dev, n_rows, n_clouds, S, groups, cloud_ptr_out, sel_ptr_out);
  int64_t per_cloud = (S > 0 && S < n_rows) ? S : n_rows;
  if (per_cloud > 0) {
    dim3 grid((unsigned)((per_cloud + 255) / 256), (unsigned)n_clouds);
    subsample_kernel<<<grid, 256, 0, st>>>(cloud_ptr_out, sel_ptr_out, n_clouds, S, groups,
                                           per_cloud * (n_clouds / groups), seed, sel_out);
  }
  count_launches(2);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_mutual_filter(const int64_t* idx01, const int64_t* idx10, const int64_t* a_ptr, const int64_t* b_ptr,
                       int32_t n_pairs, int64_t n_total, int64_t* pairs_out, int64_t* pair_ptr, void* workspace,
                       void* stream) {
  GCLB_CHECK_ARG(a_ptr && b_ptr && pair_ptr && workspace && n_pairs >= 1, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_total == 0) {
    cudaMemsetAsync(pair_ptr, 0, (size_t)(n_pairs + 1) * 8, st);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  GCLB_CHECK_ARG(idx01 && idx10 && pairs_out, "null pointer");
  int32_t* counts = (int32_t*)workspace;
  int64_t nb = compact_blocks(n_total);
  mutual_count_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(idx01, idx10, a_ptr, b_ptr, n_pairs, n_total, counts);
  launch_scan_block_counts(counts, nb, nullptr, st);
  mutual_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(idx01, idx10, a_ptr, b_ptr, n_pairs, n_total, counts,
                                                                pairs_out, pair_ptr);
  count_launches(3);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
