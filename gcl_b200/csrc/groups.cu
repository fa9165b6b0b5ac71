// Positive-group construction for GCL training: util/pointcloud.py:69-132 get_matching_indices_colocation.
// For every point of the centre cloud: radius search (nearest first, at most K) in the centre cloud itself and in each
// neighbour cloud brought into the centre frame; the neighbour lists are appended with running index offsets; the "finest"
// member is the first neighbour hit of the cloud whose nearest hit lies closest to ITS sensor; centres without any
// neighbour-cloud hit are dropped.  The reference does this with 1 + J Open3D KD-trees in a Python loop over points.
//
// Here the clouds' own voxel hashes (K1: the loader's clouds are one point per voxel, so table row == point index) replace
// the KD-trees: a ball of radius r around q touches at most (2r/v + 2)^3 voxels, each holding at most one point, and the
// quad slots answer four x-consecutive voxels per 32-byte probe.  For neighbour cloud j the QUERY is moved into the cloud's
// sensor frame (q' = T_j^-1 p) to find the candidate voxels; the accept test and the ordering then use the reference's own
// quantity, |T_j x - p|^2 < r^2 in float64.  One warp per centre point: lanes over (quad, y, z) probes, candidates collected
// in shared memory, K rounds of warp arg-min.  Variable-length output by size -> scan -> scatter (ordered, deterministic).
#include "common.cuh"

namespace gclb {

constexpr int kGroupWarps = 8;
constexpr int kCandCap = 128;       // candidates within the ball, per (point, cloud)
constexpr int kMaxSpan = 8;         // voxels per axis a ball may span

struct GroupParams {
  const float* c_xyz; int64_t n_center; HashTable c_table;
  const float* nb_xyz; const int64_t* nb_ptr; HashTable nb_table;
  const double* trans;       // [J][16] cloud j -> centre frame
  const double* inv_trans;   // [J][16] centre frame -> cloud j
  int J;
  double voxel, radius;
  int K, kcap;               // K <= 0: unlimited (up to kcap)
  int32_t* tmp_list;         // [n_center][(1+J)*kcap]
  int32_t* tmp_size;         // [n_center] 0 = dropped
  int32_t* tmp_finest;       // [n_center]
  int32_t* status;
};

__device__ __forceinline__ void warp_argmin(double& d, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (od < d || (od == d && oi < i)) { d = od; i = oi; }
  }
}

// radius search of `p` (centre frame, float64) in one cloud: returns the number of hits written to out[] (nearest first,
// ties by smallest index, at most kmax), identical in every lane.
__device__ int ball_query(const HashTable& t, int batch, const float* __restrict__ xyz, int64_t row0, const double* T /* or null */,
                          const double q[3] /* query in the cloud's frame */, const double p[3], double voxel, double r2,
                          double radius, int kmax, bool unlimited, double* s_d, int* s_i, int* s_n, int32_t* out,
                          int32_t* status) {
  const int lane = threadIdx.x & 31;
  int lo[3], span[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double m = voxel * (1e-4 + fabs(q[a] / voxel) * 2.4e-7);     // fp32 voxelisation of the stored points vs fp64 here
    lo[a] = (int)floor((q[a] - radius - m) / voxel);
    span[a] = (int)floor((q[a] + radius + m) / voxel) - lo[a] + 1;
  }
  if (span[0] > kMaxSpan || span[1] > kMaxSpan || span[2] > kMaxSpan) {
    if (lane == 0) atomicOr(status, GCLB_ST_RANGE);
    return 0;
  }
  if (lane == 0) *s_n = 0;
  __syncwarp();
  const int cx0 = lo[0] + kAxisBias;                       // biased cell index (tensor stride 1 tables)
  const int g0 = cx0 >> 2, ng = ((cx0 + span[0] - 1) >> 2) - g0 + 1;
  const int items = ng * span[1] * span[2];
  for (int it = lane; it < items; it += 32) {
    const int g = it % ng, yy = (it / ng) % span[1], zz = it / (ng * span[1]);
    const int y = lo[1] + yy, z = lo[2] + zz;
    if (!coord_in_range(batch, lo[0], y, z) || !coord_in_range(batch, lo[0] + span[0] - 1, y, z)) continue;
    const uint64_t gkey = ((uint64_t)(unsigned)batch << 54) | ((uint64_t)(unsigned)(g0 + g) << 36) |
                          ((uint64_t)(unsigned)(y + kAxisBias) << 18) | (uint64_t)(unsigned)(z + kAxisBias);
    int v[4];
    quad_find(t, gkey, v);
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int cell = ((g0 + g) << 2) + sub;
      if (v[sub] < 0 || cell < cx0 || cell >= cx0 + span[0]) continue;
      const float* px = xyz + (int64_t)v[sub] * 3;          // table rows are global over the concatenated clouds
      double x[3] = {(double)__ldg(px), (double)__ldg(px + 1), (double)__ldg(px + 2)};
      if (T) {   // open3d PointCloud.transform: x' = R x + t in float64
        const double a0 = T[0] * x[0] + T[1] * x[1] + T[2] * x[2] + T[3];
        const double a1 = T[4] * x[0] + T[5] * x[1] + T[6] * x[2] + T[7];
        const double a2 = T[8] * x[0] + T[9] * x[1] + T[10] * x[2] + T[11];
        x[0] = a0; x[1] = a1; x[2] = a2;
      }
      const double d0 = x[0] - p[0], d1 = x[1] - p[1], d2 = x[2] - p[2];
      // (d0^2 + d1^2) + d2^2 with every operation rounded separately (no FMA contraction): the order of near-tied
      // neighbours then follows plain float64 arithmetic, like the oracle's
      const double dd = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
      if (dd < r2) {                                       // nanoflann RadiusResultSet: strict
        const int pos = atomicAdd(s_n, 1);
        if (pos < kCandCap) { s_d[pos] = dd; s_i[pos] = (int)(v[sub] - row0); }     // index local to the cloud
      }
    }
  }
  __syncwarp();
  int n = *s_n;
  if (n > kCandCap) {
    if (lane == 0) atomicOr(status, GCLB_ST_FULL);
    n = kCandCap;
  }
  if (unlimited && n > kmax && lane == 0) atomicOr(status, GCLB_ST_FULL);   // "every hit" was asked for but does not fit
  const int take = n < kmax ? n : kmax;
  for (int k = 0; k < take; ++k) {     // selection sort by warp arg-min: (distance, index) ascending
    double bd = 1e300;
    int bi = 0x7fffffff, bpos = -1;
    for (int e = lane; e < n; e += 32) {
      const double d = s_d[e];
      const int i = s_i[e];
      if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; bpos = e; }
    }
    double wd = bd;
    int wi = bi;
    warp_argmin(wd, wi);
    if (bpos >= 0 && bd == wd && bi == wi) s_d[bpos] = 1e301;     // taken (indices are unique within a cloud)
    if (lane == 0) out[k] = wi;
    __syncwarp();
  }
  return take;
}

__global__ void __launch_bounds__(kGroupWarps * 32) colocation_search_kernel(GroupParams g) {
  __shared__ double s_d[kGroupWarps][kCandCap];
  __shared__ int s_i[kGroupWarps][kCandCap];
  __shared__ int s_n[kGroupWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = (1 + g.J) * g.kcap;
  const int kmax = g.K > 0 ? g.K : g.kcap;
  const double r2 = g.radius * g.radius;
  for (int64_t i = (int64_t)blockIdx.x * kGroupWarps + w; i < g.n_center; i += (int64_t)gridDim.x * kGroupWarps) {
    const double p[3] = {(double)__ldg(g.c_xyz + 3 * i), (double)__ldg(g.c_xyz + 3 * i + 1), (double)__ldg(g.c_xyz + 3 * i + 2)};
    int32_t* list = g.tmp_list + i * L;
    double closest = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);           // :97 own sensor distance
    int len = ball_query(g.c_table, 0, g.c_xyz, 0, nullptr, p, p, g.voxel, r2, g.radius, kmax, g.K <= 0, s_d[w], s_i[w], &s_n[w],
                         list, g.status);
    const int n_own = len;
    int finest = 0;
    int64_t start = g.n_center;
    for (int j = 0; j < g.J; ++j) {
      const double* Ti = g.inv_trans + 16 * j;
      const double q[3] = {Ti[0] * p[0] + Ti[1] * p[1] + Ti[2] * p[2] + Ti[3], Ti[4] * p[0] + Ti[5] * p[1] + Ti[6] * p[2] + Ti[7],
                           Ti[8] * p[0] + Ti[9] * p[1] + Ti[10] * p[2] + Ti[11]};
      const int64_t row0 = __ldg(g.nb_ptr + j), nj = __ldg(g.nb_ptr + j + 1) - row0;
      __syncwarp();
      const int got = ball_query(g.nb_table, j, g.nb_xyz, row0, g.trans + 16 * j, q, p, g.voxel, r2, g.radius, kmax, g.K <= 0,
                                 s_d[w], s_i[w], &s_n[w], list + len, g.status);
      __syncwarp();
      if (got > 0) {
        const int first = list[len];
        __syncwarp();                                                        // everybody has read it before it is offset
        const float* x = g.nb_xyz + (row0 + first) * 3;                      // :113 float32 norm of the un-transformed point
        const float x0 = __ldg(x), x1 = __ldg(x + 1), x2 = __ldg(x + 2);
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
        if ((double)nrm < closest) {
          closest = (double)nrm;
          finest = len;
        }
        if (lane < got) list[len + lane] += (int32_t)start;                  // :117 running offset (got <= kcap <= 32)
        len += got;
      }
      start += nj;
    }
    if (lane == 0) {
      g.tmp_size[i] = (len == n_own) ? 0 : len;                               // :121 no neighbour-cloud hit => dropped
      g.tmp_finest[i] = finest;
    }
    __syncwarp();
  }
}

// per 1024-point block: total list length and number of kept groups
__global__ void __launch_bounds__(kCompactBlock) group_block_sums_kernel(const int32_t* __restrict__ size, int64_t n,
                                                                         int32_t* sum_len, int32_t* sum_cnt) {
  __shared__ int total;
  const int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  const int s = i < n ? size[i] : 0;
  block_exclusive_scan(s, &total);
  const int c = __syncthreads_count(s > 0);
  if (threadIdx.x == 0) {
    sum_len[blockIdx.x] = total;
    sum_cnt[blockIdx.x] = c;
  }
}

__global__ void __launch_bounds__(kCompactBlock) group_scatter_kernel(const int32_t* __restrict__ size,
                                                                      const int32_t* __restrict__ finest,
                                                                      const int32_t* __restrict__ list, int L, int64_t n,
                                                                      const int32_t* __restrict__ off_len,
                                                                      const int32_t* __restrict__ off_cnt, int64_t* group_out,
                                                                      int64_t* index_out, uint8_t* finest_out) {
  __shared__ int total;
  const int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  const int s = i < n ? size[i] : 0;
  const int pos = off_len[blockIdx.x] + block_exclusive_scan(s, &total);
  const int rank = off_cnt[blockIdx.x] + block_exclusive_scan(s > 0 ? 1 : 0, &total);
  if (s > 0) {
    group_out[rank] = s;
    const int f = finest[i];
    for (int e = 0; e < s; ++e) {
      index_out[pos + e] = list[i * L + e];
      finest_out[pos + e] = (e == f);
    }
  }
}

// ---- _exhaustive_hash (util/misc.py:29-36): symmetric keys min(a + b*M, a*M + b) of every unordered pair inside a group,
// in the reference's order (group by group, i ascending, then the members after i).  Thread per group: pair counts ->
// ordered scan -> write.
__global__ void __launch_bounds__(kCompactBlock) pair_block_sums_kernel(const int64_t* __restrict__ group_ptr, int64_t G,
                                                                        int32_t* sums) {
  __shared__ int total;
  const int64_t g = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int c = 0;
  if (g < G) {
    const int64_t n = group_ptr[g + 1] - group_ptr[g];
    c = (int)(n * (n - 1) / 2);
  }
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kCompactBlock) pair_hash_kernel(const int64_t* __restrict__ group_ptr,
                                                                  const int64_t* __restrict__ index, int64_t G, int64_t M,
                                                                  const int32_t* __restrict__ block_off, int64_t* out) {
  __shared__ int total;
  const int64_t g = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int64_t b = 0, n = 0;
  if (g < G) { b = group_ptr[g]; n = group_ptr[g + 1] - b; }
  int64_t pos = block_off[blockIdx.x] + block_exclusive_scan((int)(n * (n - 1) / 2), &total);
  for (int64_t i = 0; i + 1 < n; ++i) {
    const int64_t a = index[b + i];
    for (int64_t j = i + 1; j < n; ++j) {
      const int64_t c = index[b + j];
      const int64_t k1 = a + c * M, k2 = a * M + c;
      out[pos++] = k1 < k2 ? k1 : k2;
    }
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" {

size_t gclb_groups_workspace_bytes(int64_t n_center, int32_t n_clouds, int32_t kcap) {
  const int64_t L = (int64_t)(1 + n_clouds) * kcap;
  const int64_t nb = compact_blocks(n_center);
  return (size_t)(n_center * L + 2 * n_center + 2 * (nb + 1) + 8) * 4;
}

int gclb_colocation_groups(const float* center_xyz, int64_t n_center, const void* center_table, int64_t center_capacity,
                           const float* nb_xyz, const int64_t* nb_ptr, const void* nb_table, int64_t nb_capacity,
                           const double* trans, const double* inv_trans, int32_t n_clouds, float voxel, double radius,
                           int32_t K, int32_t kcap, int64_t* group_out, int64_t* index_out, uint8_t* finest_out,
                           int64_t* n_groups_out, int64_t* n_index_out, int32_t* status, void* workspace, void* stream) {
  GCLB_CHECK_ARG(center_table && nb_table && nb_ptr && trans && inv_trans && status && workspace && n_groups_out && n_index_out,
                 "null pointer");
  GCLB_CHECK_ARG(n_center == 0 || (center_xyz && group_out && index_out && finest_out), "null pointer");
  GCLB_CHECK_ARG(n_clouds >= 1 && n_clouds < 1023 && voxel > 0.f && radius > 0.0, "bad arguments");
  GCLB_CHECK_ARG(kcap >= 1 && kcap <= 32 && K <= kcap, "K must be <= kcap <= 32 (K <= 0: every hit, up to kcap)");
  GCLB_CHECK_ARG(2.0 * radius / voxel + 2.0 <= kMaxSpan, "search radius too large for the voxel size (ball spans > 8 voxels)");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_center == 0) {
    cudaMemsetAsync(n_groups_out, 0, 8, st);
    cudaMemsetAsync(n_index_out, 0, 8, st);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  const int L = (1 + n_clouds) * kcap;
  const int64_t nb = compact_blocks(n_center);
  int32_t* tmp_list = (int32_t*)workspace;
  int32_t* tmp_size = tmp_list + n_center * L;
  int32_t* tmp_finest = tmp_size + n_center;
  int32_t* sum_len = tmp_finest + n_center;
  int32_t* sum_cnt = sum_len + (nb + 1);
  GroupParams g;
  g.c_xyz = center_xyz; g.n_center = n_center; g.c_table = make_table(center_table, center_capacity, 1);
  g.nb_xyz = nb_xyz; g.nb_ptr = nb_ptr; g.nb_table = make_table(nb_table, nb_capacity, 1);
  g.trans = trans; g.inv_trans = inv_trans; g.J = n_clouds;
  g.voxel = (double)voxel; g.radius = radius; g.K = K; g.kcap = kcap;
  g.tmp_list = tmp_list; g.tmp_size = tmp_size; g.tmp_finest = tmp_finest; g.status = status;
  int64_t blocks = (n_center + kGroupWarps - 1) / kGroupWarps;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  colocation_search_kernel<<<(unsigned)blocks, kGroupWarps * 32, 0, st>>>(g);
  group_block_sums_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(tmp_size, n_center, sum_len, sum_cnt);
  launch_scan_block_counts(sum_len, nb, n_index_out, st);
  launch_scan_block_counts(sum_cnt, nb, n_groups_out, st);
  group_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(tmp_size, tmp_finest, tmp_list, L, n_center, sum_len, sum_cnt,
                                                               group_out, index_out, finest_out);
  count_launches(5);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

size_t gclb_exhaustive_hash_workspace_bytes(int64_t n_groups) { return (size_t)(compact_blocks(n_groups) + 8) * 4; }

int gclb_exhaustive_hash(const int64_t* group_ptr, const int64_t* index, int64_t n_groups, int64_t M, int64_t* keys_out,
                         int64_t* n_keys_out, void* workspace, void* stream) {
  GCLB_CHECK_ARG(n_keys_out && workspace && (n_groups == 0 || (group_ptr && index && keys_out)), "null pointer");
  GCLB_CHECK_ARG(M >= 1, "M must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_groups == 0) {
    cudaMemsetAsync(n_keys_out, 0, 8, st);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  const int64_t nb = compact_blocks(n_groups);
  int32_t* sums = (int32_t*)workspace;
  pair_block_sums_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(group_ptr, n_groups, sums);
  launch_scan_block_counts(sums, nb, n_keys_out, st);
  pair_hash_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(group_ptr, index, n_groups, M, sums, keys_out);
  count_launches(3);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
