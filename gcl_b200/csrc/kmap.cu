// K2: kernel maps in output-stationary form.  One warp per output row (4 rows per warp for transposed maps), so the
// coordinate row is read once (broadcast) and all probes of a row go out together.  Three probing schemes, chosen by the
// map geometry: quad probes (3x3x3 stepping by one table cell: 18 one-sector probes answer 27 offsets), aligned-candidate
// probes (transposed maps: <= 8 candidates per row) and the generic lane-per-offset scheme.
#include "common.cuh"

namespace gclb {

__global__ void __launch_bounds__(256) kmap_build_kernel(HashTable t, const int32_t* __restrict__ out_c4,
                                                         int64_t n_out, int ksize, int K, int step, int sign,
                                                         int in_stride, int32_t* __restrict__ nbr,
                                                         int32_t* __restrict__ pair_count, uint8_t* __restrict__ row_keys,
                                                         uint32_t* __restrict__ row_masks, int32_t* __restrict__ key_hist,
                                                         int64_t hist_blocks) {
  extern __shared__ int s_count[];  // [K]
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_count[k] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int half = (ksize & 1) ? ksize / 2 : 0;
  const int reach = (ksize - 1) * step;                 // |delta| never exceeds this on any axis
  const int amask = (1 << t.shift) - 1;                 // tensor strides are powers of two
  // the first 32 offsets (all of a 3x3x3 map) never change: keep their deltas in registers, no div/mod per row
  const LaneOffset f0 = lane_offset(lane < K ? lane : 0, ksize, half, sign * step);
  // 3x3x3 map whose offsets step by exactly one cell of the probed table (self maps and coarse-from-fine maps): the three
  // x-neighbours of a column (dy, dz) live in at most two quads, so 18 lanes (2 quads x 9 columns) fetch the whole
  // neighbourhood with 18 one-sector probes instead of 27
  const bool quad_path = (ksize == 3) && (step == (1 << t.shift));
  const int qg = lane / 9, qcol = lane % 9;               // quad 0/1 of the column, column = (dy, dz)
  const int qdy = qcol % 3 - 1, qdz = qcol / 3 - 1;       // actual coordinate deltas in units of `step`
  const int64_t o0 = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t ostep = (int64_t)gridDim.x * warps_per_block;
  int4 c_next = o0 < n_out ? __ldg(reinterpret_cast<const int4*>(out_c4) + o0) : make_int4(0, 0, 0, 0);
  for (int64_t o = o0; o < n_out; o += ostep) {
    const int4 c = c_next;                       // software pipeline: the next row's coordinates are already in flight
    if (o + ostep < n_out) c_next = __ldg(reinterpret_cast<const int4*>(out_c4) + o + ostep);
    // every neighbour of this row keeps its packed fields in range => neighbour key = base key + per-lane constant
    const int lim = kAxisBias - reach;
    const bool safe = (unsigned)c.x < 1023u && c.y >= -lim && c.y < lim && c.z >= -lim && c.z < lim && c.w >= -lim && c.w < lim;
    const uint64_t base = safe ? pack_key(c.x, c.y, c.z, c.w) : 0ull;
    int key = 0;
    unsigned rmask = 0;
    if (quad_path && safe) {
      const int cx = (c.y + kAxisBias) >> t.shift;          // biased cell index along x
      const int g0 = (cx - 1) >> 2;
      unsigned bits = 0, dirs = 0;
      if (lane < 18 && (qg == 0 || ((cx + 1) >> 2) != g0)) {
        const int gx = g0 + qg;
        const uint64_t gkey = ((uint64_t)(unsigned)c.x << 54) | ((uint64_t)(unsigned)gx << 36) |
                              ((uint64_t)(unsigned)(c.z + qdy * step + kAxisBias) << 18) |
                              (uint64_t)(unsigned)(c.w + qdz * step + kAxisBias);
        int v[4];
        quad_find(t, gkey, v);
#pragma unroll
        for (int j = -1; j <= 1; ++j) {
          const int cell = cx + j;
          if ((cell >> 2) != gx) continue;
          const int r = v[cell & 3];
          // table column of the actual delta (j, qdy, qdz): offsets are sign * (i - 1)
          const int ix = sign * j + 1, iy = sign * qdy + 1, iz = sign * qdz + 1;
          const int k = ix + 3 * iy + 9 * iz;
          nbr[o * 27 + k] = r;
          if (r >= 0) {
            bits |= 1u << k;
            dirs |= (ix < 1 ? 1 : 0) | (ix > 1 ? 2 : 0) | (iy < 1 ? 4 : 0) | (iy > 1 ? 8 : 0) | (iz < 1 ? 16 : 0) | (iz > 1 ? 32 : 0);
            if (pair_count) atomicAdd(&s_count[k], 1);
          }
        }
      }
      if (row_masks) rmask = __reduce_or_sync(0xffffffffu, bits);
      if (row_keys) key = (int)__reduce_or_sync(0xffffffffu, dirs);
    } else {
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const LaneOffset f = (k0 == 0) ? f0 : lane_offset(k < K ? k : 0, ksize, half, sign * step);
        int r = -1;
        if (k < K) {
          const int x = c.y + f.dx, y = c.z + f.dy, z = c.w + f.dz;
          // rows of the probed map sit on multiples of ITS tensor stride: a misaligned candidate (19 of the 27 offsets
          // of every fine voxel of a transposed map) cannot exist and is rejected without touching the table
          const bool aligned = ((x | y | z) & amask) == 0;
          if (aligned) {
            if (safe) r = hash_find(t, base + (uint64_t)f.dkey);
            else if (coord_in_range(c.x, x, y, z)) r = hash_find(t, pack_key(c.x, x, y, z));
          }
          nbr[o * K + k] = r;
          if (r >= 0 && pair_count) atomicAdd(&s_count[k], 1);
        }
        if (row_masks && k0 == 0) rmask = __ballot_sync(0xffffffffu, r >= 0);   // populated offsets of this row (K <= 32)
        // 6-bit neighbour-direction key of the row (see gclb_kmap_sort_rows), for free while the row is in registers
        if (row_keys) key |= (int)__reduce_or_sync(0xffffffffu, (unsigned)(r >= 0 ? f.dirbits : 0));
      }
    }
    if (row_keys && lane == 0) {
      row_keys[o] = (uint8_t)key;
      if (key_hist) atomicAdd(&key_hist[(int64_t)key * hist_blocks + (o >> 10)], 1);   // [bucket][1024-row block] counts
    }
    if (row_masks && lane == 0) row_masks[o] = rmask;
  }
  __syncthreads();
  if (pair_count)
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      if (s_count[k]) atomicAdd(&pair_count[k], s_count[k]);
}

// Transposed 3x3x3 maps (out = fine map, probed table = the coarse map at twice the offset step): per axis a fine voxel
// has ONE aligned candidate when its coordinate is even on the coarse lattice (offset 0) and TWO when it is odd (offsets
// +-1), so a row has at most 8 candidates out of 27.  8 lanes per row (one per candidate), 4 rows per warp; the table is
// pre-filled with -1 by a memset and only hits are written.  ~4x fewer warp iterations than the lane-per-offset kernel.
__global__ void __launch_bounds__(256) kmap_build_up_kernel(HashTable t, const int32_t* __restrict__ out_c4, int64_t n_out,
                                                            int step, int32_t* __restrict__ nbr,
                                                            int32_t* __restrict__ pair_count, uint8_t* __restrict__ row_keys,
                                                            uint32_t* __restrict__ row_masks, int32_t* __restrict__ key_hist,
                                                            int64_t hist_blocks) {
  __shared__ int s_count[27];
  if (threadIdx.x < 27) s_count[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const unsigned gmask = 0xffu << (8 * grp);
  const int sh = log2_pow2(step);
  const int64_t rows_per_block = (blockDim.x >> 5) * 4;
  for (int64_t o = (int64_t)blockIdx.x * rows_per_block + (threadIdx.x >> 5) * 4 + grp; o < n_out;
       o += (int64_t)gridDim.x * rows_per_block) {
    const int4 c = __ldg(reinterpret_cast<const int4*>(out_c4) + o);
    const int coord[3] = {c.y, c.z, c.w};
    int cand[3], idx[3];
    bool active = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int odd = (coord[a] >> sh) & 1;            // parity on the coarse lattice (two's complement: floor semantics)
      const int bit = (sub >> a) & 1;
      if (!odd) {                                       // even: the voxel sits on a coarse cell, offset 0 only
        active = active && bit == 0;
        idx[a] = 1;
        cand[a] = coord[a];
      } else {                                          // odd: coarse neighbours at +-step (neighbour = c - (i - 1) * step)
        idx[a] = bit ? 2 : 0;
        cand[a] = coord[a] + (bit ? -step : step);
      }
    }
    int r = -1;
    const int k = idx[0] + 3 * idx[1] + 9 * idx[2];
    if (active && coord_in_range(c.x, cand[0], cand[1], cand[2])) r = hash_find(t, pack_key(c.x, cand[0], cand[1], cand[2]));
    if (r >= 0) {
      nbr[o * 27 + k] = r;
      if (pair_count) atomicAdd(&s_count[k], 1);
    }
    const unsigned dirs = r >= 0 ? (unsigned)((idx[0] < 1 ? 1 : 0) | (idx[0] > 1 ? 2 : 0) | (idx[1] < 1 ? 4 : 0) | (idx[1] > 1 ? 8 : 0) |
                                              (idx[2] < 1 ? 16 : 0) | (idx[2] > 1 ? 32 : 0)) : 0u;
    const unsigned key = __reduce_or_sync(gmask, dirs);
    const unsigned rmask = __reduce_or_sync(gmask, r >= 0 ? (1u << k) : 0u);
    if (sub == 0) {
      if (row_keys) {
        row_keys[o] = (uint8_t)key;
        if (key_hist) atomicAdd(&key_hist[(int64_t)key * hist_blocks + (o >> 10)], 1);
      }
      if (row_masks) row_masks[o] = rmask;
    }
  }
  __syncthreads();
  if (pair_count && threadIdx.x < 27 && s_count[threadIdx.x]) atomicAdd(&pair_count[threadIdx.x], s_count[threadIdx.x]);
}

// pair lists: flat ordered compaction over e = k*n_out + o
__global__ void __launch_bounds__(kCompactBlock) kmap_count_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int K,
                                                                   int32_t* counts) {
  int64_t e = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int f = 0;
  if (e < n_out * K) {
    int64_t k = e / n_out, o = e - k * n_out;
    f = nbr[o * K + k] >= 0;
  }
  int c = __syncthreads_count(f);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(kCompactBlock) kmap_scatter_kernel(const int32_t* __restrict__ nbr, int64_t n_out, int K,
                                                                     const int32_t* counts, int32_t* in_idx,
                                                                     int32_t* out_idx, int64_t* offset_ptr) {
  __shared__ int total;
  int64_t e = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  int f = 0, i = -1;
  int64_t k = 0, o = 0;
  if (e < n_out * K) {
    k = e / n_out;
    o = e - k * n_out;
    i = nbr[o * K + k];
    f = i >= 0;
  }
  int pos = counts[blockIdx.x] + block_exclusive_scan(f, &total);
  if (f) {
    in_idx[pos] = i;
    out_idx[pos] = (int)o;
  }
  if (e < n_out * K && o == 0) offset_ptr[k] = pos;         // first element of column k
  if (e == n_out * K - 1) offset_ptr[K] = pos + f;
}

// ---- row bucketing for the tensor-core kernel -----------------------------------------------------------------------
// A 128-row tile pays one pipeline stage per kernel offset that is populated ANYWHERE in the tile.  Rows are therefore
// grouped (stable counting sort, 64 buckets) by a 6-bit key: "has a neighbour with dx<0 / dx>0 / dy<0 / dy>0 / dz<0 / dz>0".
// Ground-like rows (no vertical neighbours), wall-like rows and, for transposed convolutions, the parity classes of the
// fine voxels end up in separate tiles: measured 21.5 -> 8.6 stages per tile for stride-1 maps, 17.7 -> 2.3 for transposed.
constexpr int kBuckets = 64;

__device__ __forceinline__ int row_key(const int32_t* __restrict__ row, int ksize, int K) {
  int key = 0;
  for (int k = 0; k < K; ++k) {
    if (__ldg(row + k) < 0) continue;
    int ix = k % ksize, iy = (k / ksize) % ksize, iz = k / (ksize * ksize);
    int half = ksize / 2;
    key |= (ix < half) ? 1 : 0;
    key |= (ix > half) ? 2 : 0;
    key |= (iy < half) ? 4 : 0;
    key |= (iy > half) ? 8 : 0;
    key |= (iz < half) ? 16 : 0;
    key |= (iz > half) ? 32 : 0;
  }
  return key;
}

__global__ void __launch_bounds__(kCompactBlock) rowkey_hist_kernel(const int32_t* __restrict__ nbr, int64_t n, int ksize,
                                                                    int K, const uint8_t* __restrict__ keys_in,
                                                                    uint8_t* __restrict__ keys,
                                                                    int32_t* __restrict__ hist, int64_t nblocks) {
  __shared__ int h[kBuckets];
  if (threadIdx.x < kBuckets) h[threadIdx.x] = 0;
  __syncthreads();
  int64_t o = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  if (o < n) {
    int key = keys_in ? (int)keys_in[o] : row_key(nbr + o * K, ksize, K);
    keys[o] = (uint8_t)key;
    atomicAdd(&h[key], 1);
  }
  __syncthreads();
  if (threadIdx.x < kBuckets) hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // bucket-major
}

// Stable scatter with the scan fused in: every block derives its own bucket bases from the raw [bucket][block] histogram
// (16 threads per bucket sum the blocks before it and all blocks), so no separate scan launch is needed.
__global__ void __launch_bounds__(kCompactBlock) rowkey_scatter_kernel(const uint8_t* __restrict__ keys, int64_t n,
                                                                       const int32_t* __restrict__ hist, int64_t nblocks,
                                                                       int32_t* __restrict__ perm,
                                                                       const uint32_t* __restrict__ row_masks,
                                                                       uint32_t* __restrict__ tile_mask) {
  __shared__ int warp_hist[kCompactBlock / 32][kBuckets];
  __shared__ int s_prefix[kBuckets], s_total[kBuckets], s_base[kBuckets];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < (kCompactBlock / 32) * kBuckets; e += kCompactBlock) (&warp_hist[0][0])[e] = 0;
  {
    const int bucket = threadIdx.x >> 4, sub = threadIdx.x & 15;       // 1024 threads = 64 buckets x 16
    int pre = 0, tot = 0;
    for (int64_t blk = sub; blk < nblocks; blk += 16) {
      const int v = __ldg(&hist[(int64_t)bucket * nblocks + blk]);
      tot += v;
      if (blk < blockIdx.x) pre += v;
    }
#pragma unroll
    for (int m = 1; m < 16; m <<= 1) {
      pre += __shfl_xor_sync(0xffffffffu, pre, m);
      tot += __shfl_xor_sync(0xffffffffu, tot, m);
    }
    if (sub == 0) { s_prefix[bucket] = pre; s_total[bucket] = tot; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int bkt = 0; bkt < kBuckets; ++bkt) { s_base[bkt] = run + s_prefix[bkt]; run += s_total[bkt]; }
  }
  int64_t o = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  const bool valid = o < n;
  const int key = valid ? keys[o] : kBuckets;          // invalid lanes form their own group
  const unsigned same = __match_any_sync(0xffffffffu, key);
  const int rank = __popc(same & ((1u << lane) - 1));
  if (valid && rank == 0) warp_hist[wid][key] = __popc(same);
  __syncthreads();
  if (valid) {
    int base = s_base[key];
    for (int w = 0; w < wid; ++w) base += warp_hist[w][key];
    perm[base + rank] = (int32_t)o;                    // stable: buckets keep the original row order
    if (row_masks && tile_mask) {                      // populated offsets of the 128-row tile this row lands in
      const unsigned m = __ldg(row_masks + o);
      if (m) atomicOr(&tile_mask[(base + rank) >> 7], m);
    }
  }
}

// one thread per table ENTRY: coalesced writes, reads in 4*K-byte runs; tile masks through a warp-level OR when the whole
// warp sits in one 128-row tile (the common case: a warp spans ~1.2 rows)
template <int KC>   // KC = K when known at compile time (27), 0 = runtime K
__global__ void __launch_bounds__(256) permute_rows_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                           int64_t n, int K_rt, int32_t* __restrict__ out,
                                                           uint32_t* __restrict__ tile_mask) {
  const int K = KC ? KC : K_rt;
  const int64_t total = n * K;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = e < total;
  int64_t t = 0;
  int k = 0, v = -1;
  if (in) {
    t = e / K;
    k = (int)(e - t * K);
    v = __ldg(&nbr[(int64_t)__ldg(&perm[t]) * K + k]);
    out[e] = v;
  }
  if (tile_mask) {
    const int64_t tile = in ? (t >> 7) : -1;
    const unsigned bit = (in && v >= 0) ? (1u << k) : 0u;
    const int64_t tile0 = __shfl_sync(0xffffffffu, tile, 0);
    if (__all_sync(0xffffffffu, tile == tile0)) {
      const unsigned m = __reduce_or_sync(0xffffffffu, bit);
      if ((threadIdx.x & 31) == 0 && m && tile0 >= 0) atomicOr(&tile_mask[tile0], m);
    } else if (bit) {
      atomicOr(&tile_mask[tile], bit);
    }
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_kmap_build(const void* in_table, int64_t in_capacity, const int32_t* out_coords4, int64_t n_out,
                    int32_t ksize, int32_t offset_stride, int32_t dilation, int32_t sign, int32_t in_tensor_stride,
                    int32_t* nbr, int32_t* pair_count, uint8_t* row_keys, uint32_t* row_masks, int32_t* key_hist,
                    void* stream) {
  GCLB_CHECK_ARG(in_table && (n_out == 0 || (out_coords4 && nbr)), "null pointer");
  GCLB_CHECK_ARG(in_capacity >= 2 && (in_capacity & (in_capacity - 1)) == 0, "bad capacity");
  GCLB_CHECK_ARG(row_masks == nullptr || ksize * ksize * ksize <= 32, "row masks need ksize^3 <= 32");
  GCLB_CHECK_ARG(key_hist == nullptr || row_keys != nullptr, "key_hist needs row_keys");
  GCLB_CHECK_ARG(ksize >= 1 && ksize <= 7 && offset_stride >= 1 && dilation >= 1 && (sign == 1 || sign == -1) &&
                     in_tensor_stride >= 1 && (in_tensor_stride & (in_tensor_stride - 1)) == 0,
                 "bad kernel geometry");
  if (n_out == 0) return GCLB_OK;
  int K = ksize * ksize * ksize;
  const int step = offset_stride * dilation;
  if (ksize == 3 && sign == -1 && in_tensor_stride == 2 * step && (step & (step - 1)) == 0) {
    // transposed map onto the next finer level: <= 8 candidates per row (see kmap_build_up_kernel)
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(nbr, 0xff, (size_t)n_out * 27 * sizeof(int32_t), st);
    int64_t blocks = (n_out + 31) / 32;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    kmap_build_up_kernel<<<(unsigned)blocks, 256, 0, st>>>(make_table(in_table, in_capacity, in_tensor_stride), out_coords4,
                                                           n_out, step, nbr, pair_count, row_keys, row_masks, key_hist,
                                                           compact_blocks(n_out));
    count_launches(1);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  int64_t blocks = (n_out + 7) / 8;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;  // grid-stride beyond 16 CTAs/SM
  kmap_build_kernel<<<(unsigned)blocks, 256, K * sizeof(int), (cudaStream_t)stream>>>(
      make_table(in_table, in_capacity, in_tensor_stride), out_coords4, n_out, ksize, K, offset_stride * dilation, sign, in_tensor_stride,
      nbr, pair_count, row_keys, row_masks, key_hist, compact_blocks(n_out));
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_kmap_pairs(const int32_t* nbr, int64_t n_out, int32_t K, int32_t* in_idx, int32_t* out_idx,
                    int64_t* offset_ptr, void* workspace, void* stream) {
  GCLB_CHECK_ARG(offset_ptr && workspace && K >= 1, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n = n_out * K;
  if (n == 0) {
    cudaMemsetAsync(offset_ptr, 0, (size_t)(K + 1) * 8, st);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  GCLB_CHECK_ARG(nbr && in_idx && out_idx, "null pointer");
  GCLB_CHECK_ARG(n < (1ll << 31), "kernel map too large for int32 positions");
  int32_t* counts = (int32_t*)workspace;
  int64_t nb = compact_blocks(n);
  kmap_count_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(nbr, n_out, K, counts);
  launch_scan_block_counts(counts, nb, nullptr, st);
  kmap_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(nbr, n_out, K, counts, in_idx, out_idx, offset_ptr);
  count_launches(3);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

size_t gclb_kmap_sort_workspace_bytes(int64_t n_out) {
  int64_t nb = compact_blocks(n_out);
  return (size_t)(((n_out + 15) & ~15ll) + (kBuckets * nb + 8) * 4);
}

int gclb_kmap_sort_rows(const int32_t* nbr, int64_t n_out, int32_t ksize, const uint8_t* row_keys,
                        const uint32_t* row_masks, const int32_t* key_hist, int32_t* perm_out, int32_t* nbr_sorted_out,
                        uint32_t* tile_mask_out, void* workspace, void* stream) {
  GCLB_CHECK_ARG(workspace && ksize >= 1 && ksize <= 7, "bad arguments");
  if (n_out == 0) return GCLB_OK;
  GCLB_CHECK_ARG(nbr && perm_out, "null pointer");
  GCLB_CHECK_ARG(nbr_sorted_out || !tile_mask_out || row_masks, "tile masks need either the permuted copy or row_masks");
  GCLB_CHECK_ARG(n_out < (1ll << 31), "too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int K = ksize * ksize * ksize;
  const int64_t nb = compact_blocks(n_out);
  uint8_t* keys = (uint8_t*)workspace;
  int32_t* hist = (int32_t*)(keys + ((n_out + 15) & ~15ll));
  GCLB_CHECK_ARG(key_hist == nullptr || row_keys != nullptr, "key_hist comes with row_keys (both from gclb_kmap_build)");
  int launches = nbr_sorted_out ? 2 : 1;
  if (key_hist) {        // keys and their per-block histogram were produced by gclb_kmap_build: nothing to recompute
    keys = const_cast<uint8_t*>(row_keys);
    hist = const_cast<int32_t*>(key_hist);
  } else {
    rowkey_hist_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(nbr, n_out, ksize, K, row_keys, keys, hist, nb);
    ++launches;
  }
  if (tile_mask_out) {
    GCLB_CHECK_ARG(K <= 32, "tile masks need ksize^3 <= 32");
    cudaMemsetAsync(tile_mask_out, 0, (size_t)((n_out + 127) / 128) * 4, st);
  }
  const bool masks_in_scatter = tile_mask_out && row_masks;
  rowkey_scatter_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(keys, n_out, hist, nb, perm_out,
                                                                masks_in_scatter ? row_masks : nullptr,
                                                                masks_in_scatter ? tile_mask_out : nullptr);
  if (nbr_sorted_out) {   // optional physical copy of the table in sorted order (the conv kernel can also read through perm)
    const int64_t blocks = (n_out * K + 255) / 256;
    GCLB_CHECK_ARG(blocks < (1ll << 31), "kernel map too large");
    uint32_t* tm = masks_in_scatter ? nullptr : tile_mask_out;
    if (K == 27) permute_rows_kernel<27><<<(unsigned)blocks, 256, 0, st>>>(nbr, perm_out, n_out, K, nbr_sorted_out, tm);
    else permute_rows_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(nbr, perm_out, n_out, K, nbr_sorted_out, tm);
  }
  count_launches(launches);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
