// K2: kernel maps in output-stationary form.  One warp per output row, lanes over kernel offsets, so the
// coordinate row is read once (broadcast), the 27 / 125 probes of a row go out together, and the
// nbr[o*K + k] writes are coalesced.  The hash table of a scan (<= a few MB) lives in L2.
#include "common.cuh"

namespace gclb {

__global__ void __launch_bounds__(256) kmap_build_kernel(HashTable t, const int32_t* __restrict__ out_c4,
                                                         int64_t n_out, int ksize, int K, int step, int sign,
                                                         int32_t* __restrict__ nbr, int32_t* __restrict__ pair_count) {
  extern __shared__ int s_count[];  // [K]
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_count[k] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int half = (ksize & 1) ? ksize / 2 : 0;
  for (int64_t o = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); o < n_out;
       o += (int64_t)gridDim.x * warps_per_block) {
    int4 c = __ldg(reinterpret_cast<const int4*>(out_c4) + o);
    for (int k = lane; k < K; k += 32) {
      int ix = k % ksize, iy = (k / ksize) % ksize, iz = k / (ksize * ksize);
      int x = c.y + sign * (ix - half) * step;
      int y = c.z + sign * (iy - half) * step;
      int z = c.w + sign * (iz - half) * step;
      int r = coord_in_range(c.x, x, y, z) ? hash_find(t, pack_key(c.x, x, y, z)) : -1;
      nbr[o * K + k] = r;
      if (r >= 0 && pair_count) atomicAdd(&s_count[k], 1);
    }
  }
  __syncthreads();
  if (pair_count)
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      if (s_count[k]) atomicAdd(&pair_count[k], s_count[k]);
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

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_kmap_build(const void* in_table, int64_t in_capacity, const int32_t* out_coords4, int64_t n_out,
                    int32_t ksize, int32_t offset_stride, int32_t dilation, int32_t sign, int32_t* nbr,
                    int32_t* pair_count, void* stream) {
  GCLB_CHECK_ARG(in_table && (n_out == 0 || (out_coords4 && nbr)), "null pointer");
  GCLB_CHECK_ARG(in_capacity >= 2 && (in_capacity & (in_capacity - 1)) == 0, "bad capacity");
  GCLB_CHECK_ARG(ksize >= 1 && ksize <= 7 && offset_stride >= 1 && dilation >= 1 && (sign == 1 || sign == -1),
                 "bad kernel geometry");
  if (n_out == 0) return GCLB_OK;
  int K = ksize * ksize * ksize;
  int64_t blocks = (n_out + 7) / 8;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;  // grid-stride beyond 16 CTAs/SM
  kmap_build_kernel<<<(unsigned)blocks, 256, K * sizeof(int), (cudaStream_t)stream>>>(
      make_table(in_table, in_capacity), out_coords4, n_out, ksize, K, offset_stride * dilation, sign, nbr, pair_count);
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

}  // extern "C"
