// K1: voxelisation, coordinate hashing, strided coordinate maps.  HBM/L2-bound integer work: one thread per
// row, 64-bit packed keys, open addressing with linear probing, first-occurrence winner via atomicMin on the
// row index, order-preserving compaction (block counts -> scan -> scatter).
#include <stdarg.h>

#include "common.cuh"

namespace gclb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }


// ------------------------------------------------------------------------------------------------------------
// row sources: each yields the (b,x,y,z) a row is keyed on
// ------------------------------------------------------------------------------------------------------------
struct XyzSource {  // float points divided by the voxel size, floor()ed    (sparse_quantize(xyz / voxel))
  const float* xyz;
  const int64_t* cloud_ptr;
  int n_clouds;
  float voxel;
  __device__ __forceinline__ bool get(int64_t i, int& b, int& x, int& y, int& z) const {
    // IEEE round-to-nearest division then floor: bit-identical to torch's fp32 `xyz / voxel` + torch.floor
    float fx = floorf(__fdiv_rn(xyz[3 * i + 0], voxel));
    float fy = floorf(__fdiv_rn(xyz[3 * i + 1], voxel));
    float fz = floorf(__fdiv_rn(xyz[3 * i + 2], voxel));
    int lo = 0, hi = n_clouds;  // cloud_ptr[lo] <= i < cloud_ptr[hi]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (__ldg(&cloud_ptr[mid]) <= i) lo = mid; else hi = mid;
    }
    b = lo;
    const float lim = 131072.f;
    bool ok = fx > -lim && fx < lim && fy > -lim && fy < lim && fz > -lim && fz < lim;  // also rejects NaN/inf
    x = ok ? (int)fx : 0;
    y = ok ? (int)fy : 0;
    z = ok ? (int)fz : 0;
    return ok;
  }
};
struct RowsSource {  // already discrete rows, width 3 (x,y,z) or 4 (b,x,y,z)
  const int32_t* rows;
  int width;
  __device__ __forceinline__ bool get(int64_t i, int& b, int& x, int& y, int& z) const {
    const int32_t* r = rows + i * width;
    if (width == 4) { b = r[0]; x = r[1]; y = r[2]; z = r[3]; }
    else { b = 0; x = r[0]; y = r[1]; z = r[2]; }
    return true;
  }
};
struct StrideSource {  // parent coordinate map rows floored to a coarser lattice
  const int32_t* c4;
  int s;
  __device__ __forceinline__ bool get(int64_t i, int& b, int& x, int& y, int& z) const {
    int4 c = __ldg(reinterpret_cast<const int4*>(c4) + i);
    b = c.x;
    x = floor_div(c.y, s) * s;
    y = floor_div(c.z, s) * s;
    z = floor_div(c.w, s) * s;
    return true;
  }
};

template <class Src>
__global__ void __launch_bounds__(256) insert_rows_kernel(Src src, int64_t n, const int64_t* n_dev, HashTable t,
                                                          int32_t* slot_of, int32_t* status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  if (i >= n) return;
  int b, x, y, z;
  bool ok = src.get(i, b, x, y, z) && coord_in_range(b, x, y, z);
  int slot = -1;
  if (ok) {
    slot = hash_insert_min(t, pack_key(b, x, y, z), (int)i);
    if (slot < 0) atomicOr(status, GCLB_ST_FULL);
  } else {
    atomicOr(status, GCLB_ST_RANGE);
  }
  slot_of[i] = slot;
}

__device__ __forceinline__ bool is_winner(const HashTable& t, const int32_t* slot_of, int64_t i) {
  int s = slot_of[i];
  return s >= 0 && table_val(t, s) == (int)i;
}

__global__ void __launch_bounds__(kCompactBlock) count_winners_kernel(int64_t n, const int64_t* n_dev, HashTable t,
                                                                      const int32_t* slot_of, int32_t* counts) {
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  int f = (i < n) && is_winner(t, slot_of, i);
  int c = __syncthreads_count(f);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) scan_block_counts_kernel(int32_t* counts, int64_t nblocks, int64_t* total_out) {
  __shared__ int total;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < nblocks; base += 1024) {
    int64_t i = base + threadIdx.x;
    int v = (i < nblocks) ? counts[i] : 0;
    int ex = block_exclusive_scan(v, &total);
    int c = carry;
    if (i < nblocks) counts[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[nblocks] = carry;
    if (total_out) *total_out = (int64_t)carry;
  }
}

void launch_scan_block_counts(int32_t* counts, int64_t nblocks, int64_t* total_out, cudaStream_t st) {
  scan_block_counts_kernel<<<1, 1024, 0, st>>>(counts, nblocks, total_out);
}

template <class Src>
__global__ void __launch_bounds__(kCompactBlock) scatter_winners_kernel(Src src, int64_t n, const int64_t* n_dev,
                                                                        HashTable t, const int32_t* slot_of,
                                                                        const int32_t* counts, int32_t* coords4_out,
                                                                        int64_t* unique_map_out) {
  __shared__ int total;
  int64_t i = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  int f = (i < n) && is_winner(t, slot_of, i);
  int pos = counts[blockIdx.x] + block_exclusive_scan(f, &total);
  if (f) {
    int b, x, y, z;
    src.get(i, b, x, y, z);
    reinterpret_cast<int4*>(coords4_out)[pos] = make_int4(b, x, y, z);
    if (unique_map_out) unique_map_out[pos] = i;
    // the table now maps coordinate -> compacted row.  Safe against concurrent is_winner() of duplicates:
    // pos <= i < (any losing row index), so a loser can never read its own index here.
    table_val_ref(t, slot_of[i]) = (unsigned int)pos;
  }
}

__global__ void __launch_bounds__(256) inverse_map_kernel(int64_t n, const int64_t* n_dev, HashTable t,
                                                          const int32_t* slot_of, int32_t* inverse) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, __ldg(n_dev));
  if (i >= n) return;
  int s = slot_of[i];
  inverse[i] = s >= 0 ? table_val(t, s) : -1;
}

template <class Src>
static int dedupe_rows(Src src, int64_t n, const int64_t* n_dev, void* table, int64_t capacity, int tensor_stride,
                       int32_t* coords4_out,
                       int64_t* unique_map_out, int32_t* inverse_out, int64_t* n_out, int32_t* status,
                       void* workspace, cudaStream_t st) {
  if (capacity < 2 || (capacity & (capacity - 1)) != 0 || capacity < n) {
    set_error("hash capacity must be a power of two >= the number of rows");
    return GCLB_ERR_ARG;
  }
  HashTable t = make_table(table, capacity, tensor_stride);
  cudaMemsetAsync(t.slots, 0xff, (size_t)capacity * sizeof(HashSlot), st);
  if (n == 0) {
    cudaMemsetAsync(n_out, 0, 8, st);
    GCLB_CHECK_LAUNCH();
    return GCLB_OK;
  }
  int32_t* slot_of = (int32_t*)workspace;
  int32_t* counts = slot_of + ((n + 3) & ~3ll);
  int64_t nb = compact_blocks(n);
  insert_rows_kernel<Src><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, n, n_dev, t, slot_of, status);
  count_winners_kernel<<<(unsigned)nb, kCompactBlock, 0, st>>>(n, n_dev, t, slot_of, counts);
  scan_block_counts_kernel<<<1, 1024, 0, st>>>(counts, nb, n_out);
  scatter_winners_kernel<Src><<<(unsigned)nb, kCompactBlock, 0, st>>>(src, n, n_dev, t, slot_of, counts, coords4_out,
                                                                      unique_map_out);
  if (inverse_out) inverse_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, n_dev, t, slot_of, inverse_out);
  count_launches(inverse_out ? 5 : 4);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

__global__ void __launch_bounds__(256) hash_build_kernel(const int32_t* c4, int64_t n, HashTable t, int32_t* status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = __ldg(reinterpret_cast<const int4*>(c4) + i);
  const int am = (1 << t.shift) - 1;     // rows of a map at tensor stride s sit on multiples of s
  if (!coord_in_range(c.x, c.y, c.z, c.w) || ((c.y | c.z | c.w) & am)) { atomicOr(status, GCLB_ST_RANGE); return; }
  unsigned int old = kValEmpty;
  int slot = hash_insert_min(t, pack_key(c.x, c.y, c.z, c.w), (int)i, &old);
  if (slot < 0) atomicOr(status, GCLB_ST_FULL);
  else if (old != kValEmpty) atomicOr(status, GCLB_ST_DUPLICATE);
}

__global__ void __launch_bounds__(256) hash_query_kernel(HashTable t, const int32_t* q4, int64_t nq, int32_t* rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int4 c = __ldg(reinterpret_cast<const int4*>(q4) + i);
  rows[i] = coord_in_range(c.x, c.y, c.z, c.w) ? hash_find(t, pack_key(c.x, c.y, c.z, c.w)) : -1;
}

}  // namespace gclb

using namespace gclb;

extern "C" {

const char* gclb_last_error(void) { return g_err; }
int gclb_version(void) { return 100; }
int64_t gclb_kernel_launches(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int64_t gclb_hash_capacity(int64_t n_rows) {
  // capacity counts 32-byte QUAD slots (4 x-consecutive cells each).  2 quads per row: <= 0.5 load even if every row
  // sat in its own quad, ~0.25-0.3 on LiDAR maps (1.8-2 rows share a quad).  Most kernel-map probes are misses, and an
  // unsuccessful linear-probe search costs (1 + 1/(1-a)^2)/2 slots (2.5 at a = 0.5, 1.4 at 0.25); measured: sparser
  // tables beat smaller, denser ones even when the denser table fits L2 better
  int64_t c = 1024;
  while (c < 2 * n_rows) c <<= 1;
  return c;
}
size_t gclb_hash_bytes(int64_t capacity) { return (size_t)capacity * sizeof(HashSlot); }

size_t gclb_compact_workspace_bytes(int64_t n) {
  return (size_t)(((n + 3) & ~3ll) + compact_blocks(n) + 8) * 4;
}

int gclb_hash_build(void* table, int64_t capacity, int32_t tensor_stride, const int32_t* coords4, int64_t n, int32_t* status,
                    void* stream) {
  GCLB_CHECK_ARG(table && status && (n == 0 || coords4), "null pointer");
  GCLB_CHECK_ARG(capacity >= 2 && (capacity & (capacity - 1)) == 0 && capacity >= n, "bad capacity");
  GCLB_CHECK_ARG(tensor_stride >= 1 && (tensor_stride & (tensor_stride - 1)) == 0, "tensor stride must be a power of two");
  cudaStream_t st = (cudaStream_t)stream;
  HashTable t = make_table(table, capacity, tensor_stride);
  cudaMemsetAsync(t.slots, 0xff, (size_t)capacity * sizeof(HashSlot), st);
  if (n > 0) { hash_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(coords4, n, t, status); count_launches(1); }
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_hash_query(const void* table, int64_t capacity, int32_t tensor_stride, const int32_t* q4, int64_t nq,
                    int32_t* rows_out, void* stream) {
  GCLB_CHECK_ARG(table && (nq == 0 || (q4 && rows_out)), "null pointer");
  GCLB_CHECK_ARG(capacity >= 2 && (capacity & (capacity - 1)) == 0, "bad capacity");
  GCLB_CHECK_ARG(tensor_stride >= 1 && (tensor_stride & (tensor_stride - 1)) == 0, "tensor stride must be a power of two");
  if (nq > 0) count_launches(1);
  if (nq > 0)
    hash_query_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(make_table(table, capacity, tensor_stride),
                                                                                     q4, nq, rows_out);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_voxelize(const float* xyz, int64_t P, const int64_t* cloud_ptr, int32_t n_clouds, float voxel, void* table,
                  int64_t capacity, int32_t* coords4_out, int64_t* unique_map_out, int32_t* inverse_map_out,
                  int64_t* n_out, int32_t* status, void* workspace, void* stream) {
  GCLB_CHECK_ARG(table && n_out && status && workspace && cloud_ptr, "null pointer");
  GCLB_CHECK_ARG(P == 0 || (xyz && coords4_out), "null pointer");
  GCLB_CHECK_ARG(voxel > 0.f && n_clouds >= 1 && n_clouds < 1023, "bad voxel size or cloud count");
  GCLB_CHECK_ARG(P < (1ll << 31), "too many points for int32 row indices");
  XyzSource src{xyz, cloud_ptr, n_clouds, voxel};
  return dedupe_rows(src, P, nullptr, table, capacity, 1, coords4_out, unique_map_out, inverse_map_out, n_out, status,
                     workspace, (cudaStream_t)stream);
}

int gclb_quantize_rows(const int32_t* rows, int64_t P, int32_t width, void* table, int64_t capacity,
                       int32_t* coords4_out, int64_t* unique_map_out, int32_t* inverse_map_out, int64_t* n_out,
                       int32_t* status, void* workspace, void* stream) {
  GCLB_CHECK_ARG(table && n_out && status && workspace, "null pointer");
  GCLB_CHECK_ARG(P == 0 || (rows && coords4_out), "null pointer");
  GCLB_CHECK_ARG(width == 3 || width == 4, "width must be 3 or 4");
  GCLB_CHECK_ARG(P < (1ll << 31), "too many rows for int32 row indices");
  RowsSource src{rows, width};
  return dedupe_rows(src, P, nullptr, table, capacity, 1, coords4_out, unique_map_out, inverse_map_out, n_out, status,
                     workspace, (cudaStream_t)stream);
}

int gclb_stride_map(const int32_t* in_coords4, int64_t n_in, const int64_t* n_in_dev, int32_t new_stride,
                    void* out_table, int64_t out_capacity, int32_t* out_coords4, int32_t* parent_row_out, int64_t* n_out,
                    int32_t* status, void* workspace, void* stream) {
  GCLB_CHECK_ARG(out_table && n_out && status && workspace, "null pointer");
  GCLB_CHECK_ARG(n_in == 0 || (in_coords4 && out_coords4), "null pointer");
  GCLB_CHECK_ARG(new_stride >= 1 && (new_stride & (new_stride - 1)) == 0, "tensor stride must be a power of two");
  StrideSource src{in_coords4, new_stride};
  return dedupe_rows(src, n_in, n_in_dev, out_table, out_capacity, new_stride, out_coords4, nullptr, parent_row_out, n_out,
                     status, workspace, (cudaStream_t)stream);
}

}  // extern "C"
