// Shared device/host helpers for libgclb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gclb200.h"
#include "../../include/gclb200_debug.h"

namespace gclb {

void set_error(const char* fmt, ...);
void count_launches(int n);   // process-wide tally of kernels enqueued by this library (gclb_kernel_launches)

#define GCLB_CHECK_ARG(cond, msg)                 \
  do {                                            \
    if (!(cond)) {                                \
      gclb::set_error("%s: %s", __func__, msg);   \
      return GCLB_ERR_ARG;                        \
    }                                             \
  } while (0)

#define GCLB_CHECK_LAUNCH()                                                              \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      gclb::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(e__));          \
      return GCLB_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

constexpr int kNumSMs = 148;  // B200

struct ConvParams {   // sparse-conv forward arguments shared by the fp32 (spconv.cu) and tcgen05 (spconv_tc.cu) kernels
  const float* in0; const float* in1;
  int c0, c1;
  const float* W;     // fp32 path: [K][cin][cout] ; tcgen05 path: [K][cout][cin] (gclb_weights_to_tc)
  int K, cout;
  const int32_t* nbr;
  const int32_t* perm;   // tcgen05 path: table row t describes output row perm[t] (NULL = identity)
  const uint32_t* tile_mask;   // tcgen05 path: per 128-row tile, bit k set iff offset k is populated (NULL = scan)
  const float* scale; const float* shift; const float* residual;
  int relu;
  float* out;
  int64_t n_out;
  uint32_t* range_mon = nullptr;   // fp16-range monitor of this launch (gclb_spconv_set_range_monitor) or NULL
};
// range monitor words: [0] flags (bit 0: a value outside the finite fp16 range was produced, NaN included), [1] max |y| as
// float bits.  Kernels that store fp16 fold every value in before the saturating conversion.
constexpr uint32_t kHalfMaxBits = 0x477FE000u;   // 65504.0f
uint32_t* current_range_monitor();                // thread-local (spconv.cu)
__device__ __forceinline__ void range_mon_flush(uint32_t* mon, uint32_t abs_bits_max) {   // one call per warp
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) abs_bits_max = max(abs_bits_max, __shfl_xor_sync(0xffffffffu, abs_bits_max, m));
  if ((threadIdx.x & 31) == 0 && abs_bits_max != 0u) {
    atomicMax(mon + 1, abs_bits_max);
    if (abs_bits_max > kHalfMaxBits) atomicOr(mon, 1u);
  }
}

// ---- packed coordinate keys: [ batch:10 | x:18 | y:18 | z:18 ], biased by 2^17 ---------------------------
constexpr int kAxisBits = 18;
constexpr int kAxisBias = 1 << 17;
constexpr uint64_t kEmptyKey = ~0ull;  // batch 1023 is reserved for it

__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return (unsigned)b < 1023u && (unsigned)(x + kAxisBias) < (1u << kAxisBits) &&
         (unsigned)(y + kAxisBias) < (1u << kAxisBits) && (unsigned)(z + kAxisBias) < (1u << kAxisBits);
}
__host__ __device__ __forceinline__ uint64_t pack_key(int b, int x, int y, int z) {
  return ((uint64_t)(unsigned)b << 54) | ((uint64_t)(unsigned)(x + kAxisBias) << 36) |
         ((uint64_t)(unsigned)(y + kAxisBias) << 18) | (uint64_t)(unsigned)(z + kAxisBias);
}
__device__ __forceinline__ uint32_t hash_key(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

// One 32-byte slot (= one memory sector) holds a QUAD: four cells that are consecutive along x at the table's tensor stride
//   { uint64 group key, uint32 val[0..1] | uint32 val[2..3], 8 bytes unused }        val = row index, ~0 = no such cell
// group key = the packed cell key with its x field replaced by (x_field >> shift) >> 2, sub-cell = (x_field >> shift) & 3,
// shift = log2(tensor stride).  A kernel-map probe of the 3 (5) x-neighbours of a voxel therefore touches 1-2 sectors
// instead of 3 (5) -- the kernel-map and fused-probe kernels are bound by L2 sector traffic, see DESIGN.md -- while
// single-cell insert / find keep their semantics (first-occurrence winner via atomicMin on val[sub]).
// The whole table is initialised by a single memset(0xFF) (empty key = ~0, empty val = UINT_MAX = -1 as a row index).
struct __align__(32) HashSlot {
  unsigned long long key;
  unsigned int val[4];
  unsigned long long pad;
};
struct HashTable {
  HashSlot* slots;
  uint32_t mask;  // number of quad slots - 1
  int shift;      // log2(tensor stride of the coordinates stored)
};
constexpr unsigned int kValEmpty = 0xffffffffu;
constexpr uint64_t kXFieldMask = 0x3FFFFull << 36;
__host__ __device__ __forceinline__ int log2_pow2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}
__host__ __device__ __forceinline__ HashTable make_table(const void* buf, int64_t capacity, int tensor_stride) {
  HashTable t;
  t.slots = (HashSlot*)buf;
  t.mask = (uint32_t)(capacity - 1);
  t.shift = tensor_stride > 1 ? log2_pow2(tensor_stride) : 0;
  return t;
}
// cell key -> (group key, sub-cell)
__device__ __forceinline__ uint64_t group_of(const HashTable& t, uint64_t key, int& sub) {
  const uint32_t cell = (uint32_t)((key >> 36) & 0x3FFFFu) >> t.shift;
  sub = (int)(cell & 3u);
  return (key & ~kXFieldMask) | ((uint64_t)(cell >> 2) << 36);
}
// handle = quad * 4 + sub: what insert returns and the compaction passes store per input row
__device__ __forceinline__ unsigned int& table_val_ref(const HashTable& t, int handle) {
  return t.slots[handle >> 2].val[handle & 3];
}
__device__ __forceinline__ int table_val(const HashTable& t, int handle) { return (int)table_val_ref(t, handle); }

// insert key, value = min(existing, row).  Returns the handle, or -1 when the table is full.  *old_val (optional) receives
// the previous value (kValEmpty when the cell was fresh).
__device__ __forceinline__ int hash_insert_min(const HashTable& t, uint64_t key, int row, unsigned int* old_val = nullptr) {
  int sub;
  const uint64_t g = group_of(t, key, sub);
  uint32_t slot = hash_key(g) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    unsigned long long prev = t.slots[slot].key;
    if (prev == kEmptyKey) prev = atomicCAS(&t.slots[slot].key, kEmptyKey, (unsigned long long)g);
    if (prev == kEmptyKey || prev == g) {
      unsigned int o = atomicMin(&t.slots[slot].val[sub], (unsigned int)row);
      if (old_val) *old_val = o;
      return (int)(slot * 4u + (uint32_t)sub);
    }
    slot = (slot + 1) & t.mask;
  }
  return -1;
}
// the four row indices of a quad (all -1 when the group is absent): ONE sector, two 128-bit loads
__device__ __forceinline__ void quad_find(const HashTable& t, uint64_t gkey, int (&v)[4]) {
  uint32_t slot = hash_key(gkey) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(t.slots + slot));   // key + val[0..1]
    const unsigned long long k = ((unsigned long long)a.y << 32) | a.x;
    if (k == gkey) {
      const uint2 b = __ldg(reinterpret_cast<const uint2*>(t.slots + slot) + 2);   // val[2..3], same sector
      v[0] = (int)a.z; v[1] = (int)a.w; v[2] = (int)b.x; v[3] = (int)b.y;
      return;
    }
    if (k == kEmptyKey) break;
    slot = (slot + 1) & t.mask;
  }
  v[0] = v[1] = v[2] = v[3] = -1;
}
__device__ __forceinline__ int hash_find(const HashTable& t, uint64_t key) {
  int sub;
  const uint64_t g = group_of(t, key, sub);
  uint32_t slot = hash_key(g) & t.mask;
  for (uint32_t probe = 0; probe <= t.mask; ++probe) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(t.slots + slot));
    const unsigned long long k = ((unsigned long long)a.y << 32) | a.x;
    if (k == g) {
      if (sub == 0) return (int)a.z;
      if (sub == 1) return (int)a.w;
      return (int)__ldg(&t.slots[slot].val[sub]);
    }
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & t.mask;
  }
  return -1;
}

// per-offset constants of a lane: coordinate deltas, the same deltas as ONE 64-bit addend of the packed key
// (fields are biased, so base + delta is the packed key of the neighbour as long as no field leaves its 18 bits), and
// the direction bits of the row key
struct LaneOffset {
  int dx, dy, dz;
  long long dkey;
  int dirbits;
};
__device__ __forceinline__ LaneOffset lane_offset(int k, int ksize, int half, int scale) {
  LaneOffset f;
  const int ix = k % ksize, iy = (k / ksize) % ksize, iz = k / (ksize * ksize);
  f.dx = (ix - half) * scale;
  f.dy = (iy - half) * scale;
  f.dz = (iz - half) * scale;
  f.dkey = (long long)f.dx * (1ll << 36) + (long long)f.dy * (1ll << 18) + (long long)f.dz;
  f.dirbits = (ix < half ? 1 : 0) | (ix > half ? 2 : 0) | (iy < half ? 4 : 0) | (iy > half ? 8 : 0) | (iz < half ? 16 : 0) |
              (iz > half ? 32 : 0);
  return f;
}

__device__ __forceinline__ int floor_div(int a, int s) { return (a >= 0) ? a / s : -((-a + s - 1) / s); }

// ---- ordered stream compaction of N flags (3 tiny launches, deterministic) -------------------------------
// workspace: int32 block_counts[nblocks+1]
constexpr int kCompactBlock = 1024;
inline int64_t compact_blocks(int64_t n) { return (n + kCompactBlock - 1) / kCompactBlock; }

// block-wide exclusive scan of one int per thread (blockDim.x == kCompactBlock); returns exclusive prefix,
// total in *total.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += n;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += n;
    }
    warp_sums[lane] = wi - w;  // exclusive
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  int r = warp_sums[wid] + incl - v;
  __syncthreads();
  return r;
}

// scan of per-block counts by one block: counts[i] <- exclusive prefix; counts[nblocks] <- total; also
// optionally stores the total as int64.  (defined in hash.cu)
void launch_scan_block_counts(int32_t* counts, int64_t nblocks, int64_t* total_out, cudaStream_t st);

}  // namespace gclb
