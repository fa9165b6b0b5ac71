// K4 on the tensor cores: nearest neighbour in 32-d feature space as a distance GEMM with a fused arg-min.
//
//   d2(i, j) = |a_i|^2 + |b_j|^2 - 2 a_i . b_j ,   a_i . b_j on tcgen05 (kind::tf32, fp32 accumulate in TMEM)
//
// Precision: the reference computes sum_c (a_c - b_c)^2 in fp32 (lib/metrics.py:22-29) and takes arg-min indices, so a plain
// tf32 product (10-bit mantissa) would flip near-ties.  Every operand is split into TWO tf32-representable parts
//   a = a1 + a2 (+ e),  a1 = the top 11 significant bits (exact), a2 = the residual rounded to nearest tf32, |e| <= 2^-24 |a|
// and the three product groups (1,1) (1,2) (2,1) -- every partial product exact in fp32 -- are accumulated in TMEM.  What is
// dropped ((2,2) and e) is <= 2^-22 of |a||b| per term: for unit descriptors the dot product is off by <= 3e-7 in the worst
// case (~4e-8 typical), the size of the fp32 rounding of the reference's own direct-difference sum; indices can differ from the
// oracle only where best and second-best distance are closer than 1e-6 (the documented near-tie rule, tests/test_gpu_parity.py).
// Round 1 used a three-way split with six product groups: twice the tensor-core instructions for bits below fp32 noise.
//
// Issue-rate aware shape (tools/umma_rate.py: one thread issues a tcgen05.mma every ~100 cycles whatever N is): a candidate tile
// is 256 rows, so an MMA (M = 128 queries, N = 256 candidates, K = 8) occupies the tensor pipe for 128 cycles and the pipe,
// not the issuing thread, is the limit: 12 instructions per 128 x 256 distances (round 1: 48).
//
// Structure: two row-stationary passes (queries = A over candidates B, then B over A): a CTA owns 128 query rows, keeps their two
// part-images resident in shared memory and streams 256-row candidate super-blocks (two pre-swizzled part images + norms, ONE
// cp.async.bulk each) through a 2-stage ring into a double-buffered TMEM accumulator (2 x 256 columns); four epilogue warps
// (thread <-> query row) read the 256 dot products back and keep the running (min, first index) in registers.  No atomics, no
// N x M matrix, bit-reproducible.
#include "tc_common.cuh"

namespace gclb {

constexpr int NN_C = 32;                       // channels: one 128-byte swizzle row
constexpr int NN_SB = 256;                     // rows per super-block (candidate tile)
constexpr int PART_BYTES = NN_SB * 128;        // 32 KB: one part image of a super-block
constexpr int BLOCK_IMG = 2 * PART_BYTES + 1024;   // two parts + 256 norms
constexpr int Q_PART = TM * 128;               // 16 KB: one part image of the 128 query rows
constexpr int Q_IMG = 2 * Q_PART + 1024;       // query block in shared memory (norms padded, keeps 1024-byte alignment)
constexpr int NN_STAGES = 2;
constexpr int kNnThreads = 6 * 32;             // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

// ---- split + swizzle: feature rows -> per-256-row super-block images ------------------------------------------------
// grid (superblocks_per_seg, n_pairs); block 256 threads = 32 rows x 8 chunks per pass
__global__ void __launch_bounds__(256) nn_split_kernel(const float* __restrict__ X, const int64_t* __restrict__ rows,
                                                       const int64_t* __restrict__ ptr, int blocks_per_seg,
                                                       unsigned char* __restrict__ img) {
  const int p = blockIdx.y, blk = blockIdx.x;
  const int64_t seg0 = ptr[p], n = ptr[p + 1] - seg0;
  if ((int64_t)blk * NN_SB >= n && blk > 0) return;   // nothing of this segment lives here (block 0 is always written)
  unsigned char* out = img + ((size_t)p * blocks_per_seg + blk) * BLOCK_IMG;
  float* norms = reinterpret_cast<float*>(out + 2 * PART_BYTES);
  const int j = threadIdx.x & 7;
  for (int r = threadIdx.x >> 3; r < NN_SB; r += 32) {
    const int64_t local = (int64_t)blk * NN_SB + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool valid = local < n;
    if (valid) {
      const int64_t gr = rows ? __ldg(rows + seg0 + local) : seg0 + local;
      v = __ldg(reinterpret_cast<const float4*>(X + (size_t)gr * NN_C) + j);
    }
    float x[4] = {v.x, v.y, v.z, v.w}, p1[4], p2[4];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      p1[q] = __uint_as_float(__float_as_uint(x[q]) & 0xFFFFE000u);        // top 11 significant bits
      const float r1 = x[q] - p1[q];                                        // exact
      const uint32_t u = __float_as_uint(r1);
      p2[q] = __uint_as_float((u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);   // residual, round to nearest even at 11 bits
      ss = fmaf(x[q], x[q], ss);
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(out + off) = make_float4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<float4*>(out + PART_BYTES + off) = make_float4(p2[0], p2[1], p2[2], p2[3]);
    if (j == 0) norms[r] = valid ? ss : __int_as_float(0x7f800000);        // padding rows can never win
  }
}

struct NnShared {
  uint64_t q_full;
  uint64_t full[NN_STAGES], empty[NN_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// grid (128-row query blocks, n_pairs, directions); blocks_a / blocks_b = super-blocks per segment
__global__ void __launch_bounds__(kNnThreads, 1) nn_tc_kernel(const unsigned char* __restrict__ imgA,
                                                              const unsigned char* __restrict__ imgB,
                                                              const int64_t* __restrict__ a_ptr,
                                                              const int64_t* __restrict__ b_ptr, int blocks_a,
                                                              int blocks_b, int64_t* __restrict__ idx01,
                                                              float* __restrict__ d01, int64_t* __restrict__ idx10,
                                                              float* __restrict__ d10) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ NnShared sh;
  const int p = blockIdx.y, dir = blockIdx.z, qblk = blockIdx.x;
  const unsigned char* q_img = dir == 0 ? imgA : imgB;
  const unsigned char* c_img = dir == 0 ? imgB : imgA;
  const int64_t* q_ptr = dir == 0 ? a_ptr : b_ptr;
  const int64_t* c_ptr = dir == 0 ? b_ptr : a_ptr;
  const int q_blocks = dir == 0 ? blocks_a : blocks_b, c_blocks = dir == 0 ? blocks_b : blocks_a;
  int64_t* idx_out = dir == 0 ? idx01 : idx10;
  float* d_out = dir == 0 ? d01 : d10;
  const int64_t q0 = q_ptr[p], nq = q_ptr[p + 1] - q0, nc = c_ptr[p + 1] - c_ptr[p];
  if (qblk >= 2 * q_blocks || (int64_t)qblk * TM >= nq) return;        // whole CTA: no query rows here
  const int n_tiles = (int)((nc + NN_SB - 1) / NN_SB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* q_s = smem_dyn;                                       // 2 parts + norms of the query block
  unsigned char* ring = smem_dyn + Q_IMG;

  if (tid == 0) {
    if ((smem_u32(smem_dyn) & 1023u) != 0) { printf("gclb nn_tc: shared memory not 1024-byte aligned\n"); __trap(); }
    mbar_init(&sh.q_full, 1);
    for (int s = 0; s < NN_STAGES; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1 + 4); }   // MMA commit + 4 epilogue warps
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const uint32_t q_u32 = smem_u32(q_s), ring_u32 = smem_u32(ring);

  if (warp == 0) {
    // ================= producer: the query half-block (3 pieces of its super-block), then one bulk copy per candidate tile =======
    if (lane == 0) {
      const unsigned char* qsb = q_img + ((size_t)p * q_blocks + (qblk >> 1)) * BLOCK_IMG;
      const int half = qblk & 1;
      mbar_arrive_expect_tx(&sh.q_full, 2 * Q_PART + 512);
      bulk_g2s(q_u32, qsb + half * Q_PART, Q_PART, &sh.q_full);
      bulk_g2s(q_u32 + Q_PART, qsb + PART_BYTES + half * Q_PART, Q_PART, &sh.q_full);
      bulk_g2s(q_u32 + 2 * Q_PART, qsb + 2 * PART_BYTES + half * 512, 512, &sh.q_full);
      for (int t = 0; t < n_tiles; ++t) {
        const int stage = t % NN_STAGES;
        mbar_wait(&sh.empty[stage], ((t / NN_STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&sh.full[stage], BLOCK_IMG);
        bulk_g2s(ring_u32 + stage * BLOCK_IMG, c_img + ((size_t)p * c_blocks + t) * BLOCK_IMG, BLOCK_IMG, &sh.full[stage]);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(NN_SB);      // N = 256
      mbar_wait(&sh.q_full, 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int stage = t % NN_STAGES, b = t & 1;
        mbar_wait(&sh.acc_empty[b], ((t >> 1) & 1) ^ 1);
        mbar_wait(&sh.full[stage], (t / NN_STAGES) & 1);
        tc_fence_after();
        const uint32_t c_s = ring_u32 + stage * BLOCK_IMG;
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * NN_SB);
        // smallest terms first: (2,1) (1,2) (1,1)
        const int qa[3] = {1, 0, 0}, cb[3] = {0, 1, 0};
#pragma unroll
        for (int g = 0; g < 3; ++g) {
#pragma unroll
          for (int ks = 0; ks < NN_C / 8; ++ks)
            umma_tf32(d_tmem, make_desc_sw128(q_u32 + qa[g] * Q_PART + ks * 32),
                      make_desc_sw128(c_s + cb[g] * PART_BYTES + ks * 32), idesc, (g | ks) ? 1u : 0u);
        }
        umma_commit(&sh.acc_full[b]);       // accumulator ready for the epilogue
        umma_commit(&sh.empty[stage]);      // MMA no longer reads the stage (the epilogue warps release it too: norms)
      }
    }
  } else {
    // ================= epilogue: thread <-> query row, running arg-min over all candidate tiles =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const float* qn = reinterpret_cast<const float*>(q_s + 2 * Q_PART);
    mbar_wait(&sh.q_full, 0);
    const float na = qn[row];
    float best = __int_as_float(0x7f800000);
    int best_j = -1;
    for (int t = 0; t < n_tiles; ++t) {
      const int stage = t % NN_STAGES, b = t & 1;
      mbar_wait(&sh.full[stage], (t / NN_STAGES) & 1);      // the tile's norms (bulk-copied with the operands) are visible
      mbar_wait(&sh.acc_full[b], (t >> 1) & 1);
      tc_fence_after();
      // (copying the norms out to release the stage before the epilogue was measured slower: 374 vs 346 us -- the extra
      // barrier among the epilogue warps costs more than the exposed refill)
      const float* cn = reinterpret_cast<const float*>(ring + stage * BLOCK_IMG + 2 * PART_BYTES);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * NN_SB);
      // two register buffers: the tcgen05.ld of chunk i + 1 is in flight while chunk i is reduced (TMEM reads, 64 B/clk per SM,
      // are this kernel's floor: every dot product crosses that port once per direction)
      uint32_t va[32], vb[32];
      auto consume = [&](const uint32_t* v, int n0) {
        // a serial `if (d < best)` over the 256 candidates of a tile is a 256-long dependency chain (measured: 3.9k cycles per
        // tile, 2.5x the MMAs).  Instead: the 32 distances of a chunk are independent FMAs, their minimum a 5-deep FMNMX tree,
        // and the index search runs only when the chunk improves the running best (rarely after the first tiles).
        float d[32];
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cn + n0 + q);
          d[q] = fmaf(-2.f, __uint_as_float(v[q]), c4.x);                      // |b|^2 - 2 a.b   (+|a|^2 at the end)
          d[q + 1] = fmaf(-2.f, __uint_as_float(v[q + 1]), c4.y);
          d[q + 2] = fmaf(-2.f, __uint_as_float(v[q + 2]), c4.z);
          d[q + 3] = fmaf(-2.f, __uint_as_float(v[q + 3]), c4.w);
        }
        float m[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) m[q] = fminf(d[q], d[q + 16]);
#pragma unroll
        for (int w = 8; w > 0; w >>= 1)
#pragma unroll
          for (int q = 0; q < w; ++q) m[q] = fminf(m[q], m[q + w]);
        if (m[0] < best) {                                                     // strict: an earlier chunk keeps ties
          best = m[0];
          int j = 31;
#pragma unroll
          for (int q = 30; q >= 0; --q) j = (d[q] == m[0]) ? q : j;            // first (smallest) index inside the chunk
          best_j = t * NN_SB + n0 + j;
        }
      };
      tmem_ld32_issue(t_addr, va);
      tmem_ld_wait();
#pragma unroll 1
      for (int n0 = 0; n0 < NN_SB; n0 += 64) {
        tmem_ld32_issue(t_addr + (uint32_t)(n0 + 32), vb);
        consume(va, n0);
        tmem_ld_wait();
        if (n0 + 64 < NN_SB) tmem_ld32_issue(t_addr + (uint32_t)(n0 + 64), va);
        consume(vb, n0 + 32);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&sh.acc_empty[b]); mbar_arrive(&sh.empty[stage]); }
    }
    const int64_t local = (int64_t)qblk * TM + row;
    if (local < nq) {
      idx_out[q0 + local] = best_j;                                          // -1 when there are no candidates
      if (d_out) d_out[q0 + local] = best_j >= 0 ? fmaxf(na + best, 0.f) : __int_as_float(0x7f800000);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool nn_tc_supported(int C) { return C == NN_C; }

size_t nn_tc_workspace_bytes(int n_pairs, int64_t max_n, int64_t max_m) {
  const int64_t ba = (max_n + NN_SB - 1) / NN_SB > 0 ? (max_n + NN_SB - 1) / NN_SB : 1;
  const int64_t bb = (max_m + NN_SB - 1) / NN_SB > 0 ? (max_m + NN_SB - 1) / NN_SB : 1;
  return (size_t)n_pairs * (size_t)(ba + bb) * BLOCK_IMG + 2048;
}

// idx/d outputs are written directly (no packed best arrays)
int nn_tc(const float* A, const float* B, int C, const int64_t* a_ptr, const int64_t* b_ptr, int n_pairs,
          const int64_t* a_rows, const int64_t* b_rows, int64_t max_n, int64_t max_m, int64_t* idx01, float* d01,
          int64_t* idx10, float* d10, void* workspace, cudaStream_t st) {
  if (C != NN_C) return GCLB_ERR_UNSUPPORTED;
  const int ba = (int)((max_n + NN_SB - 1) / NN_SB) > 0 ? (int)((max_n + NN_SB - 1) / NN_SB) : 1;   // super-blocks per segment
  const int bb = (int)((max_m + NN_SB - 1) / NN_SB) > 0 ? (int)((max_m + NN_SB - 1) / NN_SB) : 1;
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  unsigned char* imgA = ws;
  unsigned char* imgB = ws + (size_t)n_pairs * ba * BLOCK_IMG;
  nn_split_kernel<<<dim3(ba, n_pairs), 256, 0, st>>>(A, a_rows, a_ptr, ba, imgA);
  nn_split_kernel<<<dim3(bb, n_pairs), 256, 0, st>>>(B, b_rows, b_ptr, bb, imgB);
  const size_t smem = (size_t)Q_IMG + (size_t)NN_STAGES * BLOCK_IMG;
  cudaError_t e = cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("nn_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  const int qb = 2 * (ba > bb ? ba : bb);            // 128-row query blocks
  nn_tc_kernel<<<dim3(qb, n_pairs, idx10 ? 2 : 1), kNnThreads, smem, st>>>(imgA, imgB, a_ptr, b_ptr, ba, bb, idx01, d01,
                                                                           idx10, d10);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("nn_tc: CUDA error: %s", cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  count_launches(3);
  return GCLB_OK;
}

}  // namespace gclb
