// K4 on the tensor cores: nearest neighbour in 32-d feature space as a distance GEMM with a fused arg-min.
//
//   d2(i, j) = |a_i|^2 + |b_j|^2 - 2 a_i . b_j ,   a_i . b_j on tcgen05 (kind::tf32, fp32 accumulate in TMEM)
//
// Exactness: the reference computes sum_c (a_c - b_c)^2 in fp32 (lib/metrics.py:22-29) and takes arg-min indices, so a
// plain tf32 product (10-bit mantissa) would flip near-ties.  Every operand is therefore split EXACTLY into three
// tf32-representable parts  a = a1 + a2 + a3  (11 + 11 + 2 significant bits); all partial products a_p * b_q are then
// exact in fp32, and the six groups (1,1) (1,2) (2,1) (2,2) (1,3) (3,1) accumulated in TMEM reproduce the fp32 dot
// product to ~2^-33 relative (the dropped groups) plus fp32 accumulation rounding.
//
// Structure: two row-stationary passes (queries = A over candidates B, then B over A): a CTA owns 128 query rows, keeps
// their three part-images resident in shared memory and streams 128-row candidate tiles (three pre-swizzled part
// images + norms, ONE cp.async.bulk per tile) through a 3-stage ring; one thread issues 24 MMAs (M = 128, N = 128,
// K = 8) per tile into a double-buffered TMEM accumulator; four epilogue warps (thread <-> query row) read the 128
// dot products back and keep the running (min, first index) in registers.  No atomics, no N x M matrix, bit-reproducible.
#include "tc_common.cuh"

namespace gclb {

constexpr int NN_C = 32;                       // channels: one 128-byte swizzle row
constexpr int PART_BYTES = TM * 128;           // 16 KB: one part image of a 128-row block
constexpr int BLOCK_IMG = 3 * PART_BYTES + 1024;   // three parts + 128 norms (padded to keep 1024-byte alignment)
constexpr int NN_STAGES = 3;
constexpr int kNnThreads = 6 * 32;             // warp 0 producer, warp 1 MMA, warps 2-5 epilogue

// ---- split + swizzle: feature rows -> per-128-row-block images -----------------------------------------------------
// grid (blocks_per_seg, n_pairs); block 256 threads = 32 rows x 8 chunks per pass
__global__ void __launch_bounds__(256) nn_split_kernel(const float* __restrict__ X, const int64_t* __restrict__ rows,
                                                       const int64_t* __restrict__ ptr, int blocks_per_seg,
                                                       unsigned char* __restrict__ img) {
  const int p = blockIdx.y, blk = blockIdx.x;
  const int64_t seg0 = ptr[p], n = ptr[p + 1] - seg0;
  if ((int64_t)blk * TM >= n && blk > 0) return;     // nothing of this segment lives here (block 0 is always written)
  unsigned char* out = img + ((size_t)p * blocks_per_seg + blk) * BLOCK_IMG;
  float* norms = reinterpret_cast<float*>(out + 3 * PART_BYTES);
  const int j = threadIdx.x & 7;
  for (int r = threadIdx.x >> 3; r < TM; r += 32) {
    const int64_t local = (int64_t)blk * TM + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool valid = local < n;
    if (valid) {
      const int64_t gr = rows ? __ldg(rows + seg0 + local) : seg0 + local;
      v = __ldg(reinterpret_cast<const float4*>(X + (size_t)gr * NN_C) + j);
    }
    float x[4] = {v.x, v.y, v.z, v.w}, p1[4], p2[4], p3[4];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      p1[q] = __uint_as_float(__float_as_uint(x[q]) & 0xFFFFE000u);        // top 11 significant bits
      const float r1 = x[q] - p1[q];                                        // exact
      p2[q] = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
      p3[q] = r1 - p2[q];                                                   // exact, <= 2-3 significant bits
      ss = fmaf(x[q], x[q], ss);
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(out + off) = make_float4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<float4*>(out + PART_BYTES + off) = make_float4(p2[0], p2[1], p2[2], p2[3]);
    *reinterpret_cast<float4*>(out + 2 * PART_BYTES + off) = make_float4(p3[0], p3[1], p3[2], p3[3]);
    if (j == 0) norms[r] = valid ? ss : __int_as_float(0x7f800000);        // padding rows can never win
  }
}

struct NnShared {
  uint64_t q_full;
  uint64_t full[NN_STAGES], empty[NN_STAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// grid (blocks_per_seg, n_pairs, directions)
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
  if (qblk >= q_blocks || (int64_t)qblk * TM >= nq) return;            // whole CTA: no query rows here
  const int n_tiles = (int)((nc + TM - 1) / TM);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* q_s = smem_dyn;                                       // 3 parts + norms of the query block
  unsigned char* ring = smem_dyn + BLOCK_IMG;

  if (tid == 0) {
    if ((smem_u32(smem_dyn) & 1023u) != 0) { printf("gclb nn_tc: shared memory not 1024-byte aligned\n"); __trap(); }
    mbar_init(&sh.q_full, 1);
    for (int s = 0; s < NN_STAGES; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1 + 4); }   // MMA commit + 4 epilogue warps
    for (int b = 0; b < 2; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const uint32_t q_u32 = smem_u32(q_s), ring_u32 = smem_u32(ring);

  if (warp == 0) {
    // ================= producer: one bulk copy per block image =================
    if (lane == 0) {
      mbar_arrive_expect_tx(&sh.q_full, BLOCK_IMG);
      bulk_g2s(q_u32, q_img + ((size_t)p * q_blocks + qblk) * BLOCK_IMG, BLOCK_IMG, &sh.q_full);
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
      constexpr uint32_t idesc = make_idesc_tf32(TM);      // N = 128
      mbar_wait(&sh.q_full, 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int stage = t % NN_STAGES, b = t & 1;
        mbar_wait(&sh.acc_empty[b], ((t >> 1) & 1) ^ 1);
        mbar_wait(&sh.full[stage], (t / NN_STAGES) & 1);
        tc_fence_after();
        const uint32_t c_s = ring_u32 + stage * BLOCK_IMG;
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * TM);
        // smallest terms first: (3,1) (1,3) (2,2) (2,1) (1,2) (1,1)
        const int qa[6] = {2, 0, 1, 1, 0, 0}, cb[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
        for (int g = 0; g < 6; ++g) {
#pragma unroll
          for (int ks = 0; ks < NN_C / 8; ++ks)
            umma_tf32(d_tmem, make_desc_sw128(q_u32 + qa[g] * PART_BYTES + ks * 32),
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
    const float* qn = reinterpret_cast<const float*>(q_s + 3 * PART_BYTES);
    mbar_wait(&sh.q_full, 0);
    const float na = qn[row];
    float best = __int_as_float(0x7f800000);
    int best_j = -1;
    for (int t = 0; t < n_tiles; ++t) {
      const int stage = t % NN_STAGES, b = t & 1;
      mbar_wait(&sh.full[stage], (t / NN_STAGES) & 1);      // the tile's norms (bulk-copied with the operands) are visible
      mbar_wait(&sh.acc_full[b], (t >> 1) & 1);
      tc_fence_after();
      const float* cn = reinterpret_cast<const float*>(ring + stage * BLOCK_IMG + 3 * PART_BYTES);
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * TM);
#pragma unroll 1
      for (int n0 = 0; n0 < TM; n0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_addr + (uint32_t)n0, v);
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const float d = fmaf(-2.f, __uint_as_float(v[q]), cn[n0 + q]);     // |b|^2 - 2 a.b   (+|a|^2 at the end)
          if (d < best) { best = d; best_j = t * TM + n0 + q; }             // strict: first (smallest) index wins ties
        }
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

bool nn_tc_supported(int C) { return C == NN_C; }

size_t nn_tc_workspace_bytes(int n_pairs, int64_t max_n, int64_t max_m) {
  const int64_t ba = (max_n + TM - 1) / TM > 0 ? (max_n + TM - 1) / TM : 1, bb = (max_m + TM - 1) / TM > 0 ? (max_m + TM - 1) / TM : 1;
  return (size_t)n_pairs * (size_t)(ba + bb) * BLOCK_IMG + 2048;
}

// idx/d outputs are written directly (no packed best arrays)
int nn_tc(const float* A, const float* B, int C, const int64_t* a_ptr, const int64_t* b_ptr, int n_pairs,
          const int64_t* a_rows, const int64_t* b_rows, int64_t max_n, int64_t max_m, int64_t* idx01, float* d01,
          int64_t* idx10, float* d10, void* workspace, cudaStream_t st) {
  if (C != NN_C) return GCLB_ERR_UNSUPPORTED;
  const int ba = (int)((max_n + TM - 1) / TM) > 0 ? (int)((max_n + TM - 1) / TM) : 1;
  const int bb = (int)((max_m + TM - 1) / TM) > 0 ? (int)((max_m + TM - 1) / TM) : 1;
  unsigned char* ws = reinterpret_cast<unsigned char*>(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  unsigned char* imgA = ws;
  unsigned char* imgB = ws + (size_t)n_pairs * ba * BLOCK_IMG;
  nn_split_kernel<<<dim3(ba, n_pairs), 256, 0, st>>>(A, a_rows, a_ptr, ba, imgA);
  nn_split_kernel<<<dim3(bb, n_pairs), 256, 0, st>>>(B, b_rows, b_ptr, bb, imgB);
  const size_t smem = (size_t)(1 + NN_STAGES) * BLOCK_IMG;
  cudaError_t e = cudaFuncSetAttribute(nn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("nn_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  const int qb = ba > bb ? ba : bb;
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
