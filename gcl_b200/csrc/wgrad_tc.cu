// K3 wgrad on tcgen05 (sm_100a):  gW[k][c][n] = sum over table rows o with i = nbr[o][k] >= 0 of  X[i][c] * G[o][n]
// (autograd of ME.MinkowskiConvolution; the reference's ConvolutionBackward).
//
// The contraction runs over ROWS, so both operands are "MN-major" for the tensor core: the gathered row tiles
//   A' = X[nbr[rows][k]]  (rows x Cin)      B' = G[perm[rows]]  (rows x Cout)
// are fetched by TMA tile::gather4 exactly like the forward kernel's A operand (128-byte swizzled lines, one line = one row
// x 32 channels, missing neighbours zero-filled by the out-of-bounds rule); only the swizzle differs: tcgen05 takes MN-major
// tf32 operands in ONE shared-memory layout, SWIZZLE_128B_BASE32B (32-byte chunks XOR line % 4, 4 rows per 512-byte atom
// along K), which the TMA writes when the tensor map says SWIZZLE_128B_ATOM_32B (plain SWIZZLE_128B + MN-major tf32
// silently accumulates zeros -- measured).  32 channels are contiguous along M (or N), 32-channel slabs LBO bytes
// apart: no transpose pass anywhere.
//
//   D[c (128 TMEM lanes), n (Cout columns)] += A'[8 rows, 128 c]^T * B'[8 rows, Cout]     one tcgen05.mma.kind::tf32 per 8 rows
//
// Work item = (kernel offset k, 1/P of the 128-row tiles): a CTA walks the tiles of its share whose tile mask has bit k,
// accumulates ALL of them in TMEM (Cin/128 accumulators of Cout columns), and flushes once with vector atomics into gW.
// grid = K x P items, about four per SM, handed out by the hardware scheduler as CTAs retire (offsets are unevenly
// populated: the centre offset touches every tile, corner offsets few).
// 13 warps: 0-7 producers (warp w owns ring slot w), 8 MMA issuer + TMEM owner, 9-12 epilogue.
#include "tc_common.cuh"

namespace gclb {

constexpr int kWgMmaWarp = 8;
constexpr int kWgThreads = 13 * 32;
constexpr int kWgMaxTiles = 2048;     // active tiles one CTA can list

// 3x3x3 offsets (index = (dx+1) + 3(dy+1) + 9(dz+1), the table's column order) sorted by |d|_1: centre, 6 faces, 12 edges, 8 corners
__constant__ int kPopularity27[27] = {13, 4, 10, 12, 14, 16, 22, 1, 3, 5, 7, 9, 11, 15, 17, 19, 21, 23, 25, 0, 2, 6, 8, 18, 20, 24, 26};

struct WgShared {
  uint64_t full[8], empty[8];
  uint64_t acc_full;
  uint32_t tmem_base;
  int n_tiles;
  int tiles[kWgMaxTiles];
};

struct WgParams {
  const int32_t* nbr;        // row-bucketed table [n_out, K]; NULL = identity (K == 1)
  const int32_t* perm;       // tile row -> output row; NULL = identity
  const uint32_t* tile_mask; // per 128-row tile: populated offsets; NULL = every tile
  float* gW;                 // [K, cin, cout], accumulated into
  int64_t n_out;
  int K, cin, cout;
  int P, num_tiles;
  int rows;                  // rows per pipeline stage: 32 / 64 / 128
  int stages;
  int tmem_cols;
};

// MN-major descriptor, layout SWIZZLE_128B_BASE32B (type 1; the only shared-memory layout tcgen05 takes for MN-major
// tf32 operands): 128-byte lines of 32 channels, 32-byte chunks XOR (line % 4), 4 lines (rows) per 512-byte atom along K;
// LBO = bytes between 32-channel slabs along M/N, SBO = bytes between 4-row atoms along K.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  const uint32_t sbo = 512;
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}

__global__ void __launch_bounds__(kWgThreads, 1) spconv_wgrad_tc_kernel(WgParams p, const __grid_constant__ CUtensorMap map_x,
                                                                        const __grid_constant__ CUtensorMap map_g) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ WgShared sh;
  const int tid = threadIdx.x, warp = (int)warp_uniform((uint32_t)(threadIdx.x >> 5)), lane = tid & 31;
  // items are handed out in blockIdx order: start the heavy offsets first (centre, then faces, edges, corners of the
  // 3x3x3 stencil -- populated in that order of frequency) so the light ones fill the tail
  int k = blockIdx.x % p.K;
  int part = blockIdx.x / p.K;
  if (p.K == 27) {
    part = blockIdx.x % p.P;
    k = blockIdx.x / p.P;   // all parts of one offset are adjacent, offsets in popularity order
    k = kPopularity27[k];
  }
  const int S = p.stages;
  const int rows = p.rows, chunks = TM / rows, quads = rows / 4;
  const int slabsA = p.cin / 32, slabsB = p.cout / 32;
  const uint32_t slab_bytes = (uint32_t)rows * 128u;
  const uint32_t a_bytes = slabsA * slab_bytes;
  const uint32_t stage_bytes = (slabsA + slabsB) * slab_bytes;
  const uint32_t ring_u32 = smem_u32(smem_dyn);

  if (tid == 0) {
    if ((ring_u32 & 1023u) != 0) { printf("gclb wgrad_tc: operand ring not 1024-byte aligned\n"); __trap(); }
    for (int s = 0; s < 8; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
    mbar_init(&sh.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWgMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == 0) {   // ordered list of this item's active tiles: part, part + P, part + 2P, ... with bit k set
    int cnt = 0;
    for (int64_t j0 = 0; part + j0 * p.P < p.num_tiles; j0 += 32) {
      const int64_t t = part + (j0 + lane) * p.P;
      const bool act = t < p.num_tiles && (p.tile_mask ? ((__ldg(p.tile_mask + t) >> k) & 1u) : true);
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (act) sh.tiles[cnt + __popc(m & ((1u << lane) - 1))] = (int)t;
      cnt += __popc(m);
    }
    if (lane == 0) sh.n_tiles = cnt;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const int n_tiles = sh.n_tiles;
  const int n_stages = n_tiles * chunks;

  if (warp < 8) {
    if (warp < S) {
      // ===================================== producers: TMA gather of both operands ==============================
      // stage `it` = (active tile it / chunks, row chunk it % chunks) lives in ring slot it % S = this warp.  Lane l
      // owns row quad l % quads: it loads that quad's 4 input-row and 4 output-row indices once, then issues one
      // gather4 per 32-channel slab it is responsible for (slabs l / quads, l / quads + 32 / quads, ...).
      for (int it = warp; it < n_stages; it += S) {
        const int tile = sh.tiles[it / chunks];
        const int quad = lane % quads;
        const int64_t row0 = (int64_t)tile * TM + (it % chunks) * rows + quad * 4;
        int ra[4], rb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t row = row0 + q;
          const bool in = row < p.n_out;
          ra[q] = !in ? -1 : (p.nbr ? __ldg(p.nbr + row * p.K + k) : (int)row);
          rb[q] = !in ? -1 : (p.perm ? __ldg(p.perm + row) : (int)row);
        }
        mbar_wait(&sh.empty[warp], (((uint32_t)(it / S)) & 1u) ^ 1u);
        const uint32_t base = ring_u32 + warp * stage_bytes;
        if (lane == 0) mbar_arrive_expect_tx(&sh.full[warp], stage_bytes);
        __syncwarp();
        const int n_instr = quads * (slabsA + slabsB);
        for (int j = lane; j < n_instr; j += 32) {
          const int slab = j / quads;
          const uint32_t dst = base + slab * slab_bytes + quad * 512;
          if (slab < slabsA) tma_gather4(dst, &map_x, &sh.full[warp], slab * 32, ra[0], ra[1], ra[2], ra[3]);
          else tma_gather4(dst, &map_g, &sh.full[warp], (slab - slabsA) * 32, rb[0], rb[1], rb[2], rb[3]);
        }
      }
    }
  } else if (warp == kWgMmaWarp) {
    // ======================================= MMA issuer (warp-uniform loop, one elected lane) ================
    // all 32 lanes walk the loop on warp-uniform values, only the tcgen05 instructions sit under elect.sync: the descriptors
    // stay in uniform registers and the MMAs of a stage issue back to back (tc_common.cuh: 59 instead of 95-110 cycles each)
    {
      const uint32_t idesc = make_idesc_tf32(p.cout) | (1u << 15) | (1u << 16);   // A and B both MN-major
      const int mchunks = (p.cin + 127) / 128;
      const uint32_t tmem_u = warp_uniform(tmem_base);
      const int n_st = (int)warp_uniform((uint32_t)n_stages);
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < n_st; ++it) {
        mbar_wait(&sh.full[stage], phase);
        tc_fence_after();
        const uint32_t a_s = ring_u32 + stage * stage_bytes;
        const uint32_t b_s = a_s + a_bytes;
        if (elect_one()) {
          for (int ks = 0; ks < rows / 8; ++ks) {
            const uint64_t bdesc = make_desc_mn_sw128(b_s + ks * 1024, slab_bytes);
            for (int mc = 0; mc < mchunks; ++mc)
              umma_tf32(tmem_u + (uint32_t)(mc * p.cout), make_desc_mn_sw128(a_s + mc * 4 * slab_bytes + ks * 1024, slab_bytes),
                        bdesc, idesc, (it | ks) ? 1u : 0u);
          }
          umma_commit(&sh.empty[stage]);
        }
        __syncwarp();
        if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
      }
      if (n_st > 0 && elect_one()) umma_commit(&sh.acc_full);
      __syncwarp();
    }
  } else if (n_stages > 0) {
    // ======================================= epilogue: thread <-> input channel ===============================
    const int quarter = warp & 3;
    mbar_wait(&sh.acc_full, 0);
    tc_fence_after();
    const int mchunks = (p.cin + 127) / 128;
    for (int mc = 0; mc < mchunks; ++mc) {
      const int c = mc * 128 + quarter * 32 + lane;
      float* dst = p.gW + ((size_t)k * p.cin + c) * p.cout;
      for (int n0 = 0; n0 < p.cout; n0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mc * p.cout + n0), v);
        if (c < p.cin) {
#pragma unroll
          for (int q = 0; q < 32; q += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + n0 + q),
                      make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                                  __uint_as_float(v[q + 3])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" int gclb_spconv_wgrad_tc(const float* in, int32_t cin, int64_t n_in, const float* gout, int32_t cout,
                                    int64_t n_out, const int32_t* nbr_sorted, const int32_t* row_perm,
                                    const int32_t* tile_mask, int32_t K, float* gW, void* stream) {
  GCLB_CHECK_ARG(gW && K >= 1, "bad arguments");
  GCLB_CHECK_ARG(cin >= 32 && cin <= 256 && cin % 32 == 0, "tensor-core wgrad needs cin in {32, 64, ..., 256}");
  GCLB_CHECK_ARG(cout >= 32 && cout <= 256 && cout % 32 == 0, "tensor-core wgrad needs cout in {32, 64, ..., 256}");
  GCLB_CHECK_ARG(nbr_sorted || K == 1, "nbr may be NULL only for K == 1");
  GCLB_CHECK_ARG(!tile_mask || K <= 32, "tile masks cover at most 32 offsets");
  GCLB_CHECK_ARG(n_out == 0 || (in && gout), "null pointer");
  if (n_out == 0 || n_in == 0) return GCLB_OK;
  WgParams p;
  p.nbr = nbr_sorted;
  p.perm = row_perm;
  p.tile_mask = reinterpret_cast<const uint32_t*>(tile_mask);
  p.gW = gW;
  p.n_out = n_out;
  p.K = K;
  p.cin = cin;
  p.cout = cout;
  p.num_tiles = (int)((n_out + TM - 1) / TM);
  const int sum = cin + cout;
  p.rows = sum <= 128 ? 128 : (sum <= 256 ? 64 : 32);                       // 32 or 64 KB per stage
  const size_t stage_bytes = (size_t)p.rows * sum * 4;
  const int mchunks = (cin + 127) / 128;
  // the M = 128 instruction always reads four 32-channel slabs per accumulator; when cin is not a multiple of 128 the
  // surplus slabs (rows of D nobody looks at) alias whatever follows A' -- keep that inside the allocation
  const int overflow = mchunks * 4 - (cin + cout) / 32;
  const size_t slack = overflow > 0 ? (size_t)overflow * p.rows * 128 : 0;
  const size_t avail = (216u << 10) - slack;                                // 227 KB minus static shared and margin
  p.stages = (int)(avail / stage_bytes > 8 ? 8 : avail / stage_bytes);
  int cols = 32;
  while (cols < mchunks * cout) cols <<= 1;
  p.tmem_cols = cols;
  int P = (4 * kNumSMs + K - 1) / K;
  if (P > p.num_tiles) P = p.num_tiles;
  const int need = (p.num_tiles + kWgMaxTiles - 1) / kWgMaxTiles;
  if (P < need) P = need;
  p.P = P;
  const size_t smem = (size_t)p.stages * stage_bytes + slack;
  cudaError_t e = cudaFuncSetAttribute(spconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("spconv_wgrad_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  CUtensorMap map_x, map_g;
  int rc = make_rows_tensor_map_sw(&map_x, in, n_in, cin, 1, true);
  if (rc == GCLB_OK) rc = make_rows_tensor_map_sw(&map_g, gout, n_out, cout, 1, true);
  if (rc != GCLB_OK) return rc;
  spconv_wgrad_tc_kernel<<<(unsigned)(K * P), kWgThreads, smem, (cudaStream_t)stream>>>(p, map_x, map_g);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}
