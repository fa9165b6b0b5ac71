// K3 tensor-core path: sparse convolution as an output-stationary implicit GEMM on tcgen05 (sm_100a).
//
//   D[128 out rows, Cout] (fp32, TMEM)  +=  A[128 gathered rows, one slab] (smem)  x  B[Cout, one slab] (smem)
//   one pipeline stage per (populated kernel offset k, channel slab); a slab is one 128-byte line of a row: 64 fp16 channels
//   (kind::f16, K = 16) or 32 fp32 channels (kind::tf32, K = 8) -- 4 tcgen05.mma per stage -- or a 64-byte line of 32 fp16
//   channels (2 per stage) for tensors whose width is a multiple of 32 only; see the operand modes below
//
// Persistent, warp-specialised CTA (one per SM), 448 threads:
//   warps 0-7   A producers, a WARP per stage (ring slot s belongs to warp s % 8): the valid neighbour rows (x 128 B)
//               of a stage are fetched by up to 32 TMA tile::gather4 instructions (one per lane, 4 rows each) straight into
//               the canonical K-major SWIZZLE_128B layout.  MISSING neighbours are not fetched (an out-of-bounds gather row is
//               zero-filled but costs ~4x an in-bounds one): the stage carries a 128-bit row mask and its MMAs run with those
//               output lanes disabled.  Lane 0 also pulls the weight slab with ONE cp.async.bulk (weights are stored
//               pre-swizzled, see gclb_weights_to_tc) onto the same mbarrier.  No LSU traffic, no thread waits for data.
//   warp  8     issues tcgen05.mma from one elected lane of a warp-uniform loop (descriptors stay in uniform registers);
//               tcgen05.commit recycles smem slots and publishes accumulators.
//   warp  9     prefetches the next tiles' slices of the neighbour table (one bulk copy per tile) and lists their populated offsets.
//   warps 10-13 epilogue: tcgen05.ld (thread <-> output row), fused scale/shift (+residual) (+ReLU) (+L2 normalise),
//               overlapped with the next tiles' main loops through a ring of TMEM accumulators, which it hands back zeroed;
//               output row ids are requested two tiles ahead, residual rows one tile ahead.
// Every output row is produced by exactly one CTA: no atomics, bit-reproducible.
// All mbarrier waits are bounded spins that trap on a protocol bug instead of hanging the GPU.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace gclb {

constexpr int KSLAB = 32;             // tf32: channels per stage = one 128-byte swizzle row
// operand modes: 0 = fp32 activations, kind::tf32, 32 channels per 128-byte row (SWIZZLE_128B)
//                1 = fp16 activations, kind::f16,  64 channels per 128-byte row (SWIZZLE_128B)
//                2 = fp16 activations, kind::f16,  32 channels per  64-byte row (SWIZZLE_64B): sources whose width is a
//                    multiple of 32 but not of 64 (the 32-channel stride-1 level of the ResUNet)
// CTAS = CTAs per SM.  1: the original layout (8 gather warps | MMA | nbr | 4 epilogue = 448 threads, the whole shared memory).
// 2: two independent half-size CTAs per SM (4 gather warps | MMA | nbr | 4 epilogue = 320 threads, half the ring each).  The cycle
// accounting of round 2 (profiles/r02_conv_ablation.md) showed the kernel bound by the serial latency chains of its roles -- one
// thread issuing a tcgen05.mma every ~100 cycles, barrier round trips, the epilogue of a tile -- not by a throughput resource:
// with no data movement at all it still took 85 % of its time.  Two CTAs per SM interleave two such chains on the same SM.
// kEpiWarps: 4 = one epilogue group; 8 = two groups of four warps draining alternate tiles.  Measured (64 -> 64, 641 k rows):
// two groups with FOUR gather warps lower the per-tile floor (88 -> 76 us at <= 4 offsets per tile) but four warps cannot issue
// the gathers fast enough (276 instead of 172 ns per stage): 173 us; two groups with EIGHT gather warps (576 threads => 96
// registers, spills): 129 us; one group, eight gather warps: 125 us.  The code below handles any of these.
template <int CTAS> struct TcRoles {
  static constexpr int kGatherWarps = CTAS == 2 ? 4 : 8;
  static constexpr int kEpiWarps = 4;
  static constexpr int kMmaWarp = kGatherWarps, kNbrWarp = kGatherWarps + 1, kEpiWarp0 = kGatherWarps + 2;
  static constexpr int kThreads = (kGatherWarps + 2 + kEpiWarps) * 32;     // 448 / 320
};

template <int COUT, int MODE = 0, int CTAS = 1>
struct TcCfg {
  static constexpr int ROWB = MODE == 2 ? 64 : 128;       // bytes per operand row
  static constexpr int A_BYTES = TM * ROWB;               // 16 KB / 8 KB
  static constexpr int B_BYTES = COUT * ROWB;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  // CTAS == 1: deepest ring that leaves room for two neighbour tiles (2 x 14 KB) in 227 KB; CTAS == 2: ~112 KB and 256 TMEM
  // columns per CTA
  static constexpr int STAGES = CTAS == 2 ? (COUT >= 128 ? 2 : (STAGE > 16384 ? 3 : 4))
                                          : (COUT >= 256 ? 4 : (COUT >= 128 ? 5 : (COUT >= 64 ? 7 : 8)));
  static constexpr int NBUF = (CTAS == 2 || COUT >= 256) ? 2 : 4;  // neighbour-tile ring (tiles prefetched ahead)
  static constexpr int NACC = CTAS == 2 ? (COUT >= 128 ? 2 : 4) : (COUT >= 256 ? 2 : 4);   // TMEM accumulator ring
  static constexpr int TMEM_COLS = NACC * COUT;                   // powers of two; <= 256 per CTA when two share an SM
};

struct TcShared {   // static shared: barriers + small per-tile metadata
  uint64_t full[8], empty[8];
  uint64_t acc_full[4], acc_empty[4];
  uint64_t nbr_full[4], nbr_empty[4];
  uint32_t tmem_base;
  int n_act[4];
  int acc_n_act[4];        // per accumulator: number of populated offsets of the tile it holds (0 => treat as zeros)
  int act_k[4][32];
  uint4 lane_mask[8];      // per ring slot: rows (TMEM lanes) whose neighbour is missing at this stage's offset => MMA output switched off
};

template <int COUT, int KVOL, int MODE, int CTAS>
__global__ void __launch_bounds__(TcRoles<CTAS>::kThreads, CTAS) spconv_fwd_tc_kernel(ConvParams p, int num_tiles, int normalize, int dbg,
                                                                      const __grid_constant__ CUtensorMap map0,
                                                                      const __grid_constant__ CUtensorMap map1) {
  using Cfg = TcCfg<COUT, MODE, CTAS>;
  constexpr int kGatherWarps = TcRoles<CTAS>::kGatherWarps, kMmaWarp = TcRoles<CTAS>::kMmaWarp, kNbrWarp = TcRoles<CTAS>::kNbrWarp;
  constexpr bool HALF = MODE != 0;
  constexpr int A_BYTES = Cfg::A_BYTES;
  constexpr int ROWB = Cfg::ROWB;
  constexpr int S = Cfg::STAGES;
  constexpr int NBUF = Cfg::NBUF, NACC = Cfg::NACC;
  constexpr int NBR_INTS = TM * KVOL;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ TcShared sh;
  unsigned char* ring = smem_dyn;
  int* nbr_buf = reinterpret_cast<int*>(smem_dyn + S * Cfg::STAGE);    // [NBUF][NBR_INTS]

  const int tid = threadIdx.x, warp = (int)warp_uniform((uint32_t)(threadIdx.x >> 5)), lane = tid & 31;   // role index: provably warp-uniform
  // HALF: activations (in0, in1, residual) are IEEE fp16 in HBM, weights an fp16 image, MMA kind::f16 -- a 128-byte
  // operand row then holds 64 channels instead of 32: half the gather bytes, same 10-bit mantissa as kind::tf32
  constexpr int KCH = MODE == 1 ? 64 : 32;
  const int cin = p.c0 + p.c1;
  const int slabs = cin / KCH;
  const bool identity = (p.nbr == nullptr);     // K == 1 `mm` path: nbr[o] = o
  const bool nbr_sorted = (p.relu & 4) != 0;    // nbr is the physically re-ordered copy (else: read rows through perm)

  if (tid == 0) {
    if ((smem_u32(ring) & 1023u) != 0) { printf("gclb spconv_tc: operand ring not 1024-byte aligned\n"); __trap(); }
    for (int s = 0; s < S; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }   // one arrive.expect_tx per stage
    for (int b = 0; b < 4; ++b) {
      mbar_init(&sh.acc_full[b], 1);
      mbar_init(&sh.acc_empty[b], 4);
      mbar_init(&sh.nbr_full[b], 1);
      mbar_init(&sh.nbr_empty[b], (S < kGatherWarps ? S : kGatherWarps) + 1);   // gather warps + MMA thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const uint32_t ring_u32 = smem_u32(ring);
  // the row-masked MMAs only ever accumulate: every accumulator starts (and is handed back by the epilogue) as zeros
  if (warp >= TcRoles<CTAS>::kEpiWarp0 && warp < TcRoles<CTAS>::kEpiWarp0 + 4) {
    for (int c0 = 0; c0 < Cfg::TMEM_COLS; c0 += 32) tmem_st32_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kGatherWarps) {
    if (warp < S) {
    // ======================================= A producers: TMA gather, one WARP per ring slot ==================
    // Warp w owns ring slot w (stages it with it % S == w; S <= 8).  Per stage: lane 0 arms the slot's mbarrier with the
    // stage's byte count and pulls the weight slab with one bulk copy; then every lane issues ONE tile::gather4 (rows
    // 4*lane .. 4*lane+3 of the tile, 32 channels): 32 TMA instructions move the whole 128 x 128 B operand, swizzled by
    // the hardware, missing neighbours (-1) zero-filled by the TMA's out-of-bounds rule.  Nothing goes through the
    // LSU / L1TEX, no thread waits for data: completion arrives on the mbarrier (complete_tx).
    uint32_t it = 0;                  // global stage counter (identical in every role)
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt % NBUF;
      const int tile_m = tile * TM;
      mbar_wait(&sh.nbr_full[b], (lt / NBUF) & 1);
      const int* nb = nbr_buf + b * NBR_INTS;
      const int n_iter = sh.n_act[b] * slabs;
      for (int i = 0; i < n_iter; ++i, ++it) {
        if ((int)((it % S) % kGatherWarps) != warp) continue;   // a warp walks its slots in stage order => no phase aliasing
        const int k = sh.act_k[b][i / slabs];
        const int c = (i % slabs) * KCH;
        const int stage = it % S;
        int r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (identity) { const int64_t t_row = (int64_t)tile_m + 4 * lane + q; r[q] = t_row < p.n_out ? (int)t_row : -1; }
          else r[q] = nb[(4 * lane + q) * KVOL + k];
        }
        // Missing neighbours are NOT fetched: an out-of-bounds row of a gather4 is zero-filled correctly but costs ~4x an
        // in-bounds row (measured, profiles/r02_conv_ablation.md).  Instead the MMAs of this stage run with those rows'
        // output lanes switched off.  A quad with no valid row issues nothing; inside a partly valid quad the missing slots
        // re-fetch a valid row of the same quad (same line, L2-hot; its product is masked out).
        const uint32_t nib = (uint32_t)(r[0] >= 0) | ((uint32_t)(r[1] >= 0) << 1) | ((uint32_t)(r[2] >= 0) << 2) | ((uint32_t)(r[3] >= 0) << 3);
        const int fv = (nib & 1u) ? r[0] : (nib & 2u) ? r[1] : (nib & 4u) ? r[2] : r[3];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = r[q] >= 0 ? r[q] : fv;
        const uint32_t quads = __ballot_sync(0xffffffffu, nib != 0u);
        const uint32_t off_bits = (nib ^ 15u) << (4 * (lane & 7));      // smem row 4*lane + q <-> TMEM lane 4*lane + q
        uint4 dm;
        dm.x = __reduce_or_sync(0xffffffffu, (lane >> 3) == 0 ? off_bits : 0u);
        dm.y = __reduce_or_sync(0xffffffffu, (lane >> 3) == 1 ? off_bits : 0u);
        dm.z = __reduce_or_sync(0xffffffffu, (lane >> 3) == 2 ? off_bits : 0u);
        dm.w = __reduce_or_sync(0xffffffffu, (lane >> 3) == 3 ? off_bits : 0u);
        mbar_wait(&sh.empty[stage], ((it / S) & 1u) ^ 1u);   // passes immediately during the first round
        const uint32_t a_s = ring_u32 + stage * Cfg::STAGE;
        if (lane == 0) {
          sh.lane_mask[stage] = dm;                          // published by the arrive below (release) to the MMA warp
          const uint32_t bytes = ((dbg & 1) ? 0u : (uint32_t)__popc(quads) * (4u * ROWB)) + ((dbg & 2) ? 0u : (uint32_t)Cfg::B_BYTES);   // ablation: GCLB_TC_DBG
          if (bytes) mbar_arrive_expect_tx(&sh.full[stage], bytes);
          else mbar_arrive(&sh.full[stage]);
          if (!(dbg & 2))
            bulk_g2s(a_s + A_BYTES, reinterpret_cast<const unsigned char*>(p.W) + ((size_t)k * slabs + c / KCH) * Cfg::B_BYTES,
                     Cfg::B_BYTES, &sh.full[stage]);
        }
        __syncwarp();                                        // the barrier is armed before any gather can complete on it
        if ((dbg & 1) || nib == 0u) continue;
        // every lane issues the gather4 of its own quad.  (A TMA instruction takes its coordinates from uniform registers, so
        // this compiles to a 32-trip vote loop, ~70 cycles per trip; walking the quads in a warp-uniform loop with shuffled
        // row ids, four instructions per block on rotating uniform registers, measured SLOWER (85 cycles per instruction).
        // Eight gather warps reach ~170 ns per 128-row stage, which is also what in-bounds sequential rows reach: the TMA
        // unit's own rate, ~2.5 cycles per 128-byte row.)
        if (c < p.c0) tma_gather4(a_s + lane * (4 * ROWB), &map0, &sh.full[stage], c, r[0], r[1], r[2], r[3]);
        else tma_gather4(a_s + lane * (4 * ROWB), &map1, &sh.full[stage], c - p.c0, r[0], r[1], r[2], r[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.nbr_empty[b]);   // this warp no longer reads the neighbour tile
    }
    }
  } else if (warp == kMmaWarp) {
    // ======================================= MMA issuer (warp-uniform loop, one elected lane) ================
    // The whole warp walks the tile / stage loop on warp-uniform values; only the tcgen05 instructions and the barrier
    // arrivals sit under elect.sync.  That keeps the descriptors in uniform registers (see tc_common.cuh: 59 instead of
    // 110 cycles per tcgen05.mma at N = 64 -- this thread's issue chain bounds the kernel, profiles/r02_conv_ablation.md).
    {
      constexpr uint32_t idesc = HALF ? make_idesc_f16(COUT) : make_idesc_tf32(COUT);
      constexpr uint32_t HI = MODE == 2 ? kDescHiSw64 : kDescHiSw128;
      const uint32_t tmem_u = warp_uniform(tmem_base);
      const uint32_t ring_lo = desc_lo(ring_u32);
      uint32_t stage = 0, phase = 0;       // ring position of the global stage counter (identical in every role)
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int b = lt % NBUF, ab = lt % NACC;
        mbar_wait(&sh.nbr_full[b], (lt / NBUF) & 1);
        const int n_act = (int)warp_uniform((uint32_t)sh.n_act[b]);      // every lane has read it before the arrive below
        const int n_iter = n_act * slabs;
        if (elect_one()) mbar_arrive(&sh.nbr_empty[b]);
        mbar_wait(&sh.acc_empty[ab], ((lt / NACC) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        if (elect_one()) *reinterpret_cast<volatile int*>(&sh.acc_n_act[ab]) = n_act;   // read by the epilogue after acc_full
        const uint32_t d_tmem = tmem_u + (uint32_t)(ab * COUT);
        for (int i = 0; i < n_iter; ++i) {
          mbar_wait(&sh.full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = ring_lo + stage * (uint32_t)(Cfg::STAGE >> 4);
          const uint32_t b_lo = a_lo + (uint32_t)(A_BYTES >> 4);
          uint4 dmv;                                   // ordered after the barrier's acquire by the asm's memory clobber
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(dmv.x), "=r"(dmv.y), "=r"(dmv.z), "=r"(dmv.w) : "r"(smem_u32(&sh.lane_mask[stage])) : "memory");
          const uint32_t m0 = warp_uniform(dmv.x), m1 = warp_uniform(dmv.y), m2 = warp_uniform(dmv.z), m3 = warp_uniform(dmv.w);
          if (elect_one()) {
            if (dbg & 4) {   // ablation: no MMAs (the commits still recycle the ring)
            } else if constexpr (HALF) {
#pragma unroll
              for (int ks = 0; ks < ROWB / 32; ++ks) umma_f16_um<HI>(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, idesc, m0, m1, m2, m3);
            } else {
#pragma unroll
              for (int ks = 0; ks < ROWB / 32; ++ks) umma_tf32_um<HI>(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, idesc, m0, m1, m2, m3);
            }
            umma_commit(&sh.empty[stage]);            // frees the slot once these MMAs have read it
          }
          __syncwarp();
          if (++stage == (uint32_t)S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) {
          if (n_iter > 0) umma_commit(&sh.acc_full[ab]);
          else mbar_arrive(&sh.acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else if (warp == kNbrWarp) {
    // ======================================= neighbour-tile prefetch ========================================
    // Fast path (bucket-sorted table with precomputed tile masks, the engine's case): a whole tile's slice of the table is ONE
    // bulk copy that completes on nbr_full itself, the populated-offset list comes from the tile mask, and the mask of the next
    // tile is requested a tile ahead -- this warp never waits for memory, up to NBUF table tiles are in flight.  (Before:
    // mask load -> cp.async -> wait, one tile at a time: ~1.5 us of exposed latency per tile, the floor of the whole kernel
    // once the MMA issue chain and the out-of-bounds gathers were gone, profiles/r02_conv_ablation.md.)
    const bool bulk_ok = !identity && !(p.perm && !nbr_sorted) && p.tile_mask != nullptr &&
                         (reinterpret_cast<uintptr_t>(p.nbr) & 15u) == 0 && !(dbg & 8);
    unsigned m_next = (bulk_ok && (int)blockIdx.x < num_tiles) ? __ldg(p.tile_mask + blockIdx.x) : 0u;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt % NBUF;
      const unsigned m_cur = m_next;
      if (bulk_ok && tile + (int)gridDim.x < num_tiles) m_next = __ldg(p.tile_mask + tile + gridDim.x);
      mbar_wait(&sh.nbr_empty[b], ((lt / NBUF) & 1) ^ 1);
      int* nb = nbr_buf + b * NBR_INTS;
      if (bulk_ok && (int64_t)(tile + 1) * TM <= p.n_out) {
        if ((m_cur >> lane) & 1u) sh.act_k[b][__popc(m_cur & ((1u << lane) - 1))] = lane;
        if (lane == 0) sh.n_act[b] = __popc(m_cur);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_expect_tx(&sh.nbr_full[b], (uint32_t)NBR_INTS * 4u);
          bulk_g2s(smem_u32(nb), p.nbr + (int64_t)tile * NBR_INTS, (uint32_t)NBR_INTS * 4u, &sh.nbr_full[b]);
        }
        continue;
      }
      int n_act = 1;
      if (identity) {
        if (lane == 0) sh.act_k[b][0] = 0;
      } else {
        const uint32_t nb_u32 = smem_u32(nb);
        if (dbg & 8) {   // ablation: the neighbour tile is not loaded (combine with bit 0: the indices are garbage)
        } else if (p.perm && !nbr_sorted) {
          // the table is in its original order: tile row t is table row perm[tile*128 + t].  Each lane fetches its 4
          // rows (27 ints = 108 bytes each, 4-byte aligned) with 4-byte cp.async; rows past the end are zero-filled.
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int t = 4 * lane + q;
            const int64_t t_row = (int64_t)tile * TM + t;
            const bool in = t_row < p.n_out;
            const int32_t* src = p.nbr + (in ? (int64_t)__ldg(p.perm + t_row) * KVOL : 0);
#pragma unroll
            for (int k = 0; k < KVOL; ++k) cp_async4(nb_u32 + (t * KVOL + k) * 4, src + k, in ? 4u : 0u);
          }
        } else {
          const int64_t first = (int64_t)tile * NBR_INTS;                // int index of the tile's first table entry
          const int64_t total = p.n_out * (int64_t)KVOL;
          for (int ch = lane; ch < NBR_INTS / 4; ch += 32) {             // 16-byte chunks; the table slice is contiguous
            int64_t e = first + (int64_t)ch * 4;
            int64_t left = total - e;                                    // ints still inside the table
            uint32_t bytes = left >= 4 ? 16u : (left > 0 ? (uint32_t)left * 4u : 0u);
            cp_async16(nb_u32 + ch * 16, bytes ? (const void*)(p.nbr + e) : (const void*)p.nbr, bytes);
          }
        }
        cp_async_commit();
        unsigned m = p.tile_mask ? __ldg(p.tile_mask + tile) : 0u;       // populated offsets, precomputed at sort time
        cp_async_wait<0>();
        __syncwarp();
        const int64_t rows_left = p.n_out - (int64_t)tile * TM;
        const int rows = rows_left < TM ? (int)rows_left : TM;
        if (rows < TM) {   // rows past the table were zero-filled (= row 0): mark them missing so nothing is gathered.
          // Every lane patches exactly the words its OWN cp.async wrote (ordered by its own wait above): no cross-lane
          // write-after-write on the zero-filled chunks
          if (p.perm && !nbr_sorted) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int t = 4 * lane + q;
              if (t >= rows) for (int k = 0; k < KVOL; ++k) nb[t * KVOL + k] = -1;
            }
          } else {
            for (int ch = lane; ch < NBR_INTS / 4; ch += 32)
#pragma unroll
              for (int j = 0; j < 4; ++j) if (ch * 4 + j >= rows * KVOL) nb[ch * 4 + j] = -1;
          }
          __syncwarp();
        }
        if (!p.tile_mask) {   // no precomputed mask: scan the tile (one lane per offset)
          int f = 0;
          if (lane < KVOL) {
#pragma unroll 8
            for (int r = 0; r < rows; ++r) f |= (nb[r * KVOL + lane] >= 0);
          }
          m = __ballot_sync(0xffffffffu, f);
        }
        const int f = (m >> lane) & 1u;
        if (f) sh.act_k[b][__popc(m & ((1u << lane) - 1))] = lane;
        n_act = __popc(m);
      }
      if (lane == 0) sh.n_act[b] = n_act;
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.nbr_full[b]);
    }
  } else {
    // ======================================= epilogue: thread <-> output row ================================
    const int quarter = warp & 3;                      // TMEM lanes this warp may read
    uint32_t amax_bits = 0u;                           // fp16-range monitor: max |y| (float bits; NaN sorts above inf)
    // row ids two tiles ahead, first residual chunk one tile ahead (tc_epilogue.cuh: the epilogue is a per-tile serial chain)
    constexpr int kGroups = TcRoles<CTAS>::kEpiWarps / 4;         // epilogue groups, each drains every kGroups-th tile of the CTA
    const int grp = (warp - TcRoles<CTAS>::kEpiWarp0) >> 2;
    const int64_t g_step = (int64_t)kGroups * gridDim.x;
    EpiState<HALF> est;
    est.o_cur = epi_row_id(p, (int64_t)blockIdx.x + (int64_t)grp * gridDim.x, quarter * 32 + lane, num_tiles);
    est.o_next = epi_row_id(p, (int64_t)blockIdx.x + (int64_t)grp * gridDim.x + g_step, quarter * 32 + lane, num_tiles);
    epi_residual_first<COUT, HALF>(p, est.o_cur, est.rc);
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      if (lt % kGroups != grp) continue;
      const int ab = lt % NACC;
      const int64_t t2 = (int64_t)tile + 2 * g_step;
      tc_epilogue_tile<COUT, HALF, false, true, true>(p, tile, quarter, lane, normalize, &sh.acc_full[ab], (lt / NACC) & 1,
                                                      &sh.acc_empty[ab], &sh.acc_n_act[ab], tmem_base + (uint32_t)(ab * COUT),
                                                      amax_bits, dbg, &est, t2 < num_tiles ? (int)t2 : num_tiles, num_tiles);
    }
    if (p.range_mon) range_mon_flush(p.range_mon, amax_bits);   // saturation / NaN / tiny tensors are reported, never silent
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// W [K][cin][cout] (ME layout) -> tensor-core image: for every (k, 32-channel slab) one contiguous block of cout x 128 B
// that is byte-for-byte the SWIZZLE_128B K-major shared-memory tile (row n, 16-byte chunk j at (n/8)*1024 + (n%8)*128 +
// ((j ^ n%8)*16)), values rounded to nearest-even tf32.  One cp.async.bulk per pipeline stage then stages it.
// dgrad != 0: (cin, cout) are the DATA-GRADIENT convolution's widths and element (k, c, n) is read from the forward weights
// Wf [K][cout][cin] at (flip ? K-1-k : k, n, c)
__global__ void __launch_bounds__(256) weights_to_tc_kernel(const float* __restrict__ W, int K, int cin, int cout,
                                                            float* __restrict__ Wimg, int dgrad, int flip) {
  const int64_t total = (int64_t)K * cin * cout;
  const int slabs = cin / KSLAB;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    // dgrad: walk the elements with c fastest so that the reads of Wf (cin fastest there) stay coalesced
    const int n = dgrad ? (int)((e / cin) % cout) : (int)(e % cout);
    const int c = dgrad ? (int)(e % cin) : (int)((e / cout) % cin);
    const int k = (int)(e / ((int64_t)cout * cin));
    const int64_t src = dgrad ? ((int64_t)(flip ? K - 1 - k : k) * cout + n) * cin + c : e;
    uint32_t u = __float_as_uint(W[src]);
    u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;     // round to nearest even at 10 mantissa bits
    const int sl = c / KSLAB, cc = c % KSLAB, j = cc >> 2, w = cc & 3;
    const int64_t blk = ((int64_t)k * slabs + sl) * ((int64_t)cout * KSLAB);
    const int off = (n >> 3) * 256 + (n & 7) * 32 + ((j ^ (n & 7)) << 2) + w;
    Wimg[blk + off] = __uint_as_float(u);
  }
}

// fp16 image: per (k, 64-channel slab) one cout x 128 B block (row n = 64 halves, 16-byte chunk j at (j ^ n%8)),
// values rounded to nearest-even fp16 (saturating).
// slab == 32: per (k, 32-channel slab) one cout x 64 B block in the SWIZZLE_64B layout (16-byte chunk j at j ^ ((n/2)%4)).
__global__ void __launch_bounds__(256) weights_to_tc_f16_kernel(const float* __restrict__ W, int K, int cin, int cout,
                                                                int slab, __half* __restrict__ Wimg) {
  const int64_t total = (int64_t)K * cin * cout;
  const int slabs = cin / slab;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(e % cout);
    const int c = (int)((e / cout) % cin);
    const int k = (int)(e / ((int64_t)cout * cin));
    const int sl = c / slab, cc = c % slab, j = cc >> 3, w = cc & 7;
    const int64_t blk = ((int64_t)k * slabs + sl) * ((int64_t)cout * slab);
    const int off = slab == 64 ? (n >> 3) * 512 + (n & 7) * 64 + ((j ^ (n & 7)) << 3) + w     // in halves
                               : n * 32 + ((j ^ ((n >> 1) & 3)) << 3) + w;
    Wimg[blk + off] = __float2half_rn(fminf(fmaxf(W[e], -65504.f), 65504.f));
  }
}

template <int COUT, int KVOL, int MODE, int CTAS>
static int launch_tc_n(const ConvParams& p, int64_t n_in, cudaStream_t st) {
  using Cfg = TcCfg<COUT, MODE, CTAS>;
  constexpr size_t smem = (size_t)Cfg::STAGES * Cfg::STAGE + (size_t)Cfg::NBUF * TM * KVOL * 4;
  static_assert(smem + sizeof(TcShared) + 256 <= (CTAS == 2 ? 113u : 227u) * 1024u, "operand ring + neighbour tiles + staging exceed shared memory");
  auto kern = spconv_fwd_tc_kernel<COUT, KVOL, MODE, CTAS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("spconv_fwd_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  const int num_tiles = (int)((p.n_out + TM - 1) / TM);
  const int grid = num_tiles < kNumSMs * CTAS ? num_tiles : kNumSMs * CTAS;     // persistent: CTAS CTAs per SM
  CUtensorMap map0, map1;
  int rc = make_rows_tensor_map_ex(&map0, p.in0, n_in, p.c0, MODE, false);
  if (rc == GCLB_OK) rc = p.c1 ? make_rows_tensor_map_ex(&map1, p.in1, n_in, p.c1, MODE, false) : (map1 = map0, GCLB_OK);
  if (rc != GCLB_OK) return rc;
  // ablation switches for tools/tileprof.py (WRONG results on purpose): 1 no activation gathers, 2 no weight slabs, 4 no MMAs,
  // 8 no neighbour-table loads (use with 1), 16 no epilogue math / stores, 32 no epilogue stores
  static const int dbg = getenv("GCLB_TC_DBG") ? atoi(getenv("GCLB_TC_DBG")) : 0;
  kern<<<grid, TcRoles<CTAS>::kThreads, smem, st>>>(p, num_tiles, (p.relu >> 1) & 1, dbg, map0, map1);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spconv_fwd_tc: CUDA error: %s", cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  count_launches(1);
  return GCLB_OK;
}

template <int COUT, int KVOL, int MODE>
static int launch_tc(const ConvParams& p, int64_t n_in, cudaStream_t st) {
  // GCLB_TC_CTAS=2 selects two half-size CTAs per SM (COUT <= 128).  Measured on B200 (16-pair step, 641 k rows): 64->64
  // 194.8 vs 201.3 us, 32->32 173.3 vs 175.7 us, 128->128 392 vs 346 us (its ring shrinks to 2 stages) -- no net gain, so the
  // single-CTA layout stays the default; the variant is kept for A/B measurements (profiles/r02_conv_ablation.md)
  static const int want = getenv("GCLB_TC_CTAS") ? atoi(getenv("GCLB_TC_CTAS")) : 1;
  if constexpr (COUT <= 128) {
    if (want == 2) return launch_tc_n<COUT, KVOL, MODE, 2>(p, n_in, st);
  }
  return launch_tc_n<COUT, KVOL, MODE, 1>(p, n_in, st);
}

bool spconv_tc_supported(const ConvParams& p) {
  const int cin = p.c0 + p.c1;
  const int kch = ((p.relu & 8) && !(p.relu & 32)) ? 64 : KSLAB;     // fp16 operands: 64 channels per 128-byte row unless bit 5
  if ((p.relu & 32) && !(p.relu & 8)) return false;
  if (p.c0 % kch != 0 || p.c1 % kch != 0 || cin < kch) return false;
  if ((p.relu & 16) && ((p.relu >> 1) & 1)) return false;   // the fused L2 normalise writes fp32 descriptors
  if (!(p.cout == 32 || p.cout == 64 || p.cout == 128 || p.cout == 256)) return false;
  if (!(p.K == 27 || p.K == 1)) return false;
  if (((p.relu >> 1) & 1) && p.cout != 32) return false;   // fused L2 normalise needs the whole row in one TMEM read
  return true;
}

int spconv_fwd_tc(const ConvParams& p, int64_t n_in, cudaStream_t st) {
#define GCLB_TC_CASE(C)                                                                                \
  case C:                                                                                              \
    if (p.relu & 32) return p.K == 27 ? launch_tc<C, 27, 2>(p, n_in, st) : launch_tc<C, 1, 2>(p, n_in, st); \
    if (p.relu & 8) return p.K == 27 ? launch_tc<C, 27, 1>(p, n_in, st) : launch_tc<C, 1, 1>(p, n_in, st);   \
    return p.K == 27 ? launch_tc<C, 27, 0>(p, n_in, st) : launch_tc<C, 1, 0>(p, n_in, st);
  switch (p.cout) {
    GCLB_TC_CASE(32)
    GCLB_TC_CASE(64)
    GCLB_TC_CASE(128)
    GCLB_TC_CASE(256)
  }
#undef GCLB_TC_CASE
  return GCLB_ERR_UNSUPPORTED;
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_has_tcgen05(void) {
  // the tcgen05 / TMEM / TMA kernels exist for sm_100 only: answer for the CURRENT device instead of failing at launch
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (dev == cached_dev) return cached;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  cached_dev = dev;
  cached = (major == 10) ? 1 : 0;
  return cached;
}

int gclb_weights_to_tc(const float* W, int32_t K, int32_t cin, int32_t cout, float* Wt, void* stream) {
  GCLB_CHECK_ARG(W && Wt && K >= 1 && cin >= 1 && cout >= 1, "bad arguments");
  GCLB_CHECK_ARG(cin % KSLAB == 0 && cout % 8 == 0, "tensor-core image needs cin % 32 == 0 and cout % 8 == 0");
  int64_t total = (int64_t)K * cin * cout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  weights_to_tc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, K, cin, cout, Wt, 0, 0);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_weights_to_tc_dgrad(const float* W, int32_t K, int32_t cin, int32_t cout, int32_t flip, float* Wt, void* stream) {
  GCLB_CHECK_ARG(W && Wt && K >= 1 && cin >= 1 && cout >= 1, "bad arguments");
  GCLB_CHECK_ARG(cout % KSLAB == 0 && cin % 8 == 0, "data-gradient image needs cout % 32 == 0 and cin % 8 == 0");
  int64_t total = (int64_t)K * cin * cout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  // the data-gradient convolution maps cout -> cin channels
  weights_to_tc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, K, cout, cin, Wt, 1, flip ? 1 : 0);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_weights_to_tc_f16(const float* W, int32_t K, int32_t cin, int32_t cout, int32_t slab_channels, void* Wt,
                           void* stream) {
  GCLB_CHECK_ARG(W && Wt && K >= 1 && cin >= 1 && cout >= 1, "bad arguments");
  GCLB_CHECK_ARG(slab_channels == 64 || slab_channels == 32, "slab_channels must be 64 or 32");
  GCLB_CHECK_ARG(cin % slab_channels == 0 && cout % 8 == 0, "fp16 tensor-core image needs cin % slab_channels == 0 and cout % 8 == 0");
  int64_t total = (int64_t)K * cin * cout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  weights_to_tc_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(W, K, cin, cout, slab_channels,
                                                                                   reinterpret_cast<__half*>(Wt));
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
