// K3 with operand reuse inside the SM ("halo staging"): the tcgen05 sparse convolution for SAME-MAP 3x3x3 convolutions
// (the residual blocks of the ResUNet: 16 of the 22 convolution launches of a forward).
//
// The direct-gather kernel (spconv_tc.cu) fetches, for every populated kernel offset of a 128-row output tile, the 128
// neighbour rows from L2 again: every input row crosses the L2->SM fabric ~7x (ncu, round 1: 15x the tensor's bytes in TMA
// loads, latency-bound at the shared-memory capacity).  Neighbouring output rows share most of their input rows, so here
//   1. a precomputed HALO LIST per tile group (gclb_kmap_halo_build) names the DISTINCT input rows the group's offsets touch
//      (<= 384 rows); ONE set of TMA tile::gather4 instructions stages them in shared memory once per channel slab;
//   2. eight copy warps assemble the per-offset A operand (128 rows x 128 B, canonical K-major SWIZZLE_128B) from the staged
//      rows with 16-byte shared->shared moves (conflict-free: 8 threads move one row), indexed by a uint16 local-index table;
//   3. tcgen05.mma / TMEM accumulators / fused epilogue exactly as in spconv_tc.cu.
// L2->SM traffic per tile drops from (populated offsets x 16 KB) to (distinct rows x 128 B); the kernel becomes bound by
// shared-memory bandwidth (copy + MMA operand reads) instead of the fabric.
//
// Work item = (tile, group): the builder splits a tile's populated offsets into groups whose distinct rows fit the halo
// buffer (one group for most tiles).  All groups and slabs of a tile accumulate into the same TMEM accumulator.
//
// Warp roles (480 threads, one persistent CTA per SM):
//   0-7  copy warps (A assembly; warp 0 lane 0 also pulls the weight slab of the stage with one cp.async.bulk)
//   8    MMA issuer (one thread) + TMEM allocation        9   group-metadata prefetch (one cp.async.bulk per group)
//   10   halo loader (TMA gather4 of the distinct rows)     11-14 epilogue (thread <-> output row)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace gclb {

constexpr int HMAX = 384;                 // distinct input rows per group (halo buffer: 48 KB of 128-byte rows)
constexpr int HALO_MAXG = 16;             // groups per tile (a closed group holds > HMAX - 128 rows => <= 14)
// a group record (variable size, 16-byte granules, packed back to back in the caller's buffer):
//   header 64 B | int32 halo[n_halo rounded up to 4] | uint16 loc[n_off][128]
constexpr int HG_HDR = 64;
constexpr int HG_HALO_OFF = HG_HDR;
constexpr int HG_BYTES = HG_HDR + HMAX * 4 + 27 * TM * 2;     // largest record: 8512 bytes
constexpr int HG_STRIDE = 8576;                               // shared-memory slot of the metadata ring (67 * 128)
// worst case per tile: 27 offsets x 256 B of local indices + 3456 distinct rows (+ padding) x 4 B + 14 headers
constexpr int HG_TILE_WORST = 27 * 256 + (27 * TM + 3 * HALO_MAXG) * 4 + HALO_MAXG * HG_HDR;

struct HaloHdr {       // first 64 bytes of a group slot
  int32_t n_off, n_halo, tile, reserved;
  uint8_t offs[32];
  int32_t pad[4];
};
static_assert(sizeof(HaloHdr) == HG_HDR, "header layout");

// ------------------------------------------------------------------------------------------------------------------
// builder: one CTA per 128-row tile of the (bucket-sorted) kernel map
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t halo_hash(int v) { return ((uint32_t)v * 2654435761u) >> 22; }   // 10 bits

__global__ void __launch_bounds__(TM) halo_build_kernel(const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                        int64_t n_out, unsigned char* __restrict__ slots, int64_t capacity16,
                                                        int32_t* __restrict__ tile_groups, int32_t* __restrict__ tile_ngroups,
                                                        unsigned long long* __restrict__ counter, int32_t* __restrict__ status) {
  __shared__ int s_keys[1024];
  __shared__ uint16_t s_vals[1024];
  __shared__ int s_halo[HMAX];
  __shared__ __align__(16) uint16_t s_loc[27][TM];
  __shared__ int s_offs[32];
  __shared__ int s_warp[4];
  __shared__ long long s_slot;
  const int r = threadIdx.x, lane = r & 31, wid = r >> 5;
  const int tile = blockIdx.x;
  const int64_t t_row = (int64_t)tile * TM + r;
  int nb[27];
  {
    const int64_t row = t_row < n_out ? (perm ? (int64_t)__ldg(perm + t_row) : t_row) : -1;
#pragma unroll
    for (int k = 0; k < 27; ++k) nb[k] = row >= 0 ? __ldg(nbr + row * 27 + k) : -1;
  }
  for (int i = r; i < 1024; i += TM) s_keys[i] = -1;
  int count = 0, n_off = 0, n_groups = 0;
  __syncthreads();

  auto emit = [&]() {     // all threads; writes the group collected so far as one packed record
    const int c4 = (count + 3) & ~3;     // halo list padded with -1 (out of bounds for the TMA: zero rows, never referenced)
    const int size16 = (HG_HDR + c4 * 4 + n_off * 2 * TM) >> 4;
    if (r == 0) {
      long long off = (long long)atomicAdd(counter, (unsigned long long)size16);
      if (off + size16 > capacity16 || n_groups >= HALO_MAXG) { atomicOr(status, GCLB_ST_FULL); off = -1; }
      s_slot = off;
      if (n_groups < HALO_MAXG) {
        tile_groups[((int64_t)tile * HALO_MAXG + n_groups) * 2] = (int32_t)off;
        tile_groups[((int64_t)tile * HALO_MAXG + n_groups) * 2 + 1] = size16;
      }
    }
    __syncthreads();
    const long long off = s_slot;
    if (off >= 0) {
      unsigned char* g = slots + (size_t)off * 16;
      if (r < 16) {   // header: 16 ints
        int v = 0;
        if (r == 0) v = n_off;
        else if (r == 1) v = count;
        else if (r == 2) v = tile;
        else if (r >= 4 && r < 12) {
          const int b = (r - 4) * 4;
          v = (s_offs[b] & 0xff) | ((s_offs[b + 1] & 0xff) << 8) | ((s_offs[b + 2] & 0xff) << 16) | ((s_offs[b + 3] & 0xff) << 24);
        }
        reinterpret_cast<int*>(g)[r] = v;
      }
      for (int i = r; i < c4; i += TM) reinterpret_cast<int*>(g + HG_HALO_OFF)[i] = i < count ? s_halo[i] : -1;
      // loc [n_off][128] uint16, 256 bytes per offset: 16-byte stores
      const uint4* src = reinterpret_cast<const uint4*>(&s_loc[0][0]);
      uint4* dst = reinterpret_cast<uint4*>(g + HG_HALO_OFF + c4 * 4);
      for (int i = r; i < n_off * 16; i += TM) dst[i] = src[i];
    }
    ++n_groups;
    __syncthreads();
    for (int i = r; i < 1024; i += TM) s_keys[i] = -1;
    count = 0;
    n_off = 0;
    __syncthreads();
  };

#pragma unroll 1
  for (int k = 0; k < 27; ++k) {
    int v = -1;
#pragma unroll
    for (int q = 0; q < 27; ++q) v = (q == k) ? nb[q] : v;      // register select (nb stays in registers)
    if (__syncthreads_or(v >= 0) == 0) continue;                // offset not populated in this tile
    for (int attempt = 0; attempt < 2; ++attempt) {
      int idx = -1;
      if (v >= 0) {
        uint32_t h = halo_hash(v);
        while (true) {
          const int key = s_keys[h];
          if (key == v) { idx = s_vals[h]; break; }
          if (key == -1) break;
          h = (h + 1) & 1023u;
        }
      }
      const bool is_new = v >= 0 && idx < 0;
      // deterministic local indices: count + rank of the row among the new rows of this offset (tile-row order)
      const unsigned bal = __ballot_sync(0xffffffffu, is_new);
      if (lane == 0) s_warp[wid] = __popc(bal);
      __syncthreads();
      int base = 0, n_new = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) { const int c = s_warp[w]; base += (w < wid) ? c : 0; n_new += c; }
      __syncthreads();                                           // s_warp is reused by the next offset
      if (count + n_new > HMAX && attempt == 0) { emit(); continue; }   // close the group, retry this offset in a fresh one
      if (is_new) {
        idx = count + base + __popc(bal & ((1u << lane) - 1u));
        s_halo[idx] = v;
        uint32_t h = halo_hash(v);
        while (atomicCAS(&s_keys[h], -1, v) != -1) h = (h + 1) & 1023u;   // rows of one offset are distinct: no duplicates race
        s_vals[h] = (uint16_t)idx;
      }
      s_loc[n_off][r] = v >= 0 ? (uint16_t)idx : (uint16_t)0xFFFF;
      if (r == 0) s_offs[n_off] = k;
      count += n_new;
      ++n_off;
      __syncthreads();
      break;
    }
  }
  if (n_off > 0) emit();
  if (r == 0) tile_ngroups[tile] = n_groups < HALO_MAXG ? n_groups : HALO_MAXG;
}

// ------------------------------------------------------------------------------------------------------------------
// convolution
// ------------------------------------------------------------------------------------------------------------------
constexpr int kHCopyWarps = 8;
constexpr int kHMmaWarp = 8, kHMetaWarp = 9, kHHaloWarp = 10, kHEpiWarp0 = 11;
constexpr int kHThreads = 15 * 32;   // 480

template <int COUT, int MODE>
struct HaloCfg {
  static constexpr int ROWB = MODE == 2 ? 64 : 128;
  static constexpr int A_BYTES = TM * ROWB;
  static constexpr int B_BYTES = COUT * ROWB;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int HALO_BYTES = HMAX * ROWB;
  static constexpr int NH = 2, NM = 2;
  static constexpr int STAGES = MODE == 2 ? 8 : (COUT >= 256 ? 2 : (COUT >= 128 ? 3 : 4));
  static constexpr int NACC = COUT >= 256 ? 2 : 4;
  static constexpr int TMEM_COLS = NACC * COUT;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE + (size_t)NH * HALO_BYTES + (size_t)NM * HG_STRIDE;
};

struct HaloShared {
  uint64_t full[8], empty[8];
  uint64_t acc_full[4], acc_empty[4];
  uint64_t meta_full[2], meta_empty[2];
  uint64_t halo_full[2], halo_empty[2];
  uint32_t tmem_base;
  int acc_n_act[4];
};

// cycle accounting of the roles (debug builds of a run: GCLB_HALO_DBG bit 9), one row per CTA:
//  0 copy: wait meta  1 copy: wait halo  2 copy: wait empty  3 copy: copy+fence+arrive  4 copy: total
//  5 mma: wait meta   6 mma: wait acc_empty  7 mma: wait full  8 mma: total
//  9 epi: wait acc_full 10 epi: total  11 halo: wait meta 12 halo: wait halo_empty 13 halo: total 14 tiles 15 stages
__device__ unsigned long long g_halo_prof[kNumSMs][16];

struct HaloArgs {
  const unsigned char* slots;      // packed group records (16-byte granules)
  const int32_t* tile_groups;      // [tiles][HALO_MAXG][2] = (offset, size) of every group record, in granules
  const int32_t* tile_ngroups;
};

template <int COUT, int MODE>
__global__ void __launch_bounds__(kHThreads, 1) spconv_fwd_halo_kernel(ConvParams p, HaloArgs ha, int num_tiles, int normalize, int dbg,
                                                                      const __grid_constant__ CUtensorMap map0,
                                                                      const __grid_constant__ CUtensorMap map1) {
  using Cfg = HaloCfg<COUT, MODE>;
  constexpr int ROWB = Cfg::ROWB, A_BYTES = Cfg::A_BYTES, S = Cfg::STAGES, NACC = Cfg::NACC, NH = Cfg::NH, NM = Cfg::NM;
  constexpr int KCH = MODE == 1 ? 64 : 32;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ HaloShared sh;
  unsigned char* ring = smem_dyn;
  unsigned char* halo = smem_dyn + (size_t)S * Cfg::STAGE;               // [NH][HALO_BYTES], 1024-byte aligned
  unsigned char* meta = halo + (size_t)NH * Cfg::HALO_BYTES;             // [NM][HG_STRIDE]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slabs = (p.c0 + p.c1) / KCH;

  if (tid == 0) {
    if ((smem_u32(ring) & 1023u) != 0) { printf("gclb spconv_halo: operand ring not 1024-byte aligned\n"); __trap(); }
    for (int s = 0; s < S; ++s) { mbar_init(&sh.full[s], kHCopyWarps + 1); mbar_init(&sh.empty[s], 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(&sh.acc_full[b], 1); mbar_init(&sh.acc_empty[b], 4); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&sh.meta_full[b], 1);
      mbar_init(&sh.meta_empty[b], kHCopyWarps + 2);      // copy warps + MMA thread + halo loader
      mbar_init(&sh.halo_full[b], 1);
      mbar_init(&sh.halo_empty[b], kHCopyWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kHMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh.tmem_base;
  const uint32_t ring_u32 = smem_u32(ring), halo_u32 = smem_u32(halo);
  // warp-collective wait: one lane polls the barrier, the rest of the warp parks at the warp barrier
  auto wwait = [&](uint64_t* bar, uint32_t parity) {
    if (dbg & 64) { if (lane == 0) mbar_wait(bar, parity); __syncwarp(); }
    else mbar_wait(bar, parity);
  };

  if (warp < kHCopyWarps) {
    // ======================================= copy warps: A assembly from the staged rows =========================
    uint32_t it = 0, ih = 0, im = 0;
    const bool prof = (dbg & 512) && tid == 0;
    long long t_meta = 0, t_halo = 0, t_empty = 0, t_copy = 0, t_all = clock64(), t0 = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ng = __ldg(ha.tile_ngroups + tile);
      for (int g = 0; g < ng; ++g, ++im) {
        const int mb = im % NM;
        if (prof) t0 = clock64();
        wwait(&sh.meta_full[mb], (im / NM) & 1u);
        if (prof) t_meta += clock64() - t0;
        const unsigned char* mg = meta + (size_t)mb * HG_STRIDE;
        const HaloHdr* hdr = reinterpret_cast<const HaloHdr*>(mg);
        const int n_off = hdr->n_off;
        const uint16_t* loc = reinterpret_cast<const uint16_t*>(mg + HG_HALO_OFF + ((hdr->n_halo + 3) & ~3) * 4);
        for (int sl = 0; sl < slabs; ++sl, ++ih) {
          const int hb = ih % NH;
          if (prof) t0 = clock64();
          wwait(&sh.halo_full[hb], (ih / NH) & 1u);
          if (prof) t_halo += clock64() - t0;
          const uint32_t hbase = halo_u32 + hb * Cfg::HALO_BYTES;
          for (int j = 0; j < n_off; ++j, ++it) {
            const int stage = it % S;
            if (prof) t0 = clock64();
            wwait(&sh.empty[stage], ((it / S) & 1u) ^ 1u);
            if (prof) { const long long t1 = clock64(); t_empty += t1 - t0; t0 = t1; }
            const uint32_t a_s = ring_u32 + stage * Cfg::STAGE;
            if (tid == 0) {
              const int k = hdr->offs[j];
              if (dbg & 1) mbar_arrive(&sh.full[stage]);
              else {
              mbar_arrive_expect_tx(&sh.full[stage], Cfg::B_BYTES);
              bulk_g2s(a_s + A_BYTES, reinterpret_cast<const unsigned char*>(p.W) + ((size_t)k * slabs + sl) * Cfg::B_BYTES,
                       Cfg::B_BYTES, &sh.full[stage]);
              }
            }
            if (!(dbg & 2)) {
            const uint16_t* lj = loc + j * TM;
            if (MODE == 2) {           // 64-byte rows: 4 threads per row, 2 rounds of 64 rows
              const int c = tid & 3, r0 = tid >> 2;
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int r = r0 + 64 * i;
                const uint32_t li = lj[r];
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (li != 0xFFFFu) {
                  const uint32_t src = hbase + li * 64u + (((uint32_t)c ^ ((li >> 1) & 3u)) << 4);
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
                }
                const uint32_t dst = a_s + (uint32_t)r * 64u + (((uint32_t)c ^ (((uint32_t)r >> 1) & 3u)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
            } else {                   // 128-byte rows: 8 threads per row, 4 rounds of 32 rows
              const int c = tid & 7, r0 = tid >> 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = r0 + 32 * i;
                const uint32_t li = lj[r];
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (li != 0xFFFFu) {
                  const uint32_t src = hbase + li * 128u + (((uint32_t)c ^ (li & 7u)) << 4);
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src));
                }
                const uint32_t dst = a_s + (uint32_t)r * 128u + (((uint32_t)c ^ ((uint32_t)r & 7u)) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              }
            }
            }
            if (!(dbg & 8)) fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.full[stage]);
            if (prof) t_copy += clock64() - t0;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh.halo_empty[hb]);      // this warp no longer reads the staged rows
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.meta_empty[mb]);
      }
    }
    if (prof) {
      unsigned long long* o = g_halo_prof[blockIdx.x];
      o[0] = t_meta; o[1] = t_halo; o[2] = t_empty; o[3] = t_copy; o[4] = clock64() - t_all; o[15] = it;
    }
  } else if (warp == kHMmaWarp) {
    // ======================================= MMA issuer (one thread) =============================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(COUT);
      uint32_t it = 0, im = 0;
      int lt = 0;
      const bool prof = (dbg & 512) != 0;
      long long t_meta = 0, t_acc = 0, t_full = 0, t_all = clock64(), t0 = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int ab = lt % NACC;
        const int ng = __ldg(ha.tile_ngroups + tile);
        if (prof) t0 = clock64();
        mbar_wait(&sh.acc_empty[ab], ((lt / NACC) & 1) ^ 1);
        if (prof) t_acc += clock64() - t0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * COUT);
        int issued = 0;
        for (int g = 0; g < ng; ++g, ++im) {
          const int mb = im % NM;
          if (prof) t0 = clock64();
          mbar_wait(&sh.meta_full[mb], (im / NM) & 1u);
          if (prof) t_meta += clock64() - t0;
          const int n_off = reinterpret_cast<const HaloHdr*>(meta + (size_t)mb * HG_STRIDE)->n_off;
          mbar_arrive(&sh.meta_empty[mb]);
          const int n_iter = n_off * slabs;
          for (int i = 0; i < n_iter; ++i, ++it, ++issued) {
            const int stage = it % S;
            if (prof) t0 = clock64();
            mbar_wait(&sh.full[stage], (it / S) & 1u);
            if (prof) t_full += clock64() - t0;
            tc_fence_after();
            const uint32_t a_s = ring_u32 + stage * Cfg::STAGE;
            const uint32_t b_s = a_s + A_BYTES;
            if (!(dbg & 32)) {
              constexpr uint32_t HI = MODE == 2 ? kDescHiSw64 : kDescHiSw128;
              const uint32_t a_lo = desc_lo(a_s), b_lo = desc_lo(b_s);
              if (issued == 0) umma_f16_lo<HI, false>(d_tmem, a_lo, b_lo, idesc);
              else umma_f16_lo<HI, true>(d_tmem, a_lo, b_lo, idesc);
#pragma unroll
              for (int ks = 1; ks < ROWB / 32; ++ks) umma_f16_lo<HI, true>(d_tmem, a_lo + 2 * ks, b_lo + 2 * ks, idesc);
            }
            umma_commit(&sh.empty[stage]);
          }
        }
        *reinterpret_cast<volatile int*>(&sh.acc_n_act[ab]) = issued;
        if (issued > 0) umma_commit(&sh.acc_full[ab]);
        else mbar_arrive(&sh.acc_full[ab]);
      }
      if (prof) {
        unsigned long long* o = g_halo_prof[blockIdx.x];
        o[5] = t_meta; o[6] = t_acc; o[7] = t_full; o[8] = clock64() - t_all; o[14] = lt;
      }
    }
  } else if (warp == kHMetaWarp) {
    // ======================================= group metadata prefetch =============================================
    if (lane == 0) {
      uint32_t im = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ng = __ldg(ha.tile_ngroups + tile);
        for (int g = 0; g < ng; ++g, ++im) {
          const int mb = im % NM;
          const int2 rec = __ldg(reinterpret_cast<const int2*>(ha.tile_groups) + (int64_t)tile * HALO_MAXG + g);
          if (rec.x < 0 || rec.y * 16 > HG_BYTES) { printf("gclb spconv_halo: invalid group record (tile %d)\n", tile); __trap(); }
          mbar_wait(&sh.meta_empty[mb], ((im / NM) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sh.meta_full[mb], (uint32_t)rec.y * 16u);
          bulk_g2s(smem_u32(meta + (size_t)mb * HG_STRIDE), ha.slots + (size_t)rec.x * 16, (uint32_t)rec.y * 16u, &sh.meta_full[mb]);
        }
      }
    }
  } else if (warp == kHHaloWarp) {
    // ======================================= halo loader: the distinct rows, once per slab =======================
    uint32_t ih = 0, im = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ng = __ldg(ha.tile_ngroups + tile);
      for (int g = 0; g < ng; ++g, ++im) {
        const int mb = im % NM;
        wwait(&sh.meta_full[mb], (im / NM) & 1u);
        const unsigned char* mg = meta + (size_t)mb * HG_STRIDE;
        const int n4 = (reinterpret_cast<const HaloHdr*>(mg)->n_halo + 3) >> 2;
        const int4* rows4 = reinterpret_cast<const int4*>(mg + HG_HALO_OFF);
        for (int sl = 0; sl < slabs; ++sl, ++ih) {
          const int hb = ih % NH;
          wwait(&sh.halo_empty[hb], ((ih / NH) & 1u) ^ 1u);
          const uint32_t hbase = halo_u32 + hb * Cfg::HALO_BYTES;
          if (lane == 0) { if (dbg & 4) mbar_arrive(&sh.halo_full[hb]); else mbar_arrive_expect_tx(&sh.halo_full[hb], (uint32_t)n4 * 4u * ROWB); }
          __syncwarp();
          const int c = sl * KCH;
          for (int q = lane; q < ((dbg & 4) ? 0 : n4); q += 32) {
            const int4 r = rows4[q];
            if (c < p.c0) tma_gather4(hbase + q * (4 * ROWB), &map0, &sh.halo_full[hb], c, r.x, r.y, r.z, r.w);
            else tma_gather4(hbase + q * (4 * ROWB), &map1, &sh.halo_full[hb], c - p.c0, r.x, r.y, r.z, r.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.meta_empty[mb]);
      }
    }
  } else {
    // ======================================= epilogue ============================================================
    const int quarter = warp & 3;
    uint32_t amax_bits = 0u;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int ab = lt % NACC;
      tc_epilogue_tile<COUT, true>(p, tile, quarter, lane, normalize, &sh.acc_full[ab], (lt / NACC) & 1, &sh.acc_empty[ab],
                                   &sh.acc_n_act[ab], tmem_base + (uint32_t)(ab * COUT), amax_bits, dbg);
    }
    if (p.range_mon) range_mon_flush(p.range_mon, amax_bits);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kHMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int COUT, int MODE>
static int launch_halo(const ConvParams& p, const HaloArgs& ha, int64_t n_in, cudaStream_t st) {
  using Cfg = HaloCfg<COUT, MODE>;
  auto kern = spconv_fwd_halo_kernel<COUT, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) {
    set_error("spconv_fwd_halo: cannot reserve %zu bytes of shared memory: %s", (size_t)Cfg::SMEM, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  const int num_tiles = (int)((p.n_out + TM - 1) / TM);
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  CUtensorMap map0, map1;
  int rc = make_rows_tensor_map_ex(&map0, p.in0, n_in, p.c0, MODE, false);
  if (rc == GCLB_OK) rc = p.c1 ? make_rows_tensor_map_ex(&map1, p.in1, n_in, p.c1, MODE, false) : (map1 = map0, GCLB_OK);
  if (rc != GCLB_OK) return rc;
  static const int dbg = getenv("GCLB_HALO_DBG") ? atoi(getenv("GCLB_HALO_DBG")) : 0;   // ablation switches (wrong results!)
  kern<<<grid, kHThreads, Cfg::SMEM, st>>>(p, ha, num_tiles, (p.relu >> 1) & 1, dbg, map0, map1);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spconv_fwd_halo: CUDA error: %s", cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  count_launches(1);
  return GCLB_OK;
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_debug_halo_prof(unsigned long long* out_host /* [148][16] */) {
  cudaError_t e = cudaMemcpyFromSymbol(out_host, g_halo_prof, sizeof(g_halo_prof));
  return e == cudaSuccess ? GCLB_OK : GCLB_ERR_CUDA;
}

size_t gclb_kmap_halo_bytes(int64_t n_out) { return (size_t)((n_out + TM - 1) / TM) * HG_TILE_WORST + 16; }
int32_t gclb_kmap_halo_max_groups(void) { return HALO_MAXG; }

int gclb_kmap_halo_build(const int32_t* nbr, int64_t n_out, const int32_t* row_perm, void* records, size_t record_bytes,
                         int32_t* tile_groups, int32_t* tile_ngroups, uint64_t* counter, int32_t* status, void* stream) {
  GCLB_CHECK_ARG(nbr && records && tile_groups && tile_ngroups && counter && status, "bad arguments");
  GCLB_CHECK_ARG(((uintptr_t)records & 15) == 0 && record_bytes < ((size_t)1 << 34), "records must be 16-byte aligned and < 16 GiB");
  if (n_out == 0) return GCLB_OK;
  const int64_t tiles = (n_out + TM - 1) / TM;
  halo_build_kernel<<<(unsigned)tiles, TM, 0, (cudaStream_t)stream>>>(nbr, row_perm, n_out, (unsigned char*)records,
                                                                      (int64_t)(record_bytes >> 4), tile_groups, tile_ngroups,
                                                                      reinterpret_cast<unsigned long long*>(counter), status);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_spconv_fwd_halo(const void* in0, int32_t c0, const void* in1, int32_t c1, int64_t n_in, const void* W, int32_t cout,
                         const void* slots, const int32_t* tile_groups, const int32_t* tile_ngroups, const int32_t* row_perm,
                         const float* scale, const float* shift, const void* residual, int32_t flags, void* out,
                         int64_t n_out, void* stream) {
  GCLB_CHECK_ARG(W && slots && tile_groups && tile_ngroups && (n_out == 0 || (in0 && out)), "null pointer");
  GCLB_CHECK_ARG((c1 == 0) == (in1 == nullptr), "in1 / c1 mismatch");
  GCLB_CHECK_ARG((flags & 8) != 0, "the halo kernel runs on fp16 activations (flag bit 3)");
  GCLB_CHECK_ARG((flags & ~(1 | 2 | 8 | 16 | 32)) == 0, "flags: bit 0 ReLU, 1 L2-normalise, 3 fp16 in, 4 fp16 out, 5 32-channel rows");
  const int kch = (flags & 32) ? 32 : 64;
  GCLB_CHECK_ARG(c0 >= kch && c0 % kch == 0 && c1 % kch == 0, "channel counts must be multiples of the slab width");
  GCLB_CHECK_ARG(!((flags & 2) && (cout != 32 || (flags & 16))), "fused L2 normalise: cout == 32, fp32 output");
  if (n_out == 0) return GCLB_OK;
  ConvParams p{static_cast<const float*>(in0), static_cast<const float*>(in1), c0, c1, static_cast<const float*>(W), 27, cout,
               nullptr, row_perm, nullptr, scale, shift, static_cast<const float*>(residual), flags, static_cast<float*>(out), n_out};
  p.range_mon = current_range_monitor();
  HaloArgs ha{static_cast<const unsigned char*>(slots), tile_groups, tile_ngroups};
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & 32) {
    if (cout == 32) return launch_halo<32, 2>(p, ha, n_in, st);
    if (cout == 64) return launch_halo<64, 2>(p, ha, n_in, st);
  } else {
    if (cout == 32) return launch_halo<32, 1>(p, ha, n_in, st);
    if (cout == 64) return launch_halo<64, 1>(p, ha, n_in, st);
    if (cout == 128) return launch_halo<128, 1>(p, ha, n_in, st);
    if (cout == 256) return launch_halo<256, 1>(p, ha, n_in, st);
  }
  set_error("gclb_spconv_fwd_halo: cout=%d is not covered (32, 64, 128, 256)", cout);
  return GCLB_ERR_UNSUPPORTED;
}

}  // extern "C"
