// Epilogue of the tcgen05 sparse-conv kernels (spconv_tc.cu, spconv_halo.cu): one 128-row output tile, thread <-> output row.
//   tcgen05.ld the fp32 accumulator -> y = acc * scale + shift (+ residual) -> ReLU -> (L2 normalise) -> fp16 or fp32 store
#pragma once
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace gclb {

// acc_full / acc_empty: the accumulator's mbarriers; acc_n_act: #populated offsets of the tile it holds (0 => zeros);
// tmem_acc: TMEM address of the accumulator (lane 0, first column); amax_bits: running fp16-range monitor (common.cuh)
// DUAL: two MMA issuers accumulated alternate pipeline stages into two accumulators (tmem_acc and tmem_acc + COUT; acc_n_act
// is int[2]: stages each one issued, 0 => that accumulator holds stale data); the epilogue adds them in a fixed order.
// ZERO: the accumulator is handed back ZEROED (tcgen05.st after every chunk is read): the direct kernel's row-masked MMAs
// always accumulate and never touch the rows whose neighbour is missing.
// PIPE: the four epilogue warps drain ONE tile at a time, so whatever latency a tile's epilogue exposes is paid once per tile
// by the whole CTA.  Fetching the output row id (perm) and then the residual row (address depends on it) inside the tile's own
// epilogue is two dependent L2/DRAM round trips = 2-3 us per tile -- more than the tile's MMAs.  With PIPE the caller keeps an
// EpiState across tiles: row ids are fetched two tiles ahead, the first residual chunk one tile ahead.
template <bool HALF>
struct EpiState {
  static constexpr int RV = HALF ? 4 : 8;   // 16-byte vectors per 32 residual columns
  int64_t o_cur, o_next;                    // output row of this thread in the current / next tile (n_out = none)
  float4 rc[RV];                            // first 32 residual columns of the current tile
};
__device__ __forceinline__ int64_t epi_row_id(const ConvParams& p, int64_t tile, int row, int num_tiles) {
  const int64_t t_row = tile * TM + row;
  if (tile >= num_tiles || t_row >= p.n_out) return p.n_out;
  return p.perm ? (int64_t)__ldg(p.perm + t_row) : t_row;
}
template <int COUT, bool HALF>
__device__ __forceinline__ void epi_residual_first(const ConvParams& p, int64_t o, float4* rc) {
  constexpr int RV = HALF ? 4 : 8;
  const bool have = p.residual && o < p.n_out;
  const float4* res = reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(p.residual) + (size_t)o * COUT * (HALF ? 2 : 4));
#pragma unroll
  for (int q = 0; q < RV; ++q) rc[q] = have ? __ldg(res + q) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int COUT, bool HALF, bool DUAL = false, bool ZERO = false, bool PIPE = false>
__device__ __forceinline__ void tc_epilogue_tile(const ConvParams& p, int tile, int quarter, int lane, int normalize,
                                                 uint64_t* acc_full, uint32_t acc_parity, uint64_t* acc_empty,
                                                 const int* acc_n_act, uint32_t tmem_acc, uint32_t& amax_bits, int dbg = 0,
                                                 EpiState<HALF>* st = nullptr, int tile_next2 = 0, int num_tiles = 0) {
  const int row = quarter * 32 + lane;
  // everything that does not depend on the accumulator is fetched BEFORE waiting for it: the output row id and
  // the first 32 residual values, so their DRAM latency overlaps the tile's main loop
  const int64_t t_row = (int64_t)tile * TM + row;
  int64_t o = p.n_out;                                   // rows past the end are never stored
  // residual has the dtype of the inputs (it IS a block's input): 32 columns = 8 (fp32) or 4 (fp16) 16-byte loads
  constexpr int RV = HALF ? 4 : 8;
  constexpr int RES_B = HALF ? 2 : 4;
  float4 rc[RV];
  int64_t o_next2 = 0;
  float4 rc_next[RV];
  if constexpr (PIPE) {
    o = st->o_cur;
#pragma unroll
    for (int q = 0; q < RV; ++q) rc[q] = st->rc[q];
    o_next2 = epi_row_id(p, tile_next2, row, num_tiles);          // consumed two tiles from now
    epi_residual_first<COUT, HALF>(p, st->o_next, rc_next);      // its row id was requested a whole tile ago
  } else {
    if (t_row < p.n_out) o = p.perm ? (int64_t)__ldg(p.perm + t_row) : t_row;
  }
  const bool live = o < p.n_out;
  const unsigned char* res = (p.residual && live) ? reinterpret_cast<const unsigned char*>(p.residual) + (size_t)o * COUT * RES_B
                                                   : nullptr;
  const bool out_half = (p.relu & 16) != 0;
  if constexpr (!PIPE) {
#pragma unroll
    for (int q = 0; q < RV; ++q) rc[q] = res ? __ldg(reinterpret_cast<const float4*>(res) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  mbar_wait(acc_full, acc_parity);
  tc_fence_after();
  const bool has0 = *reinterpret_cast<const volatile int*>(acc_n_act) != 0;
  const bool has1 = DUAL && *reinterpret_cast<const volatile int*>(acc_n_act + 1) != 0;
  const bool empty_tile = !has0 && !has1;
  const uint32_t t_addr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
  for (int n0 = 0; n0 < COUT; n0 += 32) {
    uint32_t v[32];
    tmem_ld32(t_addr + (uint32_t)n0, v);
    if constexpr (DUAL) {
      uint32_t v2[32];
      tmem_ld32(t_addr + (uint32_t)(COUT + n0), v2);
#pragma unroll
      for (int q = 0; q < 32; ++q)
        v[q] = __float_as_uint((has0 ? __uint_as_float(v[q]) : 0.f) + (has1 ? __uint_as_float(v2[q]) : 0.f));
    }
    float4 rn[RV];                                      // residual of the NEXT 32 columns, in flight during this chunk
    const bool more = (n0 + 32 < COUT);
#pragma unroll
    for (int q = 0; q < RV; ++q)
      rn[q] = (more && res) ? __ldg(reinterpret_cast<const float4*>(res + (size_t)(n0 + 32) * RES_B) + q)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (ZERO) tmem_st32_zero(t_addr + (uint32_t)n0);
    if (!more) {                                        // last TMEM read of this tile: hand the accumulator back
      if constexpr (ZERO) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (live && !(dbg & 16)) {
      // ~1000 instructions per thread and tile made the epilogue the tile-rate limit of the narrow layers (profiles/
      // r02_conv_ablation.md), so the fp16 path is kept lean: no residual arithmetic when there is no residual, ReLU +
      // saturation + packing in ONE conversion per column pair (cvt.rn[.relu].satfinite.f16x2.f32), the range monitor as a
      // 3-input NaN-propagating max over pre-saturation magnitudes (16 instead of 64 instructions per 32 columns).
      float y[32];
      const bool has_res = res != nullptr;
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n0 + q)) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 sf = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n0 + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
        // ZERO: an empty tile's accumulator already holds zeros
        const float a0 = (!ZERO && empty_tile) ? 0.f : __uint_as_float(v[q + 0]), a1 = (!ZERO && empty_tile) ? 0.f : __uint_as_float(v[q + 1]);
        const float a2 = (!ZERO && empty_tile) ? 0.f : __uint_as_float(v[q + 2]), a3 = (!ZERO && empty_tile) ? 0.f : __uint_as_float(v[q + 3]);
        y[q + 0] = fmaf(a0, sc.x, sf.x);
        y[q + 1] = fmaf(a1, sc.y, sf.y);
        y[q + 2] = fmaf(a2, sc.z, sf.z);
        y[q + 3] = fmaf(a3, sc.w, sf.w);
      }
      if (has_res) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float4 r;
          if constexpr (HALF) {   // 8 halves per 16-byte vector: columns q..q+3 are the low or high half of vector q / 8
            const float4 raw = rc[q >> 3];
            const uint32_t w0 = __float_as_uint((q & 4) ? raw.z : raw.x), w1 = __float_as_uint((q & 4) ? raw.w : raw.y);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&w0));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&w1));
            r = make_float4(lo.x, lo.y, hi.x, hi.y);
          } else {
            r = rc[q >> 2];
          }
          y[q + 0] += r.x; y[q + 1] += r.y; y[q + 2] += r.z; y[q + 3] += r.w;
        }
      }
      const bool relu = (p.relu & 1) != 0;
      if (out_half && !(dbg & 32)) {                // fp16 activations for the next layer (saturating, round to nearest)
        float amax_f = 0.f;
        uint32_t w[16];
        if (relu) {
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(amax_f) : "f"(y[q]), "f"(y[q + 1]));       // negatives become 0 below
            asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[q >> 1]) : "f"(y[q + 1]), "f"(y[q]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(amax_f) : "f"(fabsf(y[q])), "f"(fabsf(y[q + 1])));
            asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[q >> 1]) : "f"(y[q + 1]), "f"(y[q]));
          }
        }
        amax_bits = max(amax_bits, __float_as_uint(amax_f) & 0x7fffffffu);    // NaN (0x7fffffff) sorts above inf
        __half* dst = reinterpret_cast<__half*>(p.out) + (size_t)o * COUT + n0;
#pragma unroll
        for (int q = 0; q < 16; q += 4) *reinterpret_cast<uint4*>(dst + 2 * q) = make_uint4(w[q], w[q + 1], w[q + 2], w[q + 3]);
      } else {
        if (relu) {
#pragma unroll
          for (int q = 0; q < 32; ++q) y[q] = fmaxf(y[q], 0.f);
        }
        if (COUT == 32 && normalize) {               // F / ||F||_2 per row (model/resunet.py:226-230, no eps)
          float ss = 0.f;
#pragma unroll
          for (int q = 0; q < 32; ++q) ss = fmaf(y[q], y[q], ss);
          const float nrm = sqrtf(ss);
#pragma unroll
          for (int q = 0; q < 32; ++q) y[q] = y[q] / nrm;
        }
        if (dbg & 32) {                             // ablation: math only (kept alive through the range monitor)
#pragma unroll
          for (int q = 0; q < 32; ++q) amax_bits = max(amax_bits, __float_as_uint(y[q]) & 0x7fffffffu);
        } else {
          float* dst = reinterpret_cast<float*>(p.out) + (size_t)o * COUT + n0;
#pragma unroll
          for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RV; ++q) rc[q] = rn[q];
  }
  if constexpr (PIPE) {
    st->o_cur = st->o_next;
    st->o_next = o_next2;
#pragma unroll
    for (int q = 0; q < RV; ++q) st->rc[q] = rc_next[q];
  }
}

}  // namespace gclb
