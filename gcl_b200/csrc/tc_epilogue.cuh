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
template <int COUT, bool HALF, bool DUAL = false>
__device__ __forceinline__ void tc_epilogue_tile(const ConvParams& p, int tile, int quarter, int lane, int normalize,
                                                 uint64_t* acc_full, uint32_t acc_parity, uint64_t* acc_empty,
                                                 const int* acc_n_act, uint32_t tmem_acc, uint32_t& amax_bits, int dbg = 0) {
  const int row = quarter * 32 + lane;
  // everything that does not depend on the accumulator is fetched BEFORE waiting for it: the output row id and
  // the first 32 residual values, so their DRAM latency overlaps the tile's main loop
  const int64_t t_row = (int64_t)tile * TM + row;
  int64_t o = p.n_out;                                   // rows past the end are never stored
  if (t_row < p.n_out) o = p.perm ? (int64_t)__ldg(p.perm + t_row) : t_row;
  const bool live = o < p.n_out;
  // residual has the dtype of the inputs (it IS a block's input): 32 columns = 8 (fp32) or 4 (fp16) 16-byte loads
  constexpr int RV = HALF ? 4 : 8;
  constexpr int RES_B = HALF ? 2 : 4;
  const unsigned char* res = (p.residual && live) ? reinterpret_cast<const unsigned char*>(p.residual) + (size_t)o * COUT * RES_B
                                                   : nullptr;
  const bool out_half = (p.relu & 16) != 0;
  float4 rc[RV];
#pragma unroll
  for (int q = 0; q < RV; ++q) rc[q] = res ? __ldg(reinterpret_cast<const float4*>(res) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    if (!more) {                                        // last TMEM read of this tile: hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
    if (live && !(dbg & 16)) {
      float y[32];
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n0 + q)) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 sf = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n0 + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        y[q + 0] = fmaf(empty_tile ? 0.f : __uint_as_float(v[q + 0]), sc.x, sf.x) + r.x;
        y[q + 1] = fmaf(empty_tile ? 0.f : __uint_as_float(v[q + 1]), sc.y, sf.y) + r.y;
        y[q + 2] = fmaf(empty_tile ? 0.f : __uint_as_float(v[q + 2]), sc.z, sf.z) + r.z;
        y[q + 3] = fmaf(empty_tile ? 0.f : __uint_as_float(v[q + 3]), sc.w, sf.w) + r.w;
      }
      if (p.relu & 1) {
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
      if (out_half) {                               // fp16 activations for the next layer (saturating, round to nearest)
#pragma unroll
        for (int q = 0; q < 32; ++q) amax_bits = max(amax_bits, __float_as_uint(y[q]) & 0x7fffffffu);
        __half* dst = reinterpret_cast<__half*>(p.out) + (size_t)o * COUT + n0;
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = fminf(fmaxf(y[q + 2 * j], -65504.f), 65504.f);
            const float b = fminf(fmaxf(y[q + 2 * j + 1], -65504.f), 65504.f);
            const __half2 h = __floats2half2_rn(a, b);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(dst + q) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else {
        float* dst = reinterpret_cast<float*>(p.out) + (size_t)o * COUT + n0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(y[q], y[q + 1], y[q + 2], y[q + 3]);
      }
    }
#pragma unroll
    for (int q = 0; q < RV; ++q) rc[q] = rn[q];
  }
}

}  // namespace gclb
