// K3 wgrad: gW[k, c, n] = sum over pairs (i = nbr[o,k] >= 0) in[i, c] * gout[o, n]        (fp32 exact)
// grid = (row chunks, K, c-tiles * n-tiles).  A CTA scans its rows 256 at a time, compacts the populated ones
// (ballot + prefix), then contracts them 32 at a time into a 64 x 64 register tile; one atomicAdd per element
// per CTA merges the row chunks.
#include "common.cuh"

namespace gclb {

constexpr int WT = 64;          // tile edge in c and n
constexpr int WROWS = 32;       // rows per contraction step
constexpr int WCHUNK = 4096;    // rows per CTA

__global__ void __launch_bounds__(256) spconv_wgrad_kernel(const float* __restrict__ in, int cin,
                                                           const float* __restrict__ gout, int cout, int64_t n_out,
                                                           const int32_t* __restrict__ nbr, int K,
                                                           float* __restrict__ gW) {
  __shared__ __align__(16) float As[WROWS][WT + 4];
  __shared__ __align__(16) float Gs[WROWS][WT + 4];
  __shared__ int list_i[256], list_o[256];
  __shared__ int warp_cnt[8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = blockIdx.y;
  const int n_tiles = (cout + WT - 1) / WT;
  const int c0 = (blockIdx.z / n_tiles) * WT, n0 = (blockIdx.z % n_tiles) * WT;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r_begin = (int64_t)blockIdx.x * WCHUNK, r_end = min(n_out, r_begin + WCHUNK);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t w0 = r_begin; w0 < r_end; w0 += 256) {
    int64_t o = w0 + tid;
    int i = -1;
    if (o < r_end) i = nbr ? __ldg(&nbr[o * K + k]) : (int)o;
    unsigned m = __ballot_sync(0xffffffffu, i >= 0);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < wid) base += warp_cnt[w];
      total += warp_cnt[w];
    }
    if (i >= 0) {
      int pos = base + __popc(m & ((1u << lane) - 1));
      list_i[pos] = i;
      list_o[pos] = (int)(o - w0);
    }
    __syncthreads();
    for (int g0 = 0; g0 < total; g0 += WROWS) {
      const int R = min(WROWS, total - g0);
      // gather: 32 rows x 64 channels of `in` and of `gout` (zero padded)
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        int e = pass * 256 + tid;          // 512 float4 slots per operand
        int r = e / 16, v = (e % 16) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
        if (r < R) {
          const float* pa = in + (size_t)list_i[g0 + r] * cin + c0 + v;
          const float* pg = gout + (size_t)(w0 + list_o[g0 + r]) * cout + n0 + v;
          if ((cin & 3) == 0 && c0 + v + 3 < cin) a = __ldg(reinterpret_cast<const float4*>(pa));
          else {
            if (c0 + v + 0 < cin) a.x = __ldg(pa + 0);
            if (c0 + v + 1 < cin) a.y = __ldg(pa + 1);
            if (c0 + v + 2 < cin) a.z = __ldg(pa + 2);
            if (c0 + v + 3 < cin) a.w = __ldg(pa + 3);
          }
          if ((cout & 3) == 0 && n0 + v + 3 < cout) g = __ldg(reinterpret_cast<const float4*>(pg));
          else {
            if (n0 + v + 0 < cout) g.x = __ldg(pg + 0);
            if (n0 + v + 1 < cout) g.y = __ldg(pg + 1);
            if (n0 + v + 2 < cout) g.z = __ldg(pg + 2);
            if (n0 + v + 3 < cout) g.w = __ldg(pg + 3);
          }
        }
        *reinterpret_cast<float4*>(&As[r][v]) = a;
        *reinterpret_cast<float4*>(&Gs[r][v]) = g;
      }
      __syncthreads();
      for (int r = 0; r < R; ++r) {
        float4 a = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
        float4 g = *reinterpret_cast<const float4*>(&Gs[r][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int i2 = 0; i2 < 4; ++i2)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i2][j] = fmaf(av[i2], gv[j], acc[i2][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + ty * 4 + i;
    if (c >= cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < cout && acc[i][j] != 0.f) atomicAdd(&gW[((size_t)k * cin + c) * cout + n], acc[i][j]);
    }
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" int gclb_spconv_wgrad(const float* in, int32_t cin, int64_t n_in, const float* gout, int32_t cout,
                                 int64_t n_out, const int32_t* nbr, int32_t K, float* gW, void* stream) {
  GCLB_CHECK_ARG(gW && cin >= 1 && cout >= 1 && K >= 1, "bad arguments");
  GCLB_CHECK_ARG(nbr || K == 1, "nbr may be NULL only for K == 1");
  GCLB_CHECK_ARG(n_out == 0 || (in && gout), "null pointer");
  (void)n_in;
  if (n_out == 0) return GCLB_OK;
  dim3 grid((unsigned)((n_out + WCHUNK - 1) / WCHUNK), (unsigned)K,
            (unsigned)(((cin + WT - 1) / WT) * ((cout + WT - 1) / WT)));
  spconv_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, cin, gout, cout, n_out, nbr, K, gW);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}
