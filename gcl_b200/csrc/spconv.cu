// K3 (fp32 CUDA-core path): sparse convolution as an output-stationary implicit GEMM.
//   out[o,:] = act( (sum_k in[nbr[o,k],:] @ W[k]) * scale + shift + residual[o,:] )
// A CTA owns BM output rows x BN output channels; for every kernel offset that has at least one neighbour in
// the tile it gathers the BM x BK input slab into shared memory (zeros where nbr < 0), streams the matching
// BK x BN weight slab and accumulates in registers, so every output row is written exactly once (no
// atomics, deterministic) with BatchNorm(eval)/bias, residual add and ReLU fused into the store.
// This path is exact fp32 and also serves dgrad (conv with the transposed table / transposed weights).
// The tensor-core path (tcgen05 kind::tf32, accumulators in TMEM) lives in spconv_tc.cu.
#include <cuda_fp16.h>

#include "common.cuh"

namespace gclb {

constexpr int BK = 32;
constexpr int kConvThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// VEC: channel counts are multiples of 4 (all ResUNet layers) -> 128-bit gathers
template <int BM, int BN, bool VEC>
__global__ void __launch_bounds__(kConvThreads) spconv_fwd_f32_kernel(ConvParams p) {
  constexpr int TX = BN / 4;       // threads along N
  constexpr int APAD = BM + 4, BPAD = BN + 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);                 // [BK][APAD]   (transposed: channel-major)
  float* Bs = As + BK * APAD;                                     // [BK][BPAD]
  int* nbr_s = reinterpret_cast<int*>(Bs + BK * BPAD);            // [BM][K]
  int* act_k = nbr_s + BM * p.K;                                  // [K] list of active offsets
  __shared__ int n_act;

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t tile_m = (int64_t)blockIdx.x * BM;
  const int tile_n = blockIdx.y * BN;
  const int K = p.K, cin = p.c0 + p.c1;

  // ---- stage the neighbour tile, find which offsets are populated in this tile
  for (int k = tid; k < K; k += kConvThreads) act_k[k] = 0;
  __syncthreads();
  for (int e = tid; e < BM * K; e += kConvThreads) {
    int64_t o = tile_m + e / K;
    int v = -1;
    if (o < p.n_out) v = p.nbr ? __ldg(&p.nbr[tile_m * K + e]) : (int)o;
    nbr_s[e] = v;
    if (v >= 0) act_k[e % K] = 1;
  }
  __syncthreads();
  if (tid < 32) {  // warp 0 compacts the flags into an ascending list (keeps the fp32 summation order fixed)
    int base = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
      int k = k0 + tid;
      int f = (k < K) ? act_k[k] : 0;
      __syncwarp();
      unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) act_k[base + __popc(m & ((1u << tid) - 1))] = k;   // base+rank <= k: never clobbers an unread flag
      base += __popc(m);
      __syncwarp();
    }
    if (tid == 0) n_act = base;
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nact = n_act;
  for (int ai = 0; ai < nact; ++ai) {
    const int k = act_k[ai];
    const float* Wk = p.W + (size_t)k * cin * p.cout;
    for (int cs = 0; cs < cin; cs += BK) {
      // ---- gather A slab: BM rows x BK channels, stored channel-major
#pragma unroll
      for (int pass = 0; pass < BM / 32; ++pass) {
        int r = pass * 32 + tid / 8;
        int cv = (tid % 8) * 4;
        int idx = nbr_s[r * K + k];
        int c = cs + cv;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx >= 0) {
          if (VEC) {
            if (c < cin) v = (c < p.c0) ? ld4(p.in0 + (size_t)idx * p.c0 + c) : ld4(p.in1 + (size_t)idx * p.c1 + (c - p.c0));
          } else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              int cc = c + j;
              t[j] = (cc < cin) ? ((cc < p.c0) ? __ldg(p.in0 + (size_t)idx * p.c0 + cc)
                                               : __ldg(p.in1 + (size_t)idx * p.c1 + (cc - p.c0)))
                                : 0.f;
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        As[(cv + 0) * APAD + r] = v.x;
        As[(cv + 1) * APAD + r] = v.y;
        As[(cv + 2) * APAD + r] = v.z;
        As[(cv + 3) * APAD + r] = v.w;
      }
      // ---- weight slab: BK x BN
#pragma unroll
      for (int pass = 0; pass < (BK * BN / 4 + kConvThreads - 1) / kConvThreads; ++pass) {
        int e = pass * kConvThreads + tid;
        if (e < BK * BN / 4) {
          int kk = e / (BN / 4), nv = (e % (BN / 4)) * 4;
          int c = cs + kk, n = tile_n + nv;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < cin) {
            const float* src = Wk + (size_t)c * p.cout + n;
            if (VEC && n + 3 < p.cout) v = ld4(src);
            else {
              if (n + 0 < p.cout) v.x = __ldg(src + 0);
              if (n + 1 < p.cout) v.y = __ldg(src + 1);
              if (n + 2 < p.cout) v.z = __ldg(src + 2);
              if (n + 3 < p.cout) v.w = __ldg(src + 3);
            }
          }
          *reinterpret_cast<float4*>(&Bs[kk * BPAD + nv]) = v;
        }
      }
      __syncthreads();
      const int kmax = min(BK, cin - cs);
#pragma unroll 8
      for (int kk = 0; kk < kmax; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk * APAD + ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk * BPAD + tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- fused epilogue
  const int n0 = tile_n + tx * 4;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int n = n0 + j;
    sc[j] = (p.scale && n < p.cout) ? __ldg(p.scale + n) : 1.f;
    sh[j] = (p.shift && n < p.cout) ? __ldg(p.shift + n) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t o = tile_m + ty * 4 + i;
    if (o >= p.n_out) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaf(acc[i][j], sc[j], sh[j]);
    float* dst = p.out + (size_t)o * p.cout + n0;
    const float* res = p.residual ? p.residual + (size_t)o * p.cout + n0 : nullptr;
    if (VEC && n0 + 3 < p.cout) {
      if (res) { float4 r = ld4(res); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + j < p.cout) {
          float t = v[j] + (res ? __ldg(res + j) : 0.f);
          dst[j] = p.relu ? fmaxf(t, 0.f) : t;
        }
    }
  }
}

// ---- tiny-Cin layers (conv1: Cin = 1, K = 125, Cout = 32): HBM/L2-bound on the neighbour table, not a GEMM.
// One warp per output row; lanes = output channels; weights in shared memory; the warp reads 32 table entries
// at a time and walks the populated ones.
template <int CIN>
__global__ void __launch_bounds__(256) spconv_fwd_small_cin_kernel(ConvParams p) {
  extern __shared__ float Ws[];  // [K*CIN][cout]
  const int K = p.K, cout = p.cout;
  for (int e = threadIdx.x; e < K * CIN * cout; e += blockDim.x) Ws[e] = __ldg(p.W + e);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t o = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); o < p.n_out; o += (int64_t)gridDim.x * wpb) {
    for (int n0 = 0; n0 < cout; n0 += 32) {
      const int n = n0 + lane;
      float acc = 0.f;
      for (int k0 = 0; k0 < K; k0 += 32) {
        int k = k0 + lane;
        int idx = (k < K) ? (p.nbr ? __ldg(&p.nbr[o * K + k]) : (int)o) : -1;
        unsigned m = __ballot_sync(0xffffffffu, idx >= 0);
        while (m) {
          int src = __ffs(m) - 1;
          m &= m - 1;
          int i = __shfl_sync(0xffffffffu, idx, src);
          int kk = k0 + src;
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            float f = __ldg(p.in0 + (size_t)i * CIN + c);
            if (n < cout) acc = fmaf(f, Ws[(kk * CIN + c) * cout + n], acc);
          }
        }
      }
      if (n < cout) {
        float v = acc * (p.scale ? __ldg(p.scale + n) : 1.f) + (p.shift ? __ldg(p.shift + n) : 0.f);
        if (p.residual) v += __ldg(p.residual + (size_t)o * cout + n);
        p.out[(size_t)o * cout + n] = p.relu ? fmaxf(v, 0.f) : v;
      }
    }
  }
}

// Same layer with the kernel map FUSED: instead of reading a [n, K] neighbour table the warp probes the coordinate hash
// itself (32 offsets at a time).  For conv1 (K = 125) this removes the 4*K*n-byte table write + read (163 MB per 325k
// voxels) and its separate build pass; the probes hit the L2-resident hash table.
template <int CIN>
__global__ void __launch_bounds__(256, 4) spconv_fwd_probe_small_cin_kernel(ConvParams p, HashTable t,
                                                                          const int32_t* __restrict__ coords4, int ksize,
                                                                          int step, int32_t* __restrict__ nbr3,
                                                                          uint8_t* __restrict__ row_keys,
                                                                          uint32_t* __restrict__ row_masks,
                                                                          int32_t* __restrict__ key_hist, int64_t hist_blocks) {
  extern __shared__ float Ws[];  // [K*CIN][cout]
  const int K = p.K, cout = p.cout;
  for (int e = threadIdx.x; e < K * CIN * cout; e += blockDim.x) Ws[e] = __ldg(p.W + e);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int half = (ksize & 1) ? ksize / 2 : 0;
  const int reach = (ksize - 1) * step;
  const int lim = kAxisBias - reach;
  const bool quad5 = (ksize == 5) && (step == (1 << t.shift));
  // quad path: lane l probes item l (round 0) and item 32 + l (round 1, lanes 0-17) of the 2 quads x 25 columns
  int q_g[2], q_dy[2], q_dz[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int item = (r * 32 + lane) % 50, col = item % 25;
    q_g[r] = item / 25;
    q_dy[r] = col % 5 - 2;
    q_dz[r] = col / 5 - 2;
  }
  const int64_t o_first = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5), o_step = (int64_t)gridDim.x * wpb;
  int4 c_next = o_first < p.n_out ? __ldg(reinterpret_cast<const int4*>(coords4) + o_first) : make_int4(0, 0, 0, 0);
  uint32_t amax_bits = 0u;                        // fp16-range monitor (see common.cuh)
  for (int64_t o = o_first; o < p.n_out; o += o_step) {
    const int4 c = c_next;                        // the next row's coordinates are already in flight
    if (o + o_step < p.n_out) c_next = __ldg(reinterpret_cast<const int4*>(coords4) + o + o_step);
    const bool safe = (unsigned)c.x < 1023u && c.y >= -lim && c.y < lim && c.z >= -lim && c.z < lim && c.w >= -lim && c.w < lim;
    const uint64_t base = safe ? pack_key(c.x, c.y, c.z, c.w) : 0ull;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};          // output channels lane, lane+32, lane+64, lane+96
    unsigned key3 = 0, mask3 = 0;
    if (quad5 && safe) {
      // 5x5x5 kernel stepping by one cell of the table: the five x-neighbours of a column (dy, dz) live in exactly two
      // quads => 50 one-sector probes (2 quads x 25 columns, two rounds of lanes) instead of 125.  The row is
      // latency-bound (table probe -> feature load), so both rounds' probes are issued together, then all eight feature
      // loads, and only then the FMA loops.
      const int cx = (c.y + kAxisBias) >> t.shift;
      const int g0 = (cx - 2) >> 2;
      unsigned kbits = 0, mbits = 0;
      int v[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gx = g0 + q_g[r];
        const uint64_t gkey = ((uint64_t)(unsigned)c.x << 54) | ((uint64_t)(unsigned)gx << 36) |
                              ((uint64_t)(unsigned)(c.z + q_dy[r] * step + kAxisBias) << 18) |
                              (uint64_t)(unsigned)(c.w + q_dz[r] * step + kAxisBias);
        if (r == 0 || lane < 18) quad_find(t, gkey, v[r]);
        else v[r][0] = v[r][1] = v[r][2] = v[r][3] = -1;
      }
      int idx[2][4], kk_[2][4];
      float fv[2][4][CIN];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int gx = g0 + q_g[r];
        const int dy = q_dy[r], dz = q_dz[r];
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          const int j = (gx << 2) + sub - cx;                    // x delta of this cell, in cells
          const bool inwin = (r == 0 || lane < 18) && j >= -2 && j <= 2;
          idx[r][sub] = inwin ? v[r][sub] : -1;
          kk_[r][sub] = (j + 2) + 5 * (dy + 2) + 25 * (dz + 2);
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) fv[r][sub][ci] = idx[r][sub] >= 0 ? __ldg(p.in0 + (size_t)idx[r][sub] * CIN + ci) : 0.f;
          if (nbr3 && inwin && j >= -1 && j <= 1 && dy >= -1 && dy <= 1 && dz >= -1 && dz <= 1) {
            const int k3 = (j + 1) + 3 * (dy + 1) + 9 * (dz + 1);
            nbr3[o * 27 + k3] = idx[r][sub];
            if (idx[r][sub] >= 0) {
              mbits |= 1u << k3;
              kbits |= (j < 0 ? 1 : 0) | (j > 0 ? 2 : 0) | (dy < 0 ? 4 : 0) | (dy > 0 ? 8 : 0) | (dz < 0 ? 16 : 0) | (dz > 0 ? 32 : 0);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          unsigned m = __ballot_sync(0xffffffffu, idx[r][sub] >= 0);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int kk = __shfl_sync(0xffffffffu, kk_[r][sub], src);
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
              const float xv = __shfl_sync(0xffffffffu, fv[r][sub][ci], src);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int n = lane + 32 * q;
                if (n < cout) acc[q] = fmaf(xv, Ws[(kk * CIN + ci) * cout + n], acc[q]);
              }
            }
          }
        }
      }
      if (nbr3) {
        key3 = __reduce_or_sync(0xffffffffu, kbits);
        mask3 = __reduce_or_sync(0xffffffffu, mbits);
      }
    } else {
    auto round = [&](int k0, const LaneOffset& f, int k3) {
      const int k = k0 + lane;
      int idx = -1;
      if (k < K) {
        if (safe) idx = hash_find(t, base + (uint64_t)f.dkey);
        else {
          const int x = c.y + f.dx, y = c.z + f.dy, z = c.w + f.dz;
          if (coord_in_range(c.x, x, y, z)) idx = hash_find(t, pack_key(c.x, x, y, z));
        }
      }
      if (nbr3) {    // uniform branch
        if (k3 >= 0) nbr3[o * 27 + k3] = idx;
        const bool hit = k3 >= 0 && idx >= 0;
        key3 |= __reduce_or_sync(0xffffffffu, hit ? (unsigned)f.dirbits : 0u);
        mask3 |= __reduce_or_sync(0xffffffffu, hit ? (1u << k3) : 0u);
      }
      float fv[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) fv[ci] = idx >= 0 ? __ldg(p.in0 + (size_t)idx * CIN + ci) : 0.f;
      unsigned m = __ballot_sync(0xffffffffu, idx >= 0);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int kk = k0 + src;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float v = __shfl_sync(0xffffffffu, fv[ci], src);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = lane + 32 * q;
            if (n < cout) acc[q] = fmaf(v, Ws[(kk * CIN + ci) * cout + n], acc[q]);
          }
        }
      }
    };
    for (int k0 = 0; k0 < K; k0 += 32) {      // generic kernels: offsets recomputed per round (not the hot path)
      const int k = k0 + lane;
      const LaneOffset f = lane_offset(k < K ? k : 0, ksize, half, step);
      const int ax = f.dx / step, ay = f.dy / step, az = f.dz / step;
      const bool inner = k < K && ax >= -1 && ax <= 1 && ay >= -1 && ay <= 1 && az >= -1 && az <= 1;
      round(k0, f, inner ? (ax + 1) + 3 * (ay + 1) + 9 * (az + 1) : -1);
    }
    }
    if (nbr3 && lane == 0) {
      if (row_keys) row_keys[o] = (uint8_t)key3;
      if (row_masks) row_masks[o] = mask3;
      if (key_hist) atomicAdd(&key_hist[(int64_t)key3 * hist_blocks + (o >> 10)], 1);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int n = lane + 32 * q;
      if (n < cout) {
        float v = acc[q] * (p.scale ? __ldg(p.scale + n) : 1.f) + (p.shift ? __ldg(p.shift + n) : 0.f);
        if (p.residual) v += __ldg(p.residual + (size_t)o * cout + n);
        v = (p.relu & 1) ? fmaxf(v, 0.f) : v;
        if (p.relu & 16) {
          amax_bits = max(amax_bits, __float_as_uint(v) & 0x7fffffffu);
          reinterpret_cast<__half*>(p.out)[(size_t)o * cout + n] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        } else p.out[(size_t)o * cout + n] = v;
      }
    }
  }
  if (p.range_mon && (p.relu & 16)) range_mon_flush(p.range_mon, amax_bits);
}

template <int BM, int BN, bool VEC>
static cudaError_t launch_f32(const ConvParams& p, cudaStream_t st) {
  size_t smem = (size_t)(BK * (BM + 4) + BK * (BN + 4)) * 4 + (size_t)(BM * p.K + p.K) * 4;
  auto kern = spconv_fwd_f32_kernel<BM, BN, VEC>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((unsigned)((p.n_out + BM - 1) / BM), (unsigned)((p.cout + BN - 1) / BN));
  kern<<<grid, kConvThreads, smem, st>>>(p);
  return cudaGetLastError();
}

int spconv_fwd_tc(const ConvParams& p, int64_t n_in, cudaStream_t st);  // spconv_tc.cu; returns GCLB_ERR_UNSUPPORTED if shape not covered
bool spconv_tc_supported(const ConvParams& p);

// ---- fused pointwise tail: out = l2norm( relu([in0|in1] W1) W2 + bias )
constexpr int kTailRows = 64;
__global__ void __launch_bounds__(256) pointwise_tail_kernel(const float* __restrict__ in0, int c0,
                                                             const float* __restrict__ in1, int c1, int64_t n,
                                                             const float* __restrict__ W1, int cmid,
                                                             const float* __restrict__ W2,
                                                             const float* __restrict__ bias, int cout, int normalize,
                                                             float* __restrict__ out) {
  extern __shared__ float sm[];
  const int cin = c0 + c1;
  float* Xs = sm;                               // [64][cin+1]
  float* W1s = Xs + kTailRows * (cin + 1);      // [cin][cmid]
  float* Hs = W1s + cin * cmid;                 // [64][cmid+1]
  float* W2s = Hs + kTailRows * (cmid + 1);     // [cmid][cout]
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kTailRows;
  for (int e = tid; e < cin * cmid; e += 256) W1s[e] = __ldg(W1 + e);
  for (int e = tid; e < cmid * cout; e += 256) W2s[e] = __ldg(W2 + e);
  for (int e = tid; e < kTailRows * cin; e += 256) {
    int r = e / cin, c = e % cin;
    int64_t o = row0 + r;
    float v = 0.f;
    if (o < n) v = (c < c0) ? __ldg(in0 + (size_t)o * c0 + c) : __ldg(in1 + (size_t)o * c1 + (c - c0));
    Xs[r * (cin + 1) + c] = v;
  }
  __syncthreads();
  for (int e = tid; e < kTailRows * cmid; e += 256) {   // hidden = relu(X W1)
    int r = e / cmid, m = e % cmid;
    float a = 0.f;
    for (int c = 0; c < cin; ++c) a = fmaf(Xs[r * (cin + 1) + c], W1s[c * cmid + m], a);
    Hs[r * (cmid + 1) + m] = fmaxf(a, 0.f);
  }
  __syncthreads();
  // 4 threads per row, each cout/4 (strided) outputs; row norm by shuffle over the 4 threads
  const int r = tid / 4, q = tid % 4;
  const int64_t o = row0 + r;
  float ss = 0.f;
  float vals[32];
  int cnt = 0;
  for (int nidx = q; nidx < cout && cnt < 32; nidx += 4, ++cnt) {
    float a = bias ? __ldg(bias + nidx) : 0.f;
    for (int m = 0; m < cmid; ++m) a = fmaf(Hs[r * (cmid + 1) + m], W2s[m * cout + nidx], a);
    vals[cnt] = a;
    ss = fmaf(a, a, ss);
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  float inv = normalize ? 1.f / sqrtf(ss) : 1.f;
  if (o < n) {
    cnt = 0;
    for (int nidx = q; nidx < cout && cnt < 32; nidx += 4, ++cnt)
      out[(size_t)o * cout + nidx] = normalize ? vals[cnt] / sqrtf(ss) : vals[cnt];
  }
  (void)inv;
}

__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, int64_t total, int c,
                                                         const float* __restrict__ scale,
                                                         const float* __restrict__ shift,
                                                         const float* __restrict__ residual, int relu,
                                                         float* __restrict__ y) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int ch = (int)(e % c);
    float v = x[e] * (scale ? __ldg(scale + ch) : 1.f) + (shift ? __ldg(shift + ch) : 0.f);
    if (residual) v += residual[e];
    y[e] = relu ? fmaxf(v, 0.f) : v;
  }
}

static thread_local uint32_t* g_range_mon = nullptr;
uint32_t* current_range_monitor() { return g_range_mon; }

// one warp: fold the per-layer monitor words into the caller's status word
__global__ void range_check_kernel(const uint32_t* __restrict__ mons, int n_layers, int32_t* __restrict__ status) {
  int bits = 0;
  for (int l = threadIdx.x; l < n_layers; l += 32) {
    const uint32_t flags = mons[2 * l], amax = mons[2 * l + 1];
    if (flags & 1u) bits |= GCLB_ST_FP16_OVERFLOW;
    // the largest magnitude of a whole tensor sits below 2^-11: its entries are fp16 subnormals or close to it
    if (amax != 0u && amax < 0x3A000000u) bits |= GCLB_ST_FP16_UNDERFLOW;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, m);
  if (threadIdx.x == 0 && bits) atomicOr(status, bits);
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_spconv_set_range_monitor(uint32_t* mon) {
  g_range_mon = mon;
  return GCLB_OK;
}

int gclb_range_check(const uint32_t* mons, int32_t n_layers, int32_t* status, void* stream) {
  GCLB_CHECK_ARG(mons && status && n_layers >= 1, "bad arguments");
  range_check_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mons, n_layers, status);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_spconv_fwd(const void* in0_, int32_t c0, const void* in1_, int32_t c1, int64_t n_in, const void* W_,
                    int32_t K, int32_t cout, const int32_t* nbr, const int32_t* row_perm, const uint32_t* tile_mask,
                    const float* scale, const float* shift, const void* residual_, int32_t relu, void* out_,
                    int64_t n_out, int32_t algo, void* stream) {
  // fp32 unless flag bits 3 / 4 say fp16 (tcgen05 path); the kernels re-interpret
  const float* in0 = static_cast<const float*>(in0_);
  const float* in1 = static_cast<const float*>(in1_);
  const float* W = static_cast<const float*>(W_);
  const float* residual = static_cast<const float*>(residual_);
  float* out = static_cast<float*>(out_);
  GCLB_CHECK_ARG(W && (n_out == 0 || out), "null pointer");
  GCLB_CHECK_ARG(c0 >= 1 && c1 >= 0 && K >= 1 && cout >= 1, "bad shape");
  GCLB_CHECK_ARG((c1 == 0) == (in1 == nullptr), "in1 / c1 mismatch");
  GCLB_CHECK_ARG(nbr || K == 1, "nbr may be NULL only for K == 1");
  GCLB_CHECK_ARG(n_in == 0 || in0, "null input");
  GCLB_CHECK_ARG(algo >= 0 && algo <= 2, "algo must be 0, 1 or 2");
  GCLB_CHECK_ARG(relu >= 0 && relu <= 63,
                 "relu flags: bit 0 = ReLU, bit 1 = L2-normalise rows, bit 2 = nbr is the re-ordered copy, bit 3 = fp16 "
                 "inputs/residual/weights, bit 4 = fp16 output, bit 5 = 32-channel fp16 rows (bits 1-5: tcgen05 path)");
  GCLB_CHECK_ARG(algo == 2 || (relu & 62) == 0, "flag bits 1-5 exist on the tcgen05 path only");
  GCLB_CHECK_ARG((relu & 4) == 0 || row_perm != nullptr, "flag bit 2 needs row_perm");
  GCLB_CHECK_ARG(row_perm == nullptr || (algo == 2 && nbr != nullptr), "row_perm is a tcgen05-path option and needs nbr");
  GCLB_CHECK_ARG(tile_mask == nullptr || (algo == 2 && nbr != nullptr && K <= 32), "tile_mask is a tcgen05-path option");
  if (n_out == 0) return GCLB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ConvParams p{in0, in1, c0, c1, W, K, cout, nbr, row_perm, tile_mask, scale, shift, residual, relu, out, n_out};
  p.range_mon = current_range_monitor();
  const int cin = c0 + c1;
  cudaError_t e;
  if (algo == 2) {   // W is in the tensor-core layout [K][cout][cin]
    if (spconv_tc_supported(p)) return spconv_fwd_tc(p, n_in, st);
    set_error("gclb_spconv_fwd: shape (c0=%d, c1=%d, cout=%d, K=%d) is not covered by the tcgen05 kernel", c0, c1, cout, K);
    return GCLB_ERR_UNSUPPORTED;
  }
  if (c1 == 0 && cin <= 4 && (size_t)K * cin * cout * 4 <= 96 * 1024) {
    size_t smem = (size_t)K * cin * cout * 4;
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      int64_t blocks = (n_out + 7) / 8;
      if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
      kern<<<(unsigned)blocks, 256, smem, st>>>(p);
    };
    if (cin == 1) launch(spconv_fwd_small_cin_kernel<1>);
    else if (cin == 2) launch(spconv_fwd_small_cin_kernel<2>);
    else if (cin == 3) launch(spconv_fwd_small_cin_kernel<3>);
    else launch(spconv_fwd_small_cin_kernel<4>);
    e = cudaGetLastError();
  } else {
    bool vec = (c0 % 4 == 0) && (c1 % 4 == 0) && (cout % 4 == 0);
    if (cout <= 32) e = vec ? launch_f32<128, 32, true>(p, st) : launch_f32<128, 32, false>(p, st);
    else e = vec ? launch_f32<64, 64, true>(p, st) : launch_f32<64, 64, false>(p, st);
  }
  if (e != cudaSuccess) {
    set_error("gclb_spconv_fwd: CUDA error: %s", cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  count_launches(1);
  return GCLB_OK;
}

int gclb_spconv_fwd_probe(const float* in, int32_t cin, const float* W, int32_t ksize, int32_t cout, const void* table,
                          int64_t capacity, const int32_t* coords4, int64_t n, int32_t tensor_stride, int32_t dilation,
                          const float* scale, const float* shift, const float* residual, int32_t relu, void* out_,
                          int32_t* nbr3_out, uint8_t* row_keys, uint32_t* row_masks, int32_t* key_hist, void* stream) {
  float* out = static_cast<float*>(out_);
  GCLB_CHECK_ARG(W && table && (n == 0 || (in && coords4 && out)), "null pointer");
  GCLB_CHECK_ARG(cin >= 1 && cin <= 4 && cout >= 1 && cout <= 128, "fused-probe convolution covers cin <= 4, cout <= 128");
  GCLB_CHECK_ARG(ksize >= 1 && ksize <= 7 && tensor_stride >= 1 && dilation >= 1, "bad kernel geometry");
  GCLB_CHECK_ARG(capacity >= 2 && (capacity & (capacity - 1)) == 0, "bad capacity");
  GCLB_CHECK_ARG((relu & ~17) == 0, "relu flags: bit 0 = ReLU, bit 4 = fp16 output");
  GCLB_CHECK_ARG(!nbr3_out || ((ksize & 1) && ksize >= 3), "the 3x3x3 table can be emitted by an odd kernel >= 3 only");
  GCLB_CHECK_ARG(nbr3_out || (!row_keys && !row_masks && !key_hist), "row keys / masks / histogram describe nbr3_out");
  GCLB_CHECK_ARG(!key_hist || row_keys, "key_hist needs row_keys");
  if (n == 0) return GCLB_OK;
  const int K = ksize * ksize * ksize;
  const size_t smem = (size_t)K * cin * cout * 4;
  GCLB_CHECK_ARG(smem <= 200 * 1024, "weights do not fit in shared memory");
  ConvParams p{in, nullptr, cin, 0, W, K, cout, nullptr, nullptr, nullptr, scale, shift, residual, relu, out, n};
  p.range_mon = current_range_monitor();
  GCLB_CHECK_ARG(tensor_stride >= 1 && (tensor_stride & (tensor_stride - 1)) == 0, "tensor stride must be a power of two");
  HashTable t = make_table(table, capacity, tensor_stride);
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (n + 7) / 8;
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    kern<<<(unsigned)blocks, 256, smem, st>>>(p, t, coords4, ksize, tensor_stride * dilation, nbr3_out, row_keys, row_masks,
                                              key_hist, (n + 1023) / 1024);
  };
  if (cin == 1) launch(spconv_fwd_probe_small_cin_kernel<1>);
  else if (cin == 2) launch(spconv_fwd_probe_small_cin_kernel<2>);
  else if (cin == 3) launch(spconv_fwd_probe_small_cin_kernel<3>);
  else launch(spconv_fwd_probe_small_cin_kernel<4>);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_pointwise_tail(const float* in0, int32_t c0, const float* in1, int32_t c1, int64_t n, const float* W1,
                        int32_t cmid, const float* W2, const float* bias, int32_t cout, int32_t normalize, float* out,
                        void* stream) {
  GCLB_CHECK_ARG(in0 && W1 && W2 && (n == 0 || out), "null pointer");
  GCLB_CHECK_ARG((c1 == 0) == (in1 == nullptr), "in1 / c1 mismatch");
  GCLB_CHECK_ARG(cout <= 128, "cout > 128 unsupported");
  if (n == 0) return GCLB_OK;
  int cin = c0 + c1;
  size_t smem = ((size_t)kTailRows * (cin + 1) + (size_t)cin * cmid + (size_t)kTailRows * (cmid + 1) + (size_t)cmid * cout) * 4;
  GCLB_CHECK_ARG(smem <= 200 * 1024, "tail layer too wide for shared memory");
  cudaFuncSetAttribute(pointwise_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pointwise_tail_kernel<<<(unsigned)((n + kTailRows - 1) / kTailRows), 256, smem, (cudaStream_t)stream>>>(
      in0, c0, in1, c1, n, W1, cmid, W2, bias, cout, normalize, out);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_affine_act(const float* x, int64_t n, int32_t c, const float* scale, const float* shift,
                    const float* residual, int32_t relu, float* y, void* stream) {
  GCLB_CHECK_ARG(c >= 1 && (n == 0 || (x && y)), "bad arguments");
  if (n == 0) return GCLB_OK;
  int64_t total = n * c;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  affine_act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, total, c, scale, shift, residual, relu, y);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
