// K3 tensor-core path: sparse convolution as an output-stationary implicit GEMM on tcgen05 (sm_100a).
//
//   D[128 out rows, Cout] (fp32, TMEM)  +=  A[128 gathered rows, 32 ch] (smem)  x  B[Cout, 32 ch] (smem)
//   one pipeline stage per (populated kernel offset k, 32-channel slab); 4 x tcgen05.mma.kind::tf32 (K = 8) per stage
//
// * A is gathered straight into the canonical K-major SWIZZLE_128B layout with 16-byte cp.async (zero-fill for
//   missing neighbours): row r, 16-byte chunk j of a 128-byte row lands at (r/8)*1024 + (r%8)*128 + ((j ^ (r%8))*16).
// * B = W^T[k] ([Cout][Cin], pre-transposed + rounded to tf32 once per layer by gclb_weights_to_tc) uses the same
//   layout; it is re-read by every CTA and stays L2 resident.
// * warps 0-3: producers (gather + weights, 3 cp.async groups in flight per thread), later the epilogue
//   (tcgen05.ld 32 lanes x 32 columns, fused scale/shift/residual/ReLU, one output row per thread);
//   warp 4: TMEM allocation + the single MMA-issuing thread; smem slots recycle through tcgen05.commit -> mbarrier.
// * every output row is produced by exactly one CTA: no atomics, deterministic.
#include "common.cuh"

namespace gclb {

constexpr int TM = 128;               // output rows per CTA = TMEM lanes
constexpr int KSLAB = 32;             // channels per stage = one 128-byte swizzle row
constexpr int A_BYTES = TM * 128;     // 16 KB
constexpr int kProducerThreads = 128;
constexpr int kTcThreads = 160;
constexpr int kInFlight = 3;          // cp.async groups in flight per producer thread

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a protocol bug traps (-> CUDA error surfaced to the caller) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 24)) {
      printf("gclb spconv_tc: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4 = 1024 B between 8-row groups
//   [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int COUT>
struct TcCfg {
  static constexpr int B_BYTES = COUT * 128;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int STAGES = 4;   // 4 x (16 KB + COUT*128 B): 80 / 96 / 128 / 192 KB -> 2 CTAs per SM up to COUT = 64
  static constexpr int TMEM_COLS = COUT < 32 ? 32 : COUT;
};

template <int COUT>
__global__ void __launch_bounds__(kTcThreads) spconv_fwd_tc_kernel(ConvParams p) {
  using Cfg = TcCfg<COUT>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms repeat every 1024 B)
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  int* nbr_s = reinterpret_cast<int*>(ring + S * Cfg::STAGE);   // [TM][K]
  int* act_k = nbr_s + TM * p.K;                                // [K]
  __shared__ uint64_t full_bar[S], empty_bar[S], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ int n_act_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, cin = p.c0 + p.c1;
  const int64_t tile_m = (int64_t)blockIdx.x * TM;
  const int slabs = cin / KSLAB;

  // ---- setup: barriers, TMEM, neighbour tile, populated offsets
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], kProducerThreads); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int k = tid; k < K; k += kTcThreads) act_k[k] = 0;
  __syncthreads();
  for (int e = tid; e < TM * K; e += kTcThreads) {
    int64_t o = tile_m + e / K;
    int v = -1;
    if (o < p.n_out) v = p.nbr ? __ldg(&p.nbr[tile_m * K + e]) : (int)o;
    nbr_s[e] = v;
    if (v >= 0) act_k[e % K] = 1;
  }
  __syncthreads();
  if (warp == 0) {
    int base = 0;
    for (int k0 = 0; k0 < K; k0 += 32) {
      int k = k0 + lane;
      int f = (k < K) ? act_k[k] : 0;
      __syncwarp();
      unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) act_k[base + __popc(m & ((1u << lane) - 1))] = k;
      base += __popc(m);
      __syncwarp();
    }
    if (lane == 0) n_act_s = base;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int n_iter = n_act_s * slabs;
  const uint32_t ring_u32 = smem_u32(ring);

  if (warp < 4) {
    // ================= producers =================
    const int j = tid & 7;            // 16-byte chunk inside the 128-byte row
    const int r0 = tid >> 3;          // first row handled (then +16 per pass)
    for (int it = 0; it < n_iter; ++it) {
      const int stage = it % S;
      const uint32_t round = (uint32_t)(it / S);
      mbar_wait(&empty_bar[stage], (round & 1u) ^ 1u);    // passes immediately in round 0
      const int k = act_k[it / slabs];
      const int c = (it % slabs) * KSLAB;                 // first channel of the slab
      const float* src_base;
      int src_stride;
      if (c < p.c0) { src_base = p.in0 + c + j * 4; src_stride = p.c0; }
      else { src_base = p.in1 + (c - p.c0) + j * 4; src_stride = p.c1; }
      const uint32_t a_s = ring_u32 + stage * Cfg::STAGE;
      const uint32_t b_s = a_s + A_BYTES;
#pragma unroll
      for (int pass = 0; pass < TM / 16; ++pass) {
        const int r = r0 + pass * 16;
        const int idx = nbr_s[r * K + k];
        const uint32_t dst = a_s + (r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4);
        cp_async16(dst, idx >= 0 ? (const void*)(src_base + (size_t)idx * src_stride) : (const void*)p.in0, idx >= 0 ? 16u : 0u);
      }
      const float* wsrc = p.W + ((size_t)k * COUT) * cin + c + j * 4;   // W^T[k][n][c]
#pragma unroll
      for (int pass = 0; pass < COUT / 16; ++pass) {
        const int n = r0 + pass * 16;
        const uint32_t dst = b_s + (n >> 3) * 1024 + (n & 7) * 128 + ((j ^ (n & 7)) << 4);
        cp_async16(dst, wsrc + (size_t)n * cin, 16u);
      }
      cp_async_commit();
      if (it >= kInFlight - 1) {       // the group issued kInFlight-1 iterations ago has landed
        cp_async_wait<kInFlight - 1>();
        fence_proxy_async();
        mbar_arrive(&full_bar[(it - (kInFlight - 1)) % S]);
      }
    }
    // drain the last groups in order
    cp_async_wait<0>();
    fence_proxy_async();
    for (int it = max(0, n_iter - (kInFlight - 1)); it < n_iter; ++it) mbar_arrive(&full_bar[it % S]);

    // ================= epilogue: thread <-> output row =================
    const int row = tid;                                   // TMEM lane
    const int64_t o = tile_m + row;
    if (n_iter > 0) {
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int n0 = 0; n0 < COUT; n0 += 32) {
      uint32_t v[32];
      if (n_iter > 0) tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
      else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0u;
      }
      if (o < p.n_out) {
        float* dst = p.out + (size_t)o * COUT + n0;
        const float* res = p.residual ? p.residual + (size_t)o * COUT + n0 : nullptr;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n0 + q)) : make_float4(1.f, 1.f, 1.f, 1.f);
          float4 sh = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n0 + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 r = res ? __ldg(reinterpret_cast<const float4*>(res + q)) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 y;
          y.x = fmaf(__uint_as_float(v[q + 0]), sc.x, sh.x) + r.x;
          y.y = fmaf(__uint_as_float(v[q + 1]), sc.y, sh.y) + r.y;
          y.z = fmaf(__uint_as_float(v[q + 2]), sc.z, sh.z) + r.z;
          y.w = fmaf(__uint_as_float(v[q + 3]), sc.w, sh.w) + r.w;
          if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
          *reinterpret_cast<float4*>(dst + q) = y;
        }
      }
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    constexpr uint32_t idesc = make_idesc_tf32(COUT);
    for (int it = 0; it < n_iter; ++it) {
      const int stage = it % S;
      const uint32_t round = (uint32_t)(it / S);
      mbar_wait(&full_bar[stage], round & 1u);
      tc_fence_after();
      const uint32_t a_s = ring_u32 + stage * Cfg::STAGE;
      const uint32_t b_s = a_s + A_BYTES;
#pragma unroll
      for (int ks = 0; ks < KSLAB / 8; ++ks)      // 4 MMAs of K = 8 (32 bytes) inside the 128-byte swizzle row
        umma_tf32(tmem_base, make_desc_sw128(a_s + ks * 32), make_desc_sw128(b_s + ks * 32), idesc, (it | ks) ? 1u : 0u);
      umma_commit(&empty_bar[stage]);             // frees the slot when these MMAs have read it
    }
    if (n_iter > 0) umma_commit(&accum_bar);      // accumulator complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// W [K][cin][cout] (ME layout) -> Wt [K][cout][cin], rounded to nearest-even tf32
__global__ void __launch_bounds__(256) weights_to_tc_kernel(const float* __restrict__ W, int K, int cin, int cout,
                                                            float* __restrict__ Wt) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const int c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int c = c0 + r, n = n0 + tx;
    tile[r][tx] = (c < cin && n < cout) ? W[((size_t)k * cin + c) * cout + n] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int n = n0 + r, c = c0 + tx;
    if (n < cout && c < cin) {
      uint32_t u = __float_as_uint(tile[tx][r]);
      u = (u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;     // round to nearest even at 10 mantissa bits
      Wt[((size_t)k * cout + n) * cin + c] = __uint_as_float(u);
    }
  }
}

template <int COUT>
static int launch_tc(const ConvParams& p, cudaStream_t st) {
  using Cfg = TcCfg<COUT>;
  size_t smem = 1024 + (size_t)Cfg::STAGES * Cfg::STAGE + (size_t)(TM * p.K + p.K) * 4;
  auto kern = spconv_fwd_tc_kernel<COUT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("spconv_fwd_tc: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  kern<<<(unsigned)((p.n_out + TM - 1) / TM), kTcThreads, smem, st>>>(p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("spconv_fwd_tc: CUDA error: %s", cudaGetErrorString(e));
    return GCLB_ERR_CUDA;
  }
  count_launches(1);
  return GCLB_OK;
}

bool spconv_tc_supported(const ConvParams& p) {
  const int cin = p.c0 + p.c1;
  if (p.c0 % KSLAB != 0 || p.c1 % KSLAB != 0 || cin < KSLAB) return false;
  if (!(p.cout == 32 || p.cout == 64 || p.cout == 128 || p.cout == 256)) return false;
  if (p.K > 27) return false;     // neighbour tile must fit next to the operand ring
  return true;
}

int spconv_fwd_tc(const ConvParams& p, int64_t, cudaStream_t st) {
  switch (p.cout) {
    case 32: return launch_tc<32>(p, st);
    case 64: return launch_tc<64>(p, st);
    case 128: return launch_tc<128>(p, st);
    case 256: return launch_tc<256>(p, st);
  }
  return GCLB_ERR_UNSUPPORTED;
}

bool nn_tc_supported(int) { return false; }
int nn_tc(const float*, const float*, int, const int64_t*, const int64_t*, int, int64_t, int64_t, unsigned long long*,
          unsigned long long*, cudaStream_t) {
  return GCLB_ERR_UNSUPPORTED;
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_has_tcgen05(void) { return 1; }

int gclb_weights_to_tc(const float* W, int32_t K, int32_t cin, int32_t cout, float* Wt, void* stream) {
  GCLB_CHECK_ARG(W && Wt && K >= 1 && cin >= 1 && cout >= 1, "bad arguments");
  dim3 grid((unsigned)((cout + 31) / 32), (unsigned)((cin + 31) / 32), (unsigned)K);
  weights_to_tc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, K, cin, cout, Wt);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
