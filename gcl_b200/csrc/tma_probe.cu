// Debug / bring-up entry point: one TMA tile::gather4 load of 4 rows x 32 fp32 of a [n, c] matrix into shared memory
// (SWIZZLE_128B), dumped back to global memory.  Used to pin down the tensor-map parameters the gather4 mode needs.
#include "tc_common.cuh"

namespace gclb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_rows_tensor_map_any(CUtensorMap* map, const void* base, int64_t n, int c, int box_rows, int mode, bool atom32);

// 2-D fp32 tensor map over a row-major [n, c] matrix, box = box_rows x 32 channels; atom32 selects
// SWIZZLE_128B_ATOM_32B (32-byte chunks XOR row % 4: the only layout tcgen05 accepts for MN-major tf32 operands)
// instead of the plain SWIZZLE_128B (16-byte chunks XOR row % 8) the K-major operands use.
int make_rows_tensor_map_sw(CUtensorMap* map, const float* base, int64_t n, int c, int box_rows, bool atom32) {
  return make_rows_tensor_map_any(map, base, n, c, box_rows, 0, atom32);
}
int make_rows_tensor_map_ex(CUtensorMap* map, const void* base, int64_t n, int c, int mode, bool atom32) {
  return make_rows_tensor_map_any(map, base, n, c, 1, mode, atom32);
}

int make_rows_tensor_map_any(CUtensorMap* map, const void* base, int64_t n, int c, int box_rows, int mode, bool atom32) {
  const bool half = mode != 0;
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return GCLB_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)c, (cuuint64_t)n};
  cuuint64_t gstride[1] = {(cuuint64_t)c * (half ? 2 : 4)};
  cuuint32_t box[2] = {mode == 1 ? 64u : 32u, (cuuint32_t)box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mode == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return GCLB_ERR_CUDA; }
  return GCLB_OK;
}

int make_rows_tensor_map(CUtensorMap* map, const float* base, int64_t n, int c, int box_rows) {
  return make_rows_tensor_map_sw(map, base, n, c, box_rows, false);
}

__global__ void __launch_bounds__(128) tma_gather4_probe_kernel(const __grid_constant__ CUtensorMap map, int col, int r0, int r1,
                                                                int r2, int r3, float* out /* [256] */) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  float* s = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = -777.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_proxy_async();
    mbar_arrive_expect_tx(&bar, 4 * 128);
    tma_gather4(smem_u32(smem), &map, &bar, col, r0, r1, r2, r3);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = s[i];
}

}  // namespace gclb

using namespace gclb;

extern "C" int gclb_debug_tma_gather4(const float* X, int64_t n, int32_t c, int32_t box_rows, int32_t col, const int32_t* rows4_host,
                                      float* out256, void* stream) {
  GCLB_CHECK_ARG(X && rows4_host && out256 && c % 32 == 0, "bad arguments");
  CUtensorMap map;
  int rc = make_rows_tensor_map(&map, X, n, c, box_rows);
  if (rc != GCLB_OK) return rc;
  tma_gather4_probe_kernel<<<1, 128, 2048, (cudaStream_t)stream>>>(map, col, rows4_host[0], rows4_host[1], rows4_host[2],
                                                                   rows4_host[3], out256);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

// bring-up micro-benchmark: issue rate of tcgen05.mma (M = 128, kind::f16, K = 16) from one thread, operands resident in
// shared memory, `n_mma` instructions accumulating into one TMEM tile; variants: commit after every `per_commit` MMAs and
// wait for it (per_commit <= 0: one commit at the end).  out[0] = cycles, out[1] = cycles of the issue loop alone.
namespace gclb {
template <int N>
__global__ void __launch_bounds__(128) umma_rate_kernel(int n_mma, int per_commit, int n_acc, int elect, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // `elect` = number of issuing warps (1..4): warp w < elect issues n_mma MMAs from its lane 0 into its own accumulator
  const int n_issuers = elect < 1 ? 1 : elect;
  const int wid = threadIdx.x >> 5;
  uint64_t& bar = bars[wid];
  if ((threadIdx.x & 31) == 0 && wid < n_issuers) {
    const uint32_t tb = tmem_base + (uint32_t)(wid * N * n_acc);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
    constexpr uint32_t idesc = make_idesc_f16(N);
    uint32_t phase = 0;
    const long long t0 = clock64();
    long long t_issue = 0;
    const uint32_t a_lo = desc_lo(a_s), b_lo = desc_lo(b_s);
    const int groups = n_mma / 4;
    for (int g = 0; g < groups; ++g) {       // per_commit: 0 = no commit inside, 1 = commit (no wait) after every 4 MMAs, 2 = commit + wait
      const uint32_t d = tb + (uint32_t)((g & (n_acc - 1)) * N);
      umma_f16_lo<kDescHiSw128, true>(d, a_lo, b_lo, idesc);
      umma_f16_lo<kDescHiSw128, true>(d, a_lo + 2, b_lo + 2, idesc);
      umma_f16_lo<kDescHiSw128, true>(d, a_lo + 4, b_lo + 4, idesc);
      umma_f16_lo<kDescHiSw128, true>(d, a_lo + 6, b_lo + 6, idesc);
      if (per_commit >= 1) umma_commit(&bar);
      if (per_commit == 2) { mbar_wait(&bar, phase); phase ^= 1; }
    }
    if (per_commit == 1) {   // drain: the barrier completed `groups` phases; resynchronise the parity bookkeeping
      phase = (uint32_t)(groups & 1);
    }
    t_issue = clock64() - t0;
    umma_commit(&bar);
    mbar_wait(&bar, phase);
    out[2 * wid] = clock64() - t0;
    out[2 * wid + 1] = t_issue;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// the same measurement with a WARP-UNIFORM issue loop: all 32 lanes of the issuing warp run the loop, descriptors are
// warp-uniform values, and only the tcgen05 instructions sit under elect.sync -- the compiler can then keep the descriptors
// in uniform registers instead of moving them there (R2UR) once per instruction from a divergent thread
template <int N>
__global__ void __launch_bounds__(128) umma_rate_uniform_kernel(int n_mma, int per_commit, int n_acc, int elect, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int n_issuers = elect < 1 ? 1 : elect;
  const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  uint64_t& bar = bars[wid];
  if (wid < n_issuers) {
    const uint32_t tb = tmem_base + (uint32_t)(wid * N * n_acc);
    const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
    constexpr uint32_t idesc = make_idesc_f16(N);
    uint32_t phase = 0;
    const long long t0 = clock64();
    long long t_issue = 0;
    const uint32_t a_lo = desc_lo(a_s), b_lo = desc_lo(b_s);
    const int groups = n_mma / 4;
    for (int g = 0; g < groups; ++g) {
      const uint32_t d = tb + (uint32_t)((g & (n_acc - 1)) * N);
      if (elect_one()) {
        umma_f16_lo<kDescHiSw128, true>(d, a_lo, b_lo, idesc);
        umma_f16_lo<kDescHiSw128, true>(d, a_lo + 2, b_lo + 2, idesc);
        umma_f16_lo<kDescHiSw128, true>(d, a_lo + 4, b_lo + 4, idesc);
        umma_f16_lo<kDescHiSw128, true>(d, a_lo + 6, b_lo + 6, idesc);
        if (per_commit >= 1) umma_commit(&bar);
      }
      __syncwarp();
      if (per_commit == 2) { mbar_wait(&bar, phase); phase ^= 1; }
    }
    if (per_commit == 1) phase = (uint32_t)(groups & 1);
    t_issue = clock64() - t0;
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, phase);
    if ((threadIdx.x & 31) == 0) {
      out[2 * wid] = clock64() - t0;
      out[2 * wid + 1] = t_issue;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}
}  // namespace gclb

extern "C" int gclb_debug_umma_rate(int32_t n, int32_t n_mma, int32_t per_commit, int32_t n_acc, int32_t elect, long long* out_dev, void* stream) {
  GCLB_CHECK_ARG(out_dev && (n == 32 || n == 64 || n == 128 || n == 256) && n_acc >= 1 && n_acc * n * (elect < 1 ? 1 : elect) <= 512 && elect <= 4, "bad arguments");
  const size_t smem = 16384 + (size_t)n * 128;
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<1, 128, smem, (cudaStream_t)stream>>>(n_mma, per_commit, n_acc, elect, out_dev);
  };
  if (per_commit >= 16) {   // bit 4: the warp-uniform issue loop (elect.sync), low bits as before
    per_commit -= 16;
    if (n == 32) go(gclb::umma_rate_uniform_kernel<32>);
    else if (n == 64) go(gclb::umma_rate_uniform_kernel<64>);
    else if (n == 128) go(gclb::umma_rate_uniform_kernel<128>);
    else go(gclb::umma_rate_uniform_kernel<256>);
  } else if (n == 32) go(gclb::umma_rate_kernel<32>);
  else if (n == 64) go(gclb::umma_rate_kernel<64>);
  else if (n == 128) go(gclb::umma_rate_kernel<128>);
  else go(gclb::umma_rate_kernel<256>);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}
