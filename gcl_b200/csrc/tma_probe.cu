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
