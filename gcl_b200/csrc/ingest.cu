// Data ingest: KITTI / nuScenes velodyne records (float32 x, y, z, reflectance) already in device memory ->
// the float32 [P, 3] point matrix K1 voxelises, with the loaders' optional augmentation applied on the way:
//   lib/complement_data_loader.py:358-361   xyzr = np.fromfile(fname, dtype=np.float32).reshape(-1, 4); return xyzr[:, :3]
//   :65-70, :753-781                        pts @ R.T + T per cloud (float32), then scale * pts (random rotation / scale)
// One pass, 16 B read + 12 B written per point (HBM-bound); the records of all clouds of a batch sit in ONE buffer that came
// over in one pinned-memory H2D copy (gcl_b200/ingest.py).
#include "common.cuh"

namespace gclb {

__global__ void __launch_bounds__(256) ingest_points_kernel(const float* __restrict__ rec, int64_t P, int width,
                                                            const int64_t* __restrict__ cloud_ptr, int n_clouds,
                                                            const float* __restrict__ trans, const float* __restrict__ scales,
                                                            float* __restrict__ xyz) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (int64_t)gridDim.x * blockDim.x) {
    float x, y, z;
    if (width == 4) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(rec) + i);
      x = r.x; y = r.y; z = r.z;
    } else {
      x = rec[i * 3]; y = rec[i * 3 + 1]; z = rec[i * 3 + 2];
    }
    if (trans || scales) {
      int lo = 0, hi = n_clouds;                 // cloud of point i: cloud_ptr[lo] <= i < cloud_ptr[lo + 1]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(cloud_ptr + mid) <= i) lo = mid; else hi = mid; }
      if (trans) {
        const float* T = trans + (size_t)lo * 16;
        // float32, un-fused multiply / add in numpy's left-to-right order (pts @ R.T + T): row . (x, y, z) then + t
        const float nx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, T[0]), __fmul_rn(y, T[1])), __fmul_rn(z, T[2])), T[3]);
        const float ny = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, T[4]), __fmul_rn(y, T[5])), __fmul_rn(z, T[6])), T[7]);
        const float nz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, T[8]), __fmul_rn(y, T[9])), __fmul_rn(z, T[10])), T[11]);
        x = nx; y = ny; z = nz;
      }
      if (scales) { const float s = __ldg(scales + lo); x = __fmul_rn(s, x); y = __fmul_rn(s, y); z = __fmul_rn(s, z); }
    }
    xyz[i * 3] = x; xyz[i * 3 + 1] = y; xyz[i * 3 + 2] = z;
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" int gclb_ingest_points(const float* records, int64_t P, int32_t width, const int64_t* cloud_ptr, int32_t n_clouds,
                                  const float* transforms, const float* scales, float* xyz_out, void* stream) {
  GCLB_CHECK_ARG((width == 4 || width == 3) && (P == 0 || (records && xyz_out)), "bad arguments");
  GCLB_CHECK_ARG(!(transforms || scales) || (cloud_ptr && n_clouds >= 1), "per-cloud transforms / scales need cloud_ptr");
  GCLB_CHECK_ARG(width != 4 || ((uintptr_t)records & 15) == 0, "float32 x 4 records must be 16-byte aligned");
  if (P == 0) return GCLB_OK;
  int64_t blocks = (P + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  ingest_points_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(records, P, width, cloud_ptr, n_clouds, transforms, scales,
                                                                          xyz_out);
  count_launches(1);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}
