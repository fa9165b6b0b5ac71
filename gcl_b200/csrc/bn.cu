// MinkowskiBatchNorm in TRAINING mode (model/common.py:6 -> torch.nn.BatchNorm1d over the rows of .F; forward at
// lib/colocation_trainer.py:846, backward through loss.backward() :879), as three hand-written kernels instead of ATen's
// batch_norm_collect_statistics / transform / backward_reduce / backward_elemt + a separate ReLU:
//   bn_colsum    per-channel sums of two row-wise quantities (x, x^2 | dy, dy * x_hat) in fp64: each block reduces a slab of
//                rows in registers + shared memory and adds ONE partial per channel to the global accumulators
//   bn_fwd_apply y = (x - mean) * invstd * gamma + beta (-> ReLU when fused); block 0 also writes mean / invstd for the
//                backward pass and updates running_mean / running_var (unbiased variance, momentum) like BatchNorm1d
//   bn_bwd_apply dx = gamma * invstd * (dy' - mean(dy') - x_hat * mean(dy' * x_hat)), dy' = dy masked by the fused ReLU;
//                block 0 writes dgamma, dbeta
// Rows are [N, C] row-major with C a multiple of 4 (32 ... 256 in the ResUNet): threads run along channels (coalesced
// float4), 8.75 M parameters never matter here -- the pass is HBM-bound: 2 reads + 1 write of the activations per direction.
#include "common.cuh"

namespace gclb {

constexpr int kBnRows = 128;     // rows per block of the column-sum kernel

// sums[0][c] += sum_r a(r, c), sums[1][c] += sum_r b(r, c)
//   MODE 0: a = x, b = x^2            MODE 1: a = dy', b = dy' * x_hat  (x_hat = (x - mean) * invstd; dy' = y > 0 ? dy : 0 if relu)
template <int MODE>
__global__ void __launch_bounds__(256) bn_colsum_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ y, int64_t n, int c,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        int relu, double* __restrict__ sums) {
  extern __shared__ double s_part[];        // [2][blockDim.y-groups][c]  -> reduced over row groups
  const int c4 = c >> 2;                    // float4 columns
  const int tx = threadIdx.x % c4 < c4 ? threadIdx.x % c4 : 0;
  const int groups = blockDim.x / c4;       // row groups working in parallel inside the block
  const int g = threadIdx.x / c4;
  const int64_t r0 = (int64_t)blockIdx.x * kBnRows;
  const int64_t r1 = min(n, r0 + kBnRows);
  double a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
  if (g < groups) {
    float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
    if (MODE == 1) { mu = __ldg(reinterpret_cast<const float4*>(mean) + tx); is = __ldg(reinterpret_cast<const float4*>(invstd) + tx); }
    for (int64_t r = r0 + g; r < r1; r += groups) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + r * c) + tx);
      if (MODE == 0) {
        a[0] += xv.x; a[1] += xv.y; a[2] += xv.z; a[3] += xv.w;
        b[0] += (double)xv.x * xv.x; b[1] += (double)xv.y * xv.y; b[2] += (double)xv.z * xv.z; b[3] += (double)xv.w * xv.w;
      } else {
        float4 d = __ldg(reinterpret_cast<const float4*>(dy + r * c) + tx);
        if (relu) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + r * c) + tx);
          d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
        }
        a[0] += d.x; a[1] += d.y; a[2] += d.z; a[3] += d.w;
        b[0] += (double)d.x * ((xv.x - mu.x) * is.x); b[1] += (double)d.y * ((xv.y - mu.y) * is.y);
        b[2] += (double)d.z * ((xv.z - mu.z) * is.z); b[3] += (double)d.w * ((xv.w - mu.w) * is.w);
      }
    }
    double* pa = s_part + (size_t)g * c + 4 * tx;
    double* pb = s_part + (size_t)(groups + g) * c + 4 * tx;
#pragma unroll
    for (int q = 0; q < 4; ++q) { pa[q] = a[q]; pb[q] = b[q]; }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < 2 * c; ch += blockDim.x) {      // fixed-order reduction over the row groups
    const int which = ch / c, cc = ch % c;
    double t = 0.0;
    for (int q = 0; q < groups; ++q) t += s_part[(size_t)(which * groups + q) * c + cc];
    atomicAdd(&sums[(size_t)which * c + cc], t);
  }
}

__global__ void __launch_bounds__(256) bn_fwd_apply_kernel(const float* __restrict__ x, int64_t n, int c,
                                                           const double* __restrict__ sums, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, float momentum,
                                                           float* __restrict__ running_mean, float* __restrict__ running_var,
                                                           int relu, float* __restrict__ y, float* __restrict__ save_mean,
                                                           float* __restrict__ save_invstd) {
  extern __shared__ float s_ab[];           // scale[c] | shift[c]
  const double inv_n = 1.0 / (double)n;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const double m = sums[ch] * inv_n;
    double var = sums[c + ch] * inv_n - m * m;        // biased variance (what normalises the batch)
    var = var > 0.0 ? var : 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = gamma ? __ldg(gamma + ch) : 1.f, be = beta ? __ldg(beta + ch) : 0.f;
    s_ab[ch] = is * ga;
    s_ab[c + ch] = be - (float)m * is * ga;
    if (blockIdx.x == 0) {
      save_mean[ch] = (float)m;
      save_invstd[ch] = is;
      if (running_mean) {
        const double unbiased = n > 1 ? var * ((double)n / (double)(n - 1)) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)m;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
      }
    }
  }
  __syncthreads();
  const int c4 = c >> 2;
  const int64_t total4 = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(e % c4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + e);
    float4 o;
    o.x = fmaf(v.x, s_ab[cc], s_ab[c + cc]); o.y = fmaf(v.y, s_ab[cc + 1], s_ab[c + cc + 1]);
    o.z = fmaf(v.z, s_ab[cc + 2], s_ab[c + cc + 2]); o.w = fmaf(v.w, s_ab[cc + 3], s_ab[c + cc + 3]);
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    reinterpret_cast<float4*>(y)[e] = o;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           const float* __restrict__ y, int64_t n, int c,
                                                           const double* __restrict__ sums, const float* __restrict__ gamma,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           int relu, float* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
  extern __shared__ float s_k[];            // per channel: a = gamma*invstd, mean, invstd, m1 = mean(dy'), m2 = mean(dy' x_hat)
  const double inv_n = 1.0 / (double)n;
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    const float ga = gamma ? __ldg(gamma + ch) : 1.f;
    s_k[ch] = ga * __ldg(invstd + ch);
    s_k[c + ch] = __ldg(mean + ch);
    s_k[2 * c + ch] = __ldg(invstd + ch);
    s_k[3 * c + ch] = (float)(sums[ch] * inv_n);
    s_k[4 * c + ch] = (float)(sums[c + ch] * inv_n);
    if (blockIdx.x == 0) {
      if (dbeta) dbeta[ch] = (float)sums[ch];
      if (dgamma) dgamma[ch] = (float)sums[c + ch];
    }
  }
  __syncthreads();
  const int c4 = c >> 2;
  const int64_t total4 = n * c4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(e % c4) * 4;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + e);
    float4 d = __ldg(reinterpret_cast<const float4*>(dy) + e);
    if (relu) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + e);
      d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
    }
    const float xin[4] = {xv.x, xv.y, xv.z, xv.w}, din[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xh = (xin[q] - s_k[c + cc + q]) * s_k[2 * c + cc + q];
      o[q] = s_k[cc + q] * (din[q] - s_k[3 * c + cc + q] - xh * s_k[4 * c + cc + q]);
    }
    reinterpret_cast<float4*>(dx)[e] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static int bn_block(int c) {       // threads: a multiple of the float4 columns, <= 256
  const int c4 = c >> 2;
  int t = (256 / c4) * c4;
  return t < c4 ? c4 : t;
}

}  // namespace gclb

using namespace gclb;

extern "C" {

int gclb_bn_train_fwd(const float* x, int64_t n, int32_t c, const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, int32_t relu, float* y, float* save_mean, float* save_invstd,
                      double* sums /* [2, c] workspace, zeroed here */, void* stream) {
  GCLB_CHECK_ARG(x && y && save_mean && save_invstd && sums && n >= 1, "bad arguments");
  GCLB_CHECK_ARG(c >= 4 && c % 4 == 0 && c <= 1024, "channel count must be a multiple of 4, <= 1024");
  GCLB_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "running_mean / running_var go together");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
  const int threads = bn_block(c), groups = threads / (c >> 2);
  bn_colsum_kernel<0><<<(unsigned)((n + kBnRows - 1) / kBnRows), threads, sizeof(double) * 2 * groups * c, st>>>(
      x, nullptr, nullptr, n, c, nullptr, nullptr, 0, sums);
  int64_t blocks = (n * (c >> 2) + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  bn_fwd_apply_kernel<<<(unsigned)blocks, 256, sizeof(float) * 2 * c, st>>>(x, n, c, sums, gamma, beta, eps, momentum, running_mean,
                                                                         running_var, relu, y, save_mean, save_invstd);
  count_launches(2);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_bn_train_bwd(const float* x, const float* dy, const float* y, int64_t n, int32_t c, const float* gamma,
                      const float* save_mean, const float* save_invstd, int32_t relu, float* dx, float* dgamma, float* dbeta,
                      double* sums /* [2, c] workspace, zeroed here */, void* stream) {
  GCLB_CHECK_ARG(x && dy && dx && save_mean && save_invstd && sums && n >= 1, "bad arguments");
  GCLB_CHECK_ARG(!relu || y, "the fused ReLU needs the forward output to rebuild its mask");
  GCLB_CHECK_ARG(c >= 4 && c % 4 == 0 && c <= 1024, "channel count must be a multiple of 4, <= 1024");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * c, st);
  const int threads = bn_block(c), groups = threads / (c >> 2);
  bn_colsum_kernel<1><<<(unsigned)((n + kBnRows - 1) / kBnRows), threads, sizeof(double) * 2 * groups * c, st>>>(
      x, dy, y, n, c, save_mean, save_invstd, relu, sums);
  int64_t blocks = (n * (c >> 2) + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
  bn_bwd_apply_kernel<<<(unsigned)blocks, 256, sizeof(float) * 5 * c, st>>>(x, dy, y, n, c, sums, gamma, save_mean, save_invstd, relu,
                                                                         dx, dgamma, dbeta);
  count_launches(2);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

}  // extern "C"
