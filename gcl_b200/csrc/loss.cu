// K5: GCL group-wise contrastive loss, forward + backward (lib/colocation_trainer.py:430-535, :734-809).
//   positive / finest terms : one warp per selected group (mean, variance-to-mean, finest-to-mean)
//   hardest-negative term   : one warp per sampled row: exact fp32 distances to the second sample, arg-min,
//                             self / positive-pair masking by binary search in the sorted pair-hash list
//   finalize                : one block, fixed-order sums -> deterministic losses
//   backward                : analytic gradients scattered into gradF with atomicAdd
// Latency-bound (tens of MFLOP): 4 launches per step instead of ~10 ATen launches x 1024 groups.
#include "common.cuh"

namespace gclb {

constexpr int kMaxC = 128;   // channels per row handled per lane: kMaxC / 32
constexpr int CPL = kMaxC / 32;
constexpr float kEps = 1e-7f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

struct LossParams {
  const float* F; int64_t N; int C;
  const int64_t* group_ptr; const int64_t* index; const int32_t* finest_pos;
  const int64_t* pos_sel; int64_t n_sel;
  const int64_t* sel1; const int64_t* sel2; int64_t n_hn;
  const int64_t* keys; int64_t n_keys;
  float pos_t, fin_t, neg_t; int square;
  float w_pos, w_fin, w_neg;
  const float* w_dev;      // device float[3] (upstream gradients of the three losses) or NULL: use w_pos / w_fin / w_neg
  float* pos_vals; float* fin_vals; float* neg_vals; int32_t* neg_j; float* neg_D;
  float* losses; float* gradF;
};

__global__ void __launch_bounds__(256) group_pos_kernel(LossParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= p.n_sel) return;
  const int64_t g = p.pos_sel[s];
  const int64_t b = p.group_ptr[g], e = p.group_ptr[g + 1];
  const int n = (int)(e - b);
  const float inv_n = 1.f / (float)n;
  float mu[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) mu[q] = 0.f;
  for (int64_t m = b; m < e; ++m) {
    const float* f = p.F + (size_t)p.index[m] * p.C;
#pragma unroll
    for (int q = 0; q < CPL; ++q) { int c = lane + 32 * q; if (c < p.C) mu[q] += __ldg(f + c); }
  }
#pragma unroll
  for (int q = 0; q < CPL; ++q) mu[q] *= inv_n;

  // pass 2: m = mean_i d2_i (square) or mean_i sqrt(d2_i + eps); v = sum_i (mu - f_i) / s_i (non-square bwd)
  float msum = 0.f;
  float v[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) v[q] = 0.f;
  for (int64_t m = b; m < e; ++m) {
    const float* f = p.F + (size_t)p.index[m] * p.C;
    float d[CPL], d2 = 0.f;
#pragma unroll
    for (int q = 0; q < CPL; ++q) { int c = lane + 32 * q; d[q] = (c < p.C) ? mu[q] - __ldg(f + c) : 0.f; d2 = fmaf(d[q], d[q], d2); }
    d2 = warp_sum(d2);
    if (p.square) msum += d2;
    else {
      float si = sqrtf(d2 + kEps);
      msum += si;
#pragma unroll
      for (int q = 0; q < CPL; ++q) v[q] += d[q] / si;
    }
  }
  const float mval = msum * inv_n;
  const float pos = fmaxf(mval - p.pos_t, 0.f);
  // finest term
  float fin = 0.f, e2 = 0.f, dfin[CPL];
  int64_t fin_row = -1;
  if (p.finest_pos) {
    const int fp = p.finest_pos[g];
    fin_row = p.index[b + ((fp >= 0 && fp < n) ? fp : 0)];     // the host rejects groups without a finest member
    const float* f = p.F + (size_t)fin_row * p.C;
#pragma unroll
    for (int q = 0; q < CPL; ++q) { int c = lane + 32 * q; dfin[q] = (c < p.C) ? mu[q] - __ldg(f + c) : 0.f; e2 = fmaf(dfin[q], dfin[q], e2); }
    e2 = warp_sum(e2);
    fin = p.square ? fmaxf(e2 - p.fin_t, 0.f) : fmaxf(sqrtf(e2 + kEps) - p.fin_t, 0.f);
  }
  if (lane == 0) { p.pos_vals[s] = pos; p.fin_vals[s] = fin; }
  if (!p.gradF) return;

  const float w_pos = p.w_dev ? __ldg(p.w_dev) : p.w_pos, w_fin = p.w_dev ? __ldg(p.w_dev + 1) : p.w_fin;
  const float gs_pos = (pos > 0.f) ? w_pos / (float)p.n_sel : 0.f;
  float gs_fin = 0.f;
  if (p.finest_pos && fin > 0.f) gs_fin = w_fin / (float)p.n_sel * (p.square ? 2.f : 1.f / sqrtf(e2 + kEps));
  if (gs_pos == 0.f && gs_fin == 0.f) return;
  for (int64_t m = b; m < e; ++m) {
    const int64_t row = p.index[m];
    const float* f = p.F + (size_t)row * p.C;
    float d2 = 0.f, d[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) { int c = lane + 32 * q; d[q] = (c < p.C) ? mu[q] - __ldg(f + c) : 0.f; d2 = fmaf(d[q], d[q], d2); }
    float si = 1.f;
    if (!p.square) { d2 = warp_sum(d2); si = sqrtf(d2 + kEps); }
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      int c = lane + 32 * q;
      if (c >= p.C) continue;
      float gacc = 0.f;
      if (gs_pos != 0.f) {
        // square: d m / d f_j = (2/n)(f_j - mu);   non-square: (1/n) [ (1/n) v - (mu - f_j)/s_j ]
        gacc += gs_pos * (p.square ? -2.f * inv_n * d[q] : inv_n * (inv_n * v[q] - d[q] / si));
      }
      if (gs_fin != 0.f) gacc += gs_fin * dfin[q] * (inv_n - ((m - b) == p.finest_pos[g] ? 1.f : 0.f));
      if (gacc != 0.f) atomicAdd(p.gradF + (size_t)row * p.C + c, gacc);
    }
  }
}

__device__ __forceinline__ bool key_in_sorted(const int64_t* keys, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int64_t v = __ldg(&keys[mid]);
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(&keys[lo]) == key;
}

__global__ void __launch_bounds__(256) hardneg_kernel(LossParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= p.n_hn) return;
  const int64_t arow = p.sel1[r];
  const float* fa = p.F + (size_t)arow * p.C;
  unsigned long long best = ~0ull;
  for (int64_t j = lane; j < p.n_hn; j += 32) {
    const float* fb = p.F + (size_t)p.sel2[j] * p.C;
    float d2 = 0.f;
    for (int c = 0; c < p.C; ++c) { float d = __ldg(fa + c) - __ldg(fb + c); d2 = fmaf(d, d, d2); }
    unsigned long long cand = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)j;
    best = min(best, cand);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)best, m);
    unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(best >> 32), m);
    best = min(best, ((unsigned long long)hi << 32) | lo);
  }
  if (lane == 0) {
    int j = (int)(best & 0xffffffffu);
    float D = sqrtf(__uint_as_float((unsigned)(best >> 32)) + kEps);
    int64_t brow = p.sel2[j];
    int64_t k1 = arow * p.N + brow, k2 = arow + brow * p.N;
    bool valid = (arow != brow) && !key_in_sorted(p.keys, p.n_keys, k1 < k2 ? k1 : k2);
    float h = fmaxf(p.neg_t - D, 0.f);
    p.neg_vals[r] = valid ? h * h : 0.f;
    p.neg_j[r] = valid ? j : -1;
    p.neg_D[r] = D;
  }
}

__global__ void __launch_bounds__(1024) loss_finalize_kernel(LossParams p) {
  __shared__ float red[3][32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float a = 0.f, b = 0.f, c = 0.f, cnt = 0.f;
  for (int64_t i = tid; i < p.n_sel; i += 1024) { a += p.pos_vals[i]; b += p.fin_vals[i]; }
  for (int64_t i = tid; i < p.n_hn; i += 1024) { c += p.neg_vals[i]; cnt += (p.neg_j[i] >= 0) ? 1.f : 0.f; }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c); cnt = warp_sum(cnt);
  __shared__ float red4[32];
  if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = c; red4[wid] = cnt; }
  __syncthreads();
  if (wid == 0) {
    a = warp_sum(red[0][lane]); b = warp_sum(red[1][lane]); c = warp_sum(red[2][lane]); cnt = warp_sum(red4[lane]);
    if (lane == 0) {
      p.losses[0] = a / (float)p.n_sel;
      p.losses[1] = p.finest_pos ? b / (float)p.n_sel : 0.f;
      p.losses[2] = c / cnt;            // empty selection -> NaN, like torch's mean of an empty tensor
      p.losses[3] = cnt;
    }
  }
}

__global__ void __launch_bounds__(256) hardneg_bwd_kernel(LossParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= p.n_hn) return;
  const int j = p.neg_j[r];
  if (j < 0) return;
  const float D = p.neg_D[r];
  const float h = fmaxf(p.neg_t - D, 0.f);
  if (h == 0.f) return;
  // L = mean relu(t - D)^2 ; dL/dD = -2 h / n_valid ; dD/da = (a - b) / D
  const float coef = (p.w_dev ? __ldg(p.w_dev + 2) : p.w_neg) * (-2.f * h) / p.losses[3] / D;
  const int64_t arow = p.sel1[r], brow = p.sel2[j];
  for (int c = lane; c < p.C; c += 32) {
    float d = __ldg(p.F + (size_t)arow * p.C + c) - __ldg(p.F + (size_t)brow * p.C + c);
    atomicAdd(p.gradF + (size_t)arow * p.C + c, coef * d);
    atomicAdd(p.gradF + (size_t)brow * p.C + c, -coef * d);
  }
}

}  // namespace gclb

using namespace gclb;

extern "C" {

size_t gclb_loss_workspace_bytes(int64_t n_sel, int64_t n_hn) { return (size_t)(2 * n_sel + 3 * n_hn + 16) * 4; }

static int group_loss_impl(const float* F, int64_t N, int32_t C, const int64_t* group_ptr, const int64_t* index,
                    const int32_t* finest_pos, const int64_t* pos_sel, int64_t n_sel, const int64_t* sel_hn1,
                    const int64_t* sel_hn2, int64_t n_hn, const int64_t* pos_keys_sorted, int64_t n_keys,
                    float pos_thresh, float finest_thresh, float neg_thresh, int32_t square_loss,
                    const float* weights, const float* upstream_dev, float* losses_out, float* gradF, void* workspace,
                    void* stream) {
  GCLB_CHECK_ARG(F && group_ptr && index && pos_sel && sel_hn1 && sel_hn2 && losses_out && workspace, "null pointer");
  GCLB_CHECK_ARG(C >= 1 && C <= kMaxC, "C must be in 1..128");
  GCLB_CHECK_ARG(n_sel >= 1 && n_hn >= 1, "empty selection");
  GCLB_CHECK_ARG(n_keys == 0 || pos_keys_sorted, "null pointer");
  GCLB_CHECK_ARG(!gradF || weights || upstream_dev, "weights (host) or upstream (device) are required with gradF");
  cudaStream_t st = (cudaStream_t)stream;
  LossParams p;
  p.F = F; p.N = N; p.C = C; p.group_ptr = group_ptr; p.index = index; p.finest_pos = finest_pos;
  p.pos_sel = pos_sel; p.n_sel = n_sel; p.sel1 = sel_hn1; p.sel2 = sel_hn2; p.n_hn = n_hn;
  p.keys = pos_keys_sorted; p.n_keys = n_keys;
  p.pos_t = pos_thresh; p.fin_t = finest_thresh; p.neg_t = neg_thresh; p.square = square_loss;
  p.w_pos = p.w_fin = p.w_neg = 0.f;
  p.w_dev = gradF ? upstream_dev : nullptr;
  if (gradF && !upstream_dev) { p.w_pos = weights[0]; p.w_fin = weights[1]; p.w_neg = weights[2]; }   // host scalars
  float* ws = (float*)workspace;
  p.pos_vals = ws; p.fin_vals = ws + n_sel; p.neg_vals = ws + 2 * n_sel;
  p.neg_j = (int32_t*)(ws + 2 * n_sel + n_hn); p.neg_D = ws + 2 * n_sel + 2 * n_hn;
  p.losses = losses_out; p.gradF = gradF;
  group_pos_kernel<<<(unsigned)((n_sel + 7) / 8), 256, 0, st>>>(p);
  hardneg_kernel<<<(unsigned)((n_hn + 7) / 8), 256, 0, st>>>(p);
  loss_finalize_kernel<<<1, 1024, 0, st>>>(p);
  if (gradF) hardneg_bwd_kernel<<<(unsigned)((n_hn + 7) / 8), 256, 0, st>>>(p);
  count_launches(gradF ? 4 : 3);
  GCLB_CHECK_LAUNCH();
  return GCLB_OK;
}

int gclb_group_loss(const float* F, int64_t N, int32_t C, const int64_t* group_ptr, const int64_t* index,
                    const int32_t* finest_pos, const int64_t* pos_sel, int64_t n_sel, const int64_t* sel_hn1,
                    const int64_t* sel_hn2, int64_t n_hn, const int64_t* pos_keys_sorted, int64_t n_keys,
                    float pos_thresh, float finest_thresh, float neg_thresh, int32_t square_loss,
                    const float* weights, float* losses_out, float* gradF, void* workspace, void* stream) {
  return group_loss_impl(F, N, C, group_ptr, index, finest_pos, pos_sel, n_sel, sel_hn1, sel_hn2, n_hn, pos_keys_sorted,
                         n_keys, pos_thresh, finest_thresh, neg_thresh, square_loss, weights, nullptr, losses_out, gradF,
                         workspace, stream);
}

int gclb_group_loss_bwd(const float* F, int64_t N, int32_t C, const int64_t* group_ptr, const int64_t* index,
                        const int32_t* finest_pos, const int64_t* pos_sel, int64_t n_sel, const int64_t* sel_hn1,
                        const int64_t* sel_hn2, int64_t n_hn, const int64_t* pos_keys_sorted, int64_t n_keys,
                        float pos_thresh, float finest_thresh, float neg_thresh, int32_t square_loss,
                        const float* upstream_dev, float* losses_scratch, float* gradF, void* workspace, void* stream) {
  GCLB_CHECK_ARG(upstream_dev && gradF && losses_scratch, "null pointer");
  return group_loss_impl(F, N, C, group_ptr, index, finest_pos, pos_sel, n_sel, sel_hn1, sel_hn2, n_hn, pos_keys_sorted,
                         n_keys, pos_thresh, finest_thresh, neg_thresh, square_loss, nullptr, upstream_dev, losses_scratch,
                         gradF, workspace, stream);
}

}  // extern "C"
