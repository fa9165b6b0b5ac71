/* gclb200.h -- C ABI of libgclb200.so: the B200-native (sm_100a) replacement for the MinkowskiEngine /
 * ATen work on the FCGF/GCL feature-extraction + matching hot path of liuQuan98/GCL.
 *
 * The reference has no FFI of its own on this path: it calls the un-vendored MinkowskiEngine Python package
 * (requirements.txt:8) and ATen.  Each entry point below names the reference call site(s) whose native work
 * it replaces (file:line under /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error; gclb_last_error() gives a thread-local message.
 *  - all buffers are CALLER-ALLOCATED DEVICE pointers (e.g. PyTorch's caching allocator); no hidden
 *    allocation, no host synchronisation, no host callbacks: every call only enqueues work on `stream`.
 *    Results whose size is data dependent are written into caller buffers sized for the worst case and the
 *    count is written to a device int64 the caller reads when it needs it.
 *  - `stream` is a cudaStream_t passed as void*.
 *  - coordinates are int32 rows (batch, x, y, z).  Valid range: 0 <= batch < 1023, |x|,|y|,|z| < 2^17;
 *    rows outside it set bit GCLB_ST_RANGE in the caller's `status` word (device int32, caller-zeroed).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with GCLB_ERR_CUDA.
 */
#ifndef GCLB200_H_
#define GCLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCLB_OK 0
#define GCLB_ERR_ARG (-1)
#define GCLB_ERR_CUDA (-2)
#define GCLB_ERR_UNSUPPORTED (-3)

/* bits of the device status word */
#define GCLB_ST_RANGE 1      /* a coordinate was outside the packable range */
#define GCLB_ST_FULL 2       /* hash table full (capacity too small) */
#define GCLB_ST_DUPLICATE 4  /* gclb_hash_build saw a duplicated coordinate row */
#define GCLB_ST_FP16_OVERFLOW 8    /* an fp16-stored activation left the finite fp16 range (or was NaN) and was saturated */
#define GCLB_ST_FP16_UNDERFLOW 16  /* a whole fp16-stored tensor has max |y| < 2^-11: its values are (near-)subnormal in fp16 */

const char* gclb_last_error(void);
int gclb_version(void);
/* process-wide count of CUDA kernels this library has enqueued so far (bench.py reports the delta) */
int64_t gclb_kernel_launches(void);
/* 1 when the CURRENT CUDA device can run the tcgen05 (UTCxMMA) / TMEM / TMA kernels (compute capability 10.x), else 0 */
int gclb_has_tcgen05(void);

/* ------------------------------------------------------------------------------------------------------
 * Coordinate hash table: open addressing, linear probing, 64-bit packed keys, 32-bit values (row index).
 * Layout in the caller's buffer: `capacity` 32-byte QUAD slots { uint64 group key; uint32 value[4]; 8 bytes unused }
 * (32-byte aligned = one memory sector).  A quad holds the four cells (b, x, y, z) with the same (b, floor(x/4s), y, z),
 * s = the map's tensor stride (a power of two; rows sit on multiples of s): the x-neighbours a kernel-map probe wants
 * arrive with one sector instead of one sector each.  Every entry point that takes a table also takes the tensor
 * stride it was built with.
 * Replaces ME's CoordinateMapGPU insert/find (SparseTensor construction: scripts/test_kitti.py:143-148,
 * lib/colocation_trainer.py:843-845, util/misc.py:128).
 * ---------------------------------------------------------------------------------------------------- */
int64_t gclb_hash_capacity(int64_t n_rows);           /* quad slots: power of two >= 2*n_rows, >= 1024 */
size_t gclb_hash_bytes(int64_t capacity);
/* insert N unique rows; vals = row index.  Duplicates keep the smallest row index and set GCLB_ST_DUPLICATE. */
int gclb_hash_build(void* table, int64_t capacity, int32_t tensor_stride, const int32_t* coords4, int64_t n,
                    int32_t* status, void* stream);
/* rows_out[q] = row index of q4[q] or -1 */
int gclb_hash_query(const void* table, int64_t capacity, int32_t tensor_stride, const int32_t* q4, int64_t nq,
                    int32_t* rows_out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K1 voxelisation = ME.utils.sparse_quantize(xyz / voxel, return_index=True) + floor().int() + sparse_collate
 * (lib/complement_data_loader.py:788-789,809-812,1310-1311; lib/colocation_data_loader.py:379,388,400-401,446).
 *   xyz        float32 [P,3]   points of all clouds, concatenated
 *   cloud_ptr  int64 [n_clouds+1]  (device) start row of each cloud in xyz; batch index = cloud number
 *   voxel      the divisor; the kernel computes floorf(__fdiv_rn(x, voxel)) (== torch fp32 `xyz / voxel`, floor)
 *   table      hash buffer of gclb_hash_capacity(P) slots; on return maps coordinate -> compacted row
 *   coords4_out int32 [P,4], unique_map_out int64 [P]: first V rows valid (ascending first-occurrence order)
 *   inverse_map_out int32 [P] or NULL: compacted row of every input point
 *   n_out      device int64[1] = V
 *   workspace  gclb_compact_workspace_bytes(P) bytes
 * ---------------------------------------------------------------------------------------------------- */
size_t gclb_compact_workspace_bytes(int64_t n);
int gclb_voxelize(const float* xyz, int64_t P, const int64_t* cloud_ptr, int32_t n_clouds, float voxel,
                  void* table, int64_t capacity, int32_t* coords4_out, int64_t* unique_map_out,
                  int32_t* inverse_map_out, int64_t* n_out, int32_t* status, void* workspace, void* stream);
/* same de-duplication for already discretised int32 rows of any width in {3,4} (sparse_quantize on ints,
 * util/misc.py:117-118).  width==3 rows get batch 0. */
int gclb_quantize_rows(const int32_t* rows, int64_t P, int32_t width, void* table, int64_t capacity,
                       int32_t* coords4_out, int64_t* unique_map_out, int32_t* inverse_map_out, int64_t* n_out,
                       int32_t* status, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Strided coordinate map (ME CoordinateManager::stride; model/resunet.py:62-95 stride-2 convolutions):
 * out = unique(floor(c / new_stride) * new_stride), rows ordered by first appearance in the parent map.
 *   n_in_dev    optional device int64: the true parent row count when only an upper bound n_in is known on the host
 *               (lets several levels be chained without a host synchronisation)
 *   out_table   hash buffer of gclb_hash_capacity(n_in) slots -> coordinate -> new row
 *   out_coords4 int32 [n_in,4] (first *n_out rows valid);  parent_row_out int32 [n_in] or NULL: new row of
 *   each parent row.
 * ---------------------------------------------------------------------------------------------------- */
int gclb_stride_map(const int32_t* in_coords4, int64_t n_in, const int64_t* n_in_dev, int32_t new_stride,
                    void* out_table, int64_t out_capacity, int32_t* out_coords4, int32_t* parent_row_out,
                    int64_t* n_out, int32_t* status, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K2 kernel map in output-stationary form ("neighbour table"):
 *   nbr[o*K + k] = input row i with coord_in[i] == coord_out[o] + sign * off_k, else -1,
 *   off_k = ((ix,iy,iz) - ksize/2) * dilation * offset_stride, k = ix + ksize*iy + ksize^2*iz   (odd ksize;
 *   even ksize: (ix,iy,iz) * dilation * offset_stride).
 * Forward conv (model/resunet.py:38-95, residual_block.py:23-33): in_table = input map, out_coords = output
 * map, offset_stride = input tensor stride, sign=+1.  Transposed conv (model/resunet.py:101-134): in_table =
 * coarse map, out_coords = fine map, offset_stride = fine tensor stride, sign=-1.
 *   in_tensor_stride: tensor stride the probed (input) table was built with (power of two, >= 1; REQUIRED: it fixes
 *            the quad layout): candidates not aligned to it are rejected without a table access (most offsets of a
 *            transposed map); 3x3x3 maps whose offsets step by exactly this stride fetch a whole column's x-neighbours
 *            with 1-2 quad probes (18 sectors per row instead of 27).
 *   pair_count int32 [K] or NULL: number of valid entries per offset (caller-zeroed).
 *   row_keys uint8 [n_out] or NULL: 6-bit neighbour-direction key of every row, computed for free during the build
 *            and accepted by gclb_kmap_sort_rows.
 *   key_hist int32 [64, ceil(n_out/1024)] or NULL (caller-zeroed, needs row_keys): histogram of the row keys per
 *            1024-row block, accumulated during the build; gclb_kmap_sort_rows then needs no counting pass.
 *   row_masks uint32 [n_out] or NULL (ksize^3 <= 32): bit k set iff nbr[o, k] >= 0; lets gclb_kmap_sort_rows derive
 *            the per-tile masks without re-reading the table.
 * ---------------------------------------------------------------------------------------------------- */
int gclb_kmap_build(const void* in_table, int64_t in_capacity, const int32_t* out_coords4, int64_t n_out,
                    int32_t ksize, int32_t offset_stride, int32_t dilation, int32_t sign, int32_t in_tensor_stride,
                    int32_t* nbr, int32_t* pair_count, uint8_t* row_keys, uint32_t* row_masks, int32_t* key_hist,
                    void* stream);
/* expand a neighbour table into ME-style per-offset pair lists, canonical order (k ascending, out row
 * ascending): in_idx/out_idx int32 [n_out*K] (first offset_ptr[K] valid), offset_ptr int64 [K+1]. */
int gclb_kmap_pairs(const int32_t* nbr, int64_t n_out, int32_t K, int32_t* in_idx, int32_t* out_idx,
                    int64_t* offset_ptr, void* workspace, void* stream);

/* group the rows of a neighbour table by a 6-bit neighbour-direction key (stable counting sort): perm_out int32 [n_out]
 * (sorted position -> original row), nbr_sorted_out int32 [n_out, ksize^3] = nbr[perm_out]; tile_mask_out uint32
 * [ceil(n_out/128)] or NULL (ksize^3 <= 32): bit k set iff offset k is populated in that 128-row tile.  Pure re-ordering: pass
 * both to gclb_spconv_fwd(algo=2); results are identical, the kernel just runs ~2-8x fewer pipeline stages. */
size_t gclb_kmap_sort_workspace_bytes(int64_t n_out);
int gclb_kmap_sort_rows(const int32_t* nbr, int64_t n_out, int32_t ksize, const uint8_t* row_keys /* or NULL */,
                        const uint32_t* row_masks /* or NULL */, const int32_t* key_hist /* or NULL */, int32_t* perm_out,
                        int32_t* nbr_sorted_out /* or NULL */, uint32_t* tile_mask_out, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K3 sparse convolution forward, output-stationary implicit GEMM with fused epilogue
 *   out[o, :] = act( (sum_k sum_c in[nbr[o,k], c] * W[k, c, :]) * scale + shift + residual[o, :] )
 * Replaces MinkowskiConvolution / MinkowskiConvolutionTranspose forward (+ MinkowskiBatchNorm in eval mode,
 * `out += residual`, MEF.relu, ME.cat when fused): model/resunet.py:173-224, model/residual_block.py:37-53.
 *   in0 [n_in, c0], in1 [n_in, c1] or NULL : the input is the channel concatenation [in0 | in1] (ME.cat fused)
 *   W  float32 [K, c0+c1, cout]            : ME layout (Appendix A10)
 *   nbr int32 [n_out, K] or NULL (K==1: identity map, the kernel_size==1 `F.mm` path)
 *   row_perm int32 [n_out] or NULL: tcgen05 path only -- tile row t processes OUTPUT row row_perm[t] (ordering from
 *            gclb_kmap_sort_rows so that a 128-row tile touches few kernel offsets).  relu bit 2 (value 4) says `nbr`
 *            is the physically re-ordered copy (row t of nbr describes output row row_perm[t]); without it `nbr` is the
 *            original table and the kernel reads row row_perm[t] of it.
 *   tile_mask uint32 [ceil(n_out/128)] or NULL: tcgen05 path only -- populated-offset bit mask per 128-row tile of
 *            `nbr` (from gclb_kmap_sort_rows); saves the kernel an in-tile scan
 *   scale, shift float32 [cout] or NULL    : folded eval-mode BatchNorm / bias
 *   residual float32 [n_out, cout] or NULL ; relu (flags): bit 0 = ReLU, bit 1 = divide every output row by its L2 norm
 *   afterwards (model/resunet.py:226-230; tcgen05 path with cout == 32 only)
 *   algo: 0/1 = fp32 CUDA-core kernel (exact fp32; any channel counts);
 *         2   = tcgen05 kind::tf32 tensor-core kernel, fp32 accumulate in TMEM.  W must then be in the tensor-core
 *               image produced by gclb_weights_to_tc; needs c0 % 32 == 0, c1 % 32 == 0,
 *               cout in {32, 64, 128, 256}, K <= 27 (otherwise GCLB_ERR_UNSUPPORTED).
 *   fp16 activations (tcgen05 path, flags): bit 3 (8) = in0, in1 and residual are IEEE fp16 [n, c] and W is the fp16
 *               image from gclb_weights_to_tc_f16; the MMA is kind::f16 (fp32 accumulate; fp16 has the 10-bit mantissa
 *               of tf32 but is rounded to nearest when stored, so it is the MORE accurate of the two -- see
 *               tools/precision_study.py) and a gathered 128-byte row carries 64 channels instead of 32: needs
 *               c0 % 64 == 0, c1 % 64 == 0.  bit 5 (32, with bit 3) = the fp16 rows are gathered 32 channels (64 bytes,
 *               SWIZZLE_64B) at a time, for sources whose width is a multiple of 32 only (W image with slab_channels 32).
 *               bit 4 (16) = `out` is fp16 (saturating round-to-nearest), independent of bit 3, so a layer can convert
 *               in either direction.  scale/shift are always fp32.
 * ---------------------------------------------------------------------------------------------------- */
/* W [K, cin, cout] (ME layout) -> tensor-core image of the same size: per (k, 32-channel slab) one contiguous
 * cout x 128 B block laid out exactly like the SWIZZLE_128B K-major shared-memory tile, values rounded to nearest tf32
 * (done once per layer; needs cin % 32 == 0, cout % 8 == 0) */
int gclb_weights_to_tc(const float* W, int32_t K, int32_t cin, int32_t cout, float* Wt, void* stream);
/* the tf32 image of the DATA-GRADIENT convolution's weights, straight from the forward weights W [K, cin, cout]: the image of
 * W'[k][c'][n'] = W[flip ? K-1-k : k][n'][c'] (cin' = cout, cout' = cin).  flip = 1 for stride-1 odd kernels, whose backward
 * table is the forward table with the offsets reversed (lib/colocation_trainer.py:879 `loss.backward()` -> MinkowskiEngine's
 * ConvolutionBackward; here one launch instead of flip + transpose + copy + image).  Needs cout % 32 == 0, cin % 8 == 0. */
int gclb_weights_to_tc_dgrad(const float* W, int32_t K, int32_t cin, int32_t cout, int32_t flip, float* Wt, void* stream);
/* same for the fp16 path: per (k, slab of slab_channels channels) one contiguous block of cout rows of fp16 (round to
 * nearest) -- slab_channels = 64: 128-byte rows, SWIZZLE_128B (flag bit 3 alone); slab_channels = 32: 64-byte rows,
 * SWIZZLE_64B (flag bits 3 + 5).  Wt holds K*cin*cout halves; needs cin % slab_channels == 0, cout % 8 == 0 */
int gclb_weights_to_tc_f16(const float* W, int32_t K, int32_t cin, int32_t cout, int32_t slab_channels, void* Wt,
                           void* stream);
int gclb_spconv_fwd(const void* in0, int32_t c0, const void* in1, int32_t c1, int64_t n_in, const void* W,
                    int32_t K, int32_t cout, const int32_t* nbr, const int32_t* row_perm, const uint32_t* tile_mask,
                    const float* scale, const float* shift, const void* residual, int32_t relu_flags, void* out,
                    int64_t n_out, int32_t algo, void* stream);
/* fp16-range monitor for the kernels that STORE fp16 (flag bit 4): the reference computes in fp32 (no range limit), so a
 * checkpoint whose activations overflow fp16 (|y| > 65504, saturated by the epilogue) or sink into its subnormals must be
 * reported, not silently degraded.  `mon` = device uint32[2], caller-zeroed: [0] bit 0 = a value outside the finite fp16 range
 * (or NaN) was produced, [1] = max |y| of the launch as float bits.  The pointer is THREAD-LOCAL state picked up by the
 * following gclb_spconv_fwd / gclb_spconv_fwd_probe calls of this thread (NULL switches monitoring off).
 * gclb_range_check folds n_layers monitors (uint32 [n_layers, 2]) into the caller's status word: GCLB_ST_FP16_OVERFLOW,
 * GCLB_ST_FP16_UNDERFLOW (0 < max |y| < 2^-11).  The remedy is the fp32-storage engine (kind::tf32). */
int gclb_spconv_set_range_monitor(uint32_t* mon);
int gclb_range_check(const uint32_t* mons, int32_t n_layers, int32_t* status, void* stream);

/* stride-1 convolution with a small input width (cin <= 4, e.g. conv1 of ResUNet: cin = 1, kernel 5^3) with the kernel
 * map FUSED into the convolution: the kernel probes the coordinate hash of the (single) coordinate map itself, so no
 * [n, K] neighbour table is built, written or read (model/resunet.py:38-45,174).  Same epilogue as gclb_spconv_fwd
 * (relu_flags: bit 0 = ReLU, bit 4 = `out` is fp16).
 * nbr3_out int32 [n, 27] or NULL (odd ksize >= 3): the probes of the inner 3x3x3 offsets ARE the stride-1 3x3x3 kernel
 * map of the same coordinate map (what the next layers, block1 of the ResUNet, convolve over), so the kernel can write
 * that table -- with row_keys / row_masks / key_hist exactly as gclb_kmap_build produces them (all optional, key_hist
 * caller-zeroed) -- and the separate gclb_kmap_build pass for it disappears. */
int gclb_spconv_fwd_probe(const float* in, int32_t cin, const float* W, int32_t ksize, int32_t cout, const void* table,
                          int64_t capacity, const int32_t* coords4, int64_t n, int32_t tensor_stride, int32_t dilation,
                          const float* scale, const float* shift, const float* residual, int32_t relu_flags, void* out,
                          int32_t* nbr3_out, uint8_t* row_keys, uint32_t* row_masks, int32_t* key_hist, void* stream);
/* ------------------------------------------------------------------------------------------------------
 * K3 with operand reuse inside the SM ("halo staging") for SAME-MAP 3x3x3 convolutions on fp16 activations
 * (model/residual_block.py:23-33,40-53: the two convolutions of every residual block; 16 of the 22 conv launches of
 * ResUNetBN2C.forward, model/resunet.py:173-232).  Neighbouring output rows share most input rows; the direct-gather kernel
 * above fetches every input row ~7x from L2.  Here a precomputed list names the DISTINCT input rows of each 128-row tile
 * (split into groups of <= 384 rows), they are staged in shared memory once per channel slab by TMA gather, and the per-offset
 * tensor-core operand is assembled from the staged rows.  Same arithmetic, same results (bit-identical to algo 2 with the
 * same row order), 3-4x fewer bytes through the L2->SM fabric.
 *
 * gclb_kmap_halo_build: nbr int32 [n_out, 27] (ORIGINAL order), row_perm int32 [n_out] from gclb_kmap_sort_rows (or NULL);
 *   records: caller buffer of record_bytes >= gclb_kmap_halo_bytes(n_out) bytes (the worst case, so the build cannot run out
 *   of space; typical use is ~1/6 of it), 16-byte aligned: packed variable-size group records
 *   { header 64 B | int32 distinct rows | uint16 local index [offsets of the group][128] };
 *   tile_groups int32 [ceil(n_out/128), gclb_kmap_halo_max_groups(), 2] = (offset, size) of every record in 16-byte granules;
 *   tile_ngroups int32 [ceil(n_out/128)]; counter uint64[1] caller-zeroed (granules used); status: GCLB_ST_FULL if the buffer
 *   was smaller than the worst case and did not suffice (the map must then not be used; the kernel traps on such a record).
 * gclb_spconv_fwd_halo: arguments as gclb_spconv_fwd (K = 27; flags bit 3 required; bits 0, 1, 4, 5 as there); W is the
 *   fp16 image from gclb_weights_to_tc_f16 with the matching slab width. */
size_t gclb_kmap_halo_bytes(int64_t n_out);
int32_t gclb_kmap_halo_max_groups(void);
int gclb_kmap_halo_build(const int32_t* nbr, int64_t n_out, const int32_t* row_perm, void* records, size_t record_bytes,
                         int32_t* tile_groups, int32_t* tile_ngroups, uint64_t* counter, int32_t* status, void* stream);
int gclb_spconv_fwd_halo(const void* in0, int32_t c0, const void* in1, int32_t c1, int64_t n_in, const void* W, int32_t cout,
                         const void* records, const int32_t* tile_groups, const int32_t* tile_ngroups, const int32_t* row_perm,
                         const float* scale, const float* shift, const void* residual, int32_t flags, void* out,
                         int64_t n_out, void* stream);

/* wgrad: gW[k, c, :] = sum over pairs in[nbr[o,k], c] * gout[o, :]   (a18; lib/colocation_trainer.py:879) */
int gclb_spconv_wgrad(const float* in, int32_t cin, int64_t n_in, const float* gout, int32_t cout, int64_t n_out,
                      const int32_t* nbr, int32_t K, float* gW, void* stream);

/* Same contraction on tcgen05 (kind::tf32 operands, fp32 accumulation in TMEM) over the row-bucketed table that
 * gclb_kmap_sort_rows produced for the forward pass: nbr_sorted [n_out, K] (NULL = identity, K == 1), row_perm [n_out]
 * (tile row -> output row, NULL = identity), tile_mask [ceil(n_out/128)] (NULL = walk every tile for every offset).
 * Needs cin, cout in {32, 64, ..., 256}.  gW [K, cin, cout] is ACCUMULATED into (caller zeroes it); partial sums of
 * different CTAs are merged with fp32 atomics, so the last bits depend on scheduling.
 * Replaces: ME ConvolutionBackward (weight gradient) reached through loss.backward(), lib/colocation_trainer.py:879. */
int gclb_spconv_wgrad_tc(const float* in, int32_t cin, int64_t n_in, const float* gout, int32_t cout, int64_t n_out,
                         const int32_t* nbr_sorted, const int32_t* row_perm, const int32_t* tile_mask, int32_t K,
                         float* gW, void* stream);
/* pointwise tail of ResUNet: out = l2normalize( relu([in0|in1] W1) W2 + bias )   (model/resunet.py:217-230) */
int gclb_pointwise_tail(const float* in0, int32_t c0, const float* in1, int32_t c1, int64_t n, const float* W1,
                        int32_t cmid, const float* W2, const float* bias, int32_t cout, int32_t normalize,
                        float* out, void* stream);

/* elementwise helpers for the op-by-op (training) path: y = act(x * scale[c] + shift[c] + residual) */
int gclb_affine_act(const float* x, int64_t n, int32_t c, const float* scale, const float* shift,
                    const float* residual, int32_t relu, float* y, void* stream);
/* MinkowskiBatchNorm in TRAINING mode (model/common.py:6 = torch.nn.BatchNorm1d over the rows of .F; forward
 * lib/colocation_trainer.py:846, backward via loss.backward() :879), optionally fused with the MEF.relu that follows it in
 * every residual block (model/residual_block.py:42; model/resunet.py:177-223).
 *   forward:  y = [relu]((x - mean_batch) / sqrt(var_batch + eps) * gamma + beta), statistics over ALL n rows (biased variance);
 *             running_mean / running_var (may be NULL) updated in place with `momentum` and the UNBIASED variance;
 *             save_mean / save_invstd float32 [c] for the backward pass
 *   backward: dx, dgamma, dbeta of the same op given dy (relu != 0: the ReLU mask is rebuilt from the forward output y)
 *   x, y, dy, dx float32 [n, c], c a multiple of 4; gamma / beta may be NULL (affine=False); sums float64 [2, c] workspace */
int gclb_bn_train_fwd(const float* x, int64_t n, int32_t c, const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, int32_t relu, float* y, float* save_mean, float* save_invstd,
                      double* sums, void* stream);
int gclb_bn_train_bwd(const float* x, const float* dy, const float* y, int64_t n, int32_t c, const float* gamma,
                      const float* save_mean, const float* save_invstd, int32_t relu, float* dx, float* dgamma, float* dbeta,
                      double* sums, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K4 nearest neighbour in feature space, both directions from one pass, N x M never materialised.
 * Replaces lib/eval.py:18-48 (find_nn_gpu) + lib/metrics.py:22-29 (pdist) and the two KD-tree queries of
 * generalization_ETH/evaluate.py:63-77 (calculate_M).  Squared L2 in the direct-difference form
 * sum_c (a_c - b_c)^2 in fp32, ties -> smallest index.
 *   A float32 [sum N_p, C], B float32 [sum M_p, C]; a_ptr/b_ptr int64 [n_pairs+1] device segment starts
 *   (n_pairs independent problems in one launch).  Indices are LOCAL to the segment.
 *   a_rows / b_rows int64 or NULL: optional row indirection (fused gather): segment row i reads A[a_rows[i]];
 *   max_n / max_m: upper bounds of the segment lengths (grid sizing; the true lengths are read from a_ptr/b_ptr)
 *   idx01 int64 [sum N], d01 float32 [sum N] ; idx10 int64 [sum M], d10 float32 [sum M] (idx10/d10 may be NULL)
 *   algo: 0 = auto (tcgen05 when C == 32), 1 = exact-fp32 CUDA-core tiles (direct-difference form, any C),
 *         2 = tcgen05 distance GEMM: operands split exactly into 3 tf32 parts, 6 product groups accumulated in TMEM
 *             (|a|^2 + |b|^2 - 2 a.b, fp32-exact products); C == 32 only
 *   workspace: gclb_nn_workspace_bytes(sum N, sum M, n_pairs, max_n, max_m)
 * ---------------------------------------------------------------------------------------------------- */
size_t gclb_nn_workspace_bytes(int64_t n_total, int64_t m_total, int32_t n_pairs, int64_t max_n, int64_t max_m);
int gclb_nn(const float* A, const float* B, int32_t C, const int64_t* a_ptr, const int64_t* b_ptr, int32_t n_pairs,
            const int64_t* a_rows, const int64_t* b_rows, int64_t n_total, int64_t m_total, int64_t max_n,
            int64_t max_m, int64_t* idx01, float* d01, int64_t* idx10, float* d10, int32_t algo, void* workspace,
            void* stream);
/* per-cloud random subsample without replacement of the rows of a batched coordinate map (find_corr's
 * `np.random.choice(len(F), 5000, replace=False)`, scripts/test_kitti.py:34-35), entirely on the device:
 *   coords4 int32 [n_rows,4] rows grouped by cloud (column 0 ascending); n_rows_dev optional device int64 row count
 *   cloud_ptr_out int64 [n_clouds+1] first row of every cloud;
 *   groups G (1 or 2 ...): cloud c belongs to group c % G as its segment c / G (G = 2: scan 0 / scan 1 of pair c/2);
 *   sel_ptr_out int64 [G, n_clouds/G + 1]: per group the CSR prefix of min(V_c, S) over its segments;
 *   sel_out int64 [G, (n_clouds/G) * cap], cap = (S > 0 && S < n_rows) ? S : n_rows: selected GLOBAL row indices, a
 *   pseudo-random permutation prefix per cloud (S <= 0 or V_c <= S keeps all rows in order).
 * Feeds gclb_nn directly: a_rows = sel_out[0], a_ptr = sel_ptr_out[0], b_rows = sel_out[1], b_ptr = sel_ptr_out[1]. */
int gclb_subsample(const int32_t* coords4, const int64_t* n_rows_dev, int64_t n_rows, int32_t n_clouds, int64_t S,
                   int32_t groups, uint64_t seed, int64_t* cloud_ptr_out, int64_t* sel_ptr_out, int64_t* sel_out,
                   void* stream);
/* mutual filter (calculate_M): pairs_out int64 [sum N, 2] = (i, idx01[i]) for rows with idx10[idx01[i]] == i,
 * ascending i inside each segment, segments concatenated; pair_ptr int64 [n_pairs+1] segment starts. */
int gclb_mutual_filter(const int64_t* idx01, const int64_t* idx10, const int64_t* a_ptr, const int64_t* b_ptr,
                       int32_t n_pairs, int64_t n_total, int64_t* pairs_out, int64_t* pair_ptr, void* workspace,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------
 * K5 GCL group-wise contrastive loss, forward + backward in one call
 * (lib/colocation_trainer.py:430-535 finest_contrastive_loss, :734-809 location_contrastive_loss).
 *   F float32 [N, C] (C <= 128);  group_ptr int64 [G+1] CSR over `index`; index int64 [sum group];
 *   finest_pos int32 [G] position of the finest member inside each group (or NULL: no finest term);
 *   pos_sel int64 [n_sel] selected groups; sel_hn1, sel_hn2 int64 [n_hn];
 *   pos_keys_sorted int64 [n_keys] ascending symmetric pair hashes (util/misc.py:29-40), hash seed = N;
 *   square_loss 0/1;  losses_out float32 [4] = pos, finest, neg, n_valid_neg;
 *   weights HOST float32[3] = (pos_w, finest_w, neg_w): gradF += d(sum_i w_i * loss_i)/dF  (gradF device float32
 *   [N, C], caller-zeroed; NULL = forward only)
 *   workspace: gclb_loss_workspace_bytes(n_sel, n_hn)
 * ---------------------------------------------------------------------------------------------------- */
size_t gclb_loss_workspace_bytes(int64_t n_sel, int64_t n_hn);
int gclb_group_loss(const float* F, int64_t N, int32_t C, const int64_t* group_ptr, const int64_t* index,
                    const int32_t* finest_pos, const int64_t* pos_sel, int64_t n_sel, const int64_t* sel_hn1,
                    const int64_t* sel_hn2, int64_t n_hn, const int64_t* pos_keys_sorted, int64_t n_keys,
                    float pos_thresh, float finest_thresh, float neg_thresh, int32_t square_loss,
                    const float* weights, float* losses_out, float* gradF, void* workspace, void* stream);
/* backward of the three losses for ARBITRARY upstream gradients (what autograd hands to the backward of
 * `(pos_w * pos / iter_size + ...).backward()`, lib/colocation_trainer.py:874-879): same inputs as gclb_group_loss,
 *   upstream DEVICE float32[3] = dL/d(pos), dL/d(finest), dL/d(neg)  (no host read: the scalars stay on the device)
 *   gradF device float32 [N, C], caller-zeroed: gradF += sum_i upstream[i] * d loss_i / dF
 *   losses_scratch device float32[4] (recomputed forward values; the kernels are latency-bound, recomputation is cheaper
 *   than keeping per-term gradient planes) */
int gclb_group_loss_bwd(const float* F, int64_t N, int32_t C, const int64_t* group_ptr, const int64_t* index,
                        const int32_t* finest_pos, const int64_t* pos_sel, int64_t n_sel, const int64_t* sel_hn1,
                        const int64_t* sel_hn2, int64_t n_hn, const int64_t* pos_keys_sorted, int64_t n_keys,
                        float pos_thresh, float finest_thresh, float neg_thresh, int32_t square_loss,
                        const float* upstream, float* losses_scratch, float* gradF, void* workspace, void* stream);


/* ------------------------------------------------------------------------------------------------------
 * Data ingest: velodyne records -> the point matrix K1 voxelises.
 * Replaces lib/complement_data_loader.py:358-361 (`np.fromfile(fname, np.float32).reshape(-1, 4)[:, :3]`) and the loaders'
 * augmentation :65-70, :753-781 (`pts @ R.T + T` per cloud in float32, then `scale * pts`).
 *   records float32 [P, width] (width 4: x, y, z, reflectance as in a KITTI .bin, 16-byte aligned; width 3: xyz) of all clouds of a
 *   batch, concatenated (device; gcl_b200/ingest.py reads the files into ONE pinned buffer and copies it once);
 *   cloud_ptr int64 [n_clouds+1] (device), transforms float32 [n_clouds, 4, 4] row-major or NULL, scales float32 [n_clouds] or NULL;
 *   xyz_out float32 [P, 3].  Without transforms / scales the output is bit-identical to the reference's slice. */
int gclb_ingest_points(const float* records, int64_t P, int32_t width, const int64_t* cloud_ptr, int32_t n_clouds,
                       const float* transforms, const float* scales, float* xyz_out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * SC2-PCR registration from putative correspondences, batched over independent problems (scan pairs).
 * Replaces Matcher.SC2_PCR, /root/reference/scripts/SC2_PCR/SC2_PCR.py:304-381 (with pick_seeds :34-58, cal_seed_trans
 * :60-168, cal_leading_eigenvector :170-196, post_refinement :235-274 and common.py:7-45 rigid_transform_3d), called right
 * after feature matching at scripts/test_kitti.py:180-182.
 *   src_xyz, tgt_xyz float32 [sum n_p, 3]: row i of src corresponds to row i of tgt; ptr int64 [n_problems+1] (device)
 *   n_max: host upper bound of the segment lengths; only the first max_points (<= 8192) rows of a segment are used (:321-324)
 *   d_thre, inlier_threshold, nms_radius, ratio, num_iterations, k1, k2 (<= 32): config_json/config_KITTI.json;
 *   refine_iters: 20 in the reference (:378); the refinement threshold is 0.10 when inlier_threshold == 0.10 else 1.2 (:249-252)
 *   trans_out float32 [n_problems, 4, 4] (src -> tgt); info_out int32 [n_problems, 4] = rows used, seeds, inliers of the best
 *   seed hypothesis, inliers of the last refinement iteration (a problem without seeds gets the identity and seeds = 0)
 *   workspace: gclb_sc2pcr_workspace_bytes(n_max, n_problems, ratio)
 * Equal scores resolve to the smaller index (torch.argsort leaves ties unspecified); rotations come from Horn's quaternion
 * form of the weighted Kabsch problem (same optimum as the SVD with the det correction). */
/* Evaluation metrics of a batch of registered pairs in one launch: scripts/test_kitti.py:188-195 (RTE, RRE with the
 * reference's clamp of the trace diagonal) and lib/trainer.py:406-409 (evaluate_hit_ratio on the correspondences).
 *   trans_est, trans_gt float32 [n_pairs, 4, 4]; src_xyz / tgt_xyz float32 [sum n, 3] + ptr int64 [n_pairs+1] (device) or all NULL;
 *   out float32 [n_pairs, 4] = RTE (m), RRE (degrees; NaN where the reference's arccos is NaN), hit ratio, #correspondences */
int gclb_pair_metrics(const float* trans_est, const float* trans_gt, const float* src_xyz, const float* tgt_xyz, const int64_t* ptr,
                      int32_t n_pairs, float hit_thresh, float* out, void* stream);
size_t gclb_sc2pcr_workspace_bytes(int64_t n_max, int32_t n_problems, double ratio);
/* the putative correspondences of a matched batch as coordinates (Matcher.match_pair's return, SC2_PCR.py:297-302), straight
 * from the outputs of gclb_subsample / gclb_nn: row i of segment p = (xyz[unique_map[sel0[i]]], xyz[unique_map[sel1[b_ptr[p] +
 * idx01[i]]]]); unique_map may be NULL (sel* then index xyz directly); n_total_bound >= a_ptr[n_pairs] sizes the grid;
 * src_out / tgt_out float32 [n_total_bound, 3] */
int gclb_corr_points(const float* xyz, const int64_t* unique_map, const int64_t* sel0, const int64_t* sel1,
                     const int64_t* a_ptr, const int64_t* b_ptr, const int64_t* idx01, int32_t n_pairs, int64_t n_total_bound,
                     float* src_out, float* tgt_out, void* stream);
int gclb_sc2pcr(const float* src_xyz, const float* tgt_xyz, const int64_t* ptr, int32_t n_problems, int64_t n_max,
                float d_thre, float inlier_threshold, float nms_radius, double ratio, int32_t num_iterations, int32_t k1,
                int32_t k2, int32_t max_points, int32_t refine_iters, float* trans_out, int32_t* info_out, void* workspace,
                void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Positive-group construction of the GCL colocation loaders:
 * util/pointcloud.py:69-132 get_matching_indices_colocation, called lib/colocation_data_loader.py:394,672.
 * Every cloud is the loader's voxel-downsampled cloud (one point per voxel of size `voxel`, in its own sensor frame), so
 * its voxel hash from gclb_voxelize (table row == point index; the caller checks that no voxel holds two points) stands in
 * for the reference's Open3D KD-tree.
 *   center_xyz float32 [n_center, 3] + center_table/center_capacity           (gclb_voxelize of the centre cloud)
 *   nb_xyz float32 [sum N_j, 3], nb_ptr int64 [J+1] (device), nb_table/nb_capacity (ONE gclb_voxelize call over all
 *   neighbour clouds, cloud index = batch index)
 *   trans, inv_trans float64 [J, 4, 4] row-major (device): cloud j -> centre frame and its inverse
 *   radius: search_voxel_size; K: neighbours kept per cloud, nearest first (K <= 0: every hit, at most kcap <= 32 --
 *   GCLB_ST_FULL if more exist)
 * Outputs (device): group_out int64 [<= n_center] sizes of the kept groups in centre-point order, index_out int64
 * [<= n_center * (1+J) * kcap] concatenated member indices (centre points first, then cloud j offset by
 * n_center + N_0 + .. + N_{j-1}), finest_out uint8 one-hot per group (util/pointcloud.py:106-116), n_groups_out /
 * n_index_out int64 [1].  A hit is a point with squared distance < radius^2 in float64 (nanoflann's rule); ties in distance
 * resolve to the smaller index (implementation-defined in Open3D).
 * ---------------------------------------------------------------------------------------------------- */
size_t gclb_groups_workspace_bytes(int64_t n_center, int32_t n_clouds, int32_t kcap);
int gclb_colocation_groups(const float* center_xyz, int64_t n_center, const void* center_table, int64_t center_capacity,
                           const float* nb_xyz, const int64_t* nb_ptr, const void* nb_table, int64_t nb_capacity,
                           const double* trans, const double* inv_trans, int32_t n_clouds, float voxel, double radius,
                           int32_t K, int32_t kcap, int64_t* group_out, int64_t* index_out, uint8_t* finest_out,
                           int64_t* n_groups_out, int64_t* n_index_out, int32_t* status, void* workspace, void* stream);

/* _exhaustive_hash (util/misc.py:29-36): for every group (members index[group_ptr[g] .. group_ptr[g+1])) the symmetric keys
 * min(a + b*M, a*M + b) of all unordered member pairs, in the reference's order; keys_out int64 [sum n_g (n_g - 1) / 2]
 * (the caller sizes it from the group sizes), n_keys_out int64 [1] (device).  Feeds gclb_group_loss's positive-pair filter. */
size_t gclb_exhaustive_hash_workspace_bytes(int64_t n_groups);
int gclb_exhaustive_hash(const int64_t* group_ptr, const int64_t* index, int64_t n_groups, int64_t M, int64_t* keys_out,
                         int64_t* n_keys_out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCLB200_H_ */
