/* gclb200_debug.h -- bring-up / measurement helpers exported by libgclb200.so next to the product ABI (include/gclb200.h).
 * None of them is on the hot path or replaces a reference call site: they exist so that the micro-architecture numbers quoted in
 * DESIGN.md and profiles/ (TMA tile::gather4 layout, tcgen05.mma issue rate, per-role cycle accounting of the halo kernel) can
 * be re-measured with tools/tma_probe.py, tools/umma_rate.py and tools/haloprof.py. */
#ifndef GCLB200_DEBUG_H_
#define GCLB200_DEBUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bring-up helper (not on the hot path): one TMA tile::gather4 of rows rows4_host[0..3] x channels [col, col+32) of the
 * fp32 matrix X [n, c] into a SWIZZLE_128B shared-memory tile, dumped to out256 (device, 256 floats). */
int gclb_debug_tma_gather4(const float* X, int64_t n, int32_t c, int32_t box_rows, int32_t col, const int32_t* rows4_host,
                           float* out256, void* stream);

/* bring-up helper: per-CTA cycle accounting of the last gclb_spconv_fwd_halo launch run with GCLB_HALO_DBG bit 9 (host uint64 [148][16]) */
int gclb_debug_halo_prof(unsigned long long* out_host);

/* bring-up helper: tcgen05.mma issue / completion cycles (M = 128, N = n, K = 16, kind::f16), out_dev int64[2] */
int gclb_debug_umma_rate(int32_t n, int32_t n_mma, int32_t per_commit, int32_t n_acc, int32_t elect, long long* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCLB200_DEBUG_H_ */
