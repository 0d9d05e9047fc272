/*
 * ditherpie_b200.h -- C ABI of the B200-native per-pixel hot path of dither_pie.
 *
 * One shared library (dither_pie_b200/libditherpie_b200.so, hand-written CUDA for sm_100a).
 * Plain C types only: pointers, sizes, an opaque palette handle and a raw cudaStream_t passed
 * as void*.  Every image pointer is a DEVICE pointer to interleaved 8-bit RGB, row-major
 * [frames][h][w][3], frames contiguous; the *_host helpers at the end take HOST buffers and
 * run the copies inside the call.  All calls are asynchronous on `stream` unless stated.
 *
 * Each entry point names the reference interface it replaces (file:line into
 * dobrosketchkun/dither_pie).  The Python binding a maintainer would add is shown in
 * INTEGRATION.md; dither_pie_b200/_capi.py is that binding.
 *
 * Output planes.  Every dither entry point writes `dst_rgb` (the palette colours, what
 * ImageDitherer.apply_dithering returns, dithering_lib.py:1984-1992) and/or `dst_idx` (one
 * palette row per dithered pixel -- the palette-index image).  Either may be NULL, not both:
 * with dst_rgb == NULL only the index plane is produced (1 byte per pixel instead of 3, which
 * is what a caller that owns the palette needs to bring back over PCIe).
 *
 * Return value: 0 on success, non-zero on error; dp_last_error() returns the message of the
 * last failing call on the calling thread.  There is no CPU fallback anywhere.
 */
#ifndef DITHERPIE_B200_H
#define DITHERPIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DP_MAX_COLORS 256

/* ---------------------------------------------------------------------------------------
 * Library / device / memory plumbing (no reference counterpart: the reference is CPU-only)
 * ------------------------------------------------------------------------------------- */
const char *dp_last_error(void);
int dp_version(void);
int dp_device_count(int *count);
int dp_set_device(int device);
int dp_malloc(void **dptr, size_t bytes);
int dp_free(void *dptr);
int dp_host_alloc(void **hptr, size_t bytes); /* pinned */
int dp_host_free(void *hptr);
int dp_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int dp_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int dp_memset(void *dst, int value, size_t bytes, void *stream);
int dp_stream_create(void **stream);
int dp_stream_destroy(void *stream);
int dp_stream_sync(void *stream); /* NULL = default stream */
/* events: ordering between the copy-in / kernel / copy-out streams of a frame pipeline (the
 * GPU-side replacement of the reference's batch loop, video_processor.py:304-346) and timing */
int dp_event_create(void **event, int timing);
int dp_event_destroy(void *event);
int dp_event_record(void *event, void *stream);
int dp_stream_wait_event(void *stream, void *event);
int dp_event_sync(void *event);
int dp_event_elapsed_ms(void *start, void *stop, float *ms);
/* page-lock a caller-owned host buffer in place (e.g. the numpy array a frame reader fills) so
 * that dp_memcpy_* on it are true asynchronous DMA transfers */
int dp_host_register(void *hptr, size_t bytes);
int dp_host_unregister(void *hptr);
int dp_host_is_pinned(const void *hptr, int *pinned); /* dp_host_alloc'ed or registered memory */
/* NVTX ranges (visible to profilers that collect NVTX; no-ops otherwise) */
int dp_range_push(const char *name);
int dp_range_pop(void);

/* ---------------------------------------------------------------------------------------
 * Palette handle.
 * Replaces the per-call palette set-up of every strategy: `palette_arr` (f32 [K,3]) plus
 * `scipy.spatial.KDTree(palette_arr)` (dithering_lib.py:339, 358, 554, 748, 1229, 1612) and
 * the gamma handling of ImageDitherer.apply_dithering (dithering_lib.py:1956-1974, 1986-1990).
 *
 *   palette      f32 [K,3], the palette in the space the search runs in (linear if gamma)
 *   out_rgb      u8  [K,3], the bytes written for each palette row (reference:
 *                `palette_arr[idx].astype(uint8)` then optional linear->sRGB, :1984-1990)
 *   in_lut       u8  [256] applied to every input byte (sRGB->linear, :1956-1959) or NULL
 *   kd_*         scipy's tree flattened in pre-order (built by the caller with scipy itself,
 *                so the in-leaf order is the installed scipy's by construction):
 *                split_dim (-1 = leaf), split, start_idx, end_idx, lesser, greater per node;
 *                indices [K]; mins/maxes [3].
 * All pointers are HOST pointers; the call copies them and may be followed by frees.
 * ------------------------------------------------------------------------------------- */
typedef struct dp_palette dp_palette;

int dp_palette_create(const float *palette, int K, const uint8_t *out_rgb, const uint8_t *in_lut,
                      int kd_nodes, const int32_t *kd_split_dim, const double *kd_split,
                      const int32_t *kd_start_idx, const int32_t *kd_end_idx,
                      const int32_t *kd_lesser, const int32_t *kd_greater,
                      const int32_t *kd_indices, const double *kd_mins, const double *kd_maxes,
                      dp_palette **out);
int dp_palette_destroy(dp_palette *pal);
int dp_palette_num_colors(const dp_palette *pal);

/* ---------------------------------------------------------------------------------------
 * Geometry shared by the per-pixel kernels: optional fused pixelization in front of the
 * dither (video_processor.py:563-577 `pixelize_regular`) and optional integer up-scale
 * behind it (video_processor.py:393-420, dither_cli.py:559-566).
 *
 *   src_h, src_w   size of each input frame
 *   ytab, xtab     DEVICE int32 tables [h], [w]: source row/column of each dithered pixel
 *                  (Pillow NEAREST mapping computed by the caller), or NULL for identity
 *                  (then h == src_h, w == src_w)
 *   h, w           size of the dithered image
 *   upscale        integer m >= 1: every dithered pixel is written as an m x m block, i.e.
 *                  output frames are [h*m, w*m] (what NEAREST resize to an exact multiple
 *                  does).  Other output sizes: dither first, then dp_resample_nearest.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int32_t src_h, src_w;
    int32_t h, w;
    int32_t upscale;
    const int32_t *ytab, *xtab;
} dp_geometry;

/* Threshold source of the ordered family. */
enum {
    DP_THRESH_NONE = 0,   /* NoDitherStrategy.dither                dithering_lib.py:333-341 */
    DP_THRESH_MATRIX = 1, /* MatrixDitherStrategy.dither (Bayer, blue noise) :355-378 and
                             PolkaDotDitherStrategy.dither                   :745-766 */
    DP_THRESH_IGN = 2     /* InterleavedGradientNoiseDitherStrategy.dither   :539-568 */
};

/*
 * dp_threshold_dither -- nearest / second-nearest palette colour chosen by a threshold.
 *   matrix      DEVICE f32 [mat_h, mat_w] (DP_THRESH_MATRIX), tiled over the dithered image
 *   ign_*       f32 constants of :546-549 already rounded to f32 by the caller:
 *               x' = (x + ign_xoff) * ign_scale,  y' = (y + ign_yoff) * ign_scale
 *   dst_rgb     DEVICE u8 [frames, h*upscale, w*upscale, 3], or NULL (index plane only)
 *   dst_idx     DEVICE u8 [frames, h, w] palette row per dithered pixel, or NULL
 */
int dp_threshold_dither(const dp_palette *pal, const uint8_t *src_rgb, int frames,
                        const dp_geometry *geo, int kind, const float *matrix, int mat_h,
                        int mat_w, float ign_xoff, float ign_yoff, float ign_scale,
                        uint8_t *dst_rgb, uint8_t *dst_idx, void *stream);

/*
 * dp_halftone -- HalftoneDitherStrategy.dither, dithering_lib.py:1597-1644, with the screen and
 * cell map of _generate_halftone_screen_with_cells (:1646-1695).
 *   screen      DEVICE f32 [h,w] or NULL to have the library compute it (only when
 *               dot_gain == 1.0, where no transcendental is needed)
 *   shape       0 circle, 1 square, 2 diamond
 *   cos_a,sin_a cos/sin(radians(angle)) as computed by the caller's libm (f64)
 * Geometry: identity only in this version (no fused pixelize / up-scale).
 */
int dp_halftone(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                int cell_size, double cos_a, double sin_a, double dot_gain, double min_dot,
                double max_dot, int shape, double sharpness, const float *screen,
                uint8_t *dst_rgb, uint8_t *dst_idx, void *stream);

/* Error-diffusion kernels of ErrorDiffusionKernel, dithering_lib.py:107-188. */
enum {
    DP_ED_FLOYD_STEINBERG = 0,
    DP_ED_JJN = 1,
    DP_ED_STUCKI = 2,
    DP_ED_BURKES = 3,
    DP_ED_ATKINSON = 4,
    DP_ED_SIERRA = 5,
    DP_ED_SIERRA_TWO_ROW = 6,
    DP_ED_SIERRA_LITE = 7
};

/*
 * dp_error_diffusion -- ErrorDiffusionDitherStrategy.dither (:631-651) with the semantics of
 * its numba core _error_diffusion_numba (:212-308): f32 state, f64 arithmetic, one f32
 * rounding per accumulation, strict '<' first-index nearest colour.
 * Frames are independent; each frame is a skewed-row wavefront (serpentine: serial rows).
 */
int dp_error_diffusion(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                       int variant, int serpentine, uint8_t *dst_rgb, uint8_t *dst_idx,
                       void *stream);

/*
 * dp_ostromoukhov -- OstromoukhovDitherStrategy.dither, the live path :1225-1269 (f32
 * arithmetic, f32-rounded weights, KD-tree nearest).  coeffs: HOST int32 [256,3] (:1170-1203).
 */
int dp_ostromoukhov(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                    const int32_t *coeffs, int serpentine, uint8_t *dst_rgb, uint8_t *dst_idx,
                    void *stream);

/*
 * dp_hybrid -- HybridDitherStrategy.dither (:1071-1155) with the semantics of its numba core
 * _hybrid_numba (:1396-1494), the canonical path when numba is importable: the Floyd-Steinberg
 * raster loop of dp_error_diffusion (f32 state, f64 arithmetic, clamp before the strict '<'
 * first-index lookup) whose error is split into a luminance part and a colour part before it
 * is distributed:  lum = (0.299 e0 + 0.587 e1) + 0.114 e2;  l_c = coef_c * lum;
 * fe_c = lum_factor * l_c + col_factor * (e_c - l_c), every operation a separately rounded f64.
 * (same wavefront skeleton, different error transform.)
 */
int dp_hybrid(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
              double lum_factor, double col_factor, uint8_t *dst_rgb, uint8_t *dst_idx,
              void *stream);

/*
 * dp_perceptual -- PerceptualDitherStrategy.dither (:1040-1066, pure Python in the reference,
 * default base_weights = Floyd-Steinberg): all-f32 diffusion, KD-tree nearest of the UNCLAMPED
 * work value, every tap scaled by the luminance factor 0.5 + 0.5 (gray / 255) of the ORIGINAL
 * pixel (f32, one rounding per operation).  Same wavefront as dp_error_diffusion; the factor
 * plane is produced by a small per-pixel kernel first.  
 */
int dp_perceptual(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                  uint8_t *dst_rgb, uint8_t *dst_idx, void *stream);

/*
 * dp_adaptive_variance -- AdaptiveVarianceDitherStrategy.dither (:989-1025, pure Python in the
 * reference): all-f32 Floyd-Steinberg on the UNCLAMPED work values with KD-tree nearest, the
 * error of a pixel being distributed only where the local variance of the original gray image
 * (scipy.ndimage.uniform_filter of gray and gray^2, size 2*window_radius+1, mode 'nearest';
 * :1021-1025) is >= var_threshold.  The gate plane is computed on the device by replaying
 * scipy's running-sum filter, then the weighted wavefront of dp_perceptual runs with factors 0/1.
 */
int dp_adaptive_variance(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                         double var_threshold, int window_radius, uint8_t *dst_rgb,
                         uint8_t *dst_idx, void *stream);

/*
 * dp_unique_colors_pyset_order -- HOST helper (no device work) for the default palette source:
 * `unique_cols = list(set(image.getdata()))` in ColorReducer.reduce_colors
 * (dithering_lib.py:1837).  Median cut depends on the ITERATION ORDER of that CPython set (stable
 * sorts keep it among equal keys), so the unique colours are returned in exactly that order by
 * replaying CPython's tuple hash and set table (see csrc/dp_pyset.cu).  SURVEY.md section 8(f) #1.
 *   rgb      HOST u8 [npix,3]
 *   out_rgb  HOST u8 [npix,3] capacity; the first *n_unique rows are written
 */
int dp_unique_colors_pyset_order(const uint8_t *rgb, int64_t npix, uint8_t *out_rgb,
                                 int64_t *n_unique);

/*
 * dp_blue_noise_from_order -- HOST helper: generate_blue_noise (dithering_lib.py:381-399), the
 * farthest-point ordering of a shuffled coordinate list, replayed in integer arithmetic.
 *   order  HOST int32 [size*size]: the shuffled list as flat indices r*size + c
 *          (np.random.RandomState(seed).shuffle, as the reference does at :388-389)
 *   out    HOST f32 [size,size]
 */
int dp_blue_noise_from_order(const int32_t *order, int size, float *out);

/*
 * dp_resample_nearest -- Image.resize(NEAREST) as used by pixelize_regular
 * (video_processor.py:576) and the final up-scale (:419, dither_cli.py:565):
 * dst[f,y,x] = src[f, ytab[y], xtab[x]].  Tables are DEVICE int32.
 */
int dp_resample_nearest(const uint8_t *src_rgb, int frames, int src_h, int src_w,
                        const int32_t *ytab, const int32_t *xtab, int dst_h, int dst_w,
                        uint8_t *dst_rgb, void *stream);

/*
 * K-means palette extraction -- the Lloyd iterations inside
 * ColorReducer.generate_kmeans_palette (dithering_lib.py:1854-1855 -> sklearn KMeans).
 *
 * dp_kmeans_accumulate: one assignment pass.  For every pixel find the nearest centre (f64,
 *   first index on ties) and add it to that centre's exact integer sums.
 *     pixels   DEVICE u8 [n,3]
 *     centers  DEVICE f64 [K,3]
 *     sums     DEVICE u64 [K,4] = (sum r, sum g, sum b, count); ACCUMULATED into (zero it
 *              first); integer sums make the result independent of shard count and order,
 *              so a multi-GPU run all-reduces `sums` (NCCL, done by the caller) and gets
 *              bit-identical centres for any number of ranks.
 * dp_kmeans_update: centres <- sums/count (empty clusters keep their centre), writes the
 *   squared centre shift (sklearn's stopping quantity) to shift2 (DEVICE f64 [1]).
 */
int dp_kmeans_accumulate(const uint8_t *pixels, int64_t n, const double *centers, int K,
                         unsigned long long *sums, void *stream);
int dp_kmeans_update(const unsigned long long *sums, int K, double *centers, double *shift2,
                     void *stream);

/*
 * dp_kmeans_lloyd -- the whole Lloyd loop of KMeans.fit (sklearn _kmeans.py:630-760 as called from
 * dithering_lib.py:1854-1855) from a given initialisation, with the stop test kept on the device.
 * K <= 32 without a communicator: ONE persistent launch runs the loop (centres from the previous
 * sums, squared centre shift, stop decision, candidate grid, assignment pass; the blocks meet at
 * grid barriers).  Otherwise per iteration one "prepare" launch and one assignment launch; with
 * a communicator the K*4+1 integer sums are all-reduced (ncclAllReduce, same stream) between
 * them, and the host reads the stop flag every `check_every` iterations only.
 *   pixels        DEVICE u8 [n,3] -- this rank's shard of the pixels
 *   centers_host  HOST f64 [K,3], in: initial centres, out: final centres (identical on all ranks)
 *   tol           stop when the squared centre shift is <= tol (sklearn: tol * mean variance)
 *   nccl_comm     communicator from dp_nccl_comm_create, or NULL (single GPU)
 *   n_iter, shift2, ties, empty_iters   HOST outputs (any may be NULL): iterations done, last
 *                 squared shift, number of samples exactly equidistant from their two nearest
 *                 centres summed over the iterations (a non-zero count marks a run that sklearn's
 *                 GEMM rounding may resolve differently), iterations that
 *                 saw an empty cluster (it keeps its centre; sklearn relocates it)
 * Synchronous: returns when the loop has stopped.
 */
int dp_kmeans_lloyd(const uint8_t *pixels, int64_t n, double *centers_host, int K, double tol,
                    int max_iter, void *nccl_comm, int check_every, int *n_iter, double *shift2,
                    unsigned long long *ties, int *empty_iters, void *stream);

/*
 * dp_kmeans_lloyd_p2p -- the same loop with the exchange done by the kernels themselves over peer
 * memory (NVLink) instead of an NCCL launch: after its assignment pass a rank stores its K*4+1
 * integers into a slot of every rank's INBOX and releases a flag; the next "prepare" step waits
 * for the `world` flags and adds the slots up (integers: identical totals on every rank).  For
 * K <= 32 this happens inside the persistent loop kernel, else in the launch pair's kernels.
 *   inboxes   HOST array of `world` DEVICE pointers; inboxes[rank] is this rank's own inbox
 *             (dp_p2p_alloc(dp_p2p_inbox_bytes(), ...)), the others are the peers' inboxes opened
 *             with dp_p2p_open from the handles the ranks exchanged
 *   epoch     a counter the caller increments per call (same value on every rank): flags only grow.
 *             The ranks must pass a barrier between two calls on the same inboxes.
 * At most 8 ranks (one NVSwitch domain).  Returns 3 if a peer did not deliver within ~10 s.
 */
int dp_kmeans_lloyd_p2p(const uint8_t *pixels, int64_t n, double *centers_host, int K, double tol,
                        int max_iter, int rank, int world, void *const *inboxes,
                        unsigned long long epoch, int check_every, int *n_iter, double *shift2,
                        unsigned long long *ties, int *empty_iters, void *stream);
/* peer-visible device memory between the processes of one box (cudaIpc) */
int dp_p2p_inbox_bytes(void);
int dp_p2p_alloc(size_t bytes, void **dptr, void *ipc_handle64);
int dp_p2p_open(const void *ipc_handle64, void **dptr);
int dp_p2p_close(void *dptr);   /* a mapping from dp_p2p_open */
int dp_p2p_free(void *dptr);    /* an allocation from dp_p2p_alloc */

/*
 * NCCL plumbing for the sharded k-means (one process per GPU).  libnccl is resolved at run time
 * (dlopen: the copy the process already loaded -- e.g. PyTorch's -- or the system one), the
 * library has no link-time dependency on it.  Rank 0 creates the id and hands the 128 bytes to
 * the other ranks by any means (torch.distributed broadcast in dither_pie_b200/distributed.py).
 */
int dp_nccl_load(const char *path_or_null);
int dp_nccl_unique_id(void *id128);
int dp_nccl_comm_create(const void *id128, int rank, int world, void **comm);
int dp_nccl_comm_destroy(void *comm);
int dp_nccl_allreduce_u64(void *buf, size_t count, void *comm, void *stream);

/* ---------------------------------------------------------------------------------------
 * Host-buffer convenience (the end-to-end path a Python caller with numpy arrays uses):
 * copies src to the device, runs the kernel, copies the result back, synchronises.
 * ------------------------------------------------------------------------------------- */
int dp_threshold_dither_host(const dp_palette *pal, const uint8_t *src_rgb_host, int frames,
                             int h, int w, int kind, const float *matrix_host, int mat_h,
                             int mat_w, float ign_xoff, float ign_yoff, float ign_scale,
                             uint8_t *dst_rgb_host);

#ifdef __cplusplus
}
#endif
#endif /* DITHERPIE_B200_H */
