// dp_resample.cu -- nearest-neighbour resampling by per-axis index tables.
//
// Replaces Image.resize(..., NEAREST) as used by pixelize_regular (video_processor.py:576) and
// the final up-scale (video_processor.py:419, dither_cli.py:565).  The caller computes the
// tables with Pillow's running-sum mapping; the GPU only gathers.
// Algorithmic bytes: 3 read per distinct source pixel touched + 3 written per output pixel.
#include "dp_common.cuh"

namespace {

// One block per (output row segment): threads walk the row in 4-byte words so that stores are
// coalesced 32-bit writes; the source row is a gather through xtab.
__global__ void __launch_bounds__(256) k_resample(const uint8_t *__restrict__ src, int src_h,
                                                  int src_w, const int *__restrict__ ytab,
                                                  const int *__restrict__ xtab, int dst_h,
                                                  int dst_w, uint8_t *__restrict__ dst,
                                                  long long rows_total)
{
    const size_t src_frame = (size_t)src_h * src_w * 3;
    const size_t row_bytes = (size_t)dst_w * 3;
    for (long long row = blockIdx.x; row < rows_total; row += gridDim.x) {
        const int f = (int)(row / dst_h);
        const int y = (int)(row - (long long)f * dst_h);
        const uint8_t *srow = src + (size_t)f * src_frame + (size_t)__ldg(ytab + y) * src_w * 3;
        uint8_t *drow = dst + (size_t)row * row_bytes;
        for (int x = threadIdx.x; x < dst_w; x += blockDim.x) {
            const uint8_t *q = srow + (size_t)__ldg(xtab + x) * 3;
            uint8_t *o = drow + (size_t)x * 3;
            o[0] = q[0];
            o[1] = q[1];
            o[2] = q[2];
        }
    }
}

}  // namespace

extern "C" int dp_resample_nearest(const uint8_t *src_rgb, int frames, int src_h, int src_w,
                                   const int32_t *ytab, const int32_t *xtab, int dst_h,
                                   int dst_w, uint8_t *dst_rgb, void *stream)
{
    DP_RANGE("dp_resample_nearest");
    DP_REQUIRE(src_rgb && dst_rgb && ytab && xtab, "null argument");
    DP_REQUIRE(frames >= 0 && src_h > 0 && src_w > 0 && dst_h >= 0 && dst_w >= 0, "bad size");
    long long rows = (long long)frames * dst_h;
    if (rows == 0 || dst_w == 0) return 0;
    int grid = (int)(rows < (long long)dp_num_sms() * 16 ? rows : (long long)dp_num_sms() * 16);
    k_resample<<<grid, 256, 0, dp_stream(stream)>>>(src_rgb, src_h, src_w, ytab, xtab, dst_h,
                                                    dst_w, dst_rgb, rows);
    DP_LAUNCH_CHECK();
    return 0;
}
