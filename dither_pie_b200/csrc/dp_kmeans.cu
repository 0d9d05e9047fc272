// dp_kmeans.cu -- Lloyd iterations for k-means palette extraction.
//
// Replaces the E/M steps inside sklearn.cluster.KMeans as called by
// ColorReducer.generate_kmeans_palette (dithering_lib.py:1854-1855).  Pixels are u8, so the
// per-cluster channel sums are exact integers: they are accumulated in u32 per block (shared
// memory atomics), flushed as u64 global atomics, and -- across GPUs -- all-reduced as integers
// by the caller (NCCL), which makes the centres independent of shard count and order.
// Algorithmic bytes: 3 read per pixel per iteration (labels are not materialised).
#include "dp_common.cuh"

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_PIX_PER_BLOCK = 16384;  // 255 * 16384 < 2^32: u32 block partials are exact

__global__ void __launch_bounds__(KM_THREADS) k_kmeans_accumulate(
    const uint8_t *__restrict__ px, long long n, const double *__restrict__ centers, int K,
    unsigned long long *__restrict__ sums)
{
    __shared__ double s_c[DP_MAX_COLORS * 3];
    __shared__ unsigned int s_sum[DP_MAX_COLORS * 4];
    for (int i = threadIdx.x; i < K * 3; i += KM_THREADS) s_c[i] = centers[i];
    const long long nchunks = (n + KM_PIX_PER_BLOCK - 1) / KM_PIX_PER_BLOCK;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        for (int i = threadIdx.x; i < K * 4; i += KM_THREADS) s_sum[i] = 0;
        __syncthreads();
        const long long p0 = chunk * KM_PIX_PER_BLOCK;
        const int cnt = (int)((n - p0) < KM_PIX_PER_BLOCK ? (n - p0) : KM_PIX_PER_BLOCK);
        for (int j = threadIdx.x; j < cnt; j += KM_THREADS) {
            const uint8_t *q = px + (size_t)(p0 + j) * 3;
            const int r = q[0], g = q[1], b = q[2];
            const double x0 = r, x1 = g, x2 = b;
            double best = 1e300;
            int bi = 0;
            for (int i = 0; i < K; ++i) {
                double d0 = x0 - s_c[3 * i], d1 = x1 - s_c[3 * i + 1], d2 = x2 - s_c[3 * i + 2];
                double d = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)),
                                     __dmul_rn(d2, d2));
                if (d < best) {
                    best = d;
                    bi = i;
                }
            }
            atomicAdd(&s_sum[4 * bi], (unsigned)r);
            atomicAdd(&s_sum[4 * bi + 1], (unsigned)g);
            atomicAdd(&s_sum[4 * bi + 2], (unsigned)b);
            atomicAdd(&s_sum[4 * bi + 3], 1u);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < K * 4; i += KM_THREADS)
            if (s_sum[i]) atomicAdd(&sums[i], (unsigned long long)s_sum[i]);
        __syncthreads();
    }
}

__global__ void k_kmeans_update(const unsigned long long *__restrict__ sums, int K,
                                double *__restrict__ centers, double *__restrict__ shift2)
{
    // single thread: K <= 256, the sum order is fixed so the result is deterministic
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double tot = 0.0;
    for (int i = 0; i < K; ++i) {
        unsigned long long c = sums[4 * i + 3];
        for (int ch = 0; ch < 3; ++ch) {
            double old = centers[3 * i + ch];
            double nw = c ? __ddiv_rn((double)sums[4 * i + ch], (double)c) : old;
            double d = nw - old;
            tot = __dadd_rn(tot, __dmul_rn(d, d));
            centers[3 * i + ch] = nw;
        }
    }
    *shift2 = tot;
}

}  // namespace

extern "C" int dp_kmeans_accumulate(const uint8_t *pixels, int64_t n, const double *centers,
                                    int K, unsigned long long *sums, void *stream)
{
    DP_REQUIRE(pixels && centers && sums, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS && n >= 0, "bad size");
    if (n == 0) return 0;
    long long chunks = (n + KM_PIX_PER_BLOCK - 1) / KM_PIX_PER_BLOCK;
    long long cap = (long long)dp_num_sms() * 8;
    int grid = (int)(chunks < cap ? chunks : cap);
    k_kmeans_accumulate<<<grid, KM_THREADS, 0, dp_stream(stream)>>>(pixels, n, centers, K, sums);
    DP_LAUNCH_CHECK();
    return 0;
}

extern "C" int dp_kmeans_update(const unsigned long long *sums, int K, double *centers,
                                double *shift2, void *stream)
{
    DP_REQUIRE(sums && centers && shift2, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS, "bad size");
    k_kmeans_update<<<1, 32, 0, dp_stream(stream)>>>(sums, K, centers, shift2);
    DP_LAUNCH_CHECK();
    return 0;
}
