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

// Per-iteration candidate grid (the argmin structure of the dither kernels, rebuilt for the moving
// centres): colour space in 16^3 boxes of 16^3 byte colours; a centre is dropped from a box when
// another centre is strictly nearer at EVERY point of the box -- the difference of two squared
// distances is linear in the point, so its maximum sits at a corner (exact test in double with a
// 1e-6 margin).  One warp per box, lane = centre.  Entry: up to four surviving centres, ascending,
// one per byte (255 = none); 0xffffffff = more than four (the pixel loop then scans all centres).
__global__ void __launch_bounds__(256) k_kmeans_grid(const double *__restrict__ centers, int K,
                                                     uint32_t *__restrict__ grid)
{
    const int cell = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (cell >= 4096) return;
    const double lo[3] = {16.0 * (cell >> 8), 16.0 * ((cell >> 4) & 15), 16.0 * (cell & 15)};
    bool alive = false;
    if (lane < K) {
        const double ci[3] = {centers[3 * lane], centers[3 * lane + 1], centers[3 * lane + 2]};
        const double ni = ci[0] * ci[0] + ci[1] * ci[1] + ci[2] * ci[2];
        alive = true;
        for (int j = 0; j < K && alive; ++j) {
            if (j == lane) continue;
            const double cj[3] = {centers[3 * j], centers[3 * j + 1], centers[3 * j + 2]};
            double mx = (cj[0] * cj[0] + cj[1] * cj[1] + cj[2] * cj[2]) - ni;   // |c_j|^2 - |c_i|^2
            for (int a = 0; a < 3; ++a) {
                const double dlt = ci[a] - cj[a];
                mx += 2.0 * (dlt > 0.0 ? lo[a] + 15.0 : lo[a]) * dlt;
            }
            if (mx < -1e-6) alive = false;   // centre j is strictly nearer everywhere in the box
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) {
        uint32_t e = 0xffffffffu;
        if (__popc(m) <= 4) {
            e = 0;
            for (int k = 0; k < 4; ++k) {
                const unsigned idx = m ? (unsigned)(__ffs(m) - 1) : 255u;
                if (m) m &= m - 1;
                e |= idx << (8 * k);
            }
        }
        grid[cell] = e;
    }
}

// K <= 32: no atomics in the pixel loop.  A warp labels 32 pixels at a time; the labels are
// screened in f32 against all centres (two smallest distances kept) and only pixels whose two
// best distances are closer than the f32 error bound repeat the exact f64 comparison (strict
// '<', first index).  The per-cluster sums of the 32 pixels are formed with warp reductions
// (REDUX) cluster by cluster -- neighbouring pixels fall into a handful of clusters -- and lane
// k keeps the running 64-bit totals of cluster k in registers; one flush per warp at the end.
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_accumulate_warp(
    const uint8_t *__restrict__ px, long long n, const double *__restrict__ centers, int K,
    unsigned long long *__restrict__ sums, const uint32_t *__restrict__ grid)
{
    __shared__ double s_c[32 * 3];
    __shared__ float4 s_cf[33];          // [32] = pad entry, far outside the colour cube
    __shared__ uint32_t s_grid[4096];
    if (threadIdx.x < K * 3) s_c[threadIdx.x] = centers[threadIdx.x];
    if (threadIdx.x < K)
        s_cf[threadIdx.x] = make_float4((float)centers[3 * threadIdx.x], (float)centers[3 * threadIdx.x + 1],
                                        (float)centers[3 * threadIdx.x + 2], 0.f);
    if (threadIdx.x == 32) s_cf[32] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
    for (int i = threadIdx.x; i < 4096; i += KM_THREADS) s_grid[i] = grid[i];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long nchunk = (n + 31) >> 5;
    const long long wstride = (long long)gridDim.x * (KM_THREADS / 32);
    unsigned long long ar = 0, ag = 0, ab = 0, an = 0;     // totals of cluster `lane`
    for (long long ch = (long long)blockIdx.x * (KM_THREADS / 32) + (threadIdx.x >> 5); ch < nchunk;
         ch += wstride) {
        const long long i = ch * 32 + lane;
        const bool ok = i < n;
        int r = 0, g = 0, b = 0, label = -1;
        if (ok) {
            const uint8_t *q = px + (size_t)i * 3;
            r = q[0];
            g = q[1];
            b = q[2];
            const float fr = (float)r, fg = (float)g, fb = (float)b;
            float d1 = 3.0e38f, d2 = 3.0e38f;
            int bi = 0;
            // the box of the pixel lists every centre that can be nearest to it (k_kmeans_grid)
            const uint32_t e = s_grid[((r >> 4) << 8) | ((g >> 4) << 4) | (b >> 4)];
            const bool few = e != 0xffffffffu;
            if (few) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {   // ascending candidates; 255 -> the pad entry
                    const int k = (int)((e >> (8 * q)) & 255u);
                    const float4 c = s_cf[k < 32 ? k : 32];
                    const float e0 = fr - c.x, e1 = fg - c.y, e2 = fb - c.z;
                    const float d = fmaf(e2, e2, fmaf(e1, e1, e0 * e0));
                    d2 = fminf(d2, fmaxf(d1, d));
                    bi = d < d1 ? k : bi;
                    d1 = fminf(d1, d);
                }
            } else {
#pragma unroll 4
                for (int k = 0; k < K; ++k) {     // branch-free: lanes of a warp disagree on every test
                    const float4 c = s_cf[k];
                    const float e0 = fr - c.x, e1 = fg - c.y, e2 = fb - c.z;
                    const float d = fmaf(e2, e2, fmaf(e1, e1, e0 * e0));
                    d2 = fminf(d2, fmaxf(d1, d));
                    bi = d < d1 ? k : bi;
                    d1 = fminf(d1, d);
                }
            }
            // centre rounded to f32 (<= 255 * 2^-24) and the rounded difference give |e32 - e| <= 3.1e-5
            // per channel, so |d32 - D| <= 1.07e-4 sqrt(D) + 1.8e-7 D; twice that (both distances)
            // is below the margin used here for every D in [0, 195075]  (d2 = pad: never ambiguous)
            if (d2 < 1.0e30f && d2 - d1 <= 3.0e-4f * sqrtf(d2) + 5.0e-7f * d2 + 3.0e-4f) {
                const double x0 = r, x1 = g, x2 = b;
                double best = 1e300;
                for (int k = 0; k < K; ++k) {
                    const double f0 = x0 - s_c[3 * k], f1 = x1 - s_c[3 * k + 1], f2 = x2 - s_c[3 * k + 2];
                    const double d = __dadd_rn(__dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1)),
                                               __dmul_rn(f2, f2));
                    if (d < best) {
                        best = d;
                        bi = k;
                    }
                }
            }
            label = bi;
        }
        // clusters present among the 32 pixels, one after the other (warp-uniform loop)
        unsigned todo = __ballot_sync(FULL, ok);
        while (todo) {
            const int src = __ffs(todo) - 1;
            const int k = __shfl_sync(FULL, label, src);
            const bool mine = ok && label == k;
            const unsigned m = __ballot_sync(FULL, mine);
            const unsigned sr = __reduce_add_sync(FULL, mine ? (unsigned)r : 0u);
            const unsigned sg = __reduce_add_sync(FULL, mine ? (unsigned)g : 0u);
            const unsigned sb = __reduce_add_sync(FULL, mine ? (unsigned)b : 0u);
            if (lane == k) {
                ar += sr;
                ag += sg;
                ab += sb;
                an += __popc(m);
            }
            todo &= ~m;
        }
    }
    // block totals first (the global accumulators are 4K addresses shared by every block: one
    // flush per block instead of one per warp keeps the L2 atomic unit out of the critical path)
    __shared__ unsigned long long s_tot[32 * 4];
    if (threadIdx.x < 32 * 4) s_tot[threadIdx.x] = 0;
    __syncthreads();
    if (lane < K && an) {
        atomicAdd(&s_tot[4 * lane], ar);
        atomicAdd(&s_tot[4 * lane + 1], ag);
        atomicAdd(&s_tot[4 * lane + 2], ab);
        atomicAdd(&s_tot[4 * lane + 3], an);
    }
    __syncthreads();
    if (threadIdx.x < K * 4 && s_tot[threadIdx.x]) atomicAdd(&sums[threadIdx.x], s_tot[threadIdx.x]);
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
    long long cap = (long long)dp_num_sms() * 8;
    if (K <= 32) {
        long long blocks = ((n + 31) / 32 + KM_THREADS / 32 - 1) / (KM_THREADS / 32);
        int grid = (int)(blocks < cap ? blocks : cap);   // 8 resident blocks of 8 warps per SM
        cudaStream_t st = dp_stream(stream);
        uint32_t *cgrid = nullptr;
        DP_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&cgrid), 4096 * sizeof(uint32_t), st));
        k_kmeans_grid<<<512, 256, 0, st>>>(centers, K, cgrid);
        k_kmeans_accumulate_warp<<<grid, KM_THREADS, 0, st>>>(pixels, n, centers, K, sums, cgrid);
        const cudaError_t le = cudaGetLastError();
        cudaFreeAsync(cgrid, st);
        DP_CUDA(le);
        return 0;
    }
    long long chunks = (n + KM_PIX_PER_BLOCK - 1) / KM_PIX_PER_BLOCK;
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
