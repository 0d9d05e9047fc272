// dp_diffusion.cu -- the error-diffusion family as a skewed-row wavefront.
//
// Replaces ErrorDiffusionDitherStrategy.dither / _error_diffusion_numba
// (dithering_lib.py:631-651, 212-308; tap tables :107-188) and the live path of
// OstromoukhovDitherStrategy.dither (:1225-1269).
//
// Parallelisation (non-serpentine).  A warp owns a band of 32 consecutive rows, lane = row.
// Lane l works S pixels behind lane l-1 (the skew S is the smallest the tap footprint allows).
// The reference scatters each pixel's error into the f32 work buffer with one f32 rounding per
// '+='; the value a pixel finally has is therefore a CHAIN of roundings whose order is the
// raster order of its sources (row y-2, then y-1, then the two left neighbours).  That chain
// is kept intact here by passing the accumulator itself down the lanes:
//   * lane (row y) holds a sliding window d2[] of accumulators of row y+2 (it seeds them with
//     the raw pixel and adds its dy=2 taps), and a window d1[] of row y+1 (seeded with what
//     the lane above finished in ITS d2 window; this lane adds its dy=1 taps);
//   * when a window entry can receive nothing more from this row it moves one lane down by
//     warp shuffle (d2 -> next lane's d1, d1 -> next lane's own pixel);
//   * same-row taps (dx=1,2; dy=0) are kept as pending f64 products and applied when the
//     pixel's accumulator arrives, i.e. after all contributions of the rows above -- exactly the
//     reference's order.
// No atomics, no work buffer: the diffusion state lives in registers.  Between bands the last
// lane writes the two outgoing streams to global memory and publishes a progress counter
// (st.release); lane 0 of the next band polls it (ld.acquire) -- the flag-based hand-off.
// Bands are handed out through an atomic ticket so that a band's predecessor is always
// already running (no deadlock whatever the grid size).
//
// Arithmetic is the reference's: numba path = f32 state, f64 math, strict '<' first-index
// nearest colour; Ostromoukhov = f32 math, f32-rounded weights, KD-tree nearest.
// Serpentine scanning makes row y+1 depend on the END of row y: it is serial per frame and
// handled by a one-warp-per-frame kernel (parallel over frames only).
//
// Algorithmic bytes: 3 read + 3 written per pixel; the bound is the dependency chain
// (W + S*(H-1) pixel steps per frame) and fp64 issue, not HBM.
#include <utility>

#include "dp_search.cuh"

namespace {

constexpr int V_OSTRO = 8;

struct Tap {
    int dx, dy, w;
};

__host__ __device__ constexpr int ed_ntaps(int v)
{
    return v == DP_ED_FLOYD_STEINBERG ? 4 : v == DP_ED_JJN ? 12 : v == DP_ED_STUCKI ? 12
         : v == DP_ED_BURKES ? 7 : v == DP_ED_ATKINSON ? 6 : v == DP_ED_SIERRA ? 10
         : v == DP_ED_SIERRA_TWO_ROW ? 7 : v == DP_ED_SIERRA_LITE ? 3 : /*ostro*/ 3;
}

__host__ __device__ constexpr int ed_divisor(int v)
{
    return v == DP_ED_FLOYD_STEINBERG ? 16 : v == DP_ED_JJN ? 48 : v == DP_ED_STUCKI ? 42
         : v == DP_ED_BURKES ? 32 : v == DP_ED_ATKINSON ? 8 : v == DP_ED_SIERRA ? 32
         : v == DP_ED_SIERRA_TWO_ROW ? 16 : v == DP_ED_SIERRA_LITE ? 4 : 1;
}

// dithering_lib.py:107-188, in the reference's order
__host__ __device__ constexpr Tap ed_tap(int v, int k)
{
    constexpr Tap FS[4] = {{1, 0, 7}, {-1, 1, 3}, {0, 1, 5}, {1, 1, 1}};
    constexpr Tap JJN[12] = {{1, 0, 7}, {2, 0, 5}, {-2, 1, 3}, {-1, 1, 5}, {0, 1, 7}, {1, 1, 5},
                             {2, 1, 3}, {-2, 2, 1}, {-1, 2, 3}, {0, 2, 5}, {1, 2, 3}, {2, 2, 1}};
    constexpr Tap STU[12] = {{1, 0, 8}, {2, 0, 4}, {-2, 1, 2}, {-1, 1, 4}, {0, 1, 8}, {1, 1, 4},
                             {2, 1, 2}, {-2, 2, 1}, {-1, 2, 2}, {0, 2, 4}, {1, 2, 2}, {2, 2, 1}};
    constexpr Tap BUR[7] = {{1, 0, 8}, {2, 0, 4}, {-2, 1, 2}, {-1, 1, 4}, {0, 1, 8}, {1, 1, 4},
                            {2, 1, 2}};
    constexpr Tap ATK[6] = {{1, 0, 1}, {2, 0, 1}, {-1, 1, 1}, {0, 1, 1}, {1, 1, 1}, {0, 2, 1}};
    constexpr Tap SIE[10] = {{1, 0, 5}, {2, 0, 3}, {-2, 1, 2}, {-1, 1, 4}, {0, 1, 5}, {1, 1, 4},
                             {2, 1, 2}, {-1, 2, 2}, {0, 2, 3}, {1, 2, 2}};
    constexpr Tap S2R[7] = {{1, 0, 4}, {2, 0, 3}, {-2, 1, 1}, {-1, 1, 2}, {0, 1, 3}, {1, 1, 2},
                            {2, 1, 1}};
    constexpr Tap SLT[3] = {{1, 0, 2}, {-1, 1, 1}, {0, 1, 1}};
    // Ostromoukhov's footprint (:1258-1266); weights come from the coefficient table
    constexpr Tap OST[3] = {{1, 0, 0}, {-1, 1, 1}, {0, 1, 2}};
    return v == DP_ED_FLOYD_STEINBERG ? FS[k] : v == DP_ED_JJN ? JJN[k] : v == DP_ED_STUCKI ? STU[k]
         : v == DP_ED_BURKES ? BUR[k] : v == DP_ED_ATKINSON ? ATK[k] : v == DP_ED_SIERRA ? SIE[k]
         : v == DP_ED_SIERRA_TWO_ROW ? S2R[k] : v == DP_ED_SIERRA_LITE ? SLT[k] : OST[k];
}

__host__ __device__ constexpr int ed_extent(int v, int dy, int sign)
{
    int m = 0;
    for (int k = 0; k < ed_ntaps(v); ++k) {
        Tap t = ed_tap(v, k);
        if (t.dy == dy && sign * t.dx > m) m = sign * t.dx;
    }
    return m;
}

__host__ __device__ constexpr bool ed_has(int v, int dx, int dy)
{
    for (int k = 0; k < ed_ntaps(v); ++k) {
        Tap t = ed_tap(v, k);
        if (t.dx == dx && t.dy == dy) return true;
    }
    return false;
}

__host__ __device__ constexpr bool ed_rows3(int v)
{
    for (int k = 0; k < ed_ntaps(v); ++k)
        if (ed_tap(v, k).dy == 2) return true;
    return false;
}

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int V>
struct Spec {
    static constexpr bool OSTRO = (V == V_OSTRO);
    static constexpr int N = ed_ntaps(V);
    static constexpr bool ROWS3 = ed_rows3(V);
    static constexpr int A1 = ed_extent(V, 1, -1), B1 = ed_extent(V, 1, 1);
    static constexpr int A2 = ed_extent(V, 2, -1), B2 = ed_extent(V, 2, 1);
    static constexpr int W1 = A1 + B1 + 1, W2 = A2 + B2 + 1;
    static constexpr bool H10 = ed_has(V, 1, 0), H20 = ed_has(V, 2, 0);
    static constexpr int S = cmax(A1 + 1, ROWS3 ? A2 + B1 + 1 : 0);
    static constexpr int DA = S - A1;                      // steps between emit and use, stream A
    static constexpr int DB = ROWS3 ? S - A2 - B1 : 1;     // stream B
    static constexpr int AMAX = cmax(A1, ROWS3 ? A2 : 0);
    static constexpr int BMAX = cmax(B1, ROWS3 ? B2 : 0);
};

constexpr int WAVE_THREADS = 256;

struct WaveParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, nbands, total_units, has_lut, K;
    float *hand;        // [units][2][w][3] f32 hand-off streams
    int *progress;      // [units]
    int *ticket;
    const float *ostro_w;  // [256][4] f32 weights (c0,c1,c2)/sum, DEVICE (ostromoukhov only)
};

// Progress poll.  Relaxed (no L1 invalidation): everything the consumer reads after the poll is
// fetched with ld.global.cg (L2 only) and the producer's st.release orders its data before the
// flag in L2; loads issue in program order behind the branch on the polled value.
__device__ __forceinline__ int ld_poll(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// accumulate: f32( f64(acc) + prod )   -- the reference's `work[ny, nx, c] += err * wgt`
__device__ __forceinline__ float acc_f64(float acc, double prod)
{
    return __double2float_rn(__dadd_rn((double)acc, prod));
}

struct Search {
    const uint2 *table;      // shared, [4096] 16^3 cells
    const uint8_t *ovf;      // global overflow lists
    const double *s_pal;     // shared, [K,3]
};

__device__ __forceinline__ int cell_of(double r, double g, double b)
{
    const int ir = min(__double2int_rz(r), 255) >> 4;
    const int ig = min(__double2int_rz(g), 255) >> 4;
    const int ib = min(__double2int_rz(b), 255) >> 4;
    return (ir << 8) | (ig << 4) | ib;
}

__device__ __forceinline__ double dist_numba(const double *pp, double r, double g, double b)
{
    const double dr = __dsub_rn(r, pp[0]), dg = __dsub_rn(g, pp[1]), db = __dsub_rn(b, pp[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dg, dg)), __dmul_rn(db, db));
}

// numba path (:254-263): strict '<' over f64 distances, first index wins.  The candidate list
// of the pixel's 16^3 cell holds every row that can be nearest there, in ascending order.
__device__ __forceinline__ int nearest_first(const Search &s, double r, double g, double b)
{
    const uint2 e = s.table[cell_of(r, g, b)];
    const unsigned n = e.x & 255u;
    double best = 1e20;
    int bi = 0;
    if (n != 255u) {
        unsigned long long ev = ((unsigned long long)e.y << 32 | e.x) >> 8;
        for (unsigned j = 0; j < n; ++j, ev >>= 8) {
            const int i = (int)(ev & 255u);
            const double d = dist_numba(s.s_pal + 3 * i, r, g, b);
            if (d < best) {
                best = d;
                bi = i;
            }
        }
    } else {
        const unsigned cnt = e.x >> 8;
        const uint8_t *lst = s.ovf + e.y;
        for (unsigned j = 0; j < cnt; ++j) {
            const int i = __ldg(lst + j);
            const double d = dist_numba(s.s_pal + 3 * i, r, g, b);
            if (d < best) {
                best = d;
                bi = i;
            }
        }
    }
    return bi;
}

__device__ __forceinline__ double dist_scipy(const double *pp, double r, double g, double b)
{
    const double d0 = __dsub_rn(pp[0], r), d1 = __dsub_rn(pp[1], g), d2 = __dsub_rn(pp[2], b);
    return __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(d0, d0)), __dmul_rn(d1, d1)),
                     __dmul_rn(d2, d2));
}

// KD-tree nearest (:1243): unique minimum among the candidates, else replay scipy.
__device__ __forceinline__ int nearest_kd(const PalDev *P, const Search &s, double r, double g,
                                          double b)
{
    const uint2 e = s.table[cell_of(r, g, b)];
    const unsigned n = e.x & 255u;
    double best = DP_INF_F64;
    int bi = 0;
    bool tie = false;
    const bool inl = n != 255u;
    const unsigned cnt = inl ? n : (e.x >> 8);
    unsigned long long ev = ((unsigned long long)e.y << 32 | e.x) >> 8;
    const uint8_t *lst = s.ovf + e.y;
    for (unsigned j = 0; j < cnt; ++j, ev >>= 8) {
        const int i = inl ? (int)(ev & 255u) : (int)__ldg(lst + j);
        const double d = dist_scipy(s.s_pal + 3 * i, r, g, b);
        if (d < best) {
            best = d;
            bi = i;
            tie = false;
        } else if (d == best) {
            tie = true;
        }
    }
    if (tie) {
        int oi[1];
        double os[1];
        kd_emulate<1>(P, r, g, b, oi, os);
        bi = oi[0];
    }
    return bi;
}

// One tap, everything about it known at compile time (weight = f64(f32(w)) / divisor, the
// reference's `weights[k] / divisor`, :280).
template <int V, int K, int W1, int W2>
__device__ __forceinline__ void apply_tap(const double (&e)[3], double (&q10)[3],
                                          double (&q20a)[3], double (&q20b)[3],
                                          float (&d1)[W1][3], float (&d2)[W2][3])
{
    using SP = Spec<V>;
    constexpr Tap tp = ed_tap(V, K);
    constexpr double wgt = (double)(float)tp.w / (double)ed_divisor(V);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double pr = __dmul_rn(e[c], wgt);
        if (tp.dy == 0) {
            if (tp.dx == 1) {
                q10[c] = pr;
            } else {
                q20a[c] = q20b[c];
                q20b[c] = pr;
            }
        } else if (tp.dy == 1) {
            d1[tp.dx + SP::A1][c] = acc_f64(d1[tp.dx + SP::A1][c], pr);
        } else {
            d2[(tp.dy == 2 ? tp.dx + SP::A2 : 0)][c] =
                acc_f64(d2[(tp.dy == 2 ? tp.dx + SP::A2 : 0)][c], pr);
        }
    }
}

template <int V, int... Ks, int W1, int W2>
__device__ __forceinline__ void apply_taps(std::integer_sequence<int, Ks...>, const double (&e)[3],
                                           double (&q10)[3], double (&q20a)[3],
                                           double (&q20b)[3], float (&d1)[W1][3],
                                           float (&d2)[W2][3])
{
    (apply_tap<V, Ks, W1, W2>(e, q10, q20a, q20b, d1, d2), ...);
}

// Per-warp staging buffers (shared memory), refilled every 32 steps ("chunk"):
//   inw  [32 rows][27] u32   the raw source bytes (aligned words) holding the 32 pixels each lane
//                            seeds its deepest window with during the chunk -- one coalesced
//                            word load per row
//   hin  [32 steps][6] f32   lane 0's two incoming streams for the chunk (from the band above,
//                            or raw rows 0/1 for the first band)
//   out  [32 rows][36] u8    the palette rows chosen during the chunk -> written back coalesced
//   hout [32 steps][6] f32   lane 31's two outgoing streams
struct WarpStage {
    unsigned inw[32][27];   // 108 raw source bytes per row: 96 wanted + alignment slack
    float hin[32][6];
    float hout[32][6];
    unsigned char out[32][36];
};

constexpr int WAVE_WARPS = WAVE_THREADS / 32;

template <int V>
__global__ void __launch_bounds__(WAVE_THREADS) k_diffuse_wave(const WaveParams p)
{
    using SP = Spec<V>;
    extern __shared__ __align__(16) unsigned char wave_smem[];
    double *s_pal = reinterpret_cast<double *>(wave_smem);                   // [256*3]
    uint2 *s_tab = reinterpret_cast<uint2 *>(s_pal + DP_MAX_COLORS * 3);      // [4096]
    float *s_palf = reinterpret_cast<float *>(s_tab + 4096);                  // [256*3]
    float *s_ow = s_palf + DP_MAX_COLORS * 3;                                 // [256*4]
    unsigned *s_orgb = reinterpret_cast<unsigned *>(s_ow + 256 * 4);          // [256]
    unsigned char *s_lut = reinterpret_cast<unsigned char *>(s_orgb + 256);   // [256]
    WarpStage *stages = reinterpret_cast<WarpStage *>(s_lut + 256);
    WarpStage &st = stages[threadIdx.x >> 5];
    const int NT = blockDim.x;

    const PalDev *P = p.P;
    for (int i = threadIdx.x; i < 4096; i += NT) s_tab[i] = P->ed_table[i];
    for (int i = threadIdx.x; i < p.K * 3; i += NT) {
        s_pal[i] = P->pal_f64[i];
        s_palf[i] = P->pal_f32[i];
    }
    for (int i = threadIdx.x; i < p.K; i += NT) {
        const uint8_t *o = P->out_rgb + 4 * i;
        s_orgb[i] = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16);
    }
    for (int i = threadIdx.x; i < 256; i += NT) s_lut[i] = P->in_lut[i];
    if (SP::OSTRO)
        for (int i = threadIdx.x; i < 256 * 4; i += NT) s_ow[i] = p.ostro_w[i];
    __syncthreads();

    Search srch;
    srch.table = s_tab;
    srch.ovf = P->ed_ovf;
    srch.s_pal = s_pal;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int W = p.w, H = p.h;
    const size_t frame_px = (size_t)W * H;
    const int T = W + SP::AMAX + SP::BMAX + SP::S * 31;
    const int NCH = (T + 31) >> 5;
    constexpr int DYF = SP::ROWS3 ? 2 : 1;             // row offset of the raw-pixel feed
    constexpr int BF = SP::ROWS3 ? SP::B2 : SP::B1;    // its column offset

    for (;;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(p.ticket, 1);
        unit = __shfl_sync(FULL, unit, 0);
        if (unit >= p.total_units) break;
        // tickets run band-major over the frames (band 0 of every frame, then band 1, ...): the
        // band above always holds an earlier ticket, and with many frames resident warps are
        // busy instead of waiting for their turn in one frame's wavefront
        const int band = unit / p.frames;
        const int f = unit - band * p.frames;
        unit = f * p.nbands + band;    // storage index of the hand-off streams
        const int y0 = band * 32;
        const int y = y0 + lane;
        const bool rowok = y < H;
        const bool has_next = (band + 1) < p.nbands;
        const uint8_t *src_f = p.src + frame_px * 3 * f;
        uint8_t *dst_f = p.dst + frame_px * 3 * f;
        uint8_t *idx_f = p.dst_idx ? p.dst_idx + frame_px * f : nullptr;
        const float *hin = p.hand + (size_t)(band > 0 ? unit - 1 : unit) * 2 * W * 3;
        float *hout = p.hand + (size_t)unit * 2 * W * 3;
        const int *prog_in = p.progress + (band > 0 ? unit - 1 : unit);
        int *prog_out = p.progress + unit;
        int avail = 0;

        float d1[SP::W1][3], d2[SP::W2][3];
        float fifoA[SP::DA > 1 ? SP::DA - 1 : 1][3], fifoB[SP::DB > 1 ? SP::DB - 1 : 1][3];
        float emitA[3] = {0.f, 0.f, 0.f}, emitB[3] = {0.f, 0.f, 0.f};
        double q10[3] = {0., 0., 0.}, q20a[3] = {0., 0., 0.}, q20b[3] = {0., 0., 0.};
        float oq10[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < SP::W1; ++j) d1[j][0] = d1[j][1] = d1[j][2] = 0.f;
#pragma unroll
        for (int j = 0; j < SP::W2; ++j) d2[j][0] = d2[j][1] = d2[j][2] = 0.f;
#pragma unroll
        for (int j = 0; j < (SP::DA > 1 ? SP::DA - 1 : 1); ++j) fifoA[j][0] = fifoA[j][1] = fifoA[j][2] = 0.f;
#pragma unroll
        for (int j = 0; j < (SP::DB > 1 ? SP::DB - 1 : 1); ++j) fifoB[j][0] = fifoB[j][1] = fifoB[j][2] = 0.f;

        // byte offsets are relative to p.src; reads are whole aligned words inside the batch
        const long long src_mis = (long long)(reinterpret_cast<uintptr_t>(p.src) & 3);
        const long long src_hi = ((long long)(frame_px * 3) * p.frames + src_mis + 3) & ~3ll;
        int my_o = 0;

#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
            const int t0 = ch << 5;
            const int x00 = t0 - SP::BMAX;        // lane 0's x at the first step of the chunk
            // this lane's row: byte offset (from p.src) of its first wanted column
            const long long gb0_lane = (long long)(frame_px * 3 * f) +
                                       ((long long)(y0 + lane + DYF) * W +
                                        (x00 - SP::S * lane + BF)) * 3 + src_mis;

            // ---- wait until the band above has published everything this chunk reads ------
            if (band > 0) {
                const int need = min(x00 + 31 + SP::S, W + SP::AMAX);
                if (need > avail) {
                    int v = 0;
                    for (unsigned spins = 0;; ++spins) {
                        if (lane == 0) v = ld_poll(prog_in);
                        v = __shfl_sync(FULL, v, 0);
                        if (v >= need) break;
                        __nanosleep(200);
                        if (spins > (1u << 24)) __trap();  // protocol bug -> error, not a hang
                    }
                    avail = v;
                }
            }

            // ---- stage the chunk's inputs ---------------------------------------------
            {
                // lane 0's streams: column of step j is x00 + j
                const int ca = x00 + lane, cb = ca + SP::B1;
                float ha[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (band == 0) {
                    if (ca >= 0 && ca < W) {
                        const uint8_t *q = src_f + ((size_t)y0 * W + ca) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[c] = (float)s_lut[q[c]];
                    }
                    if (SP::ROWS3 && y0 + 1 < H && cb >= 0 && cb < W) {
                        const uint8_t *q = src_f + ((size_t)(y0 + 1) * W + cb) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[3 + c] = (float)s_lut[q[c]];
                    }
                } else {
                    if (ca >= 0 && ca < W) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[c] = __ldcg(hin + 3 * ca + c);
                    }
                    if (SP::ROWS3 && cb >= 0 && cb < W) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[3 + c] = __ldcg(hin + 3 * (W + cb) + c);
                    }
                }
#pragma unroll
                for (int c = 0; c < 6; ++c) st.hin[lane][c] = ha[c];
                // raw-pixel feed: for row r the 96 bytes of columns col0_r .. col0_r+31, fetched
                // as the (up to 25) aligned words that contain them; bytes of columns outside the
                // image belong to targets that do not exist and are never used
                {
                    const long long row_step = (long long)W * 3 - 3 * SP::S;
                    long long gb = (long long)(frame_px * 3 * f) +
                                   ((long long)(y0 + DYF) * W + (x00 + BF)) * 3 + src_mis;
#pragma unroll 4
                    for (int r = 0; r < 32; ++r, gb += row_step) {
                        const long long wa = (gb & ~3ll) + 4 * lane;
                        unsigned v = 0;
                        if (lane < 25 && wa >= 0 && wa + 4 <= src_hi)
                            v = __ldg(reinterpret_cast<const unsigned *>(p.src - src_mis + wa));
                        if (lane < 27) st.inw[r][lane] = v;
                    }
                    my_o = (int)((gb0_lane) & 3);
                }
            }
            __syncwarp();

            // ---- 32 pixel steps --------------------------------------------------------
#pragma unroll 1
            for (int sidx = 0; sidx < 32; ++sidx) {
                const int x = x00 + sidx - SP::S * lane;
                float fa[3], fb[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ra = __shfl_up_sync(FULL, emitA[c], 1);
                    const float rb = __shfl_up_sync(FULL, emitB[c], 1);
                    if (SP::DA > 1) {
                        fa[c] = fifoA[0][c];
#pragma unroll
                        for (int j = 0; j + 1 < SP::DA - 1; ++j) fifoA[j][c] = fifoA[j + 1][c];
                        fifoA[SP::DA - 2][c] = ra;
                    } else {
                        fa[c] = ra;
                    }
                    if (SP::DB > 1) {
                        fb[c] = fifoB[0][c];
#pragma unroll
                        for (int j = 0; j + 1 < SP::DB - 1; ++j) fifoB[j][c] = fifoB[j + 1][c];
                        fifoB[SP::DB - 2][c] = rb;
                    } else {
                        fb[c] = rb;
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        fa[c] = st.hin[sidx][c];
                        fb[c] = st.hin[sidx][3 + c];
                    }
                }
                {
                    const unsigned char *pb =
                        reinterpret_cast<const unsigned char *>(st.inw[lane]) + my_o + 3 * sidx;
                    unsigned b0 = pb[0], b1 = pb[1], b2 = pb[2];
                    if (p.has_lut) {
                        b0 = s_lut[b0];
                        b1 = s_lut[b1];
                        b2 = s_lut[b2];
                    }
                    float *top = SP::ROWS3 ? d2[SP::W2 - 1] : d1[SP::W1 - 1];
                    if (SP::ROWS3) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) d1[SP::W1 - 1][c] = fb[c];
                    }
                    top[0] = (float)b0;
                    top[1] = (float)b1;
                    top[2] = (float)b2;
                }

                const bool active = rowok && x >= 0 && x < W;
                if (active) {
                    int bi;
                    if (!SP::OSTRO) {
                        double v[3], e[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float a = fa[c];
                            if (SP::H20) a = acc_f64(a, q20a[c]);
                            if (SP::H10) a = acc_f64(a, q10[c]);
                            a = fminf(fmaxf(a, 0.f), 255.f);
                            v[c] = (double)a;
                        }
                        bi = nearest_first(srch, v[0], v[1], v[2]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) e[c] = __dsub_rn(v[c], (double)s_palf[3 * bi + c]);
                        apply_taps<V>(std::make_integer_sequence<int, SP::N>{}, e, q10, q20a, q20b,
                                      d1, d2);
                    } else {
                        float ov[3], er[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float a = __fadd_rn(fa[c], oq10[c]);
                            ov[c] = fminf(fmaxf(a, 0.f), 255.f);
                        }
                        bi = nearest_kd(P, srch, (double)ov[0], (double)ov[1], (double)ov[2]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) er[c] = __fsub_rn(ov[c], s_palf[3 * bi + c]);
                        float lum = __fmul_rn(0.299f, ov[0]);
                        lum = __fadd_rn(lum, __fmul_rn(0.587f, ov[1]));
                        lum = __fadd_rn(lum, __fmul_rn(0.114f, ov[2]));
                        lum = fminf(fmaxf(lum, 0.f), 255.f);
                        const int li = (int)lum;
                        const float w0 = s_ow[4 * li], w1 = s_ow[4 * li + 1], w2 = s_ow[4 * li + 2];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            oq10[c] = __fmul_rn(er[c], w0);
                            d1[1][c] = __fadd_rn(d1[1][c], __fmul_rn(er[c], w2));
                            d1[0][c] = __fadd_rn(d1[0][c], __fmul_rn(er[c], w1));
                        }
                    }
                    st.out[lane][sidx] = (unsigned char)bi;
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        q10[c] = 0.0;
                        q20a[c] = q20b[c];
                        q20b[c] = 0.0;
                        oq10[c] = 0.f;
                    }
                }

#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    emitA[c] = d1[0][c];
                    emitB[c] = SP::ROWS3 ? d2[0][c] : 0.f;
                }
                if (lane == 31) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        st.hout[sidx][c] = emitA[c];
                        st.hout[sidx][3 + c] = emitB[c];
                    }
                }
#pragma unroll
                for (int j = 0; j + 1 < SP::W1; ++j) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) d1[j][c] = d1[j + 1][c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) d1[SP::W1 - 1][c] = 0.f;
                if (SP::ROWS3) {
#pragma unroll
                    for (int j = 0; j + 1 < SP::W2; ++j) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) d2[j][c] = d2[j + 1][c];
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) d2[SP::W2 - 1][c] = 0.f;
                }
            }
            __syncwarp();

            // ---- write the chunk back ----------------------------------------------------
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
                const int yr = y0 + r;
                const int col = x00 + lane - SP::S * r;
                if (yr < H && col >= 0 && col < W) {
                    const unsigned bi = st.out[r][lane];
                    const unsigned oc = s_orgb[bi];
                    uint8_t *o = dst_f + ((size_t)yr * W + col) * 3;
                    o[0] = (uint8_t)oc;
                    o[1] = (uint8_t)(oc >> 8);
                    o[2] = (uint8_t)(oc >> 16);
                    if (idx_f) idx_f[(size_t)yr * W + col] = (uint8_t)bi;
                }
            }
            if (has_next) {
                // lane 31's x at step j of this chunk
                const int x31 = x00 + lane - SP::S * 31;
                const int xa = x31 - SP::A1;
                if (xa >= 0 && xa < W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) __stcg(hout + 3 * xa + c, st.hout[lane][c]);
                }
                if (SP::ROWS3) {
                    const int xb = x31 - SP::A2;
                    if (xb >= 0 && xb < W) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) __stcg(hout + 3 * (W + xb) + c, st.hout[lane][3 + c]);
                    }
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    const int prog = (ch == NCH - 1) ? 0x3fffffff : (x00 + 32 - SP::S * 31);
                    st_release(prog_out, prog);
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------
// Serpentine (or any) scan, serial per frame: one warp per frame, lane 0 walks the reference's
// loop on a 3-row f32 ring kept in global memory (L1/L2 resident); the other lanes only help
// moving rows in and out.  Correct for every variant; throughput comes from frames in flight.
// ---------------------------------------------------------------------------------------
struct SerialParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, K, serpentine, variant;
    float *ring;           // [frames][3][w][3]
    const float *ostro_w;  // ostromoukhov only
};

template <bool OSTRO>
__global__ void __launch_bounds__(32) k_diffuse_serial(const SerialParams p)
{
    __shared__ double s_pal[DP_MAX_COLORS * 3];
    __shared__ float s_palf[DP_MAX_COLORS * 3];
    __shared__ uint8_t s_orgb[DP_MAX_COLORS * 4];
    __shared__ float s_lutf[256];
    __shared__ int s_tdx[12], s_tdy[12];
    __shared__ double s_tw[12];
    const PalDev *P = p.P;
    const int lane = threadIdx.x;
    for (int i = lane; i < p.K * 3; i += 32) {
        s_pal[i] = P->pal_f64[i];
        s_palf[i] = P->pal_f32[i];
    }
    for (int i = lane; i < p.K * 4; i += 32) s_orgb[i] = P->out_rgb[i];
    for (int i = lane; i < 256; i += 32) s_lutf[i] = (float)P->in_lut[i];
    int ntaps = 0;
    if (!OSTRO) {
        // runtime copy of the constexpr tables (this kernel is not specialised per variant)
        ntaps = ed_ntaps(p.variant);
        if (lane < ntaps) {
            Tap tp = ed_tap(p.variant, lane);
            s_tdx[lane] = tp.dx;
            s_tdy[lane] = tp.dy;
            s_tw[lane] = (double)(float)tp.w / (double)ed_divisor(p.variant);
        }
    }
    __syncthreads();
    Search srch;
    srch.table = P->ed_table;
    srch.ovf = P->ed_ovf;
    srch.s_pal = s_pal;

    const int W = p.w, H = p.h;
    const size_t frame_px = (size_t)W * H;
    for (int f = blockIdx.x; f < p.frames; f += gridDim.x) {
        const uint8_t *src_f = p.src + frame_px * 3 * f;
        uint8_t *dst_f = p.dst + frame_px * 3 * f;
        uint8_t *idx_f = p.dst_idx ? p.dst_idx + frame_px * f : nullptr;
        float *ring = p.ring + (size_t)f * 3 * W * 3;
        // rows 0..2 into the ring
        for (int rr = 0; rr < 3 && rr < H; ++rr)
            for (int i = lane; i < W * 3; i += 32)
                ring[(size_t)rr * W * 3 + i] = s_lutf[src_f[(size_t)rr * W * 3 + i]];
        __syncwarp();
        for (int y = 0; y < H; ++y) {
            float *r0 = ring + (size_t)(y % 3) * W * 3;
            float *r1 = ring + (size_t)((y + 1) % 3) * W * 3;
            float *r2 = ring + (size_t)((y + 2) % 3) * W * 3;
            if (lane == 0) {
                const int dir = (p.serpentine && (y & 1)) ? -1 : 1;
                int x = dir > 0 ? 0 : W - 1;
                for (int n = 0; n < W; ++n, x += dir) {
                    float *px = r0 + 3 * x;
                    int bi;
                    if (!OSTRO) {
                        double v[3], e[3];
                        for (int c = 0; c < 3; ++c) {
                            double tv = (double)px[c];
                            v[c] = tv < 0.0 ? 0.0 : (tv > 255.0 ? 255.0 : tv);
                        }
                        bi = nearest_first(srch, v[0], v[1], v[2]);
                        for (int c = 0; c < 3; ++c) e[c] = __dsub_rn(v[c], (double)s_palf[3 * bi + c]);
                        for (int k = 0; k < ntaps; ++k) {
                            const int nx = x + s_tdx[k] * dir;
                            const int dy = s_tdy[k];
                            if (nx < 0 || nx >= W || y + dy >= H) continue;
                            float *q = (dy == 0 ? r0 : dy == 1 ? r1 : r2) + 3 * nx;
                            for (int c = 0; c < 3; ++c) q[c] = acc_f64(q[c], __dmul_rn(e[c], s_tw[k]));
                        }
                    } else {
                        float ov[3], er[3];
                        for (int c = 0; c < 3; ++c) {
                            float a = px[c];
                            ov[c] = a < 0.f ? 0.f : (a > 255.f ? 255.f : a);
                        }
                        bi = nearest_kd(P, srch, (double)ov[0], (double)ov[1], (double)ov[2]);
                        for (int c = 0; c < 3; ++c) er[c] = __fsub_rn(ov[c], s_palf[3 * bi + c]);
                        float lum = __fmul_rn(0.299f, ov[0]);
                        lum = __fadd_rn(lum, __fmul_rn(0.587f, ov[1]));
                        lum = __fadd_rn(lum, __fmul_rn(0.114f, ov[2]));
                        lum = lum < 0.f ? 0.f : (lum > 255.f ? 255.f : lum);
                        const int li = (int)lum;
                        const float w0 = p.ostro_w[4 * li], w1 = p.ostro_w[4 * li + 1],
                                    w2 = p.ostro_w[4 * li + 2];
                        int nx = x + dir;
                        if (nx >= 0 && nx < W)
                            for (int c = 0; c < 3; ++c)
                                r0[3 * nx + c] = __fadd_rn(r0[3 * nx + c], __fmul_rn(er[c], w0));
                        if (y + 1 < H) {
                            nx = x - dir;
                            if (nx >= 0 && nx < W)
                                for (int c = 0; c < 3; ++c)
                                    r1[3 * nx + c] = __fadd_rn(r1[3 * nx + c], __fmul_rn(er[c], w1));
                            for (int c = 0; c < 3; ++c)
                                r1[3 * x + c] = __fadd_rn(r1[3 * x + c], __fmul_rn(er[c], w2));
                        }
                    }
                    uint8_t *o = dst_f + ((size_t)y * W + x) * 3;
                    o[0] = s_orgb[4 * bi];
                    o[1] = s_orgb[4 * bi + 1];
                    o[2] = s_orgb[4 * bi + 2];
                    if (idx_f) idx_f[(size_t)y * W + x] = (uint8_t)bi;
                }
            }
            __syncwarp();
            // row y's slot becomes row y+3
            if (y + 3 < H)
                for (int i = lane; i < W * 3; i += 32)
                    r0[i] = s_lutf[src_f[(size_t)(y + 3) * W * 3 + i]];
            __syncwarp();
        }
    }
}

struct Workspace {
    void *ptr = nullptr;
    cudaStream_t st;
    int alloc(size_t bytes, cudaStream_t s)
    {
        st = s;
        DP_CUDA(cudaMallocAsync(&ptr, bytes ? bytes : 1, s));
        return 0;
    }
    ~Workspace()
    {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

template <int V>
int launch_wave(const WaveParams &p, cudaStream_t st)
{
    // Few bands (a single image, a small batch): 4-warp blocks so that the bands spread over
    // the SM sub-partitions (the kernel is latency-bound per warp).  Many bands: 8-warp blocks,
    // which share the tables and reach the occupancy limit set by registers.
    const int sms = dp_num_sms();
    const int warps = (p.total_units <= sms * 12) ? 4 : WAVE_WARPS;
    const size_t smem = DP_MAX_COLORS * 3 * 8 + 4096 * 8 + DP_MAX_COLORS * 3 * 4 + 256 * 4 * 4 +
                        256 * 4 + 256 + sizeof(WarpStage) * warps;
    DP_CUDA(cudaFuncSetAttribute(k_diffuse_wave<V>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    int per_sm = 0;
    DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_diffuse_wave<V>, warps * 32,
                                                          smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = ((long long)p.total_units + warps - 1) / warps;
    long long cap = (long long)sms * per_sm;
    int grid = (int)(blocks < cap ? blocks : cap);
    k_diffuse_wave<V><<<grid, warps * 32, smem, st>>>(p);
    DP_LAUNCH_CHECK();
    return 0;
}

int run_diffusion(const dp_palette *pal, const uint8_t *src, int frames, int h, int w, int variant,
                  int serpentine, const float *ostro_w_host, uint8_t *dst, uint8_t *dst_idx,
                  cudaStream_t st)
{
    const bool ostro = (variant == V_OSTRO);
    Workspace ws_ow;
    const float *ostro_w = nullptr;
    if (ostro) {
        if (ws_ow.alloc(256 * 4 * sizeof(float), st)) return 1;
        DP_CUDA(cudaMemcpyAsync(ws_ow.ptr, ostro_w_host, 256 * 4 * sizeof(float),
                                cudaMemcpyHostToDevice, st));
        // the host array is a temporary of the caller: make the copy complete before returning
        DP_CUDA(cudaStreamSynchronize(st));
        ostro_w = static_cast<const float *>(ws_ow.ptr);
    }
    if (serpentine || h < 2) {
        SerialParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.P = reinterpret_cast<const PalDev *>(pal->blob);
        sp.src = src;
        sp.dst = dst;
        sp.dst_idx = dst_idx;
        sp.frames = frames;
        sp.h = h;
        sp.w = w;
        sp.K = pal->dev.K;
        sp.serpentine = serpentine;
        sp.variant = variant;
        sp.ostro_w = ostro_w;
        Workspace ring;
        if (ring.alloc((size_t)frames * 3 * w * 3 * sizeof(float), st)) return 1;
        sp.ring = static_cast<float *>(ring.ptr);
        int grid = frames < dp_num_sms() * 16 ? frames : dp_num_sms() * 16;
        if (ostro)
            k_diffuse_serial<true><<<grid, 32, 0, st>>>(sp);
        else
            k_diffuse_serial<false><<<grid, 32, 0, st>>>(sp);
        DP_LAUNCH_CHECK();
        return 0;
    }
    WaveParams p;
    memset(&p, 0, sizeof(p));
    p.P = reinterpret_cast<const PalDev *>(pal->blob);
    p.src = src;
    p.dst = dst;
    p.dst_idx = dst_idx;
    p.frames = frames;
    p.h = h;
    p.w = w;
    p.nbands = (h + 31) / 32;
    long long units = (long long)p.nbands * frames;
    DP_REQUIRE(units < (1ll << 30), "too many row bands in one call");
    p.total_units = (int)units;
    p.has_lut = pal->has_lut;
    p.K = pal->dev.K;
    p.ostro_w = ostro_w;
    Workspace hand, flags;
    if (hand.alloc((size_t)units * 2 * w * 3 * sizeof(float), st)) return 1;
    if (flags.alloc((size_t)(units + 1) * sizeof(int), st)) return 1;
    DP_CUDA(cudaMemsetAsync(flags.ptr, 0, (size_t)(units + 1) * sizeof(int), st));
    p.hand = static_cast<float *>(hand.ptr);
    p.progress = static_cast<int *>(flags.ptr);
    p.ticket = p.progress + units;
    switch (variant) {
        case DP_ED_FLOYD_STEINBERG: return launch_wave<DP_ED_FLOYD_STEINBERG>(p, st);
        case DP_ED_JJN: return launch_wave<DP_ED_JJN>(p, st);
        case DP_ED_STUCKI: return launch_wave<DP_ED_STUCKI>(p, st);
        case DP_ED_BURKES: return launch_wave<DP_ED_BURKES>(p, st);
        case DP_ED_ATKINSON: return launch_wave<DP_ED_ATKINSON>(p, st);
        case DP_ED_SIERRA: return launch_wave<DP_ED_SIERRA>(p, st);
        case DP_ED_SIERRA_TWO_ROW: return launch_wave<DP_ED_SIERRA_TWO_ROW>(p, st);
        case DP_ED_SIERRA_LITE: return launch_wave<DP_ED_SIERRA_LITE>(p, st);
        default: return launch_wave<V_OSTRO>(p, st);
    }
}

}  // namespace

extern "C" int dp_error_diffusion(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h,
                                  int w, int variant, int serpentine, uint8_t *dst_rgb,
                                  uint8_t *dst_idx, void *stream)
{
    DP_REQUIRE(pal && src_rgb && dst_rgb, "null argument");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    DP_REQUIRE(variant >= DP_ED_FLOYD_STEINBERG && variant <= DP_ED_SIERRA_LITE,
               "unknown error-diffusion variant");
    if (frames == 0 || h == 0 || w == 0) return 0;
    return run_diffusion(pal, src_rgb, frames, h, w, variant, serpentine ? 1 : 0, nullptr, dst_rgb,
                         dst_idx, dp_stream(stream));
}

extern "C" int dp_ostromoukhov(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h,
                               int w, const int32_t *coeffs, int serpentine, uint8_t *dst_rgb,
                               uint8_t *dst_idx, void *stream)
{
    DP_REQUIRE(pal && src_rgb && dst_rgb && coeffs, "null argument");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    if (frames == 0 || h == 0 || w == 0) return 0;
    // weights as the reference forms them: float32(c / (c0+c1+c2)), c/d in f64 (:1252-1266);
    // a zero divisor skips the distribution (:1254-1255) == all-zero weights
    float wts[256 * 4];
    for (int i = 0; i < 256; ++i) {
        int c0 = coeffs[3 * i], c1 = coeffs[3 * i + 1], c2 = coeffs[3 * i + 2];
        int d = c0 + c1 + c2;
        wts[4 * i + 0] = d ? (float)((double)c0 / (double)d) : 0.f;
        wts[4 * i + 1] = d ? (float)((double)c1 / (double)d) : 0.f;
        wts[4 * i + 2] = d ? (float)((double)c2 / (double)d) : 0.f;
        wts[4 * i + 3] = 0.f;
    }
    return run_diffusion(pal, src_rgb, frames, h, w, V_OSTRO, serpentine ? 1 : 0, wts, dst_rgb,
                         dst_idx, dp_stream(stream));
}
