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
// Bands are handed out through a READY QUEUE in global memory: band 0 of every frame is seeded,
// and a band enqueues its successor once it is far enough ahead for the successor to start.  A
// resident warp therefore always holds runnable work (with many frames in flight every warp
// slot of the GPU is busy), and a band's predecessor is always already running -- no deadlock
// whatever the grid size.
//
// Arithmetic is the reference's: numba path = f32 state, f64 math, strict '<' first-index
// nearest colour; Ostromoukhov = f32 math, f32-rounded weights, KD-tree nearest.
// Serpentine scanning makes row y+1 depend on the END of row y: it is serial per frame and
// handled by a one-warp-per-frame kernel (parallel over frames only).
//
// Algorithmic bytes: 3 read + 3 written per pixel; the bound is the dependency chain
// (W + S*(H-1) pixel steps per frame) and fp64 issue, not HBM.
#include <stdlib.h>

#include <mutex>
#include <utility>

#include "dp_search.cuh"

namespace {

constexpr int V_OSTRO = 8;
constexpr int V_HYBRID = 9;   // Floyd-Steinberg footprint, luminance/colour split of the error
constexpr int V_WEIGHTED = 10; // Floyd-Steinberg footprint, all-f32, per-pixel weight factor, KD-tree nearest
                               // on the UNCLAMPED value (perceptual: dithering_lib.py:1030-1066)

// footprint a variant diffuses with
__host__ __device__ constexpr int ed_base(int v)
{
    return (v == V_HYBRID || v == V_WEIGHTED) ? DP_ED_FLOYD_STEINBERG : v;
}

struct Tap {
    int dx, dy, w;
};

__host__ __device__ constexpr int ed_ntaps(int v0)
{
    const int v = ed_base(v0);
    return v == DP_ED_FLOYD_STEINBERG ? 4 : v == DP_ED_JJN ? 12 : v == DP_ED_STUCKI ? 12
         : v == DP_ED_BURKES ? 7 : v == DP_ED_ATKINSON ? 6 : v == DP_ED_SIERRA ? 10
         : v == DP_ED_SIERRA_TWO_ROW ? 7 : v == DP_ED_SIERRA_LITE ? 3 : /*ostro*/ 3;
}

__host__ __device__ constexpr int ed_divisor(int v0)
{
    const int v = ed_base(v0);
    return v == DP_ED_FLOYD_STEINBERG ? 16 : v == DP_ED_JJN ? 48 : v == DP_ED_STUCKI ? 42
         : v == DP_ED_BURKES ? 32 : v == DP_ED_ATKINSON ? 8 : v == DP_ED_SIERRA ? 32
         : v == DP_ED_SIERRA_TWO_ROW ? 16 : v == DP_ED_SIERRA_LITE ? 4 : 1;
}

// dithering_lib.py:107-188, in the reference's order
__host__ __device__ constexpr Tap ed_tap(int v0, int k)
{
    const int v = ed_base(v0);
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
    static constexpr bool WEIGHTED = (V == V_WEIGHTED);
    static constexpr bool F32 = OSTRO || WEIGHTED;          // all-f32 state and arithmetic
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

#ifdef DP_WAVE_TIMING
__device__ unsigned long long g_wave_timing[128 * 4 + 4];   // + counters: slow pixels, slow warp-steps
#define DP_TICK(k)                                                     \
    do {                                                               \
        const long long _now = clock64();                              \
        if (lane == 0 && f == 0 && band < 128)                         \
            g_wave_timing[band * 4 + (k)] += (unsigned long long)(_now - _tick); \
        _tick = _now;                                                  \
    } while (0)
__device__ unsigned long long g_step_timing[16];
#define DP_STICK(k)                                                    \
    do {                                                               \
        const long long _snow = clock64();                             \
        if (lane == 5 && f == 0 && band == 10)                         \
            g_step_timing[(k)] += (unsigned long long)(_snow - _stick); \
        _stick = _snow;                                                \
    } while (0)
#else
#define DP_TICK(k) do { } while (0)
#define DP_STICK(k) do { } while (0)
#endif

struct WaveParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, nbands, total_units, has_lut, K;
    int bytes;          // identity input LUT and palette rows == output bytes: BYTES instantiation
    float *hand;        // [units][6 planes][TPAD] f32 hand-off streams (index = consumer step)
    int *progress;      // [units]
    int *qctrl;         // [0] queue head (consumers), [1] queue tail (producers)
    int *queue;         // [units] unit + 1, 0 = not yet enqueued
    int discard;        // hand-off stream lines are 128-byte aligned: consumers may discard them from L2
    int slack;          // extra chunks of lead before the successor band is enqueued
    int pat_smem;       // number of candidate patterns kept in shared memory (all or none)
    int l1_smem;        // first table level in shared memory (else read through L1)
    const float *ostro_w;  // [256][4] f32 weights (c0,c1,c2)/sum, DEVICE (ostromoukhov only)
    double hyb_lum, hyb_col;   // lum_factor, col_factor (hybrid only)
    const float *plane;        // [frames][h][w] per-pixel weight factor, DEVICE (weighted only)
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
// Spin-wait guard.  The protocol cannot deadlock (a band's predecessor is always running), so a
// wait is bounded by the predecessor's progress; the guard only turns a protocol BUG into an
// error instead of a hang.  Every spin sleeps >= ~100 ns, so 2^30 spins are at least 100 s of
// waiting (far more under compute-sanitizer or a debugger, where each spin is slower): a slow
// predecessor is not mistaken for a bug.
#define DP_SPIN_LIMIT (1u << 30)
// Drop one 128-byte line from L2 without writing it back (the address must be 128-byte aligned).
__device__ __forceinline__ void discard_l2_line(const void *p)
{
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}
// Executed by EVERY lane once a poll has succeeded (the polled value reaches the lanes by
// shuffle): orders this thread's later loads of the producer's data after the observation of the
// flag -- the acquire side of the producer's fence + st.release, per the PTX memory model
// (fence-fence synchronisation through the relaxed load that observed the release).
__device__ __forceinline__ void acquire_fence()
{
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
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

// The wavefront kernel keeps the f32 work values of the numba path as DOUBLES that are exactly
// representable in f32, because f32<->f64 conversions issue at 1/8 rate on sm_100 (measured:
// 8.7 cycles per warp instruction, 18 cycles latency; tools/ubench/lat.cu) and the reference's
// `+=` needs two of them per tap and channel.  Rounding an f64 to the nearest f32 (ties to even)
// without a conversion, in two fused multiply-adds (Veltkamp's splitting with the constant
// 2^29 - 1):   t = fl(x * 2^29 - x),   r = x * 2^29 - t.
// x * 2^29 is exact, t is x * (2^29 - 1) rounded to 53 bits, i.e. to a multiple of the f32 ulp of
// x's binade (2^(e-23)), so r = x * 2^29 - t -- exact by Sterbenz -- is x rounded to that grid;
// the tie rule is the adder's round-to-even (in a tie the low 29 bits of x are 10...0, so
// x * 2^29 is an even multiple of the grid and the parity of t is the parity of r).  With the
// MINUS sign the sum can only fall into the binade below (for mantissas within 2^-29 of 1.0), where
// the finer grid still rounds such x to 2^e; the plus sign would spill into the binade above and
// round too coarsely.  Checked against (double)(float)x on 8e7 adversarial doubles (ties, ties +- 1
// ulp, mantissas next to 1 and 2, both signs): identical.  Valid for 0 and for magnitudes in the
// normal f32 range (work values are bounded by a few hundred; |x| < 2^-126 would need three
// consecutive exact cancellations).  The sign of a zero result may differ (+0 for -0), which no
// comparison, clamp or product downstream can observe in the chosen palette rows.
__device__ __forceinline__ double round_to_f32(double x)
{
    const double t = __fma_rn(x, 536870912.0, -x);
    return __fma_rn(x, 536870912.0, -t);
}
__device__ __forceinline__ double acc_r(double acc, double prod)
{
    return round_to_f32(__dadd_rn(acc, prod));
}
// clamp to [0,255] (the numba loop's `if r < 0 ... elif r > 255`, :247-252)
__device__ __forceinline__ double clamp255(double a)
{
    a = (__double2hiint(a) < 0) ? 0.0 : a;
    return (a > 255.0) ? 255.0 : a;
}
// f32 bit pattern of a double in [0,255] that is exactly representable in f32 (integer ops only)
__device__ __forceinline__ float narrow_nonneg(double v)
{
    const int hi = __double2hiint(v);
    const unsigned b = __funnelshift_l((unsigned)__double2loint(v), (unsigned)(hi - 0x38000000), 3);
    return __int_as_float(hi < 0x38100000 ? 0 : (int)b);
}

// Loads from the read-only shared tables through 32-bit shared-space addresses: with generic
// pointers the compiler re-derives the shared window base (S2R SR_CgaCtaId) inside the pixel
// loop, on the critical path of every step.
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds_f32x4(unsigned a)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned a)
{
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

__device__ __forceinline__ unsigned lds_u16(unsigned a)
{
    unsigned short v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(unsigned a)
{
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// Byte -> work value without a table (BYTES instantiations: identity input LUT, palette rows equal
// to their output bytes).  The shared-memory pipe is the busiest unit of the wavefront kernel
// (ncu: l1tex data-pipe wavefronts 68-74 % of peak, profiles/r2p_*): the byte -> double LUT and
// the f64 palette rows cost ~21 of its ~99 wavefronts per warp step, all of them bank-conflicted
// gathers.  2^52 + b is exact and so is the subtraction (one DADD / FADD on an idle pipe).
__device__ __forceinline__ double byte_f64(unsigned b)
{
    return __dsub_rn(__hiloint2double(0x43300000, (int)b), 4503599627370496.0);
}
__device__ __forceinline__ float byte_f32(unsigned b)
{
    return __fsub_rn(__int_as_float((int)(0x4b000000u | b)), 8388608.0f);
}

struct Search {
    const uint16_t *l1;      // global: cell -> pattern (format: dp_common.cuh)
    const uint4 *pat;        // global: candidate patterns
    const uint4 *flat;       // global: pattern of every cell (one load; saturating batches)
    unsigned l1_a;           // shared copy of l1 (wavefront kernel)
    unsigned pat_a;          // shared copy of the patterns, 0 if they stay in global memory (L1)
    const PalDev *P;         // overflow lists (rare)
    const double *s_pal;     // shared, [K,3]
    unsigned rows_a;         // shared address of float4 [257]: (r, g, b, row index bits); 256 = pad
};

__device__ __forceinline__ int cell_of(float r, float g, float b)
{
    // values are clamped to [0,255] by the caller; x + 2^23 rounded toward zero leaves trunc(x)
    // in the low mantissa bits (F2I is a quarter-rate op)
    const int ir = __float_as_int(__fadd_rz(r, 8388608.0f));
    const int ig = __float_as_int(__fadd_rz(g, 8388608.0f));
    const int ib = __float_as_int(__fadd_rz(b, 8388608.0f));
    return ((ir & 0xf8) << 7) | ((ig & 0xf8) << 2) | ((ib >> 3) & 0x1f);
}

// Screening pass of the nearest-row search, all in f32: the (up to seven) candidate rows of the
// cell are evaluated side by side, each distance is truncated to 16 mantissa bits and the row
// index put in the freed byte, and the two smallest keys are kept.  The f32 distance is within
// 5 * 2^-24 (relative) of the exact one and the truncation takes at most 2^-15 off, so a row
// whose key is more than 2048 units (>= 1.2e-4 relative) above the smallest cannot be the
// exact minimum: if the runner-up is that far away the smallest key's row IS the answer of the
// reference's f64 comparison loop.  Otherwise (near-tie, exact tie, or a cell with more than
// seven candidates) the caller takes the exact path below.
// NSLOT = 4: palettes whose cells (almost) never hold more than four candidates -- the usual
// 2..32-colour palettes -- evaluate only the first four slots; a cell with a fifth candidate is
// treated like a near-tie (exact path).
template <int NSLOT>
__device__ __forceinline__ int nearest_screen(const Search &s, const uint4 e, float r, float g,
                                              float b, bool &sure)
{
    const unsigned off[7] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16,
                             e.z & 0xffffu, e.z >> 16, e.w & 0xffffu};
    int key[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) {
        const float4 p = lds_f32x4(s.rows_a + off[j]);
        const float dr = __fsub_rn(r, p.x), dg = __fsub_rn(g, p.y), db = __fsub_rn(b, p.z);
        const float d = __fmaf_rn(db, db, __fmaf_rn(dg, dg, __fmul_rn(dr, dr)));
        key[j] = (__float_as_int(d) & (int)0xffffff00) | __float_as_int(p.w);
    }
    // the two smallest keys as a tree (the running pair costs three dependent operations per key):
    // pair the keys, then  smallest = min of the pair minima,  runner-up = min(second smallest of
    // the pair minima, smallest pair maximum) -- a pair maximum of another pair is never below
    // that pair's minimum, so taking all of them cannot undercut the true runner-up
    int k1, k2;
    if (NSLOT == 4) {
        const int lo01 = min(key[0], key[1]), hi01 = max(key[0], key[1]);
        const int lo23 = min(key[2], key[3]), hi23 = max(key[2], key[3]);
        k1 = min(lo01, lo23);
        k2 = min(max(lo01, lo23), min(hi01, hi23));
    } else {
        const int lo01 = min(key[0], key[1]), hi01 = max(key[0], key[1]);
        const int lo23 = min(key[2], key[3]), hi23 = max(key[2], key[3]);
        const int lo45 = min(key[4], key[5]), hi45 = max(key[4], key[5]);
        const int a = min(lo01, lo23), bq = max(lo01, lo23), c = min(lo45, key[6]), dq = max(lo45, key[6]);
        k1 = min(a, c);
        k2 = min(min(max(a, c), min(bq, dq)), min(min(hi01, hi23), hi45));
    }
    sure = ((e.w >> 16) != DP_ED_OVERFLOW) && (k2 - k1 > 2048);
    if (NSLOT < 7) sure = sure && (off[NSLOT] == DP_ED_PAD);   // slots fill in order
    return k1 & 255;
}

__device__ __forceinline__ double dist_numba(const double *pp, double r, double g, double b)
{
    const double dr = __dsub_rn(r, pp[0]), dg = __dsub_rn(g, pp[1]), db = __dsub_rn(b, pp[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dg, dg)), __dmul_rn(db, db));
}

// Candidate list of a cell for the exact paths: `n` rows, ascending; row j via cand_at().
struct CandList {
    uint4 e;
    const uint8_t *ovf;   // non-null: the cell has more than seven candidates
    int n;
};

__device__ __forceinline__ uint4 fetch_pattern(const Search &s, int cell)
{
    return __ldg(s.pat + __ldg(s.l1 + cell));
}

// `e`: the cell's pattern (the screening pass has it in registers already)
__device__ __forceinline__ CandList cand_list(const Search &s, int cell, const uint4 e)
{
    CandList L;
    L.e = e;
    L.ovf = nullptr;
    if ((L.e.w >> 16) == DP_ED_OVERFLOW) {
        const PalDev *P = s.P;
        int lo = 0, hi = P->ed_novf - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (P->ed_ovf_cells[mid] < cell) lo = mid + 1; else hi = mid;
        }
        L.ovf = P->ed_ovf + P->ed_ovf_off[lo];
        L.n = (int)(P->ed_ovf_off[lo + 1] - P->ed_ovf_off[lo]);
    } else {
        const unsigned w[4] = {L.e.x, L.e.y, L.e.z, L.e.w};
        int n = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const unsigned o = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu);
            n += (o != DP_ED_PAD) ? 1 : 0;
        }
        L.n = n;
    }
    return L;
}

__device__ __forceinline__ int cand_at(const CandList &L, int j)
{
    if (L.ovf) return (int)L.ovf[j];
    const unsigned w = j < 2 ? L.e.x : j < 4 ? L.e.y : j < 6 ? L.e.z : L.e.w;
    return (int)(((j & 1) ? (w >> 16) : (w & 0xffffu)) >> 4);
}

// numba path (:254-263), exact: strict '<' over f64 distances, first index wins.  The candidate
// list of the pixel's cell holds every row that can be nearest there, in ascending order.
__device__ __noinline__ int nearest_first_exact(const Search &s, int cell, const uint4 e, double r, double g,
                                                double b)
{
    const CandList L = cand_list(s, cell, e);
    double best = 1e20;
    int bi = 0;
    for (int j = 0; j < L.n; ++j) {
        const int i = cand_at(L, j);
        const double d = dist_numba(s.s_pal + 3 * i, r, g, b);
        if (d < best) {
            best = d;
            bi = i;
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

// KD-tree nearest (:1243), exact: unique minimum among the candidates, else replay scipy.
__device__ __noinline__ int nearest_kd_exact(const PalDev *P, const Search &s, int cell, const uint4 e,
                                             double r, double g, double b)
{
    const CandList L = cand_list(s, cell, e);
    double best = DP_INF_F64;
    int bi = 0;
    bool tie = false;
    for (int j = 0; j < L.n; ++j) {
        const int i = cand_at(L, j);
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

// KD-tree nearest of an arbitrary point (outside the colour cube the per-cell candidate lists do
// not apply): every row, unique minimum else scipy's traversal.
__device__ __noinline__ int nearest_kd_full(const PalDev *P, const double *s_pal, int K, double r, double g,
                                            double b)
{
    double best = DP_INF_F64;
    int bi = 0;
    bool tie = false;
    for (int i = 0; i < K; ++i) {
        const double d = dist_scipy(s_pal + 3 * i, r, g, b);
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

// work values (clamped to [0,255]) -> palette row, the reference's answer.  (r,g,b) are the f32
// screening copies of the exact values (xr,xg,xb).
// f32 screening of EVERY row (same keys and margin as nearest_screen): for cells whose candidate
// list overflows the pattern -- frequent only in the unbounded outer cells of the unclamped
// modes -- this beats walking the overflow list (a binary search through global memory).
__device__ __noinline__ int nearest_all_rows(const PalDev *P, const Search &s, int K, float r, float g, float b,
                                             double xr, double xg, double xb)
{
    int k1 = 0x7fffffff, k2 = 0x7fffffff;
    for (int i = 0; i < K; ++i) {
        const float4 p = lds_f32x4(s.rows_a + 16u * i);
        const float dr = __fsub_rn(r, p.x), dg = __fsub_rn(g, p.y), db = __fsub_rn(b, p.z);
        const float d = __fmaf_rn(db, db, __fmaf_rn(dg, dg, __fmul_rn(dr, dr)));
        const int key = (__float_as_int(d) & (int)0xffffff00) | __float_as_int(p.w);
        const int hi = max(k1, key);
        k1 = min(k1, key);
        k2 = min(k2, hi);
    }
    if (k2 - k1 > 2048) return k1 & 255;
    return nearest_kd_full(P, s.s_pal, K, xr, xg, xb);
}

template <bool KD, int NSLOT, bool ALLROWS = false>
__device__ __forceinline__ int nearest_row_in(const PalDev *P, const Search &s, int cell, float r, float g,
                                              float b, double xr, double xg, double xb, int K = 0)
{
    bool sure;
    uint4 e;
    if (s.l1_a) {
        const unsigned pi = lds_u16(s.l1_a + 2u * cell);
        e = s.pat_a ? lds_u32x4(s.pat_a + 16u * pi) : __ldg(s.pat + pi);
    } else {
        e = __ldg(s.flat + cell);
    }
    int bi = nearest_screen<NSLOT>(s, e, r, g, b, sure);
#ifdef DP_WAVE_TIMING
    if (!sure) atomicAdd(&g_wave_timing[128 * 4], 1ull);
    if (__any_sync(__activemask(), !sure) && (threadIdx.x & 31) == (__ffs(__activemask()) - 1))
        atomicAdd(&g_wave_timing[128 * 4 + 1], 1ull);
#endif
    if (!sure) {
        if (ALLROWS && (e.w >> 16) == DP_ED_OVERFLOW)
            bi = nearest_all_rows(P, s, K, r, g, b, xr, xg, xb);
        else
            bi = KD ? nearest_kd_exact(P, s, cell, e, xr, xg, xb) : nearest_first_exact(s, cell, e, xr, xg, xb);
    }
    return bi;
}

template <bool KD, int NSLOT>
__device__ __forceinline__ int nearest_row(const PalDev *P, const Search &s, float r, float g,
                                           float b, double xr, double xg, double xb)
{
    return nearest_row_in<KD, NSLOT>(P, s, cell_of(r, g, b), r, g, b, xr, xg, xb);
}

// One tap, everything about it known at compile time (weight = f64(f32(w)) / divisor, the
// reference's `weights[k] / divisor`, :280).
template <int V, int K, int W1, int W2>
__device__ __forceinline__ void apply_tap(const double (&e)[3], double (&q10)[3],
                                          double (&q20a)[3], double (&q20b)[3],
                                          double (&d1)[W1][3], double (&d2)[W2][3])
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
            d1[tp.dx + SP::A1][c] = acc_r(d1[tp.dx + SP::A1][c], pr);
        } else {
            d2[(tp.dy == 2 ? tp.dx + SP::A2 : 0)][c] =
                acc_r(d2[(tp.dy == 2 ? tp.dx + SP::A2 : 0)][c], pr);
        }
    }
}

template <int V, int... Ks, int W1, int W2>
__device__ __forceinline__ void apply_taps(std::integer_sequence<int, Ks...>, const double (&e)[3],
                                           double (&q10)[3], double (&q20a)[3],
                                           double (&q20b)[3], double (&d1)[W1][3],
                                           double (&d2)[W2][3])
{
    (apply_tap<V, Ks, W1, W2>(e, q10, q20a, q20b, d1, d2), ...);
}

// State type of the wavefront: doubles holding f32 values for the numba variants (see
// round_to_f32), plain f32 for Ostromoukhov (whose reference arithmetic is f32 throughout).
template <int V> struct StateOf { using T = double; };
template <> struct StateOf<V_OSTRO> { using T = float; };
template <> struct StateOf<V_WEIGHTED> { using T = float; };

// Per-warp staging buffers (shared memory), one "chunk" = 32 pixel steps:
//   inw  [2][32 rows][25] u32  raw source bytes (aligned words) holding the 32 pixels each lane
//                              seeds its deepest window with during a chunk; double-buffered,
//                              the next chunk's rows are in flight (cp.async) during this one
//   hin  [32 steps][3|6]       lane 0's two incoming streams for the chunk (from the band above,
//                              or raw rows 0/1 for the first band)
//   hout [32 steps][3|6]       lane 31's two outgoing streams
//   outb [32 rows][108] u8     the output bytes of the chunk, each row shifted by the (constant)
//                              misalignment m of its global address so that whole aligned words
//                              can be stored; word 24 (the trailing m bytes) is carried into
//                              word 0 of the next chunk
template <typename T, int NS>
struct WarpStageT {
    unsigned inw[2][32][25];       // 99 bytes per row are read at most
    T hin[32][NS];
    T hout[32][NS];
    unsigned char outb[32][108];
};

// Warps per block, from the registers ptxas needs without spilling the window state: 16 (128
// registers) for the 3x3 and two-row footprints, 12 (168) for Sierra and Ostromoukhov.  The 5x5
// footprints (JJN, Stucki) want 221 registers = 8 warps; capped at 168 (32 bytes of spills) they
// run 12 warps, which is 4 % faster on saturating batches and 9 % slower on a single image, so
// they are compiled both ways (BIG) and the launch picks.
template <int V, bool BIG>
constexpr int wave_max_warps()
{
    return (V == DP_ED_JJN || V == DP_ED_STUCKI) ? (BIG ? 12 : 8)
         : (V == DP_ED_SIERRA || V == V_OSTRO || V == V_WEIGHTED) ? 12 : 16;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 4 : 0;   // src-size 0: nothing is read, the word is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float to_f32(double v) { return __double2float_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }

#ifndef DP_WAVE_UNROLL_OTHER
#define DP_WAVE_UNROLL_OTHER 2
#endif
#ifndef DP_WAVE_UNROLL_5X5
#define DP_WAVE_UNROLL_5X5 2
#endif
__host__ __device__ constexpr int wave_step_unroll(int v)
{
    return (v == DP_ED_JJN || v == DP_ED_STUCKI) ? DP_WAVE_UNROLL_5X5 : DP_WAVE_UNROLL_OTHER;
}

template <int V, bool BIG, int NSLOT, bool BYTES>
__global__ void __launch_bounds__(wave_max_warps<V, BIG>() * 32, 1) k_diffuse_wave(const WaveParams p)
{
    using SP = Spec<V>;
    using T = typename StateOf<V>::T;
    using Stage = WarpStageT<T, SP::ROWS3 ? 6 : 3>;
    extern __shared__ __align__(16) unsigned char wave_smem[];
    double *s_pal = reinterpret_cast<double *>(wave_smem);                    // [256*3]
    float4 *s_rows = reinterpret_cast<float4 *>(s_pal + DP_MAX_COLORS * 3);   // [257]
    double *s_lutd = reinterpret_cast<double *>(s_rows + DP_MAX_COLORS + 1);  // [256] (as T)
    float *s_palf = reinterpret_cast<float *>(s_lutd + 256);                  // [256*3] (Ostromoukhov)
    float *s_ow = s_palf + (SP::F32 ? DP_MAX_COLORS * 3 : 0);                 // [256*4] (Ostromoukhov)
    unsigned *s_orgb = reinterpret_cast<unsigned *>(s_ow + (SP::OSTRO ? 256 * 4 : 0));   // [256]
    uint16_t *s_l1 = reinterpret_cast<uint16_t *>(s_orgb + 256);              // [32768] if l1_smem
    uint4 *s_pat = reinterpret_cast<uint4 *>(s_l1 + (p.l1_smem ? 32768 : 0)); // [p.pat_smem]
    Stage *stages = reinterpret_cast<Stage *>(s_pat + p.pat_smem);
    Stage &st = stages[threadIdx.x >> 5];
    T *s_lut = reinterpret_cast<T *>(s_lutd);   // source byte -> work value (gamma LUT folded in)
    const int NT = blockDim.x;

    const PalDev *P = p.P;
    {
        const uint4 *src4 = reinterpret_cast<const uint4 *>(P->ed_l1);
        uint4 *dst4 = reinterpret_cast<uint4 *>(s_l1);
        if (p.l1_smem)
            for (int i = threadIdx.x; i < 4096; i += NT) dst4[i] = __ldg(src4 + i);
        for (int i = threadIdx.x; i < p.pat_smem; i += NT) s_pat[i] = __ldg(P->ed_pat + i);
    }
    for (int i = threadIdx.x; i < p.K * 3; i += NT) {
        s_pal[i] = P->pal_f64[i];
        if (SP::F32) s_palf[i] = P->pal_f32[i];
    }
    for (int i = threadIdx.x; i < p.K; i += NT) {
        const uint8_t *o = P->out_rgb + 4 * i;
        s_orgb[i] = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16);
    }
    for (int i = threadIdx.x; i <= DP_MAX_COLORS; i += NT) {
        // rows past K (and the pad row 256) sit far outside the colour cube
        float4 rw = make_float4(1e18f, 1e18f, 1e18f, __int_as_float(0));
        if (i < p.K)
            rw = make_float4(P->pal_f32[3 * i], P->pal_f32[3 * i + 1], P->pal_f32[3 * i + 2],
                             __int_as_float(i));
        s_rows[i] = rw;
    }
    for (int i = threadIdx.x; i < 256; i += NT) s_lut[i] = (T)P->in_lut[i];
    if (SP::OSTRO)
        for (int i = threadIdx.x; i < 256 * 4; i += NT) s_ow[i] = p.ostro_w[i];
    __syncthreads();

    Search srch;
    srch.l1 = P->ed_l1;
    srch.pat = P->ed_pat;
    srch.flat = P->ed_flat;
    srch.l1_a = p.l1_smem ? smem_u32(s_l1) : 0u;
    srch.pat_a = p.pat_smem ? smem_u32(s_pat) : 0u;
    srch.P = P;
    srch.s_pal = s_pal;
    srch.rows_a = smem_u32(s_rows);
    const unsigned pal_a = smem_u32(s_pal), palf_a = smem_u32(s_palf), lut_a = smem_u32(s_lut),
                   orgb_a = smem_u32(s_orgb), ow_a = smem_u32(s_ow);

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int W = p.w, H = p.h;
    const size_t frame_px = (size_t)W * H;
    const int T_STEPS = W + SP::AMAX + SP::BMAX + SP::S * 31;
    const int NCH = (T_STEPS + 31) >> 5;
    const int TPAD = NCH << 5;                         // stream length (consumer step index)
    constexpr int DYF = SP::ROWS3 ? 2 : 1;             // row offset of the raw-pixel feed
    constexpr int BF = SP::ROWS3 ? SP::B2 : SP::B1;    // its column offset
    constexpr int NS = SP::ROWS3 ? 6 : 3;              // stream values per step
    constexpr int NFA = SP::DA > 1 ? SP::DA - 1 : 1, NFB = SP::DB > 1 ? SP::DB - 1 : 1;
    // byte offsets are relative to p.src; reads are whole aligned words inside the batch
    const long long src_mis = (long long)(reinterpret_cast<uintptr_t>(p.src) & 3);
    const long long src_hi = ((long long)(frame_px * 3) * p.frames + src_mis + 3) & ~3ll;
    const long long row_step = (long long)W * 3 - 3 * SP::S;

    for (;;) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(p.qctrl, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot >= p.total_units) break;
        int unit = 0;
        for (unsigned spins = 1;; ++spins) {
            if (lane == 0) unit = ld_poll(p.queue + slot);
            unit = __shfl_sync(FULL, unit, 0);
            if (unit > 0) break;
            __nanosleep(200);
            if (spins > DP_SPIN_LIMIT) __trap();
        }
        acquire_fence();
        __syncwarp();
        unit -= 1;                          // storage index of the hand-off streams: f * nbands + band
        const int f = unit / p.nbands;
        const int band = unit - f * p.nbands;
        const int y0 = band * 32;
        const int y = y0 + lane;
        const bool rowok = y < H;
        const bool has_next = (band + 1) < p.nbands;
        const uint8_t *src_f = p.src + frame_px * 3 * f;
        uint8_t *dst_f = p.dst + frame_px * 3 * f;
        uint8_t *idx_f = p.dst_idx ? p.dst_idx + frame_px * f : nullptr;
        const float *hin = p.hand + (size_t)(band > 0 ? unit - 1 : unit) * 6 * TPAD;
        float *hout = p.hand + (size_t)unit * 6 * TPAD;
        const int *prog_in = p.progress + (band > 0 ? unit - 1 : unit);
        int *prog_out = p.progress + unit;
        int avail = 0;

        T d1[SP::W1][3], d2[SP::W2][3];
        T fifoA[NFA][3], fifoB[NFB][3];
        T emitA[3] = {0, 0, 0}, emitB[3] = {0, 0, 0};
        double q10[3] = {0., 0., 0.}, q20a[3] = {0., 0., 0.}, q20b[3] = {0., 0., 0.};
        float oq10[3] = {0.f, 0.f, 0.f};
        float fac_nxt = 0.f;   // weighted variant: the factor of the NEXT pixel of this lane's row
#pragma unroll
        for (int j = 0; j < SP::W1; ++j) d1[j][0] = d1[j][1] = d1[j][2] = 0;
#pragma unroll
        for (int j = 0; j < SP::W2; ++j) d2[j][0] = d2[j][1] = d2[j][2] = 0;
#pragma unroll
        for (int j = 0; j < NFA; ++j) fifoA[j][0] = fifoA[j][1] = fifoA[j][2] = 0;
#pragma unroll
        for (int j = 0; j < NFB; ++j) fifoB[j][0] = fifoB[j][1] = fifoB[j][2] = 0;

        // this lane's rows: (a) the feed row y+DYF, first wanted byte at chunk 0 (relative to
        // p.src, advances 96 bytes per chunk); (b) the output row y, byte offset in the frame of
        // the pixel processed at step 0 of chunk 0 (may be negative), also +96 per chunk
        const long long feed0 = (long long)(frame_px * 3 * f) +
                                ((long long)(y + DYF) * W + (-SP::BMAX - SP::S * lane + BF)) * 3 +
                                src_mis;
        const int my_o = (int)(feed0 & 3);
        const long long out0 = ((long long)y * W + (-SP::BMAX - SP::S * lane)) * 3;
        const int my_m = (int)((reinterpret_cast<uintptr_t>(dst_f) + (unsigned long long)out0) & 3);
        const bool rgb_out = p.dst != nullptr;   // null: index-plane-only output (idx_f is set)

        // rows of chunk `c` -> st.inw[c & 1] (asynchronous; completion via cp_async_wait)
        auto issue_rows = [&](int c) {
            long long gb = (long long)(frame_px * 3 * f) +
                           ((long long)(y0 + DYF) * W + ((c << 5) - SP::BMAX + BF)) * 3 + src_mis;
            unsigned *dstw = &st.inw[c & 1][0][lane];
#pragma unroll 4
            for (int r = 0; r < 32; ++r, gb += row_step, dstw += 25) {
                const long long wa = (gb & ~3ll) + 4 * lane;
                const bool ok = wa >= 0 && wa + 4 <= src_hi;
                if (lane < 25) cp_async4(dstw, p.src - src_mis + (ok ? wa : 0), ok);
            }
            cp_async_commit();
        };
        issue_rows(0);

#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
#ifdef DP_WAVE_TIMING
            long long _tick = clock64();
#endif
            const int t0 = ch << 5;
            const int x00 = t0 - SP::BMAX;        // lane 0's x at the first step of the chunk
            if (ch + 1 < NCH)
                issue_rows(ch + 1);
            else
                cp_async_commit();                // keep one group per chunk in flight

            // ---- wait until the band above has published everything this chunk reads ------
            if (band > 0) {
                const int need = min(t0 + 32, W + SP::BMAX);
                if (need > avail) {
                    int v = 0;
                    for (unsigned spins = 1;; ++spins) {
                        if (lane == 0) v = ld_poll(prog_in);
                        v = __shfl_sync(FULL, v, 0);
                        if (v >= need) break;
                        __nanosleep(100);
                        if (spins > DP_SPIN_LIMIT) __trap();
                    }
                    acquire_fence();
                    __syncwarp();
                    avail = v;
                }
            }
            DP_TICK(0);

            // ---- lane 0's streams for the chunk: step j of the chunk is handled by lane j ----
            {
                const int ca = x00 + lane, cb = ca + SP::B1;
                const bool oka = ca >= 0 && ca < W;
                const bool okb = SP::ROWS3 && cb >= 0 && cb < W;
                T ha[6] = {0, 0, 0, 0, 0, 0};
                if (band == 0) {
                    if (oka) {
                        const uint8_t *q = src_f + ((size_t)y0 * W + ca) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[c] = s_lut[q[c]];
                    }
                    if (okb && y0 + 1 < H) {
                        const uint8_t *q = src_f + ((size_t)(y0 + 1) * W + cb) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) ha[3 + c] = s_lut[q[c]];
                    }
                } else {
                    float hv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int c = 0; c < NS; ++c) hv[c] = __ldcg(hin + (size_t)c * TPAD + t0 + lane);
                    // This chunk of the hand-off streams (one 128-byte line per plane) has now been
                    // consumed and is never read again: drop the dirty lines from L2 instead of
                    // letting them be written back to HBM (the streams are 12-25 % of the kernel's
                    // DRAM traffic otherwise).  The loads above must have completed first.
                    if (p.discard) {
                        float keep = hv[0];
#pragma unroll
                        for (int c = 1; c < NS; ++c) keep += hv[c];
                        if (__any_sync(FULL, keep == keep) && lane < NS)   // always true; orders the discard after the loads
                            discard_l2_line(hin + (size_t)lane * TPAD + t0);
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        ha[c] = oka ? (T)hv[c] : (T)0;
                        if (SP::ROWS3) ha[3 + c] = okb ? (T)hv[3 + c] : (T)0;
                    }
                }
#pragma unroll
                for (int c = 0; c < NS; ++c) st.hin[lane][c] = ha[c];
            }
            cp_async_wait<1>();   // this chunk's rows have landed (the next chunk's may still fly)
            __syncwarp();
            DP_TICK(1);

            // ---- 32 pixel steps --------------------------------------------------------
            const unsigned char *feed = reinterpret_cast<const unsigned char *>(st.inw[ch & 1][lane]) + my_o;
            unsigned char *ob = st.outb[lane] + my_m;
            // the 5x5 footprints shift two five-entry register windows every step (60 moves): two
            // steps per loop iteration let the compiler rename across the pair instead
#pragma unroll(wave_step_unroll(V))
            for (int sidx = 0; sidx < 32; ++sidx) {
                const int x = x00 + sidx - SP::S * lane;
#ifdef DP_WAVE_TIMING
                long long _stick = clock64();
#endif
                float fac = 0.f;
                if constexpr (SP::WEIGHTED) {
                    // one step ahead: the load is in flight during this step's search and taps
                    fac = fac_nxt;
                    const int xn = x + 1;
                    fac_nxt = (rowok && xn >= 0 && xn < W) ? __ldg(p.plane + ((size_t)f * H + y) * W + xn) : 0.f;
                }
                T fa[3], fb[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const T ra = __shfl_up_sync(FULL, emitA[c], 1);
                    if (SP::DA > 1) {
                        fa[c] = fifoA[0][c];
#pragma unroll
                        for (int j = 0; j + 1 < SP::DA - 1; ++j) fifoA[j][c] = fifoA[j + 1][c];
                        fifoA[NFA - 1][c] = ra;
                    } else {
                        fa[c] = ra;
                    }
                    if (SP::ROWS3) {
                        const T rb = __shfl_up_sync(FULL, emitB[c], 1);
                        if (SP::DB > 1) {
                            fb[c] = fifoB[0][c];
#pragma unroll
                            for (int j = 0; j + 1 < SP::DB - 1; ++j) fifoB[j][c] = fifoB[j + 1][c];
                            fifoB[NFB - 1][c] = rb;
                        } else {
                            fb[c] = rb;
                        }
                    } else {
                        fb[c] = 0;
                    }
                }
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        fa[c] = st.hin[sidx][c];
                        if (SP::ROWS3) fb[c] = st.hin[sidx][3 + c];
                    }
                }
                {
                    const unsigned char *pb = feed + 3 * sidx;
                    T *top = SP::ROWS3 ? d2[SP::W2 - 1] : d1[SP::W1 - 1];
                    if (SP::ROWS3) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) d1[SP::W1 - 1][c] = fb[c];
                    }
                    if constexpr (sizeof(T) == 8) {
                        top[0] = BYTES ? byte_f64(pb[0]) : lds_f64(lut_a + 8u * pb[0]);
                        top[1] = BYTES ? byte_f64(pb[1]) : lds_f64(lut_a + 8u * pb[1]);
                        top[2] = BYTES ? byte_f64(pb[2]) : lds_f64(lut_a + 8u * pb[2]);
                    } else {
                        top[0] = BYTES ? byte_f32(pb[0]) : lds_f32(lut_a + 4u * pb[0]);
                        top[1] = BYTES ? byte_f32(pb[1]) : lds_f32(lut_a + 4u * pb[1]);
                        top[2] = BYTES ? byte_f32(pb[2]) : lds_f32(lut_a + 4u * pb[2]);
                    }
                }

                DP_STICK(0);
                const bool active = rowok && x >= 0 && x < W;
                if (active) {
                    int bi;
                    unsigned oc;   // the row's output bytes (BYTES: also its palette values)
                    if constexpr (!SP::F32) {
                        double v[3], e[3];
                        float vf[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            double a = fa[c];
                            if (SP::H20) a = acc_r(a, q20a[c]);
                            if (SP::H10) a = acc_r(a, q10[c]);
                            v[c] = clamp255(a);
                            vf[c] = narrow_nonneg(v[c]);
                        }
                        DP_STICK(1);
                        bi = nearest_row<false, NSLOT>(P, srch, vf[0], vf[1], vf[2], v[0], v[1], v[2]);
                        DP_STICK(2);
                        oc = lds_u32(orgb_a + 4u * bi);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            e[c] = __dsub_rn(v[c], BYTES ? byte_f64(__byte_perm(oc, 0u, 0x4440u + c))
                                                         : lds_f64(pal_a + 24u * bi + 8u * c));
                        if constexpr (V == V_HYBRID) {
                            // _hybrid_numba :1447-1455, one rounding per operation, in its order
                            const double lum = __dadd_rn(
                                __dadd_rn(__dmul_rn(0.299, e[0]), __dmul_rn(0.587, e[1])),
                                __dmul_rn(0.114, e[2]));
                            const double l0 = __dmul_rn(0.299, lum), l1 = __dmul_rn(0.587, lum),
                                         l2 = __dmul_rn(0.114, lum);
                            e[0] = __dadd_rn(__dmul_rn(p.hyb_lum, l0), __dmul_rn(p.hyb_col, __dsub_rn(e[0], l0)));
                            e[1] = __dadd_rn(__dmul_rn(p.hyb_lum, l1), __dmul_rn(p.hyb_col, __dsub_rn(e[1], l1)));
                            e[2] = __dadd_rn(__dmul_rn(p.hyb_lum, l2), __dmul_rn(p.hyb_col, __dsub_rn(e[2], l2)));
                        }
                        apply_taps<V>(std::make_integer_sequence<int, SP::N>{}, e, q10, q20a, q20b,
                                      d1, d2);
                        DP_STICK(3);
                    } else if constexpr (SP::WEIGHTED) {
                        // perceptual (:1042-1063): no clamp, KD-tree nearest of the raw work value
                        // (NaN cannot arise: the work values stay finite),
                        // f32 error, taps scaled by the pixel's factor: err * f32(wgt * factor)
                        float ov[3], er[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) ov[c] = __fadd_rn(fa[c], oq10[c]);
                        // p.P is the descriptor whose outer cells are unbounded: the cell of the
                        // CLAMPED value lists every row that can be nearest to the value itself
                        bi = nearest_row_in<true, NSLOT, true>(
                            P, srch, cell_of(fminf(fmaxf(ov[0], 0.f), 255.f), fminf(fmaxf(ov[1], 0.f), 255.f),
                                             fminf(fmaxf(ov[2], 0.f), 255.f)),
                            ov[0], ov[1], ov[2], (double)ov[0], (double)ov[1], (double)ov[2], p.K);
                        oc = lds_u32(orgb_a + 4u * bi);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            er[c] = __fsub_rn(ov[c], BYTES ? byte_f32(__byte_perm(oc, 0u, 0x4440u + c))
                                                           : lds_f32(palf_a + 12u * bi + 4u * c));
                        const float w0 = __fmul_rn(0.4375f, fac), w1 = __fmul_rn(0.1875f, fac),
                                    w2 = __fmul_rn(0.3125f, fac), w3 = __fmul_rn(0.0625f, fac);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            oq10[c] = __fmul_rn(er[c], w0);
                            d1[0][c] = __fadd_rn(d1[0][c], __fmul_rn(er[c], w1));
                            d1[1][c] = __fadd_rn(d1[1][c], __fmul_rn(er[c], w2));
                            d1[2][c] = __fadd_rn(d1[2][c], __fmul_rn(er[c], w3));
                        }
                    } else {
                        float ov[3], er[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float a = __fadd_rn(fa[c], oq10[c]);
                            ov[c] = fminf(fmaxf(a, 0.f), 255.f);
                        }
                        bi = nearest_row<true, NSLOT>(P, srch, ov[0], ov[1], ov[2], (double)ov[0],
                                               (double)ov[1], (double)ov[2]);
                        oc = lds_u32(orgb_a + 4u * bi);
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            er[c] = __fsub_rn(ov[c], BYTES ? byte_f32(__byte_perm(oc, 0u, 0x4440u + c))
                                                           : lds_f32(palf_a + 12u * bi + 4u * c));
                        float lum = __fmul_rn(0.299f, ov[0]);
                        lum = __fadd_rn(lum, __fmul_rn(0.587f, ov[1]));
                        lum = __fadd_rn(lum, __fmul_rn(0.114f, ov[2]));
                        lum = fminf(fmaxf(lum, 0.f), 255.f);
                        const int li = __float_as_int(__fadd_rz(lum, 8388608.0f)) & 255;  // int(lum)
                        const float4 ow4 = lds_f32x4(ow_a + 16u * li);
                        const float w0 = ow4.x, w1 = ow4.y, w2 = ow4.z;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            oq10[c] = __fmul_rn(er[c], w0);
                            d1[1][c] = __fadd_rn(d1[1][c], __fmul_rn(er[c], w2));
                            d1[0][c] = __fadd_rn(d1[0][c], __fmul_rn(er[c], w1));
                        }
                    }
                    ob[3 * sidx] = (unsigned char)oc;
                    ob[3 * sidx + 1] = (unsigned char)(oc >> 8);
                    ob[3 * sidx + 2] = (unsigned char)(oc >> 16);
                    if (idx_f) idx_f[(size_t)y * W + x] = (unsigned char)bi;   // optional index plane
                    DP_STICK(4);
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
                    emitB[c] = SP::ROWS3 ? d2[0][c] : (T)0;
                }
                if (lane == 31) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        st.hout[sidx][c] = emitA[c];
                        if (SP::ROWS3) st.hout[sidx][3 + c] = emitB[c];
                    }
                }
#pragma unroll
                for (int j = 0; j + 1 < SP::W1; ++j) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) d1[j][c] = d1[j + 1][c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) d1[SP::W1 - 1][c] = 0;
                if (SP::ROWS3) {
#pragma unroll
                    for (int j = 0; j + 1 < SP::W2; ++j) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) d2[j][c] = d2[j + 1][c];
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) d2[SP::W2 - 1][c] = 0;
                }
                DP_STICK(5);
            }
            __syncwarp();
            DP_TICK(2);

            // ---- hand the outgoing streams to the band below first (it may be waiting) --------
            if (has_next) {
                // lane j holds step j: lane 31 was at column x31; stream A carries column
                // x31 - A1, stream B column x31 - A2; both are stored at the CONSUMER's step index
                const int x31 = x00 + lane - SP::S * 31;
                const int xa = x31 - SP::A1;
                if (xa >= 0 && xa < W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        __stcg(hout + (size_t)c * TPAD + xa + SP::BMAX, to_f32(st.hout[lane][c]));
                }
                if (SP::ROWS3) {
                    const int xb = x31 - SP::A2;
                    if (xb >= 0 && xb < W) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            __stcg(hout + (size_t)(3 + c) * TPAD + xb + SP::BMAX - SP::B1,
                                   to_f32(st.hout[lane][3 + c]));
                    }
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    // every consumer step index below this one is complete in both streams
                    const int prog = (ch == NCH - 1) ? 0x3fffffff : (t0 + 32 - SP::S * 32 + 1);
                    st_release(prog_out, prog);
                    // the band below can run its first chunk once chunk S of this band is out
                    if (ch == min(SP::S + p.slack, NCH - 1)) {
                        const int pos = atomicAdd(p.qctrl + 1, 1);
                        st_release(p.queue + pos, unit + 2);   // (unit + 1) + 1
                    }
                }
            }

            // ---- write the chunk's pixels: each lane stores its own row -------------------
            {
                const int xfirst = x00 - SP::S * lane;          // column of this lane's step 0
                const long long obyte = out0 + 96ll * ch;       // its byte offset in the frame
                unsigned *ow = reinterpret_cast<unsigned *>(st.outb[lane]);
                if (rgb_out && rowok && xfirst + 31 >= 0 && xfirst - 1 < W) {
                    const bool lead_ok = xfirst >= 1 || (xfirst == 0 && my_m == 0);
                    if (lead_ok && xfirst + 32 < W) {
                        unsigned *g = reinterpret_cast<unsigned *>(dst_f + (obyte - my_m));
#pragma unroll
                        for (int wd = 0; wd < 24; ++wd) g[wd] = ow[wd];
                    } else {
                        // row start / row end: byte by byte; k < my_m are the carried bytes of
                        // pixel xfirst-1
                        const unsigned char *sb = st.outb[lane];
                        for (int k = 0; k < my_m + 96; ++k) {
                            const int j = k - my_m;
                            const int xp = xfirst + (j >= 0 ? j / 3 : -1);
                            if (xp >= 0 && xp < W) dst_f[obyte + j] = sb[k];
                        }
                    }
                }
                ow[0] = ow[24];   // trailing partial word -> leading bytes of the next chunk
            }
            __syncwarp();
            DP_TICK(3);
        }
        cp_async_wait<0>();
    }
}

// ---------------------------------------------------------------------------------------
// Serpentine (or any) scan, serial per frame: row y+1 starts where row y ends, so there is no
// wavefront -- one WARP per frame walks the reference's loop pixel by pixel and spends its lanes
// inside the pixel: lane j evaluates candidate j of the pixel's cell in f64 (warp minimum, lowest
// lane among equals = first index), lane k applies tap k.  The 3-row f32 work ring lives in shared
// memory (in global memory for rows wider than ~6000 pixels).  Throughput comes from frames in
// flight; a single frame runs at the latency of this chain (~0.4 us per pixel).
// ---------------------------------------------------------------------------------------
struct SerialParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, K, serpentine, variant;
    float *ring;           // [frames][3][w][3], or null: the ring is in shared memory
    const float *ostro_w;  // ostromoukhov only
    double hyb_lum, hyb_col;   // hybrid only
    const float *plane;        // weighted only
};

__device__ __forceinline__ double warp_min_f64(double v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// nearest_first_exact by the whole warp: strict '<' in ascending candidate order = the minimum,
// lowest position among equals.  Every lane returns the row.
__device__ __forceinline__ int nearest_first_exact_warp(const Search &s, int cell, const uint4 e, double r, double g,
                                                        double b, int lane)
{
    const CandList L = cand_list(s, cell, e);
    double best = 1e20;
    int bi = 0;
    for (int base = 0; base < L.n; base += 32) {
        const int j = base + lane;
        int row = 0;
        double d = 1e300;
        if (j < L.n) {
            row = cand_at(L, j);
            d = dist_numba(s.s_pal + 3 * row, r, g, b);
        }
        const double m = warp_min_f64(d);
        if (m < best) {
            const unsigned eq = __ballot_sync(0xffffffffu, d == m);
            bi = __shfl_sync(0xffffffffu, row, __ffs(eq) - 1);
            best = m;
        }
    }
    return bi;
}

// nearest_kd_exact / nearest_kd_full by the whole warp: unique minimum, else scipy's traversal.
// `L` null: every palette row (points outside the colour cube).
__device__ __forceinline__ int nearest_kd_warp(const PalDev *P, const Search &s, const CandList *L, int K, double r,
                                               double g, double b, int lane)
{
    const int n = L ? L->n : K;
    double best = DP_INF_F64;
    int bi = 0;
    bool tie = false;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        int row = 0;
        double d = DP_INF_F64;
        if (j < n) {
            row = L ? cand_at(*L, j) : j;
            d = dist_scipy(s.s_pal + 3 * row, r, g, b);
        }
        const double m = warp_min_f64(d);
        const unsigned eq = __ballot_sync(0xffffffffu, j < n && d == m);
        if (m < best) {
            bi = __shfl_sync(0xffffffffu, row, __ffs(eq) - 1);
            best = m;
            tie = __popc(eq) > 1;
        } else if (m == best && eq) {
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

template <bool OSTRO, bool RING_SMEM>
__global__ void __launch_bounds__(32) k_diffuse_serial(const SerialParams p)
{
    extern __shared__ __align__(16) float s_ring[];
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
    srch.l1 = P->ed_l1;
    srch.pat = P->ed_pat;
    srch.flat = P->ed_flat;
    srch.l1_a = 0;
    srch.pat_a = 0;
    srch.P = P;
    srch.s_pal = s_pal;
    srch.rows_a = 0;   // exact search only

    const int W = p.w, H = p.h;
    const size_t frame_px = (size_t)W * H;
    for (int f = blockIdx.x; f < p.frames; f += gridDim.x) {
        const uint8_t *src_f = p.src + frame_px * 3 * f;
        uint8_t *dst_f = p.dst + frame_px * 3 * f;
        uint8_t *idx_f = p.dst_idx ? p.dst_idx + frame_px * f : nullptr;
        float *ring = RING_SMEM ? s_ring : p.ring + (size_t)f * 3 * W * 3;
        // rows 0..2 into the ring
        for (int rr = 0; rr < 3 && rr < H; ++rr)
            for (int i = lane; i < W * 3; i += 32)
                ring[(size_t)rr * W * 3 + i] = s_lutf[src_f[(size_t)rr * W * 3 + i]];
        __syncwarp();
        for (int y = 0; y < H; ++y) {
            float *r0 = ring + (size_t)(y % 3) * W * 3;
            float *r1 = ring + (size_t)((y + 1) % 3) * W * 3;
            float *r2 = ring + (size_t)((y + 2) % 3) * W * 3;
            const int dir = (p.serpentine && (y & 1)) ? -1 : 1;
            int x = dir > 0 ? 0 : W - 1;
            // every lane walks the row; the values it reads are the same in all lanes
            for (int n = 0; n < W; ++n, x += dir) {
                const float *px = r0 + 3 * x;
                const float pv[3] = {px[0], px[1], px[2]};
                int bi;
                if (!OSTRO && p.variant == V_WEIGHTED) {
                    // perceptual (:1042-1063), see the wavefront kernel
                    float er[3];
                    const bool inside = pv[0] >= 0.f && pv[0] <= 255.f && pv[1] >= 0.f && pv[1] <= 255.f &&
                                        pv[2] >= 0.f && pv[2] <= 255.f;
                    if (inside) {
                        const int wcell = cell_of(pv[0], pv[1], pv[2]);
                        const CandList L = cand_list(srch, wcell, __ldg(srch.flat + wcell));
                        bi = nearest_kd_warp(P, srch, &L, p.K, (double)pv[0], (double)pv[1], (double)pv[2], lane);
                    } else {
                        bi = nearest_kd_warp(P, srch, nullptr, p.K, (double)pv[0], (double)pv[1], (double)pv[2], lane);
                    }
                    for (int c = 0; c < 3; ++c) er[c] = __fsub_rn(pv[c], s_palf[3 * bi + c]);
                    const float fac = p.plane[((size_t)f * H + y) * W + x];
                    // lane t applies tap t: (x+1, y) 7/16, (x-1, y+1) 3/16, (x, y+1) 5/16, (x+1, y+1) 1/16
                    if (lane < 4) {
                        const float wt = __fmul_rn(lane == 0 ? 0.4375f : lane == 1 ? 0.1875f : lane == 2 ? 0.3125f : 0.0625f,
                                                   fac);
                        const int nx = lane == 1 ? x - 1 : lane == 2 ? x : x + 1;
                        const bool ok = nx >= 0 && nx < W && (lane == 0 || y + 1 < H);
                        if (ok) {
                            float *q = (lane == 0 ? r0 : r1) + 3 * nx;
                            for (int c = 0; c < 3; ++c) q[c] = __fadd_rn(q[c], __fmul_rn(er[c], wt));
                        }
                    }
                } else if (!OSTRO) {
                    double v[3], e[3];
                    for (int c = 0; c < 3; ++c) {
                        const double tv = (double)pv[c];
                        v[c] = tv < 0.0 ? 0.0 : (tv > 255.0 ? 255.0 : tv);
                    }
                    const int scell = cell_of((float)v[0], (float)v[1], (float)v[2]);
                    bi = nearest_first_exact_warp(srch, scell, __ldg(srch.flat + scell), v[0], v[1], v[2], lane);
                    for (int c = 0; c < 3; ++c) e[c] = __dsub_rn(v[c], (double)s_palf[3 * bi + c]);
                    if (p.variant == V_HYBRID) {   // _hybrid_numba :1447-1455
                        const double lum = __dadd_rn(
                            __dadd_rn(__dmul_rn(0.299, e[0]), __dmul_rn(0.587, e[1])),
                            __dmul_rn(0.114, e[2]));
                        const double cf[3] = {0.299, 0.587, 0.114};
                        for (int c = 0; c < 3; ++c) {
                            const double l = __dmul_rn(cf[c], lum);
                            e[c] = __dadd_rn(__dmul_rn(p.hyb_lum, l), __dmul_rn(p.hyb_col, __dsub_rn(e[c], l)));
                        }
                    }
                    if (lane < ntaps) {   // the taps of one pixel hit distinct pixels: one lane each
                        const int nx = x + s_tdx[lane] * dir;
                        const int dy = s_tdy[lane];
                        if (nx >= 0 && nx < W && y + dy < H) {
                            float *q = (dy == 0 ? r0 : dy == 1 ? r1 : r2) + 3 * nx;
                            const double tw = s_tw[lane];
                            for (int c = 0; c < 3; ++c) q[c] = acc_f64(q[c], __dmul_rn(e[c], tw));
                        }
                    }
                } else {
                    float ov[3], er[3];
                    for (int c = 0; c < 3; ++c) ov[c] = pv[c] < 0.f ? 0.f : (pv[c] > 255.f ? 255.f : pv[c]);
                    const int ocell = cell_of(ov[0], ov[1], ov[2]);
                    const CandList L = cand_list(srch, ocell, __ldg(srch.flat + ocell));
                    bi = nearest_kd_warp(P, srch, &L, p.K, (double)ov[0], (double)ov[1], (double)ov[2], lane);
                    for (int c = 0; c < 3; ++c) er[c] = __fsub_rn(ov[c], s_palf[3 * bi + c]);
                    float lum = __fmul_rn(0.299f, ov[0]);
                    lum = __fadd_rn(lum, __fmul_rn(0.587f, ov[1]));
                    lum = __fadd_rn(lum, __fmul_rn(0.114f, ov[2]));
                    lum = lum < 0.f ? 0.f : (lum > 255.f ? 255.f : lum);
                    const int li = (int)lum;
                    // lane t applies tap t: (x+dir, y) c0, (x-dir, y+1) c1, (x, y+1) c2
                    if (lane < 3) {
                        const float wt = p.ostro_w[4 * li + lane];
                        const int nx = lane == 0 ? x + dir : lane == 1 ? x - dir : x;
                        const bool ok = nx >= 0 && nx < W && (lane == 0 || y + 1 < H);
                        if (ok) {
                            float *q = (lane == 0 ? r0 : r1) + 3 * nx;
                            for (int c = 0; c < 3; ++c) q[c] = __fadd_rn(q[c], __fmul_rn(er[c], wt));
                        }
                    }
                }
                if (lane == 0) {
                    if (p.dst) {   // (null: index-plane-only output)
                        uint8_t *o = dst_f + ((size_t)y * W + x) * 3;
                        o[0] = s_orgb[4 * bi];
                        o[1] = s_orgb[4 * bi + 1];
                        o[2] = s_orgb[4 * bi + 2];
                    }
                    if (idx_f) idx_f[(size_t)y * W + x] = (uint8_t)bi;
                }
                __syncwarp();   // the taps are in the ring before any lane reads the next pixel
            }
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
        int dev = 0;
        DP_CUDA(cudaGetDevice(&dev));
        if (dp_retain_pool(dev)) return 1;
        DP_CUDA(cudaMallocAsync(&ptr, bytes ? bytes : 1, s));
        return 0;
    }
    ~Workspace()
    {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

__global__ void k_wave_init(int *progress, int *qctrl, int *queue, int units, int frames, int nbands)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < units; i += gridDim.x * blockDim.x) {
        progress[i] = 0;
        queue[i] = i < frames ? i * nbands + 1 : 0;   // band 0 of frame i, stored as unit + 1
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        qctrl[0] = 0;
        qctrl[1] = frames;
    }
}

// BIG selects the kernel instantiation (register budget); `big` the launch shape; NSLOT the
// number of candidate slots the screening pass evaluates.
template <int V, bool BIG, int NSLOT, bool BYTES>
int launch_wave_as(const WaveParams &p0, cudaStream_t st, int npat, bool big)
{
    using Stage = WarpStageT<typename StateOf<V>::T, Spec<V>::ROWS3 ? 6 : 3>;
    WaveParams p = p0;
    const int sms = dp_num_sms();
    const size_t base = DP_MAX_COLORS * 3 * 8 + (DP_MAX_COLORS + 1) * 16 + 256 * 8 +
                        (V == V_OSTRO ? DP_MAX_COLORS * 3 * 4 + 256 * 4 * 4 : 0) +
                        (V == V_WEIGHTED ? DP_MAX_COLORS * 3 * 4 : 0) + 256 * 4;
    const size_t limit = 227 * 1024;
    constexpr int maxw = wave_max_warps<V, BIG>();
    int warps = big ? maxw : 4;
    if (const char *ev = getenv("DP_WAVE_WARPS")) {   // tuning knob (tools/): 1..max warps per block
        const int w = atoi(ev);
        if (w >= 1 && w <= maxw) warps = w;
    }
    // First table level (64 KB) in shared memory: always for latency-bound launches; for
    // saturating batches only if it does not cost warps (occupancy is worth more there).
    p.l1_smem = (!big || base + 65536 + sizeof(Stage) * warps <= limit) ? 1 : 0;
    if (const char *ev = getenv("DP_WAVE_L1SMEM")) p.l1_smem = atoi(ev) ? 1 : 0;   // tuning knob
    const size_t fixed = base + (p.l1_smem ? 65536 : 0);
    while (warps > 1 && fixed + sizeof(Stage) * warps > limit) --warps;
    // the patterns join it when they fit beside the stages
    p.pat_smem = (p.l1_smem && fixed + sizeof(Stage) * warps + (size_t)npat * 16 <= limit) ? npat : 0;
    const size_t smem = fixed + sizeof(Stage) * warps + (size_t)p.pat_smem * 16;
    DP_CUDA(cudaFuncSetAttribute(k_diffuse_wave<V, BIG, NSLOT, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    int per_sm = 0;
    DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_diffuse_wave<V, BIG, NSLOT, BYTES>, warps * 32,
                                                          smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = ((long long)p.total_units + warps - 1) / warps;
    long long cap = (long long)sms * per_sm;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (const char *ev = getenv("DP_WAVE_GRID")) {   // test knob: fewer blocks than the device holds
        const int g = atoi(ev);
        if (g >= 1 && g < grid) grid = g;
    }
    k_diffuse_wave<V, BIG, NSLOT, BYTES><<<grid, warps * 32, smem, st>>>(p);
    DP_LAUNCH_CHECK();
    return 0;
}

// Few bands (a single image, a small batch): 4-warp blocks so that the bands spread over the SM
// sub-partitions (a lone warp is latency-bound).  Many bands: one block per SM with as many
// warps as registers (wave_max_warps) and shared memory allow.  Only the 5x5 footprints are
// compiled differently for the two regimes; for the others BIG merely selects the launch shape.
template <int V, int NSLOT, bool BYTES>
int launch_wave_n(const WaveParams &p, cudaStream_t st, int npat)
{
    const bool big = p.total_units > dp_num_sms() * 8;
    if constexpr (V == DP_ED_JJN || V == DP_ED_STUCKI) {
        return big ? launch_wave_as<V, true, NSLOT, BYTES>(p, st, npat, true)
                   : launch_wave_as<V, false, NSLOT, BYTES>(p, st, npat, false);
    } else {
        return launch_wave_as<V, true, NSLOT, BYTES>(p, st, npat, big);
    }
}

// p.bytes: identity input LUT and palette rows equal to their output bytes (every plain byte
// palette without gamma): the BYTES instantiations form work values and palette values from the
// bytes themselves instead of gathering them from shared-memory tables.
template <int V>
int launch_wave(const WaveParams &p, cudaStream_t st, int npat, bool four_slots)
{
    if (p.bytes)
        return four_slots ? launch_wave_n<V, 4, true>(p, st, npat) : launch_wave_n<V, 7, true>(p, st, npat);
    return four_slots ? launch_wave_n<V, 4, false>(p, st, npat) : launch_wave_n<V, 7, false>(p, st, npat);
}

template <int V>
constexpr int wave_extra_steps()
{
    return Spec<V>::AMAX + Spec<V>::BMAX + Spec<V>::S * 31;
}

// stream length = the kernel's step count rounded up to whole chunks (see k_diffuse_wave)
int wave_tpad(int variant, int w)
{
    int extra;
    switch (variant) {
        case DP_ED_FLOYD_STEINBERG: extra = wave_extra_steps<DP_ED_FLOYD_STEINBERG>(); break;
        case DP_ED_JJN: extra = wave_extra_steps<DP_ED_JJN>(); break;
        case DP_ED_STUCKI: extra = wave_extra_steps<DP_ED_STUCKI>(); break;
        case DP_ED_BURKES: extra = wave_extra_steps<DP_ED_BURKES>(); break;
        case DP_ED_ATKINSON: extra = wave_extra_steps<DP_ED_ATKINSON>(); break;
        case DP_ED_SIERRA: extra = wave_extra_steps<DP_ED_SIERRA>(); break;
        case DP_ED_SIERRA_TWO_ROW: extra = wave_extra_steps<DP_ED_SIERRA_TWO_ROW>(); break;
        case DP_ED_SIERRA_LITE: extra = wave_extra_steps<DP_ED_SIERRA_LITE>(); break;
        case V_HYBRID: extra = wave_extra_steps<V_HYBRID>(); break;
        case V_WEIGHTED: extra = wave_extra_steps<V_WEIGHTED>(); break;
        default: extra = wave_extra_steps<V_OSTRO>(); break;
    }
    return ((w + extra + 31) >> 5) << 5;
}

int run_diffusion(const dp_palette *pal, const uint8_t *src, int frames, int h, int w, int variant,
                  int serpentine, const float *ostro_w_host, uint8_t *dst, uint8_t *dst_idx,
                  cudaStream_t st, double hyb_lum = 0.0, double hyb_col = 0.0, const float *plane = nullptr)
{
    const bool ostro = (variant == V_OSTRO);
    const float *ostro_w = nullptr;
    if (ostro) {
        // The weight table depends only on the caller's coefficient table: keep one device copy
        // per device (re-uploaded, synchronously, only when the contents change) so that a call
        // neither allocates nor synchronises the stream -- a frame pipeline keeps overlapping.
        static std::mutex mu;
        static float *dev_w[64] = {nullptr};
        static float host_w[64][256 * 4];
        int dev = 0;
        DP_CUDA(cudaGetDevice(&dev));
        DP_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
        std::lock_guard<std::mutex> lk(mu);
        if (!dev_w[dev] || memcmp(host_w[dev], ostro_w_host, sizeof(host_w[dev])) != 0) {
            if (!dev_w[dev]) DP_CUDA(cudaMalloc(reinterpret_cast<void **>(&dev_w[dev]), sizeof(host_w[dev])));
            else DP_CUDA(cudaDeviceSynchronize());   // a running kernel may still read the old table
            DP_CUDA(cudaMemcpy(dev_w[dev], ostro_w_host, sizeof(host_w[dev]), cudaMemcpyHostToDevice));
            memcpy(host_w[dev], ostro_w_host, sizeof(host_w[dev]));
        }
        ostro_w = dev_w[dev];
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
        sp.hyb_lum = hyb_lum;
        sp.hyb_col = hyb_col;
        sp.plane = plane;
        // the 3-row ring in shared memory when it fits beside the kernel's static tables
        const size_t ring_bytes = (size_t)3 * w * 3 * sizeof(float);
        const bool ring_smem = ring_bytes <= 200 * 1024;
        Workspace ring;
        if (!ring_smem) {
            if (ring.alloc((size_t)frames * ring_bytes, st)) return 1;
            sp.ring = static_cast<float *>(ring.ptr);
        }
        const size_t smem = ring_smem ? ring_bytes : 0;
        int grid = frames < dp_num_sms() * 16 ? frames : dp_num_sms() * 16;
        if (ring_smem) {
            if (ostro) {
                DP_CUDA(cudaFuncSetAttribute(k_diffuse_serial<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024));
                k_diffuse_serial<true, true><<<grid, 32, smem, st>>>(sp);
            } else {
                DP_CUDA(cudaFuncSetAttribute(k_diffuse_serial<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024));
                k_diffuse_serial<false, true><<<grid, 32, smem, st>>>(sp);
            }
        } else if (ostro) {
            k_diffuse_serial<true, false><<<grid, 32, 0, st>>>(sp);
        } else {
            k_diffuse_serial<false, false><<<grid, 32, 0, st>>>(sp);
        }
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
    p.bytes = (!pal->has_lut && pal->pal_is_out && !getenv("DP_WAVE_NO_BYTES")) ? 1 : 0;
    p.K = pal->dev.K;
    p.ostro_w = ostro_w;
    p.hyb_lum = hyb_lum;
    p.hyb_col = hyb_col;
    p.plane = plane;
    Workspace hand, flags;
    if (hand.alloc((size_t)units * 6 * (size_t)wave_tpad(variant, w) * sizeof(float), st)) return 1;
    if (flags.alloc((size_t)(2 * units + 2) * sizeof(int), st)) return 1;
    p.hand = static_cast<float *>(hand.ptr);
    p.progress = static_cast<int *>(flags.ptr);
    p.qctrl = p.progress + units;
    p.queue = p.qctrl + 2;
    p.discard = ((reinterpret_cast<uintptr_t>(p.hand) & 127) == 0 && !getenv("DP_WAVE_NO_DISCARD")) ? 1 : 0;
    p.slack = 0;   // measured: extra lead only delays the successor (9.6 ms vs 11.0 ms at 32 frames)
    if (const char *ev = getenv("DP_WAVE_SLACK")) p.slack = atoi(ev) > 0 ? atoi(ev) : 0;   // tuning knob
    k_wave_init<<<(int)((units + 255) / 256 < 1024 ? (units + 255) / 256 : 1024), 256, 0, st>>>(
        p.progress, p.qctrl, p.queue, (int)units, frames, p.nbands);
    DP_LAUNCH_CHECK();
    int npat = pal->dev.ed_npat, gt4 = pal->dev.ed_gt4;
    if (variant == V_WEIGHTED && dp_palette_ext(const_cast<dp_palette *>(pal), &p.P, &npat, &gt4)) return 1;
    // four candidate slots are enough when at most 0.1 % of the cells hold a fifth candidate
    const bool four = gt4 <= 32 && !getenv("DP_WAVE_SEVEN");
    switch (variant) {
        case DP_ED_FLOYD_STEINBERG: return launch_wave<DP_ED_FLOYD_STEINBERG>(p, st, npat, four);
        case DP_ED_JJN: return launch_wave<DP_ED_JJN>(p, st, npat, four);
        case DP_ED_STUCKI: return launch_wave<DP_ED_STUCKI>(p, st, npat, four);
        case DP_ED_BURKES: return launch_wave<DP_ED_BURKES>(p, st, npat, four);
        case DP_ED_ATKINSON: return launch_wave<DP_ED_ATKINSON>(p, st, npat, four);
        case DP_ED_SIERRA: return launch_wave<DP_ED_SIERRA>(p, st, npat, four);
        case DP_ED_SIERRA_TWO_ROW: return launch_wave<DP_ED_SIERRA_TWO_ROW>(p, st, npat, four);
        case DP_ED_SIERRA_LITE: return launch_wave<DP_ED_SIERRA_LITE>(p, st, npat, four);
        case V_HYBRID: return launch_wave<V_HYBRID>(p, st, npat, four);
        case V_WEIGHTED: return launch_wave<V_WEIGHTED>(p, st, npat, four);
        default: return launch_wave<V_OSTRO>(p, st, npat, four);
    }
}

}  // namespace

#ifdef DP_WAVE_TIMING
extern "C" int dp_debug_wave_timing(unsigned long long *out, int reset)
{
    DP_CUDA(cudaMemcpyFromSymbol(out, g_wave_timing, sizeof(unsigned long long) * (128 * 4 + 4)));
    DP_CUDA(cudaMemcpyFromSymbol(out + 128 * 4 + 4, g_step_timing, sizeof(unsigned long long) * 16));
    if (reset) {
        static unsigned long long zeros[128 * 4 + 4];
        DP_CUDA(cudaMemcpyToSymbol(g_wave_timing, zeros, sizeof(zeros)));
        DP_CUDA(cudaMemcpyToSymbol(g_step_timing, zeros, sizeof(unsigned long long) * 16));
    }
    return 0;
}
#endif

extern "C" int dp_error_diffusion(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h,
                                  int w, int variant, int serpentine, uint8_t *dst_rgb,
                                  uint8_t *dst_idx, void *stream)
{
    DP_RANGE("dp_error_diffusion");
    DP_REQUIRE(pal && src_rgb, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
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
    DP_RANGE("dp_ostromoukhov");
    DP_REQUIRE(pal && src_rgb && coeffs, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
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

extern "C" int dp_hybrid(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                         double lum_factor, double col_factor, uint8_t *dst_rgb, uint8_t *dst_idx,
                         void *stream)
{
    DP_RANGE("dp_hybrid");
    DP_REQUIRE(pal && src_rgb, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    if (frames == 0 || h == 0 || w == 0) return 0;
    return run_diffusion(pal, src_rgb, frames, h, w, V_HYBRID, 0, nullptr, dst_rgb, dst_idx,
                         dp_stream(stream), lum_factor, col_factor);
}

namespace {

// perceptual factor plane (:1037, :1051-1052), all f32 with one rounding per operation:
// gray = ((0.299 R + 0.587 G) + 0.114 B) of the ORIGINAL pixel, factor = 0.5 + 0.5 (gray / 255)
__global__ void __launch_bounds__(256) k_perceptual_plane(const PalDev *P, const uint8_t *src, float *plane,
                                                          long long n)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = P->in_lut[threadIdx.x];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const uint8_t *q = src + 3 * i;
        const float r = (float)s_lut[q[0]], g = (float)s_lut[q[1]], b = (float)s_lut[q[2]];
        const float gray = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
        plane[i] = __fadd_rn(0.5f, __fmul_rn(0.5f, __fdiv_rn(gray, 255.0f)));
    }
}

}  // namespace

extern "C" int dp_perceptual(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                             uint8_t *dst_rgb, uint8_t *dst_idx, void *stream)
{
    DP_RANGE("dp_perceptual");
    DP_REQUIRE(pal && src_rgb, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    if (frames == 0 || h == 0 || w == 0) return 0;
    cudaStream_t st = dp_stream(stream);
    Workspace plane;
    const long long n = (long long)frames * h * w;
    if (plane.alloc((size_t)n * sizeof(float), st)) return 1;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)dp_num_sms() * 16;
    k_perceptual_plane<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(
        reinterpret_cast<const PalDev *>(pal->blob), src_rgb, static_cast<float *>(plane.ptr), n);
    DP_LAUNCH_CHECK();
    return run_diffusion(pal, src_rgb, frames, h, w, V_WEIGHTED, 0, nullptr, dst_rgb, dst_idx, st, 0.0, 0.0,
                         static_cast<const float *>(plane.ptr));
}

namespace {

// ---- adaptive variance: the gate plane (AdaptiveVarianceDitherStrategy, :989-1025) -------------
// gray and gray^2 of the original pixels (f32, one rounding per operation; numpy squares with a
// multiply)
__global__ void __launch_bounds__(256) k_av_gray(const PalDev *P, const uint8_t *src, float *g, float *g2,
                                                 long long n)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = P->in_lut[threadIdx.x];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const uint8_t *q = src + 3 * i;
        const float r = (float)s_lut[q[0]], gg = (float)s_lut[q[1]], b = (float)s_lut[q[2]];
        const float gray = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, gg)), __fmul_rn(0.114f, b));
        g[i] = gray;
        g2[i] = __fmul_rn(gray, gray);
    }
}

// scipy.ndimage.uniform_filter1d (scipy 1.18, third-party; ni_filters.c NI_UniformFilter1D) along
// one axis, mode='nearest', origin 0, f32 in / f32 out: the line is extended by `rad` edge copies
// and walked with a RUNNING SUM in double -- tmp = sum of the first window; out[0] = tmp / size;
// then tmp += new - old; out[l] = tmp / size -- one thread per line, two planes per launch.
// `stride` = element distance along the line, lines are enumerated by `line_of`.
__global__ void __launch_bounds__(128) k_av_uniform(const float *in_a, const float *in_b, float *out_a,
                                                    float *out_b, int frames, int h, int w, int axis, int rad)
{
    const long long nlines = (long long)frames * (axis == 0 ? w : h);
    const int len = axis == 0 ? h : w;
    const long long stride = axis == 0 ? w : 1;
    const double size = (double)(2 * rad + 1);
    for (long long t = (long long)blockIdx.x * 128 + threadIdx.x; t < 2 * nlines; t += (long long)gridDim.x * 128) {
        const bool second = t >= nlines;
        const long long line = second ? t - nlines : t;
        const float *in = second ? in_b : in_a;
        float *out = second ? out_b : out_a;
        const long long f = line / (axis == 0 ? w : h);
        const long long k = line - f * (axis == 0 ? w : h);
        const long long base = f * (long long)h * w + (axis == 0 ? k : k * w);
        auto at = [&](int i) -> double {
            i = i < 0 ? 0 : (i >= len ? len - 1 : i);
            return (double)in[base + (long long)i * stride];
        };
        double tmp = 0.0;
        for (int j = -rad; j <= rad; ++j) tmp = __dadd_rn(tmp, at(j));
        out[base] = __double2float_rn(__ddiv_rn(tmp, size));
        for (int l = 1; l < len; ++l) {
            tmp = __dadd_rn(tmp, __dsub_rn(at(l + rad), at(l - 1 - rad)));
            out[base + (long long)l * stride] = __double2float_rn(__ddiv_rn(tmp, size));
        }
    }
}

// The same filter along rows (axis 1) with coalesced traffic: a warp owns 32 consecutive rows of
// one plane and walks them in tiles of AVR_TW columns.  The tile plus its halo is loaded with the
// lane as the column index (32 independent loads per column group, fully unrolled) into shared
// memory; lane r then advances the running sum of ITS row through the tile sequentially -- the
// same operations in the same order as k_av_uniform -- and the results go back through shared
// memory as coalesced stores.
constexpr int AVR_MAXRAD = 16, AVR_TW = 96;
__global__ void __launch_bounds__(32) k_av_uniform_rows(const float *in_a, const float *in_b, float *out_a,
                                                        float *out_b, int frames, int h, int w, int rad)
{
    __shared__ float s_in[32][AVR_TW + 2 * AVR_MAXRAD + 3];   // odd row length: no bank conflicts
    __shared__ float s_out[32][AVR_TW + 1];
    const int lane = threadIdx.x;
    const long long rows_total = (long long)frames * h;       // rows of consecutive frames are contiguous
    const long long ngroups = (rows_total + 31) / 32;
    const double size = (double)(2 * rad + 1);
    const int tw = AVR_TW + 2 * rad + 1;                       // tile columns x0-rad-1 .. x0+AVR_TW-1+rad
    for (long long gidx = blockIdx.x; gidx < 2 * ngroups; gidx += gridDim.x) {
        const bool second = gidx >= ngroups;
        const long long g = second ? gidx - ngroups : gidx;
        const float *in = second ? in_b : in_a;
        float *out = second ? out_b : out_a;
        const long long row0 = g * 32;
        const int nrows = (int)(rows_total - row0 < 32 ? rows_total - row0 : 32);
        const bool rowok = lane < nrows;
        double tmp = 0.0;
        for (int x0 = 0; x0 < w; x0 += AVR_TW) {
            for (int c0 = 0; c0 < tw; c0 += 32) {
                const int c = c0 + lane;
                int x = x0 - rad - 1 + c;
                x = x < 0 ? 0 : (x >= w ? w - 1 : x);          // mode='nearest'
                float v[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = (r < nrows && c < tw) ? in[(row0 + r) * w + x] : 0.f;
                if (c < tw) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) s_in[r][c] = v[r];
                }
            }
            __syncwarp();
            const int nout = w - x0 < AVR_TW ? w - x0 : AVR_TW;
            if (rowok) {
                const float *line = s_in[lane];                // line[c] = element x0 - rad - 1 + c
                for (int k = 0; k < nout; ++k) {
                    if (x0 + k == 0) {
                        tmp = 0.0;
                        for (int j = 0; j <= 2 * rad; ++j) tmp = __dadd_rn(tmp, (double)line[1 + j]);
                    } else {
                        tmp = __dadd_rn(tmp, __dsub_rn((double)line[k + 2 * rad + 1], (double)line[k]));
                    }
                    s_out[lane][k] = __double2float_rn(__ddiv_rn(tmp, size));
                }
            }
            __syncwarp();
            for (int c0 = 0; c0 < nout; c0 += 32) {
                const int c = c0 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    if (r < nrows && c < nout) out[(row0 + r) * w + x0 + c] = s_out[r][c];
            }
            __syncwarp();
        }
    }
}

// var = max(0, mean_sq - mean^2) (f32), gate = var >= f32(threshold) -> factor 1 or 0
__global__ void __launch_bounds__(256) k_av_gate(const float *mean, const float *mean_sq, float thr, float *plane,
                                                 long long n)
{
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float m = mean[i];
        const float var = fmaxf(0.0f, __fsub_rn(mean_sq[i], __fmul_rn(m, m)));
        plane[i] = var >= thr ? 1.0f : 0.0f;
    }
}

}  // namespace

extern "C" int dp_adaptive_variance(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                                    double var_threshold, int window_radius, uint8_t *dst_rgb,
                                    uint8_t *dst_idx, void *stream)
{
    DP_RANGE("dp_adaptive_variance");
    DP_REQUIRE(pal && src_rgb, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    DP_REQUIRE(window_radius >= 0 && window_radius <= 64, "window radius out of range");
    if (frames == 0 || h == 0 || w == 0) return 0;
    cudaStream_t st = dp_stream(stream);
    const long long n = (long long)frames * h * w;
    Workspace wa, wb, wc, wd;
    if (wa.alloc((size_t)n * 4, st) || wb.alloc((size_t)n * 4, st) || wc.alloc((size_t)n * 4, st) ||
        wd.alloc((size_t)n * 4, st))
        return 1;
    float *A = static_cast<float *>(wa.ptr), *B = static_cast<float *>(wb.ptr), *C = static_cast<float *>(wc.ptr),
          *D = static_cast<float *>(wd.ptr);
    const long long cap = (long long)dp_num_sms() * 16;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < cap ? blocks : cap);
    k_av_gray<<<grid, 256, 0, st>>>(reinterpret_cast<const PalDev *>(pal->blob), src_rgb, A, B, n);
    DP_LAUNCH_CHECK();
    const float *mean = A, *mean_sq = B;
    if (window_radius >= 1) {   // size = 2 r + 1 > 1 on both axes: axis 0, then axis 1 on its f32 result
        const long long l0 = 2ll * frames * w, l1 = 2ll * frames * h;
        k_av_uniform<<<(int)((l0 + 127) / 128 < cap ? (l0 + 127) / 128 : cap), 128, 0, st>>>(A, B, C, D, frames, h, w,
                                                                                          0, window_radius);
        DP_LAUNCH_CHECK();
        if (window_radius <= AVR_MAXRAD && !getenv("DP_AV_SIMPLE_ROWS")) {
            const long long groups = 2 * (((long long)frames * h + 31) / 32);
            k_av_uniform_rows<<<(int)(groups < cap * 2 ? groups : cap * 2), 32, 0, st>>>(C, D, A, B, frames, h, w,
                                                                                     window_radius);
        } else {
            k_av_uniform<<<(int)((l1 + 127) / 128 < cap ? (l1 + 127) / 128 : cap), 128, 0, st>>>(C, D, A, B, frames, h,
                                                                                              w, 1, window_radius);
        }
        DP_LAUNCH_CHECK();
    }
    // python float threshold meets an f32 array element: numpy (NEP 50) compares in f32
    k_av_gate<<<grid, 256, 0, st>>>(mean, mean_sq, (float)var_threshold, C, n);
    DP_LAUNCH_CHECK();
    return run_diffusion(pal, src_rgb, frames, h, w, V_WEIGHTED, 0, nullptr, dst_rgb, dst_idx, st, 0.0, 0.0, C);
}
