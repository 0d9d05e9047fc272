// dp_halftone.cu -- newspaper halftone.
//
// Replaces HalftoneDitherStrategy.dither (dithering_lib.py:1597-1644) and
// _generate_halftone_screen_with_cells (:1646-1695).
//
//   pass 0 (frame-invariant)  rotated-grid cell id and dot screen per pixel, in f64 with one
//                             rounding per numpy ufunc (no transcendental when dot_gain == 1);
//                             the screen value s is stored as the GRAY THRESHOLD g*(s) = the
//                             smallest f32 gray with 1 - gray/255 <= s (bisection over the f32
//                             bit patterns with the reference's own division and subtraction):
//                             both rounded operations are monotone, so
//                             `1 - gray/255 > s`  <=>  `gray < g*(s)` exactly, and pass 3 needs
//                             neither the division nor the subtraction.  Maps are cached per
//                             device and parameter set across calls (video: same size every batch)
//   pass 1                    exact integer RGB sums and counts per cell (atomics)
//   pass 2                    per cell: mean colour in f64 -> KD-tree nearest palette row
//   pass 3                    per pixel: ink (cell colour) where 1 - gray/255 > screen, else paper
// Algorithmic bytes: 3 read + 3 written per pixel (the maps are an implementation cost).
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "dp_search.cuh"

namespace {

struct HtParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, npix, K;
    int cell_size, shape;
    double ca, sa, min_dot, span, sharp;
    int sharpen;
    int cx_min, cy_min, ncx, ncells;
    float *screen;              // gray thresholds g*(screen value), see the file header
    const float *user_screen;   // caller-provided screen values (dot_gain != 1) or null
    int *cell;
    unsigned long long *sums;   // [frames][ncells][2]: (sum r | sum g << 32), (sum b | count << 32)
    uint8_t *cell_pal;   // [frames][ncells]
    int paper;
    int make_screen;
    int has_lut;   // the palette carries a gamma input LUT (else the LUT is the identity)
    int vec16;     // rows and buffers allow 128-bit strip loads (w % 16 == 0, 16-byte aligned)
};

// numpy's floored modulo for doubles (npy_divmod): fmod, then fix the sign.
__device__ __forceinline__ double np_mod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m = __dadd_rn(m, b);
    } else {
        m = copysign(0.0, b);
    }
    return m;
}

// smallest non-negative f32 g with  1 - g/255 <= s  (the reference's f32 ops, :1638-1640);
// 0 <= s <= 1 after the reference's clip, so g = 256 always satisfies it
__device__ __forceinline__ float gray_threshold(float s)
{
    // Bit patterns of non-negative floats are ordered like the values, and both rounded operations
    // are monotone, so  fl(1 - fl(g / 255)) <= s  <=>  fl(g / 255) >= q*  with q* the smallest q
    // satisfying fl(1 - q) <= s.  Stage 1 bisects q* (an add per step, no division); stage 2 finds
    // the smallest g whose quotient reaches q* among the few floats around q* * 255 (the division
    // is correctly rounded, so the answer lies within a couple of ulps of the product).
    unsigned lo = 0u, hi = 0x40000000u;              // 0.0f .. 2.0f; fl(1 - 2) = -1 <= s always
#pragma unroll 1
    for (int it = 0; it < 31; ++it) {
        const unsigned mid = (lo + hi) >> 1;
        const bool le = __fsub_rn(1.0f, __uint_as_float(mid)) <= s;
        hi = (lo < hi && le) ? mid : hi;
        lo = (lo < hi && !le) ? mid + 1u : lo;
    }
    const float qs = __uint_as_float(hi);
    if (hi == 0u) return 0.0f;                       // s >= 1: every gray qualifies, ink never
    const unsigned g0 = __float_as_uint(__fmul_rn(qs, 255.0f));
    unsigned g = g0 > 8u ? g0 - 8u : 0u;
    if (g > 0u && __fdiv_rn(__uint_as_float(g), 255.0f) >= qs) {
        // (never observed) the window does not bracket the answer: whole-range bisection
        unsigned a = 0u, b = 0x44000000u;            // 512.0f / 255 > 2 >= q*
#pragma unroll 1
        while (a < b) {
            const unsigned mid = (a + b) >> 1;
            if (__fdiv_rn(__uint_as_float(mid), 255.0f) >= qs) b = mid;
            else a = mid + 1u;
        }
        return __uint_as_float(b);
    }
#pragma unroll 1
    while (__fdiv_rn(__uint_as_float(g), 255.0f) < qs) ++g;   // <= 16 steps
    return __uint_as_float(g);
}

__global__ void __launch_bounds__(256) k_ht_maps(const HtParams p)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= p.npix) return;
    const int y = i / p.w, x = i - y * p.w;
    const double xd = (double)x, yd = (double)y, cs = (double)p.cell_size;
    const double xr = __dsub_rn(__dmul_rn(xd, p.ca), __dmul_rn(yd, p.sa));
    const double yr = __dadd_rn(__dmul_rn(xd, p.sa), __dmul_rn(yd, p.ca));
    const int cx = (int)floor(__ddiv_rn(xr, cs));
    const int cy = (int)floor(__ddiv_rn(yr, cs));
    p.cell[i] = (cy - p.cy_min) * p.ncx + (cx - p.cx_min);
    if (!p.make_screen) {
        p.screen[i] = gray_threshold(__ldg(p.user_screen + i));
        return;
    }
    const double dx = __dsub_rn(__ddiv_rn(np_mod(xr, cs), cs), 0.5);
    const double dy = __dsub_rn(__ddiv_rn(np_mod(yr, cs), cs), 0.5);
    double dist, dmax;
    if (p.shape == 1) {
        dist = fmax(fabs(dx), fabs(dy));
        dmax = 0.5;
    } else if (p.shape == 2) {
        dist = __dadd_rn(fabs(dx), fabs(dy));
        dmax = 1.0;
    } else {
        dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        dmax = 0.5;
    }
    double t = __ddiv_rn(dist, dmax);
    t = fmin(fmax(t, 0.0), 1.0);
    // dot_gain == 1.0: x ** 1.0 == x
    t = __dadd_rn(p.min_dot, __dmul_rn(t, p.span));
    if (p.sharpen) t = __dadd_rn(0.5, __dmul_rn(__dsub_rn(t, 0.5), p.sharp));
    t = fmin(fmax(t, 0.0), 1.0);
    p.screen[i] = gray_threshold(__double2float_rn(t));
}

// Per-cell integer RGB sums and pixel counts.  A warp walks 32 consecutive pixels; consecutive
// pixels mostly share a cell (a cell of size c covers runs of up to ~1.4 c pixels of a row), so
// only the last lane of each run touches global memory: two 64-bit reductions per run instead of
// four 32-bit atomics per pixel.  Run sums come from ONE unsegmented inclusive warp scan of the
// packed channel values (r | g << 16 and b; 32 * 255 < 2^16) -- prefix at the run's tail minus
// prefix just before its head -- with the run boundaries taken from a ballot of "cell changed".
// Exact as long as a cell's channel sum is below 2^32.
__global__ void __launch_bounds__(256) k_ht_sums(const HtParams p)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.y;
    const uint8_t *src = p.src + (size_t)f * p.npix * 3;
    unsigned long long *sums = p.sums + (size_t)f * p.ncells * 2;
    const int nchunk = (p.npix + 31) >> 5;
    const int wpb = 256 / 32;
    for (int ch = blockIdx.x * wpb + (threadIdx.x >> 5); ch < nchunk; ch += gridDim.x * wpb) {
        const int i = ch * 32 + lane;
        const bool ok = i < p.npix;
        int c = -1 - lane;                       // distinct, never a cell id
        unsigned p0 = 0, p1 = 0;
        if (ok) {
            const uint8_t *q = src + (size_t)i * 3;
            c = __ldg(p.cell + i);
            p0 = (unsigned)s_lut[q[0]] | ((unsigned)s_lut[q[1]] << 16);
            p1 = (unsigned)s_lut[q[2]];
        }
        const int pc = __shfl_up_sync(FULL, c, 1);
        const unsigned heads = __ballot_sync(FULL, lane == 0 || pc != c);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t0 = __shfl_up_sync(FULL, p0, d), t1 = __shfl_up_sync(FULL, p1, d);
            if (lane >= d) {
                p0 += t0;
                p1 += t1;
            }
        }
        // head of this lane's run = highest head bit at or below the lane
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const unsigned b0 = __shfl_sync(FULL, p0, (start + 31) & 31), b1 = __shfl_sync(FULL, p1, (start + 31) & 31);
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (ok && tail) {
            const unsigned s0 = p0 - (start ? b0 : 0u), s1 = p1 - (start ? b1 : 0u);   // fields never borrow
            const unsigned long long a = (unsigned long long)(s0 & 0xffffu) |
                                         ((unsigned long long)(s0 >> 16) << 32);
            const unsigned long long b = (unsigned long long)s1 |
                                         ((unsigned long long)(unsigned)(lane - start + 1) << 32);
            atomicAdd(sums + 2 * (size_t)c, a);
            atomicAdd(sums + 2 * (size_t)c + 1, b);
        }
    }
}

// Tiled variant (the default): a block owns a 64 x 128 pixel tile and adds the run sums into a
// shared-memory image of the cell grid the tile touches (32-bit shared atomics: r, g, b, count),
// then flushes every touched cell ONCE with two 64-bit global reductions.  A cell of size c is
// met by ~c rows of a tile, so this issues ~c/1.5 times fewer global reductions than one per
// run (they are bound by the L2 atomic units).  The tile's cell range comes from its four corners
// (the rotated coordinates are monotone in x and y, also after rounding); tiles whose range does
// not fit the shared grid (tiny cells) are handled by k_ht_sums.
constexpr int HT_TW = 128, HT_TH = 64, HT_CAP = 2048;

__device__ __forceinline__ void ht_cell_xy(const HtParams &p, int x, int y, int &cx, int &cy)
{
    const double xd = (double)x, yd = (double)y, cs = (double)p.cell_size;
    const double xr = __dsub_rn(__dmul_rn(xd, p.ca), __dmul_rn(yd, p.sa));
    const double yr = __dadd_rn(__dmul_rn(xd, p.sa), __dmul_rn(yd, p.ca));
    cx = (int)floor(__ddiv_rn(xr, cs));
    cy = (int)floor(__ddiv_rn(yr, cs));
}

// MODE 0: four 32-bit shared atomics per run; 1: two (16-bit fields: a cell holds < 257 pixels,
// cell_size <= 14); 2: no shared grid, two 64-bit global reductions per run
template <int MODE>
__global__ void __launch_bounds__(256) k_ht_sums_tile(const HtParams p, int tiles_x)
{
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned s_grid[HT_CAP * 4];
    __shared__ int s_geo[4];   // base cell id, local width, local height
    s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * HT_TW, y0 = ty * HT_TH;
    const int x1 = min(x0 + HT_TW, p.w) - 1, y1 = min(y0 + HT_TH, p.h) - 1;
    if (threadIdx.x < 32) {
        // one corner per lane 0..3, range by warp shuffles
        const unsigned FULL = 0xffffffffu;
        const int k = threadIdx.x & 3;
        int cx, cy;
        ht_cell_xy(p, (k & 1) ? x1 : x0, (k & 2) ? y1 : y0, cx, cy);
        int cxl = cx, cxh = cx, cyl = cy, cyh = cy;
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            cxl = min(cxl, __shfl_xor_sync(FULL, cxl, d));
            cxh = max(cxh, __shfl_xor_sync(FULL, cxh, d));
            cyl = min(cyl, __shfl_xor_sync(FULL, cyl, d));
            cyh = max(cyh, __shfl_xor_sync(FULL, cyh, d));
        }
        if (threadIdx.x == 0) {
            s_geo[0] = (cyl - p.cy_min) * p.ncx + (cxl - p.cx_min);
            s_geo[1] = cxh - cxl + 1;
            s_geo[2] = cyh - cyl + 1;
        }
    }
    __syncthreads();
    const int base = s_geo[0], lw = s_geo[1], lh = s_geo[2];
    const int nloc = lw * lh;           // <= HT_CAP (checked by the host for the worst tile)
    if (MODE != 2) {
        if (nloc > HT_CAP) __trap();    // host bound violated: an error, never a silent overrun
        for (int i = threadIdx.x; i < nloc * (MODE == 1 ? 2 : 4); i += 256) s_grid[i] = 0;
        __syncthreads();
    }

    // A thread walks strips of 16 consecutive pixels of a row (48 source bytes = three 128-bit
    // loads, 16 cell ids = four) and keeps the sums of the current run of equal cell ids in
    // registers; a run is added to the shared grid when the cell changes and at the strip's end.
    const int f = blockIdx.y;
    const uint8_t *src = p.src + (size_t)f * p.npix * 3;
    const unsigned mdiv = (unsigned)((0x100000000ull + (unsigned)p.ncx - 1) / (unsigned)p.ncx);
    const unsigned skip = (unsigned)(p.ncx - lw);
    const bool lut = p.has_lut != 0;
    unsigned long long *gsums = p.sums + (size_t)f * p.ncells * 2;
    auto flush = [&](int cell, unsigned rg, unsigned bn) {
        if (MODE == 2) {
            atomicAdd(gsums + 2 * (size_t)cell, (unsigned long long)(rg & 0xffffu) | ((unsigned long long)(rg >> 16) << 32));
            atomicAdd(gsums + 2 * (size_t)cell + 1, (unsigned long long)(bn & 0xffffu) | ((unsigned long long)(bn >> 16) << 32));
            return;
        }
        const unsigned d = (unsigned)(cell - base);
        const unsigned ly = __umulhi(d, mdiv);       // d / ncx, exact for d < 2^32 / ncx
        if (MODE == 1) {
            unsigned *g = s_grid + 2 * (d - ly * skip);
            atomicAdd(g, rg);
            atomicAdd(g + 1, bn);
        } else {
            unsigned *g = s_grid + 4 * (d - ly * skip);  // ly * lw + lx
            atomicAdd(g, rg & 0xffffu);
            atomicAdd(g + 1, rg >> 16);
            atomicAdd(g + 2, bn & 0xffffu);
            atomicAdd(g + 3, bn >> 16);
        }
    };
    constexpr int SPR = HT_TW / 16;                  // strips per tile row
    for (int sidx = threadIdx.x; sidx < HT_TH * SPR; sidx += 256) {
        const int r = sidx / SPR, y = y0 + r;
        const int xs = x0 + (sidx - r * SPR) * 16;
        if (y > y1 || xs > x1) continue;
        const int i0 = y * p.w + xs;
        const int n = min(16, x1 + 1 - xs);
        unsigned wsrc[12];
        int cid[16];
        if (n == 16 && p.vec16) {
            const uint4 *q = reinterpret_cast<const uint4 *>(src + (size_t)i0 * 3);
            const uint4 a = __ldcs(q), b = __ldcs(q + 1), c = __ldcs(q + 2);
            wsrc[0] = a.x; wsrc[1] = a.y; wsrc[2] = a.z; wsrc[3] = a.w;
            wsrc[4] = b.x; wsrc[5] = b.y; wsrc[6] = b.z; wsrc[7] = b.w;
            wsrc[8] = c.x; wsrc[9] = c.y; wsrc[10] = c.z; wsrc[11] = c.w;
            const int4 *cq = reinterpret_cast<const int4 *>(p.cell + i0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int4 t = __ldg(cq + k);
                cid[4 * k] = t.x; cid[4 * k + 1] = t.y; cid[4 * k + 2] = t.z; cid[4 * k + 3] = t.w;
            }
        } else {
            // ragged strips / unaligned rows: byte loads, fully unrolled (no local-memory arrays)
            const uint8_t *q = src + (size_t)i0 * 3;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                unsigned wv = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * k + j < 3 * n) wv |= (unsigned)q[4 * k + j] << (8 * j);
                wsrc[k] = wv;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) cid[k] = k < n ? __ldg(p.cell + i0 + k) : -1;
        }
        int cur = cid[0];
        unsigned rg = 0, bn = 0;                     // r | g << 16, b | count << 16 (<= 16 pixels)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            // pixel k: bytes 3k .. 3k+2 of the strip
            const unsigned lo = wsrc[(3 * k) >> 2], hi = wsrc[((3 * k) >> 2) + ((3 * k) % 4 > 1 ? 1 : 0)];
            const unsigned v = (3 * k) % 4 > 1 ? __funnelshift_r(lo, hi, 8 * ((3 * k) % 4)) : lo >> (8 * ((3 * k) % 4));
            unsigned prg, pbn;                        // r | g << 16,  b | 1 << 16
            if (lut) {
                prg = (unsigned)s_lut[v & 255u] | ((unsigned)s_lut[(v >> 8) & 255u] << 16);
                pbn = (unsigned)s_lut[(v >> 16) & 255u] | 0x10000u;
            } else {
                prg = __byte_perm(v, 0u, 0x4140);
                pbn = __byte_perm(v, 1u, 0x5452);
            }
            if (k && cid[k] != cur) {
                if (cur >= 0) flush(cur, rg, bn);
                rg = bn = 0;
                cur = cid[k];
            }
            if (k < n) {
                rg += prg;
                bn += pbn;
            }
        }
        if (cur >= 0) flush(cur, rg, bn);
    }
    if (MODE == 2) return;
    __syncthreads();
    for (int i = threadIdx.x; i < nloc; i += 256) {
        uint4 v;
        if (MODE == 1) {
            const uint2 t = reinterpret_cast<const uint2 *>(s_grid)[i];
            v = make_uint4(t.x & 0xffffu, t.x >> 16, t.y & 0xffffu, t.y >> 16);
        } else {
            v = reinterpret_cast<const uint4 *>(s_grid)[i];
        }
        if (!v.w) continue;
        const int ly = i / lw, lx = i - ly * lw;
        const size_t c = (size_t)(base + ly * p.ncx + lx);
        atomicAdd(gsums + 2 * c, (unsigned long long)v.x | ((unsigned long long)v.y << 32));
        atomicAdd(gsums + 2 * c + 1, (unsigned long long)v.z | ((unsigned long long)v.w << 32));
    }
}

// KD-tree k=1 on an arbitrary f64 point (:1633)
__device__ __noinline__ int nearest_kd_f64(const PalDev *P, double x0, double x1, double x2)
{
    const int K = P->K;
    double best = DP_INF_F64;
    int bi = 0;
    bool tie = false;
    for (int i = 0; i < K; ++i) {
        const double *pp = P->pal_f64 + 3 * i;
        const double d0 = __dsub_rn(pp[0], x0), d1 = __dsub_rn(pp[1], x1), d2 = __dsub_rn(pp[2], x2);
        const double d = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(d0, d0)), __dmul_rn(d1, d1)),
                                   __dmul_rn(d2, d2));
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
        kd_emulate<1>(P, x0, x1, x2, oi, os);
        bi = oi[0];
    }
    return bi;
}

__global__ void __launch_bounds__(128) k_ht_cells(const HtParams p)
{
    const int f = blockIdx.y;
    const unsigned long long *sums = p.sums + (size_t)f * p.ncells * 2;
    uint8_t *cp = p.cell_pal + (size_t)f * p.ncells;
    for (int c = blockIdx.x * 128 + threadIdx.x; c < p.ncells; c += gridDim.x * 128) {
        const unsigned long long sa = sums[2 * c], sb = sums[2 * c + 1];
        const unsigned n = (unsigned)(sb >> 32);
        if (!n) continue;
        const double dn = (double)n;
        const double m0 = __ddiv_rn((double)(unsigned)sa, dn);
        const double m1 = __ddiv_rn((double)(unsigned)(sa >> 32), dn);
        const double m2 = __ddiv_rn((double)(unsigned)sb, dn);
        cp[c] = (uint8_t)nearest_kd_f64(p.P, m0, m1, m2);
    }
}

__global__ void __launch_bounds__(256) k_ht_select(const HtParams p)
{
    __shared__ uint8_t s_lut[256];
    __shared__ uint8_t s_orgb[DP_MAX_COLORS * 4];
    s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    for (int i = threadIdx.x; i < p.K * 4; i += 256) s_orgb[i] = p.P->out_rgb[i];
    __syncthreads();
    const int f = blockIdx.y;
    const uint8_t *src = p.src + (size_t)f * p.npix * 3;
    uint8_t *dst = p.dst + (size_t)f * p.npix * 3;
    const uint8_t *cp = p.cell_pal + (size_t)f * p.ncells;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < p.npix; i += gridDim.x * 256) {
        const uint8_t *q = src + (size_t)i * 3;
        const float r = (float)s_lut[q[0]], g = (float)s_lut[q[1]], b = (float)s_lut[q[2]];
        // (:1605-1606, :1638-1639) f32, one rounding per operation
        float gray = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)),
                               __fmul_rn(0.114f, b));
        const int idx = (gray < __ldg(p.screen + i)) ? cp[__ldg(p.cell + i)] : p.paper;
        if (p.dst) {   // (null: index-plane-only output)
            uint8_t *o = dst + (size_t)i * 3;
            o[0] = s_orgb[4 * idx];
            o[1] = s_orgb[4 * idx + 1];
            o[2] = s_orgb[4 * idx + 2];
        }
        if (p.dst_idx) p.dst_idx[(size_t)f * p.npix + i] = (uint8_t)idx;
    }
}

// Four pixels per thread with word accesses (frames whose pixel count is a multiple of 4 and
// 4-byte aligned buffers): 12 bytes in, 12 bytes out, the threshold and cell maps as 128-bit loads.
// Byte -> float without a conversion instruction: 2^23 + byte as a bit pattern (one PRMT), and
// fl(c * byte) = fma(c, 2^23 + byte, -c * 2^23): the fma rounds the exact product c * byte once,
// like the reference's multiply (c * 2^23 is exact).
// (Measured alternatives at 1080p x64: sixteen pixels per thread with 128-bit source loads 0.27 ms,
// the same staged through shared memory with cp.async at two blocks per SM 0.43 ms, this kernel
// 0.22 ms -- the per-thread chain load -> decide -> gather -> colour -> store wants many warps.)
__global__ void __launch_bounds__(256) k_ht_select4(const HtParams p)
{
    __shared__ uint8_t s_lut[256];
    __shared__ unsigned s_orgb[DP_MAX_COLORS];
    s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    for (int i = threadIdx.x; i < p.K; i += 256) {
        const uint8_t *o = p.P->out_rgb + 4 * i;
        s_orgb[i] = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16);
    }
    __syncthreads();
    const int f = blockIdx.y;
    const unsigned *src = reinterpret_cast<const unsigned *>(p.src + (size_t)f * p.npix * 3);
    unsigned *dst = reinterpret_cast<unsigned *>(p.dst + (size_t)f * p.npix * 3);
    const uint8_t *cp = p.cell_pal + (size_t)f * p.ncells;
    const int ngroups = p.npix >> 2;
    const bool lut = p.has_lut != 0;
    for (int gi = blockIdx.x * 256 + threadIdx.x; gi < ngroups; gi += gridDim.x * 256) {
        const unsigned w0 = __ldcs(src + 3 * gi), w1 = __ldcs(src + 3 * gi + 1), w2 = __ldcs(src + 3 * gi + 2);
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.screen) + gi);
        const int4 ce = __ldg(reinterpret_cast<const int4 *>(p.cell) + gi);
        const unsigned v[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
        const float scr[4] = {sc.x, sc.y, sc.z, sc.w};
        const int cel[4] = {ce.x, ce.y, ce.z, ce.w};
        unsigned col[4], idx4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned mr, mg, mb;   // 2^23 + byte as a float bit pattern
            if (lut) {
                mr = 0x4b000000u | s_lut[v[k] & 255u];
                mg = 0x4b000000u | s_lut[(v[k] >> 8) & 255u];
                mb = 0x4b000000u | s_lut[(v[k] >> 16) & 255u];
            } else {
                mr = __byte_perm(v[k], 0x4b000000u, 0x7540);
                mg = __byte_perm(v[k], 0x4b000000u, 0x7541);
                mb = __byte_perm(v[k], 0x4b000000u, 0x7542);
            }
            const float gray = __fadd_rn(
                __fadd_rn(__fmaf_rn(0.299f, __uint_as_float(mr), -0.299f * 8388608.0f),
                          __fmaf_rn(0.587f, __uint_as_float(mg), -0.587f * 8388608.0f)),
                __fmaf_rn(0.114f, __uint_as_float(mb), -0.114f * 8388608.0f));
            int idx = p.paper;
            if (gray < scr[k]) idx = (int)__ldg(cp + cel[k]);
            col[k] = s_orgb[idx];
            idx4 |= (unsigned)idx << (8 * k);
        }
        if (p.dst) {   // (null: index-plane-only output)
            __stcs(dst + 3 * gi, col[0] | (col[1] << 24));
            __stcs(dst + 3 * gi + 1, (col[1] >> 8) | (col[2] << 16));
            __stcs(dst + 3 * gi + 2, (col[2] >> 16) | (col[3] << 8));
        }
        if (p.dst_idx) reinterpret_cast<unsigned *>(p.dst_idx + (size_t)f * p.npix)[gi] = idx4;
    }
}

// ---------------------------------------------------------------------------------------
// v2 kernels (round 2): rows of a multiple of 16 pixels, 16-byte aligned buffers, cell_size <= 14,
// tile cell range within HT_CAP.  Three more frame-invariant maps make the per-pixel work
// branch-free and keep every cell look-up in shared memory:
//   tgeo  [tiles]     int4 (base cell id, local grid width, local grid height) of a 64 x 128 tile
//   loc16 [npix] u16  4 * (tile-local cell number ly * lw + lx): the byte offset of the cell's
//                     entry in the select kernel's colour table, half the offset of its sums entry
//   bmask [npix/16]   bit k: pixel k of the 16-pixel strip starts a new run of equal cells
// Global sums are ONE u64 per cell (r | g << 16 | b << 32 | n << 48; a cell of size <= 14 holds at
// most 225 pixels), the cells kernel leaves rgb | row << 24 per cell.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ht_tilegeo(const HtParams p, int tiles_x, int ntiles, int4 *tgeo)
{
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ntiles) return;
    const int ty = t / tiles_x, tx = t - ty * tiles_x;
    const int x0 = tx * HT_TW, y0 = ty * HT_TH;
    const int x1 = min(x0 + HT_TW, p.w) - 1, y1 = min(y0 + HT_TH, p.h) - 1;
    int cxl = 0, cxh = 0, cyl = 0, cyh = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        int cx, cy;
        ht_cell_xy(p, (k & 1) ? x1 : x0, (k & 2) ? y1 : y0, cx, cy);
        if (k == 0) {
            cxl = cxh = cx;
            cyl = cyh = cy;
        } else {
            cxl = min(cxl, cx); cxh = max(cxh, cx);
            cyl = min(cyl, cy); cyh = max(cyh, cy);
        }
    }
    tgeo[t] = make_int4((cyl - p.cy_min) * p.ncx + (cxl - p.cx_min), cxh - cxl + 1, cyh - cyl + 1, 0);
}

// one thread per 16-pixel strip: tile-local cell offsets and the run-start mask from the cell map
__global__ void __launch_bounds__(256) k_ht_locmap(const HtParams p, int tiles_x, const int4 *__restrict__ tgeo,
                                                   uint16_t *__restrict__ loc16, uint16_t *__restrict__ bmask)
{
    const int spr = p.w >> 4;
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= spr * p.h) return;
    const int y = s / spr, xs = (s - y * spr) << 4;
    const int4 geo = __ldg(tgeo + (y / HT_TH) * tiles_x + xs / HT_TW);
    const int i0 = y * p.w + xs;
    unsigned packed[8];
    unsigned mask = 0;
    int prev = -1;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int d = __ldg(p.cell + i0 + k) - geo.x;
        const int ly = d / p.ncx, lx = d - ly * p.ncx;
        const int loc = ly * geo.y + lx;
        if (d < 0 || lx >= geo.y || ly >= geo.z || loc >= HT_CAP) __trap();   // never: the corners bound the range
        if (k && loc != prev) mask |= 1u << k;
        prev = loc;
        const unsigned v = (unsigned)(loc * 4) & 0xffffu;
        if (k & 1) packed[k >> 1] |= v << 16;
        else packed[k >> 1] = v;
    }
    uint4 *o = reinterpret_cast<uint4 *>(loc16 + i0);
    o[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    o[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    bmask[i0 >> 4] = (uint16_t)mask;
}

// 128-bit loads as volatile asm: they keep their program order, so a strip's loads are all in
// flight before the first use (left to itself the compiler threads them through the pixel loop
// to save registers and every pixel group waits for its own load)
__device__ __forceinline__ uint4 ht_ld_stream(const void *q)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q));
    return v;
}
__device__ __forceinline__ uint4 ht_ld_map(const void *q)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q));
    return v;
}

// bytes 3k .. 3k+2 of a 48-byte strip held in twelve words, as r | g << 16 and b | 1 << 16
__device__ __forceinline__ void ht_px_packed(const unsigned (&w)[12], int k, unsigned &prg, unsigned &pbn)
{
    const int o = 3 * k, j = o >> 2, sh = o & 3;
    if (sh == 0) {
        prg = __byte_perm(w[j], 0u, 0x4140);
        pbn = __byte_perm(w[j], 1u, 0x5452);
    } else if (sh == 1) {
        prg = __byte_perm(w[j], 0u, 0x4241);
        pbn = __byte_perm(w[j], 1u, 0x5453);
    } else if (sh == 2) {
        prg = __byte_perm(w[j], 0u, 0x4342);
        pbn = __byte_perm(w[j + 1], 1u, 0x5450);
    } else {
        const unsigned v = __funnelshift_r(w[j], w[j + 1], 24);
        prg = __byte_perm(v, 0u, 0x4140);
        pbn = __byte_perm(v, 1u, 0x5452);
    }
}

__device__ __forceinline__ unsigned ht_byte(const unsigned (&w)[12], int b)
{
    return (w[b >> 2] >> (8 * (b & 3))) & 255u;
}

template <bool LUT, int OCC>
__global__ void __launch_bounds__(256, OCC) k_ht_sums_v2(const HtParams p, int tiles_x, const int4 *__restrict__ tgeo,
                                                    const uint16_t *__restrict__ loc16,
                                                    const uint16_t *__restrict__ bmask,
                                                    unsigned long long *__restrict__ sums64)
{
    __shared__ uint8_t s_lut[256];
    __shared__ __align__(16) unsigned s_grid[HT_CAP * 2];
    if (LUT) s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * HT_TW, y0 = ty * HT_TH;
    const int x1 = min(x0 + HT_TW, p.w) - 1, y1 = min(y0 + HT_TH, p.h) - 1;
    const int4 geo = __ldg(tgeo + blockIdx.x);
    const int nloc = geo.y * geo.z;
    for (int i = threadIdx.x; i < nloc * 2; i += 256) s_grid[i] = 0;
    __syncthreads();

    const int f = blockIdx.y;
    const uint8_t *src = p.src + (size_t)f * p.npix * 3;
    char *grid = reinterpret_cast<char *>(s_grid);
    constexpr int SPR = HT_TW / 16;
#pragma unroll 1
    for (int sidx = threadIdx.x; sidx < HT_TH * SPR; sidx += 256) {
        const int r = sidx / SPR, y = y0 + r;
        const int xs = x0 + (sidx - r * SPR) * 16;
        if (y > y1 || xs > x1) continue;
        const int i0 = y * p.w + xs;
        unsigned w[12], lo[8];
        {
            const uint4 *q = reinterpret_cast<const uint4 *>(src + (size_t)i0 * 3);
            const uint4 a = ht_ld_stream(q), b = ht_ld_stream(q + 1), c = ht_ld_stream(q + 2);
            const uint4 *lq = reinterpret_cast<const uint4 *>(loc16 + i0);
            const uint4 l0 = ht_ld_map(lq), l1 = ht_ld_map(lq + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
            lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w;
            lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
        }
        const unsigned m = __ldg(bmask + (i0 >> 4));
        unsigned rg = 0, bn = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            unsigned prg, pbn;
            if (LUT) {
                prg = (unsigned)s_lut[ht_byte(w, 3 * k)] | ((unsigned)s_lut[ht_byte(w, 3 * k + 1)] << 16);
                pbn = (unsigned)s_lut[ht_byte(w, 3 * k + 2)] | 0x10000u;
            } else {
                ht_px_packed(w, k, prg, pbn);
            }
            if (k && ((m >> k) & 1u)) {
                // the run that ended at pixel k-1: entry at 2 * loc16 bytes
                const unsigned off = ((k - 1) & 1) ? (lo[(k - 1) >> 1] >> 16) : (lo[(k - 1) >> 1] & 0xffffu);
                unsigned *g = reinterpret_cast<unsigned *>(grid + 2 * off);
                atomicAdd(g, rg);
                atomicAdd(g + 1, bn);
                rg = 0;
                bn = 0;
            }
            rg += prg;
            bn += pbn;
        }
        {
            unsigned *g = reinterpret_cast<unsigned *>(grid + 2 * (lo[7] >> 16));
            atomicAdd(g, rg);
            atomicAdd(g + 1, bn);
        }
    }
    __syncthreads();
    unsigned long long *gs = sums64 + (size_t)f * p.ncells;
    for (int i = threadIdx.x; i < nloc; i += 256) {
        const uint2 t = reinterpret_cast<const uint2 *>(s_grid)[i];
        if (!(t.y >> 16)) continue;
        const int ly = i / geo.y, lx = i - ly * geo.y;
        atomicAdd(gs + (size_t)(geo.x + ly * p.ncx + lx), (unsigned long long)t.x | ((unsigned long long)t.y << 32));
    }
}

// per cell: mean colour in f64 -> nearest palette row (scipy query(k=1) semantics), stored as
// out_rgb | row << 24.  All K distances are screened in f32 first: with palette and mean in
// [0, 255] the absolute error of an f32 squared distance is below 0.1, so a runner-up more than
// 0.5 above the minimum makes the minimum the exact (and untied) answer; otherwise the exact f64
// search with the KD-tree tie replay decides.
__global__ void __launch_bounds__(128) k_ht_cells_v2(const HtParams p, const unsigned long long *__restrict__ sums64,
                                                     unsigned *__restrict__ cell_col)
{
    __shared__ float4 s_pal[DP_MAX_COLORS];
    for (int i = threadIdx.x; i < p.K; i += 128) {
        const float *q = p.P->pal_f32 + 3 * i;
        s_pal[i] = make_float4(q[0], q[1], q[2], 0.f);
    }
    __syncthreads();
    const int f = blockIdx.y;
    const unsigned long long *sums = sums64 + (size_t)f * p.ncells;
    unsigned *cc = cell_col + (size_t)f * p.ncells;
    for (int c = blockIdx.x * 128 + threadIdx.x; c < p.ncells; c += gridDim.x * 128) {
        const unsigned long long s = sums[c];
        const unsigned n = (unsigned)(s >> 48);
        if (!n) continue;
        const double dn = (double)n;
        const double m0 = __ddiv_rn((double)(unsigned)(s & 0xffffu), dn);
        const double m1 = __ddiv_rn((double)(unsigned)((s >> 16) & 0xffffu), dn);
        const double m2 = __ddiv_rn((double)(unsigned)((s >> 32) & 0xffffu), dn);
        const float x0 = (float)m0, x1 = (float)m1, x2 = (float)m2;
        float b1 = 3.0e38f, b2 = 3.0e38f;
        int bi = 0;
        for (int i = 0; i < p.K; ++i) {
            const float4 q = s_pal[i];
            const float d0 = q.x - x0, d1 = q.y - x1, d2 = q.z - x2;
            const float d = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
            if (d < b1) {
                b2 = b1;
                b1 = d;
                bi = i;
            } else if (d < b2) {
                b2 = d;
            }
        }
        if (!(b2 - b1 > 0.5f)) bi = nearest_kd_f64(p.P, m0, m1, m2);
        const uint8_t *o = p.P->out_rgb + 4 * bi;
        cc[c] = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16) | ((unsigned)bi << 24);
    }
}

// ink / paper per pixel: a block owns a 64 x 128 tile, the tile's cell colours sit in shared
// memory (one LDS per pixel at the offset the loc16 map holds), a thread owns 16-pixel strips
// (three 128-bit loads, three 128-bit stores, four + two for the maps).
template <bool LUT, bool RGB, bool IDX, int OCC>
__global__ void __launch_bounds__(256, OCC) k_ht_select_v2(const HtParams p, int tiles_x, const int4 *__restrict__ tgeo,
                                                      const uint16_t *__restrict__ loc16,
                                                      const unsigned *__restrict__ cell_col)
{
    __shared__ uint8_t s_lut[256];
    __shared__ __align__(16) unsigned s_col[HT_CAP];
    __shared__ unsigned s_paper;
    if (LUT) s_lut[threadIdx.x] = p.P->in_lut[threadIdx.x];
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * HT_TW, y0 = ty * HT_TH;
    const int x1 = min(x0 + HT_TW, p.w) - 1, y1 = min(y0 + HT_TH, p.h) - 1;
    const int4 geo = __ldg(tgeo + blockIdx.x);
    const int nloc = geo.y * geo.z;
    const int f = blockIdx.y;
    const unsigned *cc = cell_col + (size_t)f * p.ncells;
    for (int i = threadIdx.x; i < nloc; i += 256) {
        const int ly = i / geo.y, lx = i - ly * geo.y;
        s_col[i] = __ldg(cc + (geo.x + ly * p.ncx + lx));
    }
    if (threadIdx.x == 0) {
        const uint8_t *o = p.P->out_rgb + 4 * p.paper;
        s_paper = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16) | ((unsigned)p.paper << 24);
    }
    __syncthreads();
    const unsigned paper = s_paper;
    const uint8_t *src = p.src + (size_t)f * p.npix * 3;
    uint8_t *dst = RGB ? p.dst + (size_t)f * p.npix * 3 : nullptr;
    uint8_t *dix = IDX ? p.dst_idx + (size_t)f * p.npix : nullptr;
    const char *colt = reinterpret_cast<const char *>(s_col);
    constexpr int SPR = HT_TW / 16;
#pragma unroll 1
    for (int sidx = threadIdx.x; sidx < HT_TH * SPR; sidx += 256) {
        const int r = sidx / SPR, y = y0 + r;
        const int xs = x0 + (sidx - r * SPR) * 16;
        if (y > y1 || xs > x1) continue;
        const int i0 = y * p.w + xs;
        unsigned w[12], lo[8];
        float th[16];
        {
            const uint4 *q = reinterpret_cast<const uint4 *>(src + (size_t)i0 * 3);
            const uint4 a = ht_ld_stream(q), b = ht_ld_stream(q + 1), c = ht_ld_stream(q + 2);
            const uint4 *lq = reinterpret_cast<const uint4 *>(loc16 + i0);
            const uint4 l0 = ht_ld_map(lq), l1 = ht_ld_map(lq + 1);
            const uint4 *tq = reinterpret_cast<const uint4 *>(p.screen + i0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 t = ht_ld_map(tq + k);
                th[4 * k] = __uint_as_float(t.x); th[4 * k + 1] = __uint_as_float(t.y);
                th[4 * k + 2] = __uint_as_float(t.z); th[4 * k + 3] = __uint_as_float(t.w);
            }
            // every load of the strip is issued before the first use
            asm volatile("" ::"r"(a.x), "r"(b.x), "r"(c.x), "r"(l0.x), "r"(l1.x), "f"(th[0]), "f"(th[4]), "f"(th[8]),
                         "f"(th[12]));
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
            lo[0] = l0.x; lo[1] = l0.y; lo[2] = l0.z; lo[3] = l0.w;
            lo[4] = l1.x; lo[5] = l1.y; lo[6] = l1.z; lo[7] = l1.w;
        }
        unsigned col[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            unsigned mr, mg, mb;   // 2^23 + byte as a float bit pattern
            if (LUT) {
                mr = 0x4b000000u | s_lut[ht_byte(w, 3 * k)];
                mg = 0x4b000000u | s_lut[ht_byte(w, 3 * k + 1)];
                mb = 0x4b000000u | s_lut[ht_byte(w, 3 * k + 2)];
            } else {
                mr = __byte_perm(w[(3 * k) >> 2], 0x4b000000u, 0x7540 + ((3 * k) & 3));
                mg = __byte_perm(w[(3 * k + 1) >> 2], 0x4b000000u, 0x7540 + ((3 * k + 1) & 3));
                mb = __byte_perm(w[(3 * k + 2) >> 2], 0x4b000000u, 0x7540 + ((3 * k + 2) & 3));
            }
            // (:1605-1606, :1638-1639) f32, one rounding per operation: fl(c * byte) as one fma
            const float gray = __fadd_rn(
                __fadd_rn(__fmaf_rn(0.299f, __uint_as_float(mr), -0.299f * 8388608.0f),
                          __fmaf_rn(0.587f, __uint_as_float(mg), -0.587f * 8388608.0f)),
                __fmaf_rn(0.114f, __uint_as_float(mb), -0.114f * 8388608.0f));
            const unsigned off = (k & 1) ? (lo[k >> 1] >> 16) : (lo[k >> 1] & 0xffffu);
            unsigned c = paper;
            if (gray < th[k]) c = *reinterpret_cast<const unsigned *>(colt + off);
            col[k] = c;
        }
        if (RGB) {
            uint4 *o = reinterpret_cast<uint4 *>(dst + (size_t)i0 * 3);
            unsigned ow[12];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                ow[3 * g] = __byte_perm(col[4 * g], col[4 * g + 1], 0x4210);
                ow[3 * g + 1] = __byte_perm(col[4 * g + 1], col[4 * g + 2], 0x5421);
                ow[3 * g + 2] = __byte_perm(col[4 * g + 2], col[4 * g + 3], 0x6542);
            }
            __stcs(o, make_uint4(ow[0], ow[1], ow[2], ow[3]));
            __stcs(o + 1, make_uint4(ow[4], ow[5], ow[6], ow[7]));
            __stcs(o + 2, make_uint4(ow[8], ow[9], ow[10], ow[11]));
        }
        if (IDX) {
            unsigned iw[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const unsigned t01 = __byte_perm(col[4 * g], col[4 * g + 1], 0x0073);
                const unsigned t23 = __byte_perm(col[4 * g + 2], col[4 * g + 3], 0x0073);
                iw[g] = __byte_perm(t01, t23, 0x5410);
            }
            __stcs(reinterpret_cast<uint4 *>(dix + i0), make_uint4(iw[0], iw[1], iw[2], iw[3]));
        }
    }
}

struct Ws {
    void *ptr = nullptr;
    cudaStream_t st = nullptr;
    ~Ws()
    {
        if (ptr) cudaFreeAsync(ptr, st);
    }
};

// Frame-invariant maps, kept per device and parameter set between calls (a video hands over the
// same geometry batch after batch).  Entries own plain cudaMalloc memory; `ready` orders the map
// kernel before any consumer on another stream.  Never destroyed at exit (no CUDA calls from
// static destructors).
struct MapEntry {
    int dev, h, w, cell_size, shape, sharpen;
    double ca, sa, min_dot, span, sharp;
    float *gthr;
    int *cell;
    // v2 maps (null until a call that can use them asks)
    uint16_t *loc16, *bmask;
    int4 *tgeo;
    cudaEvent_t ready;
    unsigned long long stamp;
};
struct V2Maps {
    uint16_t *loc16 = nullptr, *bmask = nullptr;
    int4 *tgeo = nullptr;
};
struct MapCache {
    std::mutex mu;
    std::vector<MapEntry> ent;
    unsigned long long clock = 0;
};
MapCache &map_cache()
{
    static MapCache *c = new MapCache;
    return *c;
}
constexpr size_t HT_MAP_CACHE = 4;

// tile geometry, tile-local cell offsets and run-start masks from the cell map (on `st`)
int build_v2_maps(const HtParams &p, const V2Maps &v, int tiles_x, int ntiles, cudaStream_t st)
{
    k_ht_tilegeo<<<(ntiles + 255) / 256, 256, 0, st>>>(p, tiles_x, ntiles, v.tgeo);
    DP_LAUNCH_CHECK();
    const int strips = p.npix / 16;
    k_ht_locmap<<<(strips + 255) / 256, 256, 0, st>>>(p, tiles_x, v.tgeo, v.loc16, v.bmask);
    DP_LAUNCH_CHECK();
    return 0;
}

// returns 0 and fills p.screen / p.cell (and `v2` when asked); launches the map kernels on `st`
// on a miss
int cached_maps(HtParams &p, cudaStream_t st, V2Maps *v2, int tiles_x, int ntiles)
{
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    MapCache &mc = map_cache();
    std::lock_guard<std::mutex> lk(mc.mu);
    for (MapEntry &e : mc.ent) {
        if (e.dev == dev && e.h == p.h && e.w == p.w && e.cell_size == p.cell_size && e.shape == p.shape &&
            e.sharpen == p.sharpen && e.ca == p.ca && e.sa == p.sa && e.min_dot == p.min_dot &&
            e.span == p.span && e.sharp == p.sharp) {
            e.stamp = ++mc.clock;
            p.screen = e.gthr;
            p.cell = e.cell;
            DP_CUDA(cudaStreamWaitEvent(st, e.ready, 0));
            if (v2) {
                if (!e.loc16) {
                    V2Maps n;
                    DP_CUDA(cudaMalloc(&n.loc16, (size_t)p.npix * 2));
                    DP_CUDA(cudaMalloc(&n.bmask, (size_t)(p.npix / 16) * 2));
                    DP_CUDA(cudaMalloc(&n.tgeo, (size_t)ntiles * sizeof(int4)));
                    if (build_v2_maps(p, n, tiles_x, ntiles, st)) return 1;
                    DP_CUDA(cudaEventRecord(e.ready, st));
                    e.loc16 = n.loc16;
                    e.bmask = n.bmask;
                    e.tgeo = n.tgeo;
                }
                v2->loc16 = e.loc16;
                v2->bmask = e.bmask;
                v2->tgeo = e.tgeo;
            }
            return 0;
        }
    }
    if (mc.ent.size() >= HT_MAP_CACHE) {
        size_t v = 0;
        for (size_t i = 1; i < mc.ent.size(); ++i)
            if (mc.ent[i].stamp < mc.ent[v].stamp) v = i;
        // cudaFree waits for the device: no in-flight kernel can still read the evicted maps
        cudaFree(mc.ent[v].gthr);
        cudaFree(mc.ent[v].cell);
        cudaFree(mc.ent[v].loc16);
        cudaFree(mc.ent[v].bmask);
        cudaFree(mc.ent[v].tgeo);
        cudaEventDestroy(mc.ent[v].ready);
        mc.ent.erase(mc.ent.begin() + v);
    }
    MapEntry e;
    e.dev = dev; e.h = p.h; e.w = p.w; e.cell_size = p.cell_size; e.shape = p.shape; e.sharpen = p.sharpen;
    e.ca = p.ca; e.sa = p.sa; e.min_dot = p.min_dot; e.span = p.span; e.sharp = p.sharp;
    e.gthr = nullptr;
    e.cell = nullptr;
    e.loc16 = e.bmask = nullptr;
    e.tgeo = nullptr;
    e.stamp = ++mc.clock;
    DP_CUDA(cudaMalloc(&e.gthr, (size_t)p.npix * 4));
    if (cudaMalloc(&e.cell, (size_t)p.npix * 4) != cudaSuccess) {
        cudaFree(e.gthr);
        DP_REQUIRE(false, "out of device memory for the halftone maps");
    }
    DP_CUDA(cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming));
    p.screen = e.gthr;
    p.cell = e.cell;
    k_ht_maps<<<(p.npix + 255) / 256, 256, 0, st>>>(p);
    DP_LAUNCH_CHECK();
    if (v2) {
        V2Maps n;
        DP_CUDA(cudaMalloc(&n.loc16, (size_t)p.npix * 2));
        DP_CUDA(cudaMalloc(&n.bmask, (size_t)(p.npix / 16) * 2));
        DP_CUDA(cudaMalloc(&n.tgeo, (size_t)ntiles * sizeof(int4)));
        if (build_v2_maps(p, n, tiles_x, ntiles, st)) return 1;
        e.loc16 = v2->loc16 = n.loc16;
        e.bmask = v2->bmask = n.bmask;
        e.tgeo = v2->tgeo = n.tgeo;
    }
    DP_CUDA(cudaEventRecord(e.ready, st));
    mc.ent.push_back(e);
    return 0;
}

}  // namespace

extern "C" int dp_halftone(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                           int cell_size, double cos_a, double sin_a, double dot_gain,
                           double min_dot, double max_dot, int shape, double sharpness,
                           const float *screen, uint8_t *dst_rgb, uint8_t *dst_idx, void *stream)
{
    DP_RANGE("dp_halftone");
    DP_REQUIRE(pal && src_rgb, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0 && cell_size >= 1, "bad size");
    DP_REQUIRE(shape >= 0 && shape <= 2, "unknown dot shape");
    DP_REQUIRE(screen || dot_gain == 1.0,
               "dot_gain != 1 needs a pow(): pass the screen computed by the caller");
    if (frames == 0 || h == 0 || w == 0) return 0;
    DP_REQUIRE((long long)h * w < (1ll << 31), "frame too large");
    cudaStream_t st = dp_stream(stream);

    HtParams p;
    memset(&p, 0, sizeof(p));
    p.P = reinterpret_cast<const PalDev *>(pal->blob);
    p.src = src_rgb;
    p.dst = dst_rgb;
    p.dst_idx = dst_idx;
    p.frames = frames;
    p.h = h;
    p.w = w;
    p.npix = h * w;
    p.K = pal->dev.K;
    p.cell_size = cell_size;
    p.shape = shape;
    p.ca = cos_a;
    p.sa = sin_a;
    p.min_dot = min_dot;
    p.span = max_dot - min_dot;
    p.sharp = sharpness;
    p.sharpen = sharpness != 1.0;
    p.make_screen = screen == nullptr;
    p.has_lut = pal->has_lut;
    p.vec16 = (w % 16 == 0) && (reinterpret_cast<uintptr_t>(src_rgb) & 15) == 0;

    // cell index range: the rotated coordinates are linear in (x, y), so the extremes are at
    // the image corners; same operation order as the device (separately rounded ops)
    {
        int cxmin = 0, cxmax = 0, cymin = 0, cymax = 0;
        bool first = true;
        for (int k = 0; k < 4; ++k) {
            volatile double xd = (k & 1) ? (double)(w - 1) : 0.0;
            volatile double yd = (k & 2) ? (double)(h - 1) : 0.0;
            volatile double a = xd * cos_a, b = yd * sin_a, c = xd * sin_a, d = yd * cos_a;
            volatile double xr = a - b, yr = c + d;
            volatile double qx = xr / (double)cell_size, qy = yr / (double)cell_size;
            int cx = (int)floor(qx), cy = (int)floor(qy);
            if (first || cx < cxmin) cxmin = cx;
            if (first || cx > cxmax) cxmax = cx;
            if (first || cy < cymin) cymin = cy;
            if (first || cy > cymax) cymax = cy;
            first = false;
        }
        p.cx_min = cxmin;
        p.cy_min = cymin;
        p.ncx = cxmax - cxmin + 1;
        long long nc = (long long)p.ncx * (cymax - cymin + 1);
        DP_REQUIRE(nc < (1ll << 28), "too many halftone cells");
        p.ncells = (int)nc;
    }
    // paper = first argmax of f32 palette luma (:1609-1610)
    {
        int best = 0;
        float bl = 0.f;
        for (int i = 0; i < p.K; ++i) {
            volatile float a = 0.299f * pal->host_pal[3 * i];
            volatile float b = 0.587f * pal->host_pal[3 * i + 1];
            volatile float c = 0.114f * pal->host_pal[3 * i + 2];
            volatile float ab = a + b;
            volatile float l = ab + c;
            if (i == 0 || l > bl) {
                bl = l;
                best = i;
            }
        }
        p.paper = best;
    }

    Ws w_screen, w_cell, w_sums, w_cp, w_loc, w_mask, w_geo;
    w_screen.st = w_cell.st = w_sums.st = w_cp.st = w_loc.st = w_mask.st = w_geo.st = st;
    const int sms = dp_num_sms();
    // worst-case cell range of a tile: |d(xr)| <= TW |cos| + TH |sin| across a tile (+2 for
    // the floor at both ends and rounding), likewise yr
    const double ex = (HT_TW * fabs(cos_a) + HT_TH * fabs(sin_a)) / cell_size + 2.0;
    const double ey = (HT_TW * fabs(sin_a) + HT_TH * fabs(cos_a)) / cell_size + 2.0;
    const int tiles_x = (w + HT_TW - 1) / HT_TW, tiles_y = (h + HT_TH - 1) / HT_TH;
    const int ntiles = tiles_x * tiles_y;
    int mode = (ceil(ex) * ceil(ey) <= (double)HT_CAP) ? (cell_size <= 14 ? 1 : 0) : 2;
    if (const char *ev = getenv("DP_HT_MODE")) mode = atoi(ev);   // tuning knob (tools/)
    // the v2 kernels: 16-pixel strips, 128-bit accesses, 16-bit sums per tile cell
    const bool v2 = mode == 1 && cell_size <= 14 && ceil(ex) * ceil(ey) <= (double)HT_CAP && p.vec16 &&
                    ((reinterpret_cast<uintptr_t>(dst_rgb) | reinterpret_cast<uintptr_t>(dst_idx)) & 15) == 0 &&
                    !getenv("DP_HT_V1");
    V2Maps vm;
    if (p.make_screen && !getenv("DP_HT_NO_MAP_CACHE")) {
        if (cached_maps(p, st, v2 ? &vm : nullptr, tiles_x, ntiles)) return 1;
    } else {
        DP_CUDA(cudaMallocAsync(&w_screen.ptr, (size_t)p.npix * 4, st));
        p.screen = static_cast<float *>(w_screen.ptr);
        p.user_screen = screen;
        DP_CUDA(cudaMallocAsync(&w_cell.ptr, (size_t)p.npix * 4, st));
        p.cell = static_cast<int *>(w_cell.ptr);
        k_ht_maps<<<(p.npix + 255) / 256, 256, 0, st>>>(p);
        DP_LAUNCH_CHECK();
        if (v2) {
            DP_CUDA(cudaMallocAsync(&w_loc.ptr, (size_t)p.npix * 2, st));
            DP_CUDA(cudaMallocAsync(&w_mask.ptr, (size_t)(p.npix / 16) * 2, st));
            DP_CUDA(cudaMallocAsync(&w_geo.ptr, (size_t)ntiles * sizeof(int4), st));
            vm.loc16 = static_cast<uint16_t *>(w_loc.ptr);
            vm.bmask = static_cast<uint16_t *>(w_mask.ptr);
            vm.tgeo = static_cast<int4 *>(w_geo.ptr);
            if (build_v2_maps(p, vm, tiles_x, ntiles, st)) return 1;
        }
    }
    if (v2) {
        const size_t sums_bytes = (size_t)frames * p.ncells * sizeof(unsigned long long);
        DP_CUDA(cudaMallocAsync(&w_sums.ptr, sums_bytes, st));
        unsigned long long *sums64 = static_cast<unsigned long long *>(w_sums.ptr);
        DP_CUDA(cudaMallocAsync(&w_cp.ptr, (size_t)frames * p.ncells * 4, st));
        unsigned *cell_col = static_cast<unsigned *>(w_cp.ptr);
        DP_CUDA(cudaMemsetAsync(sums64, 0, sums_bytes, st));
        const dim3 tg(ntiles, frames);
        // OCC: blocks per SM the register allocation aims at (4: a strip's loads all in flight before
        // the first use, 6: ptxas threads the loads through the pixel loop to save registers)
        int occ_sums = 6, occ_sel = 4;
        if (const char *ev = getenv("DP_HT_OCC")) {   // tuning knob (tools/): "<sums><select>", e.g. 64
            if (ev[0] == '4' || ev[0] == '6') occ_sums = ev[0] - '0';
            if (ev[0] && (ev[1] == '4' || ev[1] == '6')) occ_sel = ev[1] - '0';
        }
        if (p.has_lut) k_ht_sums_v2<true, 6><<<tg, 256, 0, st>>>(p, tiles_x, vm.tgeo, vm.loc16, vm.bmask, sums64);
        else if (occ_sums == 4) k_ht_sums_v2<false, 4><<<tg, 256, 0, st>>>(p, tiles_x, vm.tgeo, vm.loc16, vm.bmask, sums64);
        else k_ht_sums_v2<false, 6><<<tg, 256, 0, st>>>(p, tiles_x, vm.tgeo, vm.loc16, vm.bmask, sums64);
        DP_LAUNCH_CHECK();
        int gc = (p.ncells + 127) / 128;
        if (gc > sms * 8) gc = sms * 8;
        k_ht_cells_v2<<<dim3(gc, frames), 128, 0, st>>>(p, sums64, cell_col);
        DP_LAUNCH_CHECK();
#define HT_SEL(L, R, I)                                                                              \
    do {                                                                                             \
        if (occ_sel == 4) k_ht_select_v2<L, R, I, 4><<<tg, 256, 0, st>>>(p, tiles_x, vm.tgeo, vm.loc16, cell_col); \
        else k_ht_select_v2<L, R, I, 6><<<tg, 256, 0, st>>>(p, tiles_x, vm.tgeo, vm.loc16, cell_col);              \
    } while (0)
        const int variant = (p.has_lut ? 4 : 0) | (dst_rgb ? 2 : 0) | (dst_idx ? 1 : 0);
        switch (variant) {
        case 1: HT_SEL(false, false, true); break;
        case 2: HT_SEL(false, true, false); break;
        case 3: HT_SEL(false, true, true); break;
        case 5: HT_SEL(true, false, true); break;
        case 6: HT_SEL(true, true, false); break;
        default: HT_SEL(true, true, true); break;
        }
#undef HT_SEL
        DP_LAUNCH_CHECK();
        return 0;
    }
    size_t sums_bytes = (size_t)frames * p.ncells * 2 * sizeof(unsigned long long);
    DP_CUDA(cudaMallocAsync(&w_sums.ptr, sums_bytes, st));
    p.sums = static_cast<unsigned long long *>(w_sums.ptr);
    DP_CUDA(cudaMallocAsync(&w_cp.ptr, (size_t)frames * p.ncells, st));
    p.cell_pal = static_cast<uint8_t *>(w_cp.ptr);
    DP_CUDA(cudaMemsetAsync(p.sums, 0, sums_bytes, st));

    int gx = (p.npix + 255) / 256;
    if (gx > sms * 8) gx = sms * 8;
    {
        const dim3 tg(ntiles, frames);
        if (mode == 0) k_ht_sums_tile<0><<<tg, 256, 0, st>>>(p, tiles_x);
        else if (mode == 1) k_ht_sums_tile<1><<<tg, 256, 0, st>>>(p, tiles_x);
        else if (mode == 2) k_ht_sums_tile<2><<<tg, 256, 0, st>>>(p, tiles_x);
        else k_ht_sums<<<dim3(gx, frames), 256, 0, st>>>(p);
    }
    DP_LAUNCH_CHECK();
    int gc = (p.ncells + 127) / 128;
    if (gc > sms * 8) gc = sms * 8;
    k_ht_cells<<<dim3(gc, frames), 128, 0, st>>>(p);
    DP_LAUNCH_CHECK();
    const bool al16 = ((reinterpret_cast<uintptr_t>(src_rgb) | reinterpret_cast<uintptr_t>(dst_rgb) |
                        reinterpret_cast<uintptr_t>(dst_idx) | reinterpret_cast<uintptr_t>(p.screen)) & 15) == 0;
    if (al16 && p.npix % 4 == 0) {
        int g4 = (p.npix / 4 + 255) / 256;
        if (g4 > sms * 8) g4 = sms * 8;
        k_ht_select4<<<dim3(g4, frames), 256, 0, st>>>(p);
    } else {
        k_ht_select<<<dim3(gx, frames), 256, 0, st>>>(p);
    }
    DP_LAUNCH_CHECK();
    return 0;
}
