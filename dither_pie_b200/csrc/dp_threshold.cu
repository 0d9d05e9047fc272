// dp_threshold.cu -- ordered / threshold family and nearest-colour quantisation.
//
// Replaces NoDitherStrategy.dither (dithering_lib.py:333-341), MatrixDitherStrategy.dither
// (:355-378; Bayer :402-448, blue noise :451-499), InterleavedGradientNoiseDitherStrategy.dither
// (:539-568) and PolkaDotDitherStrategy.dither (:745-766), optionally fused with
// pixelize_regular (video_processor.py:563-577) in front and the integer up-scale
// (video_processor.py:393-420) behind.
//
// Data layout: interleaved u8 RGB, frames contiguous.  The identity-geometry kernel streams
// 4096-pixel tiles (12 KB in, 12 KB out) through shared memory with 128-bit global accesses;
// the palette coefficients, threshold matrix, gamma LUT and output colours live in shared
// memory for the lifetime of the (persistent) block.
// Algorithmic bytes: 3 read + 3 written per pixel; HBM-bound by design.
#include <stdlib.h>

#include "dp_search.cuh"

namespace {

struct FastDiv {
    uint32_t mul, shr, d;
};

FastDiv make_fastdiv(uint32_t d)
{
    FastDiv f;
    f.d = d;
    if (d <= 1) {
        f.mul = 0;
        f.shr = 0;
        f.d = 1;
        return f;
    }
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;
    uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}

__device__ __forceinline__ uint32_t fd_div(const FastDiv f, uint32_t n)
{
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shr);  // n < 2^31
}

constexpr int TILE_PX = 4096;
constexpr int THREADS = 256;
constexpr int TILE_BYTES = TILE_PX * 3;        // 12288, a multiple of 16
constexpr int TILE_BUF = TILE_BYTES + 32;      // room for the 16-byte misalignment shift

struct ThreshParams {
    const PalDev *P;
    const uint8_t *src;
    uint8_t *dst;
    uint8_t *dst_idx;
    int frames, h, w, npix;
    const float *matrix;
    int mh, mw;
    FastDiv dw, dmh, dmw;
    float ign_xoff, ign_yoff, ign_scale;
    int tiles_per_frame;
    int total_tiles;
    int K, integral, has_lut;
    // geometry kernel only
    int src_h, src_w, upscale;
    const int *ytab, *xtab;
    int fast;        // 0 generic kernel, 1 fast kernel with the candidate table in shared, 2 in global
    int thr_cells;
    int wm;          // v4: width of the widened threshold matrix in shared memory (multiple of 16)
    int sub_bytes;   // v4: shared bytes of the sub-cell table (multiple of 16)
    FastDiv dwm, dnpix;
    FastDiv dspr, dh;   // geom2: strips per row, rows per frame
    int geom_table;     // geom2: the compact candidate tables exist (K <= 30, integral, no LUT)
    int wide;           // v4: the 31..256-colour table format (PalDev::thr4_wide)
    int defer;          // v4 wide thresholds: flagged pixels are listed per warp and fixed 32 at a time
};

// IGN threshold (:541-549): f32, one rounding per numpy ufunc, no contraction.
__device__ __forceinline__ float ign_threshold(const ThreshParams &p, int x, int y)
{
    float xv = __fmul_rn(__fadd_rn((float)x, p.ign_xoff), p.ign_scale);
    float yv = __fmul_rn(__fadd_rn((float)y, p.ign_yoff), p.ign_scale);
    float t = __fadd_rn(__fmul_rn(xv, 0.06711056f), __fmul_rn(yv, 0.00583715f));
    t = __fsub_rn(t, floorf(t));
    float u = __fmul_rn(t, 52.9829189f);
    return __fsub_rn(u, floorf(u));
}

template <int KIND>
__device__ __forceinline__ float threshold_at(const ThreshParams &p, const float *s_mat,
                                              bool mat_in_smem, int x, int y)
{
    if (KIND == DP_THRESH_MATRIX) {
        uint32_t ym = (uint32_t)y - fd_div(p.dmh, (uint32_t)y) * p.mh;
        uint32_t xm = (uint32_t)x - fd_div(p.dmw, (uint32_t)x) * p.mw;
        uint32_t o = ym * p.mw + xm;
        return mat_in_smem ? s_mat[o] : __ldg(p.matrix + o);
    } else if (KIND == DP_THRESH_IGN) {
        return ign_threshold(p, x, y);
    }
    return 0.0f;
}

// ---- integer palette: exact search + decision ------------------------------------------
template <int KIND>
__device__ __forceinline__ int pick_int(const PalDev *P, const int4 *s_coef, int K, int r,
                                        int g, int b, float thr)
{
    if (K == 1) return 0;
    Top3 t;
    top3_init(t);
#pragma unroll 4
    for (int i = 0; i < K; ++i) top3_push(t, key_of(s_coef[i], r, g, b));
    int i1 = t.m1 & 255, i2 = t.m2 & 255;
    int s1 = t.m1 >> 8, s2 = t.m2 >> 8, s3 = t.m3 >> 8;
    bool amb = (s1 == s2);
    if (KIND != DP_THRESH_NONE) amb = amb || (K >= 3 && s2 == s3);
    if (amb) {
        int oi[2];
        if (KIND == DP_THRESH_NONE)
            tie_answer<1>(P, r, g, b, oi);
        else
            tie_answer<2>(P, r, g, b, oi);
        i1 = oi[0];
        if (KIND != DP_THRESH_NONE) i2 = oi[1];
        // the multiset of distances is unchanged: (s1, s2) stay valid
    }
    if (KIND == DP_THRESH_NONE) return i1;
    int vv = r * r + g * g + b * b;
    return factor_le_int(s1 + vv, s2 + vv, thr) ? i1 : i2;
}

// ---- general palette (gamma: non-integral f32 rows): f64 search + f64 decision -----------
template <int KIND>
__device__ __noinline__ int pick_f64(const PalDev *P, int K, int r, int g, int b, float thr)
{
    double x0 = (double)r, x1 = (double)g, x2 = (double)b;
    double s1 = DP_INF_F64, s2 = DP_INF_F64, s3 = DP_INF_F64;
    int i1 = K, i2 = K;
    for (int i = 0; i < K; ++i) {
        const double *pp = P->pal_f64 + 3 * i;
        double d0 = __dsub_rn(pp[0], x0), d1 = __dsub_rn(pp[1], x1), d2 = __dsub_rn(pp[2], x2);
        double s = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(d0, d0)), __dmul_rn(d1, d1)),
                             __dmul_rn(d2, d2));
        if (s < s1) {
            s3 = s2;
            s2 = s1;
            i2 = i1;
            s1 = s;
            i1 = i;
        } else if (s < s2) {
            s3 = s2;
            s2 = s;
            i2 = i;
        } else if (s < s3) {
            s3 = s;
        }
    }
    bool amb = (K >= 2 && s1 == s2);
    if (KIND != DP_THRESH_NONE) amb = amb || (K >= 3 && s2 == s3);
    if (amb) {
        int oi[2];
        double os[2];
        if (KIND == DP_THRESH_NONE)
            kd_emulate<1>(P, x0, x1, x2, oi, os);
        else
            kd_emulate<2>(P, x0, x1, x2, oi, os);
        i1 = oi[0];
        if (KIND != DP_THRESH_NONE) i2 = oi[1];
    }
    if (KIND == DP_THRESH_NONE || K == 1) return i1;
    return factor_le_f64(s1, s2, thr) ? i1 : min(i2, K - 1);
}

// ---------------------------------------------------------------------------------------
// Identity geometry: persistent blocks, 4096-pixel tiles staged through shared memory.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(THREADS) k_thresh_tile(const ThreshParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_in = smem;                       // TILE_BUF
    uint8_t *s_out = s_in + TILE_BUF;           // TILE_BUF
    uint8_t *s_idx = s_out + TILE_BUF;          // TILE_PX
    uint8_t *s_lut = s_idx + TILE_PX;           // 256
    uint8_t *s_orgb = s_lut + 256;              // 1024
    int4 *s_coef = reinterpret_cast<int4 *>(s_orgb + 1024);  // K
    float *s_mat = reinterpret_cast<float *>(s_coef + p.K);  // <= 1024 floats

    const PalDev *P = p.P;
    const int tid = threadIdx.x;
    const int K = p.K;
    const bool mat_in_smem = (KIND == DP_THRESH_MATRIX) && (p.mh * p.mw <= 1024);

    s_lut[tid] = P->in_lut[tid];
    for (int i = tid; i < K * 4; i += THREADS) s_orgb[i] = P->out_rgb[i];
    if (p.integral)
        for (int i = tid; i < K; i += THREADS) s_coef[i] = P->coef[i];
    if (mat_in_smem)
        for (int i = tid; i < p.mh * p.mw; i += THREADS) s_mat[i] = p.matrix[i];
    __syncthreads();

    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int f = tile / p.tiles_per_frame;
        const int tin = tile - f * p.tiles_per_frame;
        const int px0 = tin * TILE_PX;
        const int npx = min(TILE_PX, p.npix - px0);
        const int nbytes = npx * 3;
        const size_t goff = ((size_t)f * p.npix + px0) * 3;

        // ---- load: aligned 128-bit reads; byte b of the tile lands at s_in[mis_in + b]
        const uint8_t *gsrc = p.src + goff;
        const int mis_in = (int)(reinterpret_cast<uintptr_t>(gsrc) & 15);
        {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(gsrc - mis_in);
            const int n16 = (mis_in + nbytes + 15) >> 4;
            uint4 *s4 = reinterpret_cast<uint4 *>(s_in);
            for (int i = tid; i < n16; i += THREADS) s4[i] = __ldcs(g4 + i);
        }
        uint8_t *gdst = p.dst + goff;
        const int mis_out = (int)(reinterpret_cast<uintptr_t>(gdst) & 15);
        __syncthreads();

        // ---- compute: thread t owns pixels t, t+256, ... (conflict-free byte accesses)
#pragma unroll 1
        for (int j = tid; j < npx; j += THREADS) {
            const uint8_t *q = s_in + mis_in + 3 * j;
            int r = q[0], g = q[1], b = q[2];
            if (p.has_lut) {
                r = s_lut[r];
                g = s_lut[g];
                b = s_lut[b];
            }
            float thr = 0.0f;
            if (KIND != DP_THRESH_NONE) {
                uint32_t pi = (uint32_t)(px0 + j);
                uint32_t y = fd_div(p.dw, pi);
                uint32_t x = pi - y * p.w;
                thr = threshold_at<KIND>(p, s_mat, mat_in_smem, (int)x, (int)y);
            }
            int idx = p.integral ? pick_int<KIND>(P, s_coef, K, r, g, b, thr)
                                 : pick_f64<KIND>(P, K, r, g, b, thr);
            uint8_t *o = s_out + mis_out + 3 * j;
            o[0] = s_orgb[4 * idx];
            o[1] = s_orgb[4 * idx + 1];
            o[2] = s_orgb[4 * idx + 2];
            s_idx[j] = (uint8_t)idx;
        }
        __syncthreads();

        // ---- store: 128-bit writes for the aligned interior, bytes for the fringes
        {
            if (p.dst) {   // (null: index-plane-only output)
                const int head = (16 - mis_out) & 15;           // bytes before the first aligned word
                const int hb = min(head, nbytes);
                if (tid < hb) gdst[tid] = s_out[mis_out + tid];
                const int nmid = (nbytes - hb) >> 4;
                uint4 *g4 = reinterpret_cast<uint4 *>(gdst + hb);
                const uint4 *s4 = reinterpret_cast<const uint4 *>(s_out + mis_out + hb);
                for (int i = tid; i < nmid; i += THREADS) __stcs(g4 + i, s4[i]);
                const int tail0 = hb + (nmid << 4);
                if (tid < nbytes - tail0) gdst[tail0 + tid] = s_out[mis_out + tail0 + tid];
            }
            if (p.dst_idx) {
                uint8_t *gi = p.dst_idx + (size_t)f * p.npix + px0;
                for (int i = tid; i < npx; i += THREADS) gi[i] = s_idx[i];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Fast identity-geometry kernel for integral palettes (the common case: no gamma).
//
// Per pixel the exact top-3 is taken over the CANDIDATES of the pixel's colour cell only -- the
// palette rows that are nearest or second nearest (ties included) for at least one colour of
// the cell, found by exhaustive enumeration at palette creation (k_thr_masks).  For 16 colours
// the lists hold 2.3-2.7 rows on average, so the O(K) scan becomes ~3 distance evaluations:
//   key_i = (|p_i|^2 << 8 | i) - 512 * dp4a(v, p_i)      (one LDS.64, one IDP4A, one IMAD)
// Work is organised per WARP: a warp streams 512-pixel tiles (1536 B) through its private
// shared-memory buffer with 128-bit coalesced loads/stores; lanes then own groups of four
// pixels (three aligned words, conflict-free stride-3 access) and write the result in place.
// No block-wide barrier in the loop.  Rare cases (exact distance ties, threshold equality) go
// to out-of-line slow paths.
// ---------------------------------------------------------------------------------------
struct FastCtx {
    const uint2 *table;   // shared or global
    const uint8_t *ovf;   // global
    const int2 *ent;      // shared: (packed rgb, |p|^2<<8 | i)
    int shift, ncell;     // cell = ((r>>shift)*ncell + (g>>shift))*ncell + (b>>shift)
};

constexpr int WTILE_PX = 512;
constexpr int WTILE_BYTES = WTILE_PX * 3;   // 1536 = 96 x 16

__device__ __forceinline__ int fast_key(const FastCtx &c, unsigned v, unsigned i)
{
    const int2 e = c.ent[i];
    return e.y - 512 * (int)__dp4a(v, (unsigned)e.x, 0u);
}

// exact ties between candidate distances: replay scipy, then decide (out of line, rare)
template <int KIND>
__device__ __noinline__ int pick_tie(const PalDev *P, unsigned v, float thr)
{
    const int r = v & 255u, g = (v >> 8) & 255u, b = (v >> 16) & 255u;
    int oi[2];
    if (KIND == DP_THRESH_NONE) {
        tie_answer<1>(P, r, g, b, oi);
        return oi[0];
    }
    tie_answer<2>(P, r, g, b, oi);
    // the two smallest distances as a multiset do not depend on the tie order
    const int4 c0 = P->coef[oi[0]], c1 = P->coef[oi[1]];
    const int vv = r * r + g * g + b * b;
    const int n0 = (key_of(c0, r, g, b) >> 8) + vv, n1 = (key_of(c1, r, g, b) >> 8) + vv;
    return factor_le_int(min(n0, n1), max(n0, n1), thr) ? oi[0] : oi[1];
}

__device__ __noinline__ bool factor_le_f64_slow(int n1, int n2, float thr)
{
    return factor_le_f64((double)n1, (double)n2, thr);
}

// same decision as factor_le_int, with the (very rare) f64 sequence kept out of line
__device__ __forceinline__ bool factor_le_fast(int n1, int n2, float thr)
{
    const float N = (float)(n1 + n2);
    const float p = __fmul_rn(thr, N);
    const float e = __fmaf_rn(thr, N, -p);
    const float d = __fsub_rn((float)n1, p);
    if (n1 == 0) return true;
    if (thr >= 1e-6f && d != e) return d < e;
    return factor_le_f64_slow(n1, n2, thr);
}

template <int KIND>
__device__ __forceinline__ int pick_fast(const PalDev *P, const FastCtx &c, int K, unsigned v,
                                         float thr)
{
    const unsigned r = v & 255u, g = (v >> 8) & 255u, b = v >> 16;
    const unsigned cell = ((r >> c.shift) * c.ncell + (g >> c.shift)) * c.ncell + (b >> c.shift);
    const uint2 e = c.table[cell];
    const unsigned n = e.x & 255u;
    int m1, m2, m3 = 0x7fffffff;
    if (n != 255u) {
        const int k0 = fast_key(c, v, (e.x >> 8) & 255u);
        const int k1 = fast_key(c, v, (e.x >> 16) & 255u);
        m1 = min(k0, k1);
        m2 = max(k0, k1);
        if (n > 2) {
            const int k2 = fast_key(c, v, e.x >> 24);
            const int a = max(m1, k2);
            m1 = min(m1, k2);
            m3 = max(m2, a);
            m2 = min(m2, a);
            unsigned rest = e.y;
            for (unsigned j = 3; j < n; ++j, rest >>= 8) {
                const int k = fast_key(c, v, rest & 255u);
                const int a2 = max(m1, k);
                m1 = min(m1, k);
                const int b2 = max(m2, a2);
                m2 = min(m2, a2);
                m3 = min(m3, b2);
            }
        }
    } else {
        const unsigned cnt = e.x >> 8;
        const uint8_t *lst = c.ovf + e.y;
        m1 = m2 = 0x7fffffff;
        for (unsigned j = 0; j < cnt; ++j) {
            const int k = fast_key(c, v, __ldg(lst + j));
            const int a2 = max(m1, k);
            m1 = min(m1, k);
            const int b2 = max(m2, a2);
            m2 = min(m2, a2);
            m3 = min(m3, b2);
        }
    }
    const int s1 = m1 >> 8, s2 = m2 >> 8, s3 = m3 >> 8;
    bool amb = (s1 == s2);
    if (KIND != DP_THRESH_NONE) amb = amb || (s2 == s3 && K >= 3);
    if (amb) return pick_tie<KIND>(P, v, thr);
    if (KIND == DP_THRESH_NONE) return m1 & 255;
    const int vv = (int)__dp4a(v, v, 0u);
    return factor_le_fast(s1 + vv, s2 + vv, thr) ? (m1 & 255) : (m2 & 255);
}

template <int KIND, bool TABLE_SMEM, bool MAT_POW2>
__global__ void __launch_bounds__(THREADS) k_thresh_fast(const ThreshParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint4 *s_io = reinterpret_cast<uint4 *>(smem);                           // [warps][96]
    unsigned *s_orgb = reinterpret_cast<unsigned *>(smem + (THREADS / 32) * WTILE_BYTES);
    int2 *s_ent = reinterpret_cast<int2 *>(s_orgb + DP_MAX_COLORS);          // K
    uint2 *s_table = reinterpret_cast<uint2 *>(s_ent + p.K);                 // thr_cells or 0
    const bool mat_in_smem = (KIND == DP_THRESH_MATRIX) && (p.mh * p.mw <= 4096);
    float *s_mat = reinterpret_cast<float *>(s_table + (TABLE_SMEM ? p.thr_cells : 0));
    uint8_t *s_lut = reinterpret_cast<uint8_t *>(s_mat + (mat_in_smem ? p.mh * p.mw : 0));

    const PalDev *P = p.P;
    const int tid = threadIdx.x;
    const int K = p.K;
    s_lut[tid] = P->in_lut[tid];
    for (int i = tid; i < K; i += THREADS) {
        const uint8_t *o = P->out_rgb + 4 * i;
        s_orgb[i] = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16);
        const int4 cf = P->coef[i];
        const int pr = -cf.x >> 9, pg = -cf.y >> 9, pb = -cf.z >> 9;   // coef = -2p << 8
        s_ent[i] = make_int2(pr | (pg << 8) | (pb << 16), cf.w);
    }
    if (mat_in_smem)
        for (int i = tid; i < p.mh * p.mw; i += THREADS) s_mat[i] = p.matrix[i];
    FastCtx ctx;
    ctx.shift = P->thr_shift;
    ctx.ncell = 256 >> ctx.shift;
    ctx.ovf = P->thr_ovf;
    ctx.ent = s_ent;
    if (TABLE_SMEM) {
        const int cells = P->thr_cells;
        for (int i = tid; i < cells; i += THREADS) s_table[i] = P->thr_table[i];
        ctx.table = s_table;
    } else {
        ctx.table = P->thr_table;
    }
    __syncthreads();

    const int lane = tid & 31;
    const int wib = tid >> 5;
    uint4 *io4 = s_io + wib * (WTILE_BYTES / 16);
    unsigned *iow = reinterpret_cast<unsigned *>(io4);
    const int wpf = p.tiles_per_frame;              // warp tiles per frame
    const int nwarps = gridDim.x * (THREADS / 32);
    const unsigned mwm = (unsigned)p.mw - 1u, mhm = (unsigned)p.mh - 1u;

    for (int wt = blockIdx.x * (THREADS / 32) + wib; wt < p.total_tiles; wt += nwarps) {
        const int f = wt / wpf;
        const int tin = wt - f * wpf;
        const int px0 = tin * WTILE_PX;
        const int npx = min(WTILE_PX, p.npix - px0);
        const int n16 = (npx * 3) >> 4;                 // npix % 16 == 0 on this path
        const size_t goff = ((size_t)f * p.npix + px0) * 3;
        const uint4 *g4 = reinterpret_cast<const uint4 *>(p.src + goff);
        uint4 *d4 = reinterpret_cast<uint4 *>(p.dst + goff);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = j * 32 + lane;
            if (i < n16) io4[i] = __ldcs(g4 + i);
        }
        __syncwarp();
        const int ngroups = npx >> 2;
#pragma unroll 1
        for (int gi = lane; gi < ngroups; gi += 32) {
            // 12 bytes = 4 pixels, walked as a 96-bit shift register (keeps the loop rolled: the
            // body is the whole per-pixel path and has to stay resident in the instruction cache)
            unsigned a = iow[3 * gi], bw = iow[3 * gi + 1], cw = iow[3 * gi + 2];
            unsigned oa = 0, ob = 0, oc = 0, idx4 = 0;
            uint32_t y = 0, x = 0;
            const float *mrow = s_mat;
            if (KIND != DP_THRESH_NONE) {
                const uint32_t pi = (uint32_t)(px0 + 4 * gi);
                y = fd_div(p.dw, pi);
                x = pi - y * p.w;
                if (KIND == DP_THRESH_MATRIX) {
                    const uint32_t ym = MAT_POW2 ? (y & mhm) : (y - fd_div(p.dmh, y) * p.mh);
                    mrow = (mat_in_smem ? s_mat : p.matrix) + ym * p.mw;
                }
            }
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                unsigned v = a & 0xffffffu;
                a = __funnelshift_r(a, bw, 24);
                bw = __funnelshift_r(bw, cw, 24);
                cw >>= 24;
                if (p.has_lut)
                    v = (unsigned)s_lut[v & 255u] | ((unsigned)s_lut[(v >> 8) & 255u] << 8) |
                        ((unsigned)s_lut[v >> 16] << 16);
                float thr = 0.0f;
                if (KIND == DP_THRESH_MATRIX) {
                    const uint32_t xm = MAT_POW2 ? (x & mwm) : (x - fd_div(p.dmw, x) * p.mw);
                    thr = mat_in_smem ? mrow[xm] : __ldg(mrow + xm);
                } else if (KIND == DP_THRESH_IGN) {
                    thr = ign_threshold(p, (int)x, (int)y);
                }
                if (KIND != DP_THRESH_NONE) {
                    if (++x == (uint32_t)p.w) {   // row wrap inside the group (w % 4 != 0)
                        x = 0;
                        ++y;
                        if (KIND == DP_THRESH_MATRIX) {
                            const uint32_t ym = MAT_POW2 ? (y & mhm) : (y - fd_div(p.dmh, y) * p.mh);
                            mrow = (mat_in_smem ? s_mat : p.matrix) + ym * p.mw;
                        }
                    }
                }
                const unsigned idx = (unsigned)pick_fast<KIND>(P, ctx, K, v, thr);
                const unsigned col = s_orgb[idx];
                oa = __funnelshift_r(oa, ob, 24);
                ob = __funnelshift_r(ob, oc, 24);
                oc = (oc >> 24) | (col << 8);
                idx4 = (idx4 >> 8) | (idx << 24);
            }
            iow[3 * gi] = oa;
            iow[3 * gi + 1] = ob;
            iow[3 * gi + 2] = oc;
            if (p.dst_idx)
                reinterpret_cast<unsigned *>(p.dst_idx + (size_t)f * p.npix + px0)[gi] = idx4;
        }
        __syncwarp();
        if (p.dst) {   // (null: index-plane-only output)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i = j * 32 + lane;
                if (i < n16) __stcs(d4 + i, io4[i]);
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// v4: identity geometry, integral palette with K <= 30, image width a multiple of 16.
//
// Uniform cost per pixel, no data-dependent control flow outside the rare slow path:
//   * the 32^3 top-2 candidate table (one u32 = four row offsets per 8x8x8 colour cell, 128 KB)
//     is resident in shared memory: one random LDS.32 per pixel;
//   * the four candidate rows are always evaluated:  key = (|p|^2 << 8 | row) - 512 * dp4a(v, p)
//     (PRMT, LDS.64, IDP.4A, IMAD), pad rows lose every comparison;
//   * a 10-op min/max network yields the three smallest keys; ties in the upper 24 bits (exact
//     distance ties), cells with more than four candidates and exact threshold equality go to
//     an out-of-line exact path;
//   * the reference's f64 `factor <= T` test is the sign of ONE f32 fma: T*N - n1 with
//     n1 < N < 2^19 exact in f32; the fma rounds the exact value once, so its sign is the sign of
//     T*(n1+n2) - n1 unless that is exactly 0.  A non-zero difference is a multiple of
//     ulp(T)/2^23-ish relative 2^-43 of N, far above the 2^-50 relative error of the reference's
//     f64 sqrt/square/divide chain, so the decisions agree; exact equality replays that chain.
// A warp streams 512-pixel tiles through a private, double-buffered 1.5 KB shared buffer: TMA
// bulk copies (cp.async.bulk, one instruction per tile and direction, issued by lane 0) bring
// the next tile in while the current one is processed and take the finished tile out; a lane
// owns 16 consecutive pixels (48 bytes) in registers.
// ---------------------------------------------------------------------------------------
// 24 warps per SM for the threshold kinds (80 registers); plain quantisation needs fewer
// registers and no matrix, so 31 warps fit beside the table (30 with the wide format's row tables)
template <int KIND, bool WIDE = false>
__host__ __device__ constexpr int v4_threads()
{
    // (the skewed candidate table and the 256-row tables of the wide format leave room for 31 / 30
    // warps)
    return KIND == DP_THRESH_NONE ? (WIDE ? 960 : 992) : 768;
}
// The 32^3 cell table in shared memory, SKEWED: the entry of cell (r5, g5, b5) sits at word
// r5 * 1057 + g5 * 33 + b5, so that its bank is (r5 + g5 + b5) mod 32 instead of b5.  Neighbouring
// lanes hold similar colours: with the plain layout the gather of the cell entries was ~7-way
// bank-conflicted (ncu: 6.9 wavefronts per LDS, profiles/r2t_thresh_v4_none_K16.txt), the largest
// item on the kernel's busiest unit.  The address stays linear in the three 5-bit fields -- the
// same IMAD + IDP.4A, other constants -- and the table grows by 4 KB (33 822 words).
#define V4_TABLE_WORDS 33824                    /* 31 * 1057 + 31 * 33 + 31 + 1, rounded up to 16 bytes */
#define V4_TABLE_BYTES (V4_TABLE_WORDS * 4)

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_s32x2(unsigned a)
{
    int2 v;
    asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(unsigned a)
{
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion ---------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst_smem, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, unsigned src_smem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

struct V4Ctx {
    unsigned table_a, ent_a, orgb_a, sub_a;   // shared-space addresses
    const uint32_t *sub_g;                    // nearest-only table: sub-cell entries (global, L1)
    const uint32_t *table_g, *subt_g;         // TG variant: both table levels in global memory (L1)
    const PalDev *P;
    int K;
};

// one pixel: v = r | g<<8 | b<<16 (bits 24..31 arbitrary) -> palette row; `slow` is set when
// the exact path has to decide (see above).  The byte extractions and the table address use
// IDP.4A / IMAD so that the ALU pipe (64 lanes/clk, the bottleneck) and the IMAD pipe share
// the work.
// TG = the table is read from global memory through L1 (kernels that cannot afford 128 KB of
// shared memory for it, e.g. the fused geometry kernel) instead of from shared memory.
// WIDE = the 31..256-colour table format (PalDev::thr4_wide): plain row numbers, fillers instead of
// pad rows, sub-cell entries in global memory.
template <int KIND, bool TG = false, bool WIDE = false>
__device__ __forceinline__ unsigned v4_pick(const V4Ctx &c, unsigned v, float thr, bool &slow)
{
    constexpr unsigned AW = WIDE ? 8u : 1u;            // byte j of the entry -> row offset in the int2 array
    constexpr unsigned MARK4 = WIDE ? 0x03000000u : 0xf8000000u;
    // byte offset of the pixel's cell in the u32 table: (r>>3)*4096 + (g>>3)*128 + (b>>3)*4 in global
    // memory (TG), (r>>3)*4228 + (g>>3)*132 + (b>>3)*4 in the skewed shared copy (V4_TABLE_WORDS)
    const unsigned a5 = (v >> 3) & 0x1f1f1fu;
    const unsigned ta = TG ? (a5 & 0x1fu) * 4096u + __dp4a(a5, 0x00048000u, 0u)
                           : (a5 & 0x1fu) * 4228u + __dp4a(a5, 0x00048400u, c.table_a);
    unsigned e = TG ? __ldg(reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(c.table_g) + ta))
                    : lds_u32(ta);
    if (KIND == DP_THRESH_NONE) {
        // plain quantisation: the table holds the rows that can be NEAREST in the cell, three
        // slots; the answer is the smallest key unless the two smallest tie
        if (WIDE ? (e < 0xff000000u) : (e >= 0xf8000000u)) {
            const unsigned sc = ((v >> 2) & 1u) * 4u + ((v >> 10) & 1u) * 2u + ((v >> 18) & 1u);
            e = __ldg(c.sub_g + (e & 0xffffu) * 8u + sc);
        }
        const int2 q0 = lds_s32x2(__dp4a(e, 0x00000001u * AW, c.ent_a));
        const int2 q1 = lds_s32x2(__dp4a(e, 0x00000100u * AW, c.ent_a));
        const int2 q2 = lds_s32x2(__dp4a(e, 0x00010000u * AW, c.ent_a));
        const int k0 = q0.y - 512 * (int)__dp4a(v, (unsigned)q0.x, 0u);
        const int k1 = q1.y - 512 * (int)__dp4a(v, (unsigned)q1.x, 0u);
        const int k2 = q2.y - 512 * (int)__dp4a(v, (unsigned)q2.x, 0u);
        const int m1 = min(min(k0, k1), k2);
        const int m3 = max(max(k0, k1), k2);
        const int m2 = (int)((unsigned)k0 + (unsigned)k1 + (unsigned)k2 - (unsigned)m1 - (unsigned)m3);
        slow = (WIDE ? (e < 0xff000000u) : (e >= 0xf8000000u)) || ((unsigned)(m1 ^ m2) < 256u);
        return (unsigned)m1 & 255u;
    }
    if (WIDE ? (e < MARK4) : (e >= MARK4)) {   // more than four candidates in the 8^3 cell: refine to the 4^3 sub-cell
        const unsigned sc = ((v >> 2) & 1u) * 4u + ((v >> 10) & 1u) * 2u + ((v >> 18) & 1u);
        e = (TG || WIDE) ? __ldg(c.subt_g + (e & 0xffffu) * 8u + sc) : lds_u32(c.sub_a + ((e & 0xffffu) * 8u + sc) * 4u);
    }
    const int2 q0 = lds_s32x2(__dp4a(e, 0x00000001u * AW, c.ent_a));   // base + row offset of byte j of e
    const int2 q1 = lds_s32x2(__dp4a(e, 0x00000100u * AW, c.ent_a));
    const int2 q2 = lds_s32x2(__dp4a(e, 0x00010000u * AW, c.ent_a));
    const int2 q3 = lds_s32x2(__dp4a(e, 0x01000000u * AW, c.ent_a));
    const int k0 = q0.y - 512 * (int)__dp4a(v, (unsigned)q0.x, 0u);
    const int k1 = q1.y - 512 * (int)__dp4a(v, (unsigned)q1.x, 0u);
    const int k2 = q2.y - 512 * (int)__dp4a(v, (unsigned)q2.x, 0u);
    const int k3 = q3.y - 512 * (int)__dp4a(v, (unsigned)q3.x, 0u);
    // the three smallest of the four keys (real keys are distinct; two pads may coincide)
    const int lo01 = min(k0, k1), hi01 = max(k0, k1), lo23 = min(k2, k3), hi23 = max(k2, k3);
    const int m1 = min(lo01, lo23);
    const int x = max(lo01, lo23);
    const int m2 = min(min(x, hi01), hi23);
    slow = WIDE ? (e < MARK4) : (e >= MARK4);                   // still more than four candidates
    const int mx = max(max(x, hi01), hi23);
    // the median of {x, hi01, hi23}; modular arithmetic, the pad key 0x7fffffff may wrap
    const int m3 = (int)((unsigned)x + (unsigned)hi01 + (unsigned)hi23 - (unsigned)m2 - (unsigned)mx);
    slow = slow || (min((unsigned)(m1 ^ m2), (unsigned)(m2 ^ m3)) < 256u);
    const unsigned vm = v & 0xffffffu;
    const int vv = (int)__dp4a(vm, vm, 0u);
    if (WIDE) {
        // rows up to 255: the two row bytes may carry when the keys are added, so shift first
        const int n1w = (vv * 256 + m1) >> 8, n2w = (vv * 256 + m2) >> 8;
        const float s = __fmaf_rn(thr, __int2float_rn(n1w + n2w), -__int2float_rn(n1w));
        slow = slow || (s == 0.0f);
        return (unsigned)(s > 0.0f ? m1 : m2) & 255u;
    }
    const int n1 = (vv * 256 + m1) >> 8;                        // exact squared distances
    const int nn = (vv * 512 + m1 + m2) >> 8;                   // n1 + n2 (row bits never carry: K <= 30)
    const float s = __fmaf_rn(thr, __int2float_rn(nn), -__int2float_rn(n1));
    slow = slow || (s == 0.0f);
    return (unsigned)(s > 0.0f ? m1 : m2) & 255u;
}

// pick_int over the kernel's SHARED row table (packed rgb, |p|^2 << 8 | row): the exact top-3 of all
// K rows without touching global memory; only a genuine distance tie goes to the exception table
// (two dependent loads through its bucket index).
template <int KIND>
__device__ __forceinline__ int pick_int_ent(const PalDev *P, const int2 *s_ent, int K, unsigned v, float thr)
{
    if (K == 1) return 0;
    Top3 t;
    top3_init(t);
#pragma unroll 4
    for (int i = 0; i < K; ++i) {
        const int2 e = s_ent[i];
        top3_push(t, e.y - 512 * (int)__dp4a(v, (unsigned)e.x, 0u));
    }
    const int r = v & 255u, g = (v >> 8) & 255u, b = (v >> 16) & 255u;
    int i1 = t.m1 & 255, i2 = t.m2 & 255;
    const int s1 = t.m1 >> 8, s2 = t.m2 >> 8, s3 = t.m3 >> 8;
    bool amb = (s1 == s2);
    if (KIND != DP_THRESH_NONE) amb = amb || (K >= 3 && s2 == s3);
    if (amb) {
        int oi[2];
        if (KIND == DP_THRESH_NONE)
            tie_answer<1>(P, r, g, b, oi);
        else
            tie_answer<2>(P, r, g, b, oi);
        i1 = oi[0];
        if (KIND != DP_THRESH_NONE) i2 = oi[1];
        // the multiset of distances is unchanged: (s1, s2) stay valid
    }
    if (KIND == DP_THRESH_NONE) return i1;
    const int vv = r * r + g * g + b * b;
    return factor_le_int(s1 + vv, s2 + vv, thr) ? i1 : i2;
}

// exact decision for one pixel v4_pick flagged: most flagged pixels are exact distance ties, whose
// answer is in the exception table (two dependent loads), no scan of the rows is needed
template <int KIND, bool WIDE>
__device__ __forceinline__ int v4_exact_idx(const ThreshParams &p, const FastCtx &fc, const int2 *s_ent,
                                            unsigned r, unsigned g, unsigned b, float thr)
{
    const PalDev *P = p.P;
    const unsigned vq = r | (g << 8) | (b << 16);
    int oi[2];
    if (tie_lookup<KIND == DP_THRESH_NONE ? 1 : 2>(P, (int)r, (int)g, (int)b, oi)) {
        int idx = oi[0];
        if (KIND != DP_THRESH_NONE) {
            const int vv = (int)(r * r + g * g + b * b);
            const int2 ea = s_ent[oi[0]], eb = s_ent[oi[1]];
            const int n0 = ((ea.y - 512 * (int)__dp4a(vq, (unsigned)ea.x, 0u)) >> 8) + vv;
            const int n1 = ((eb.y - 512 * (int)__dp4a(vq, (unsigned)eb.x, 0u)) >> 8) + vv;
            // the two smallest distances as a multiset do not depend on the tie order
            idx = factor_le_int(min(n0, n1), max(n0, n1), thr) ? oi[0] : oi[1];
        }
        return idx;
    }
    return WIDE ? pick_fast<KIND>(P, fc, p.K, vq, thr) : pick_int_ent<KIND>(P, s_ent, p.K, vq, thr);
}

// exact decision for the pixels v4_pick flagged (rare): re-reads the pixel from global memory,
// patches its output bytes in the warp's staging buffer and its index byte in global memory
template <int KIND, bool WM_POW2, bool WIDE>
__device__ __forceinline__ void v4_fix(const ThreshParams &p, unsigned slowmask, long long gp, uint32_t x,
                                            uint32_t y, unsigned ra, uint8_t *out_bytes, const unsigned *s_orgb,
                                            const int2 *s_ent)
{
    const PalDev *P = p.P;
    FastCtx fc;     // WIDE: the exact pick over the cell's full candidate list (a K-row scan is too long)
    fc.table = P->thr_table;
    fc.ovf = P->thr_ovf;
    fc.ent = s_ent;
    fc.shift = P->thr_shift;
    fc.ncell = 256 >> fc.shift;
    while (slowmask) {
        const int j = __ffs(slowmask) - 1;
        slowmask &= slowmask - 1;
        const uint8_t *q = p.src + (gp + j) * 3;
        float thr = 0.0f;
        if (KIND == DP_THRESH_MATRIX) {
            float t;
            asm("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(ra + 4u * j));
            thr = t;
        } else if (KIND == DP_THRESH_IGN) {
            thr = ign_threshold(p, (int)x + j, (int)y);
        }
        const int idx = v4_exact_idx<KIND, WIDE>(p, fc, s_ent, q[0], q[1], q[2], thr);
        const unsigned col = s_orgb[idx];
        out_bytes[3 * j] = (uint8_t)col;
        out_bytes[3 * j + 1] = (uint8_t)(col >> 8);
        out_bytes[3 * j + 2] = (uint8_t)(col >> 16);
        if (p.dst_idx) p.dst_idx[gp + j] = (uint8_t)idx;
    }
}

// The wide format flags ~1 % of the pixels (distance ties and crowded sub-cells of a 256-colour
// palette), i.e. a few per 512-pixel tile: fixed tile by tile (v4_fix), almost every tile pays a
// chain of five dependent global loads for one or two pixels of one or two lanes, a third of the
// kernel's time.  Instead the warp collects the flagged pixels of its tiles (batch pixel index)
// in a small shared list and fixes them 32 at a time, one per lane, directly in global memory --
// after the bulk stores of their tiles have completed.
#define V4_DEFER_CAP 64
template <int KIND, bool WM_POW2, bool WIDE>
__device__ __forceinline__ void v4_fix_deferred(const ThreshParams &p, const uint32_t *list, int n, int lane,
                                             const float *s_mat, const unsigned *s_orgb, const int2 *s_ent)
{
    const PalDev *P = p.P;
    FastCtx fc;
    fc.table = P->thr_table;
    fc.ovf = P->thr_ovf;
    fc.ent = s_ent;
    fc.shift = P->thr_shift;
    fc.ncell = 256 >> fc.shift;
    for (int i = lane; i < n; i += 32) {
        const uint32_t gpp = list[i];
        const uint8_t *q = p.src + (size_t)gpp * 3;
        const uint32_t pin = gpp - fd_div(p.dnpix, gpp) * (uint32_t)p.npix;
        const uint32_t y = fd_div(p.dw, pin), x = pin - y * p.w;
        float thr = 0.0f;
        if (KIND == DP_THRESH_MATRIX) {
            const uint32_t ym = y - fd_div(p.dmh, y) * p.mh;
            const uint32_t xm = WM_POW2 ? (x & (uint32_t)(p.wm - 1)) : (x - fd_div(p.dwm, x) * p.wm);
            thr = s_mat[ym * p.wm + xm];
        } else if (KIND == DP_THRESH_IGN) {
            thr = ign_threshold(p, (int)x, (int)y);
        }
        const int idx = v4_exact_idx<KIND, WIDE>(p, fc, s_ent, q[0], q[1], q[2], thr);
        const unsigned col = s_orgb[idx];
        if (p.dst) {
            uint8_t *o = p.dst + (size_t)gpp * 3;
            o[0] = (uint8_t)col;
            o[1] = (uint8_t)(col >> 8);
            o[2] = (uint8_t)(col >> 16);
        }
        if (p.dst_idx) p.dst_idx[gpp] = (uint8_t)idx;
    }
}

__host__ __device__ constexpr int v4_ent_bytes(bool wide) { return wide ? 2048 : 272; }     // int2 [256] / [34]
__host__ __device__ constexpr int v4_orgb_bytes(bool wide) { return wide ? 1024 : 128; }    // u32 [256] / [32]

template <int KIND, bool WM_POW2, bool WIDE, bool DEFER>
__global__ void __launch_bounds__(v4_threads<KIND, WIDE>(), 1) k_thresh_v4(const ThreshParams p)
{
    constexpr int V4_THREADS = v4_threads<KIND, WIDE>();
    constexpr int V4_WARPS = V4_THREADS / 32;
    constexpr int ENTB = v4_ent_bytes(WIDE), ORGBB = v4_orgb_bytes(WIDE);
    extern __shared__ __align__(16) uint8_t smem[];
    const int P_nsub = WIDE ? 0 : p.P->thr4_nsub;          // WIDE: sub-cell entries stay in global memory
    uint32_t *s_table = reinterpret_cast<uint32_t *>(smem);                    // [V4_TABLE_WORDS], skewed
    int2 *s_ent = reinterpret_cast<int2 *>(smem + V4_TABLE_BYTES);             // [34] / [256]
    unsigned *s_orgb = reinterpret_cast<unsigned *>(smem + V4_TABLE_BYTES + ENTB);     // [32] / [256]
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem + V4_TABLE_BYTES + ENTB + ORGBB);   // [warps][2]
    uint4 *s_io = reinterpret_cast<uint4 *>(smem + V4_TABLE_BYTES + ENTB + ORGBB + 512);  // [warps][2][96]
    uint32_t *s_sub = reinterpret_cast<uint32_t *>(s_io + V4_WARPS * 192);     // [8 * nsub]
    float *s_mat = reinterpret_cast<float *>(s_sub + (KIND == DP_THRESH_NONE ? 0 : ((8 * P_nsub + 3) & ~3)));   // [mh][wm]
    // DEFER instantiations: per-warp list of flagged pixels (the launch found room for it)
    uint32_t *s_defer = reinterpret_cast<uint32_t *>(s_mat + (KIND == DP_THRESH_MATRIX ? p.mh * p.wm : 0));

    const PalDev *P = p.P;
    const int tid = threadIdx.x;
    const int K = p.K;
    {
        const uint4 *src4 = reinterpret_cast<const uint4 *>(KIND == DP_THRESH_NONE ? P->near3_table
                                                                                   : P->thr4_table);
        // four consecutive cells (b5 = 4k .. 4k+3 of one (r5, g5)) per load, scattered to the skewed
        // positions r5 * 1057 + g5 * 33 + b5
        for (int i = tid; i < 8192; i += V4_THREADS) {
            const uint4 e4 = __ldg(src4 + i);
            const int r5 = i >> 8, g5 = (i >> 3) & 31, b5 = (i & 7) << 2;
            uint32_t *d = s_table + r5 * 1057 + g5 * 33 + b5;
            d[0] = e4.x; d[1] = e4.y; d[2] = e4.z; d[3] = e4.w;
        }
        if (KIND != DP_THRESH_NONE)
            for (int i = tid; i < 8 * P_nsub; i += V4_THREADS) s_sub[i] = __ldg(P->thr4_sub + i);
    }
    if (tid < (WIDE ? 256 : 34)) {
        int2 en = make_int2(0, 0x7fffff00 | 255);       // pad rows never win
        if (tid < K) {
            const int4 cf = P->coef[tid];
            const int pr = -cf.x >> 9, pg = -cf.y >> 9, pb = -cf.z >> 9;   // coef = -2p << 8
            en = make_int2(pr | (pg << 8) | (pb << 16), cf.w);
        }
        s_ent[tid] = en;
    }
    if (tid < (WIDE ? 256 : 32)) {
        unsigned col = 0;
        if (tid < K) {
            const uint8_t *o = P->out_rgb + 4 * tid;
            col = (unsigned)o[0] | ((unsigned)o[1] << 8) | ((unsigned)o[2] << 16);
        }
        s_orgb[tid] = col;
    }
    if (KIND == DP_THRESH_MATRIX) {
        // widened copy: row r holds the matrix row repeated up to a multiple of 16 columns
        const int n = p.mh * p.wm;
        for (int i = tid; i < n; i += V4_THREADS) {
            const int r = i / p.wm, cc = i - r * p.wm;
            s_mat[i] = p.matrix[r * p.mw + cc % p.mw];
        }
    }
    __syncthreads();

    V4Ctx ctx;
    ctx.table_a = smem_u32(s_table);
    ctx.ent_a = smem_u32(s_ent);
    ctx.orgb_a = smem_u32(s_orgb);
    ctx.sub_a = smem_u32(s_sub);
    ctx.sub_g = P->near3_sub;
    ctx.subt_g = P->thr4_sub;
    ctx.table_g = nullptr;
    ctx.P = P;
    ctx.K = K;
    const unsigned mat_a = smem_u32(s_mat);

    const int lane = tid & 31;
    const int wib = tid >> 5;
    uint4 *io4 = s_io + wib * 192;
    // 32-bit indexing: the host splits batches so that frames * npix < 2^31
    const uint32_t total_px = (uint32_t)p.frames * (uint32_t)p.npix;  // frames are contiguous
    const uint32_t ntiles = (total_px + 511u) >> 9;
    const uint32_t wstride = gridDim.x * V4_WARPS;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(p.src);
    uint4 *d4 = reinterpret_cast<uint4 *>(p.dst);
    const uint32_t n16_total = (total_px >> 4) * 3u;                  // 16-byte units in the batch

    // tile `t` -> staging buffer `buf` of this warp: ONE TMA bulk copy issued by lane 0, completion
    // on the buffer's mbarrier.  The buffer was the source of a bulk store two tiles ago; lane 0
    // (which committed that store) first waits until the store has finished reading it.
    const unsigned bar_a = smem_u32(s_bar + 2 * wib);
    const unsigned io_a = smem_u32(io4);
    if (lane == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_a + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](uint32_t t, int buf) {
        if (lane == 0) {
            const uint32_t b16 = t * 96u;
            const uint32_t n16 = min(96u, n16_total - b16);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_expect_tx(bar_a + 8 * buf, n16 * 16u);
            bulk_load(io_a + buf * 1536, g4 + b16, n16 * 16u, bar_a + 8 * buf);
        }
    };
    uint32_t wt = blockIdx.x * V4_WARPS + wib;
    int buf = 0;
    unsigned it = 0;    // tiles done by this warp: buffer it & 1, barrier parity (it >> 1) & 1
    uint32_t *dlist = s_defer + wib * V4_DEFER_CAP;
    int ndefer = 0;     // warp-uniform
    if (wt < ntiles) issue(wt, 0);
    // position of this lane's first pixel inside its frame, advanced incrementally
    uint32_t pin = 0, pstep = 0;
    if (KIND != DP_THRESH_NONE) {
        const uint32_t gp0 = (wt << 9) + 16u * lane, adv = wstride << 9;
        pin = gp0 - fd_div(p.dnpix, gp0) * (uint32_t)p.npix;
        pstep = adv - fd_div(p.dnpix, adv) * (uint32_t)p.npix;
    }
    for (; wt < ntiles; wt += wstride, buf ^= 1, ++it) {
        const uint32_t b16 = wt * 96u;
        if (wt + wstride < ntiles) issue(wt + wstride, buf ^ 1);
        mbar_wait(bar_a + 8 * buf, (it >> 1) & 1u);
        uint4 *cur = io4 + buf * 96;
        // this lane's 16 pixels
        unsigned w[12];
        {
            const uint4 a = cur[3 * lane], b = cur[3 * lane + 1], cq = cur[3 * lane + 2];
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
            w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            w[8] = cq.x; w[9] = cq.y; w[10] = cq.z; w[11] = cq.w;
        }
        const uint32_t gp = (wt << 9) + 16u * lane;                  // first pixel (batch index)
        const bool live = gp < total_px;
        uint32_t x = 0, y = 0;
        if (KIND != DP_THRESH_NONE) {
            y = fd_div(p.dw, pin);
            x = pin - y * p.w;                                       // multiple of 16
            pin += pstep;
            if (pin >= (uint32_t)p.npix) pin -= (uint32_t)p.npix;
        }
        unsigned ra = 0;   // shared address of this lane's 16 thresholds
        if (KIND == DP_THRESH_MATRIX) {
            const uint32_t ym = y - fd_div(p.dmh, y) * p.mh;
            const uint32_t xm = WM_POW2 ? (x & (uint32_t)(p.wm - 1)) : (x - fd_div(p.dwm, x) * p.wm);
            ra = mat_a + 4u * (ym * p.wm + xm);
        }
        unsigned idxw[4] = {0, 0, 0, 0};
        unsigned slowmask = 0;
        if (live) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                const unsigned a = w[3 * gq], b = w[3 * gq + 1], cq = w[3 * gq + 2];
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (KIND == DP_THRESH_MATRIX) {
                    t = lds_f32x4(ra + 16u * gq);
                } else if (KIND == DP_THRESH_IGN) {
                    t.x = ign_threshold(p, (int)x + 4 * gq, (int)y);
                    t.y = ign_threshold(p, (int)x + 4 * gq + 1, (int)y);
                    t.z = ign_threshold(p, (int)x + 4 * gq + 2, (int)y);
                    t.w = ign_threshold(p, (int)x + 4 * gq + 3, (int)y);
                }
                bool s0, s1, s2, s3;
                const unsigned i0 = v4_pick<KIND, false, WIDE>(ctx, a, t.x, s0);
                const unsigned i1 = v4_pick<KIND, false, WIDE>(ctx, __funnelshift_r(a, b, 24), t.y, s1);
                const unsigned i2 = v4_pick<KIND, false, WIDE>(ctx, __funnelshift_r(b, cq, 16), t.z, s2);
                const unsigned i3 = v4_pick<KIND, false, WIDE>(ctx, cq >> 8, t.w, s3);
                slowmask |= ((s0 ? 1u : 0u) | (s1 ? 2u : 0u) | (s2 ? 4u : 0u) | (s3 ? 8u : 0u)) << (4 * gq);
                const unsigned c0 = lds_u32(ctx.orgb_a + 4u * i0), c1 = lds_u32(ctx.orgb_a + 4u * i1),
                               c2 = lds_u32(ctx.orgb_a + 4u * i2), c3 = lds_u32(ctx.orgb_a + 4u * i3);
                w[3 * gq] = c0 | (c1 << 24);
                w[3 * gq + 1] = (c1 >> 8) | (c2 << 16);
                w[3 * gq + 2] = (c2 >> 16) | (c3 << 8);
                idxw[gq] = i0 | (i1 << 8) | (i2 << 16) | (i3 << 24);
            }
            if (p.dst_idx)
                *reinterpret_cast<uint4 *>(p.dst_idx + gp) = make_uint4(idxw[0], idxw[1], idxw[2], idxw[3]);
        }
        cur[3 * lane] = make_uint4(w[0], w[1], w[2], w[3]);      // own slots: no hazard with other lanes
        cur[3 * lane + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        cur[3 * lane + 2] = make_uint4(w[8], w[9], w[10], w[11]);
        if (DEFER) {
            // (inline on purpose, and so are the fix functions it uses: any call in the tile loop
            // makes ptxas keep the loop's state on the stack -- five exposed local-memory loads per
            // tile, 10 % of the stall samples.  IGN K=256: everything behind one call 0.66 ms, the
            // fix functions called 0.554 ms, all inline 0.520 ms)
            if (__ballot_sync(0xffffffffu, slowmask != 0)) {
                const int cnt = __popc(slowmask);
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                if (total > V4_DEFER_CAP) {     // pathological tile: fixed in place, lane by lane
                    if (slowmask)
                        v4_fix<KIND, WM_POW2, WIDE>(p, slowmask, gp, x, y, ra,
                                                    reinterpret_cast<uint8_t *>(cur + 3 * lane), s_orgb, s_ent);
                } else {
                    if (ndefer + total > V4_DEFER_CAP) {
                        // every listed pixel belongs to an earlier tile: wait until the bulk stores
                        // of those tiles have completed (this tile's store is not issued yet), then
                        // patch them in global memory
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        __syncwarp();
                        // the patches are generic-proxy stores to bytes the async proxy wrote
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        v4_fix_deferred<KIND, WM_POW2, WIDE>(p, dlist, ndefer, lane, s_mat, s_orgb, s_ent);
                        __syncwarp();
                        ndefer = 0;
                    }
                    int pos = ndefer + incl - cnt;
                    unsigned m = slowmask;
                    while (m) {
                        dlist[pos++] = gp + (uint32_t)(__ffs(m) - 1);
                        m &= m - 1;
                    }
                    ndefer += total;
                }
            }
        } else if (slowmask) {
            // (inline: a call in the tile loop makes ptxas keep the loop's state on the stack and
            // costs more than the fix's registers -- measured for every kind and both formats:
            // `none` K=16 0.296 -> 0.277 ms, K=256 0.408 -> 0.389 ms, Bayer K=16 0.365 -> 0.359 ms)
            v4_fix<KIND, WM_POW2, WIDE>(p, slowmask, gp, x, y, ra, reinterpret_cast<uint8_t *>(cur + 3 * lane),
                                        s_orgb, s_ent);
        }
        // generic-proxy writes -> visible to the async proxy, then one bulk store by lane 0
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && p.dst) {   // (null dst: index-plane-only output)
            const uint32_t n16 = min(96u, n16_total - b16);
            bulk_store(d4 + b16, io_a + buf * 1536, n16 * 16u);
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (DEFER) {
        __syncwarp();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        v4_fix_deferred<KIND, WM_POW2, WIDE>(p, dlist, ndefer, lane, s_mat, s_orgb, s_ent);
    }
}

// ---------------------------------------------------------------------------------------
// Fused geometry: gather (pixelize) -> dither -> m x m block store (up-scale).
// One thread per dithered pixel.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(THREADS) k_thresh_geom(const ThreshParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_lut = smem;
    uint8_t *s_orgb = s_lut + 256;
    int4 *s_coef = reinterpret_cast<int4 *>(s_orgb + 1024);
    float *s_mat = reinterpret_cast<float *>(s_coef + p.K);

    const PalDev *P = p.P;
    const int tid = threadIdx.x;
    const int K = p.K;
    const bool mat_in_smem = (KIND == DP_THRESH_MATRIX) && (p.mh * p.mw <= 1024);
    s_lut[tid] = P->in_lut[tid];
    for (int i = tid; i < K * 4; i += THREADS) s_orgb[i] = P->out_rgb[i];
    if (p.integral)
        for (int i = tid; i < K; i += THREADS) s_coef[i] = P->coef[i];
    if (mat_in_smem)
        for (int i = tid; i < p.mh * p.mw; i += THREADS) s_mat[i] = p.matrix[i];
    __syncthreads();

    const int m = p.upscale;
    const size_t src_frame = (size_t)p.src_h * p.src_w * 3;
    const size_t out_w = (size_t)p.w * m;
    const size_t dst_frame = (size_t)p.h * m * out_w * 3;
    const long long total = (long long)p.frames * p.npix;
    for (long long gi = (long long)blockIdx.x * THREADS + tid; gi < total;
         gi += (long long)gridDim.x * THREADS) {
        const int f = (int)(gi / p.npix);
        const uint32_t pi = (uint32_t)(gi - (long long)f * p.npix);
        const uint32_t y = fd_div(p.dw, pi);
        const uint32_t x = pi - y * p.w;
        const int sy = p.ytab ? __ldg(p.ytab + y) : (int)y;
        const int sx = p.xtab ? __ldg(p.xtab + x) : (int)x;
        const uint8_t *q = p.src + (size_t)f * src_frame + ((size_t)sy * p.src_w + sx) * 3;
        int r = q[0], g = q[1], b = q[2];
        if (p.has_lut) {
            r = s_lut[r];
            g = s_lut[g];
            b = s_lut[b];
        }
        float thr = threshold_at<KIND>(p, s_mat, mat_in_smem, (int)x, (int)y);
        int idx = p.integral ? pick_int<KIND>(P, s_coef, K, r, g, b, thr)
                             : pick_f64<KIND>(P, K, r, g, b, thr);
        const uint8_t o0 = s_orgb[4 * idx], o1 = s_orgb[4 * idx + 1], o2 = s_orgb[4 * idx + 2];
        uint8_t *d = p.dst + (size_t)f * dst_frame + ((size_t)y * m * out_w + (size_t)x * m) * 3;
        for (int yy = 0; p.dst && yy < m; ++yy) {   // (null dst: index-plane-only output)
            uint8_t *row = d + (size_t)yy * out_w * 3;
            for (int xx = 0; xx < m; ++xx) {
                row[3 * xx] = o0;
                row[3 * xx + 1] = o1;
                row[3 * xx + 2] = o2;
            }
        }
        if (p.dst_idx) p.dst_idx[gi] = (uint8_t)idx;
    }
}

// ---------------------------------------------------------------------------------------
// Fused geometry, coalesced stores (config 4: pixelize -> dither -> x m up-scale).  The work is
// dominated by the m*m-times larger output, so the kernel is organised around the stores: a warp
// takes a strip of 32 consecutive dithered pixels of one row, every lane replicates its colour m
// times into a shared row image (96*m bytes), and the warp writes that image to each of the m
// output rows with 128-bit stores.  The lane's source pixel is fetched one strip ahead.
// Needs 16-byte aligned output rows; otherwise k_thresh_geom.
// The bound of this path is the WRITE-ONLY bandwidth of HBM, 3.9 TB/s on B200 (memset/fill of
// 0.4-1.6 GB, tools/ubench/write_bw.py) against 6.5 TB/s for a copy: 398 MB of output per 64
// 1080p frames cannot take less than 0.10 ms.  Measured alternatives: three staged kernels
// (coalesced gather 0.036 ms, k_thresh_v4 on the low-resolution frames 0.053 ms, a pure
// replication kernel 0.098 ms = 4.06 TB/s, i.e. at the write roofline) sum to the same 0.187 ms
// as this fused kernel (whose dither work overlaps its stores); capping it at 64 registers for a
// fourth resident block did not help either.
// ---------------------------------------------------------------------------------------
constexpr int GEOM2_MAX_M = 8;
constexpr int GEOM2_MAT_SMEM = 4096;   // threshold matrices up to 64 x 64 (blue noise) live in shared memory

template <int KIND>
__global__ void __launch_bounds__(THREADS) k_thresh_geom2(const ThreshParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_lut = smem;
    uint8_t *s_orgb = s_lut + 256;
    int4 *s_coef = reinterpret_cast<int4 *>(s_orgb + 1024);
    float *s_mat = reinterpret_cast<float *>(s_coef + p.K);
    const bool mat_in_smem = (KIND == DP_THRESH_MATRIX) && (p.mh * p.mw <= GEOM2_MAT_SMEM);
    uint8_t *s_row = reinterpret_cast<uint8_t *>(s_mat + (mat_in_smem ? p.mh * p.mw : 0));
    s_row += (16 - (reinterpret_cast<uintptr_t>(s_row) & 15)) & 15;

    const PalDev *P = p.P;
    const int tid = threadIdx.x;
    const int K = p.K;
    s_lut[tid] = P->in_lut[tid];
    for (int i = tid; i < K * 4; i += THREADS) s_orgb[i] = P->out_rgb[i];
    if (p.integral)
        for (int i = tid; i < K; i += THREADS) s_coef[i] = P->coef[i];
    if (mat_in_smem)
        for (int i = tid; i < p.mh * p.mw; i += THREADS) s_mat[i] = p.matrix[i];
    __syncthreads();

    const int lane = tid & 31, wib = tid >> 5;
    const int m = p.upscale;
    uint8_t *rowimg = s_row + wib * (96 * GEOM2_MAX_M);       // this warp's row image
    // candidate-table pick (K <= 30, integral palette): the 32^3 table through L1, rows in shared
    __shared__ int2 s_ent[34];
    V4Ctx ctx;
    const bool use_table = p.geom_table != 0;
    if (use_table) {
        if (tid < 34) {
            int2 en = make_int2(0, 0x7fffff00 | 255);
            if (tid < K) {
                const int4 cf = P->coef[tid];
                const int pr = -cf.x >> 9, pg = -cf.y >> 9, pb = -cf.z >> 9;
                en = make_int2(pr | (pg << 8) | (pb << 16), cf.w);
            }
            s_ent[tid] = en;
        }
        __syncthreads();
        ctx.ent_a = smem_u32(s_ent);
        ctx.table_a = ctx.orgb_a = ctx.sub_a = 0;
        ctx.table_g = KIND == DP_THRESH_NONE ? P->near3_table : P->thr4_table;
        ctx.subt_g = P->thr4_sub;
        ctx.sub_g = P->near3_sub;
        ctx.P = P;
        ctx.K = K;
    }
    const int strips_per_row = (p.w + 31) >> 5;
    const uint32_t total = (uint32_t)p.frames * p.h * strips_per_row;     // < 2^31 (host check)
    const size_t src_frame = (size_t)p.src_h * p.src_w * 3;
    const size_t out_w3 = (size_t)p.w * m * 3;
    const size_t dst_frame = (size_t)p.h * m * out_w3;
    const uint32_t stride = gridDim.x * (THREADS / 32);

    // strip id -> (frame, row, strip in the row): divisions once, then advanced incrementally by
    // the (constant) stride of the grid
    struct Pos {
        int f, y, strip;
    };
    auto locate = [&](uint32_t st) {
        Pos q;
        const uint32_t rowid = fd_div(p.dspr, st);
        q.strip = (int)(st - rowid * strips_per_row);
        q.f = (int)fd_div(p.dh, rowid);
        q.y = (int)(rowid - (uint32_t)q.f * p.h);
        return q;
    };
    const Pos dpos = locate(stride);        // stride = (df frames, dy rows, dstrip strips)
    auto advance = [&](Pos &q) {
        q.strip += dpos.strip;
        const int c = q.strip >= strips_per_row ? 1 : 0;
        q.strip -= c * strips_per_row;
        q.y += dpos.y + c;
        const int c2 = q.y >= p.h ? 1 : 0;
        q.y -= c2 * p.h;
        q.f += dpos.f + c2;
    };
    // the three source bytes of this lane's pixel of a strip, as raw loads: nothing is done
    // with them until the next iteration, so the loads stay in flight behind this strip's work
    auto fetch = [&](uint32_t st, const Pos &q, unsigned &b0, unsigned &b1, unsigned &b2) {
        b0 = b1 = b2 = 0u;
        if (st >= total) return;
        const int x = q.strip * 32 + lane;
        if (x >= p.w) return;
        const int sy = p.ytab ? __ldg(p.ytab + q.y) : q.y;
        const int sx = p.xtab ? __ldg(p.xtab + x) : x;
        const uint8_t *src = p.src + (size_t)q.f * src_frame + ((size_t)sy * p.src_w + sx) * 3;
        b0 = __ldg(src);
        b1 = __ldg(src + 1);
        b2 = __ldg(src + 2);
    };
    // x4 up-scale, full strips: the 4 x 384 output bytes of a strip are 96 sixteen-byte pieces,
    // three per lane (piece j = lane + 32 i: output row j / 24, piece j % 24 of the row image)
    size_t x4_off[3];
    int x4_col[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = lane + 32 * i;
        x4_col[i] = j % 24;
        x4_off[i] = (size_t)(j / 24) * out_w3 + (size_t)(j % 24) * 16;
    }
    uint32_t st = blockIdx.x * (THREADS / 32) + wib;
    Pos cur = locate(st < total ? st : 0u), nxt = cur;
    unsigned n0, n1, n2;
    fetch(st, cur, n0, n1, n2);
    for (; st < total; st += stride) {
        int r = (int)n0, g = (int)n1, b = (int)n2;
        cur = nxt;
        advance(nxt);
        fetch(st + stride, nxt, n0, n1, n2);
        const int f = cur.f, y = cur.y, strip = cur.strip;
        const int x = strip * 32 + lane;
        const int nvalid = min(32, p.w - strip * 32);
        if (x < p.w) {
            if (p.has_lut) {
                r = s_lut[r];
                g = s_lut[g];
                b = s_lut[b];
            }
            const float thr = threshold_at<KIND>(p, s_mat, mat_in_smem, x, y);
            int idx;
            if (use_table) {
                bool slow;
                idx = (int)v4_pick<KIND, true>(ctx, (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16), thr, slow);
                if (slow) idx = pick_int<KIND>(P, s_coef, K, r, g, b, thr);
            } else {
                idx = p.integral ? pick_int<KIND>(P, s_coef, K, r, g, b, thr)
                                 : pick_f64<KIND>(P, K, r, g, b, thr);
            }
            const unsigned c = (unsigned)s_orgb[4 * idx] | ((unsigned)s_orgb[4 * idx + 1] << 8) |
                               ((unsigned)s_orgb[4 * idx + 2] << 16);
            if (m == 4) {   // 12 bytes = three aligned words (lane * 12 bytes into the row image)
                unsigned *dw = reinterpret_cast<unsigned *>(rowimg) + lane * 3;
                dw[0] = c | (c << 24);
                dw[1] = (c >> 8) | (c << 16);
                dw[2] = (c >> 16) | (c << 8);
            } else {
                uint8_t *d = rowimg + lane * m * 3;
                for (int k = 0; k < m; ++k) {
                    d[3 * k] = (uint8_t)c;
                    d[3 * k + 1] = (uint8_t)(c >> 8);
                    d[3 * k + 2] = (uint8_t)(c >> 16);
                }
            }
            if (p.dst_idx) p.dst_idx[((size_t)f * p.h + y) * p.w + x] = (uint8_t)idx;
        }
        __syncwarp();
        if (!p.dst) continue;                 // index-plane-only output (warp-uniform)
        const int nbytes = nvalid * m * 3;
        const int n16 = nbytes >> 4;          // <= 6 * GEOM2_MAX_M = 48 sixteen-byte pieces
        uint8_t *drow = p.dst + (size_t)f * dst_frame + (size_t)y * m * out_w3 + (size_t)strip * 96 * m;
        const uint4 *img4 = reinterpret_cast<const uint4 *>(rowimg);
        if (m == 4 && nvalid == 32) {         // warp-uniform
#pragma unroll
            for (int i = 0; i < 3; ++i) __stcs(reinterpret_cast<uint4 *>(drow + x4_off[i]), img4[x4_col[i]]);
            __syncwarp();
            continue;
        }
        const uint4 v0 = lane < n16 ? img4[lane] : make_uint4(0, 0, 0, 0);
        const uint4 v1 = lane + 32 < n16 ? img4[lane + 32] : make_uint4(0, 0, 0, 0);
        for (int rr = 0; rr < m; ++rr) {
            uint4 *o = reinterpret_cast<uint4 *>(drow + (size_t)rr * out_w3);
            if (lane < n16) __stcs(o + lane, v0);
            if (lane + 32 < n16) __stcs(o + lane + 32, v1);
            if ((n16 << 4) != nbytes)     // ragged last strip of a row: the few bytes left over
                for (int j = (n16 << 4) + lane; j < nbytes; j += 32) (drow + (size_t)rr * out_w3)[j] = rowimg[j];
        }
        __syncwarp();
    }
}


template <int KIND>
int launch_kind(const ThreshParams &p, bool geom, cudaStream_t st)
{
    int sms = dp_num_sms();
    size_t mat_bytes = (KIND == DP_THRESH_MATRIX && p.mh * p.mw <= 1024) ? (size_t)p.mh * p.mw * 4 : 0;
    if (!geom && p.fast == 4) {
        const bool pow2 = (p.wm & (p.wm - 1)) == 0;
        const bool wide = p.wide != 0;
        const int V4_THREADS = wide ? v4_threads<KIND, true>() : v4_threads<KIND, false>();
        const int V4_WARPS = V4_THREADS / 32;
        size_t smem = V4_TABLE_BYTES + v4_ent_bytes(wide) + v4_orgb_bytes(wide) + 512 + (size_t)V4_WARPS * 3072 +
                      (KIND == DP_THRESH_NONE ? 0 : (size_t)p.sub_bytes) +
                      (KIND == DP_THRESH_MATRIX ? (size_t)p.mh * p.wm * 4 : 0);
        ThreshParams q = p;
        // wide thresholds: deferred fixes when the per-warp lists fit beside everything else
        const size_t defer_bytes = (size_t)V4_WARPS * V4_DEFER_CAP * 4;
        // (plain quantisation flags nearest-row ties only: deferring them costs more than it saves,
        // K=256 0.407 -> 0.421 ms;
        // the narrow format flags too few pixels for it to pay: PICO-8 0.367 -> 0.383 ms; only a
        // tie-heavy palette like the C64's gains, 0.476 -> 0.455 ms: p.defer says which;
        // DP_THRESH_DEFER_ALL / DP_THRESH_NO_DEFER for tools/ and tests)
        q.defer = ((p.defer || getenv("DP_THRESH_DEFER_ALL")) && KIND != DP_THRESH_NONE &&
                   smem + defer_bytes <= 227 * 1024 && !getenv("DP_THRESH_NO_DEFER")) ? 1 : 0;
        if (q.defer) smem += defer_bytes;
        void (*kern)(ThreshParams);
        if constexpr (KIND != DP_THRESH_NONE) {
            if (q.defer)
                kern = wide ? (pow2 ? k_thresh_v4<KIND, true, true, true> : k_thresh_v4<KIND, false, true, true>)
                            : (pow2 ? k_thresh_v4<KIND, true, false, true> : k_thresh_v4<KIND, false, false, true>);
            else
                kern = wide ? (pow2 ? k_thresh_v4<KIND, true, true, false> : k_thresh_v4<KIND, false, true, false>)
                            : (pow2 ? k_thresh_v4<KIND, true, false, false> : k_thresh_v4<KIND, false, false, false>);
        } else {
            kern = wide ? (pow2 ? k_thresh_v4<KIND, true, true, false> : k_thresh_v4<KIND, false, true, false>)
                        : (pow2 ? k_thresh_v4<KIND, true, false, false> : k_thresh_v4<KIND, false, false, false>);
        }
        DP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long ntiles = ((long long)p.frames * p.npix + 511) >> 9;
        const long long want = (ntiles + V4_WARPS - 1) / V4_WARPS;   // frames * npix < 2^31 here
        const int grid = (int)(want < sms ? want : sms);
        kern<<<grid, V4_THREADS, smem, st>>>(q);
    } else if (!geom && p.fast) {
        const bool tsm = p.fast == 1;
        const bool pow2 = ((p.mw & (p.mw - 1)) == 0) && ((p.mh & (p.mh - 1)) == 0);
        const size_t mat_sm = (KIND == DP_THRESH_MATRIX && p.mh * p.mw <= 4096) ? (size_t)p.mh * p.mw * 4 : 0;
        size_t smem = (THREADS / 32) * WTILE_BYTES + DP_MAX_COLORS * 4 + (size_t)p.K * 8 +
                      (tsm ? (size_t)p.thr_cells * 8 : 0) + mat_sm + 256;
        void (*kern)(ThreshParams) =
            tsm ? (pow2 ? k_thresh_fast<KIND, true, true> : k_thresh_fast<KIND, true, false>)
                : (pow2 ? k_thresh_fast<KIND, false, true> : k_thresh_fast<KIND, false, false>);
        DP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        ThreshParams q = p;
        q.tiles_per_frame = (p.npix + WTILE_PX - 1) / WTILE_PX;
        long long tt = (long long)q.tiles_per_frame * p.frames;
        DP_REQUIRE(tt < (1ll << 31), "too many tiles in one call");
        q.total_tiles = (int)tt;
        long long want = (tt + (THREADS / 32) - 1) / (THREADS / 32);
        long long cap = (long long)sms * per_sm;
        int grid = (int)(want < cap ? want : cap);
        kern<<<grid, THREADS, smem, st>>>(q);
    } else if (!geom) {
        size_t smem = 2 * TILE_BUF + TILE_PX + 256 + 1024 + (size_t)p.K * 16 + mat_bytes;
        DP_CUDA(cudaFuncSetAttribute(k_thresh_tile<KIND>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_thresh_tile<KIND>,
                                                              THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        int grid = sms * per_sm;
        if (grid > p.total_tiles) grid = p.total_tiles;
        k_thresh_tile<KIND><<<grid, THREADS, smem, st>>>(p);
    } else if (p.upscale <= GEOM2_MAX_M && ((size_t)p.w * p.upscale * 3) % 16 == 0 &&
               (reinterpret_cast<uintptr_t>(p.dst) & 15) == 0 &&
               ((size_t)p.h * p.upscale * p.w * p.upscale * 3) % 16 == 0 &&
               (long long)p.frames * p.h * ((p.w + 31) / 32) < (1ll << 31)) {
        // output rows are 16-byte aligned: strip kernel with 128-bit stores
        ThreshParams q = p;
        q.dspr = make_fastdiv((uint32_t)((p.w + 31) / 32));
        q.dh = make_fastdiv((uint32_t)p.h);
        q.geom_table = p.geom_table;
        const size_t mat2 = (KIND == DP_THRESH_MATRIX && p.mh * p.mw <= GEOM2_MAT_SMEM) ? (size_t)p.mh * p.mw * 4 : 0;
        size_t smem = 256 + 1024 + (size_t)p.K * 16 + mat2 + 16 +
                      (size_t)(THREADS / 32) * 96 * GEOM2_MAX_M;
        long long strips = (long long)p.frames * p.h * ((p.w + 31) / 32);
        long long want = (strips + THREADS / 32 - 1) / (THREADS / 32);
        int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        k_thresh_geom2<KIND><<<grid, THREADS, smem, st>>>(q);
    } else {
        size_t smem = 256 + 1024 + (size_t)p.K * 16 + mat_bytes;
        long long total = (long long)p.frames * p.npix;
        long long want = (total + THREADS - 1) / THREADS;
        int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
        k_thresh_geom<KIND><<<grid, THREADS, smem, st>>>(p);
    }
    DP_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" int dp_threshold_dither(const dp_palette *pal, const uint8_t *src_rgb, int frames,
                                   const dp_geometry *geo, int kind, const float *matrix,
                                   int mat_h, int mat_w, float ign_xoff, float ign_yoff,
                                   float ign_scale, uint8_t *dst_rgb, uint8_t *dst_idx,
                                   void *stream)
{
    DP_RANGE("dp_threshold_dither");
    DP_REQUIRE(pal && src_rgb && geo, "null argument");
    DP_REQUIRE(dst_rgb || dst_idx, "no output: dst_rgb and dst_idx are both null");
    DP_REQUIRE(frames >= 0 && geo->h >= 0 && geo->w >= 0, "negative size");
    DP_REQUIRE(kind >= DP_THRESH_NONE && kind <= DP_THRESH_IGN, "unknown threshold kind");
    if (frames == 0 || geo->h == 0 || geo->w == 0) return 0;
    DP_REQUIRE((long long)geo->h * geo->w < (1ll << 31), "frame too large");
    if (kind == DP_THRESH_MATRIX)
        DP_REQUIRE(matrix && mat_h > 0 && mat_w > 0, "threshold matrix missing");
    const int m = geo->upscale < 1 ? 1 : geo->upscale;
    const bool geom = geo->ytab || geo->xtab || m != 1;
    if (!geom)
        DP_REQUIRE(geo->src_h == geo->h && geo->src_w == geo->w, "identity geometry size mismatch");

    ThreshParams p;
    memset(&p, 0, sizeof(p));
    p.P = reinterpret_cast<const PalDev *>(pal->blob);
    p.src = src_rgb;
    p.dst = dst_rgb;
    p.dst_idx = dst_idx;
    p.frames = frames;
    p.h = geo->h;
    p.w = geo->w;
    p.npix = geo->h * geo->w;
    p.matrix = matrix;
    p.mh = mat_h > 0 ? mat_h : 1;
    p.mw = mat_w > 0 ? mat_w : 1;
    p.dw = make_fastdiv((uint32_t)geo->w);
    p.dmh = make_fastdiv((uint32_t)p.mh);
    p.dmw = make_fastdiv((uint32_t)p.mw);
    p.ign_xoff = ign_xoff;
    p.ign_yoff = ign_yoff;
    p.ign_scale = ign_scale;
    p.tiles_per_frame = (p.npix + TILE_PX - 1) / TILE_PX;
    long long tt = (long long)p.tiles_per_frame * frames;
    DP_REQUIRE(tt < (1ll << 31), "too many tiles in one call");
    p.total_tiles = (int)tt;
    p.K = pal->dev.K;
    p.integral = pal->dev.integral;
    p.has_lut = pal->has_lut;
    p.src_h = geo->src_h;
    p.src_w = geo->src_w;
    p.upscale = m;
    p.ytab = geo->ytab;
    p.xtab = geo->xtab;
    p.thr_cells = pal->dev.thr_cells;
    p.fast = 0;
    if (!geom && pal->dev.integral && pal->dev.K >= 2 && pal->dev.thr_table &&
        ((reinterpret_cast<uintptr_t>(src_rgb) | reinterpret_cast<uintptr_t>(dst_rgb)) & 15) == 0 &&
        p.npix % 16 == 0 && (!dst_idx || (reinterpret_cast<uintptr_t>(dst_idx) & 3) == 0))
        p.fast = (pal->dev.thr_cells <= 4096) ? 1 : 2;
    const bool wide = pal->dev.thr4_wide != 0;
    if (p.fast && (kind == DP_THRESH_NONE ? pal->dev.near3_table : pal->dev.thr4_table) && (pal->dev.K <= 30 || wide) && !pal->has_lut && geo->w % 16 == 0 &&
        (!dst_idx || (reinterpret_cast<uintptr_t>(dst_idx) & 15) == 0) && !getenv("DP_THRESH_NO_V4")) {
        // widened matrix width: the smallest common multiple of mat_w and 16
        int wm = 16;
        if (kind == DP_THRESH_MATRIX) {
            int a = p.mw, b = 16;
            while (b) { int t = a % b; a = b; b = t; }
            wm = p.mw / a * 16;
        }
        const int sub_bytes = wide ? 0 : ((8 * pal->dev.thr4_nsub + 3) & ~3) * 4;
        const long long need = V4_TABLE_BYTES + v4_ent_bytes(wide) + v4_orgb_bytes(wide) + 512 +
                               (kind == DP_THRESH_NONE ? (wide ? 30ll : 31ll) * 3072 : 24ll * 3072 + sub_bytes) +
                               (kind == DP_THRESH_MATRIX ? (long long)p.mh * wm * 4 : 0);
        if (need <= 227 * 1024 && (long long)frames * p.npix < (1ll << 31) &&
            (kind != DP_THRESH_MATRIX || (long long)p.mh * wm <= 4096)) {
            p.fast = 4;
            p.wide = wide ? 1 : 0;
            // deferred fixes: always for the wide format; for the narrow one only when the palette
            // is tie-heavy (>= 0.2 % of the byte colours in its exception table, e.g. the C64's)
            p.defer = (wide || pal->dev.tie_n >= 32768) ? 1 : 0;
            p.sub_bytes = sub_bytes;
            p.wm = wm;
            p.dwm = make_fastdiv((uint32_t)wm);
            p.dnpix = make_fastdiv((uint32_t)p.npix);
        }
    }
    p.geom_table = (pal->dev.integral && pal->dev.K >= 2 && pal->dev.K <= 30 && !wide && !pal->has_lut &&
                    pal->dev.thr4_table && pal->dev.near3_table) ? 1 : 0;
    cudaStream_t st = dp_stream(stream);
    switch (kind) {
        case DP_THRESH_NONE: return launch_kind<DP_THRESH_NONE>(p, geom, st);
        case DP_THRESH_MATRIX: return launch_kind<DP_THRESH_MATRIX>(p, geom, st);
        default: return launch_kind<DP_THRESH_IGN>(p, geom, st);
    }
}

// Host-buffer convenience: H2D, kernel, D2H and a stream synchronise inside the call.  The call
// runs on a private non-blocking stream (never the legacy default stream, which would serialise
// with every other stream of the process) and takes its device buffers from the retained
// stream-ordered pool (no cudaMalloc / cudaFree per call).
extern "C" int dp_threshold_dither_host(const dp_palette *pal, const uint8_t *src_rgb_host,
                                        int frames, int h, int w, int kind,
                                        const float *matrix_host, int mat_h, int mat_w,
                                        float ign_xoff, float ign_yoff, float ign_scale,
                                        uint8_t *dst_rgb_host)
{
    DP_RANGE("dp_threshold_dither_host");
    DP_REQUIRE(pal && src_rgb_host && dst_rgb_host, "null argument");
    DP_REQUIRE(frames >= 0 && h >= 0 && w >= 0, "negative size");
    if (kind == DP_THRESH_MATRIX) DP_REQUIRE(matrix_host && mat_h > 0 && mat_w > 0, "threshold matrix missing");
    const size_t bytes = (size_t)frames * h * w * 3;
    if (bytes == 0) return 0;
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    if (dp_retain_pool(dev)) return 1;
    cudaStream_t st = nullptr;
    DP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint8_t *dsrc = nullptr, *ddst = nullptr;
    float *dmat = nullptr;
    auto fail = [&](const char *what) {
        dp_set_error("dp_threshold_dither_host: %s: %s", what, cudaGetErrorString(cudaGetLastError()));
        if (dsrc) cudaFreeAsync(dsrc, st);
        if (ddst) cudaFreeAsync(ddst, st);
        if (dmat) cudaFreeAsync(dmat, st);
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        return 1;
    };
    if (cudaMallocAsync(reinterpret_cast<void **>(&dsrc), bytes, st) != cudaSuccess ||
        cudaMallocAsync(reinterpret_cast<void **>(&ddst), bytes, st) != cudaSuccess)
        return fail("device allocation");
    if (kind == DP_THRESH_MATRIX &&
        (cudaMallocAsync(reinterpret_cast<void **>(&dmat), (size_t)mat_h * mat_w * 4, st) != cudaSuccess ||
         cudaMemcpyAsync(dmat, matrix_host, (size_t)mat_h * mat_w * 4, cudaMemcpyHostToDevice, st) != cudaSuccess))
        return fail("threshold matrix upload");
    if (cudaMemcpyAsync(dsrc, src_rgb_host, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail("H2D copy");
    dp_geometry geo;
    memset(&geo, 0, sizeof(geo));
    geo.src_h = geo.h = h;
    geo.src_w = geo.w = w;
    geo.upscale = 1;
    int rc = dp_threshold_dither(pal, dsrc, frames, &geo, kind, dmat, mat_h, mat_w, ign_xoff, ign_yoff,
                                 ign_scale, ddst, nullptr, st);
    if (rc == 0 && (cudaMemcpyAsync(dst_rgb_host, ddst, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess))
        return fail("D2H copy");
    cudaFreeAsync(dsrc, st);
    cudaFreeAsync(ddst, st);
    if (dmat) cudaFreeAsync(dmat, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}
