/*
 * dp_oracle.c -- CPU restatement (TEST INFRASTRUCTURE, NOT PRODUCT CODE) of the serial
 * parts of dither_pie's per-pixel hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (dither_pie_b200/) never links or calls it.
 *
 * What is restated here, and the reference lines each function follows:
 *   orc_kdtree_query      scipy.spatial.cKDTree.query (scipy 1.18.1, third-party, not under
 *                         /root/reference; algorithm restated from its published source, see
 *                         SURVEY.md 5.8).  Call sites: dithering_lib.py:339-340, 358-360,
 *                         554-556, 748-749, 1243, 1633.
 *   orc_error_diffusion   dithering_lib.py:212-308 (_error_diffusion_numba): f32 state, f64
 *                         arithmetic, one f32 rounding per accumulation, strict '<' argmin.
 *   orc_hybrid            dithering_lib.py:1396-1494 (_hybrid_numba): the Floyd-Steinberg loop of
 *                         orc_error_diffusion with the error split into luminance and colour
 *                         parts (f64, one rounding per operation).
 *   orc_perceptual        dithering_lib.py:1030-1066 (PerceptualDitherStrategy, pure Python):
 *                         all-f32 Floyd-Steinberg whose taps are scaled by a luminance factor of
 *                         the ORIGINAL pixel; KD-tree nearest of the UNCLAMPED work value.
 *   orc_adaptive          dithering_lib.py:989-1015 (AdaptiveVarianceDitherStrategy, pure Python):
 *                         all-f32 Floyd-Steinberg, unclamped KD-tree lookups, the error of a
 *                         pixel distributed only where the caller's gate plane is set.
 *   orc_ostromoukhov      dithering_lib.py:1225-1269 (the live pure-Python path): all-f32
 *                         arithmetic, f32-rounded weights, KD-tree nearest (tie rules of scipy).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build.py).  -ffp-contract=off is
 * required: the reference performs every multiply and add as a separately rounded operation.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * KD-tree (flattened, pre-order node arrays exported from scipy at palette set-up time).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_nodes;
    int32_t n_points;
    const int32_t *split_dim; /* -1 => leaf */
    const double *split;
    const int32_t *start_idx;
    const int32_t *end_idx;
    const int32_t *lesser;  /* node id */
    const int32_t *greater; /* node id */
    const int32_t *indices; /* permutation of palette rows */
    const double *data;     /* n_points x 3, f64 */
    const double *mins;     /* 3 */
    const double *maxes;    /* 3 */
} orc_kdtree;

typedef struct {
    double priority;
    int32_t payload;
} hitem;

/* scipy's array binary min-heap (strict comparisons both ways). */
typedef struct {
    hitem *a;
    int n;
} heap_t;

static void heap_push(heap_t *h, hitem it)
{
    int i = h->n++;
    h->a[i] = it;
    while (i > 0 && h->a[i].priority < h->a[(i - 1) / 2].priority) {
        hitem t = h->a[(i - 1) / 2];
        h->a[(i - 1) / 2] = h->a[i];
        h->a[i] = t;
        i = (i - 1) / 2;
    }
}

static void heap_remove(heap_t *h)
{
    h->a[0] = h->a[h->n - 1];
    h->n--;
    int nn = h->n, i = 0, j = 1, k = 2;
    while ((j < nn && h->a[i].priority > h->a[j].priority) ||
           (k < nn && h->a[i].priority > h->a[k].priority)) {
        int l = (k < nn && h->a[j].priority > h->a[k].priority) ? k : j;
        hitem t = h->a[l];
        h->a[l] = h->a[i];
        h->a[i] = t;
        i = l;
        j = 2 * i + 1;
        k = 2 * i + 2;
    }
}

typedef struct {
    int32_t node;
    double min_distance;
    double side[3];
} ninfo;

#define ORC_MAX_NODES 1024

/* One query point; k in {1,2}.  out_idx[j] = n_points and out_d2[j] = +inf for missing
 * neighbours (scipy returns index n and distance inf).  out_d2 holds SQUARED distances as the
 * traversal keeps them; the caller takes sqrt where the reference does. */
static void kd_query_one(const orc_kdtree *t, const double x[3], int k, int32_t *out_idx,
                         double *out_d2)
{
    ninfo pool[ORC_MAX_NODES];
    int npool = 0;
    hitem qbuf[ORC_MAX_NODES];
    hitem nbuf[4];
    heap_t q = {qbuf, 0};
    heap_t nb = {nbuf, 0};
    double ub = INFINITY;

    ninfo *cur = &pool[npool++];
    cur->node = 0;
    cur->min_distance = 0.0;
    for (int i = 0; i < 3; ++i) {
        double s = t->mins[i] - x[i];
        double s2 = x[i] - t->maxes[i];
        if (s2 > s) s = s2;
        if (s < 0.0) s = 0.0;
        cur->side[i] = s * s;
        cur->min_distance += cur->side[i];
    }

    for (;;) {
        int32_t nd = cur->node;
        int sd = t->split_dim[nd];
        if (sd == -1) {
            for (int32_t i = t->start_idx[nd]; i < t->end_idx[nd]; ++i) {
                int32_t pi = t->indices[i];
                const double *p = t->data + 3 * (size_t)pi;
                double s = 0.0;
                for (int c = 0; c < 3; ++c) {
                    double d = p[c] - x[c];
                    s += d * d;
                }
                if (s < ub) {
                    if (nb.n == k) heap_remove(&nb);
                    hitem it = {-s, pi};
                    heap_push(&nb, it);
                    if (nb.n == k) ub = -nb.a[0].priority;
                }
            }
            if (q.n == 0) break;
            hitem top = q.a[0];
            heap_remove(&q);
            cur = &pool[top.payload];
        } else {
            if (cur->min_distance > ub) break;
            double sp = t->split[nd];
            ninfo *far = &pool[npool++];
            *far = *cur;
            if (x[sd] < sp) {
                cur->node = t->lesser[nd];
                far->node = t->greater[nd];
            } else {
                cur->node = t->greater[nd];
                far->node = t->lesser[nd];
            }
            double diff = fabs(sp - x[sd]);
            double nsd = diff * diff;
            far->min_distance += nsd - far->side[sd];
            far->side[sd] = nsd;
            ninfo *near = cur;
            if (near->min_distance > far->min_distance) {
                ninfo *tmp = near;
                near = far;
                far = tmp;
            }
            cur = near;
            if (far->min_distance <= ub) {
                hitem it = {far->min_distance, (int32_t)(far - pool)};
                heap_push(&q, it);
            }
        }
    }

    int nnb = nb.n;
    for (int j = 0; j < k; ++j) {
        out_idx[j] = t->n_points;
        out_d2[j] = INFINITY;
    }
    for (int i = nnb - 1; i >= 0; --i) {
        out_idx[i] = nb.a[0].payload;
        out_d2[i] = -nb.a[0].priority;
        heap_remove(&nb);
    }
}

/* Batch query: points f64 [n,3]; out_idx int32 [n,k]; out_d2 f64 [n,k] (squared). */
int orc_kdtree_query(const orc_kdtree *t, const double *points, int64_t n, int k,
                     int32_t *out_idx, double *out_d2)
{
    if (k < 1 || k > 2 || t->n_nodes > ORC_MAX_NODES / 2) return -1;
    for (int64_t i = 0; i < n; ++i)
        kd_query_one(t, points + 3 * i, k, out_idx + (size_t)k * i, out_d2 + (size_t)k * i);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Error diffusion, numba semantics (dithering_lib.py:212-308).
 *   work     f32 [h,w,3], in/out (state is f32)
 *   palette  f32 [K,3]
 *   offsets  int32 [ntaps,2] (dx,dy); weights f32 [ntaps]; divisor f64
 * Every per-pixel quantity is f64; each '+=' is  f32( f64(work) + err*wgt ).
 * ---------------------------------------------------------------------------------------- */
int orc_error_diffusion(float *work, int h, int w, const float *palette, int K,
                        const int32_t *offsets, const float *weights, int ntaps, double divisor,
                        int serpentine, uint8_t *out_idx /* may be NULL, [h,w] */)
{
    for (int y = 0; y < h; ++y) {
        int dir = (serpentine && (y & 1)) ? -1 : 1;
        int x = dir > 0 ? 0 : w - 1;
        for (int n = 0; n < w; ++n, x += dir) {
            float *px = work + 3 * ((size_t)y * w + x);
            double v[3];
            for (int c = 0; c < 3; ++c) {
                double t = (double)px[c];
                if (t < 0.0) t = 0.0;
                else if (t > 255.0) t = 255.0;
                v[c] = t;
            }
            int best = 0;
            double bestd = 1e20;
            for (int i = 0; i < K; ++i) {
                double dr = v[0] - (double)palette[3 * i + 0];
                double dg = v[1] - (double)palette[3 * i + 1];
                double db = v[2] - (double)palette[3 * i + 2];
                double d = dr * dr + dg * dg + db * db;
                if (d < bestd) {
                    bestd = d;
                    best = i;
                }
            }
            double e[3];
            for (int c = 0; c < 3; ++c) {
                float ch = palette[3 * best + c];
                px[c] = ch;
                e[c] = v[c] - (double)ch;
            }
            if (out_idx) out_idx[(size_t)y * w + x] = (uint8_t)best;
            for (int k = 0; k < ntaps; ++k) {
                int nx = x + offsets[2 * k] * dir;
                int ny = y + offsets[2 * k + 1];
                if (nx >= 0 && nx < w && ny >= 0 && ny < h) {
                    double wgt = (double)weights[k] / divisor;
                    float *q = work + 3 * ((size_t)ny * w + nx);
                    for (int c = 0; c < 3; ++c) q[c] = (float)((double)q[c] + e[c] * wgt);
                }
            }
        }
    }
    /* final clamp pass (:285-306); a no-op for visited pixels, kept for fidelity */
    for (size_t i = 0; i < (size_t)h * w * 3; ++i) {
        float t = work[i];
        if (t < 0.0f) t = 0.0f;
        else if (t > 255.0f) t = 255.0f;
        work[i] = t;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Hybrid (dithering_lib.py:1396-1494, _hybrid_numba -- the path HybridDitherStrategy.dither
 * takes whenever numba imports, :1116-1127).  Floyd-Steinberg footprint; the error of a pixel is
 * split into a luminance part and a colour part which are scaled separately (:1447-1455).
 * numba types: the clamp assigns 0.0/255.0 (f64) to r,g,b so they are f64 holding f32 values;
 * every product and sum below is f64 and separately rounded (no fast-math, no contraction).
 * ---------------------------------------------------------------------------------------- */
int orc_hybrid(float *work, int h, int w, const float *palette, int K, double lum_factor,
               double col_factor, uint8_t *out_idx /* may be NULL, [h,w] */)
{
    static const int dxs[4] = {1, -1, 0, 1}, dys[4] = {0, 1, 1, 1};
    static const double wts[4] = {7.0 / 16.0, 3.0 / 16.0, 5.0 / 16.0, 1.0 / 16.0};
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float *px = work + 3 * ((size_t)y * w + x);
            double v[3];
            for (int c = 0; c < 3; ++c) {
                double t = (double)px[c];
                if (t < 0.0) t = 0.0;
                else if (t > 255.0) t = 255.0;
                v[c] = t;
            }
            int best = 0;
            double bestd = 1e20;
            for (int i = 0; i < K; ++i) {
                double dr = v[0] - (double)palette[3 * i + 0];
                double dg = v[1] - (double)palette[3 * i + 1];
                double db = v[2] - (double)palette[3 * i + 2];
                double d = dr * dr + dg * dg + db * db;
                if (d < bestd) {
                    bestd = d;
                    best = i;
                }
            }
            double e[3], fe[3];
            for (int c = 0; c < 3; ++c) {
                float ch = palette[3 * best + c];
                px[c] = ch;
                e[c] = v[c] - (double)ch;
            }
            if (out_idx) out_idx[(size_t)y * w + x] = (uint8_t)best;
            {
                static const double cf[3] = {0.299, 0.587, 0.114};
                double lum = 0.299 * e[0] + 0.587 * e[1];
                lum = lum + 0.114 * e[2];
                for (int c = 0; c < 3; ++c) {
                    double l = cf[c] * lum;
                    fe[c] = lum_factor * l + col_factor * (e[c] - l);
                }
            }
            for (int k = 0; k < 4; ++k) {
                int nx = x + dxs[k], ny = y + dys[k];
                if (nx >= 0 && nx < w && ny < h) {
                    float *q = work + 3 * ((size_t)ny * w + nx);
                    for (int c = 0; c < 3; ++c) q[c] = (float)((double)q[c] + fe[c] * wts[k]);
                }
            }
        }
    }
    for (size_t i = 0; i < (size_t)h * w * 3; ++i) {   /* final clamp pass (:1474-1492) */
        float t = work[i];
        if (t < 0.0f) t = 0.0f;
        else if (t > 255.0f) t = 255.0f;
        work[i] = t;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Ostromoukhov, live pure-Python path (dithering_lib.py:1225-1269): f32 throughout,
 * nearest colour by the KD-tree (f64 distances, scipy tie rules).
 *   coeffs int32 [256,3]
 * ---------------------------------------------------------------------------------------- */
int orc_ostromoukhov(float *work, int h, int w, const float *palette, int K,
                     const orc_kdtree *tree, const int32_t *coeffs, int serpentine,
                     uint8_t *out_idx)
{
    (void)K;
    const float c299 = (float)0.299, c587 = (float)0.587, c114 = (float)0.114;
    for (int y = 0; y < h; ++y) {
        int dir = (serpentine && (y & 1)) ? -1 : 1;
        int x = dir > 0 ? 0 : w - 1;
        for (int n = 0; n < w; ++n, x += dir) {
            float *px = work + 3 * ((size_t)y * w + x);
            float old[3];
            double xq[3];
            for (int c = 0; c < 3; ++c) {
                float t = px[c];
                if (t < 0.0f) t = 0.0f;
                else if (t > 255.0f) t = 255.0f;
                old[c] = t;
                xq[c] = (double)t;
            }
            int32_t idx;
            double d2;
            kd_query_one(tree, xq, 1, &idx, &d2);
            float err[3];
            for (int c = 0; c < 3; ++c) {
                float ch = palette[3 * idx + c];
                px[c] = ch;
                err[c] = old[c] - ch;
            }
            if (out_idx) out_idx[(size_t)y * w + x] = (uint8_t)idx;
            float lum = c299 * old[0];
            lum = lum + c587 * old[1];
            lum = lum + c114 * old[2];
            if (lum < 0.0f) lum = 0.0f;
            if (lum > 255.0f) lum = 255.0f;
            int li = (int)lum;
            int c0 = coeffs[3 * li], c1 = coeffs[3 * li + 1], c2 = coeffs[3 * li + 2];
            int div = c0 + c1 + c2;
            if (div == 0) continue;
            float w0 = (float)((double)c0 / (double)div);
            float w1 = (float)((double)c1 / (double)div);
            float w2 = (float)((double)c2 / (double)div);
            int nx = x + dir;
            if (nx >= 0 && nx < w) {
                float *q = work + 3 * ((size_t)y * w + nx);
                for (int c = 0; c < 3; ++c) q[c] = q[c] + err[c] * w0;
            }
            if (y + 1 < h) {
                nx = x - dir;
                if (nx >= 0 && nx < w) {
                    float *q = work + 3 * ((size_t)(y + 1) * w + nx);
                    for (int c = 0; c < 3; ++c) q[c] = q[c] + err[c] * w1;
                }
                float *q = work + 3 * ((size_t)(y + 1) * w + x);
                for (int c = 0; c < 3; ++c) q[c] = q[c] + err[c] * w2;
            }
        }
    }
    for (size_t i = 0; i < (size_t)h * w * 3; ++i) {
        float t = work[i];
        if (t < 0.0f) t = 0.0f;
        else if (t > 255.0f) t = 255.0f;
        work[i] = t;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Perceptual (dithering_lib.py:1030-1066, pure Python; numpy >= 2 scalar promotion: every
 * quantity is f32, python floats are weak).
 *   gray     = ((0.299f R + 0.587f G) + 0.114f B) of the ORIGINAL image (:1037), f32 ops
 *   per pixel: old = work[y,x] (NOT clamped); idx = KDTree.query(old); err = old - chosen (f32);
 *              sens = 0.5f + 0.5f * (gray / 255.0f); for the FS taps: work += err * f32(wgt * sens)
 * ---------------------------------------------------------------------------------------- */
int orc_perceptual(float *work, int h, int w, const float *palette, int K, const orc_kdtree *tree,
                   uint8_t *out_idx)
{
    (void)K;
    static const int dxs[4] = {1, -1, 0, 1}, dys[4] = {0, 1, 1, 1};
    static const float wts[4] = {0.4375f, 0.1875f, 0.3125f, 0.0625f};
    const float c299 = (float)0.299, c587 = (float)0.587, c114 = (float)0.114;
    float *gray = (float *)malloc(sizeof(float) * (size_t)h * w);
    if (!gray) return 1;
    for (size_t i = 0; i < (size_t)h * w; ++i) {
        float a = c299 * work[3 * i], b = c587 * work[3 * i + 1], c = c114 * work[3 * i + 2];
        float ab = a + b;
        gray[i] = ab + c;
    }
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float *px = work + 3 * ((size_t)y * w + x);
            float old[3], err[3];
            double xq[3];
            for (int c = 0; c < 3; ++c) {
                old[c] = px[c];
                xq[c] = (double)px[c];
            }
            int32_t idx;
            double d2;
            kd_query_one(tree, xq, 1, &idx, &d2);
            for (int c = 0; c < 3; ++c) {
                float ch = palette[3 * idx + c];
                px[c] = ch;
                err[c] = old[c] - ch;
            }
            if (out_idx) out_idx[(size_t)y * w + x] = (uint8_t)idx;
            float q = gray[(size_t)y * w + x] / 255.0f;
            float hq = 0.5f * q;
            float sens = 0.5f + hq;
            for (int k = 0; k < 4; ++k) {
                int nx = x + dxs[k], ny = y + dys[k];
                if (nx >= 0 && nx < w && ny >= 0 && ny < h) {
                    float wk = wts[k] * sens;
                    float *t = work + 3 * ((size_t)ny * w + nx);
                    for (int c = 0; c < 3; ++c) {
                        float pr = err[c] * wk;
                        t[c] = t[c] + pr;
                    }
                }
            }
        }
    }
    free(gray);
    for (size_t i = 0; i < (size_t)h * w * 3; ++i) {   /* np.clip(work, 0, 255) (:1065) */
        float t = work[i];
        if (t < 0.0f) t = 0.0f;
        else if (t > 255.0f) t = 255.0f;
        work[i] = t;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Adaptive variance (dithering_lib.py:989-1015, pure Python): as orc_perceptual without the
 * luminance factor; `gate` (u8 [h,w]) = var_map >= var_threshold, computed by the caller with
 * scipy's uniform_filter exactly as the reference does (:1021-1025).
 * ---------------------------------------------------------------------------------------- */
int orc_adaptive(float *work, int h, int w, const float *palette, int K, const orc_kdtree *tree,
                 const uint8_t *gate, uint8_t *out_idx)
{
    (void)K;
    static const int dxs[4] = {1, -1, 0, 1}, dys[4] = {0, 1, 1, 1};
    static const float wts[4] = {0.4375f, 0.1875f, 0.3125f, 0.0625f};
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float *px = work + 3 * ((size_t)y * w + x);
            float old[3], err[3];
            double xq[3];
            for (int c = 0; c < 3; ++c) {
                old[c] = px[c];
                xq[c] = (double)px[c];
            }
            int32_t idx;
            double d2;
            kd_query_one(tree, xq, 1, &idx, &d2);
            for (int c = 0; c < 3; ++c) {
                float ch = palette[3 * idx + c];
                px[c] = ch;
                err[c] = old[c] - ch;
            }
            if (out_idx) out_idx[(size_t)y * w + x] = (uint8_t)idx;
            if (!gate[(size_t)y * w + x]) continue;
            for (int k = 0; k < 4; ++k) {
                int nx = x + dxs[k], ny = y + dys[k];
                if (nx >= 0 && nx < w && ny < h) {
                    float *t = work + 3 * ((size_t)ny * w + nx);
                    for (int c = 0; c < 3; ++c) {
                        float pr = err[c] * wts[k];
                        t[c] = t[c] + pr;
                    }
                }
            }
        }
    }
    for (size_t i = 0; i < (size_t)h * w * 3; ++i) {
        float t = work[i];
        if (t < 0.0f) t = 0.0f;
        else if (t > 255.0f) t = 255.0f;
        work[i] = t;
    }
    return 0;
}
