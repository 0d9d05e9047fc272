// dp_search.cuh -- nearest / second-nearest palette search, device side.
//
// The reference defines "nearest" and "second nearest" through scipy.spatial.KDTree.query
// (dithering_lib.py:339-340, 358-360, 554-556, 748-749, 1243, 1633).  For almost every pixel
// the answer is the plain top-2 by distance; when distances tie exactly the answer depends on
// scipy's traversal order and heap mechanics.  The kernels therefore
//   1. brute-force exact distances and keep a top-3 ordered by (distance, palette row),
//   2. take the top-2 when d1 < d2 < d3 strictly,
//   3. otherwise replay scipy's traversal on the exported tree (kd_emulate below).
#pragma once

#include "dp_common.cuh"

#define DP_KD_POOL 72   // >= 1 + inner nodes; checked at palette creation
#define DP_INF_F64 __longlong_as_double(0x7ff0000000000000LL)

// ---------------------------------------------------------------------------------------
// Faithful replay of cKDTree.query(x, k=KQ, p=2, eps=0) for one point.
// All distances are squared sums s = ((0 + d0^2) + d1^2) + d2^2 in f64, no fused multiply-add.
// Results: oi[j] (palette row; K if missing), os[j] (squared distance; +inf if missing).
// ---------------------------------------------------------------------------------------
template <int KQ>
__device__ __noinline__ void kd_emulate(const PalDev *__restrict__ P, double x0, double x1,
                                        double x2, int *oi, double *os)
{
    double p_md[DP_KD_POOL];
    double p_sd[DP_KD_POOL][3];
    int p_node[DP_KD_POOL];
    double q_pri[DP_KD_POOL];
    int q_pay[DP_KD_POOL];
    int qn = 0, npool = 0;
    double n_pri[2];
    int n_pay[2];
    int nn = 0;
    double ub = DP_INF_F64;
    double x[3] = {x0, x1, x2};

    int cur = npool++;
    p_node[cur] = 0;
    {
        double md = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double s = __dsub_rn(P->kd_mins[i], x[i]);
            double s2 = __dsub_rn(x[i], P->kd_maxes[i]);
            if (s2 > s) s = s2;
            if (s < 0.0) s = 0.0;
            double sq = __dmul_rn(s, s);
            p_sd[cur][i] = sq;
            md = __dadd_rn(md, sq);
        }
        p_md[cur] = md;
    }

    for (;;) {
        int nd = p_node[cur];
        int sd = P->kd_split_dim[nd];
        if (sd == -1) {
            int e = P->kd_end[nd];
            for (int i = P->kd_start[nd]; i < e; ++i) {
                int pi = P->kd_indices[i];
                const double *pp = P->pal_f64 + 3 * pi;
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double d = __dsub_rn(pp[c], x[c]);
                    s = __dadd_rn(s, __dmul_rn(d, d));
                }
                if (s < ub) {
                    if (nn == KQ) {  // remove the top (largest distance)
                        n_pri[0] = n_pri[nn - 1];
                        n_pay[0] = n_pay[nn - 1];
                        nn--;
                    }
                    int i2 = nn++;
                    n_pri[i2] = -s;
                    n_pay[i2] = pi;
                    if (KQ == 2 && i2 == 1 && n_pri[1] < n_pri[0]) {
                        double tp = n_pri[0];
                        int ti = n_pay[0];
                        n_pri[0] = n_pri[1];
                        n_pay[0] = n_pay[1];
                        n_pri[1] = tp;
                        n_pay[1] = ti;
                    }
                    if (nn == KQ) ub = -n_pri[0];
                }
            }
            if (qn == 0) break;
            cur = q_pay[0];
            // heap remove
            q_pri[0] = q_pri[qn - 1];
            q_pay[0] = q_pay[qn - 1];
            qn--;
            int i = 0, j = 1, k = 2;
            while ((j < qn && q_pri[i] > q_pri[j]) || (k < qn && q_pri[i] > q_pri[k])) {
                int l = (k < qn && q_pri[j] > q_pri[k]) ? k : j;
                double tp = q_pri[l];
                int ti = q_pay[l];
                q_pri[l] = q_pri[i];
                q_pay[l] = q_pay[i];
                q_pri[i] = tp;
                q_pay[i] = ti;
                i = l;
                j = 2 * i + 1;
                k = 2 * i + 2;
            }
        } else {
            if (p_md[cur] > ub) break;
            double sp = P->kd_split[nd];
            int far = npool++;
            p_md[far] = p_md[cur];
            p_sd[far][0] = p_sd[cur][0];
            p_sd[far][1] = p_sd[cur][1];
            p_sd[far][2] = p_sd[cur][2];
            if (x[sd] < sp) {
                p_node[cur] = P->kd_lesser[nd];
                p_node[far] = P->kd_greater[nd];
            } else {
                p_node[cur] = P->kd_greater[nd];
                p_node[far] = P->kd_lesser[nd];
            }
            double diff = fabs(__dsub_rn(sp, x[sd]));
            double nsd = __dmul_rn(diff, diff);
            p_md[far] = __dadd_rn(p_md[far], __dsub_rn(nsd, p_sd[far][sd]));
            p_sd[far][sd] = nsd;
            int near = cur;
            if (p_md[near] > p_md[far]) {
                int t = near;
                near = far;
                far = t;
            }
            cur = near;
            if (p_md[far] <= ub) {
                int i = qn++;
                q_pri[i] = p_md[far];
                q_pay[i] = far;
                while (i > 0 && q_pri[i] < q_pri[(i - 1) / 2]) {
                    int pa = (i - 1) / 2;
                    double tp = q_pri[pa];
                    int ti = q_pay[pa];
                    q_pri[pa] = q_pri[i];
                    q_pay[pa] = q_pay[i];
                    q_pri[i] = tp;
                    q_pay[i] = ti;
                    i = pa;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        oi[j] = P->K;
        os[j] = DP_INF_F64;
    }
    // pop: fills positions nn-1 .. 0
    if (nn == 2) {
        oi[1] = n_pay[0];
        os[1] = -n_pri[0];
        oi[0] = n_pay[1];
        os[0] = -n_pri[1];
    } else if (nn == 1) {
        oi[0] = n_pay[0];
        os[0] = -n_pri[0];
    }
}

// ---------------------------------------------------------------------------------------
// Exact-tie answers for byte colours and integral palettes: binary search in the palette's
// exception table (built by k_tie_scan at palette creation), KD-tree replay if there is none.
// ---------------------------------------------------------------------------------------
// Look a byte colour up in the exception table only: true (and scipy's answers in oi) if the
// colour has an exact tie among its three nearest rows, false if it is not in the table or the
// table was not built.
template <int KQ>
__device__ __forceinline__ bool tie_lookup(const PalDev *__restrict__ P, int r, int g, int b, int *oi)
{
    const int n = P->tie_n;
    if (n <= 0 || !P->tie_idx) return false;
    const unsigned key = (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16);
    int lo = (int)__ldg(P->tie_idx + (key >> 8));
    int hi = (int)__ldg(P->tie_idx + (key >> 8) + 1) - 1;
    if (hi < lo) return false;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((__ldg(&P->tie_table[mid].x) & 0xffffffu) < key) lo = mid + 1; else hi = mid;
    }
    const uint2 e = __ldg(P->tie_table + lo);
    if ((e.x & 0xffffffu) != key) return false;
    if (KQ == 1) {
        oi[0] = (int)(e.x >> 24);
    } else {
        oi[0] = (int)(e.y & 255u);
        oi[1] = (int)((e.y >> 8) & 255u);
    }
    return true;
}

template <int KQ>
__device__ __forceinline__ void tie_answer(const PalDev *__restrict__ P, int r, int g, int b, int *oi)
{
    const int n = P->tie_n;
    if (n > 0) {
        const unsigned key = (unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16);
        int lo = 0, hi = n - 1;
        if (P->tie_idx) {          // the bucket of (g, b): usually zero or one entry
            lo = (int)__ldg(P->tie_idx + (key >> 8));
            hi = (int)__ldg(P->tie_idx + (key >> 8) + 1) - 1;
            if (hi < lo) hi = lo = (lo < n ? lo : n - 1);   // empty bucket: one compare that fails
        }
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((__ldg(&P->tie_table[mid].x) & 0xffffffu) < key) lo = mid + 1; else hi = mid;
        }
        const uint2 e = __ldg(P->tie_table + lo);
        if ((e.x & 0xffffffu) == key) {
            if (KQ == 1) {
                oi[0] = (int)(e.x >> 24);
            } else {
                oi[0] = (int)(e.y & 255u);
                oi[1] = (int)((e.y >> 8) & 255u);
            }
            return;
        }
    }
    double os[2];
    kd_emulate<KQ>(P, (double)r, (double)g, (double)b, oi, os);
}

// ---------------------------------------------------------------------------------------
// Integer fast path: keys  key_i = (score_i << 8) | i  with
//   score_i = |p_i|^2 - 2 v.p_i = dist_i - |v|^2   (exact, |score| <= 195075)
// computed as three IMADs from the precomputed coefficients; top-3 kept by min/max.
// ---------------------------------------------------------------------------------------
struct Top3 {
    int m1, m2, m3;
};

__device__ __forceinline__ void top3_init(Top3 &t) { t.m1 = t.m2 = t.m3 = 0x7fffffff; }

__device__ __forceinline__ void top3_push(Top3 &t, int key)
{
    int a = max(t.m1, key);
    t.m1 = min(t.m1, key);
    int b = max(t.m2, a);
    t.m2 = min(t.m2, a);
    t.m3 = min(t.m3, b);
}

__device__ __forceinline__ int key_of(const int4 c, int r, int g, int b)
{
    return r * c.x + (g * c.y + (b * c.z + c.w));
}

// The f64 sequence of the reference's factor test (dithering_lib.py:361-365, 376):
//   dn = fl(fl(sqrt(n1))^2), ds likewise, factor = dn / (dn + ds) (0 if the sum is 0),
//   nearest iff factor <= threshold (f32 promoted to f64).
__device__ __forceinline__ bool factor_le_f64(double n1, double n2, float thr)
{
    double a = __dsqrt_rn(n1);
    double b = __dsqrt_rn(n2);
    double dn = __dmul_rn(a, a);
    double ds = __dmul_rn(b, b);
    double tot = __dadd_rn(dn, ds);
    double f = (tot == 0.0) ? 0.0 : __ddiv_rn(dn, tot);
    return f <= (double)thr;
}

// Exact decision for integer squared distances 1 <= n1 <= n2 < 2^19 against an f32 threshold:
// sign of n1 - T*(n1+n2) from an error-free product; only exact equality (or a threshold small
// enough to underflow the product) needs the f64 sequence above.
__device__ __forceinline__ bool factor_le_int(int n1, int n2, float thr)
{
    if (n1 == 0) return true;  // factor == 0 <= any threshold the reference can produce (>= 0)
    float N = (float)(n1 + n2);
    if (!(thr >= 1e-6f)) return factor_le_f64((double)n1, (double)n2, thr);
    float p = __fmul_rn(thr, N);
    float e = __fmaf_rn(thr, N, -p);      // exact residual T*N - p
    float d = __fsub_rn((float)n1, p);    // exact when n1 and p are within 2x, else sign-safe
    if (d < e) return true;               // n1/(n1+n2) < T
    if (d > e) return false;
    return factor_le_f64((double)n1, (double)n2, thr);
}
