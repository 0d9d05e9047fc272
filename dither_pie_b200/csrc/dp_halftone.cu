// dp_halftone.cu -- newspaper halftone.
//
// Replaces HalftoneDitherStrategy.dither (dithering_lib.py:1597-1644) and
// _generate_halftone_screen_with_cells (:1646-1695).
//
//   pass 0 (frame-invariant)  rotated-grid cell id and dot screen per pixel, in f64 with one
//                             rounding per numpy ufunc (no transcendental when dot_gain == 1)
//   pass 1                    exact integer RGB sums and counts per cell (atomics)
//   pass 2                    per cell: mean colour in f64 -> KD-tree nearest palette row
//   pass 3                    per pixel: ink (cell colour) where 1 - gray/255 > screen, else paper
// Algorithmic bytes: 3 read + 3 written per pixel (the maps are an implementation cost).
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
    float *screen;
    int *cell;
    unsigned long long *sums;   // [frames][ncells][2]: (sum r | sum g << 32), (sum b | count << 32)
    uint8_t *cell_pal;   // [frames][ncells]
    int paper;
    int make_screen;
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
    if (!p.make_screen) return;
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
    p.screen[i] = __double2float_rn(t);
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
        const float dark = __fsub_rn(1.0f, __fdiv_rn(gray, 255.0f));
        const int idx = (dark > __ldg(p.screen + i)) ? cp[__ldg(p.cell + i)] : p.paper;
        uint8_t *o = dst + (size_t)i * 3;
        o[0] = s_orgb[4 * idx];
        o[1] = s_orgb[4 * idx + 1];
        o[2] = s_orgb[4 * idx + 2];
        if (p.dst_idx) p.dst_idx[(size_t)f * p.npix + i] = (uint8_t)idx;
    }
}

// Four pixels per thread with word accesses (frames whose pixel count is a multiple of 4 and
// 4-byte aligned buffers): 12 bytes in, 12 bytes out, the screen and cell maps as 128-bit loads.
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
            const float r = (float)s_lut[v[k] & 255u], g = (float)s_lut[(v[k] >> 8) & 255u],
                        b = (float)s_lut[(v[k] >> 16) & 255u];
            const float gray = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)),
                                         __fmul_rn(0.114f, b));
            const float dark = __fsub_rn(1.0f, __fdiv_rn(gray, 255.0f));
            const int idx = (dark > scr[k]) ? (int)__ldg(cp + cel[k]) : p.paper;
            col[k] = s_orgb[idx];
            idx4 |= (unsigned)idx << (8 * k);
        }
        __stcs(dst + 3 * gi, col[0] | (col[1] << 24));
        __stcs(dst + 3 * gi + 1, (col[1] >> 8) | (col[2] << 16));
        __stcs(dst + 3 * gi + 2, (col[2] >> 16) | (col[3] << 8));
        if (p.dst_idx) reinterpret_cast<unsigned *>(p.dst_idx + (size_t)f * p.npix)[gi] = idx4;
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

}  // namespace

extern "C" int dp_halftone(const dp_palette *pal, const uint8_t *src_rgb, int frames, int h, int w,
                           int cell_size, double cos_a, double sin_a, double dot_gain,
                           double min_dot, double max_dot, int shape, double sharpness,
                           const float *screen, uint8_t *dst_rgb, uint8_t *dst_idx, void *stream)
{
    DP_REQUIRE(pal && src_rgb && dst_rgb, "null argument");
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

    Ws w_screen, w_cell, w_sums, w_cp;
    w_screen.st = w_cell.st = w_sums.st = w_cp.st = st;
    if (p.make_screen) {
        DP_CUDA(cudaMallocAsync(&w_screen.ptr, (size_t)p.npix * 4, st));
        p.screen = static_cast<float *>(w_screen.ptr);
    } else {
        p.screen = const_cast<float *>(screen);
    }
    DP_CUDA(cudaMallocAsync(&w_cell.ptr, (size_t)p.npix * 4, st));
    p.cell = static_cast<int *>(w_cell.ptr);
    size_t sums_bytes = (size_t)frames * p.ncells * 2 * sizeof(unsigned long long);
    DP_CUDA(cudaMallocAsync(&w_sums.ptr, sums_bytes, st));
    p.sums = static_cast<unsigned long long *>(w_sums.ptr);
    DP_CUDA(cudaMallocAsync(&w_cp.ptr, (size_t)frames * p.ncells, st));
    p.cell_pal = static_cast<uint8_t *>(w_cp.ptr);
    DP_CUDA(cudaMemsetAsync(p.sums, 0, sums_bytes, st));

    const int sms = dp_num_sms();
    k_ht_maps<<<(p.npix + 255) / 256, 256, 0, st>>>(p);
    DP_LAUNCH_CHECK();
    int gx = (p.npix + 255) / 256;
    if (gx > sms * 8) gx = sms * 8;
    k_ht_sums<<<dim3(gx, frames), 256, 0, st>>>(p);
    DP_LAUNCH_CHECK();
    int gc = (p.ncells + 127) / 128;
    if (gc > sms * 8) gc = sms * 8;
    k_ht_cells<<<dim3(gc, frames), 128, 0, st>>>(p);
    DP_LAUNCH_CHECK();
    const bool vec4 = p.npix % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(src_rgb) | reinterpret_cast<uintptr_t>(dst_rgb) |
                        reinterpret_cast<uintptr_t>(dst_idx) | reinterpret_cast<uintptr_t>(p.screen)) & 15) == 0;
    if (vec4) {
        int g4 = (p.npix / 4 + 255) / 256;
        if (g4 > sms * 8) g4 = sms * 8;
        k_ht_select4<<<dim3(g4, frames), 256, 0, st>>>(p);
    } else {
        k_ht_select<<<dim3(gx, frames), 256, 0, st>>>(p);
    }
    DP_LAUNCH_CHECK();
    return 0;
}
