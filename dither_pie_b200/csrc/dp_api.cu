// dp_api.cu -- library plumbing and the palette handle of libditherpie_b200.
#include <stdarg.h>

#include <vector>

#include <algorithm>
#include <array>
#include <map>
#include <mutex>

#include <nvtx3/nvToolsExt.h>

#include "dp_search.cuh"

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void dp_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *dp_last_error(void) { return g_err; }
extern "C" int dp_version(void) { return 100; }

int dp_num_sms()
{
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        cached_dev = dev;
    }
    return cached;
}

// The per-call workspaces (hand-off streams, halftone maps) come from the stream-ordered pool.
// Its default release threshold is 0, i.e. every synchronise hands the memory back to the OS and
// the next call pays for a fresh several-hundred-MB allocation; keep it instead.
int dp_retain_pool(int device)
{
    static thread_local unsigned long long done_mask = 0;
    if (device >= 0 && device < 64 && (done_mask >> device & 1ull)) return 0;
    cudaMemPool_t pool;
    DP_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    DP_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    if (device >= 0 && device < 64) done_mask |= 1ull << device;
    return 0;
}

extern "C" int dp_device_count(int *count)
{
    DP_REQUIRE(count, "null argument");
    DP_CUDA(cudaGetDeviceCount(count));
    return 0;
}

extern "C" int dp_set_device(int device)
{
    DP_CUDA(cudaSetDevice(device));
    int major = 0;
    DP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    DP_REQUIRE(major >= 10, "libditherpie_b200 needs an sm_100a (B200) device");
    return dp_retain_pool(device);
}

extern "C" int dp_malloc(void **dptr, size_t bytes)
{
    DP_REQUIRE(dptr, "null argument");
    DP_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return 0;
}
extern "C" int dp_free(void *dptr)
{
    DP_CUDA(cudaFree(dptr));
    return 0;
}
extern "C" int dp_host_alloc(void **hptr, size_t bytes)
{
    DP_REQUIRE(hptr, "null argument");
    DP_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}
extern "C" int dp_host_free(void *hptr)
{
    DP_CUDA(cudaFreeHost(hptr));
    return 0;
}
extern "C" int dp_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    DP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dp_stream(stream)));
    return 0;
}
extern "C" int dp_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    DP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dp_stream(stream)));
    return 0;
}
extern "C" int dp_memset(void *dst, int value, size_t bytes, void *stream)
{
    DP_CUDA(cudaMemsetAsync(dst, value, bytes, dp_stream(stream)));
    return 0;
}
extern "C" int dp_stream_create(void **stream)
{
    DP_REQUIRE(stream, "null argument");
    cudaStream_t s;
    DP_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return 0;
}
extern "C" int dp_stream_destroy(void *stream)
{
    DP_CUDA(cudaStreamDestroy(dp_stream(stream)));
    return 0;
}
extern "C" int dp_stream_sync(void *stream)
{
    DP_CUDA(cudaStreamSynchronize(dp_stream(stream)));
    return 0;
}
extern "C" int dp_event_create(void **event, int timing)
{
    DP_REQUIRE(event, "null argument");
    cudaEvent_t e;
    DP_CUDA(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
    *event = e;
    return 0;
}
extern "C" int dp_event_destroy(void *event)
{
    DP_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}
extern "C" int dp_event_record(void *event, void *stream)
{
    DP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), dp_stream(stream)));
    return 0;
}
extern "C" int dp_stream_wait_event(void *stream, void *event)
{
    DP_CUDA(cudaStreamWaitEvent(dp_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
    return 0;
}
extern "C" int dp_event_sync(void *event)
{
    DP_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}
extern "C" int dp_event_elapsed_ms(void *start, void *stop, float *ms)
{
    DP_REQUIRE(ms, "null argument");
    DP_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(stop)));
    return 0;
}
extern "C" int dp_host_register(void *hptr, size_t bytes)
{
    DP_REQUIRE(hptr && bytes, "null argument");
    DP_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return 0;
}
extern "C" int dp_host_is_pinned(const void *hptr, int *pinned)
{
    DP_REQUIRE(hptr && pinned, "null argument");
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, hptr);
    if (e != cudaSuccess) {
        cudaGetLastError();   // unregistered memory may report an error on older drivers
        *pinned = 0;
        return 0;
    }
    *pinned = at.type == cudaMemoryTypeHost ? 1 : 0;
    return 0;
}
extern "C" int dp_host_unregister(void *hptr)
{
    DP_CUDA(cudaHostUnregister(hptr));
    return 0;
}
// NVTX ranges around the entry points (the reference has no tracing at all);
// no-ops unless a profiler that collects NVTX is attached.
extern "C" int dp_range_push(const char *name)
{
    nvtxRangePushA(name ? name : "dp");
    return 0;
}
extern "C" int dp_range_pop(void)
{
    nvtxRangePop();
    return 0;
}

namespace {

// Exhaustive top-2 candidate masks: one block per cell of (1<<shift)^3 byte colours.  A palette
// row is a candidate of the cell if, for SOME colour of the cell, its distance is <= the
// second-smallest distance (so every row that can be nearest or second nearest, ties included);
// with nearest_only: <= the smallest distance (every row that can be nearest, ties included).
__global__ void __launch_bounds__(256) k_thr_masks(const int4 *__restrict__ coef, int K, int shift,
                                                   uint32_t *__restrict__ masks,
                                                   int nearest_only = 0)
{
    __shared__ int4 s_coef[DP_MAX_COLORS];
    __shared__ uint32_t s_mask[8];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_coef[i] = coef[i];   // (launched with 64..256 threads)
    if (threadIdx.x < 8) s_mask[threadIdx.x] = 0;
    __syncthreads();
    const int n = 256 >> shift;            // cells per axis
    const int side = 1 << shift;           // colours per axis in a cell
    const int cell = blockIdx.x;
    const int cr = cell / (n * n), cg = (cell / n) % n, cb = cell % n;
    uint32_t local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int total = side * side * side;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int r = (cr << shift) + t / (side * side);
        const int g = (cg << shift) + (t / side) % side;
        const int b = (cb << shift) + t % side;
        int m1 = 0x7fffffff, m2 = 0x7fffffff;
        for (int i = 0; i < K; ++i) {
            const int4 c = s_coef[i];
            const int s = (r * c.x + g * c.y + b * c.z + c.w) >> 8;  // score (idx stripped)
            const int a = max(m1, s);
            m1 = min(m1, s);
            m2 = min(m2, a);
        }
        for (int i = 0; i < K; ++i) {
            const int4 c = s_coef[i];
            const int s = (r * c.x + g * c.y + b * c.z + c.w) >> 8;
            if (s <= (nearest_only ? m1 : m2)) local[i >> 5] |= 1u << (i & 31);
        }
    }
    for (int wd = 0; wd < 8; ++wd)
        if (local[wd]) atomicOr(&s_mask[wd], local[wd]);
    __syncthreads();
    if (threadIdx.x < 8) masks[(size_t)cell * 8 + threadIdx.x] = s_mask[threadIdx.x];
}

// Nearest-row candidates per 8x8x8 box of colour space (table format: dp_common.cuh) by
// pairwise dominance: row j dominates row i on the box iff
//   max_{x in box} (|x-p_j|^2 - |x-p_i|^2) < 0;
// the expression is linear in x, so the maximum sits at the corner picked coordinate-wise by
// the sign of (p_j - p_i).  One block per cell; the result is a 256-bit mask of survivors.
// `unbounded`: the boxes of the outermost cells extend to infinity on their outer sides, so that
// the cell of a CLAMPED point lists every row that can be nearest to the unclamped point (the
// modes that look up unclamped work values: perceptual, adaptive variance).  The maximum over a
// half-infinite box is +inf unless the expression does not grow in that direction.
__global__ void __launch_bounds__(256) k_ed_masks(const double *__restrict__ pal, int K,
                                                  uint32_t *__restrict__ masks, int unbounded)
{
    __shared__ double s_p[DP_MAX_COLORS * 3];
    __shared__ double s_n[DP_MAX_COLORS];
    __shared__ uint32_t s_mask[8];
    const int cell = blockIdx.x;
    for (int i = threadIdx.x; i < K; i += 256) {
        const double a = pal[3 * i], b = pal[3 * i + 1], c = pal[3 * i + 2];
        s_p[3 * i] = a;
        s_p[3 * i + 1] = b;
        s_p[3 * i + 2] = c;
        s_n[i] = a * a + b * b + c * c;
    }
    if (threadIdx.x < 8) s_mask[threadIdx.x] = 0;
    __syncthreads();
    const double lo[3] = {8.0 * (cell >> 10), 8.0 * ((cell >> 5) & 31), 8.0 * (cell & 31)};
    const int ci[3] = {cell >> 10, (cell >> 5) & 31, cell & 31};
    const int i = threadIdx.x;
    if (i < K) {
        bool dominated = false;
        for (int j = 0; j < K && !dominated; ++j) {
            if (j == i) continue;
            double mx = s_n[j] - s_n[i];
            for (int c = 0; c < 3; ++c) {
                const double dlt = s_p[3 * j + c] - s_p[3 * i + c];
                const double x = dlt > 0.0 ? lo[c] : lo[c] + 8.0;
                mx += -2.0 * x * dlt;
                if (unbounded && ((dlt > 0.0 && ci[c] == 0) || (dlt < 0.0 && ci[c] == 31))) mx = 1e300;
            }
            dominated = mx < -1e-6;
        }
        if (!dominated) atomicOr(&s_mask[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    if (threadIdx.x < 8) masks[(size_t)cell * 8 + threadIdx.x] = s_mask[threadIdx.x];
}

// masks -> device tables (host side; palette creation is set-up)
int build_ed_table(const double *d_pal64, int K, PalDev &d, void **out_table, void **out_ovf, int unbounded)
{
    const int cells = 32768;
    uint32_t *dmask = nullptr;
    std::vector<uint32_t> hmask((size_t)cells * 8);
    if (cudaMalloc(&dmask, (size_t)cells * 32) != cudaSuccess) return 1;
    k_ed_masks<<<cells, 256>>>(d_pal64, K, dmask, unbounded);
    bool ok = cudaMemcpy(hmask.data(), dmask, (size_t)cells * 32, cudaMemcpyDeviceToHost) ==
              cudaSuccess;
    cudaFree(dmask);
    if (!ok) return 1;
    std::vector<uint16_t> l1(cells);
    int gt4 = 0;
    std::vector<uint4> pats;
    std::map<std::array<uint32_t, 4>, int> seen;
    std::vector<int> ocells;
    std::vector<uint32_t> ooff;
    std::vector<uint8_t> olist;
    for (int c = 0; c < cells; ++c) {
        uint8_t cand[DP_MAX_COLORS];
        int cnt = 0;
        for (int i = 0; i < K; ++i)
            if (hmask[(size_t)c * 8 + (i >> 5)] >> (i & 31) & 1u) cand[cnt++] = (uint8_t)i;
        unsigned slot[8];
        for (int k = 0; k < 8; ++k) slot[k] = (k < cnt) ? (unsigned)cand[k] * 16u : DP_ED_PAD;
        if (cnt > 4) ++gt4;
        if (cnt > 7) {
            slot[7] = DP_ED_OVERFLOW;
            ocells.push_back(c);
            ooff.push_back((uint32_t)olist.size());
            olist.insert(olist.end(), cand, cand + cnt);
        }
        const std::array<uint32_t, 4> key = {slot[0] | (slot[1] << 16), slot[2] | (slot[3] << 16),
                                             slot[4] | (slot[5] << 16), slot[6] | (slot[7] << 16)};
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, (int)pats.size()).first;
            pats.push_back(make_uint4(key[0], key[1], key[2], key[3]));
        }
        l1[c] = (uint16_t)it->second;   // at most 32768 patterns
    }
    ooff.push_back((uint32_t)olist.size());
    const size_t n = ocells.size();
    const size_t o_cells = 0, o_off = (n * 4 + 15) / 16 * 16, o_list = o_off + ((n + 1) * 4 + 15) / 16 * 16;
    std::vector<uint8_t> blob(o_list + olist.size() + 16, 0);
    if (n) memcpy(blob.data() + o_cells, ocells.data(), n * 4);
    memcpy(blob.data() + o_off, ooff.data(), (n + 1) * 4);
    if (!olist.empty()) memcpy(blob.data() + o_list, olist.data(), olist.size());
    void *dt = nullptr, *dov = nullptr;
    const size_t l1_bytes = (size_t)cells * 2;
    const size_t pat_bytes = pats.size() * 16;
    std::vector<uint4> flat(cells);
    for (int c = 0; c < cells; ++c) flat[c] = pats[l1[c]];
    ok = cudaMalloc(&dt, l1_bytes + pat_bytes + (size_t)cells * 16) == cudaSuccess &&
         cudaMemcpy(static_cast<uint8_t *>(dt) + l1_bytes + pat_bytes, flat.data(), (size_t)cells * 16,
                    cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMalloc(&dov, blob.size()) == cudaSuccess &&
         cudaMemcpy(dt, l1.data(), l1_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(static_cast<uint8_t *>(dt) + l1_bytes, pats.data(), pats.size() * 16,
                    cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(dov, blob.data(), blob.size(), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        if (dt) cudaFree(dt);
        if (dov) cudaFree(dov);
        return 1;
    }
    d.ed_l1 = static_cast<const uint16_t *>(dt);
    d.ed_pat = reinterpret_cast<const uint4 *>(static_cast<uint8_t *>(dt) + l1_bytes);
    d.ed_npat = (int)pats.size();
    d.ed_gt4 = gt4;
    d.ed_flat = reinterpret_cast<const uint4 *>(static_cast<uint8_t *>(dt) + l1_bytes + pat_bytes);
    d.ed_ovf_cells = reinterpret_cast<const int *>(static_cast<uint8_t *>(dov) + o_cells);
    d.ed_ovf_off = reinterpret_cast<const uint32_t *>(static_cast<uint8_t *>(dov) + o_off);
    d.ed_ovf = static_cast<uint8_t *>(dov) + o_list;
    d.ed_novf = (int)n;
    *out_table = dt;
    *out_ovf = dov;
    return 0;
}

// Every byte colour with an exact distance tie among its three nearest rows -> scipy's answers
// for query(k=1) and query(k=2) (format: dp_common.cuh, tie_table).  One block per (r, g), one
// thread per b.  `count` may run past `cap`; the host then discards the table.
__global__ void __launch_bounds__(256) k_tie_scan(const PalDev *__restrict__ P, int K, uint2 *out,
                                                  unsigned cap, unsigned *count)
{
    __shared__ int4 s_coef[DP_MAX_COLORS];
    for (int i = threadIdx.x; i < K; i += 256) s_coef[i] = P->coef[i];
    __syncthreads();
    const int r = blockIdx.x >> 8, g = blockIdx.x & 255, b = threadIdx.x;
    Top3 t;
    top3_init(t);
    for (int i = 0; i < K; ++i) top3_push(t, key_of(s_coef[i], r, g, b));
    const int s1 = t.m1 >> 8, s2 = t.m2 >> 8, s3 = t.m3 >> 8;
    if (!(s1 == s2 || (K >= 3 && s2 == s3))) return;
    int o1[1], o2[2];
    double os[2];
    kd_emulate<1>(P, (double)r, (double)g, (double)b, o1, os);
    kd_emulate<2>(P, (double)r, (double)g, (double)b, o2, os);
    const unsigned pos = atomicAdd(count, 1u);
    if (pos < cap)
        out[pos] = make_uint2((unsigned)r | ((unsigned)g << 8) | ((unsigned)b << 16) |
                                  ((unsigned)o1[0] << 24),
                              (unsigned)o2[0] | ((unsigned)o2[1] << 8));
}

int build_tie_table(PalDev &d, dp_palette *h, const void *dev_paldev, int K)
{
    d.tie_table = nullptr;
    d.tie_idx = nullptr;
    d.tie_n = -1;
    const unsigned cap = 1u << 21;   // 2M tie colours (16 MB); beyond that replay in the kernels
    uint2 *dout = nullptr;
    unsigned *dcnt = nullptr;
    if (cudaMalloc(&dout, (size_t)cap * 8) != cudaSuccess) return 0;
    if (cudaMalloc(&dcnt, 4) != cudaSuccess) {
        cudaFree(dout);
        return 0;
    }
    cudaMemset(dcnt, 0, 4);
    k_tie_scan<<<65536, 256>>>(static_cast<const PalDev *>(dev_paldev), K, dout, cap, dcnt);
    unsigned n = 0;
    bool ok = cudaMemcpy(&n, dcnt, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(dcnt);
    if (!ok || n > cap) {
        cudaFree(dout);
        return ok ? 0 : 1;
    }
    std::vector<uint2> host(n);
    if (n && cudaMemcpy(host.data(), dout, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaFree(dout);
        return 1;
    }
    cudaFree(dout);
    std::sort(host.begin(), host.end(),
              [](const uint2 &a, const uint2 &b) { return (a.x & 0xffffffu) < (b.x & 0xffffffu); });
    void *dt = nullptr;
    if (cudaMalloc(&dt, (size_t)(n ? n : 1) * 8) != cudaSuccess) return 0;
    if (n && cudaMemcpy(dt, host.data(), (size_t)n * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(dt);
        return 1;
    }
    d.tie_table = static_cast<const uint2 *>(dt);
    d.tie_n = (int)n;
    h->tie_table = dt;
    // bucket index by the two high bytes of the colour key (the table is sorted by the key)
    if (n) {
        std::vector<uint32_t> idx(65537, 0);
        for (unsigned i = 0; i < n; ++i) idx[((host[i].x & 0xffffffu) >> 8) + 1]++;
        for (int k = 0; k < 65536; ++k) idx[k + 1] += idx[k];
        void *di = nullptr;
        if (cudaMalloc(&di, idx.size() * 4) == cudaSuccess &&
            cudaMemcpy(di, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess) {
            d.tie_idx = static_cast<const uint32_t *>(di);
            h->tie_idx = di;
        } else if (di) {
            cudaFree(di);
        }
    }
    return 0;
}

// Compact 32^3 candidate table for k_thresh_v4 (format: dp_common.cuh): `slots` candidates per
// u32 entry, cells with more get eight 4x4x4 sub-cell entries.  K <= 30.  Returns false if the
// table cannot be built (the older kernels then do the work).
bool build_compact_table(const int4 *d_coef, int K, int nearest_only, int slots, void **out_table,
                         void **out_sub, int *out_nsub)
{
    std::vector<uint32_t> m3((size_t)32768 * 8);
    uint32_t *dm = nullptr;
    bool ok = cudaMalloc(&dm, (size_t)32768 * 32) == cudaSuccess;
    if (ok) {
        k_thr_masks<<<32768, 256>>>(d_coef, K, 3, dm, nearest_only);
        ok = cudaMemcpy(m3.data(), dm, (size_t)32768 * 32, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (dm) cudaFree(dm);
    if (!ok) return false;
    auto pack = [K, slots](uint32_t w0, bool &over) {
        unsigned slot[4] = {(unsigned)K * 8u, (unsigned)K * 8u, (unsigned)K * 8u, 0u};
        if (slots == 4) slot[3] = (unsigned)K * 8u;
        int cnt = 0;
        for (int i = 0; i < K; ++i)
            if (w0 >> i & 1u) {
                if (cnt < slots) slot[cnt] = (unsigned)i * 8u;
                ++cnt;
            }
        over = cnt > slots;
        if (over) slot[3] = 0xf8u;   // row 31: a pad row, and the overflow mark
        return slot[0] | (slot[1] << 8) | (slot[2] << 16) | (slot[3] << 24);
    };
    std::vector<uint32_t> t4(32768), sub;
    std::vector<int> ocell;
    for (int c = 0; c < 32768; ++c) {
        bool over;
        t4[c] = pack(m3[(size_t)c * 8], over);
        if (over) {
            t4[c] = 0xf8000000u | (uint32_t)ocell.size();
            ocell.push_back(c);
        }
    }
    if (ocell.size() > 8192) return false;
    if (!ocell.empty()) {
        // masks of the 4x4x4 sub-cells (64^3 grid), only read for the crowded cells
        uint32_t *dm2 = nullptr;
        std::vector<uint32_t> m2((size_t)262144 * 8);
        ok = cudaMalloc(&dm2, (size_t)262144 * 32) == cudaSuccess;
        if (ok) {
            k_thr_masks<<<262144, 64>>>(d_coef, K, 2, dm2, nearest_only);
            ok = cudaMemcpy(m2.data(), dm2, (size_t)262144 * 32, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        if (dm2) cudaFree(dm2);
        if (!ok) return false;
        sub.resize(ocell.size() * 8);
        for (size_t n = 0; n < ocell.size(); ++n) {
            const int c = ocell[n], cr = c >> 10, cg = (c >> 5) & 31, cb = c & 31;
            for (int s8 = 0; s8 < 8; ++s8) {
                const int fr = cr * 2 + (s8 >> 2), fg = cg * 2 + ((s8 >> 1) & 1), fb = cb * 2 + (s8 & 1);
                bool over;
                sub[n * 8 + s8] = pack(m2[((size_t)(fr * 64 + fg) * 64 + fb) * 8], over);
            }
        }
    }
    void *d4 = nullptr, *dsub = nullptr;
    ok = cudaMalloc(&d4, 32768 * 4) == cudaSuccess &&
         cudaMemcpy(d4, t4.data(), 32768 * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMalloc(&dsub, sub.size() * 4 + 16) == cudaSuccess &&
         (sub.empty() ||
          cudaMemcpy(dsub, sub.data(), sub.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess);
    if (!ok) {
        if (d4) cudaFree(d4);
        if (dsub) cudaFree(dsub);
        return false;
    }
    *out_table = d4;
    *out_sub = dsub;
    *out_nsub = (int)ocell.size();
    return true;
}

// The same tables for 31 <= K <= 256 (format: PalDev::thr4_wide in dp_common.cuh): plain row
// numbers, `slots` distinct ascending rows per entry (filled up with non-candidates), sub-cell
// entries for the cells with more candidates.
bool build_compact_table_wide(const int4 *d_coef, int K, int nearest_only, int slots, void **out_table,
                              void **out_sub, int *out_nsub)
{
    std::vector<uint32_t> m3((size_t)32768 * 8);
    uint32_t *dm = nullptr;
    bool ok = cudaMalloc(&dm, (size_t)32768 * 32) == cudaSuccess;
    if (ok) {
        k_thr_masks<<<32768, 256>>>(d_coef, K, 3, dm, nearest_only);
        ok = cudaMemcpy(m3.data(), dm, (size_t)32768 * 32, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    if (dm) cudaFree(dm);
    if (!ok) return false;
    const uint32_t marker_top = slots == 4 ? 0u : 0u;   // markers have a top byte of 0
    auto pack = [K, slots](const uint32_t *mask, bool &over) -> uint32_t {
        int rows[DP_MAX_COLORS];
        int cnt = 0;
        for (int i = 0; i < K; ++i)
            if (mask[i >> 5] >> (i & 31) & 1u) rows[cnt++] = i;
        over = cnt > slots;
        if (over) return 0u;
        for (int i = 0; cnt < slots && i < K; ++i) {      // fill up with the lowest non-candidates
            bool have = false;
            for (int j = 0; j < cnt; ++j) have = have || rows[j] == i;
            if (!have) rows[cnt++] = i;
        }
        std::sort(rows, rows + slots);
        uint32_t e = 0;
        for (int j = 0; j < slots; ++j) e |= (uint32_t)rows[j] << (8 * j);
        if (slots == 3) e |= 0xff000000u;
        return e;
    };
    (void)marker_top;
    std::vector<uint32_t> t4(32768), sub;
    std::vector<int> ocell;
    for (int c = 0; c < 32768; ++c) {
        bool over;
        t4[c] = pack(&m3[(size_t)c * 8], over);
        if (over) {
            t4[c] = (uint32_t)ocell.size();     // < 0x03000000 (and < 0xff000000): the marker
            ocell.push_back(c);
        }
    }
    if (!ocell.empty()) {
        uint32_t *dm2 = nullptr;
        std::vector<uint32_t> m2((size_t)262144 * 8);
        ok = cudaMalloc(&dm2, (size_t)262144 * 32) == cudaSuccess;
        if (ok) {
            k_thr_masks<<<262144, 64>>>(d_coef, K, 2, dm2, nearest_only);
            ok = cudaMemcpy(m2.data(), dm2, (size_t)262144 * 32, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        if (dm2) cudaFree(dm2);
        if (!ok) return false;
        sub.resize(ocell.size() * 8);
        for (size_t n = 0; n < ocell.size(); ++n) {
            const int c = ocell[n], cr = c >> 10, cg = (c >> 5) & 31, cb = c & 31;
            for (int s8 = 0; s8 < 8; ++s8) {
                const int fr = cr * 2 + (s8 >> 2), fg = cg * 2 + ((s8 >> 1) & 1), fb = cb * 2 + (s8 & 1);
                bool over;
                sub[n * 8 + s8] = pack(&m2[((size_t)(fr * 64 + fg) * 64 + fb) * 8], over);   // 0 if still crowded
            }
        }
    }
    void *d4 = nullptr, *dsub = nullptr;
    ok = cudaMalloc(&d4, 32768 * 4) == cudaSuccess &&
         cudaMemcpy(d4, t4.data(), 32768 * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMalloc(&dsub, sub.size() * 4 + 16) == cudaSuccess &&
         (sub.empty() ||
          cudaMemcpy(dsub, sub.data(), sub.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess);
    if (!ok) {
        if (d4) cudaFree(d4);
        if (dsub) cudaFree(dsub);
        return false;
    }
    *out_table = d4;
    *out_sub = dsub;
    *out_nsub = (int)ocell.size();
    return true;
}

template <typename T>
size_t put(std::vector<uint8_t> &buf, const T *src, size_t n, size_t align = 16)
{
    size_t o = (buf.size() + align - 1) / align * align;
    buf.resize(o + n * sizeof(T));
    if (src) memcpy(buf.data() + o, src, n * sizeof(T));
    return o;
}

}  // namespace

extern "C" int dp_palette_create(const float *palette, int K, const uint8_t *out_rgb,
                                 const uint8_t *in_lut, int kd_nodes,
                                 const int32_t *kd_split_dim, const double *kd_split,
                                 const int32_t *kd_start_idx, const int32_t *kd_end_idx,
                                 const int32_t *kd_lesser, const int32_t *kd_greater,
                                 const int32_t *kd_indices, const double *kd_mins,
                                 const double *kd_maxes, dp_palette **out)
{
    DP_RANGE("dp_palette_create");
    DP_REQUIRE(palette && out_rgb && out, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS, "palette size must be 1..256");
    DP_REQUIRE(kd_nodes >= 1 && kd_split_dim && kd_split && kd_start_idx && kd_end_idx &&
                   kd_lesser && kd_greater && kd_indices && kd_mins && kd_maxes,
               "KD-tree arrays missing");
    int inner = 0;
    for (int i = 0; i < kd_nodes; ++i) {
        if (kd_split_dim[i] != -1) {
            ++inner;
            DP_REQUIRE(kd_split_dim[i] >= 0 && kd_split_dim[i] < 3, "bad split_dim");
            DP_REQUIRE(kd_lesser[i] > 0 && kd_lesser[i] < kd_nodes && kd_greater[i] > 0 &&
                           kd_greater[i] < kd_nodes, "bad child index");
        } else {
            DP_REQUIRE(kd_start_idx[i] >= 0 && kd_end_idx[i] <= K &&
                           kd_start_idx[i] <= kd_end_idx[i], "bad leaf range");
        }
    }
    DP_REQUIRE(inner + 1 <= 72, "KD-tree has too many inner nodes for the device traversal");
    for (int i = 0; i < K; ++i)
        DP_REQUIRE(kd_indices[i] >= 0 && kd_indices[i] < K, "bad kd index");

    dp_palette *h = new dp_palette();
    memset(h, 0, sizeof(*h));
    DP_CUDA(cudaGetDevice(&h->device));

    bool integral = true;
    std::vector<double> p64(K * 3);
    std::vector<int4> coef(K);
    for (int i = 0; i < K * 3; ++i) {
        float v = palette[i];
        h->host_pal[i] = v;
        p64[i] = (double)v;
        if (!(v >= 0.0f && v <= 255.0f && v == (float)(int)v)) integral = false;
    }
    if (integral) {
        for (int i = 0; i < K; ++i) {
            int r = (int)palette[3 * i], g = (int)palette[3 * i + 1], b = (int)palette[3 * i + 2];
            coef[i].x = (-2 * r) * 256;
            coef[i].y = (-2 * g) * 256;
            coef[i].z = (-2 * b) * 256;
            coef[i].w = (r * r + g * g + b * b) * 256 + i;
        }
    }
    std::vector<uint8_t> orgb(K * 4, 0);
    for (int i = 0; i < K; ++i)
        for (int c = 0; c < 3; ++c) orgb[4 * i + c] = out_rgb[3 * i + c];
    h->pal_is_out = 1;
    for (int i = 0; i < K; ++i)
        for (int c = 0; c < 3; ++c)
            if (palette[3 * i + c] != (float)out_rgb[3 * i + c]) h->pal_is_out = 0;
    uint8_t lut[256];
    h->has_lut = 0;
    for (int i = 0; i < 256; ++i) {
        lut[i] = in_lut ? in_lut[i] : (uint8_t)i;
        if (lut[i] != i) h->has_lut = 1;
    }

    std::vector<uint8_t> buf;
    PalDev d;
    memset(&d, 0, sizeof(d));
    size_t o_self = put<PalDev>(buf, nullptr, 1, 256);
    size_t o_f32 = put(buf, palette, (size_t)K * 3);
    size_t o_f64 = put(buf, p64.data(), (size_t)K * 3);
    size_t o_coef = put(buf, coef.data(), (size_t)K);
    size_t o_orgb = put(buf, orgb.data(), (size_t)K * 4);
    size_t o_lut = put(buf, lut, 256);
    size_t o_sd = put(buf, kd_split_dim, (size_t)kd_nodes);
    size_t o_sp = put(buf, kd_split, (size_t)kd_nodes);
    size_t o_st = put(buf, kd_start_idx, (size_t)kd_nodes);
    size_t o_en = put(buf, kd_end_idx, (size_t)kd_nodes);
    size_t o_le = put(buf, kd_lesser, (size_t)kd_nodes);
    size_t o_gr = put(buf, kd_greater, (size_t)kd_nodes);
    size_t o_ix = put(buf, kd_indices, (size_t)K);
    (void)o_self;

    uint8_t *blob = nullptr;
    if (cudaMalloc(&blob, buf.size()) != cudaSuccess) {
        delete h;
        dp_set_error("cudaMalloc(%zu) failed for the palette blob", buf.size());
        return 1;
    }
    d.K = K;
    d.integral = integral ? 1 : 0;
    d.kd_nodes = kd_nodes;
    d.pal_f32 = reinterpret_cast<const float *>(blob + o_f32);
    d.pal_f64 = reinterpret_cast<const double *>(blob + o_f64);
    d.coef = reinterpret_cast<const int4 *>(blob + o_coef);
    d.out_rgb = blob + o_orgb;
    d.in_lut = blob + o_lut;
    d.kd_split_dim = reinterpret_cast<const int *>(blob + o_sd);
    d.kd_split = reinterpret_cast<const double *>(blob + o_sp);
    d.kd_start = reinterpret_cast<const int *>(blob + o_st);
    d.kd_end = reinterpret_cast<const int *>(blob + o_en);
    d.kd_lesser = reinterpret_cast<const int *>(blob + o_le);
    d.kd_greater = reinterpret_cast<const int *>(blob + o_gr);
    d.kd_indices = reinterpret_cast<const int *>(blob + o_ix);
    for (int i = 0; i < 3; ++i) {
        d.kd_mins[i] = kd_mins[i];
        d.kd_maxes[i] = kd_maxes[i];
    }
    h->blob = blob;

    if (cudaMemcpy(blob, buf.data(), buf.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        dp_set_error("palette upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(blob);
        delete h;
        return 1;
    }
    if (build_ed_table(d.pal_f64, K, d, &h->ed_table, &h->ed_ovf, 0)) {
        dp_set_error("palette nearest-row table build failed: %s",
                     cudaGetErrorString(cudaGetLastError()));
        cudaFree(blob);
        delete h;
        return 1;
    }
    if (integral && K >= 2) {
        // top-2 candidate table for the threshold kernels
        const int shift = (K <= 64) ? 4 : 3;
        const int n = 256 >> shift, cells = n * n * n;
        uint32_t *dmask = nullptr;
        std::vector<uint32_t> hmask((size_t)cells * 8);
        bool ok2 = cudaMalloc(&dmask, (size_t)cells * 32) == cudaSuccess;
        if (ok2) {
            k_thr_masks<<<cells, 256>>>(d.coef, K, shift, dmask);
            ok2 = cudaMemcpy(hmask.data(), dmask, (size_t)cells * 32, cudaMemcpyDeviceToHost) ==
                  cudaSuccess;
        }
        if (dmask) cudaFree(dmask);
        std::vector<uint2> table(cells);
        std::vector<uint8_t> ovf;
        for (int c = 0; ok2 && c < cells; ++c) {
            uint8_t cand[DP_MAX_COLORS];
            int cnt = 0;
            for (int i = 0; i < K; ++i)
                if (hmask[(size_t)c * 8 + (i >> 5)] >> (i & 31) & 1u) cand[cnt++] = (uint8_t)i;
            uint2 e;
            if (cnt <= 7) {
                uint64_t v = (uint64_t)cnt;
                for (int j = 0; j < cnt; ++j) v |= (uint64_t)cand[j] << (8 * (j + 1));
                e.x = (uint32_t)v;
                e.y = (uint32_t)(v >> 32);
            } else {
                e.x = 0xffu | ((uint32_t)cnt << 8);
                e.y = (uint32_t)ovf.size();
                ovf.insert(ovf.end(), cand, cand + cnt);
            }
            table[c] = e;
        }
        void *dt = nullptr, *dovf = nullptr;
        ok2 = ok2 && cudaMalloc(&dt, (size_t)cells * 8) == cudaSuccess &&
              cudaMalloc(&dovf, ovf.size() + 16) == cudaSuccess &&
              cudaMemcpy(dt, table.data(), (size_t)cells * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
              (ovf.empty() ||
               cudaMemcpy(dovf, ovf.data(), ovf.size(), cudaMemcpyHostToDevice) == cudaSuccess);
        if (!ok2) {
            dp_set_error("palette top-2 table build failed: %s",
                         cudaGetErrorString(cudaGetLastError()));
            if (dt) cudaFree(dt);
            if (dovf) cudaFree(dovf);
            cudaFree(h->ed_table);
            cudaFree(h->ed_ovf);
            cudaFree(blob);
            delete h;
            return 1;
        }
        d.thr_table = static_cast<const uint2 *>(dt);
        d.thr_ovf = static_cast<const uint8_t *>(dovf);
        d.thr_shift = shift;
        d.thr_cells = cells;
        h->thr_table = dt;
        h->thr_ovf = dovf;
        if (K <= 30) {
            // compact 32^3 tables for k_thresh_v4: top-2 candidates (4 slots) for the threshold
            // modes, nearest candidates (3 slots) for plain quantisation
            void *t = nullptr, *sb = nullptr;
            int ns = 0;
            if (build_compact_table(d.coef, K, 0, 4, &t, &sb, &ns) && ns <= 1024) {
                d.thr4_table = static_cast<const uint32_t *>(t);
                d.thr4_sub = static_cast<const uint32_t *>(sb);
                d.thr4_nsub = ns;
                h->thr4_table = t;
                h->thr4_sub = sb;
            } else {
                if (t) cudaFree(t);
                if (sb) cudaFree(sb);
            }
            t = sb = nullptr;
            if (build_compact_table(d.coef, K, 1, 3, &t, &sb, &ns)) {
                d.near3_table = static_cast<const uint32_t *>(t);
                d.near3_sub = static_cast<const uint32_t *>(sb);
                d.near3_nsub = ns;
                h->near3_table = t;
                h->near3_sub = sb;
            }
        } else {
            // 31..256 colours: the same kernel with plain row numbers (PalDev::thr4_wide)
            void *t = nullptr, *sb = nullptr, *t3 = nullptr, *sb3 = nullptr;
            int ns = 0, ns3 = 0;
            if (build_compact_table_wide(d.coef, K, 0, 4, &t, &sb, &ns) &&
                build_compact_table_wide(d.coef, K, 1, 3, &t3, &sb3, &ns3)) {
                d.thr4_table = static_cast<const uint32_t *>(t);
                d.thr4_sub = static_cast<const uint32_t *>(sb);
                d.thr4_nsub = ns;
                d.near3_table = static_cast<const uint32_t *>(t3);
                d.near3_sub = static_cast<const uint32_t *>(sb3);
                d.near3_nsub = ns3;
                d.thr4_wide = 1;
                h->thr4_table = t;
                h->thr4_sub = sb;
                h->near3_table = t3;
                h->near3_sub = sb3;
            } else {
                if (t) cudaFree(t);
                if (sb) cudaFree(sb);
                if (t3) cudaFree(t3);
                if (sb3) cudaFree(sb3);
            }
        }
    }
    d.tie_table = nullptr;
    d.tie_idx = nullptr;
    d.tie_n = -1;
    h->dev = d;
    if (cudaMemcpy(blob, &d, sizeof(d), cudaMemcpyHostToDevice) != cudaSuccess) {
        dp_set_error("palette upload failed");
        dp_palette_destroy(h);
        return 1;
    }
    if (integral && K >= 2) {
        // the scan reads the palette through the descriptor just uploaded
        if (build_tie_table(d, h, blob, K) ||
            cudaMemcpy(blob, &d, sizeof(d), cudaMemcpyHostToDevice) != cudaSuccess) {
            dp_set_error("palette tie table build failed: %s", cudaGetErrorString(cudaGetLastError()));
            dp_palette_destroy(h);
            return 1;
        }
        h->dev = d;
    }
    *out = h;
    return 0;
}

extern "C" int dp_palette_destroy(dp_palette *pal)
{
    if (!pal) return 0;
    if (pal->ed_table) cudaFree(pal->ed_table);
    if (pal->ed_ovf) cudaFree(pal->ed_ovf);
    if (pal->ext_table) cudaFree(pal->ext_table);
    if (pal->ext_ovf) cudaFree(pal->ext_ovf);
    if (pal->ext_dev) cudaFree(pal->ext_dev);
    if (pal->tie_idx) cudaFree(pal->tie_idx);
    if (pal->thr_table) cudaFree(pal->thr_table);
    if (pal->thr_ovf) cudaFree(pal->thr_ovf);
    if (pal->thr4_table) cudaFree(pal->thr4_table);
    if (pal->thr4_sub) cudaFree(pal->thr4_sub);
    if (pal->near3_table) cudaFree(pal->near3_table);
    if (pal->near3_sub) cudaFree(pal->near3_sub);
    if (pal->tie_table) cudaFree(pal->tie_table);
    if (pal->blob) cudaFree(pal->blob);
    delete pal;
    return 0;
}

// Second device copy of the palette descriptor whose nearest-row tables were built with unbounded
// outer cells (see k_ed_masks); built on first use, under a lock (palettes are shared by threads).
int dp_palette_ext(dp_palette *pal, const PalDev **dev, int *npat, int *gt4)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!pal->ext_dev) {
        PalDev dx = pal->dev;
        void *t = nullptr, *o = nullptr, *dd = nullptr;
        if (build_ed_table(pal->dev.pal_f64, pal->dev.K, dx, &t, &o, 1)) {
            dp_set_error("unbounded nearest-row table build failed: %s", cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        if (cudaMalloc(&dd, sizeof(PalDev)) != cudaSuccess ||
            cudaMemcpy(dd, &dx, sizeof(PalDev), cudaMemcpyHostToDevice) != cudaSuccess) {
            dp_set_error("palette descriptor upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            cudaFree(t);
            cudaFree(o);
            if (dd) cudaFree(dd);
            return 1;
        }
        pal->ext_table = t;
        pal->ext_ovf = o;
        pal->ext_npat = dx.ed_npat;
        pal->ext_gt4 = dx.ed_gt4;
        pal->ext_dev = dd;
    }
    *dev = static_cast<const PalDev *>(pal->ext_dev);
    *npat = pal->ext_npat;
    *gt4 = pal->ext_gt4;
    return 0;
}

extern "C" int dp_palette_num_colors(const dp_palette *pal) { return pal ? pal->dev.K : 0; }
