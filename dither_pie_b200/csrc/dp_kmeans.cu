// dp_kmeans.cu -- Lloyd iterations for k-means palette extraction.
//
// Replaces the E/M steps inside sklearn.cluster.KMeans as called by
// ColorReducer.generate_kmeans_palette (dithering_lib.py:1854-1855).  Pixels are u8, so the
// per-cluster channel sums are exact integers: accumulated in registers / shared memory per block,
// flushed as u64 global atomics and -- across GPUs -- all-reduced as integers (NCCL, on the same
// stream), which makes the centres independent of shard count and order.
// Algorithmic bytes: 3 read per pixel per iteration (labels are not materialised).
//
// One Lloyd iteration is two launches:
//   k_kmeans_prepare     every block recomputes the new centres from the previous iteration's
//                        (all-reduced) sums -- K*4 integers, cheaper than a separate launch --,
//                        block 0 publishes them with the squared centre shift (sklearn's stop
//                        quantity) and the stop flag; then the blocks rebuild the CANDIDATE GRID
//                        for the new centres (16^3 boxes of 16^3 byte colours -> the <= 4 centres
//                        that can be nearest somewhere in the box) and the fixed-point centre table
//   k_kmeans_accum16     the assignment pass: a lane owns 16 consecutive pixels (three 128-bit
//                        loads), labels them against the box's candidates in 32-bit fixed point,
//                        and keeps run-length partial sums that are merged by warp reductions /
//                        shared atomics; pixels the fixed-point screen cannot decide go to a
//                        per-warp list that the warp resolves cooperatively in f64 (lane = centre)
// The stop test stays on the device: once the flag is set the remaining launches of the batch the
// host enqueued return immediately, and the host looks at the flag only every few iterations.
#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "dp_common.cuh"

struct DpNcclId {   // ncclUniqueId: 128 opaque bytes, passed by value to ncclCommInitRank
    char internal[128];
};

namespace {

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_PIX_PER_BLOCK = 16384;  // 255 * 16384 < 2^32: u32 block partials are exact

struct KmState {
    int done;                  // stop flag (shift <= tol, or the iteration budget is used up)
    int n_iter;                // completed Lloyd iterations
    int empty;                 // iterations that saw an empty cluster (it keeps its centre)
    int error;                 // peer exchange timed out (dp_kmeans_lloyd_p2p)
    double shift2;             // squared centre shift of the last completed iteration
    unsigned long long ties;   // samples exactly equidistant from their two nearest centres
};

// ---- exchange of the integer sums over peer memory (NVLink), dp_kmeans_lloyd_p2p -----------------
// Every rank owns an INBOX in its own memory that the peers can write (cudaIpc):
//   sums  u64 [2 parities][KM_P2P_RANKS][KM_P2P_STRIDE]   slot [parity][r] <- rank r's K*4+1 sums
//   flags u64 [2 parities][KM_P2P_RANKS]                  <- the iteration number, written last
// The assignment kernel itself PUSHES: the block that finishes last (a ticket counter) stores the
// rank's sums into slot [it & 1][rank] of every inbox (its own included) with plain stores over
// NVLink, fences at system scope and then releases the flags; the
// next prepare launch waits until all `world` flags of that parity show the iteration and adds the
// slots up -- integers, so every rank obtains the same totals whatever the order.  No collective
// launch and no reduction tree between the two kernels of an iteration: the message is 65 integers
// and the cost is one NVLink store latency.
// A parity is rewritten two iterations later, which a rank can only reach after it has seen every
// peer's flag of the iteration in between -- and a peer raises that flag only after its own
// prepare launch (the reader of the older parity) has completed: double buffering suffices.
constexpr int KM_GRID_REPLICAS = 8;
constexpr int KM_P2P_RANKS = 8;
constexpr int KM_P2P_STRIDE = DP_MAX_COLORS * 4 + 8;      // u64 per slot (>= K*4+1)
constexpr size_t KM_P2P_FLAGS_OFF = (size_t)2 * KM_P2P_RANKS * KM_P2P_STRIDE;   // in u64
constexpr size_t KM_P2P_BYTES = (KM_P2P_FLAGS_OFF + 2 * KM_P2P_RANKS) * 8;

struct KmPeers {
    unsigned long long *inbox[KM_P2P_RANKS];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct KmPush {              // by-value kernel argument; world == 0: no exchange
    KmPeers peers;
    int rank, world, parity;
    unsigned long long iter;   // the tag the flags receive
    unsigned *ticket;          // blocks-finished counter (zero between launches)
};

__device__ __forceinline__ void km_push_sums(const unsigned long long *sums, int nsum, const KmPush &ps)
{
    // constant indices into the kernel argument (unrolled): it stays in the constant bank
    const size_t slot = ((size_t)ps.parity * KM_P2P_RANKS + ps.rank) * KM_P2P_STRIDE;
    for (int k = threadIdx.x; k < nsum; k += blockDim.x) {
        const unsigned long long v = __ldcg(sums + k);
#pragma unroll
        for (int p = 0; p < KM_P2P_RANKS; ++p)
            if (p < ps.world) ps.peers.inbox[p][slot + k] = v;
    }
    __threadfence_system();
    __syncthreads();
    const size_t flag = KM_P2P_FLAGS_OFF + (size_t)ps.parity * KM_P2P_RANKS + ps.rank;
#pragma unroll
    for (int p = 0; p < KM_P2P_RANKS; ++p)
        if ((int)threadIdx.x == p && p < ps.world) st_release_sys(ps.peers.inbox[p] + flag, ps.iter);
}

// tail of an assignment kernel: the block that takes the last ticket sees every block's atomics
// (fence before the ticket, fence after) and pushes the totals of this rank
__device__ __forceinline__ void km_push_tail(const unsigned long long *sums, int nsum, const KmPush &ps)
{
    if (ps.world == 0) return;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ps.ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ps.ticket = 0;
    km_push_sums(sums, nsum, ps);
}

// a rank whose shard is empty launches no assignment kernel: it pushes its (zero) sums with this
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_push(const unsigned long long *__restrict__ sums, int nsum,
                                                            KmPush ps, const KmState *__restrict__ state)
{
    if (state->done) return;
    km_push_sums(sums, nsum, ps);
}

// ---- generic path (K > 32): one thread per pixel, shared atomics ----------------------------
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_accumulate(
    const uint8_t *__restrict__ px, long long n, const double *__restrict__ centers, int K,
    unsigned long long *__restrict__ sums, const KmState *__restrict__ state, KmPush ps, int chunk_px)
{
    if (state && state->done) return;
    __shared__ double s_c[DP_MAX_COLORS * 3];
    __shared__ unsigned int s_sum[DP_MAX_COLORS * 4];
    __shared__ unsigned int s_ties;
    for (int i = threadIdx.x; i < K * 3; i += KM_THREADS) s_c[i] = centers[i];
    if (threadIdx.x == 0) s_ties = 0;
    const long long nchunks = (n + chunk_px - 1) / chunk_px;
    for (long long chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        for (int i = threadIdx.x; i < K * 4; i += KM_THREADS) s_sum[i] = 0;
        __syncthreads();
        const long long p0 = chunk * chunk_px;
        const int cnt = (int)((n - p0) < chunk_px ? (n - p0) : chunk_px);
        for (int j = threadIdx.x; j < cnt; j += KM_THREADS) {
            const uint8_t *q = px + (size_t)(p0 + j) * 3;
            const int r = q[0], g = q[1], b = q[2];
            const double x0 = r, x1 = g, x2 = b;
            double best = 1e300;
            int bi = 0;
            bool tie = false;
            for (int i = 0; i < K; ++i) {
                double d0 = x0 - s_c[3 * i], d1 = x1 - s_c[3 * i + 1], d2 = x2 - s_c[3 * i + 2];
                double d = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)),
                                     __dmul_rn(d2, d2));
                if (d < best) {
                    best = d;
                    bi = i;
                    tie = false;
                } else if (d == best) {
                    tie = true;
                }
            }
            if (tie) atomicAdd(&s_ties, 1u);
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
    if (threadIdx.x == 0 && s_ties) atomicAdd(&sums[4 * K], (unsigned long long)s_ties);
    km_push_tail(sums, K * 4 + 1, ps);
}

// ---- centres <- sums / count (every thread that calls it gets the same values) ---------------
// sklearn: centres = sums / counts in f64; an empty cluster keeps its centre here (sklearn moves it
// to a far sample -- see DESIGN.md; the device flag `empty` reports that it happened).
__device__ __forceinline__ double km_new_center(const unsigned long long *sums, const double *old_c, int i,
                                                int ch, bool *was_empty)
{
    const unsigned long long c = sums[4 * i + 3];
    if (!c) {
        *was_empty = true;
        return old_c[3 * i + ch];
    }
    return __ddiv_rn((double)sums[4 * i + ch], (double)c);
}

// Candidate grid + fixed-point centre table for the centres in shared memory `s_c` (K <= 32).
//   grid[4096]  box = (r>>4) | (g>>4)<<4 | (b>>4)<<8; entry: up to four surviving centres,
//               ascending, one per byte (32 = none: the pad entry); 0xffffffff = more than four
//   ent[4][33]  per candidate slot: (-4 round(1024 c) per channel, round(2048 |c|^2) with the slot number
//               in its two low bits); [.][32] = pad that never wins
// A centre is dropped from a box when another centre is strictly nearer at EVERY point of the box:
// the difference of two squared distances is linear in the point, so its maximum sits at a corner
// (exact test in double with a 1e-6 margin).  One warp per box, lane = centre.
// `replicas` copies of the grid, 4096 words apart (the persistent loop: its blocks all read the
// grid at the same moment, copies spread that over more L2 lines)
__device__ void km_build_grid(const double *s_c, int K, uint32_t *__restrict__ grid, int4 *__restrict__ ent,
                              int first_box, int box_stride, int replicas = 1)
{
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x < 4 * 33) {
        // ent[slot][k]: NEGATED fixed-point centre (multiples of 4) and H with the slot in its low bits
        const int slot = threadIdx.x / 33, k = threadIdx.x % 33;
        int4 e = make_int4(0, 0, 0, 0x3ffffffc | slot);   // [32] = pad that never wins
        if (k < K) {
            const double c0 = s_c[3 * k], c1 = s_c[3 * k + 1], c2 = s_c[3 * k + 2];
            e.x = -4 * (int)llrint(c0 * 1024.0);
            e.y = -4 * (int)llrint(c1 * 1024.0);
            e.z = -4 * (int)llrint(c2 * 1024.0);
            e.w = (((int)llrint((c0 * c0 + c1 * c1 + c2 * c2) * 2048.0)) & ~3) | slot;
        }
        ent[threadIdx.x] = e;
    }
    // f32 with a margin that covers its rounding (|terms| < 4e5, a dozen operations: error < 0.2):
    // a centre is dropped only if it loses by more than 2 everywhere -- a superset of the exact
    // candidate set, which is all the assignment pass needs (its decisions are exact).
    // K <= 16: half a warp per box (lanes 16..31 take the odd box), else a warp per box.
    const int bpw = K <= 16 ? 2 : 1;
    const int sub = bpw == 2 ? lane >> 4 : 0, li = bpw == 2 ? (lane & 15) : lane;
    float cf[3] = {0.f, 0.f, 0.f}, nf = 0.f;
    if (li < K) {
        cf[0] = (float)s_c[3 * li];
        cf[1] = (float)s_c[3 * li + 1];
        cf[2] = (float)s_c[3 * li + 2];
        nf = (float)(s_c[3 * li] * s_c[3 * li] + s_c[3 * li + 1] * s_c[3 * li + 1] + s_c[3 * li + 2] * s_c[3 * li + 2]);
    }
    for (int task = first_box + wib; task * bpw < 4096; task += box_stride) {
        const int cell = task * bpw + sub;
        const float lo[3] = {16.0f * (cell & 15), 16.0f * ((cell >> 4) & 15), 16.0f * (cell >> 8)};
        bool alive = li < K;
        for (int j = 0; j < K; ++j) {
            // centre j's values come from lane j of this half (shuffles: no shared-memory traffic)
            const int srcl = bpw == 2 ? (lane & 16) | j : j;
            const float cj0 = __shfl_sync(0xffffffffu, cf[0], srcl), cj1 = __shfl_sync(0xffffffffu, cf[1], srcl),
                        cj2 = __shfl_sync(0xffffffffu, cf[2], srcl), nj = __shfl_sync(0xffffffffu, nf, srcl);
            const float d0 = cf[0] - cj0, d1 = cf[1] - cj1, d2 = cf[2] - cj2;
            float mx = nj - nf;                                   // |c_j|^2 - |c_i|^2
            mx = fmaf(2.0f * (d0 > 0.0f ? lo[0] + 15.0f : lo[0]), d0, mx);
            mx = fmaf(2.0f * (d1 > 0.0f ? lo[1] + 15.0f : lo[1]), d1, mx);
            mx = fmaf(2.0f * (d2 > 0.0f ? lo[2] + 15.0f : lo[2]), d2, mx);
            if (j != li && mx < -2.0f) alive = false;             // centre j is nearer everywhere in the box
        }
        const unsigned mall = __ballot_sync(0xffffffffu, alive);
        unsigned m = bpw == 2 ? (mall >> (16 * sub)) & 0xffffu : mall;
        if (li == 0) {
            uint32_t e = 0xffffffffu;
            if (__popc(m) <= 4) {
                e = 0;
                for (int k = 0; k < 4; ++k) {
                    const unsigned idx = m ? (unsigned)(__ffs(m) - 1) : 32u;
                    if (m) m &= m - 1;
                    e |= idx << (8 * k);
                }
            }
            for (int r = 0; r < replicas; ++r) grid[r * 4096 + cell] = e;
        }
    }
}

// prepare: (optional) centre update from `sums_prev`, stop test, grid for the new centres
//   c_prev / c_next   f64 [K,3] ping-pong centre buffers
//   sums_prev         u64 [K*4+1] of the iteration that just finished (null: first iteration)
//   sums_next         zeroed here for the accumulate pass that follows (null: nothing follows)
//   inbox / world / wait_iter   peer exchange: the previous iteration's sums are the total of the
//                     `world` slots of parity wait_iter & 1 of this rank's inbox, valid once every
//                     flag shows wait_iter (null / 0: sums_prev holds the totals already)
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_prepare(
    const double *__restrict__ c_prev, double *__restrict__ c_next, const unsigned long long *sums_prev,
    unsigned long long *__restrict__ sums_next, int K, double tol, int last, KmState *__restrict__ state,
    uint32_t *__restrict__ grid, int4 *__restrict__ ent, const unsigned long long *inbox = nullptr, int world = 0,
    unsigned long long wait_iter = 0)
{
    if (state->done) return;
    __shared__ double s_c[DP_MAX_COLORS * 3];
    __shared__ double s_d2[DP_MAX_COLORS * 3];
    __shared__ unsigned long long s_tot[DP_MAX_COLORS * 4 + 1];
    __shared__ int s_stop, s_empty, s_err;
    if (threadIdx.x == 0) s_stop = s_empty = s_err = 0;
    __syncthreads();
    if (inbox && sums_prev) {
        const int parity = (int)(wait_iter & 1ull);
        if ((int)threadIdx.x < world) {
            const unsigned long long *f = inbox + KM_P2P_FLAGS_OFF + (size_t)parity * KM_P2P_RANKS + threadIdx.x;
            unsigned spins = 0;
            while (ld_acquire_sys(f) < wait_iter) {
                __nanosleep(200);
                if (++spins > (1u << 24)) {      // ~10 s: a peer died or never launched -> error, not a hang
                    s_err = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (s_err) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                state->error = 1;
                __threadfence();
                state->done = 1;
            }
            return;
        }
        const int nsum = K * 4 + 1;
        for (int k = threadIdx.x; k < nsum; k += KM_THREADS) {
            unsigned long long t = 0;
            for (int r = 0; r < world; ++r)
                t += __ldcg(inbox + ((size_t)parity * KM_P2P_RANKS + r) * KM_P2P_STRIDE + k);
            s_tot[k] = t;
        }
        __syncthreads();
        sums_prev = s_tot;
    }
    if (sums_prev) {
        for (int i = threadIdx.x; i < K * 3; i += KM_THREADS) {
            bool emp = false;
            const double nw = km_new_center(sums_prev, c_prev, i / 3, i % 3, &emp);
            const double d = nw - c_prev[i];
            s_c[i] = nw;
            s_d2[i] = __dmul_rn(d, d);
            if (emp) s_empty = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // single thread, fixed order: deterministic and identical in every block
            double tot = 0.0;
            for (int i = 0; i < K * 3; ++i) tot = __dadd_rn(tot, s_d2[i]);
            const int stop = (tot <= tol || last) ? 1 : 0;
            s_stop = stop;
            if (blockIdx.x == 0) {
                for (int i = 0; i < K * 3; ++i) c_next[i] = s_c[i];
                state->n_iter += 1;
                state->shift2 = tot;
                state->ties += sums_prev[4 * K];
                if (s_empty) state->empty += 1;
                __threadfence();
                if (stop) state->done = 1;   // written last; later launches on the stream see it
            }
        }
        __syncthreads();
        if (s_stop) return;
    } else {
        for (int i = threadIdx.x; i < K * 3; i += KM_THREADS) {
            s_c[i] = c_prev[i];
            if (blockIdx.x == 0 && c_next != c_prev) c_next[i] = c_prev[i];
        }
        __syncthreads();
    }
    if (sums_next && blockIdx.x == 0)
        for (int i = threadIdx.x; i < K * 4 + 1; i += KM_THREADS) sums_next[i] = 0ull;
    if (K <= 32) km_build_grid(s_c, K, grid, ent, blockIdx.x * KM_WARPS, gridDim.x * KM_WARPS);
}

// NOTE on `state->done` and other blocks: block 0 may set the flag while other blocks of the SAME
// prepare launch have not started yet; they would then return at the top without building their
// part of the grid -- harmless, because done also stops every later accumulate launch.

// grid only, for dp_kmeans_accumulate (the stand-alone assignment entry point)
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_grid_only(const double *__restrict__ centers, int K,
                                                                 uint32_t *__restrict__ grid, int4 *__restrict__ ent)
{
    __shared__ double s_c[32 * 3];
    if (threadIdx.x < K * 3) s_c[threadIdx.x] = centers[threadIdx.x];
    __syncthreads();
    km_build_grid(s_c, K, grid, ent, blockIdx.x * KM_WARPS, gridDim.x * KM_WARPS);
}

// ---- the assignment pass, K <= 32 --------------------------------------------------------------
// Fixed-point screen.  The centre table holds ci = 4 * round(c * 1024) per channel (NEGATED) and
// H = round(|c|^2 * 2048) with its two low bits replaced by the candidate's slot number, one copy
// of the table per slot.  The score
//   s = H - v . ci  =  2048 * (|v - c|^2 - |v|^2) + err,   |err| <= 3 * 255 * 2 + 4 = 1534
// orders the candidates like their distances up to 3068 score units (1.5 in squared-distance
// units), and because v . ci is a multiple of 4 its two low bits still name the slot: the minimum
// of the four scores carries its own index.  If the runner-up is more than KM_MARGIN above the
// minimum the minimum IS the exact f64 argmin; otherwise (and for boxes with more than four
// candidates) the pixel goes to the warp's list and is resolved exactly in f64 over all K centres
// (strict '<', first index -- what sklearn's argmin over exact distances gives).
//
// Sums without atomics and without data-dependent branches: every LANE owns a private set of bins
// in shared memory, bins[label][lane] = (r | g << 16, b | n << 16) -- consecutive lanes are
// consecutive 8-byte words, so the accesses are conflict-free -- and adds each of its pixels with
// one 64-bit load and store.  Undecided pixels add to a dummy row.  Every 16 tiles (256 pixels per
// lane, the most the 16-bit fields can hold) the warp folds the bins into lane k's 64-bit totals
// of centre k with warp reductions.
constexpr int KM_MARGIN = 3200;          // score units, see above
constexpr int KM_SLOW_CAP = 96;          // per-warp list of undecided pixels (per 512-pixel tile)
constexpr int KM_DRAIN_TILES = 16;

struct KmWarpShared {
    unsigned slow[KM_SLOW_CAP];      // pixel offsets inside the tile of the undecided pixels
    unsigned nslow;
    unsigned pad[3];
};

// One lane screens ALL centres for a pixel with the fixed-point scores (slot 0's copy of the
// table) and keeps the two smallest.  Returns the label, or `undecided` when the runner-up is
// within the margin.
__device__ __noinline__ unsigned km_scan_all(unsigned v, unsigned ent_a, int K, unsigned undecided)
{
    const int vr = (int)(v & 255u), vg = (int)((v >> 8) & 255u), vb = (int)((v >> 16) & 255u);
    int best = 0x7fffffff, second = 0x7fffffff;
    unsigned bi = 0;
    // four centres per round, their loads in flight together (entries K..32 are pads that never win)
    for (int k0 = 0; k0 < K; k0 += 4) {
        int sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int4 en;
            asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(en.x), "=r"(en.y), "=r"(en.z), "=r"(en.w)
                : "r"(ent_a + 16u * (k0 + u)));
            sc[u] = vb * en.z + (vg * en.y + (vr * en.x + en.w));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (sc[u] < best) {
                second = best;
                best = sc[u];
                bi = (unsigned)(k0 + u);
            } else if (sc[u] < second) {
                second = sc[u];
            }
        }
    }
    return (second - best <= KM_MARGIN) ? undecided : bi;
}

// Undecided pixels inside a box that keeps more than four candidates (grid word 0xffffffff; early
// Lloyd iterations put whole image regions there): every lane screens its own such pixels against
// all centres -- 32 lanes side by side -- and adds the decided ones to its bins; what is left goes
// to the warp's list as usual.  The pixels are read again (their first pass went to the dummy row).
// (Measured: 20 iterations over a 4K frame 1.87 -> 0.83 ms; a single first-iteration-like pass over
// 16 frames 0.41 -> 0.45 ms, the lane-serial scan being slower than the warp when such pixels are few.)
__device__ __noinline__ unsigned km_lane_overflow(unsigned slowmask, const uint8_t *lane_px, unsigned grid_a,
                                                  unsigned ent_a, unsigned bins_a, int K, unsigned undecided)
{
    unsigned m = slowmask;
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const uint8_t *q = lane_px + 3 * j;
        const unsigned v = (unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16);
        const unsigned a4 = (v >> 4) & 0x0f0f0fu;
        unsigned e;
        asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(__dp4a(a4, 0x00004004u, grid_a) + ((a4 >> 6) & 0x3c00u)));
        if (e != 0xffffffffu) continue;
        const unsigned label = km_scan_all(v, ent_a, K, undecided);
        if (label != undecided) {
            slowmask &= ~(1u << j);
            const unsigned ba = bins_a + 256u * label;
            uint2 bin;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bin.x), "=r"(bin.y) : "r"(ba));
            bin.x += __byte_perm(v, 0u, 0x4140);
            bin.y += __byte_perm(v, 1u, 0x7472);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ba), "r"(bin.x), "r"(bin.y) : "memory");
        }
    }
    return slowmask;
}

// shared memory of the assignment pass (KP = 16 or 32 centres; bins rows = KP + 1 with the dummy row)
template <int KP>
struct KmSmem {
    uint2 *bins;                 // [warps][KP+1][32]
    uint32_t *grid;              // [4096]
    int4 *ent;                   // [4][33]
    double *c;                   // [32*3]
    KmWarpShared *w;             // [warps]
    unsigned long long *tot;     // [32*4+1]
    unsigned char *end;
};

template <int KP>
__device__ __forceinline__ KmSmem<KP> km_carve(unsigned char *base)
{
    KmSmem<KP> m;
    m.bins = reinterpret_cast<uint2 *>(base);
    m.grid = reinterpret_cast<uint32_t *>(m.bins + KM_WARPS * (KP + 1) * 32);
    m.ent = reinterpret_cast<int4 *>(m.grid + 4096);
    m.c = reinterpret_cast<double *>(m.ent + 4 * 33);
    m.w = reinterpret_cast<KmWarpShared *>(m.c + 32 * 3);
    m.tot = reinterpret_cast<unsigned long long *>(m.w + KM_WARPS);
    m.end = reinterpret_cast<unsigned char *>(m.tot + 32 * 4 + 1);
    return m;
}

// grid + table into shared memory (ld.global.cg: inside the persistent loop kernel the tables
// were written by other blocks of the SAME launch)
template <int KP>
__device__ __forceinline__ void km_accum_tables(const KmSmem<KP> &m, const uint32_t *grid, const int4 *ent)
{
    const int tid = threadIdx.x;
    if (tid < 4 * 33) m.ent[tid] = __ldcg(ent + tid);
    static_assert(4096 / 4 == 4 * KM_THREADS, "four 16-byte words per thread");
    uint4 g[4];   // all four loads in flight before the first store (one L2 round trip, not four)
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = __ldcg(reinterpret_cast<const uint4 *>(grid) + tid + i * KM_THREADS);
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<uint4 *>(m.grid)[tid + i * KM_THREADS] = g[i];
}

// bins, block totals and slow lists cleared -- once per launch: the assignment pass leaves them zero
template <int KP>
__device__ __forceinline__ void km_accum_zero(const KmSmem<KP> &m)
{
    const int tid = threadIdx.x;
    for (int i = tid; i < KM_WARPS * (KP + 1) * 32; i += KM_THREADS) m.bins[i] = make_uint2(0u, 0u);
    if (tid < 32 * 4 + 1) m.tot[tid] = 0ull;
    if ((tid & 31) == 0) m.w[tid >> 5].nslow = 0u;
}

// the assignment pass proper: this block's tiles -> atomics into sums[K*4+1]
template <int KP>
__device__ __forceinline__ void km_accum_body(const KmSmem<KP> &m, const uint8_t *__restrict__ px, long long n, int K,
                                              unsigned long long *__restrict__ sums)
{
    uint2 *s_bins = m.bins;
    uint32_t *s_grid = m.grid;
    int4 *s_ent = m.ent;
    double *s_c = m.c;
    unsigned long long *s_tot = m.tot;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    KmWarpShared &ws = m.w[wib];

    const unsigned FULL = 0xffffffffu;
    const unsigned ent_a = (unsigned)__cvta_generic_to_shared(s_ent);
    const unsigned grid_a = (unsigned)__cvta_generic_to_shared(s_grid);
    // this lane's column of bins: row k at bins_a + 256 * k
    const unsigned bins_a = (unsigned)__cvta_generic_to_shared(s_bins + (size_t)wib * (KP + 1) * 32 + lane);
    unsigned long long ar = 0, ag = 0, ab = 0, an = 0;     // totals of centre `lane`
    unsigned nties = 0;

    // exact resolution of one pixel by the whole warp: lane k holds the f64 distance to centre k
    auto resolve = [&](unsigned v) {
        const int r = v & 255u, g = (v >> 8) & 255u, b = (v >> 16) & 255u;
        double d = 1e300;
        if (lane < K) {
            const double f0 = (double)r - s_c[3 * lane], f1 = (double)g - s_c[3 * lane + 1],
                         f2 = (double)b - s_c[3 * lane + 2];
            d = __dadd_rn(__dadd_rn(__dmul_rn(f0, f0), __dmul_rn(f1, f1)), __dmul_rn(f2, f2));
        }
        double best = d;
#pragma unroll
        for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(FULL, best, o));
        const unsigned eq = __ballot_sync(FULL, lane < K && d == best);
        const int label = __ffs(eq) - 1;                  // first index among the minima
        if (__popc(eq) > 1 && lane == 0) ++nties;
        if (lane == label) {
            ar += (unsigned)r;
            ag += (unsigned)g;
            ab += (unsigned)b;
            an += 1;
        }
    };
    // an undecided pixel, by the whole warp.  If its box keeps more than four candidates (grid word
    // 0xffffffff) the warp first screens ALL centres with the fixed-point scores, lane = centre
    // (two integer warp reductions); only a runner-up within the margin needs the f64 resolution.
    auto resolve_any = [&](unsigned v) {
        const unsigned a4 = (v >> 4) & 0x0f0f0fu;
        if (s_grid[(a4 & 15u) | ((a4 >> 4) & 0xf0u) | ((a4 >> 8) & 0xf00u)] == 0xffffffffu) {
            int sc = 0x7fffffff;
            if (lane < K) {
                const int4 en = s_ent[lane];   // slot 0's copy of the table
                sc = (int)(v >> 16) * en.z + ((int)((v >> 8) & 255u) * en.y + ((int)(v & 255u) * en.x + en.w));
            }
            const int m1 = __reduce_min_sync(FULL, sc);
            const int win = __ffs(__ballot_sync(FULL, sc == m1)) - 1;
            const int m2 = __reduce_min_sync(FULL, lane == win ? 0x7fffffff : sc);
            if (m2 - m1 > KM_MARGIN) {
                if (lane == win) {
                    ar += v & 255u;
                    ag += (v >> 8) & 255u;
                    ab += (v >> 16) & 255u;
                    an += 1;
                }
                return;
            }
        }
        resolve(v);
    };
    // fold the lanes' private bins into the 64-bit totals (lane k owns centre k)
    auto drain = [&]() {
        for (int k = 0; k < K; ++k) {
            uint2 bin;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bin.x), "=r"(bin.y) : "r"(bins_a + 256u * k));
            asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(bins_a + 256u * k), "r"(0u) : "memory");
            const unsigned sr = __reduce_add_sync(FULL, bin.x & 0xffffu), sg = __reduce_add_sync(FULL, bin.x >> 16);
            const unsigned sb = __reduce_add_sync(FULL, bin.y & 0xffffu), sn = __reduce_add_sync(FULL, bin.y >> 16);
            if (lane == k) {
                ar += sr;
                ag += sg;
                ab += sb;
                an += sn;
            }
        }
        asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(bins_a + 256u * KP), "r"(0u) : "memory");   // dummy row
    };

    // head: pixels in front of the first 16-byte boundary; body: whole 16-pixel groups; tail
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(px) & 15u);
    long long head = (long long)((16u - mis) * 11u & 15u);    // 3 * head = -mis (mod 16)
    if (mis == 0) head = 0;
    if (head > n) head = n;
    const long long ngroups = (n - head) >> 4;
    const long long tail0 = head + (ngroups << 4);
    if (blockIdx.x == 0 && wib == 0) {
        for (long long i = 0; i < head; ++i) {
            const uint8_t *q = px + 3 * i;
            resolve((unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16));
        }
        for (long long i = tail0; i < n; ++i) {
            const uint8_t *q = px + 3 * i;
            resolve((unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16));
        }
    }

    const uint4 *g4 = reinterpret_cast<const uint4 *>(px + 3 * head);
    const long long ntiles = (ngroups + 31) >> 5;             // 32 groups = 512 pixels per warp tile
    // consecutive tiles go to different BLOCKS (then to the warps of a block): a short input
    // spreads over all SMs instead of filling the first blocks' eight warps
    const long long wstride = (long long)gridDim.x * KM_WARPS;
    long long t = (long long)blockIdx.x + (long long)gridDim.x * wib;
    uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = nx0, nx2 = nx0;
    auto fetch = [&](long long tile) {
        const long long grp = tile * 32 + lane;
        if (tile < ntiles && grp < ngroups) {
            nx0 = __ldg(g4 + 3 * grp);
            nx1 = __ldg(g4 + 3 * grp + 1);
            nx2 = __ldg(g4 + 3 * grp + 2);
        }
    };
    fetch(t);
    int since_drain = 0;
    for (; t < ntiles; t += wstride) {
        unsigned w[12] = {nx0.x, nx0.y, nx0.z, nx0.w, nx1.x, nx1.y, nx1.z, nx1.w, nx2.x, nx2.y, nx2.z, nx2.w};
        const bool live = t * 32 + lane < ngroups;
        fetch(t + wstride);                                   // next tile's loads fly during this one
        unsigned slowmask = live ? 0u : 0xffffu;              // dead lanes: every pixel to the dummy row
#pragma unroll
        for (int j = 0; j < 16; ++j) {                        // pixel j of the lane's 16
            const int wi = (3 * j) >> 2, sh = ((3 * j) & 3) * 8;
            const unsigned v = sh == 0 ? w[wi] : __funnelshift_r(w[wi], w[wi + (wi < 11 ? 1 : 0)], sh);
            // box = (r>>4) | (g>>4)<<4 | (b>>4)<<8, as a byte offset into the u32 grid
            const unsigned a4 = (v >> 4) & 0x0f0f0fu;
            unsigned e;
            asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(__dp4a(a4, 0x00004004u, grid_a) + ((a4 >> 6) & 0x3c00u)));
            const int vr = (int)__byte_perm(v, 0u, 0x4440), vg = (int)__byte_perm(v, 0u, 0x4441),
                      vb = (int)__byte_perm(v, 0u, 0x4442);
            int s[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                int4 en;   // slot c's copy of the table: ent_a + 528 * c + 16 * byte c of e
                const unsigned a = __dp4a(e, 0x10u << (8 * c), ent_a + 528u * c);
                asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(en.x), "=r"(en.y), "=r"(en.z), "=r"(en.w) : "r"(a));
                s[c] = vb * en.z + (vg * en.y + (vr * en.x + en.w));
            }
            const int lo01 = min(s[0], s[1]), hi01 = max(s[0], s[1]);
            const int lo23 = min(s[2], s[3]), hi23 = max(s[2], s[3]);
            const int m1 = min(lo01, lo23);
            const int m2 = min(max(lo01, lo23), min(hi01, hi23));
            const bool slow = (e == 0xffffffffu) || (m2 - m1 <= KM_MARGIN) || !live;
            const unsigned label = slow ? (unsigned)KP : __byte_perm(e, 0u, ((unsigned)m1 & 3u) | 0x4440u);
            if (slow) slowmask |= 1u << j;
            // bins[label][lane] += (r | g << 16, b | 1 << 16)
            const unsigned ba = bins_a + 256u * label;
            uint2 bin;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(bin.x), "=r"(bin.y) : "r"(ba));
            bin.x += __byte_perm(v, 0u, 0x4140);
            bin.y += __byte_perm(v, 1u, 0x7472);
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(ba), "r"(bin.x), "r"(bin.y) : "memory");
        }
        // undecided pixels of boxes with more than four candidates: the lanes first screen their own
        if (live && slowmask)
            slowmask = km_lane_overflow(slowmask, px + 3 * (head + t * 512 + lane * 16), grid_a, ent_a, bins_a, K,
                                        (unsigned)KP);
        // undecided pixels of the tile: onto the warp's list, then the warp resolves them one by one
        if (live && slowmask) {
            unsigned m = slowmask;
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                const unsigned pos = atomicAdd(&ws.nslow, 1u);
                if (pos < KM_SLOW_CAP) ws.slow[pos] = (unsigned)(lane * 16 + j);
            }
        }
        __syncwarp();
        const unsigned ns = ws.nslow;
        if (ns) {
            const uint8_t *tile_px = px + 3 * (head + t * 512);
            if (ns <= KM_SLOW_CAP) {
                for (unsigned i = 0; i < ns; ++i) {
                    const uint8_t *q = tile_px + 3 * ws.slow[i];
                    resolve_any((unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16));
                }
            } else {
                // list overflow (pathological inputs): redo the tile's undecided pixels by
                // re-screening every pixel of the tile, cooperatively
                const long long base = head + t * 512;
                const long long lim = min((long long)512, tail0 - base);
                for (long long i = 0; i < lim; ++i) {
                    const uint8_t *q = px + 3 * (base + i);
                    const unsigned v = (unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16);
                    const unsigned a4 = (v >> 4) & 0x0f0f0fu;
                    const unsigned e = s_grid[(a4 & 15u) | ((a4 >> 4) & 0xf0u) | ((a4 >> 8) & 0xf00u)];
                    int sc[4];
                    for (int c = 0; c < 4; ++c) {
                        const int4 en = s_ent[33 * c + ((e >> (8 * c)) & 255u) % 33u];
                        sc[c] = (int)(v >> 16) * en.z + ((int)((v >> 8) & 255u) * en.y + ((int)(v & 255u) * en.x + en.w));
                    }
                    const int lo01 = min(sc[0], sc[1]), hi01 = max(sc[0], sc[1]);
                    const int lo23 = min(sc[2], sc[3]), hi23 = max(sc[2], sc[3]);
                    const int m1 = min(lo01, lo23), m2 = min(max(lo01, lo23), min(hi01, hi23));
                    if ((e == 0xffffffffu) || (m2 - m1 <= KM_MARGIN)) resolve_any(v);
                }
            }
            __syncwarp();
            if (lane == 0) ws.nslow = 0u;
            __syncwarp();
        }
        if (++since_drain == KM_DRAIN_TILES) {
            drain();
            since_drain = 0;
        }
    }
    drain();
    // block totals first (the global accumulators are a few addresses shared by every block: one
    // flush per block keeps the L2 atomic unit out of the critical path)
    if (lane < K && an) {
        atomicAdd(&s_tot[4 * lane], ar);
        atomicAdd(&s_tot[4 * lane + 1], ag);
        atomicAdd(&s_tot[4 * lane + 2], ab);
        atomicAdd(&s_tot[4 * lane + 3], an);
    }
    if (lane == 0 && nties) atomicAdd(&s_tot[32 * 4], (unsigned long long)nties);
    __syncthreads();
    // (every word that was used goes back to zero: the persistent loop comes here again)
    if (tid < K * 4) {
        const unsigned long long v = s_tot[tid];
        if (v) {
            atomicAdd(&sums[tid], v);
            s_tot[tid] = 0ull;
        }
    }
    if (tid == 0) {
        const unsigned long long v = s_tot[32 * 4];
        if (v) {
            atomicAdd(&sums[4 * K], v);
            s_tot[32 * 4] = 0ull;
        }
    }
}

template <int KP>
__global__ void __launch_bounds__(KM_THREADS) k_kmeans_accum16(
    const uint8_t *__restrict__ px, long long n, const double *__restrict__ centers, int K,
    unsigned long long *__restrict__ sums, const uint32_t *__restrict__ grid, const int4 *__restrict__ ent,
    const KmState *__restrict__ state, KmPush ps)
{
    if (state && state->done) return;
    extern __shared__ __align__(16) unsigned char km_smem[];
    const KmSmem<KP> m = km_carve<KP>(km_smem);
    if (threadIdx.x < K * 3) m.c[threadIdx.x] = centers[threadIdx.x];
    km_accum_tables<KP>(m, grid, ent);
    km_accum_zero<KP>(m);
    __syncthreads();
    km_accum_body<KP>(m, px, n, K, sums);
    km_push_tail(sums, K * 4 + 1, ps);
}

// ---- the whole Lloyd loop as ONE persistent launch (K <= 32) -------------------------------------
// All blocks are resident (grid = SMs x occupancy) and meet at a grid barrier twice per iteration:
//   prepare(it)   every block: totals of iteration it-1 (own sums, or the ranks' inbox slots once
//                 their flags show the tag), new centres, shift, stop test -- all blocks compute
//                 the same values, so the decision to leave the loop needs no broadcast; block 0
//                 records the state and clears the sums; the 4096 grid boxes are spread over the
//                 warps of the whole grid
//   -- barrier --
//   assign(it)    as k_kmeans_accum16
//   -- barrier --  then block 0 pushes the rank's sums to the peers (peer exchange)
// Two launches and their gaps per iteration (~22 us on this stack) become two barriers (~3 us).
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// barrier over all blocks of the launch: bar[0] counts arrivals and is never reset -- barrier
// number g (0, 1, ...) is complete when the count reaches (g + 1) * blocks, and the block that
// makes it so publishes g + 1 in bar[1].  One release (the arriving atomic) and one acquire (the
// poll) per block.  Returns false after ~10 s without release (a block never became resident):
// callers bail out.
__device__ __forceinline__ bool km_grid_barrier(unsigned *bar, unsigned &gen)
{
    __shared__ int s_ok;
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ok = 1;
        unsigned old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(bar) : "memory");
        if (old + 1u == (gen + 1u) * gridDim.x) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1u) : "memory");
        } else {
            unsigned spins = 0;
            while (ld_acquire_gpu_u32(bar + 1) == gen) {
                if (++spins > 64u) __nanosleep(32);
                if (spins > (1u << 26)) {
                    s_ok = 0;
                    break;
                }
            }
        }
    }
    ++gen;
    __syncthreads();
    return s_ok != 0;
}

#ifdef DP_KM_TIMING
__device__ unsigned long long g_km_timing[8];
#define KM_TICK(k)                                                             \
    do {                                                                       \
        const long long _now = clock64();                                      \
        if (blockIdx.x == 1 && threadIdx.x == 0) g_km_timing[(k)] += (unsigned long long)(_now - _tick); \
        _tick = _now;                                                          \
    } while (0)
#else
#define KM_TICK(k) do { } while (0)
#endif

template <int KP>
__global__ void __launch_bounds__(KM_THREADS, 3) k_kmeans_loop(
    const uint8_t *__restrict__ px, long long n, double *__restrict__ c_io, int K, double tol, int max_iter,
    unsigned long long *__restrict__ sums2, KmState *__restrict__ state, uint32_t *__restrict__ grid,
    int4 *__restrict__ ent, unsigned *__restrict__ bar, KmPush ps, const unsigned long long *inbox)
{
    extern __shared__ __align__(16) unsigned char km_smem[];
    const KmSmem<KP> m = km_carve<KP>(km_smem);
    double *s_new = reinterpret_cast<double *>(m.end + 8 - ((uintptr_t)m.end & 7));   // [32*3]
    double *s_d2 = s_new + 32 * 3;                                                    // [32*3]
    unsigned long long *s_prev = reinterpret_cast<unsigned long long *>(s_d2 + 32 * 3);   // [32*4+1]
    __shared__ int s_stop, s_empty, s_err;
    const int tid = threadIdx.x;
    const int nsum = K * 4 + 1;
    if (tid < K * 3) m.c[tid] = c_io[tid];
    if (tid == 0) s_stop = s_empty = s_err = 0;
    km_accum_zero<KP>(m);
    __syncthreads();
    // this block's share of the assignment pass (same tiling as k_kmeans_accum16)
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(px) & 15u);
    long long head = mis ? (long long)((16u - mis) * 11u & 15u) : 0;
    if (head > n) head = n;
    const long long ntiles = (((n - head) >> 4) + 31) >> 5;
    const bool has_work = blockIdx.x == 0 || (long long)blockIdx.x < ntiles;
    const unsigned long long tag0 = ps.iter;
    bool failed = false;
    unsigned bar_gen = 0;
    unsigned long long ties_acc = 0;
    int empty_acc = 0;
#ifdef DP_KM_TIMING
    long long _tick = clock64();
#endif
    for (int it = 1;; ++it) {
        KM_TICK(6);
        unsigned long long *sums_cur = sums2 + (size_t)(it & 1) * nsum;
        if (it > 1) {
            // ---- totals of iteration it-1
            if (inbox) {
                const int parity = (it - 1) & 1;
                if (tid < ps.world) {
                    const unsigned long long *f = inbox + KM_P2P_FLAGS_OFF + (size_t)parity * KM_P2P_RANKS + tid;
                    const unsigned long long want = tag0 + (unsigned long long)(it - 1);
                    unsigned spins = 0;
                    while (ld_acquire_sys(f) < want) {
                        __nanosleep(100);
                        if (++spins > (1u << 25)) {
                            s_err = 1;
                            break;
                        }
                    }
                }
                __syncthreads();
                if (s_err) {
                    failed = true;
                    break;
                }
                for (int k = tid; k < nsum; k += KM_THREADS) {
                    unsigned long long t = 0;
                    for (int r = 0; r < ps.world; ++r)
                        t += __ldcg(inbox + ((size_t)parity * KM_P2P_RANKS + r) * KM_P2P_STRIDE + k);
                    s_prev[k] = t;
                }
            } else {
                const unsigned long long *sp = sums2 + (size_t)((it - 1) & 1) * nsum;
                for (int k = tid; k < nsum; k += KM_THREADS) s_prev[k] = __ldcg(sp + k);
            }
            __syncthreads();
            // ---- centres, shift, stop test (same code and order as k_kmeans_prepare)
            if (tid < K * 3) {
                bool emp = false;
                const double nw = km_new_center(s_prev, m.c, tid / 3, tid % 3, &emp);
                const double d = nw - m.c[tid];
                s_new[tid] = nw;
                s_d2[tid] = __dmul_rn(d, d);
                if (emp) s_empty = 1;
            }
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                for (int i = 0; i < K * 3; ++i) tot = __dadd_rn(tot, s_d2[i]);
                const int stop = (tot <= tol || it - 1 >= max_iter) ? 1 : 0;
                s_stop = stop;
                if (blockIdx.x == 0) {
                    // stores only (the counters live in registers): no global round trip on the
                    // path every other block waits for at the barrier
                    for (int i = 0; i < K * 3; ++i) c_io[i] = s_new[i];
                    ties_acc += s_prev[4 * K];
                    empty_acc += s_empty ? 1 : 0;
                    state->n_iter = it - 1;
                    state->shift2 = tot;
                    state->ties = ties_acc;
                    state->empty = empty_acc;
                }
                s_empty = 0;
            }
            __syncthreads();
            if (tid < K * 3) m.c[tid] = s_new[tid];
            __syncthreads();
            if (s_stop) break;
        }
        KM_TICK(0);
        // ---- tables for the new centres; clear the sums of this iteration
        if (blockIdx.x == 0)
            for (int k = tid; k < nsum; k += KM_THREADS) sums_cur[k] = 0ull;
        km_build_grid(m.c, K, grid, ent, blockIdx.x * KM_WARPS, gridDim.x * KM_WARPS, KM_GRID_REPLICAS);
        KM_TICK(1);
        if (!km_grid_barrier(bar, bar_gen)) {
            failed = true;
            break;
        }
        KM_TICK(2);
        // ---- assignment pass
        if (has_work) {
            km_accum_tables<KP>(m, grid + (blockIdx.x % KM_GRID_REPLICAS) * 4096, ent);
            __syncthreads();
            KM_TICK(3);
            km_accum_body<KP>(m, px, n, K, sums_cur);
        }
        KM_TICK(4);
        if (!km_grid_barrier(bar, bar_gen)) {
            failed = true;
            break;
        }
        KM_TICK(5);
        if (inbox && blockIdx.x == 0) {
            KmPush q = ps;
            q.parity = it & 1;
            q.iter = tag0 + (unsigned long long)it;
            km_push_sums(sums_cur, nsum, q);
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (failed) state->error = 1;
        __threadfence();
        state->done = 1;
    }
}

inline size_t km_loop_smem(int KP);

inline size_t km_accum_smem(int KP)
{
    return (size_t)KM_WARPS * (KP + 1) * 32 * 8 + 4096 * 4 + 4 * 33 * 16 + 32 * 3 * 8 + KM_WARPS * sizeof(KmWarpShared) +
           (32 * 4 + 1) * 8;
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

// ---- per-device scratch of the stand-alone accumulate entry point (grid + centre table) -----
struct KmScratch {
    uint32_t *grid = nullptr;
    int4 *ent = nullptr;
};
std::mutex g_km_mu;
KmScratch g_km_scratch[64];

int km_scratch(KmScratch **out)
{
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    DP_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    std::lock_guard<std::mutex> lk(g_km_mu);
    KmScratch &s = g_km_scratch[dev];
    if (!s.grid) {
        DP_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.grid), 4096 * sizeof(uint32_t)));
        DP_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.ent), 4 * 33 * sizeof(int4)));
    }
    *out = &s;
    return 0;
}

inline size_t km_loop_smem(int KP) { return km_accum_smem(KP) + 8 + 2 * 32 * 3 * 8 + (32 * 4 + 1) * 8; }

// launch of the persistent Lloyd loop: every block must be resident (they meet at grid barriers),
// so the grid is the occupancy of an idle device -- DP_KMEANS_LOOP_BLOCKS lowers it (tests that run
// several ranks as streams of ONE device need all their kernels resident side by side)
template <int KP>
int km_launch_loop_kp(const uint8_t *pixels, long long n, double *c_io, int K, double tol, int max_iter,
                      unsigned long long *sums2, KmState *state, uint32_t *grid, int4 *ent, unsigned *bar,
                      const KmPush &ps, const unsigned long long *inbox, cudaStream_t st)
{
    static thread_local int per_sm[64] = {0};
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    const size_t smem = km_loop_smem(KP);
    if (dev >= 0 && dev < 64 && !per_sm[dev]) {
        DP_CUDA(cudaFuncSetAttribute(k_kmeans_loop<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_kmeans_loop<KP>, KM_THREADS, smem));
        per_sm[dev] = occ < 1 ? 1 : occ;
    }
    long long blocks = (long long)dp_num_sms() * (dev >= 0 && dev < 64 ? per_sm[dev] : 1);
    if (const char *ev = getenv("DP_KMEANS_LOOP_BLOCKS")) {
        const long long b = atoll(ev);
        if (b >= 1 && b < blocks) blocks = b;
    }
    k_kmeans_loop<KP><<<(int)blocks, KM_THREADS, smem, st>>>(pixels, n, c_io, K, tol, max_iter, sums2, state, grid, ent,
                                                             bar, ps, inbox);
    DP_LAUNCH_CHECK();
    return 0;
}

// launch of the assignment pass: bins for 16 or 32 centres, as many resident blocks as fit
template <int KP>
int km_launch_accum_kp(const uint8_t *pixels, long long n, const double *centers, int K, unsigned long long *sums,
                       const uint32_t *grid, const int4 *ent, const KmState *state, cudaStream_t st, const KmPush &ps)
{
    static thread_local int per_sm[64] = {0};
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    const size_t smem = km_accum_smem(KP);
    if (dev >= 0 && dev < 64 && !per_sm[dev]) {
        DP_CUDA(cudaFuncSetAttribute(k_kmeans_accum16<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ps = 0;
        DP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_kmeans_accum16<KP>, KM_THREADS, smem));
        per_sm[dev] = ps < 1 ? 1 : ps;
    }
    const long long cap = (long long)dp_num_sms() * (dev >= 0 && dev < 64 ? per_sm[dev] : 1);
    const long long tiles = (n / 16 + 31) / 32;
    long long blocks = (tiles + KM_WARPS - 1) / KM_WARPS;
    if (blocks < 1) blocks = 1;
    k_kmeans_accum16<KP><<<(int)(blocks < cap ? blocks : cap), KM_THREADS, smem, st>>>(pixels, n, centers, K, sums, grid,
                                                                                       ent, state, ps);
    DP_LAUNCH_CHECK();
    return 0;
}

// pixels per block of the generic pass: at most KM_PIX_PER_BLOCK (exact u32 partials), less when
// the input is short so that it still spreads over the whole device (the reference's k-means
// sees 10 000 samples: one 16 384-pixel chunk would be ONE block)
inline int km_generic_chunk(long long n, long long cap, int *grid)
{
    long long chunk = (n + cap - 1) / cap;
    chunk = (chunk + KM_THREADS - 1) / KM_THREADS * KM_THREADS;
    if (chunk < KM_THREADS) chunk = KM_THREADS;
    if (chunk > KM_PIX_PER_BLOCK) chunk = KM_PIX_PER_BLOCK;
    const long long chunks = (n + chunk - 1) / chunk;
    *grid = (int)(chunks < 1 ? 1 : (chunks < cap ? chunks : cap));
    return (int)chunk;
}

KmPush km_no_push()
{
    KmPush ps;
    memset(&ps, 0, sizeof(ps));
    return ps;
}

int km_launch_accum(const uint8_t *pixels, long long n, const double *centers, int K, unsigned long long *sums,
                    const uint32_t *grid, const int4 *ent, const KmState *state, cudaStream_t st,
                    const KmPush &ps = km_no_push())
{
    return K <= 16 ? km_launch_accum_kp<16>(pixels, n, centers, K, sums, grid, ent, state, st, ps)
                   : km_launch_accum_kp<32>(pixels, n, centers, K, sums, grid, ent, state, st, ps);
}

// ---- NCCL, resolved at run time (the library has no link-time dependency on it) -----------------
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(void *id) = nullptr;
    int (*CommInitRank)(void **comm, int nranks, DpNcclId id, int rank) = nullptr;
    int (*CommDestroy)(void *comm) = nullptr;
    int (*AllReduce)(const void *send, void *recv, size_t count, int dtype, int op, void *comm,
                     cudaStream_t stream) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load(const char *path)
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return 0;
    const char *cands[] = {path, getenv("DP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *c : cands) {
        if (!c || !*c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    DP_REQUIRE(h, "libnccl.so.2 not found (set DP_NCCL_LIB or pass the path to dp_nccl_load)");
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void *)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void **, int, DpNcclId, int)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void *)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t)>(
        dlsym(h, "ncclAllReduce"));
    g_nccl.GetErrorString = reinterpret_cast<const char *(*)(int)>(dlsym(h, "ncclGetErrorString"));
    DP_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce,
               "libnccl does not export the expected symbols");
    g_nccl.handle = h;
    return 0;
}

#define DP_NCCL(call)                                                                             \
    do {                                                                                          \
        const int _r = (call);                                                                    \
        if (_r != 0) {                                                                            \
            dp_set_error("%s:%d %s -> nccl error %d (%s)", __FILE__, __LINE__, #call, _r,          \
                         g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");                 \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

constexpr int DP_NCCL_UINT64 = 5;   // ncclUint64
constexpr int DP_NCCL_SUM = 0;      // ncclSum

}  // namespace

extern "C" int dp_nccl_load(const char *path) { return nccl_load(path); }

extern "C" int dp_nccl_unique_id(void *id128)
{
    DP_REQUIRE(id128, "null argument");
    if (nccl_load(nullptr)) return 1;
    DP_NCCL(g_nccl.GetUniqueId(id128));
    return 0;
}

extern "C" int dp_nccl_comm_create(const void *id128, int rank, int world, void **comm)
{
    DP_REQUIRE(id128 && comm && world >= 1 && rank >= 0 && rank < world, "bad argument");
    if (nccl_load(nullptr)) return 1;
    DpNcclId id;
    memcpy(&id, id128, sizeof(id));
    DP_NCCL(g_nccl.CommInitRank(comm, world, id, rank));
    return 0;
}

extern "C" int dp_nccl_comm_destroy(void *comm)
{
    if (!comm) return 0;
    DP_REQUIRE(g_nccl.handle, "NCCL was never loaded");
    DP_NCCL(g_nccl.CommDestroy(comm));
    return 0;
}

extern "C" int dp_nccl_allreduce_u64(void *buf, size_t count, void *comm, void *stream)
{
    DP_REQUIRE(buf && comm, "null argument");
    DP_REQUIRE(g_nccl.handle, "NCCL was never loaded");
    DP_NCCL(g_nccl.AllReduce(buf, buf, count, DP_NCCL_UINT64, DP_NCCL_SUM, comm, dp_stream(stream)));
    return 0;
}

extern "C" int dp_kmeans_accumulate(const uint8_t *pixels, int64_t n, const double *centers,
                                    int K, unsigned long long *sums, void *stream)
{
    DP_RANGE("dp_kmeans_accumulate");
    DP_REQUIRE(pixels && centers && sums, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS && n >= 0, "bad size");
    if (n == 0) return 0;
    cudaStream_t st = dp_stream(stream);
    if (K <= 32) {
        KmScratch *sc = nullptr;
        if (km_scratch(&sc)) return 1;
        k_kmeans_grid_only<<<512, KM_THREADS, 0, st>>>(centers, K, sc->grid, sc->ent);
        DP_LAUNCH_CHECK();
        return km_launch_accum(pixels, n, centers, K, sums, sc->grid, sc->ent, nullptr, st);
    }
    int grid = 1;
    const int chunk_px = km_generic_chunk(n, (long long)dp_num_sms() * 8, &grid);
    k_kmeans_accumulate<<<grid, KM_THREADS, 0, st>>>(pixels, n, centers, K, sums, nullptr, km_no_push(), chunk_px);
    DP_LAUNCH_CHECK();
    return 0;
}

extern "C" int dp_kmeans_update(const unsigned long long *sums, int K, double *centers,
                                double *shift2, void *stream)
{
    DP_RANGE("dp_kmeans_update");
    DP_REQUIRE(sums && centers && shift2, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS, "bad size");
    k_kmeans_update<<<1, 32, 0, dp_stream(stream)>>>(sums, K, centers, shift2);
    DP_LAUNCH_CHECK();
    return 0;
}

// The whole Lloyd loop (see the header).  Synchronous: returns when the loop has stopped.
// Exchange of the sums between ranks: none, ncclAllReduce on the stream, or the peer-memory push.
namespace {

// small pool of pinned KmState blocks (one per concurrent Lloyd loop)
std::mutex g_pin_mu;
std::vector<KmState *> g_pin_free;

KmState *km_pinned_state_get()
{
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) {
            KmState *p = g_pin_free.back();
            g_pin_free.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, 64, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return static_cast<KmState *>(p);
}

void km_pinned_state_put(KmState *p)
{
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(p);
}

struct KmExchange {
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;            // peer exchange when inboxes != nullptr
    void *const *inboxes = nullptr;
    unsigned long long epoch = 0;       // tags = epoch << 32 | iteration (flags only ever grow)
};

int km_lloyd(const uint8_t *pixels, int64_t n, double *centers_host, int K, double tol, int max_iter,
             const KmExchange &ex, int check_every, int *n_iter, double *shift2, unsigned long long *ties,
             int *empty_iters, cudaStream_t st)
{
    int dev = 0;
    DP_CUDA(cudaGetDevice(&dev));
    if (dp_retain_pool(dev)) return 1;
    if (check_every < 1) check_every = 4;
    const bool p2p = ex.inboxes != nullptr && ex.world > 1;
    KmPeers peers;
    memset(&peers, 0, sizeof(peers));
    if (p2p)
        for (int r = 0; r < ex.world; ++r) peers.inbox[r] = static_cast<unsigned long long *>(ex.inboxes[r]);
    const unsigned long long *my_inbox = p2p ? peers.inbox[ex.rank] : nullptr;
    const size_t nsum = (size_t)K * 4 + 1;
    const size_t off_c = 0, off_s = off_c + 2 * (size_t)K * 3 * 8, off_state = off_s + 2 * nsum * 8,
                 off_grid = off_state + 64, off_ent = off_grid + 4096 * 4 * KM_GRID_REPLICAS, total = off_ent + 4 * 33 * 16;
    char *ws = nullptr;
    DP_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ws), total, st));
    double *cbuf[2] = {reinterpret_cast<double *>(ws + off_c), reinterpret_cast<double *>(ws + off_c) + (size_t)K * 3};
    unsigned long long *sums[2] = {reinterpret_cast<unsigned long long *>(ws + off_s),
                                   reinterpret_cast<unsigned long long *>(ws + off_s) + nsum};
    KmState *state = reinterpret_cast<KmState *>(ws + off_state);
    unsigned *ticket = reinterpret_cast<unsigned *>(ws + off_state + 48);
    static_assert(sizeof(KmState) <= 48, "the ticket lives behind the state in its 64-byte slot");
    uint32_t *grid = reinterpret_cast<uint32_t *>(ws + off_grid);
    int4 *ent = reinterpret_cast<int4 *>(ws + off_ent);
    // the state is polled through PINNED memory: a copy to pageable memory blocks inside the driver,
    // and with several ranks as threads of one process that stalls the other ranks' launches
    KmState *pinned = km_pinned_state_get();
    if (!pinned) {
        cudaFreeAsync(ws, st);
        dp_set_error("cudaHostAlloc failed for the k-means state");
        return 1;
    }
    KmState &host_state = *pinned;
    memset(&host_state, 0, sizeof(host_state));
    auto finish = [&](int code) {
        cudaFreeAsync(ws, st);
        cudaStreamSynchronize(st);
        km_pinned_state_put(pinned);
        return code;
    };
#define KM_TRY(call)                                                                          \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            dp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return finish(1);                                                                 \
        }                                                                                     \
    } while (0)
    KM_TRY(cudaMemsetAsync(ws + off_s, 0, 2 * nsum * 8 + 64, st));
    KM_TRY(cudaMemcpyAsync(cbuf[0], centers_host, (size_t)K * 3 * 8, cudaMemcpyHostToDevice, st));
    int ggrid = 1;
    const int gchunk = km_generic_chunk(n, (long long)dp_num_sms() * 8, &ggrid);
    const unsigned long long tag0 = ex.epoch << 32;
    int it = 1;
    bool stopped = false;
    // K <= 32 without an NCCL communicator: the whole loop is one persistent launch
    bool persistent = K <= 32 && !ex.nccl_comm;
    if (const char *ev = getenv("DP_KMEANS_LOOP")) persistent = persistent && atoi(ev) != 0;
    if (persistent) {
        KmPush ps = km_no_push();
        if (p2p) {
            ps.peers = peers;
            ps.rank = ex.rank;
            ps.world = ex.world;
        }
        ps.iter = tag0;
        const int rc = K <= 16 ? km_launch_loop_kp<16>(pixels, n, cbuf[0], K, tol, max_iter, sums[0], state, grid, ent,
                                                        ticket, ps, my_inbox, st)
                               : km_launch_loop_kp<32>(pixels, n, cbuf[0], K, tol, max_iter, sums[0], state, grid, ent,
                                                        ticket, ps, my_inbox, st);
        if (rc) return finish(1);
        KM_TRY(cudaMemcpyAsync(&host_state, state, sizeof(KmState), cudaMemcpyDeviceToHost, st));
        KM_TRY(cudaStreamSynchronize(st));
        if (host_state.error) {
            dp_set_error("k-means loop kernel gave up at a barrier or waiting for a peer's sums");
            return finish(3);
        }
        KM_TRY(cudaMemcpyAsync(centers_host, cbuf[0], (size_t)K * 3 * 8, cudaMemcpyDeviceToHost, st));
        KM_TRY(cudaStreamSynchronize(st));
        if (n_iter) *n_iter = host_state.n_iter;
        if (shift2) *shift2 = host_state.shift2;
        if (ties) *ties = host_state.ties;
        if (empty_iters) *empty_iters = host_state.empty;
        return finish(0);
    }
    while (!stopped) {
        const int upto = it + check_every - 1 < max_iter ? it + check_every - 1 : max_iter;
        for (; it <= upto; ++it) {
            // prepare(it): centres C_{it-1} from the sums of iteration it-1 (none for it == 1), stop
            // test, grid; then the assignment pass of iteration `it` into sums[it & 1]
            k_kmeans_prepare<<<K <= 32 ? 512 : 1, KM_THREADS, 0, st>>>(
                it == 1 ? cbuf[0] : cbuf[it & 1], cbuf[(it - 1) & 1], it > 1 ? sums[(it - 1) & 1] : nullptr,
                sums[it & 1], K, tol, 0, state, grid, ent, my_inbox, ex.world, tag0 + (unsigned long long)(it - 1));
            KmPush ps = km_no_push();
            if (p2p) {
                ps.peers = peers;
                ps.rank = ex.rank;
                ps.world = ex.world;
                ps.parity = it & 1;
                ps.iter = tag0 + (unsigned long long)it;
                ps.ticket = ticket;
            }
            if (n > 0) {
                // with peers the last block of the assignment kernel pushes the sums itself
                if (K <= 32) {
                    if (km_launch_accum(pixels, n, cbuf[(it - 1) & 1], K, sums[it & 1], grid, ent, state, st, ps))
                        return finish(1);
                } else
                    k_kmeans_accumulate<<<ggrid, KM_THREADS, 0, st>>>(pixels, n, cbuf[(it - 1) & 1], K, sums[it & 1],
                                                                      state, ps, gchunk);
            } else if (p2p)
                k_kmeans_push<<<1, KM_THREADS, 0, st>>>(sums[it & 1], (int)nsum, ps, state);
            KM_TRY(cudaGetLastError());
            if (!p2p && ex.nccl_comm) {
                const int r = g_nccl.AllReduce(sums[it & 1], sums[it & 1], nsum, DP_NCCL_UINT64, DP_NCCL_SUM, ex.nccl_comm, st);
                if (r != 0) {
                    dp_set_error("ncclAllReduce failed: %d (%s)", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
                    return finish(1);
                }
            }
        }
        if (it > max_iter) {
            // the iteration budget is used up: finalise C_{max_iter}
            k_kmeans_prepare<<<1, KM_THREADS, 0, st>>>(cbuf[it & 1], cbuf[(it - 1) & 1], sums[(it - 1) & 1], nullptr, K, tol,
                                                      1, state, grid, ent, my_inbox, ex.world,
                                                      tag0 + (unsigned long long)(it - 1));
            KM_TRY(cudaGetLastError());
        }
        KM_TRY(cudaMemcpyAsync(&host_state, state, sizeof(KmState), cudaMemcpyDeviceToHost, st));
        KM_TRY(cudaStreamSynchronize(st));
        stopped = host_state.done != 0 || it > max_iter;
    }
    if (host_state.error) {
        dp_set_error("k-means peer exchange timed out: a rank did not deliver its sums");
        return finish(3);
    }
    // final centres: C_{n_iter} lives in cbuf[n_iter & 1]
    KM_TRY(cudaMemcpyAsync(centers_host, cbuf[host_state.n_iter & 1], (size_t)K * 3 * 8, cudaMemcpyDeviceToHost, st));
    KM_TRY(cudaStreamSynchronize(st));
#undef KM_TRY
    if (n_iter) *n_iter = host_state.n_iter;
    if (shift2) *shift2 = host_state.shift2;
    if (ties) *ties = host_state.ties;
    if (empty_iters) *empty_iters = host_state.empty;
    return finish(0);
}

}  // namespace

extern "C" int dp_kmeans_lloyd(const uint8_t *pixels, int64_t n, double *centers_host, int K, double tol,
                               int max_iter, void *nccl_comm, int check_every, int *n_iter, double *shift2,
                               unsigned long long *ties, int *empty_iters, void *stream)
{
    DP_RANGE("dp_kmeans_lloyd");
    DP_REQUIRE(pixels && centers_host, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS && n >= 0 && max_iter >= 1, "bad size");
    if (nccl_comm) DP_REQUIRE(g_nccl.handle, "NCCL communicator given but NCCL was never loaded");
    KmExchange ex;
    ex.nccl_comm = nccl_comm;
    return km_lloyd(pixels, n, centers_host, K, tol, max_iter, ex, check_every, n_iter, shift2, ties, empty_iters,
                    dp_stream(stream));
}

extern "C" int dp_kmeans_lloyd_p2p(const uint8_t *pixels, int64_t n, double *centers_host, int K, double tol,
                                   int max_iter, int rank, int world, void *const *inboxes, unsigned long long epoch,
                                   int check_every, int *n_iter, double *shift2, unsigned long long *ties,
                                   int *empty_iters, void *stream)
{
    DP_RANGE("dp_kmeans_lloyd_p2p");
    DP_REQUIRE(pixels && centers_host && inboxes, "null argument");
    DP_REQUIRE(K >= 1 && K <= DP_MAX_COLORS && n >= 0 && max_iter >= 1, "bad size");
    DP_REQUIRE(world >= 1 && world <= KM_P2P_RANKS && rank >= 0 && rank < world, "bad rank / world (at most 8 ranks)");
    for (int r = 0; r < world; ++r) DP_REQUIRE(inboxes[r], "null inbox");
    KmExchange ex;
    ex.rank = rank;
    ex.world = world;
    ex.inboxes = inboxes;
    ex.epoch = epoch;
    return km_lloyd(pixels, n, centers_host, K, tol, max_iter, ex, check_every, n_iter, shift2, ties, empty_iters,
                    dp_stream(stream));
}

// ---- peer-visible memory (cudaIpc) for the inboxes ------------------------------------------------
extern "C" int dp_p2p_inbox_bytes(void) { return (int)KM_P2P_BYTES; }

extern "C" int dp_p2p_alloc(size_t bytes, void **dptr, void *ipc_handle64)
{
    DP_REQUIRE(dptr && ipc_handle64 && bytes, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    DP_CUDA(cudaMalloc(&p, bytes));
    DP_CUDA(cudaMemset(p, 0, bytes));
    DP_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        dp_set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return 1;
    }
    memcpy(ipc_handle64, &h, 64);
    *dptr = p;
    return 0;
}

extern "C" int dp_p2p_open(const void *ipc_handle64, void **dptr)
{
    DP_REQUIRE(dptr && ipc_handle64, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, 64);
    DP_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int dp_p2p_close(void *dptr)
{
    if (!dptr) return 0;
    DP_CUDA(cudaIpcCloseMemHandle(dptr));
    return 0;
}

extern "C" int dp_p2p_free(void *dptr)
{
    if (!dptr) return 0;
    DP_CUDA(cudaFree(dptr));
    return 0;
}

#ifdef DP_KM_TIMING
extern "C" int dp_debug_km_timing(unsigned long long *out, int reset)
{
    DP_CUDA(cudaMemcpyFromSymbol(out, g_km_timing, sizeof(unsigned long long) * 8));
    if (reset) {
        static unsigned long long zeros[8];
        DP_CUDA(cudaMemcpyToSymbol(g_km_timing, zeros, sizeof(zeros)));
    }
    return 0;
}
#endif
