// dp_common.cuh -- shared host/device definitions of libditherpie_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ditherpie_b200.h"

// NVTX range around an entry point of the C ABI (host side; a no-op unless a profiler is attached)
struct DpRange {
    explicit DpRange(const char *name) { nvtxRangePushA(name); }
    ~DpRange() { nvtxRangePop(); }
    DpRange(const DpRange &) = delete;
    DpRange &operator=(const DpRange &) = delete;
};
#define DP_RANGE(name) DpRange dp_range_guard_(name)

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libditherpie_b200 targets sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
void dp_set_error(const char *fmt, ...);

#define DP_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            dp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

#define DP_REQUIRE(cond, msg)                                      \
    do {                                                           \
        if (!(cond)) {                                             \
            dp_set_error("%s:%d %s", __FILE__, __LINE__, msg);     \
            return 2;                                              \
        }                                                          \
    } while (0)

#define DP_LAUNCH_CHECK() DP_CUDA(cudaGetLastError())

#define DP_ED_PAD 4096u        /* row 256 of the float4 row array: the pad row */
#define DP_ED_OVERFLOW 0xffffu

static inline cudaStream_t dp_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------
// Palette as the kernels see it (plain device pointers, passed by value)
// ---------------------------------------------------------------------------------------
struct PalDev {
    int K;
    int integral;          // every palette value is an integer in [0,255] -> exact int path
    int kd_nodes;
    const float *pal_f32;  // [K,3]
    const double *pal_f64; // [K,3]
    const int4 *coef;      // [K] integer path: (-2r<<8, -2g<<8, -2b<<8, (|p|^2<<8)|i)
    const uint8_t *out_rgb;  // [K,4] (r,g,b,0) bytes written for each palette row
    const uint8_t *in_lut;   // [256]
    // scipy KD-tree, flattened pre-order
    const int *kd_split_dim;
    const double *kd_split;
    const int *kd_start, *kd_end, *kd_lesser, *kd_greater, *kd_indices;
    double kd_mins[3], kd_maxes[3];
    // top-2 candidate table for byte-valued pixels (integral palettes only): one 8-byte entry
    // per (256>>thr_shift)^3 cell, built by exhaustive enumeration of the 2^24 colours.
    //   x & 0xff = n (<= 7): candidates in bytes 1..7 of the entry, ascending
    //   x & 0xff = 0xff    : n = x >> 8, candidates at thr_ovf[y .. y+n)
    const uint2 *thr_table;
    const uint8_t *thr_ovf;
    int thr_shift;   // 4 -> 16^3 cells, 3 -> 32^3 cells
    int thr_cells;
    // compact top-2 table for the v4 threshold kernel (integral palettes with K <= 30): one u32
    // per 8x8x8 colour cell = four candidate slots, each the byte offset row*8 of the row in the
    // kernel's int2 row array (free slots: K*8, a pad row that never wins).  A cell with more
    // than four candidates holds 0xf8000000 | n: its eight 4x4x4 sub-cells have entries of the
    // same format at thr4_sub[8n + ((r>>2)&1)*4 + ((g>>2)&1)*2 + ((b>>2)&1)]; a sub-cell that
    // still has more than four candidates has 0xf8 (row 31, also a pad row) in its last slot.
    const uint32_t *thr4_table;   // [32768] or null
    const uint32_t *thr4_sub;     // [8 * thr4_nsub]
    int thr4_nsub;
    // the same for plain nearest-colour quantisation: rows that can be NEAREST (ties included),
    // three slots per entry; sub-cell entries are read from global memory (L1)
    const uint32_t *near3_table;  // [32768] or null
    const uint32_t *near3_sub;    // [8 * near3_nsub]
    int near3_nsub;
    // thr4_wide = 1 (31 <= K <= 256): the compact tables hold plain row numbers, one per byte.
    //   top-2 table: four DISTINCT rows in ascending order (a cell with fewer candidates is filled
    //     up with other rows -- harmless: a row that is not a candidate is never among the two
    //     nearest anywhere in the cell), so the top byte is >= 3; an entry < 0x03000000 says "more
    //     than four candidates" and holds the number n of its eight sub-cell entries in thr4_sub
    //     (same format; a sub-cell entry < 0x03000000 still has more than four: exact path).
    //   nearest table: three distinct ascending rows in bytes 0..2, top byte 0xff; an entry
    //     < 0xff000000 is the refinement marker.
    //   The sub-cell entries stay in global memory (L1).
    int thr4_wide;
    // Exception table of the byte colours with an exact distance tie among their three nearest
    // rows (integral palettes): scipy's answers, replayed once at palette creation.
    //   x = colour (r | g<<8 | b<<16) | nearest row of query(k=1) << 24
    //   y = first | second << 8 row of query(k=2)
    // sorted by colour; tie_n < 0: not built (the kernels replay the KD-tree themselves).
    const uint2 *tie_table;
    int tie_n;
    // bucket index of the tie table: entries whose colour key >> 8 (= g | b << 8) equals k are
    // tie_table[tie_idx[k] .. tie_idx[k+1]) -- a lookup is one index load and a search among the
    // (on average < 1) entries of the bucket instead of ~16 dependent loads of a whole-table
    // binary search.  [65537] or null.
    const uint32_t *tie_idx;
    // nearest-row candidates for arbitrary real values in [0,255]^3 (diffusion modes): 32^3 cells
    // of 8x8x8; a row is dropped from a cell only if another row is strictly nearer at EVERY
    // point of the cell's closed box (exact linear test).  Two levels, because few candidate
    // sets are distinct (141 / 939 / 4 106 for the 16 / 64 / 256-colour test palettes):
    //   ed_l1  [32768] u16   cell -> pattern number (the kernels keep this level in shared memory)
    //   ed_pat [ed_npat]     one 16-byte pattern = eight u16 slots holding row*16 (the byte offset
    //                        of the row in the kernels' float4 row array), ascending; free slots
    //                        hold DP_ED_PAD (the offset of a far-away pad row).
    // More than seven candidates (rare): the cell gets a pattern of its own with slot 7 =
    // DP_ED_OVERFLOW and slots 0..6 the first seven; the full list is found through
    // ed_ovf_cells (sorted cell numbers) -> ed_ovf_off -> ed_ovf.
    const uint16_t *ed_l1;
    const uint4 *ed_pat;
    int ed_npat;
    int ed_gt4;                   // cells with more than four candidates
    // the same information flattened, ed_flat[cell] = ed_pat[ed_l1[cell]] (512 KB): one L1-cached
    // load per pixel for saturating batches, where shared memory is better spent on more warps
    const uint4 *ed_flat;
    const int *ed_ovf_cells;      // [ed_novf] ascending
    const uint32_t *ed_ovf_off;   // [ed_novf + 1] offsets into ed_ovf
    const uint8_t *ed_ovf;        // concatenated candidate lists
    int ed_novf;
};

struct dp_palette {
    PalDev dev;
    int has_lut;   // in_lut is not the identity
    int pal_is_out;   // every palette value equals the row's output byte (plain byte palettes)
    int device;
    void *blob;    // single device allocation backing the palette arrays and the KD-tree
    void *thr_table;
    void *thr_ovf;
    void *thr4_table;
    void *thr4_sub;
    void *near3_table;
    void *near3_sub;
    void *tie_table;
    void *tie_idx;
    void *ed_table;   // one allocation: level 1 | patterns | flat
    void *ed_ovf;     // one allocation: cells | offsets | lists
    // lazily built twin with unbounded outer cells (dp_palette_ext): tables, a device copy of
    // `dev` pointing at them, and the two counters the launch code needs
    void *ext_table;
    void *ext_ovf;
    void *ext_dev;
    int ext_npat, ext_gt4;
    float host_pal[DP_MAX_COLORS * 3];
};

// descriptor + nearest-row tables with unbounded outer cells, for the modes that look up
// UNCLAMPED work values (built on first use)
int dp_palette_ext(dp_palette *pal, const PalDev **dev, int *npat, int *gt4);
// number of SMs of the current device (cached)
int dp_num_sms();
// make the device's default stream-ordered memory pool keep its memory between calls
int dp_retain_pool(int device);
