// dp_pyset.cu -- host-side set-up helper for the default palette source (no device code).
//
// ColorReducer.reduce_colors (dithering_lib.py:1834-1843) feeds median cut with
// `list(set(image.getdata()))`: the unique colours in the ITERATION ORDER OF A CPYTHON SET of
// (r, g, b) tuples.  That order decides how equal keys fall around each median (list.sort is
// stable), so reproducing the reference's palettes needs exactly that order.  Building the set in
// the interpreter costs seconds per 1080p frame; this file replays CPython's algorithm on packed
// colours instead:
//   * tuple hash: the xxHash-style accumulator of Objects/tupleobject.c (CPython >= 3.8) over the
//     three small-int hashes (hash(n) == n for 0 <= n <= 255);
//   * table: Objects/setobject.c -- open addressing, LINEAR_PROBES = 9, perturb shift 5,
//     i = i * 5 + 1 + perturb; growth when fill * 5 >= mask * 3 to the first power of two above
//     used * 4 (used * 2 beyond 50 000 entries); a resize re-inserts the old table in slot order
//     with set_insert_clean; no deletions occur, so there are no dummy entries;
//   * iteration: slot order.
// tests/test_host_logic.py checks the replay against the running interpreter's own set.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "dp_common.cuh"

namespace {

constexpr uint64_t XXPRIME_1 = 11400714785074694791ull;
constexpr uint64_t XXPRIME_2 = 14029467366897019727ull;
constexpr uint64_t XXPRIME_5 = 2870177450012600261ull;

inline uint64_t rotl31(uint64_t x) { return (x << 31) | (x >> 33); }

// hash((r, g, b)) for ints in 0..255
inline uint64_t tuple3_hash(uint32_t c)
{
    uint64_t acc = XXPRIME_5;
    for (int k = 0; k < 3; ++k) {
        const uint64_t lane = (c >> (8 * k)) & 255u;
        acc += lane * XXPRIME_2;
        acc = rotl31(acc);
        acc *= XXPRIME_1;
    }
    acc += 3ull ^ (XXPRIME_5 ^ 3527539ull);
    if (acc == ~0ull) return 1546275796ull;
    return acc;
}

constexpr uint32_t OCC = 0x80000000u;   // slot in use; low 24 bits = r | g << 8 | b << 16
constexpr int LINEAR_PROBES = 9;
constexpr int PERTURB_SHIFT = 5;

inline void insert_clean(uint32_t *table, size_t mask, uint32_t c, uint64_t hash)
{
    size_t perturb = hash;
    size_t i = (size_t)hash & mask;
    for (;;) {
        uint32_t *e = table + i;
        int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
        do {
            if (!(*e & OCC)) {
                *e = c | OCC;
                return;
            }
            ++e;
        } while (probes--);
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

}  // namespace

extern "C" int dp_unique_colors_pyset_order(const uint8_t *rgb, int64_t npix, uint8_t *out_rgb,
                                            int64_t *n_unique)
{
    DP_RANGE("dp_unique_colors_pyset_order");
    DP_REQUIRE(rgb && out_rgb && n_unique && npix >= 0, "bad argument");
    size_t mask = 7;   // PySet_MINSIZE - 1
    uint32_t *table = static_cast<uint32_t *>(calloc(mask + 1, sizeof(uint32_t)));
    DP_REQUIRE(table, "out of host memory");
    size_t used = 0;   // == fill: nothing is ever deleted
    // adding a key the set already holds leaves the table untouched: a 2 MB bitmap of the colours
    // seen so far filters those pixels out before the (cache-unfriendly) table walk
    uint64_t *seen = static_cast<uint64_t *>(calloc((1u << 24) / 64, sizeof(uint64_t)));
    if (!seen) {
        free(table);
        DP_REQUIRE(false, "out of host memory");
    }
    for (int64_t p = 0; p < npix; ++p) {
        const uint32_t c = (uint32_t)rgb[3 * p] | ((uint32_t)rgb[3 * p + 1] << 8) | ((uint32_t)rgb[3 * p + 2] << 16);
        if (seen[c >> 6] & (1ull << (c & 63))) continue;
        seen[c >> 6] |= 1ull << (c & 63);
        const uint64_t hash = tuple3_hash(c);
        size_t perturb = hash;
        size_t i = (size_t)hash & mask;
        bool found = false;
        uint32_t *slot = nullptr;
        for (;;) {
            uint32_t *e = table + i;
            int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
            do {
                if (!(*e & OCC)) {
                    slot = e;
                    break;
                }
                if ((*e & 0xffffffu) == c) {   // equal hash and equal key
                    found = true;
                    break;
                }
                ++e;
            } while (probes--);
            if (slot || found) break;
            perturb >>= PERTURB_SHIFT;
            i = (i * 5 + 1 + perturb) & mask;
        }
        if (found) continue;
        *slot = c | OCC;
        ++used;
        if (used * 5 < mask * 3) continue;
        // set_table_resize(so, used > 50000 ? used * 2 : used * 4)
        const size_t minused = used > 50000 ? used * 2 : used * 4;
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        uint32_t *nt = static_cast<uint32_t *>(calloc(newsize, sizeof(uint32_t)));
        if (!nt) {
            free(table);
            free(seen);
            DP_REQUIRE(false, "out of host memory");
        }
        for (size_t k = 0; k <= mask; ++k)
            if (table[k] & OCC) insert_clean(nt, newsize - 1, table[k] & 0xffffffu, tuple3_hash(table[k] & 0xffffffu));
        free(table);
        table = nt;
        mask = newsize - 1;
    }
    int64_t n = 0;
    for (size_t k = 0; k <= mask; ++k) {
        if (!(table[k] & OCC)) continue;
        out_rgb[3 * n] = (uint8_t)table[k];
        out_rgb[3 * n + 1] = (uint8_t)(table[k] >> 8);
        out_rgb[3 * n + 2] = (uint8_t)(table[k] >> 16);
        ++n;
    }
    free(table);
    free(seen);
    *n_unique = n;
    return 0;
}

// generate_blue_noise (dithering_lib.py:381-399): farthest-point ordering of a shuffled
// coordinate list -- `max(coords, key=min_dist)` (first maximum in list order), value
// i / (n - 1 + 1e-9), then every remaining point's squared distance to the pick is folded into
// min_dist.  The distances are small integers (exact in the reference's f32 array), so the
// replay is integer arithmetic; the reference's pure-Python double loop takes 7 s at size 64 and
// minutes at 128.  `order`: the shuffled list as flat indices r * size + c (the caller shuffles
// with numpy's RandomState, as the reference does).
extern "C" int dp_blue_noise_from_order(const int32_t *order, int size, float *out)
{
    DP_RANGE("dp_blue_noise_from_order");
    DP_REQUIRE(order && out && size >= 1 && size <= 1024, "bad argument");
    const int n = size * size;
    int32_t *rr = static_cast<int32_t *>(malloc(sizeof(int32_t) * 3 * (size_t)n));
    DP_REQUIRE(rr, "out of host memory");
    int32_t *cc = rr + n, *mind = cc + n;
    for (int k = 0; k < n; ++k) {
        rr[k] = order[k] / size;
        cc[k] = order[k] % size;
        mind[k] = INT32_MAX;                 // +inf
    }
    const double denom = (double)(n - 1) + 1e-9;
    int live = n;                            // the list, compacted in place (order preserved)
    for (int i = 0; i < n; ++i) {
        int j = 0;
        int32_t best = mind[0];
        for (int k = 1; k < live; ++k)
            if (mind[k] > best) {            // strict: the first maximum wins
                best = mind[k];
                j = k;
            }
        const int br = rr[j], bc = cc[j];
        out[br * size + bc] = (float)((double)i / denom);
        --live;
        for (int k = j; k < live; ++k) {     // coords.remove(best)
            rr[k] = rr[k + 1];
            cc[k] = cc[k + 1];
            mind[k] = mind[k + 1];
        }
        for (int k = 0; k < live; ++k) {
            const int dr = rr[k] - br, dc = cc[k] - bc;
            const int32_t d2 = dr * dr + dc * dc;
            if (d2 < mind[k]) mind[k] = d2;
        }
    }
    free(rr);
    return 0;
}
