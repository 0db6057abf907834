// Host-side helper for vol2h5ParOSol (voxelFE.py:366-456).  The reference collects its boundary-condition
// rows (z, y, x, dof) in a Python set and writes list(set): the rows come out in the iteration order of
// CPython's hash table.  That order is a deterministic function of the insertion sequence:
//   * hash of a tuple of small non-negative ints (Objects/tupleobject.c, CPython >= 3.8: the xxHash-style
//     accumulator over the element hashes; hash(int n) = n for 0 <= n < 2^61 - 1),
//   * open addressing with 9 linear probes and the (5 i + 1 + perturb) recurrence, growth to 4x (2x above
//     50000 entries) the number of entries whenever the table is 3/5 full, re-insertion in table order
//     (Objects/setobject.c, CPython >= 3.7).
// This file replays that table in C so that a 1024^2 boundary plane (10^7 insertions) costs a fraction of a
// second instead of ten.  The Python side checks the replay against the running interpreter's own set on a
// sample before it trusts it, and falls back to the interpreter's set otherwise.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

typedef uint64_t uh;

inline uh tuple4_hash(const int64_t* r)
{
    const uh P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uh acc = P5;
    for (int i = 0; i < 4; ++i) {
        acc += (uh)r[i] * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += (uh)4 ^ (P5 ^ 3527539ULL);
    if (acc == (uh)-1) return 1546275796ULL;
    return acc;
}

struct Entry {
    uh hash;
    int64_t row;      // index of the row in the input, -1 = unused
};

const int LINEAR_PROBES = 9;
const int PERTURB_SHIFT = 5;

inline void insert_clean(Entry* table, size_t mask, int64_t row, uh hash)
{
    size_t perturb = (size_t)hash;
    size_t i = (size_t)hash & mask;
    while (true) {
        Entry* e = table + i;
        if (e->row < 0) { e->row = row; e->hash = hash; return; }
        if (i + LINEAR_PROBES <= mask) {
            for (int j = 0; j < LINEAR_PROBES; ++j) {
                ++e;
                if (e->row < 0) { e->row = row; e->hash = hash; return; }
            }
        }
        perturb >>= PERTURB_SHIFT;
        i = (i * 5 + 1 + perturb) & mask;
    }
}

}  // namespace

// rows: n x 4 int64 (0 <= value < 2^61 - 1) in insertion order, duplicates allowed.  out: the distinct rows
// in the order list(set(map(tuple, rows))) has under CPython >= 3.8 (capacity n x 4); *n_out = their number.
extern "C" int cfe_host_pyset_order(const int64_t* rows, int64_t n, int64_t* out, int64_t* n_out)
{
    *n_out = 0;
    if (n <= 0) return 0;
    for (int64_t k = 0; k < 4 * n; ++k)
        if (rows[k] < 0 || rows[k] >= ((int64_t)1 << 61) - 1) { cfe_set_error("cfe_host_pyset_order: value out of range"); return -1; }
    size_t mask = 7, fill = 0;
    std::vector<Entry> table(mask + 1);
    for (auto& e : table) { e.hash = 0; e.row = -1; }
    for (int64_t k = 0; k < n; ++k) {
        const int64_t* r = rows + 4 * k;
        const uh hash = tuple4_hash(r);
        // set_add_entry
        size_t perturb = (size_t)hash;
        size_t i = (size_t)hash & mask;
        Entry* slot = nullptr;
        bool found = false;
        while (!slot && !found) {
            Entry* e = table.data() + i;
            int probes = (i + LINEAR_PROBES <= mask) ? LINEAR_PROBES : 0;
            do {
                if (e->row < 0) { slot = e; break; }
                if (e->hash == hash && std::memcmp(rows + 4 * e->row, r, 4 * sizeof(int64_t)) == 0) { found = true; break; }
                ++e;
            } while (probes--);
            if (slot || found) break;
            perturb >>= PERTURB_SHIFT;
            i = (i * 5 + 1 + perturb) & mask;
        }
        if (found) continue;
        slot->row = k;
        slot->hash = hash;
        ++fill;
        if (fill * 5 < mask * 3) continue;
        // set_table_resize(used > 50000 ? used * 2 : used * 4); no deletions here, so used == fill
        const size_t minused = fill > 50000 ? fill * 2 : fill * 4;
        size_t newsize = 8;
        while (newsize <= minused) newsize <<= 1;
        std::vector<Entry> bigger(newsize);
        for (auto& e : bigger) { e.hash = 0; e.row = -1; }
        for (size_t q = 0; q <= mask; ++q)
            if (table[q].row >= 0) insert_clean(bigger.data(), newsize - 1, table[q].row, table[q].hash);
        table.swap(bigger);
        mask = newsize - 1;
    }
    int64_t m = 0;
    for (size_t q = 0; q <= mask; ++q)
        if (table[q].row >= 0) {
            std::memcpy(out + 4 * m, rows + 4 * table[q].row, 4 * sizeof(int64_t));
            ++m;
        }
    *n_out = m;
    return 0;
}
