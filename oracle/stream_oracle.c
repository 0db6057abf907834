/*
 * TEST INFRASTRUCTURE ONLY -- section-streamed C restatement of ciclope's vol2ugrid + mesh2voxelfe for
 * volumes whose materialised mesh (points 24 B / grid node, cells 64 B / element) does not fit the
 * test box: the deck is produced straight from the uint8 volume, section by section, into a SHA-256
 * sink (and, for small cases, into a file so that the streamed text itself can be compared byte for
 * byte with the golden decks of the unmodified reference -- tests/test_oracle_golden.py does that for
 * every golden deck case, which is what pins this file).
 *
 * Never imported, linked or executed by the product (ciclope_b200/).
 *
 * Reference lines followed (paths relative to the ciclope checkout, src/ciclope/core/voxelFE.py):
 *   element test / numbering   :140-142  (GV > GVmin, cell_i in raster order)
 *   node ids of a cell         :145-154
 *   existing nodes             :628      (np.unique(cells): a grid node is written iff one of its <= 8
 *                                         adjacent voxels is an element)
 *   *NODE lines                :629-630  ('{0:10d}, {n[0]:12.6f}, {n[1]:12.6f}, {n[2]:12.6f}\n'; the three
 *                                         fields come pre-formatted per axis from the Python wrapper, which
 *                                         evaluates the reference's own expression voxelsize[a] * index and
 *                                         '{:12.6f}'.format with the interpreter)
 *   *ELEMENT blocks            :633-647  (ascending grey value; ascending cell index inside a block)
 *   *NSET / *ELSET sections    :650-688  (sets from orc_boundary_sets = :163-192, 195-251)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* from ciclope_oracle.c */
typedef struct { i64 *ids[6]; i64 n[6]; } orc_sets6;
int orc_boundary_sets(const uint8_t *vol, i64 Z, i64 Y, i64 X, int thr, int strategy, i64 p, i64 k,
                      orc_sets6 *nodes, orc_sets6 *cells);
void orc_free_sets(orc_sets6 *s);

/* ------------------------------------------------------------------ SHA-256 (FIPS 180-4) */

typedef struct { uint32_t h[8]; uint64_t len; uint8_t buf[64]; size_t fill; } sha256;

static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};

#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))

static void sha_block(sha256 *s, const uint8_t *p)
{
    uint32_t w[64], a, b, c, d, e, f, g, h;
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    a = s->h[0]; b = s->h[1]; c = s->h[2]; d = s->h[3]; e = s->h[4]; f = s->h[5]; g = s->h[6]; h = s->h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25), ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K256[i] + w[i];
        uint32_t S0 = ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}

static void sha_init(sha256 *s)
{
    static const uint32_t h0[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(s->h, h0, sizeof h0);
    s->len = 0;
    s->fill = 0;
}

static void sha_update(sha256 *s, const uint8_t *p, size_t n)
{
    s->len += n;
    if (s->fill) {
        size_t take = 64 - s->fill < n ? 64 - s->fill : n;
        memcpy(s->buf + s->fill, p, take);
        s->fill += take; p += take; n -= take;
        if (s->fill == 64) { sha_block(s, s->buf); s->fill = 0; }
    }
    while (n >= 64) { sha_block(s, p); p += 64; n -= 64; }
    if (n) { memcpy(s->buf, p, n); s->fill = n; }
}

static void sha_final(sha256 *s, uint8_t out[32])
{
    uint64_t bits = s->len * 8;
    uint8_t pad[72] = {0x80};
    size_t padn = (s->fill < 56 ? 56 : 120) - s->fill;
    for (int i = 0; i < 8; i++) pad[padn + i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(s, pad, padn + 8);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(s->h[i] >> 24); out[4 * i + 1] = (uint8_t)(s->h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(s->h[i] >> 8); out[4 * i + 3] = (uint8_t)s->h[i];
    }
}

/* one-shot hash for the unit test of the implementation itself */
void orc_sha256(const uint8_t *p, i64 n, uint8_t out[32])
{
    sha256 s;
    sha_init(&s);
    sha_update(&s, p, (size_t)n);
    sha_final(&s, out);
}

/* ------------------------------------------------------------------ section sink */

#define SINK_BUF (1 << 20)
typedef struct { sha256 sha; i64 bytes; FILE *f; uint8_t *buf; size_t n; } sink;

static void sk_flush(sink *k)
{
    if (!k->n) return;
    sha_update(&k->sha, k->buf, k->n);
    if (k->f) fwrite(k->buf, 1, k->n, k->f);
    k->bytes += (i64)k->n;
    k->n = 0;
}
static inline uint8_t *sk_room(sink *k, size_t want)
{
    if (k->n + want > SINK_BUF) sk_flush(k);
    return k->buf + k->n;
}
static void sk_puts(sink *k, const char *s)
{
    size_t l = strlen(s);
    memcpy(sk_room(k, l), s, l);
    k->n += l;
}
/* '{}'.format(non-negative int) */
static inline size_t put_int(uint8_t *p, i64 v)
{
    char t[24];
    int l = 0;
    uint64_t u = (uint64_t)v;
    do { t[l++] = (char)('0' + u % 10); u /= 10; } while (u);
    for (int i = 0; i < l; i++) p[i] = (uint8_t)t[l - 1 - i];
    return (size_t)l;
}
static void sk_section_begin(sink *k) { sha_init(&k->sha); k->bytes = 0; k->n = 0; }

typedef struct { i64 bytes; uint8_t digest[32]; } orc_section;

static void sk_section_end(sink *k, orc_section *out)
{
    sk_flush(k);
    out->bytes = k->bytes;
    sha_final(&k->sha, out->digest);
}

/* voxelFE.py:657-666: ids ten per line, ',' after entries 1-9 of a line, '\n' after the tenth, then
 * one unconditional '\n' */
static void sk_idlist(sink *k, const i64 *ids, i64 n)
{
    int cr = 1;
    for (i64 i = 0; i < n; i++) {
        uint8_t *p = sk_room(k, 32);
        size_t l = put_int(p, ids[i]);
        if (cr < 10) p[l++] = ',';
        else { p[l++] = '\n'; cr = 0; }
        cr++;
        k->n += l;
    }
    *sk_room(k, 1) = '\n';
    k->n += 1;
}

/* Streams the deck body of mesh2voxelfe(vol2ugrid(vol, voxelsize, GVmin, ...)) from
 * '** Node coordinates from input model' to the end of the *ELSET section.
 *   thr          integer t with (GV > GVmin) <=> (GV > t)
 *   xt / yt / zt '{:12.6f}'.format(voxelsize[a] * i) for i = 0..X / 0..Y / 0..Z, `stride` bytes per
 *                entry, NUL terminated (formatted by the Python wrapper)
 *   strategy     0 "plane" (p = plane_lock_num), 1 "exact" (k = node_level_lock_num)
 *   path         NULL, or a file that receives the same bytes (preceded by nothing: the caller
 *                writes the three header comment lines)
 *   out[0..3]    node LINES (without the two lines '** Node coordinates...' / '*NODE'), the element
 *                section from its '** Elements ...' comment, the NSET section, the ELSET section
 *                (bytes 0 when the keyword is absent)
 * Returns 0, or -1 when the volume is too large for the 32-bit bucket indices. */
int orc_stream_deck(const uint8_t *vol, i64 Z, i64 Y, i64 X, int thr, const char *xt, const char *yt,
                    const char *zt, int stride, const char *eltype, int strategy, i64 p, i64 k,
                    int want_nset, int want_elset, const char *path, orc_section out[4], i64 *n_el_out,
                    i64 *n_nd_out)
{
    const i64 RN = X + 1, SN = (X + 1) * (Y + 1), NV = Z * Y * X;
    if (NV >= ((i64)1 << 32)) return -1;
    sink sk;
    memset(&sk, 0, sizeof sk);
    sk.buf = (uint8_t *)malloc(SINK_BUF);
    sk.f = path ? fopen(path, "ab") : NULL;
    if (path && !sk.f) { free(sk.buf); return -2; }
    if (sk.f) fputs("** Node coordinates from input model\n*NODE\n", sk.f);
    size_t *xl = (size_t *)malloc(sizeof(size_t) * (size_t)(X + 1));
    size_t *yl = (size_t *)malloc(sizeof(size_t) * (size_t)(Y + 1));
    size_t *zl = (size_t *)malloc(sizeof(size_t) * (size_t)(Z + 1));
    for (i64 i = 0; i <= X; i++) xl[i] = strlen(xt + i * stride);
    for (i64 i = 0; i <= Y; i++) yl[i] = strlen(yt + i * stride);
    for (i64 i = 0; i <= Z; i++) zl[i] = strlen(zt + i * stride);

    /* ---- *NODE lines: every grid node in raster order that belongs to an element (:628-630) */
    sk_section_begin(&sk);
    i64 n_nd = 0;
    for (i64 s = 0; s <= Z; s++)
        for (i64 r = 0; r <= Y; r++)
            for (i64 c = 0; c <= X; c++) {
                int used = 0;
                for (i64 ds = 0; ds < 2 && !used; ds++)
                    for (i64 dr = 0; dr < 2 && !used; dr++)
                        for (i64 dc = 0; dc < 2 && !used; dc++) {
                            i64 vs = s - ds, vr = r - dr, vc = c - dc;
                            if (vs < 0 || vs >= Z || vr < 0 || vr >= Y || vc < 0 || vc >= X) continue;
                            used = (int)vol[(vs * Y + vr) * X + vc] > thr;
                        }
                if (!used) continue;
                n_nd++;
                uint8_t *q = sk_room(&sk, 32 + 3 * (size_t)stride + 8);
                uint8_t tmp[24];
                size_t l = put_int(tmp, SN * s + RN * r + c), o = 0;
                for (size_t i = l; i < 10; i++) q[o++] = ' ';        /* '{0:10d}' */
                memcpy(q + o, tmp, l); o += l;
                q[o++] = ','; q[o++] = ' ';
                memcpy(q + o, xt + c * stride, xl[c]); o += xl[c];
                q[o++] = ','; q[o++] = ' ';
                memcpy(q + o, yt + r * stride, yl[r]); o += yl[r];
                q[o++] = ','; q[o++] = ' ';
                memcpy(q + o, zt + s * stride, zl[s]); o += zl[s];
                q[o++] = '\n';
                sk.n += o;
            }
    sk_section_end(&sk, &out[0]);

    /* ---- *ELEMENT blocks, ascending grey value, ascending cell index inside a block (:633-647) */
    i64 hist[256], first[257];
    memset(hist, 0, sizeof hist);
    i64 n_el = 0;
    for (i64 i = 0; i < NV; i++)
        if ((int)vol[i] > thr) { hist[vol[i]]++; n_el++; }
    first[0] = 0;
    for (int g = 0; g < 256; g++) first[g + 1] = first[g] + hist[g];
    uint32_t *bvox = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_el ? n_el : 1));
    uint32_t *beid = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_el ? n_el : 1));
    {
        i64 cur[256], cell_i = 0;
        memcpy(cur, first, sizeof cur);
        for (i64 i = 0; i < NV; i++)
            if ((int)vol[i] > thr) {
                bvox[cur[vol[i]]] = (uint32_t)i;
                beid[cur[vol[i]]++] = (uint32_t)cell_i++;
            }
    }
    sk_section_begin(&sk);
    sk_puts(&sk, "** Elements and Element sets from input model\n");
    for (int g = 0; g < 256; g++) {
        if (!hist[g]) continue;
        sk_puts(&sk, "*ELEMENT, TYPE=");
        sk_puts(&sk, eltype);
        sk_puts(&sk, ", ELSET=SET");
        { uint8_t *q = sk_room(&sk, 8); size_t l = put_int(q, g); q[l++] = '\n'; sk.n += l; }   /* 'SET' + str(int(GV)) */
        for (i64 j = first[g]; j < first[g + 1]; j++) {
            const i64 v = bvox[j];
            const i64 s = v / (Y * X), rem = v % (Y * X), r = rem / X, c = rem % X;
            const i64 b = SN * s + RN * r + c;
            const i64 nd[8] = {b, b + 1, b + RN + 1, b + RN, b + SN, b + SN + 1, b + SN + RN + 1, b + SN + RN};
            uint8_t *q = sk_room(&sk, 9 * 21);
            size_t o = put_int(q, (i64)beid[j] + 1);
            for (int t = 0; t < 8; t++) { q[o++] = ','; o += put_int(q + o, nd[t]); }
            q[o++] = '\n';
            sk.n += o;
        }
    }
    sk_section_end(&sk, &out[1]);
    free(bvox);
    free(beid);

    /* ---- *NSET / *ELSET (:650-688) */
    memset(&out[2], 0, 2 * sizeof(orc_section));
    if (want_nset || want_elset) {
        orc_sets6 ns, cs;
        orc_boundary_sets(vol, Z, Y, X, thr, strategy, p, k, &ns, &cs);
        static const char *NN[8] = {"NODES_X0", "NODES_X1", "NODES_Y0", "NODES_Y1", "NODES_Z0", "NODES_Z1",
                                    "NODES_X0Y0Z0", "NODES_X0Y0Z1"};
        static const char *CN[6] = {"CELLS_X0", "CELLS_X1", "CELLS_Y0", "CELLS_Y1", "CELLS_Z0", "CELLS_Z1"};
        if (want_nset) {
            sk_section_begin(&sk);
            sk_puts(&sk, "** Additional nset from voxel model. New generated nsets are:\n");
            sk_puts(&sk, "** NODES_X0, \n");                       /* '** {}, \n'.format(*names) (:652) */
            for (int q = 0; q < 8; q++) {
                sk_puts(&sk, "*NSET, NSET=");
                sk_puts(&sk, NN[q]);
                sk_puts(&sk, "\n");
                if (q < 6) sk_idlist(&sk, ns.ids[q], ns.n[q]);
                else {                                              /* NODES_Z0[0:2], NODES_Z1[0:2] (:250-251) */
                    const int src = q == 6 ? 4 : 5;
                    sk_idlist(&sk, ns.ids[src], ns.n[src] < 2 ? ns.n[src] : 2);
                }
            }
            sk_section_end(&sk, &out[2]);
        }
        if (want_elset) {
            sk_section_begin(&sk);
            sk_puts(&sk, "** Additional elset from voxel model. New generated elsets are:\n");
            sk_puts(&sk, "** CELLS_X0, \n");
            for (int q = 0; q < 6; q++) {
                sk_puts(&sk, "*ELSET, ELSET=");
                sk_puts(&sk, CN[q]);
                sk_puts(&sk, "\n");
                sk_idlist(&sk, cs.ids[q], cs.n[q]);
            }
            sk_section_end(&sk, &out[3]);
        }
        orc_free_sets(&ns);
        orc_free_sets(&cs);
    }
    if (sk.f) fclose(sk.f);
    free(sk.buf); free(xl); free(yl); free(zl);
    *n_el_out = n_el;
    *n_nd_out = n_nd;
    return 0;
}
