/* oracle.c -- CPU restatement of the reference's voxel hot path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may link, import or
 * call this file; it is the checker the CUDA path is compared against (tests/,
 * __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * Parity pin: validated against the unmodified reference run in the build
 * container (tests/test_oracle_vs_reference.py, tests/golden/), and against the
 * known-answer hashes of SURVEY.md Appendix D (tests/test_oracle_golden.py).
 *
 * Each function cites the reference lines (relative to /root/reference/) whose
 * behaviour it restates.  Grids are C-order [X][Y][Z], Z fastest.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC oracle.c -o liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* tiny open-addressing set of 64-bit keys                                     */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t *slot; uint8_t *used; size_t cap, n; } set64;

static void set_init(set64 *s, size_t cap) {
    s->cap = cap; s->n = 0;
    s->slot = (uint64_t *)malloc(cap * sizeof(uint64_t));
    s->used = (uint8_t *)calloc(cap, 1);
}
static void set_free(set64 *s) { free(s->slot); free(s->used); }
static uint64_t mix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
static void set_add(set64 *s, uint64_t k);
static void set_grow(set64 *s) {
    set64 t; set_init(&t, s->cap * 2);
    for (size_t i = 0; i < s->cap; ++i) if (s->used[i]) set_add(&t, s->slot[i]);
    set_free(s); *s = t;
}
static void set_add(set64 *s, uint64_t k) {
    if ((s->n + 1) * 2 > s->cap) set_grow(s);
    size_t i = mix64(k) & (s->cap - 1);
    while (s->used[i]) { if (s->slot[i] == k) return; i = (i + 1) & (s->cap - 1); }
    s->used[i] = 1; s->slot[i] = k; s->n++;
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* sorted keys -> caller frees */
static uint64_t *set_sorted(const set64 *s) {
    uint64_t *out = (uint64_t *)malloc((s->n ? s->n : 1) * sizeof(uint64_t));
    size_t k = 0;
    for (size_t i = 0; i < s->cap; ++i) if (s->used[i]) out[k++] = s->slot[i];
    qsort(out, k, sizeof(uint64_t), cmp_u64);
    return out;
}
/* order-preserving map int32 -> uint32 so that unsigned key order == signed tuple order */
static inline uint64_t ord32(int32_t v) { return (uint64_t)((uint32_t)v ^ 0x80000000u); }
static inline int32_t unord32(uint64_t v) { return (int32_t)((uint32_t)v ^ 0x80000000u); }

/* ------------------------------------------------------------------------- */
/* A1  binning: atomgroup_to_voxels.py:131-142                                 */
/*   fraxels = positions(float32) @ transform(float64)      (:132)             */
/*   voxels  = floor(fraxels); fraxels -= voxels            (:133-134)         */
/*   for dim in (2,1,0): voxels -= (voxels[dim] // nbox[dim,dim]) * nbox[dim]  */
/*   positions = float32((voxels + fraxels) @ inv(transform))   (:142)         */
/* The float64 products are evaluated as the FMA chain                         */
/*   fma(p2,T[2][j], fma(p1,T[1][j], p0*T[0][j]))                              */
/* which equals numpy/OpenBLAS dgemm bit-for-bit on the probed hosts           */
/* (re-checked by tests/test_oracle_vs_reference.py).         */
/* ------------------------------------------------------------------------- */
static inline int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

void orc_bin_atoms(const float *pos, int64_t n, const double *T, const double *invT,
                   const int64_t *nbox, int64_t *vox_out, float *pos_out)
{
    for (int64_t i = 0; i < n; ++i) {
        double p0 = (double)pos[3 * i], p1 = (double)pos[3 * i + 1], p2 = (double)pos[3 * i + 2];
        double f[3]; int64_t v[3];
        for (int j = 0; j < 3; ++j) {
            f[j] = fma(p2, T[6 + j], fma(p1, T[3 + j], p0 * T[j]));
            double fl = floor(f[j]);
            v[j] = (int64_t)fl;
            f[j] = f[j] - (double)v[j];
        }
        for (int dim = 2; dim >= 0; --dim) {
            int64_t s = floordiv64(v[dim], nbox[4 * dim]);
            for (int j = 0; j < 3; ++j) v[j] -= s * nbox[3 * dim + j];
        }
        if (vox_out) { vox_out[3 * i] = v[0]; vox_out[3 * i + 1] = v[1]; vox_out[3 * i + 2] = v[2]; }
        if (pos_out) {
            double g0 = (double)v[0] + f[0], g1 = (double)v[1] + f[1], g2 = (double)v[2] + f[2];
            for (int j = 0; j < 3; ++j)
                pos_out[3 * i + j] = (float)fma(g2, invT[6 + j], fma(g1, invT[3 + j], g0 * invT[j]));
        }
    }
}

/* ------------------------------------------------------------------------- */
/* A5  labels: label.py:4-19 (scipy.ndimage.label, 3x3x3 structure, both       */
/* polarities, no wrap).  Restated as a raster-ordered flood fill: components  */
/* are discovered in C-order of their first voxel, which is the numbering      */
/* scipy produces.  True -> 1,2,..  False -> -1,-2,..          */
/* ------------------------------------------------------------------------- */
int orc_label(const uint8_t *grid, int64_t X, int64_t Y, int64_t Z, int32_t *labels,
              int32_t *n_true, int32_t *n_false)
{
    int64_t N = X * Y * Z;
    memset(labels, 0, (size_t)N * sizeof(int32_t));
    int64_t *stack = (int64_t *)malloc((size_t)(N ? N : 1) * sizeof(int64_t));
    if (!stack) return -1;
    int32_t nt = 0, nf = 0;
    for (int64_t s = 0; s < N; ++s) {
        if (labels[s]) continue;
        uint8_t val = grid[s] != 0;
        int32_t lab = val ? ++nt : -(++nf);
        int64_t top = 0; stack[top++] = s; labels[s] = lab;
        while (top) {
            int64_t c = stack[--top];
            int64_t z = c % Z, y = (c / Z) % Y, x = c / (Z * Y);
            for (int dx = -1; dx <= 1; ++dx) { int64_t nx = x + dx; if (nx < 0 || nx >= X) continue;
            for (int dy = -1; dy <= 1; ++dy) { int64_t ny = y + dy; if (ny < 0 || ny >= Y) continue;
            for (int dz = -1; dz <= 1; ++dz) { int64_t nz = z + dz; if (nz < 0 || nz >= Z) continue;
                int64_t q = (nx * Y + ny) * Z + nz;
                if (!labels[q] && (grid[q] != 0) == val) { labels[q] = lab; stack[top++] = q; }
            }}}
        }
    }
    free(stack);
    *n_true = nt; *n_false = nf;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* A7  contacts: find_label_contacts.pyx:41-77.  Every voxel against its 13    */
/* lexicographically-previous 26-neighbours, NO wrap, zero labels ignored,     */
/* rows (min,max), sorted unique.  Returns the number of rows; rows written    */
/* up to cap.                                                                   */
/* ------------------------------------------------------------------------- */
int64_t orc_contacts(const int32_t *L, int64_t X, int64_t Y, int64_t Z, int32_t *rows, int64_t cap)
{
    static const int sh[13][3] = {{-1,-1,-1},{-1,-1,0},{-1,-1,1},{-1,0,-1},{-1,0,0},{-1,0,1},
                                  {-1,1,-1},{-1,1,0},{-1,1,1},{0,-1,-1},{0,-1,0},{0,-1,1},{0,0,-1}};
    set64 s; set_init(&s, 1024);
    for (int64_t x = 0; x < X; ++x) for (int64_t y = 0; y < Y; ++y) for (int64_t z = 0; z < Z; ++z) {
        int32_t a = L[(x * Y + y) * Z + z];
        if (!a) continue;
        for (int i = 0; i < 13; ++i) {
            int64_t nx = x + sh[i][0], ny = y + sh[i][1], nz = z + sh[i][2];
            if (nx < 0 || nx >= X || ny < 0 || ny >= Y || nz < 0 || nz >= Z) continue;
            int32_t b = L[(nx * Y + ny) * Z + nz];
            if (!b || b == a) continue;
            int32_t lo = a < b ? a : b, hi = a < b ? b : a;
            set_add(&s, (ord32(lo) << 32) | ord32(hi));
        }
    }
    uint64_t *k = set_sorted(&s);
    int64_t n = (int64_t)s.n;
    for (int64_t i = 0; i < n && i < cap; ++i) {
        rows[2 * i] = unord32(k[i] >> 32); rows[2 * i + 1] = unord32(k[i] & 0xffffffffu);
    }
    free(k); set_free(&s);
    return n;
}

/* ------------------------------------------------------------------------- */
/* A6  bridges: find_bridges.pyx:21-89.  Face voxels only; each of the 26      */
/* neighbours that leaves the box is wrapped, shift = +1 through the max face, */
/* -1 through the min face; row (la,lb,s) if la<=lb else (lb,la,-s); sorted    */
/* unique.  Key packs labels order-preserving (two 32-bit halves) and needs a  */
/* third word for the shift, so two sets keyed by shift code are avoided by    */
/* sorting 128-bit records instead.                                             */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t hi; uint64_t lo; } rec128;
static int cmp_rec(const void *a, const void *b) {
    const rec128 *x = (const rec128 *)a, *y = (const rec128 *)b;
    if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    return 0;
}

int64_t orc_bridges(const int32_t *L, int64_t X, int64_t Y, int64_t Z, int32_t *rows, int64_t cap)
{
    /* per shift code (27) one set of label pairs keeps memory small */
    set64 sets[27];
    for (int i = 0; i < 27; ++i) set_init(&sets[i], 64);
    for (int64_t x = 0; x < X; ++x) for (int64_t y = 0; y < Y; ++y) {
        int bx = (x == 0 || x == X - 1), by = (y == 0 || y == Y - 1);
        for (int64_t z = 0; z < Z; ++z) {
            int bz = (z == 0 || z == Z - 1);
            if (!(bx || by || bz)) { z = Z - 2; continue; }   /* interior of the row: jump to z = Z-1 */
            int32_t a = L[(x * Y + y) * Z + z];
            if (!a) continue;
            for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
                if (!dx && !dy && !dz) continue;
                int64_t nx = x + dx, ny = y + dy, nz = z + dz; int sx = 0, sy = 0, sz = 0;
                if (nx < 0) { nx = X - 1; sx = -1; } else if (nx >= X) { nx = 0; sx = 1; }
                if (ny < 0) { ny = Y - 1; sy = -1; } else if (ny >= Y) { ny = 0; sy = 1; }
                if (nz < 0) { nz = Z - 1; sz = -1; } else if (nz >= Z) { nz = 0; sz = 1; }
                if (!(sx || sy || sz)) continue;
                int32_t b = L[(nx * Y + ny) * Z + nz];
                if (!b) continue;
                int32_t l1 = a, l2 = b;
                if (a > b) { l1 = b; l2 = a; sx = -sx; sy = -sy; sz = -sz; }
                int code = (sx + 1) * 9 + (sy + 1) * 3 + (sz + 1);
                set_add(&sets[code], (ord32(l1) << 32) | ord32(l2));
            }
        }
    }
    size_t total = 0;
    for (int i = 0; i < 27; ++i) total += sets[i].n;
    rec128 *r = (rec128 *)malloc((total ? total : 1) * sizeof(rec128));
    size_t k = 0;
    for (int c = 0; c < 27; ++c)
        for (size_t i = 0; i < sets[c].cap; ++i)
            if (sets[c].used[i]) { r[k].hi = sets[c].slot[i]; r[k].lo = (uint64_t)c; ++k; }
    qsort(r, total, sizeof(rec128), cmp_rec);
    for (size_t i = 0; i < total && (int64_t)i < cap; ++i) {
        int c = (int)r[i].lo;
        rows[5 * i] = unord32(r[i].hi >> 32); rows[5 * i + 1] = unord32(r[i].hi & 0xffffffffu);
        rows[5 * i + 2] = c / 9 - 1; rows[5 * i + 3] = (c / 3) % 3 - 1; rows[5 * i + 4] = c % 3 - 1;
    }
    free(r);
    for (int i = 0; i < 27; ++i) set_free(&sets[i]);
    return (int64_t)total;
}

/* ------------------------------------------------------------------------- */
/* A3/A4  morphology: atomgroup_to_voxels.py:32-92, 242-276.  With box =       */
/* diag(shape) and span 1 the in-place blur is the periodic 3x3x3 OR; erosion  */
/* is its De Morgan dual.  op: 'd' or 'e'.  in/out are 1 byte per voxel.       */
/* ------------------------------------------------------------------------- */
int orc_morph_step(const uint8_t *in, uint8_t *out, int64_t X, int64_t Y, int64_t Z, char op)
{
    if (op != 'd' && op != 'e') return -1;
    uint8_t want = (op == 'd');   /* dilate: any neighbour 1 -> 1 ; erode: any neighbour 0 -> 0 */
    for (int64_t x = 0; x < X; ++x) for (int64_t y = 0; y < Y; ++y) for (int64_t z = 0; z < Z; ++z) {
        uint8_t hit = 0;
        for (int dx = -1; dx <= 1 && !hit; ++dx) { int64_t nx = (x + dx + X) % X;
        for (int dy = -1; dy <= 1 && !hit; ++dy) { int64_t ny = (y + dy + Y) % Y;
        for (int dz = -1; dz <= 1 && !hit; ++dz) { int64_t nz = (z + dz + Z) % Z;
            if ((in[(nx * Y + ny) * Z + nz] != 0) == want) hit = 1;
        }}}
        out[(x * Y + y) * Z + z] = hit ? want : (uint8_t)!want;
    }
    return 0;
}
