/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's variational
 * region growing, /root/reference/Code/variationalRegionGrowing.py (VRG:line).
 *
 * Same order-free, histogram-based semantics as oracle/vrg_oracle.py (which is
 * pinned to the unmodified reference through tests/golden/): this file exists so
 * the parity tests and bench.py's CPU baseline / `--impl reference` arm can run
 * the restatement at sizes NumPy cannot finish in seconds.  It is itself pinned
 * to the golden fixtures by tests/test_oracle_golden.py.  The product path never
 * links or loads it.
 *
 * Narrow-band like the reference: per iteration it scans the label volume once
 * for band voxels (VRG:79-88), applies the flips as set rules (VRG:165-230) and
 * re-classifies only the 3x3x3 neighbourhoods of executed flips.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o _build/libvrg_oracle.so vrg_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VRG_A 0.3989422804014327 /* (2*pi)**-0.5, VRG:7 */

enum { EXIT_CONVERGED = 0, EXIT_MAX_SEGMENT = 2, EXIT_MAX_ITER = 3 };
enum { ERR_BAD_LABEL = -1, ERR_EMPTY_SEED = -2, ERR_NO_BAND = -3, ERR_TOO_MANY_LEVELS = -4, ERR_NOMEM = -5 };

typedef struct {
    int64_t Z, Y, X;
    uint8_t *lab;
    uint8_t *flag;
} vol_t;

#define F_R 1
#define F_A0 2
#define F_AEX 4
#define F_DIRTY 8

/* ---- distinct intensity levels: open-addressing hash over the bit patterns ---- */
typedef struct { uint64_t *keys; uint8_t *used; int64_t cap, n; } hset_t;

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
static int64_t level_rank(const double *levels, int64_t L, double v) {
    int64_t lo = 0, hi = L - 1;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (levels[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

typedef struct { int64_t *v; int64_t n, cap; } list_t;
static void push(list_t *l, int64_t x) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 1024; l->v = (int64_t *)realloc(l->v, l->cap * sizeof(int64_t)); }
    l->v[l->n++] = x;
}

/* canonical label of voxel p from the seg-ness (lab <= 1) of its in-bounds 26 neighbours (VRG:21,139-145) */
static uint8_t classify(const vol_t *V, int64_t z, int64_t y, int64_t x) {
    int64_t p = (z * V->Y + y) * V->X + x;
    uint8_t me = V->lab[p];
    if (me == 4) return 4;
    int seg = me <= 1, hit = 0;
    for (int64_t dz = -1; dz <= 1 && !hit; dz++) {
        int64_t zz = z + dz; if (zz < 0 || zz >= V->Z) continue;
        for (int64_t dy = -1; dy <= 1 && !hit; dy++) {
            int64_t yy = y + dy; if (yy < 0 || yy >= V->Y) continue;
            for (int64_t dx = -1; dx <= 1; dx++) {
                int64_t xx = x + dx; if (xx < 0 || xx >= V->X) continue;
                int nseg = V->lab[(zz * V->Y + yy) * V->X + xx] <= 1;
                if (nseg != seg) { hit = 1; break; }
            }
        }
    }
    return seg ? (hit ? 1 : 0) : (hit ? 2 : 3);
}

/* Returns 0 or a negative error.  labels: in = initial valueMap (0/3/4), out = final canonical valueMap.
 * trace: rows of (n_flips, n_in, n_out), row 0 = init with n_flips = -1; capacity trace_cap rows.
 * quirk[4] = add_to_inside, remove_to_outside, cancel_repromoted, cancelled.
 * tables (optional): per decision, 2*L doubles (pin/n_in then pout/n_out), capacity tables_cap decisions;
 * levels_out (optional, capacity levels_cap) receives the sorted levels, *n_levels their count. */
int vrg_oracle_run(const double *data, uint8_t *labels, int64_t Z, int64_t Y, int64_t X, double H,
                   int64_t max_segment_size, int64_t iter_max, int64_t *iterations, int64_t *exit_code,
                   int64_t *trace, int64_t trace_cap, int64_t *n_trace, int64_t *quirk,
                   double *tables, int64_t tables_cap, double *levels_out, int64_t levels_cap,
                   int64_t *n_levels, int nthreads) {
    const int64_t N = Z * Y * X;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int64_t i = 0; i < N; i++)
        if (labels[i] != 0 && labels[i] != 3 && labels[i] != 4) return ERR_BAD_LABEL;

    /* levels */
    hset_t hs; hs.cap = 1 << 18; hs.n = 0;
    hs.keys = (uint64_t *)malloc(hs.cap * sizeof(uint64_t));
    hs.used = (uint8_t *)calloc(hs.cap, 1);
    for (int64_t i = 0; i < N; i++) {
        double v = data[i] + 0.0; /* -0.0 -> +0.0 */
        uint64_t k; memcpy(&k, &v, 8);
        uint64_t h = mix64(k) & (hs.cap - 1);
        while (hs.used[h] && hs.keys[h] != k) h = (h + 1) & (hs.cap - 1);
        if (!hs.used[h]) { hs.used[h] = 1; hs.keys[h] = k; if (++hs.n > 65536) { free(hs.keys); free(hs.used); return ERR_TOO_MANY_LEVELS; } }
    }
    int64_t L = hs.n;
    double *levels = (double *)malloc(L * sizeof(double));
    { int64_t j = 0; for (int64_t h = 0; h < hs.cap; h++) if (hs.used[h]) memcpy(&levels[j++], &hs.keys[h], 8); }
    free(hs.keys); free(hs.used);
    qsort(levels, L, sizeof(double), cmp_double);
    if (n_levels) *n_levels = L;
    if (levels_out) for (int64_t j = 0; j < L && j < levels_cap; j++) levels_out[j] = levels[j];
    uint16_t *lev = (uint16_t *)malloc(N * sizeof(uint16_t));
    if (!lev) return ERR_NOMEM;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) lev[i] = (uint16_t)level_rank(levels, L, data[i] + 0.0);

    vol_t V; V.Z = Z; V.Y = Y; V.X = X; V.lab = labels;
    V.flag = (uint8_t *)calloc(N, 1);
    int64_t *hist_in = (int64_t *)calloc(L, sizeof(int64_t));
    int64_t *hist_out = (int64_t *)calloc(L, sizeof(int64_t));
    double *pin = (double *)malloc(L * sizeof(double)), *pout = (double *)malloc(L * sizeof(double));
    uint8_t *dbit = (uint8_t *)malloc(L);
    int64_t n_in = 0, n_out = 0, n_excl = 0;
    list_t dirty = {0, 0, 0};

    /* init branch (VRG:129-145): absorb 4s around seeds, then classify the seeds' neighbourhoods */
    for (int64_t z = 0; z < Z; z++) for (int64_t y = 0; y < Y; y++) for (int64_t x = 0; x < X; x++) {
        int64_t p = (z * Y + y) * X + x;
        if (labels[p] != 0) continue;
        for (int64_t dz = -1; dz <= 1; dz++) { int64_t zz = z + dz; if (zz < 0 || zz >= Z) continue;
            for (int64_t dy = -1; dy <= 1; dy++) { int64_t yy = y + dy; if (yy < 0 || yy >= Y) continue;
                for (int64_t dx = -1; dx <= 1; dx++) { int64_t xx = x + dx; if (xx < 0 || xx >= X) continue;
                    int64_t q = (zz * Y + yy) * X + xx;
                    if (!(V.flag[q] & F_DIRTY)) { V.flag[q] |= F_DIRTY; push(&dirty, q); }
                } } }
    }
    for (int64_t i = 0; i < dirty.n; i++) if (labels[dirty.v[i]] == 4) labels[dirty.v[i]] = 3;
    {
        uint8_t *nl = (uint8_t *)malloc(dirty.n ? dirty.n : 1);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < dirty.n; i++) {
            int64_t p = dirty.v[i]; int64_t x = p % X, y = (p / X) % Y, z = p / (X * Y);
            nl[i] = classify(&V, z, y, x);
        }
        for (int64_t i = 0; i < dirty.n; i++) { labels[dirty.v[i]] = nl[i]; V.flag[dirty.v[i]] = 0; }
        free(nl);
    }
    int64_t n_band = 0;
    for (int64_t i = 0; i < N; i++) {
        uint8_t l = labels[i];
        if (l <= 1) { hist_in[lev[i]]++; n_in++; }
        else if (l <= 3) { hist_out[lev[i]]++; n_out++; }
        else n_excl++;
        n_band += (l == 1 || l == 2);
    }
    int rc = 0;
    if (n_in == 0) rc = ERR_EMPTY_SEED;
    else if (n_band == 0) rc = ERR_NO_BAND;
    int64_t nt = 0;
    if (trace && nt < trace_cap) { trace[0] = -1; trace[1] = n_in; trace[2] = n_out; }
    nt = 1;
    memset(quirk, 0, 4 * sizeof(int64_t));
    int64_t iter = 1, ex = EXIT_MAX_ITER;
    list_t Rl = {0, 0, 0}, Al = {0, 0, 0};
    const double mhH = -0.5 * H;

    while (rc == 0 && iter <= iter_max) {
        /* decision table over levels (VRG:79-87; Parzen sums VRG:151-155 as histogram mat-vec) */
#pragma omp parallel for schedule(static)
        for (int64_t b = 0; b < L; b++) {
            double si = 0.0, so = 0.0;
            for (int64_t c = 0; c < L; c++) {
                if (!(hist_in[c] | hist_out[c])) continue;
                double diff = levels[c] - levels[b];
                double kv = VRG_A * exp(mhH * (diff * diff));
                si += (double)hist_in[c] * kv;
                so += (double)hist_out[c] * kv;
            }
            pin[b] = si / (double)n_in;
            pout[b] = so / (double)n_out;
            dbit[b] = pin[b] >= pout[b];
        }
        if (tables && iter - 1 < tables_cap) {
            memcpy(tables + (iter - 1) * 2 * L, pin, L * sizeof(double));
            memcpy(tables + (iter - 1) * 2 * L + L, pout, L * sizeof(double));
        }
        /* flip lists: inner band leaves iff in < out, outer band enters iff in >= out (VRG:87-88) */
        Rl.n = 0; Al.n = 0;
#pragma omp parallel
        {
            list_t r = {0, 0, 0}, a = {0, 0, 0};
#pragma omp for schedule(static) nowait
            for (int64_t i = 0; i < N; i++) {
                uint8_t l = labels[i];
                if (l == 1) { if (!dbit[lev[i]]) push(&r, i); }
                else if (l == 2) { if (dbit[lev[i]]) push(&a, i); }
            }
#pragma omp critical
            {
                for (int64_t i = 0; i < r.n; i++) push(&Rl, r.v[i]);
                for (int64_t i = 0; i < a.n; i++) push(&Al, a.v[i]);
            }
            free(r.v); free(a.v);
        }
        int64_t n_flips = Rl.n + Al.n;
        if (n_flips == 0) { ex = EXIT_CONVERGED; break; }
        if (n_in >= max_segment_size) { ex = EXIT_MAX_SEGMENT; break; } /* before applying, VRG:101 */
        for (int64_t i = 0; i < Rl.n; i++) V.flag[Rl.v[i]] |= F_R;
        /* cancel rule: an addition needs a segmented neighbour that is not leaving (VRG:183-190, 198) */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < Al.n; i++) {
            int64_t p = Al.v[i]; int64_t x = p % X, y = (p / X) % Y, z = p / (X * Y);
            int ok = 0;
            for (int64_t dz = -1; dz <= 1 && !ok; dz++) { int64_t zz = z + dz; if (zz < 0 || zz >= Z) continue;
                for (int64_t dy = -1; dy <= 1 && !ok; dy++) { int64_t yy = y + dy; if (yy < 0 || yy >= Y) continue;
                    for (int64_t dx = -1; dx <= 1; dx++) { int64_t xx = x + dx; if (xx < 0 || xx >= X) continue;
                        int64_t q = (zz * Y + yy) * X + xx;
                        if (labels[q] <= 1 && !(V.flag[q] & F_R)) { ok = 1; break; }
                    } } }
            V.flag[p] |= ok ? (F_A0 | F_AEX) : F_A0;
        }
        /* apply seg-ness with temporary labels; absorb 4 -> 3 (VRG:167-168: 3^3 of every flip;
         * VRG:177-179,207-208: 5^3 of every executed flip); collect the dirty set */
        dirty.n = 0;
        for (int pass = 0; pass < 2; pass++) {
            list_t *Lst = pass ? &Al : &Rl;
            for (int64_t i = 0; i < Lst->n; i++) {
                int64_t p = Lst->v[i]; int64_t x = p % X, y = (p / X) % Y, z = p / (X * Y);
                int executed = pass ? ((V.flag[p] & F_AEX) != 0) : 1;
                int rad = executed ? 2 : 1;
                if (executed) labels[p] = pass ? 0 : 3;
                if (n_excl == 0 && !executed) continue;
                for (int64_t dz = -rad; dz <= rad; dz++) { int64_t zz = z + dz; if (zz < 0 || zz >= Z) continue;
                    for (int64_t dy = -rad; dy <= rad; dy++) { int64_t yy = y + dy; if (yy < 0 || yy >= Y) continue;
                        for (int64_t dx = -rad; dx <= rad; dx++) { int64_t xx = x + dx; if (xx < 0 || xx >= X) continue;
                            int64_t q = (zz * Y + yy) * X + xx;
                            if (labels[q] == 4) { labels[q] = 3; hist_out[lev[q]]++; n_out++; n_excl--; }
                            if (executed && dz >= -1 && dz <= 1 && dy >= -1 && dy <= 1 && dx >= -1 && dx <= 1 &&
                                !(V.flag[q] & F_DIRTY)) { V.flag[q] |= F_DIRTY; push(&dirty, q); }
                        } } }
            }
        }
        /* canonical re-classification of everything within 1 of an executed flip */
        {
            uint8_t *nl = (uint8_t *)malloc(dirty.n ? dirty.n : 1);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < dirty.n; i++) {
                int64_t p = dirty.v[i]; int64_t x = p % X, y = (p / X) % Y, z = p / (X * Y);
                nl[i] = classify(&V, z, y, x);
            }
            for (int64_t i = 0; i < dirty.n; i++) labels[dirty.v[i]] = nl[i];
            free(nl);
        }
        /* integer statistics (VRG:232-255 as histogram deltas; VRG:113-116 sizes) + quirk potentials */
        for (int64_t i = 0; i < Rl.n; i++) {
            int64_t p = Rl.v[i];
            hist_in[lev[p]]--; hist_out[lev[p]]++; n_in--; n_out++;
            if (labels[p] == 3) quirk[1]++;
        }
        for (int64_t i = 0; i < Al.n; i++) {
            int64_t p = Al.v[i];
            if (V.flag[p] & F_AEX) {
                hist_in[lev[p]]++; hist_out[lev[p]]--; n_in++; n_out--;
                if (labels[p] == 0) quirk[0]++;
            } else {
                quirk[3]++;
                int64_t x = p % X, y = (p / X) % Y, z = p / (X * Y); int hit = 0;
                for (int64_t dz = -1; dz <= 1 && !hit; dz++) { int64_t zz = z + dz; if (zz < 0 || zz >= Z) continue;
                    for (int64_t dy = -1; dy <= 1 && !hit; dy++) { int64_t yy = y + dy; if (yy < 0 || yy >= Y) continue;
                        for (int64_t dx = -1; dx <= 1; dx++) { int64_t xx = x + dx; if (xx < 0 || xx >= X) continue;
                            if (V.flag[(zz * Y + yy) * X + xx] & F_AEX) { hit = 1; break; }
                        } } }
                quirk[2] += hit;
            }
        }
        for (int64_t i = 0; i < Rl.n; i++) V.flag[Rl.v[i]] = 0;
        for (int64_t i = 0; i < Al.n; i++) V.flag[Al.v[i]] = 0;
        for (int64_t i = 0; i < dirty.n; i++) V.flag[dirty.v[i]] = 0;
        if (trace && nt < trace_cap) { trace[3 * nt] = n_flips; trace[3 * nt + 1] = n_in; trace[3 * nt + 2] = n_out; }
        nt++;
        iter++;
    }
    if (iterations) *iterations = iter;
    if (exit_code) *exit_code = ex;
    if (n_trace) *n_trace = nt < trace_cap ? nt : trace_cap;
    free(lev); free(V.flag); free(hist_in); free(hist_out); free(pin); free(pout); free(dbit);
    free(levels); free(dirty.v); free(Rl.v); free(Al.v);
    return rc;
}

/* Position-sensitive 64-bit hash of a label volume: sum over voxels of mix64(((base + i) << 3) | label) mod 2^64 -- the
 * same number arterynetwork_b200's vrg_labels_hash computes on the device (checker side of the multi-GPU parity test:
 * the hashes of z-slabs add up to the hash of the whole volume). */
uint64_t vrg_oracle_hash_labels(const uint8_t *lab, int64_t n, int64_t base, int nthreads) {
    uint64_t acc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < n; i++) acc += mix64(((uint64_t)(base + i) << 3) | (uint64_t)lab[i]);
    return acc;
}
