/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's variational region growing WITH its list order
 * ("strict" mode): /root/reference/Code/variationalRegionGrowing.py, cited as VRG:line.
 *
 * oracle/vrg_oracle.{py,c} restate the band state machine as order-free set rules.  This file keeps what those leave
 * out: the reference walks `flipedPoints` in the order of its band lists (VRG:163), tests every point against the
 * labels as they are AT ITS TURN (VRG:169,198), relabels a removed / added voxel 2 / 1 without looking at its
 * neighbours (VRG:174,202: stale band labels, Q2), selects the Parzen corrections by the labels AFTER the whole call
 * (VRG:232-233: dropped corrections, Q3) and keeps per-voxel running sums that drift from the true ones (VRG:236-247).
 * The three Python lists (VRG:126,157-158) hold every voxel at most once (list membership <=> label, see below), so
 * "list" here is an array with tombstones: append at the end, remove = blank the slot, iterate in slot order.
 *
 * Parzen sums go through the intensity levels (a voxel's kernel terms depend on its level only): the sum over a set of
 * voxels is  sum_c count[c] * A * exp(-0.5 * H * (level_c - level_v)^2).  The reference adds the same terms voxel by
 * voxel (np.sum, pairwise); the two agree to rounding (1e-13 relative), and the tests state that tolerance.
 *
 * Pinned by tests/test_strict_oracle.py against fixtures written by tests/golden/make_golden_strict.py from the
 * unmodified reference on inputs where its result DIFFERS from the order-free restatement.  The product path never
 * links or loads this file.
 *
 * Build: gcc -O2 -shared -fPIC -o _build/libvrg_strict_oracle.so vrg_strict_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VRG_A 0.3989422804014327 /* (2*pi)**-0.5, VRG:7 */

enum { EXIT_RUNNING = -1, EXIT_CONVERGED = 0, EXIT_MAX_SEGMENT = 2, EXIT_MAX_ITER = 3 };

typedef struct { int64_t *v; int64_t n, cap, live; int64_t *pos; } olist;

static void ol_init(olist *l, int64_t N) {
    l->cap = 1024; l->n = 0; l->live = 0;
    l->v = (int64_t *)malloc(l->cap * sizeof(int64_t));
    l->pos = (int64_t *)malloc(N * sizeof(int64_t));
    for (int64_t i = 0; i < N; ++i) l->pos[i] = -1;
}
static void ol_free(olist *l) { free(l->v); free(l->pos); }
static void ol_append(olist *l, int64_t x) { /* list.append */
    if (l->n == l->cap) { l->cap *= 2; l->v = (int64_t *)realloc(l->v, l->cap * sizeof(int64_t)); }
    l->pos[x] = l->n; l->v[l->n++] = x; l->live++;
}
static void ol_remove(olist *l, int64_t x) { /* list.remove: the element is unique */
    if (l->pos[x] < 0) return;
    l->v[l->pos[x]] = -1; l->pos[x] = -1; l->live--;
}
static void ol_compact(olist *l) {
    int64_t w = 0;
    for (int64_t i = 0; i < l->n; ++i) if (l->v[i] >= 0) { l->pos[l->v[i]] = w; l->v[w++] = l->v[i]; }
    l->n = w;
}

typedef struct {
    int64_t Z, Y, X, N, L;
    double H;
    const int32_t *lev; /* level index per voxel (caller's buffer) */
    double *levels;
    double *K;          /* L x L kernel values */
    uint8_t *vm, *sm;   /* valueMap (VRG:21), segmentedMap */
    double *pin, *pout; /* innerProb, outerProb (VRG:132-133), unnormalised */
    olist inner, outer, seg;
    int64_t max_seg, iter_max, iter_num, exit_code;
    int64_t n_in, n_out;
    int64_t *trace; int64_t n_trace;
    int64_t skipped, dropped; /* flipped points that were in neither band at their turn; flipped points whose label after the call is not 1 or 2 (Q3) */
    double min_margin;
    int64_t *newl; int64_t n_new, cap_new;
    int64_t *hin, *hout, *ha, *hb, *hc;
    double *tin, *tout, *ta, *tb, *tc;
} strict_t;

/* get_neighbours (VRG:263-282): offsets in C order of np.indices((3,3,3)), centre excluded, out-of-bounds dropped */
static int neighbours(const strict_t *s, int64_t v, int64_t *out) {
    const int64_t X = s->X, Y = s->Y, Z = s->Z;
    const int64_t x = v % X, y = (v / X) % Y, z = v / (X * Y);
    int n = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if (!dz && !dy && !dx) continue;
                const int64_t zz = z + dz, yy = y + dy, xx = x + dx;
                if (zz < 0 || zz >= Z || yy < 0 || yy >= Y || xx < 0 || xx >= X) continue;
                out[n++] = (zz * Y + yy) * X + xx;
            }
    return n;
}

static void count_regions(strict_t *s) { /* VRG:47-50,112-115 */
    int64_t a = 0, b = 0;
    for (int64_t i = 0; i < s->N; ++i) { a += s->vm[i] <= 1; b += s->vm[i] == 2 || s->vm[i] == 3; }
    s->n_in = a; s->n_out = b;
}

static void table(const strict_t *s, const int64_t *hist, double *out) { /* sum_c hist[c] * K[c][b] in level order */
    for (int64_t b = 0; b < s->L; ++b) {
        double acc = 0.0;
        for (int64_t c = 0; c < s->L; ++c) if (hist[c]) acc += (double)hist[c] * s->K[c * s->L + b];
        out[b] = acc;
    }
}

static void region_hists(strict_t *s) { /* innerValues / outerValues, VRG:148-149,249-250 */
    memset(s->hin, 0, s->L * sizeof(int64_t)); memset(s->hout, 0, s->L * sizeof(int64_t));
    for (int64_t i = 0; i < s->N; ++i) {
        if (s->vm[i] <= 1) s->hin[s->lev[i]]++;
        else if (s->vm[i] <= 3) s->hout[s->lev[i]]++;
    }
}

void *vrg_strict_create(const int32_t *lev, const double *levels, int64_t L, const uint8_t *vm, int64_t Z, int64_t Y,
                        int64_t X, double H, int64_t max_seg, int64_t iter_max) {
    strict_t *s = (strict_t *)calloc(1, sizeof(strict_t));
    s->Z = Z; s->Y = Y; s->X = X; s->N = Z * Y * X; s->L = L; s->H = H; s->lev = lev;
    s->max_seg = max_seg; s->iter_max = iter_max; s->iter_num = 1; s->exit_code = EXIT_RUNNING;
    s->min_margin = INFINITY;
    const int64_t N = s->N;
    s->levels = (double *)malloc(L * sizeof(double)); memcpy(s->levels, levels, L * sizeof(double));
    s->K = (double *)malloc(L * L * sizeof(double));
    for (int64_t c = 0; c < L; ++c)
        for (int64_t b = 0; b < L; ++b) { const double d = levels[c] - levels[b]; s->K[c * L + b] = VRG_A * exp(-0.5 * H * d * d); }
    s->vm = (uint8_t *)malloc(N); memcpy(s->vm, vm, N);
    s->sm = (uint8_t *)calloc(N, 1);
    s->pin = (double *)calloc(N, sizeof(double)); s->pout = (double *)calloc(N, sizeof(double));
    ol_init(&s->inner, N); ol_init(&s->outer, N); ol_init(&s->seg, N);
    s->trace = (int64_t *)calloc(3 * (iter_max + 2), sizeof(int64_t));
    s->cap_new = 1024; s->newl = (int64_t *)malloc(s->cap_new * sizeof(int64_t));
    s->hin = (int64_t *)calloc(5 * L, sizeof(int64_t)); s->hout = s->hin + L; s->ha = s->hout + L; s->hb = s->ha + L; s->hc = s->hb + L;
    s->tin = (double *)calloc(5 * L, sizeof(double)); s->tout = s->tin + L; s->ta = s->tout + L; s->tb = s->ta + L; s->tc = s->tb + L;
    /* VRG:44-46: segmented = where(valueMap == 0) in C order */
    for (int64_t i = 0; i < N; ++i) if (s->vm[i] == 0) { ol_append(&s->seg, i); s->sm[i] = 1; }
    /* init branch, VRG:129-145 */
    int64_t nb[26];
    for (int64_t k = 0; k < s->seg.n; ++k) {
        const int64_t p = s->seg.v[k];
        const int n = neighbours(s, p, nb);
        for (int j = 0; j < n; ++j) if (s->vm[nb[j]] == 4) s->vm[nb[j]] = 3; /* VRG:137 */
        for (int j = 0; j < n; ++j) {
            if (s->sm[nb[j]] == 0) {
                if (s->vm[p] != 1) { ol_append(&s->inner, p); s->vm[p] = 1; }
                if (s->vm[nb[j]] != 2) { ol_append(&s->outer, nb[j]); s->vm[nb[j]] = 2; }
            }
        }
    }
    /* VRG:148-155: full sums at every band voxel */
    region_hists(s);
    table(s, s->hin, s->tin); table(s, s->hout, s->tout);
    for (int64_t k = 0; k < s->inner.n; ++k) { const int64_t v = s->inner.v[k]; s->pin[v] = s->tin[lev[v]]; s->pout[v] = s->tout[lev[v]]; }
    for (int64_t k = 0; k < s->outer.n; ++k) { const int64_t v = s->outer.v[k]; s->pin[v] = s->tin[lev[v]]; s->pout[v] = s->tout[lev[v]]; }
    count_regions(s);
    s->trace[0] = -1; s->trace[1] = s->n_in; s->trace[2] = s->n_out; s->n_trace = 1;
    return s;
}

void vrg_strict_destroy(void *h) {
    strict_t *s = (strict_t *)h;
    if (!s) return;
    free(s->levels); free(s->K); free(s->vm); free(s->sm); free(s->pin); free(s->pout);
    ol_free(&s->inner); ol_free(&s->outer); ol_free(&s->seg);
    free(s->trace); free(s->newl); free(s->hin); free(s->tin); free(s);
}

static void new_push(strict_t *s, int64_t v) {
    if (s->n_new == s->cap_new) { s->cap_new *= 2; s->newl = (int64_t *)realloc(s->newl, s->cap_new * sizeof(int64_t)); }
    s->newl[s->n_new++] = v;
}
static void absorb(strict_t *s, const int64_t *nb, int n) { /* 4 -> 3 and the voxel joins includedPoints, VRG:165-168,177-179,207-208 */
    for (int j = 0; j < n; ++j) if (s->vm[nb[j]] == 4) { s->vm[nb[j]] = 3; s->hc[s->lev[nb[j]]]++; }
}
static int any_label(const strict_t *s, const int64_t *nb, int n, uint8_t lab) {
    for (int j = 0; j < n; ++j) if (s->vm[nb[j]] == lab) return 1;
    return 0;
}

/* one pass of the while loop, VRG:58-117; returns the exit code (EXIT_RUNNING = the update was applied) */
int64_t vrg_strict_step(void *h) {
    strict_t *s = (strict_t *)h;
    if (s->exit_code != EXIT_RUNNING) return s->exit_code;
    if (s->iter_num > s->iter_max) { s->exit_code = EXIT_MAX_ITER; return s->exit_code; } /* VRG:118-121 */
    ol_compact(&s->inner); ol_compact(&s->outer);
    /* VRG:79-88: allBnd = innerBnd ++ outerBnd; flipped = band voxels whose side disagrees with the decision (ties inside) */
    const int64_t nband = s->inner.n + s->outer.n;
    int64_t *fl = (int64_t *)malloc((nband + 1) * sizeof(int64_t));
    int64_t nf = 0;
    for (int64_t k = 0; k < nband; ++k) {
        const int64_t v = k < s->inner.n ? s->inner.v[k] : s->outer.v[k - s->inner.n];
        const double a = s->pin[v] / (double)s->n_in, b = s->pout[v] / (double)s->n_out;
        const double m = fabs(a - b) / fmax(a, b);
        if (m < s->min_margin) s->min_margin = m;
        if ((s->sm[v] != 0) != (a >= b)) fl[nf++] = v;
    }
    if (nf == 0) { free(fl); s->exit_code = EXIT_CONVERGED; return s->exit_code; }             /* VRG:91 */
    if (s->seg.live >= s->max_seg) { free(fl); s->exit_code = EXIT_MAX_SEGMENT; return s->exit_code; } /* VRG:101 */
    /* update(), VRG:156-230 */
    memset(s->hc, 0, s->L * sizeof(int64_t));
    s->n_new = 0;
    int64_t nb[26], nb2[26];
    for (int64_t k = 0; k < nf; ++k) {
        const int64_t p = fl[k];
        const int n = neighbours(s, p, nb);
        absorb(s, nb, n);                                  /* VRG:165-168 */
        if (s->vm[p] == 1) {                               /* VRG:169: inner band voxel leaves */
            ol_remove(&s->inner, p); ol_remove(&s->seg, p);
            s->sm[p] = 0; s->vm[p] = 2; ol_append(&s->outer, p);
            for (int j = 0; j < n; ++j) {
                const int64_t q = nb[j];
                const int n2 = neighbours(s, q, nb2);
                absorb(s, nb2, n2);                        /* VRG:177-179 */
                if (s->vm[q] == 2) {                       /* VRG:183-190 */
                    if (!any_label(s, nb2, n2, 1)) { s->vm[q] = 3; ol_remove(&s->outer, q); s->pin[q] = 0.0; s->pout[q] = 0.0; }
                } else if (s->vm[q] == 0) {                /* VRG:193-196: inside -> inner band */
                    s->vm[q] = 1; new_push(s, q); ol_append(&s->inner, q);
                }
            }
        } else if (s->vm[p] == 2) {                        /* VRG:198: outer band voxel enters */
            ol_remove(&s->outer, p); ol_append(&s->seg, p);
            s->sm[p] = 1; s->vm[p] = 1; ol_append(&s->inner, p);
            for (int j = 0; j < n; ++j) {
                const int64_t q = nb[j];
                const int n2 = neighbours(s, q, nb2);
                absorb(s, nb2, n2);                        /* VRG:205-208 */
                if (s->vm[q] == 3) {                       /* VRG:209-212: outside -> outer band */
                    s->vm[q] = 2; new_push(s, q); ol_append(&s->outer, q);
                } else if (s->vm[q] == 1) {                /* VRG:218-227 */
                    if (!any_label(s, nb2, n2, 2)) { s->vm[q] = 0; ol_remove(&s->inner, q); s->pin[q] = 0.0; s->pout[q] = 0.0; }
                }
            }
        } else {
            s->skipped++;                                  /* in neither band any more at its turn: nothing happens */
        }
    }
    /* VRG:232-247: corrections, selected by the labels after the call */
    memset(s->ha, 0, s->L * sizeof(int64_t)); memset(s->hb, 0, s->L * sizeof(int64_t));
    for (int64_t k = 0; k < nf; ++k) {
        const int64_t p = fl[k];
        if (s->vm[p] == 1) s->ha[s->lev[p]]++;
        else if (s->vm[p] == 2) s->hb[s->lev[p]]++;
        else s->dropped++;
    }
    table(s, s->ha, s->ta); table(s, s->hb, s->tb); table(s, s->hc, s->tc);
    for (int pass = 0; pass < 2; ++pass) {
        const olist *l = pass ? &s->outer : &s->inner;
        for (int64_t k = 0; k < l->n; ++k) {
            const int64_t v = l->v[k];
            if (v < 0) continue;
            const int32_t b = s->lev[v];
            s->pin[v] += s->ta[b]; s->pin[v] -= s->tb[b];
            s->pout[v] -= s->ta[b]; s->pout[v] += s->tb[b]; s->pout[v] += s->tc[b];
        }
    }
    /* VRG:249-255: voxels that entered a band get both sums against the whole regions (also those that left again) */
    region_hists(s);
    table(s, s->hin, s->tin); table(s, s->hout, s->tout);
    for (int64_t k = 0; k < s->n_new; ++k) { const int64_t v = s->newl[k]; s->pin[v] = s->tin[s->lev[v]]; s->pout[v] = s->tout[s->lev[v]]; }
    count_regions(s);                                     /* VRG:112-115 */
    s->trace[3 * s->n_trace] = nf; s->trace[3 * s->n_trace + 1] = s->n_in; s->trace[3 * s->n_trace + 2] = s->n_out; s->n_trace++;
    s->iter_num++;
    free(fl);
    return EXIT_RUNNING;
}

/* getters ---------------------------------------------------------------------------------------------------------- */
int64_t vrg_strict_info(void *h, int64_t *out /* [10] */) {
    strict_t *s = (strict_t *)h;
    out[0] = s->iter_num; out[1] = s->exit_code; out[2] = s->n_in; out[3] = s->n_out; out[4] = s->inner.live;
    out[5] = s->outer.live; out[6] = s->seg.live; out[7] = s->n_trace; out[8] = s->skipped; out[9] = s->dropped;
    return 0;
}
double vrg_strict_min_margin(void *h) { return ((strict_t *)h)->min_margin; }
void vrg_strict_get(void *h, uint8_t *vm, uint8_t *sm, double *pin, double *pout, int64_t *trace) {
    strict_t *s = (strict_t *)h;
    if (vm) memcpy(vm, s->vm, s->N);
    if (sm) memcpy(sm, s->sm, s->N);
    if (pin) memcpy(pin, s->pin, s->N * sizeof(double));
    if (pout) memcpy(pout, s->pout, s->N * sizeof(double));
    if (trace) memcpy(trace, s->trace, 3 * s->n_trace * sizeof(int64_t));
}
/* which: 0 innerBndList, 1 outerBndList, 2 segmentedList -- flat voxel indices in list order */
int64_t vrg_strict_list(void *h, int which, int64_t *out) {
    strict_t *s = (strict_t *)h;
    const olist *l = which == 0 ? &s->inner : which == 1 ? &s->outer : &s->seg;
    int64_t w = 0;
    for (int64_t i = 0; i < l->n; ++i) if (l->v[i] >= 0) out[w++] = l->v[i];
    return w;
}
