/* TEST INFRASTRUCTURE ONLY -- plain-C restatements of the two third-party array operations on either side of the
 * variational-region-growing path.  The product path never links or loads
 * this file.
 *
 *  edt_sq_oracle    squared Euclidean distance of every non-zero voxel to the nearest zero voxel.  The reference calls
 *                   scipy.ndimage.distance_transform_edt(mask) with default arguments
 *                   (Code/manualCorrectionGUI.py:248, Code/generateVesselVolume.py:183); SciPy is not part of
 *                   /root/reference, so this restates the published definition (exact EDT, unit spacing) as three
 *                   separable min-plus passes  out(u) = min_i (u - i)^2 + in(i)  written as the plain double loop, in
 *                   64-bit integers.  Deliberately NOT the lower-envelope algorithm the CUDA path uses.
 *  label26_oracle   connected components under 26-connectivity, numbered 1.. in raster (C) order of each component's
 *                   first voxel.  The reference calls skimage.measure.label(volume, return_num=True, connectivity=3)
 *                   (Code/generateVesselVolume.py:126; scikit-image is not installed here); restated as a breadth-first
 *                   flood fill.
 *
 * Parity status: PINNED to SciPy 1.x outputs computed in the build container (tests/golden/mask/*.npz, generator
 * tests/golden/mask/make_golden_mask.py): distance_transform_edt for the first, ndimage.label with a full 3x3x3
 * structure for the second (same components and, on every fixture, the same numbering).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o _build/libmask_oracle.so mask_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_INF ((int64_t)1 << 40)

/* one min-plus pass along an axis of length n with element stride `stride`, for `nlines` lines whose first elements are
 * listed implicitly: line l starts at (l / inner) * outer_stride + (l % inner) * inner_stride */
static void minplus_axis(const int64_t *in, int64_t *out, int64_t n, int64_t stride, int64_t nlines, int64_t inner,
                         int64_t inner_stride, int64_t outer_stride) {
#pragma omp parallel for schedule(static)
    for (int64_t l = 0; l < nlines; ++l) {
        const int64_t base = (l / inner) * outer_stride + (l % inner) * inner_stride;
        for (int64_t u = 0; u < n; ++u) {
            int64_t best = ORACLE_INF;
            for (int64_t i = 0; i < n; ++i) {
                const int64_t v = in[base + i * stride];
                if (v >= ORACLE_INF) continue;
                const int64_t d = (u - i) * (u - i) + v;
                if (d < best) best = d;
            }
            out[base + u * stride] = best;
        }
    }
}

/* mask: uint8 (Z,Y,X) C order, non-zero = foreground; out: int64 squared distances (ORACLE_INF if the mask has no zero).
 * returns 0, or -5 when out of memory */
int edt_sq_oracle(const uint8_t *mask, int64_t Z, int64_t Y, int64_t X, int64_t *out) {
    const int64_t n = Z * Y * X;
    int64_t *a = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    if (!a) return -5;
    for (int64_t p = 0; p < n; ++p) out[p] = mask[p] ? ORACLE_INF : 0;
    minplus_axis(out, a, X, 1, Z * Y, 1, 0, X);          /* along x: line l = row l */
    minplus_axis(a, out, Y, X, Z * X, X, 1, X * Y);      /* along y: lines (z, x) */
    minplus_axis(out, a, Z, X * Y, Y * X, Y * X, 1, 0);  /* along z: lines (y, x) */
    memcpy(out, a, (size_t)n * sizeof(int64_t));
    free(a);
    return 0;
}

/* binary: uint8, non-zero = foreground; labels: int32 out (0 = background, components 1..K in raster order of their
 * first voxel); sizes_out (optional): int64[cap] voxel counts of components 1..min(K,cap) at index k-1.
 * returns K, or -5 when out of memory */
int64_t label26_oracle(const uint8_t *binary, int64_t Z, int64_t Y, int64_t X, int32_t *labels, int64_t *sizes_out, int64_t cap) {
    const int64_t n = Z * Y * X;
    int64_t *queue = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    if (!queue) return -5;
    memset(labels, 0, (size_t)n * sizeof(int32_t));
    int64_t K = 0;
    for (int64_t p0 = 0; p0 < n; ++p0) {
        if (!binary[p0] || labels[p0]) continue;
        ++K;
        int64_t head = 0, tail = 0, size = 0;
        queue[tail++] = p0;
        labels[p0] = (int32_t)K;
        while (head < tail) {
            const int64_t p = queue[head++];
            ++size;
            const int64_t z = p / (Y * X), y = (p / X) % Y, x = p % X;
            for (int64_t dz = -1; dz <= 1; ++dz)
                for (int64_t dy = -1; dy <= 1; ++dy)
                    for (int64_t dx = -1; dx <= 1; ++dx) {
                        const int64_t zz = z + dz, yy = y + dy, xx = x + dx;
                        if (zz < 0 || zz >= Z || yy < 0 || yy >= Y || xx < 0 || xx >= X) continue;
                        const int64_t q = (zz * Y + yy) * X + xx;
                        if (binary[q] && !labels[q]) { labels[q] = (int32_t)K; queue[tail++] = q; }
                    }
        }
        if (sizes_out && K <= cap) sizes_out[K - 1] = size;
    }
    free(queue);
    return K;
}
