/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the eikonal distance map behind the reference's navigation fields
 * (core/steering/quickest_path.py:54-117).  The reference calls skfmm.distance (scikit-fmm, third party, version unpinned in
 * requirements.txt, NOT installed here): "parity unpinned" against skfmm itself.  What is restated is the published algorithm
 * skfmm implements -- Sethian's Fast Marching Method on the signed level set phi (+1 inside the target raster, -1 elsewhere,
 * masked cells excluded) -- in its FIRST-ORDER upwind form:
 *   initialisation  a cell with an axis neighbour of opposite sign is frozen at the distance to the interpolated zero
 *                   crossing, d_axis = h |phi| / (|phi| + |phi_nb|) = h / 2, combined over the axes as
 *                   1 / sqrt(sum 1 / d_axis^2)   (skfmm's initialize_frozen);
 *   update          T = min(a, b) + h                                if |a - b| >= h
 *                   T = (a + b + sqrt(2 h^2 - (a - b)^2)) / 2         otherwise,   a / b = smaller x / y neighbour;
 *   order           accepted in increasing T (binary heap).
 * (skfmm defaults to its second-order variant; both converge to the true distance, first order with error O(h).  The GPU
 * solver iterates the same first-order update to its fixed point, which is this function's output.)
 * Output: signed distance, positive inside the target raster, negative outside, NaN in masked cells. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { double t; int64_t c; } heap_item;
typedef struct { heap_item *a; int64_t n, cap; } heap_t;

static void heap_push(heap_t *h, double t, int64_t c) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 1024; h->a = (heap_item *)realloc(h->a, (size_t)h->cap * sizeof(heap_item)); }
    int64_t i = h->n++;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (h->a[p].t <= t) break;
        h->a[i] = h->a[p]; i = p;
    }
    h->a[i].t = t; h->a[i].c = c;
}
static heap_item heap_pop(heap_t *h) {
    heap_item top = h->a[0], last = h->a[--h->n];
    int64_t i = 0;
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        double best = last.t;
        if (l < h->n && h->a[l].t < best) { m = l; best = h->a[l].t; }
        if (r < h->n && h->a[r].t < best) { m = r; }
        if (m == i) break;
        h->a[i] = h->a[m]; i = m;
    }
    h->a[i] = last;
    return top;
}

double oracle_eikonal_update(double a, double b, double h) {
    if (isinf(a) && isinf(b)) return INFINITY;
    if (fabs(a - b) >= h || isinf(a) || isinf(b)) return fmin(a, b) + h;
    return (a + b + sqrt(2.0 * h * h - (a - b) * (a - b))) / 2.0;
}

/* target, mask: (ny, nx) uint8 rasters; out: (ny, nx) doubles */
int oracle_distance_map(const uint8_t *target, const uint8_t *mask, int64_t ny, int64_t nx, double h, double *out) {
    const int64_t n = ny * nx;
    double *T = (double *)malloc((size_t)n * sizeof(double));
    uint8_t *state = (uint8_t *)calloc((size_t)n, 1);     /* 0 far, 1 accepted, 2 masked */
    heap_t heap = {0, 0, 0};
    if (!T || !state) return 1;
    for (int64_t c = 0; c < n; ++c) { T[c] = INFINITY; if (mask && mask[c]) state[c] = 2; }
    /* frozen cells next to the zero level set */
    for (int64_t y = 0; y < ny; ++y)
        for (int64_t x = 0; x < nx; ++x) {
            const int64_t c = y * nx + x;
            if (state[c] == 2) continue;
            const int s = target[c] ? 1 : -1;
            double inv = 0.0;
            int cross_x = 0, cross_y = 0;
            if (x > 0 && state[c - 1] != 2 && (target[c - 1] ? 1 : -1) != s) cross_x = 1;
            if (x + 1 < nx && state[c + 1] != 2 && (target[c + 1] ? 1 : -1) != s) cross_x = 1;
            if (y > 0 && state[c - nx] != 2 && (target[c - nx] ? 1 : -1) != s) cross_y = 1;
            if (y + 1 < ny && state[c + nx] != 2 && (target[c + nx] ? 1 : -1) != s) cross_y = 1;
            const double d = h / 2.0;
            if (cross_x) inv += 1.0 / (d * d);
            if (cross_y) inv += 1.0 / (d * d);
            if (cross_x || cross_y) { T[c] = 1.0 / sqrt(inv); state[c] = 1; }
        }
    /* neighbours of the frozen band */
    const int64_t dx[4] = {-1, 1, 0, 0}, dy[4] = {0, 0, -1, 1};
#define NB(xx, yy) (((xx) >= 0 && (xx) < nx && (yy) >= 0 && (yy) < ny && state[(yy) * nx + (xx)] == 1) ? T[(yy) * nx + (xx)] : INFINITY)
    for (int64_t y = 0; y < ny; ++y)
        for (int64_t x = 0; x < nx; ++x) {
            const int64_t c = y * nx + x;
            if (state[c] != 0) continue;
            const double a = fmin(NB(x - 1, y), NB(x + 1, y)), b = fmin(NB(x, y - 1), NB(x, y + 1));
            const double t = oracle_eikonal_update(a, b, h);
            if (t < T[c]) { T[c] = t; heap_push(&heap, t, c); }
        }
    while (heap.n) {
        const heap_item it = heap_pop(&heap);
        const int64_t c = it.c;
        if (state[c] != 0 || it.t != T[c]) continue;     /* stale entry */
        state[c] = 1;
        const int64_t x = c % nx, y = c / nx;
        for (int k = 0; k < 4; ++k) {
            const int64_t x2 = x + dx[k], y2 = y + dy[k];
            if (x2 < 0 || x2 >= nx || y2 < 0 || y2 >= ny) continue;
            const int64_t c2 = y2 * nx + x2;
            if (state[c2] != 0) continue;
            const double a = fmin(NB(x2 - 1, y2), NB(x2 + 1, y2)), b = fmin(NB(x2, y2 - 1), NB(x2, y2 + 1));
            const double t = oracle_eikonal_update(a, b, h);
            if (t < T[c2]) { T[c2] = t; heap_push(&heap, t, c2); }
        }
    }
#undef NB
    for (int64_t c = 0; c < n; ++c)
        out[c] = state[c] == 2 ? NAN : (target[c] ? T[c] : -T[c]);
    free(T); free(state); free(heap.a);
    return 0;
}
