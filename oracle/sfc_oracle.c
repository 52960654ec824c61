/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * Plain-C restatement of the reference's *float* curve pipeline
 * (src/curves/space_filling_curves.py): the recursive generators emit double
 * cell centres, a 2x2 rotation/mirror matrix built from cos()/sin() of
 * pi/2, pi, 2*pi is applied, then floor() and the in-domain filter of
 * embed_and_prune_sfc. It deliberately keeps the reference's algorithmic shape
 * (recursion + float transform) so that it is an independent check of the
 * product's per-index integer kernels (csrc/curve_index.h).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library.
 *
 * Pinned: tests/test_oracle_golden.py checks it against hashes and vectors
 * generated from the live reference (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { double x, y; } pt_t;
typedef struct { pt_t *p; int64_t n, cap; } plist_t;

static void push(plist_t *l, double x, double y) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 1024; l->p = (pt_t *)realloc(l->p, (size_t)l->cap * sizeof(pt_t)); }
    l->p[l->n].x = x; l->p[l->n].y = y; l->n++;
}

/* reference space_filling_curves.py:181-193 (hilbert inner recursion; also :219-231 in moore_curve) */
static void hilbert_rec(plist_t *l, double x0, double y0, double xi, double xj, double yi, double yj, int n) {
    if (n <= 0) {
        push(l, x0 + (xi + yi) / 2, y0 + (xj + yj) / 2);
    } else {
        hilbert_rec(l, x0, y0, yi / 2, yj / 2, xi / 2, xj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2, y0 + xj / 2, xi / 2, xj / 2, yi / 2, yj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2 + yi / 2, y0 + xj / 2 + yj / 2, xi / 2, xj / 2, yi / 2, yj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2 + yi, y0 + xj / 2 + yj, -yi / 2, -yj / 2, -xi / 2, -xj / 2, n - 1);
    }
}

/* reference :233-245 (moore outer level) */
static void moore_top(plist_t *l, double x0, double y0, double xi, double xj, double yi, double yj, int n) {
    if (n <= 0) {
        push(l, x0 + (xi + yi) / 2, y0 + (xj + yj) / 2);
    } else {
        hilbert_rec(l, x0 + xi / 2, y0 + xj / 2, -xi / 2, xj / 2, yi / 2, yj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2 + yi / 2, y0 + xj / 2 + yj / 2, -xi / 2, xj / 2, yi / 2, yj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2 + yi, y0 + xj / 2 + yj, xi / 2, xj / 2, yi / 2, -yj / 2, n - 1);
        hilbert_rec(l, x0 + xi / 2 + yi / 2, y0 + xj / 2 + yj / 2, xi / 2, xj / 2, yi / 2, -yj / 2, n - 1);
    }
}

/* reference :147-156 */
static void z_rec(plist_t *l, double x0, double y0, double w, int n) {
    if (n == 0) {
        push(l, x0 + w / 2, y0 + w / 2);
    } else {
        double half = w / 2;
        z_rec(l, x0 + half, y0, half, n - 1);
        z_rec(l, x0, y0, half, n - 1);
        z_rec(l, x0 + half, y0 + half, half, n - 1);
        z_rec(l, x0, y0 + half, half, n - 1);
    }
}

/* reference :95-108 */
static const int PEANO_PAT[4][9][3] = {
    {{0,0,0},{1,0,1},{2,0,0},{2,1,1},{1,1,0},{0,1,1},{0,2,0},{1,2,1},{2,2,0}},
    {{2,0,1},{1,0,0},{0,0,1},{0,1,0},{1,1,1},{2,1,0},{2,2,1},{1,2,0},{0,2,1}},
    {{0,2,2},{1,2,3},{2,2,2},{2,1,3},{1,1,2},{0,1,3},{0,0,2},{1,0,3},{2,0,2}},
    {{2,2,3},{1,2,2},{0,2,3},{0,1,2},{1,1,3},{2,1,2},{2,0,3},{1,0,2},{0,0,3}},
};

/* reference :86-123: returns a list; sub-lists at idx%3==1 are reversed */
static void peano_rec(plist_t *l, double x, double y, double size, int order, int pattern) {
    if (order == 0) { push(l, x + size / 2, y + size / 2); return; }
    size /= 3;
    for (int idx = 0; idx < 9; ++idx) {
        int dx = PEANO_PAT[pattern][idx][0], dy = PEANO_PAT[pattern][idx][1], nxt = PEANO_PAT[pattern][idx][2];
        int64_t start = l->n;
        peano_rec(l, x + dx * size, y + dy * size, size, order - 1, nxt);
        if (idx % 3 == 1) {
            int64_t a = start, b = l->n - 1;
            while (a < b) { pt_t t = l->p[a]; l->p[a] = l->p[b]; l->p[b] = t; ++a; --b; }
        }
    }
}

static void apply_fin(plist_t *l, const double fin[2][2]) {
    for (int64_t k = 0; k < l->n; ++k) {
        double x = l->p[k].x, y = l->p[k].y;
        l->p[k].x = fin[0][0] * x + fin[0][1] * y;
        l->p[k].y = fin[1][0] * x + fin[1][1] * y;
    }
}

static void matmul2(const double a[2][2], const double b[2][2], double c[2][2]) {
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) c[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j];
}

enum { SFC_HILBERT = 0, SFC_Z = 1, SFC_PEANO = 2, SFC_MOORE = 3, SFC_RASTER = 4 };

/* sfc(order, size) of the reference for curve id; caller frees l->p */
static int gen_curve(int curve, int order, double size, plist_t *l) {
    memset(l, 0, sizeof(*l));
    double rot[2][2], fin[2][2];
    if (curve == SFC_HILBERT) {
        hilbert_rec(l, 0, 0, size, 0, 0, size, order);
        double deg = M_PI / 2; /* :196-202 */
        rot[0][0] = cos(deg); rot[0][1] = -sin(deg); rot[1][0] = sin(deg); rot[1][1] = cos(deg);
        const double mir[2][2] = {{-1, 0}, {0, 1}};
        matmul2(mir, rot, fin); apply_fin(l, fin);
    } else if (curve == SFC_Z) {
        z_rec(l, 0, 0, size, order);
        double deg = M_PI; /* :159-165 */
        rot[0][0] = cos(deg); rot[0][1] = -sin(deg); rot[1][0] = sin(deg); rot[1][1] = cos(deg);
        const double mir[2][2] = {{-1, 0}, {0, -1}};
        matmul2(mir, rot, fin); apply_fin(l, fin);
    } else if (curve == SFC_PEANO) {
        peano_rec(l, 0, 0, size, order, 0);
        double deg = M_PI / 2; /* :125-131 */
        rot[0][0] = cos(deg); rot[0][1] = -sin(deg); rot[1][0] = sin(deg); rot[1][1] = cos(deg);
        const double mir[2][2] = {{-1, 0}, {0, 1}};
        matmul2(mir, rot, fin); apply_fin(l, fin);
    } else if (curve == SFC_MOORE) {
        moore_top(l, 0, 0, size, 0, 0, size, order);
        double deg = M_PI * 2; /* :248-251 */
        rot[0][0] = cos(deg); rot[0][1] = -sin(deg); rot[1][0] = sin(deg); rot[1][1] = cos(deg);
        apply_fin(l, rot);
    } else if (curve == SFC_RASTER) { /* :254-271 */
        int64_t n = 1ll << order; double cell = size / (double)n;
        for (int64_t y = 0; y < n; ++y) for (int64_t x = 0; x < n; ++x) push(l, (x + 0.5) * cell, (y + 0.5) * cell);
    } else return -1;
    return 0;
}

/* reference grid_size :458-468 */
int64_t sfc_oracle_grid_size(int curve, int order) {
    int64_t g = 1; int base = (curve == SFC_PEANO) ? 3 : 2;
    for (int k = 0; k < order; ++k) g *= base;
    return g;
}

/* float centres sfc(order,size): out_xy has 2*P*P doubles. returns count or <0 */
int64_t sfc_oracle_curve_points(int curve, int order, double size, double *out_xy, int64_t cap) {
    plist_t l; if (gen_curve(curve, order, size, &l)) return -1;
    if (l.n > cap) { free(l.p); return -2; }
    for (int64_t k = 0; k < l.n; ++k) { out_xy[2 * k] = l.p[k].x; out_xy[2 * k + 1] = l.p[k].y; }
    int64_t n = l.n; free(l.p); return n;
}

/* reference embed_and_prune_sfc :471-491. out_ij: up to width*height (i,j) int64 pairs. returns count */
int64_t sfc_oracle_embed_and_prune(int curve, int width, int height, int64_t *out_ij, int64_t cap_pairs) {
    if (curve == SFC_RASTER) return -1; /* grid_size raises for raster_curve in the reference */
    int order = 0; int m = width > height ? width : height;
    while (sfc_oracle_grid_size(curve, order) < m) order++;
    int64_t P = sfc_oracle_grid_size(curve, order);
    plist_t l; if (gen_curve(curve, order, (double)P, &l)) return -1;
    int64_t cnt = 0;
    for (int64_t k = 0; k < l.n; ++k) {
        int64_t i = (int64_t)floor(l.p[k].x), j = (int64_t)floor(l.p[k].y);
        if (0 <= i && i < width && 0 <= j && j < height) {
            if (cnt >= cap_pairs) { free(l.p); return -2; }
            out_ij[2 * cnt] = i; out_ij[2 * cnt + 1] = j; cnt++;
        }
    }
    free(l.p); return cnt;
}

/* reference _2D/hilbert_embedding.py:30-78: unit-square Hilbert without the final transform,
 * int(x*grid), int(y*grid), flat = i*grid + j. grid must be a power of two (order=int(log2(grid))). */
int64_t sfc_oracle_hilbert2d_flat(int grid, int64_t *out_flat, int64_t cap) {
    int order = (int)(log2((double)grid));
    plist_t l; memset(&l, 0, sizeof(l));
    hilbert_rec(&l, 0, 0, 1.0, 0, 0, 1.0, order);
    if (l.n > cap) { free(l.p); return -2; }
    for (int64_t k = 0; k < l.n; ++k) {
        int64_t i = (int64_t)(l.p[k].x * grid), j = (int64_t)(l.p[k].y * grid);
        out_flat[k] = i * grid + j;
    }
    int64_t n = l.n; free(l.p); return n;
}
