// Per-index integer space-filling-curve descent: position d along the un-pruned
// curve on a P x P padded square -> cell (i, j) = (row, col).
//
// These replace the reference's recursive float generators
// (/root/reference/src/curves/space_filling_curves.py:74-251) with closed-form,
// data-parallel integer walks (one root-to-leaf path per index), so that one GPU
// thread can evaluate one curve position independently. Conventions follow the
// reference *after* its rotation/mirror + floor (embed_and_prune_sfc :482-489):
//   hilbert : vector recursion (:181-195) followed by a transpose (:196-202)
//   z       : quadrant order TR, TL, BR, BL (:153-156), identity transform
//   peano   : 3x3 serpentine pattern table (:95-108), middle child of each row
//             reversed (:117-119), followed by a transpose (:125-131)
//   moore   : first level = Moore child table (:239-245), deeper levels = Hilbert
//             table (:226-231), identity transform (:248-251)
// Usable from host (tests) and device (curves.cu).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SFC_HD __host__ __device__ __forceinline__
#else
#define SFC_HD inline
#endif

enum SfcCurveId : int { SFC_HILBERT = 0, SFC_Z = 1, SFC_PEANO = 2, SFC_MOORE = 3, SFC_RASTER = 4 };

// smallest order with base^order >= m (reference grid_size/embed_and_prune_sfc :458-481)
SFC_HD int sfc_order_for(int curve, int m, int64_t* P_out) {
  const int64_t base = (curve == SFC_PEANO) ? 3 : 2;
  int order = 0; int64_t P = 1;
  while (P < m) { P *= base; ++order; }
  *P_out = P;
  return order;
}

// Vector descent in doubled-integer coordinates shared by Hilbert and Moore.
// State (x0,y0) origin and (xi,xj),(yi,yj) axis vectors, all scaled by 2 so that the
// final half-steps stay integral: start (0,0,2P,0,0,2P); each level halves the vectors.
SFC_HD void sfc_vec_descent(int order, uint64_t d, bool moore_top, int64_t P, int* ox, int* oy) {
  int64_t x0 = 0, y0 = 0, xi = 2 * P, xj = 0, yi = 0, yj = 2 * P;
  for (int lvl = order - 1; lvl >= 0; --lvl) {
    const int q = (int)((d >> (2 * lvl)) & 3);
    const int64_t hxi = xi / 2, hxj = xj / 2, hyi = yi / 2, hyj = yj / 2;
    if (moore_top && lvl == order - 1) {
      if (q == 0)      { x0 += hxi;            y0 += hxj;            xi = -hxi; xj = hxj; yi = hyi; yj = hyj; }
      else if (q == 1) { x0 += hxi + hyi;      y0 += hxj + hyj;      xi = -hxi; xj = hxj; yi = hyi; yj = hyj; }
      else if (q == 2) { x0 += hxi + yi;       y0 += hxj + yj;       xi = hxi;  xj = hxj; yi = hyi; yj = -hyj; }
      else             { x0 += hxi + hyi;      y0 += hxj + hyj;      xi = hxi;  xj = hxj; yi = hyi; yj = -hyj; }
    } else {
      if (q == 0)      { xi = hyi; xj = hyj; yi = hxi; yj = hxj; }
      else if (q == 1) { x0 += hxi;            y0 += hxj;            xi = hxi; xj = hxj; yi = hyi; yj = hyj; }
      else if (q == 2) { x0 += hxi + hyi;      y0 += hxj + hyj;      xi = hxi; xj = hxj; yi = hyi; yj = hyj; }
      else             { x0 += hxi + yi;       y0 += hxj + yj;       xi = -hyi; xj = -hyj; yi = -hxi; yj = -hxj; }
    }
  }
  // leaf centre in doubled coordinates is odd; >>1 is the floor of the real centre
  *ox = (int)((x0 + (xi + yi) / 2) >> 1);
  *oy = (int)((y0 + (xj + yj) / 2) >> 1);
}

SFC_HD uint32_t sfc_compact_even_bits(uint64_t v) {
  v &= 0x5555555555555555ull;
  v = (v | (v >> 1)) & 0x3333333333333333ull;
  v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
  v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
  v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
  v = (v | (v >> 16)) & 0x00000000ffffffffull;
  return (uint32_t)v;
}

// Peano pattern table, rows 0/1 (rows 2/3 of the reference table are unreachable from pattern 0).
// entry = dx | dy<<2 | next<<4
SFC_HD int sfc_peano_entry(int pat, int e) {
  // pattern 0: (0,0)0 (1,0)1 (2,0)0 (2,1)1 (1,1)0 (0,1)1 (0,2)0 (1,2)1 (2,2)0
  // pattern 1: (2,0)1 (1,0)0 (0,0)1 (0,1)0 (1,1)1 (2,1)0 (2,2)1 (1,2)0 (0,2)1
  const int row = e / 3, col = e % 3;
  const int ser = (row & 1) ? (2 - col) : col;     // serpentine x within the row
  const int dx = pat ? (2 - ser) : ser;
  const int dy = row;
  const int nxt = (e & 1) ^ pat;
  return dx | (dy << 2) | (nxt << 4);
}

// d -> (i, j) on the padded P x P square, P = base^order.
SFC_HD void sfc_d2ij(int curve, int order, int64_t P, uint64_t d, int* i, int* j) {
  if (curve == SFC_HILBERT) {
    int x, y; sfc_vec_descent(order, d, false, P, &x, &y);
    *i = y; *j = x;                                  // transpose (:196-202)
  } else if (curve == SFC_MOORE) {
    int x, y; sfc_vec_descent(order, d, true, P, &x, &y);
    *i = x; *j = y;
  } else if (curve == SFC_Z) {
    // child 0 = (x+half, y), 1 = (x, y), 2 = (x+half, y+half), 3 = (x, y+half):
    // x-bit = !(digit & 1), y-bit = digit >> 1; identity transform
    *i = (int)(P - 1) - (int)sfc_compact_even_bits(d);
    *j = (int)sfc_compact_even_bits(d >> 1);
  } else if (curve == SFC_PEANO) {
    // base-9 digits, most significant first; state (pat, rev)
    uint64_t pw = 1; for (int l = 1; l < order; ++l) pw *= 9;
    int64_t s = P / 3; int x = 0, y = 0, pat = 0; bool rev = false;
    for (int lvl = 0; lvl < order; ++lvl) {
      const int dig = (int)((d / pw) % 9);
      const int e = rev ? 8 - dig : dig;
      const int ent = sfc_peano_entry(pat, e);
      x += (ent & 3) * (int)s; y += ((ent >> 2) & 3) * (int)s;
      if (e % 3 == 1) rev = !rev;
      pat = ent >> 4;
      pw /= 9; s /= 3;
    }
    *i = y; *j = x;                                  // transpose (:125-131)
  } else {  // raster: identity
    *i = (int)(d / (uint64_t)P); *j = (int)(d % (uint64_t)P);
  }
}
