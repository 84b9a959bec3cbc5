/* hilbert.h -- drop-in for hilbert.h:11-75.  The integer maths lives in
 * libfastsparse_b200.so (fsb_host_*; bit-exact, see tests/test_host_structure.py). */
#ifndef HILBERT_H
#define HILBERT_H
#include <stdlib.h>
#include "../fsb.h"

static inline int ceilPower2(int x) { return fsb_host_ceil_pow2(x); }
static inline long xy2d(int n, int x, int y) { return fsb_host_xy2d(n, x, y); }
static inline void d2xy(int n, long d, int* x, int* y) { fsb_host_d2xy(n, d, x, y); }
static inline long row_xy2d(int n, int x, int y) { return fsb_host_row_xy2d(n, x, y); }
static inline void row_d2xy(int n, long d, int* x, int* y) { fsb_host_row_d2xy(n, d, x, y); }

/* quadrant rotate/flip (hilbert.h:45-57), public in the reference header */
static inline void rot(int n, int* x, int* y, int rx, int ry) {
  if (ry) return;
  if (rx == 1) {
    *x = n - 1 - *x;
    *y = n - 1 - *y;
  }
  const int t = *x;
  *x = *y;
  *y = t;
}
#endif /* HILBERT_H */
