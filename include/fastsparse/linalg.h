/* linalg.h -- drop-in for libfastsparse's linalg.h: the dot / norm / Gram reductions of
 * the CG solver run on the GPU (fsb_gram_host: two-stage deterministic reduction); the
 * closed-form 2x2 solve is scalar control code and stays inline. */
#ifndef LINALG_H
#define LINALG_H

#include <math.h>

#include "../fsb.h"

static inline double dist(double* x, double* y, int n) {   /* linalg.h:6-13 */
  double d = 0.0;
  if (fsb_dist_host(&d, x, y, n)) fsb_die("dist");
  return d;
}

static inline double pnormsq(double* x, int n) {   /* linalg.h:15-22 */
  double g = 0.0;
  if (fsb_gram_host(&g, x, x, n, 1)) fsb_die("pnormsq");
  return g;
}

static inline double pdot(double* x, double* y, int n) {   /* linalg.h:51-58 */
  double g = 0.0;
  if (fsb_gram_host(&g, x, y, n, 1)) fsb_die("pdot");
  return g;
}

static inline void pnormsq2(double* normsq, double* X, int n) {   /* linalg.h:24-34 */
  double G[4];
  if (fsb_gram_host(G, X, X, n, 2)) fsb_die("pnormsq2");
  normsq[0] = G[0];
  normsq[1] = G[3];
}

/* a'a, b'b, a'b of the 2-column matrix X = [a b] (linalg.h:37-49) */
static inline void pouter2(double* outer, double* X, int n) {
  double G[4];
  if (fsb_gram_host(G, X, X, n, 2)) fsb_die("pouter2");
  outer[0] = G[0];
  outer[1] = G[3];
  outer[2] = G[1];
}

/* symmetric D = X'Y for 2-column X, Y, stored [d00, d11, d01] (linalg.h:61-73) */
static inline void pdot2sym(double* D, double* X, double* Y, int n) {
  double G[4];
  if (fsb_gram_host(G, X, Y, n, 2)) fsb_die("pdot2sym");
  D[0] = G[0];
  D[1] = G[3];
  D[2] = G[1];
}

/* A X = RHS, A = [a0 a2; a2 a1] symmetric, X and RHS column-ordered 2x2 (linalg.h:77-88) */
static inline void solve2sym(double* X, double* A, double* RHS) {
  const double det = A[0] * A[1] - A[2] * A[2];
  const double inv = 1.0 / det;
  const double p = inv * A[1], q = inv * A[0], r = -inv * A[2];
  X[0] = p * RHS[0] + r * RHS[1];
  X[1] = r * RHS[0] + q * RHS[1];
  X[2] = p * RHS[2] + r * RHS[3];
  X[3] = r * RHS[2] + q * RHS[3];
}

#endif /* LINALG_H */
