/* quickSortD.h -- drop-in for quickSortD.h:12-26: sort a[l..r] carrying the payload v. */
#ifndef QUICKSORTD_H
#define QUICKSORTD_H
#include <stdio.h>
#include "../fsb.h"

static inline void quickSortD(long a[], long l, long r, double v[]) {
  if (r > l) fsb_host_sort_keys(a + l, v + l, r - l + 1);
}
#endif /* QUICKSORTD_H */
