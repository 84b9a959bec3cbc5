/* quickSort.h -- drop-in for quickSort.h:10-24: ascending in-place sort of a[l..r]. */
#ifndef QUICKSORT_H
#define QUICKSORT_H
#include <stdio.h>
#include "../fsb.h"

static inline void quickSort(long a[], long l, long r) {
  if (r > l) fsb_host_sort_keys(a + l, NULL, r - l + 1);
}
#endif /* QUICKSORT_H */
