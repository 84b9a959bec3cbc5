/* timing.h -- drop-in for timing.h:9-19: wall clock (gettimeofday) and user CPU time. */
#ifndef FSB_TIMING_H
#define FSB_TIMING_H
#include <stdio.h>
#include <stdlib.h>
#include <sys/resource.h>
#include <sys/time.h>
#include <sys/types.h>

static inline void timing(double* wcTime, double* cpuTime) {
  struct timeval now;
  struct rusage use;
  gettimeofday(&now, NULL);
  getrusage(RUSAGE_SELF, &use);
  *wcTime = (double)now.tv_sec + (double)now.tv_usec * 1e-6;
  *cpuTime = (double)use.ru_utime.tv_sec + (double)use.ru_utime.tv_usec * 1e-6;
}
#endif /* FSB_TIMING_H */
