/* utils.h -- drop-in for libfastsparse's utils.h: read_long (utils.h:4-12) forwards to the library. */
#ifndef FSB_DROPIN_UTILS_H
#define FSB_DROPIN_UTILS_H
#include <stdio.h>

#include "../fsb.h"

/* a short read is fatal, with the reference's message and exit code */
static inline long read_long(FILE* fh) {
  int ok = 0;
  const long v = fsb_host_read_long(fh, &ok);
  if (!ok) fsb_die("read_long");
  return v;
}
#endif
