/* utils.h -- drop-in for libfastsparse's utils.h (read_long, utils.h:4-12). */
#ifndef UTILS_H
#define UTILS_H
#include <stdio.h>
#include <stdlib.h>

/* one native 8-byte long from the stream; a short read is fatal, as in the reference */
static inline long read_long(FILE* fh) {
  long value = 0;
  if (fread(&value, sizeof value, 1, fh) != 1) {
    fprintf(stderr, "File reading error for a long. File is corrupt.\n");
    exit(1);
  }
  return value;
}
#endif /* UTILS_H */
