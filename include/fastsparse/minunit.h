/* minunit.h -- drop-in for the two-macro test harness the reference tests use
 * (minunit.h:2-5): a failing assertion returns its message, a passing test prints a dot. */
#ifndef MINUNIT_H
#define MINUNIT_H
extern int tests_run;
#define mu_assert(message, test) \
  do {                           \
    if (!(test)) return message; \
  } while (0)
#define mu_run_test(test)        \
  do {                           \
    char* mu_msg_ = test();      \
    ++tests_run;                 \
    if (mu_msg_) return mu_msg_; \
    printf(".");                 \
  } while (0)
#endif /* MINUNIT_H */
