/* Test infrastructure: interposes time() for the reference degensac build so that
 * srand(time(NULL)) (degensac/exp_ranH.c:823, exp_ranF.c:832) is reproducible.
 * Linked INTO oracle/_ref/libdegensac_ref.so only (symbol resolution inside that .so is
 * forced with -Bsymbolic-free default lookup: the .so defines time() itself). */
#include <time.h>
static time_t g_pinned_time = 12345;
void orc_ref_set_time(long t) { g_pinned_time = (time_t)t; }
time_t time(time_t* p) { if (p) *p = g_pinned_time; return g_pinned_time; }
