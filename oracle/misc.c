/* misc.c — oracle helpers.  TEST INFRASTRUCTURE (see p2oracle.h). */
#include "p2oracle.h"
#ifdef _OPENMP
#include <omp.h>
unsigned p2o_num_threads(void) { return (unsigned)omp_get_max_threads(); }
#else
unsigned p2o_num_threads(void) { return 1; }
#endif
