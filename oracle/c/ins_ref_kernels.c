/* ins_ref_kernels.c -- ORACLE (test infrastructure): C/OpenMP restatement of the stencil
 * kernels of the reference's CPU path (src/operators.jl), used (a) to cross-check the NumPy
 * oracle and (b) as the multithreaded CPU baseline of bench.py (`cpu_baseline`, kind "port").
 * Nothing in the product path links or loads this file. */
#include <float.h>
#include <stddef.h>

#define REAL float
#define EPS FLT_EPSILON
#define FN(name) insref_##name##_f32
#include "kernels_body.h"
#undef REAL
#undef EPS
#undef FN

#define REAL double
#define EPS DBL_EPSILON
#define FN(name) insref_##name##_f64
#include "kernels_body.h"
#undef REAL
#undef EPS
#undef FN

#ifdef _OPENMP
#include <omp.h>
int insref_max_threads(void) { return omp_get_max_threads(); }
void insref_set_threads(int n) { omp_set_num_threads(n); }
#else
int insref_max_threads(void) { return 1; }
void insref_set_threads(int n) { (void)n; }
#endif
