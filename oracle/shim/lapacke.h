/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * Stand-in for <lapacke.h> (wanted by /root/reference/framework/include/saf_externals.h:145).
 * The convolver path calls no LAPACK routine; the types below merely let
 * saf_utility_veclib.c compile.  Its LAPACKE_* calls compile as implicit
 * declarations (gnu99) and are left unresolved in the shared object
 * (never reached from saf_matrixConv / saf_multiConv / saf_TVConv).
 */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#include <complex.h>
typedef int lapack_int;
typedef float  _Complex lapack_complex_float;
typedef double _Complex lapack_complex_double;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#endif
