/* TEST INFRASTRUCTURE ONLY (oracle build shim).
 *
 * utility_cchol / utility_schol of the reference (saf_utility_veclib.c:4103-4160) hand the CBLAS enum CblasUpper (121)
 * to LAPACKE_?potrf_work, whose `uplo` argument is the CHARACTER 'U' / 'L'.  A real LAPACKE therefore rejects the call
 * ("On entry to CPOTRF parameter number 1 had an illegal value"), utility_cchol zeroes its result, and every caller of
 * the Cholesky factor (applyDiffCovMatching, saf_hoa.c:497-604) silently produces zeros in a LAPACKE build -- the
 * CLAPACK / Fortran interfaces of the same file pass the upper triangle as intended.  This shim is linked in front of
 * OpenBLAS and translates the enum to the character, so that the UNMODIFIED reference sources run with the semantics
 * their authors wrote ("Upper").  Nothing else is intercepted.
 */
#include <complex.h>
extern void cpotrf_(const char* uplo, const int* n, float _Complex* a, const int* lda, int* info);
extern void spotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info);

static char fix_uplo(int u) { return (u == 121 || u == 'U' || u == 'u') ? 'U' : 'L'; }

int LAPACKE_cpotrf_work(int layout, int uplo, int n, float _Complex* a, int lda)
{
    (void)layout;   /* the reference always passes column-major data (veclib.c:4123-4131) */
    char u = fix_uplo(uplo); int info = 0;
    cpotrf_(&u, &n, a, &lda, &info);
    return info;
}
int LAPACKE_spotrf_work(int layout, int uplo, int n, float* a, int lda)
{
    (void)layout;
    char u = fix_uplo(uplo); int info = 0;
    spotrf_(&u, &n, a, &lda, &info);
    return info;
}
