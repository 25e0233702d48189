/* Minimal stand-in for <gsl/gsl_randist.h>: only what mh.cpp:88,92 and util.cpp:25 call. */
#ifndef PWGS_SHIM_GSL_RANDIST_H
#define PWGS_SHIM_GSL_RANDIST_H
#include <stddef.h>
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif
void gsl_ran_dirichlet(gsl_rng *r, size_t K, const double alpha[], double theta[]);
double gsl_ran_dirichlet_lnpdf(size_t K, const double alpha[], const double theta[]);
double gsl_ran_gamma(gsl_rng *r, double a, double b);
double gsl_ran_gaussian(gsl_rng *r, double sigma);
#ifdef __cplusplus
}
#endif
#endif
