/* Minimal stand-in for <gsl/gsl_rng.h>: ONLY the declarations that the reference
 * mh.cpp / util.cpp use (mh.cpp:66,95,102).  TEST INFRASTRUCTURE — libgsl is not
 * installed in this image (and cannot be, no network); see oracle/README.md.
 * Definitions live in gsl_shim.c.  Not GSL source: written from the documented API. */
#ifndef PWGS_SHIM_GSL_RNG_H
#define PWGS_SHIM_GSL_RNG_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct gsl_rng_type_s { const char *name; } gsl_rng_type;
typedef struct gsl_rng_s gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
double gsl_rng_uniform_pos(gsl_rng *r);
#ifdef __cplusplus
}
#endif
#endif
