/* Stand-in for libcerf's cerf.h -- TEST INFRASTRUCTURE ONLY (oracle/).
 * libcerf is an external, un-vendored, unpinned dependency of the reference
 * (voigt.c:5, README.md:210-218) and is absent from this image.  Only the one entry
 * point voigt.c calls (voigt.c:288) is declared; shim.c implements it with the
 * Faddeeva-package wofz that SciPy ships (the code libcerf itself wraps). */
#ifndef GPDLA_ORACLE_CERF_SHIM_H
#define GPDLA_ORACLE_CERF_SHIM_H
double voigt(double x, double sigma, double gamma);
#endif
