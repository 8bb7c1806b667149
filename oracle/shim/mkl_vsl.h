/* Stub of Intel MKL's mkl_vsl.h, used ONLY to compile the unmodified reference
 * sources (rand.cpp includes it, /root/reference/LDPC_dec/ldpc/rand.h:8).
 * Intel MKL (build logs name compilers_and_libraries_2020.0.166) is a third-party
 * dependency that is neither vendored in the reference nor installed here.
 * On the BP decode path only rand_seed() is reached (DNA_main.cpp:699-703) and
 * its stream is never consumed, so no-op stand-ins preserve behaviour.
 * TEST INFRASTRUCTURE - not part of the product. */
#pragma once
typedef void *VSLStreamStatePtr;
#define VSL_BRNG_MT2203 0
#define VSL_RNG_METHOD_GAUSSIAN_BOXMULLER 0
static inline int vslNewStream(VSLStreamStatePtr *s, int, unsigned) { *s = (void *)1; return 0; }
static inline int vslDeleteStream(VSLStreamStatePtr *s) { *s = 0; return 0; }
static inline int vslCopyStream(VSLStreamStatePtr *d, VSLStreamStatePtr s) { *d = s; return 0; }
static inline int vslSaveStreamF(VSLStreamStatePtr, const char *) { return 0; }
static inline int vslLoadStreamF(VSLStreamStatePtr *, const char *) { return 0; }
static inline int vdRngUniform(int, VSLStreamStatePtr, int n, double *r, double a, double) { for (int i = 0; i < n; i++) r[i] = a; return 0; }
static inline int vdRngGaussian(int, VSLStreamStatePtr, int n, double *r, double a, double) { for (int i = 0; i < n; i++) r[i] = a; return 0; }
static inline int viRngUniformBits(int, VSLStreamStatePtr, int n, unsigned int *r) { for (int i = 0; i < n; i++) r[i] = 0; return 0; }
