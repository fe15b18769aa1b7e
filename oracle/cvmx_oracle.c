/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 * Plain-C, order-explicit CPU restatement of the cvmatrix fold-matrix path (see
 * cvmx_oracle_body.h for the reference file:line map).  Built by oracle/Makefile into
 * oracle/_build/libcvmx_oracle.so and loaded with ctypes by tests/ only.
 * Parity status: pinned - tests/test_oracle.py checks it against the golden fixtures
 * generated from the live reference (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define REAL double
#define SUF f64
#define SQRT sqrt
#include "cvmx_oracle_body.h"
#undef REAL
#undef SUF
#undef SQRT

#define REAL float
#define SUF f32
#define SQRT sqrtf
#include "cvmx_oracle_body.h"
#undef REAL
#undef SUF
#undef SQRT
