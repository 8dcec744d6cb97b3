/* strict_math_map.h -- TEST INFRASTRUCTURE (oracle/): force-included (-include) into the four reference translation units that
 * call log / sin / cos on the tracking path -- src/MC_Segment_Outcome.cc:83, src/CollisionEvent.cc:31-32,44,
 * src/MC_SourceNow.cc:119, src/DirectionCosine.cc:11-12 -- so that the UNMODIFIED reference sources evaluate them with the
 * portable functions of csrc/qs_strict_math.h (the ones the GPU validation kernels and the oracle's strict mode use).  The
 * resulting binary, oracle/_ref/qs_dump_strict, is the reference itself in strict-math mode: its census records can be compared
 * byte for byte with the oracle chain that checks the GPU (tests/test_oracle_golden.py). */
#ifndef QSB_ORACLE_STRICT_MATH_MAP_H
#define QSB_ORACLE_STRICT_MATH_MAP_H
#include <cmath>
#include <math.h>
#include <cstdint>
#include "qs_strict_math.h"
/* qs_strict_sincos takes 0 <= phi < 8; the source's azimuth lies in (-pi, pi) (src/DirectionCosine.cc:9-12): odd / even symmetry,
 * as the strict-math mode of csrc/qs_cycle_init.h does */
static inline double qs_strict_sin_only(double x) { double s, c; qs_strict_sincos(x < 0.0 ? -x : x, &s, &c); return x < 0.0 ? -s : s; }
static inline double qs_strict_cos_only(double x) { double s, c; qs_strict_sincos(x < 0.0 ? -x : x, &s, &c); return c; }
#define log(x) qs_strict_log(x)
#define sin(x) qs_strict_sin_only(x)
#define cos(x) qs_strict_cos_only(x)
#endif
