// rtc_compat.h — the few host-library names the device sources use, for both compilers: nvcc gets the standard headers, NVRTC
// (run-time compilation of K1 against a constant program; no host headers there) gets minimal equivalents.
#pragma once
#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#ifndef FLT_EPSILON
#define FLT_EPSILON 1.1920928955078125e-07F
#endif
#ifndef INFINITY
#define INFINITY __int_as_float(0x7f800000)
#endif
#ifndef INT32_MIN
#define INT32_MIN (-2147483647 - 1)
#endif
namespace std {
struct true_type { static constexpr bool value = true; };
struct false_type { static constexpr bool value = false; };
}
#else
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <type_traits>
#endif
