// jit_rt.h -- what the run-time compiler (NVRTC, csrc/jit.cu) sees instead of the host headers:
// ops.h and group_device.h include this file when __CUDACC_RTC__ is defined. NVRTC has no libc
// headers; memcpy and the CUDA device builtins (atomics, __longlong_as_double ...) are built in.
#ifndef SSB_CSRC_JIT_RT_H_
#define SSB_CSRC_JIT_RT_H_

typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;

#define INT32_MIN (-2147483647 - 1)
#define INT64_MAX 9223372036854775807LL
#define INT64_MIN (-9223372036854775807LL - 1)

// include/supersonic_b200.h (the aggregate functions the device code switches on)
enum { SSB_AGG_SUM = 0, SSB_AGG_MIN = 1, SSB_AGG_MAX = 2, SSB_AGG_COUNT = 3, SSB_AGG_FIRST = 5, SSB_AGG_LAST = 6 };

#endif  // SSB_CSRC_JIT_RT_H_
