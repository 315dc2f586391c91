// Transfer narrowing, host side: an INT64 column whose values of one chunk all fit 32 bits
// crosses PCIe as INT32 and is widened again by a CAST at the kernel's INPUT node. This is the loop the host pool
// runs per chunk and column; it is host-memory bound, so the AVX2 form exists to leave the core to its sibling
// hyperthread, not to compute faster. Header-only so that tests/cpp/narrow_check.cc can test it on the CPU.
#ifndef SSB200_HOST_NARROW_H_
#define SSB200_HOST_NARROW_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace supersonic {
namespace narrow {

// dst[i] = int32(src[i]) for begin <= i < end. Returns non-zero when some value does not fit (dst is then garbage).
inline int64_t RangeScalar(const int64_t* src, int32_t* dst, size_t begin, size_t end) {
  int64_t bad = 0;
  for (size_t i = begin; i < end; ++i) {
    const int64_t v = src[i];
    const int32_t n = static_cast<int32_t>(v);
    dst[i] = n;
    bad |= v ^ static_cast<int64_t>(n);
  }
  return bad;
}

#if defined(__x86_64__)
// Eight values per step: the even dwords of two vectors are gathered with two shuffles; v fits 32 bits exactly
// when v + 2^31 has a zero high dword, so one add and one or per vector keep the check.
__attribute__((target("avx2"))) inline int64_t RangeAvx2(const int64_t* src, int32_t* dst, size_t begin, size_t end) {
  size_t i = begin;
  __m256i bad = _mm256_setzero_si256();
  const __m256i bias = _mm256_set1_epi64x(0x80000000LL);
  for (; i + 8 <= end; i += 8) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 4));
    const __m256 even = _mm256_shuffle_ps(_mm256_castsi256_ps(a), _mm256_castsi256_ps(b), 0x88);   // a0 a1 b0 b1 | a2 a3 b2 b3
    const __m256i packed = _mm256_permute4x64_epi64(_mm256_castps_si256(even), 0xD8);              // a0 a1 a2 a3 b0 b1 b2 b3
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + i), packed);
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_add_epi64(a, bias), _mm256_add_epi64(b, bias)));
  }
  bad = _mm256_srli_epi64(bad, 32);
  int64_t r = _mm256_testz_si256(bad, bad) ? 0 : 1;
  if (i < end) r |= RangeScalar(src, dst, i, end);
  return r;
}
#endif

inline int64_t Range(const int64_t* src, int32_t* dst, size_t begin, size_t end) {
#if defined(__x86_64__)
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) return RangeAvx2(src, dst, begin, end);
#endif
  return RangeScalar(src, dst, begin, end);
}

}  // namespace narrow
}  // namespace supersonic

#endif  // SSB200_HOST_NARROW_H_
