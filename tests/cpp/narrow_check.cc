// CPU test of supersonic_b200/host/src/narrow.h (run by tests/test_host_narrow.py): the AVX2 form against the scalar
// one on random ranges, alignments and the edge values of the 32-bit range; prints throughput for information.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "../../supersonic_b200/host/src/narrow.h"

using namespace supersonic::narrow;

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main() {
  const size_t n = 1 << 16;
  std::vector<int64_t> src(n + 64);
  std::vector<int32_t> a(n + 64), b(n + 64);
  int checked = 0;
  const int64_t edges[] = {2147483647LL, -2147483648LL, 2147483648LL, -2147483649LL, 0, -1, 1LL << 40, -(1LL << 40),
                           INT64_MAX, INT64_MIN, 4294967295LL, 4294967296LL, -4294967296LL};
  for (int trial = 0; trial < 2000; ++trial) {
    for (size_t i = 0; i < n + 64; ++i) src[i] = static_cast<int64_t>(static_cast<int32_t>(rnd()));
    const size_t begin = rnd() % 40, len = rnd() % 300 + (trial % 50 == 0 ? n - 400 : 0), end = begin + len;
    const bool plant = trial % 3 != 0 && len > 0;
    if (plant) src[begin + rnd() % len] = edges[rnd() % (sizeof(edges) / sizeof(edges[0]))];
    memset(a.data(), 0x55, a.size() * 4);
    memset(b.data(), 0x55, b.size() * 4);
    const bool bad_s = RangeScalar(src.data(), a.data(), begin, end) != 0;
    const bool bad_v = Range(src.data(), b.data(), begin, end) != 0;
    bool expect = false;
    for (size_t i = begin; i < end; ++i) expect = expect || src[i] < -2147483648LL || src[i] > 2147483647LL;
    if (bad_s != expect || bad_v != expect) { printf("FAIL fit flag trial %d: %d %d want %d\n", trial, bad_s, bad_v, expect); return 1; }
    if (!expect && memcmp(a.data(), b.data(), a.size() * 4) != 0) { printf("FAIL values trial %d\n", trial); return 1; }
    // nothing outside [begin, end) is written
    for (size_t i = 0; i < begin; ++i) if (b[i] != 0x55555555) { printf("FAIL wrote before begin\n"); return 1; }
    for (size_t i = end; i < b.size(); ++i) if (b[i] != 0x55555555) { printf("FAIL wrote past end\n"); return 1; }
    ++checked;
  }
  const size_t big = 1 << 24;
  std::vector<int64_t> s2(big);
  std::vector<int32_t> d2(big);
  for (size_t i = 0; i < big; ++i) s2[i] = static_cast<int32_t>(i * 2654435761u);
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = std::chrono::steady_clock::now(); const int64_t x = RangeScalar(s2.data(), d2.data(), 0, big);
    auto t1 = std::chrono::steady_clock::now(); const int64_t y = Range(s2.data(), d2.data(), 0, big);
    auto t2 = std::chrono::steady_clock::now();
    printf("scalar %.2f GB/s, dispatched %.2f GB/s (avx2=%d) %lld %lld\n", big * 12 / std::chrono::duration<double>(t1 - t0).count() / 1e9,
           big * 12 / std::chrono::duration<double>(t2 - t1).count() / 1e9, (int)__builtin_cpu_supports("avx2"), (long long)x, (long long)y);
  }
  printf("OK %d trials\n", checked);
  return 0;
}
