// Minimal stand-in for googletest (not in the image), enough for the reference's own client
// programs under test/guide/ to compile UNMODIFIED against this repo's supersonic.h and run:
// TEST registers a function, EXPECT_* / ASSERT_* count failures and print them, RUN_ALL_TESTS runs
// everything and returns the number of failed checks.
#ifndef SSB200_TESTS_GTEST_STUB_H_
#define SSB200_TESTS_GTEST_STUB_H_
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct Registry {
  struct Entry { const char* suite; const char* name; void (*fn)(); };
  static std::vector<Entry>& tests() { static std::vector<Entry> t; return t; }
  static int& failures() { static int f = 0; return f; }
};
struct Registrar {
  Registrar(const char* suite, const char* name, void (*fn)()) { Registry::tests().push_back({suite, name, fn}); }
};
// Streams the user's message when the check failed; swallows it otherwise.
class Message {
 public:
  Message(bool failed, const char* file, int line, const std::string& what) : failed_(failed) {
    if (failed_) { ++Registry::failures(); os_ << file << ":" << line << ": Failure: " << what << " "; }
  }
  ~Message() { if (failed_) std::cerr << os_.str() << std::endl; }
  template <typename T> Message& operator<<(const T& v) { if (failed_) os_ << v; return *this; }
 private:
  bool failed_;
  std::ostringstream os_;
};
inline void InitGoogleTest(int*, char**) {}
// Fixture base of TEST_F: a fresh object per test, SetUp() before and TearDown() after the body.
class Test {
 public:
  virtual ~Test() {}
 protected:
  Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
 public:
  void Run() { SetUp(); TestBody(); TearDown(); }
};
}  // namespace testing

#define TEST(suite, name)                                                              \
  static void suite##_##name##_body();                                                 \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, &suite##_##name##_body); \
  static void suite##_##name##_body()

#define TEST_F(fixture, name)                                                          \
  class fixture##_##name##_Test : public fixture {                                     \
   public:                                                                             \
    fixture##_##name##_Test() {}                                                       \
    virtual void TestBody();                                                           \
    static void RunIt() { fixture##_##name##_Test t; t.Run(); }                        \
  };                                                                                   \
  static ::testing::Registrar fixture##_##name##_reg(#fixture, #name, &fixture##_##name##_Test::RunIt); \
  void fixture##_##name##_Test::TestBody()

#define SSB_GT_CHECK(cond, text) ::testing::Message(!(cond), __FILE__, __LINE__, text)
#define EXPECT_TRUE(c) SSB_GT_CHECK((c), "expected true: " #c)
#define EXPECT_FALSE(c) SSB_GT_CHECK(!(c), "expected false: " #c)
#define EXPECT_EQ(a, b) SSB_GT_CHECK(((a) == (b)), "expected equal: " #a " and " #b)
#define EXPECT_NE(a, b) SSB_GT_CHECK(((a) != (b)), "expected different: " #a " and " #b)
#define EXPECT_LT(a, b) SSB_GT_CHECK(((a) < (b)), "expected " #a " < " #b)
#define EXPECT_LE(a, b) SSB_GT_CHECK(((a) <= (b)), "expected " #a " <= " #b)
#define EXPECT_GT(a, b) SSB_GT_CHECK(((a) > (b)), "expected " #a " > " #b)
#define EXPECT_GE(a, b) SSB_GT_CHECK(((a) >= (b)), "expected " #a " >= " #b)
#define EXPECT_DOUBLE_EQ(a, b) SSB_GT_CHECK(((a) == (b)), "expected equal doubles: " #a " and " #b)
#define ASSERT_TRUE EXPECT_TRUE
#define ASSERT_FALSE EXPECT_FALSE
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_NE EXPECT_NE

inline int RUN_ALL_TESTS() {
  for (const ::testing::Registry::Entry& e : ::testing::Registry::tests()) {
    const int before = ::testing::Registry::failures();
    std::printf("[ RUN      ] %s.%s\n", e.suite, e.name);
    e.fn();
    std::printf("[ %s ] %s.%s\n", ::testing::Registry::failures() == before ? "      OK" : " FAILED ", e.suite, e.name);
  }
  std::printf("[ %d failed checks ]\n", ::testing::Registry::failures());
  return ::testing::Registry::failures();
}
#endif  // SSB200_TESTS_GTEST_STUB_H_
