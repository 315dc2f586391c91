// main() for the reference's client programs built against the googletest stand-in.
#include "gtest/gtest.h"
int main(int argc, char** argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS() == 0 ? 0 : 1;
}
