// main() of the reference's test executables (GTest::gtest_main in the reference's CMakeLists.txt:105-106)
#include <gtest/gtest.h>
int main() { return ::testing::run_all(); }
