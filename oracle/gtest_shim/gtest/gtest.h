// Minimal stand-in for the part of googletest the reference's tests/*.cpp use -- TEST INFRASTRUCTURE, written for
// this repository (googletest is a FetchContent download in the reference's CMakeLists.txt:75-82; no network here).
// TEST(suite, name), EXPECT_EQ, EXPECT_FLOAT_EQ (4 ulp of float, as googletest), EXPECT_THROW, ASSERT_* aliases and a
// main() that runs every registered test, prints googletest-style lines and returns the number of failed tests.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <string>
#include <vector>

namespace testing {
struct TestCase { std::string suite, name; std::function<void()> body; };
inline std::vector<TestCase> &registry() { static std::vector<TestCase> r; return r; }
inline int &failures_in_current() { static int f = 0; return f; }
struct Registrar {
    Registrar(const char *s, const char *n, std::function<void()> b) { registry().push_back({s, n, std::move(b)}); }
};
inline void fail(const char *file, int line, const std::string &what) {
    ++failures_in_current();
    std::cout << file << ":" << line << ": Failure\n" << what << std::endl;
}
inline bool float_almost_equal(float a, float b) {
    // googletest: equal when within 4 units in the last place
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](float f) {
        std::uint32_t u; std::memcpy(&u, &f, 4);
        return (u & 0x80000000u) ? (~u + 1) : (u | 0x80000000u);
    };
    std::uint32_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}
inline int run_all() {
    int failed = 0;
    std::cout << "[==========] Running " << registry().size() << " tests." << std::endl;
    for (auto &t: registry()) {
        std::cout << "[ RUN      ] " << t.suite << "." << t.name << std::endl;
        failures_in_current() = 0;
        try { t.body(); } catch (const std::exception &e) { fail("(exception)", 0, e.what()); }
        if (failures_in_current()) { ++failed; std::cout << "[  FAILED  ] "; } else std::cout << "[       OK ] ";
        std::cout << t.suite << "." << t.name << std::endl;
    }
    std::cout << "[==========] " << registry().size() << " tests ran.\n[  PASSED  ] " << registry().size() - failed << " tests." << std::endl;
    if (failed) std::cout << "[  FAILED  ] " << failed << " tests." << std::endl;
    return failed;
}
}  // namespace testing

#define TEST(suite, name)                                                                         \
    static void suite##_##name##_body();                                                          \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body);       \
    static void suite##_##name##_body()

#define EXPECT_EQ(a, b)                                                                           \
    do {                                                                                          \
        auto &&va_ = (a); auto &&vb_ = (b);                                                       \
        if (!(va_ == vb_)) { ::testing::fail(__FILE__, __LINE__, std::string("Expected equality of: ") + #a + " and " + #b); \
            std::cout << "  " << va_ << " vs " << vb_ << std::endl; }                             \
    } while (0)
#define EXPECT_FLOAT_EQ(a, b)                                                                     \
    do {                                                                                          \
        float va_ = static_cast<float>(a), vb_ = static_cast<float>(b);                          \
        if (!::testing::float_almost_equal(va_, vb_)) { ::testing::fail(__FILE__, __LINE__, std::string("Expected float equality of: ") + #a + " and " + #b); \
            std::cout << "  " << va_ << " vs " << vb_ << std::endl; }                             \
    } while (0)
#define EXPECT_TRUE(c) do { if (!(c)) ::testing::fail(__FILE__, __LINE__, std::string("Expected true: ") + #c); } while (0)
#define EXPECT_FALSE(c) do { if (c) ::testing::fail(__FILE__, __LINE__, std::string("Expected false: ") + #c); } while (0)
#define EXPECT_THROW(stmt, exc)                                                                   \
    do {                                                                                          \
        bool thrown_ = false;                                                                     \
        try { stmt; } catch (const exc &) { thrown_ = true; } catch (...) {}                      \
        if (!thrown_) ::testing::fail(__FILE__, __LINE__, std::string("Expected: ") + #stmt + " throws " + #exc); \
    } while (0)
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_TRUE EXPECT_TRUE
#define ASSERT_FALSE EXPECT_FALSE
#define ASSERT_FLOAT_EQ EXPECT_FLOAT_EQ
#define ASSERT_THROW EXPECT_THROW
