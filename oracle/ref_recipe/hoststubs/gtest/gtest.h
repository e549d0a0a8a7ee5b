// TEST INFRASTRUCTURE ONLY. Stands in for googletest (fetched at configure time by the reference, not in
// this image) so that the reference's OWN unit tests can be compiled unmodified and run on top of the
// other stand-ins -- which is how those stand-ins are themselves checked: if GLM's operators, the
// host-memory cl.hpp or the FFT were wrong, the reference's tests of its own geometry, kernels and
// filters would say so. Only what those test files use: TEST, TEST_F, ::testing::Test,
// ASSERT_/EXPECT_ {EQ, NE, TRUE, FALSE, NEAR, LT, LE, GT, GE, NO_THROW, THROW, ANY_THROW}, a streamed
// message after an assertion, InitGoogleTest, RUN_ALL_TESTS. A failed ASSERT leaves the test body like
// googletest's does; values are not printed.
#pragma once
// googletest 1.8's gtest.h brings these in, and the reference's test files rely on that
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <vector>

namespace testing {

class Test {
public:
    virtual ~Test() = default;
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};

struct stub_case {
    std::string name;
    std::function<Test*()> make;
};
inline std::vector<stub_case>& stub_cases() {
    static std::vector<stub_case> cases;
    return cases;
}
inline int& stub_failures_in_current() {
    static int n = 0;
    return n;
}
struct stub_registrar {
    stub_registrar(const char* name, std::function<Test*()> make) { stub_cases().push_back({name, std::move(make)}); }
};

// `ASSERT_x(...) << "text"` : the streamed text is collected and printed with the failure
class stub_failure {
public:
    stub_failure(const char* file, int line, const char* what) {
        ++stub_failures_in_current();
        text_ << file << ":" << line << ": failed: " << what << " ";
    }
    template <typename T>
    stub_failure& operator<<(const T& t) {
        text_ << t;
        return *this;
    }
    ~stub_failure() { std::printf("%s\n", text_.str().c_str()); }

private:
    std::ostringstream text_;
};
struct stub_void {
    void operator=(const stub_failure&) const {}
};

inline void InitGoogleTest(int*, char**) {}

inline int stub_run_all() {
    int failed = 0;
    for (const auto& c : stub_cases()) {
        stub_failures_in_current() = 0;
        try {
            Test* t = c.make();
            t->SetUp();
            t->TestBody();
            t->TearDown();
            delete t;
        } catch (const std::exception& e) {
            ++stub_failures_in_current();
            std::printf("exception: %s\n", e.what());
        }
        const bool ok = stub_failures_in_current() == 0;
        std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", c.name.c_str());
        failed += !ok;
    }
    std::printf("REFERENCE_TESTS %zu run, %d failed\n", stub_cases().size(), failed);
    return failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::stub_run_all()

#define GTEST_STUB_CLASS(suite, name) suite##_##name##_Test
#define GTEST_STUB_TEST(suite, name, base)                                                        \
    class GTEST_STUB_CLASS(suite, name) : public base {                                           \
    public:                                                                                       \
        void TestBody() override;                                                                 \
    };                                                                                            \
    static ::testing::stub_registrar suite##_##name##_registrar(                                  \
            #suite "." #name, [] { return static_cast<::testing::Test*>(new GTEST_STUB_CLASS(suite, name)); }); \
    void GTEST_STUB_CLASS(suite, name)::TestBody()
#define TEST(suite, name) GTEST_STUB_TEST(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_STUB_TEST(fixture, name, fixture)

// fatal: leave the (void) function, like googletest
#define GTEST_STUB_ASSERT(cond, what) \
    if (cond)                         \
        ;                             \
    else                              \
        return ::testing::stub_void{} = ::testing::stub_failure(__FILE__, __LINE__, what)
// non-fatal: record and carry on
#define GTEST_STUB_EXPECT(cond, what) \
    if (cond)                         \
        ;                             \
    else                              \
        ::testing::stub_failure(__FILE__, __LINE__, what)

#define ASSERT_TRUE(x) GTEST_STUB_ASSERT(static_cast<bool>(x), #x)
#define ASSERT_FALSE(x) GTEST_STUB_ASSERT(!static_cast<bool>(x), "!(" #x ")")
#define ASSERT_EQ(a, b) GTEST_STUB_ASSERT((a) == (b), #a " == " #b)
#define ASSERT_NE(a, b) GTEST_STUB_ASSERT(!((a) == (b)), #a " != " #b)
#define ASSERT_LT(a, b) GTEST_STUB_ASSERT((a) < (b), #a " < " #b)
#define ASSERT_LE(a, b) GTEST_STUB_ASSERT((a) <= (b), #a " <= " #b)
#define ASSERT_GT(a, b) GTEST_STUB_ASSERT((a) > (b), #a " > " #b)
#define ASSERT_GE(a, b) GTEST_STUB_ASSERT((a) >= (b), #a " >= " #b)
#define ASSERT_NEAR(a, b, tol) GTEST_STUB_ASSERT(std::fabs(double(a) - double(b)) <= double(tol), #a " ~ " #b)
#define EXPECT_TRUE(x) GTEST_STUB_EXPECT(static_cast<bool>(x), #x)
#define EXPECT_FALSE(x) GTEST_STUB_EXPECT(!static_cast<bool>(x), "!(" #x ")")
#define EXPECT_EQ(a, b) GTEST_STUB_EXPECT((a) == (b), #a " == " #b)
#define EXPECT_NE(a, b) GTEST_STUB_EXPECT(!((a) == (b)), #a " != " #b)
#define EXPECT_LT(a, b) GTEST_STUB_EXPECT((a) < (b), #a " < " #b)
#define EXPECT_LE(a, b) GTEST_STUB_EXPECT((a) <= (b), #a " <= " #b)
#define EXPECT_GT(a, b) GTEST_STUB_EXPECT((a) > (b), #a " > " #b)
#define EXPECT_GE(a, b) GTEST_STUB_EXPECT((a) >= (b), #a " >= " #b)
#define EXPECT_NEAR(a, b, tol) GTEST_STUB_EXPECT(std::fabs(double(a) - double(b)) <= double(tol), #a " ~ " #b)

#define GTEST_STUB_THROWS(statement, expect_throw, kind, what)                                    \
    if ([&] {                                                                                     \
            try {                                                                                 \
                statement;                                                                        \
            } catch (...) {                                                                       \
                return expect_throw;                                                              \
            }                                                                                     \
            return !expect_throw;                                                                 \
        }())                                                                                      \
        ;                                                                                         \
    else                                                                                          \
        kind ::testing::stub_failure(__FILE__, __LINE__, what)
#define ASSERT_NO_THROW(s) GTEST_STUB_THROWS(s, false, return ::testing::stub_void{} =, "no throw: " #s)
#define ASSERT_ANY_THROW(s) GTEST_STUB_THROWS(s, true, return ::testing::stub_void{} =, "throws: " #s)
#define ASSERT_THROW(s, type) GTEST_STUB_THROWS(s, true, return ::testing::stub_void{} =, "throws " #type ": " #s)
#define EXPECT_NO_THROW(s) GTEST_STUB_THROWS(s, false, , "no throw: " #s)
#define EXPECT_ANY_THROW(s) GTEST_STUB_THROWS(s, true, , "throws: " #s)
#define EXPECT_THROW(s, type) GTEST_STUB_THROWS(s, true, , "throws " #type ": " #s)
