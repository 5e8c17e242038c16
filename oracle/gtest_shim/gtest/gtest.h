// Minimal GoogleTest-compatible shim (test infrastructure, see oracle/README.md).
// The image has no gtest and no network; the reference's test/testTensor.cu uses only TEST_F,
// testing::Test, EXPECT_{EQ,NE,LT,NEAR,TRUE,THROW} and ASSERT_{EQ,LT}. This header provides exactly that
// plus a main() with --gtest_filter and gtest-style output, so the reference test file compiles unchanged.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing {

class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
};

namespace shim {

struct Case {
    std::string suite, name;
    std::function<Test *()> make;
};

inline std::vector<Case> &registry() {
    static std::vector<Case> r;
    return r;
}

inline int &current_failures() {
    static int f = 0;
    return f;
}

struct Registrar {
    Registrar(const char *suite, const char *name, std::function<Test *()> make) {
        registry().push_back(Case{suite, name, std::move(make)});
    }
};

template<typename V>
auto show(std::ostream &os, const V &v, int) -> decltype(os << v, void()) { os << v; }
template<typename V>
void show(std::ostream &os, const V &, long) { os << "<unprintable>"; }
template<typename V>
void show(std::ostream &os, const std::vector<V> &v, int) {
    os << "{";
    for (size_t i = 0; i < v.size() && i < 32; i++) os << (i ? ", " : "") << v[i];
    if (v.size() > 32) os << ", ...";
    os << "}";
}

inline void fail(const char *file, int line, const std::string &msg) {
    current_failures()++;
    std::cout << file << ":" << line << ": Failure\n" << msg << std::endl;
}

template<typename A, typename B>
bool cmp_eq(const A &a, const B &b) {
    if constexpr (std::is_arithmetic_v<A> && std::is_arithmetic_v<B>) {
        using C = std::common_type_t<A, B>;
        return static_cast<C>(a) == static_cast<C>(b);
    } else {
        return a == b;
    }
}
template<typename A, typename B>
bool cmp_lt(const A &a, const B &b) {
    if constexpr (std::is_arithmetic_v<A> && std::is_arithmetic_v<B>) {
        using C = std::common_type_t<A, B>;
        return static_cast<C>(a) < static_cast<C>(b);
    } else {
        return a < b;
    }
}

template<typename A, typename B>
bool check_binary(bool ok, const char *op, const char *ea, const char *eb, const A &a, const B &b, const char *file, int line) {
    if (ok) return true;
    std::ostringstream os;
    os << "Expected: (" << ea << ") " << op << " (" << eb << "), actual: ";
    show(os, a, 0);
    os << " vs ";
    show(os, b, 0);
    fail(file, line, os.str());
    return false;
}

inline bool check_near(double a, double b, double tol, const char *ea, const char *eb, const char *file, int line) {
    double diff = std::fabs(a - b);
    if (diff <= tol) return true; // NaN fails, like gtest
    std::ostringstream os;
    os.precision(17);
    os << "The difference between " << ea << " and " << eb << " is " << diff << ", which exceeds " << tol << ", where\n"
       << ea << " evaluates to " << a << ",\n" << eb << " evaluates to " << b << ".";
    fail(file, line, os.str());
    return false;
}

} // namespace shim
} // namespace testing

#define GTEST_SHIM_CLASS(suite, name) suite##_##name##_Test

#define TEST_F(suite, name)                                                                                   \
    class GTEST_SHIM_CLASS(suite, name) : public suite {                                                      \
    public:                                                                                                   \
        void TestBody() override;                                                                             \
    };                                                                                                        \
    static ::testing::shim::Registrar gtest_shim_reg_##suite##_##name(                                        \
        #suite, #name, []() -> ::testing::Test * { return new GTEST_SHIM_CLASS(suite, name)(); });            \
    void GTEST_SHIM_CLASS(suite, name)::TestBody()

#define TEST(suite, name)                                                                                     \
    class GTEST_SHIM_CLASS(suite, name) : public ::testing::Test {                                            \
    public:                                                                                                   \
        void TestBody() override;                                                                             \
    };                                                                                                        \
    static ::testing::shim::Registrar gtest_shim_reg_##suite##_##name(                                        \
        #suite, #name, []() -> ::testing::Test * { return new GTEST_SHIM_CLASS(suite, name)(); });            \
    void GTEST_SHIM_CLASS(suite, name)::TestBody()

#define GTEST_SHIM_BIN(ok_expr, op, a, b, on_fail)                                                            \
    do {                                                                                                      \
        const auto &gs_a_ = (a);                                                                              \
        const auto &gs_b_ = (b);                                                                              \
        if (!::testing::shim::check_binary((ok_expr), op, #a, #b, gs_a_, gs_b_, __FILE__, __LINE__)) { on_fail; } \
    } while (0)

#define EXPECT_EQ(a, b) GTEST_SHIM_BIN(::testing::shim::cmp_eq(gs_a_, gs_b_), "==", a, b, (void) 0)
#define ASSERT_EQ(a, b) GTEST_SHIM_BIN(::testing::shim::cmp_eq(gs_a_, gs_b_), "==", a, b, return)
#define EXPECT_NE(a, b) GTEST_SHIM_BIN(!::testing::shim::cmp_eq(gs_a_, gs_b_), "!=", a, b, (void) 0)
#define ASSERT_NE(a, b) GTEST_SHIM_BIN(!::testing::shim::cmp_eq(gs_a_, gs_b_), "!=", a, b, return)
#define EXPECT_LT(a, b) GTEST_SHIM_BIN(::testing::shim::cmp_lt(gs_a_, gs_b_), "<", a, b, (void) 0)
#define ASSERT_LT(a, b) GTEST_SHIM_BIN(::testing::shim::cmp_lt(gs_a_, gs_b_), "<", a, b, return)
#define EXPECT_GT(a, b) GTEST_SHIM_BIN(::testing::shim::cmp_lt(gs_b_, gs_a_), ">", a, b, (void) 0)
#define EXPECT_LE(a, b) GTEST_SHIM_BIN(!::testing::shim::cmp_lt(gs_b_, gs_a_), "<=", a, b, (void) 0)
#define EXPECT_GE(a, b) GTEST_SHIM_BIN(!::testing::shim::cmp_lt(gs_a_, gs_b_), ">=", a, b, (void) 0)

#define EXPECT_NEAR(a, b, tol)                                                                                 \
    do {                                                                                                       \
        ::testing::shim::check_near((double) (a), (double) (b), (double) (tol), #a, #b, __FILE__, __LINE__);   \
    } while (0)
#define ASSERT_NEAR(a, b, tol)                                                                                 \
    do {                                                                                                       \
        if (!::testing::shim::check_near((double) (a), (double) (b), (double) (tol), #a, #b, __FILE__, __LINE__)) return; \
    } while (0)

#define EXPECT_TRUE(c)                                                                                         \
    do {                                                                                                       \
        if (!(c)) ::testing::shim::fail(__FILE__, __LINE__, std::string("Value of: ") + #c + "\n  Actual: false\nExpected: true"); \
    } while (0)
#define ASSERT_TRUE(c)                                                                                         \
    do {                                                                                                       \
        if (!(c)) {                                                                                            \
            ::testing::shim::fail(__FILE__, __LINE__, std::string("Value of: ") + #c + "\n  Actual: false\nExpected: true"); \
            return;                                                                                            \
        }                                                                                                      \
    } while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define ASSERT_FALSE(c) ASSERT_TRUE(!(c))

#define EXPECT_THROW(statement, exc)                                                                           \
    do {                                                                                                       \
        bool gs_caught_ = false;                                                                               \
        try {                                                                                                  \
            statement;                                                                                         \
        } catch (const exc &) {                                                                                \
            gs_caught_ = true;                                                                                 \
        } catch (...) {                                                                                        \
            ::testing::shim::fail(__FILE__, __LINE__, std::string("Expected: ") + #statement + " throws " #exc ".\n  Actual: it throws a different type."); \
            gs_caught_ = true;                                                                                 \
        }                                                                                                      \
        if (!gs_caught_)                                                                                       \
            ::testing::shim::fail(__FILE__, __LINE__, std::string("Expected: ") + #statement + " throws " #exc ".\n  Actual: it throws nothing."); \
    } while (0)

#ifndef GTEST_SHIM_NO_MAIN
int main(int argc, char **argv) {
    std::string filter = "*";
    for (int i = 1; i < argc; i++)
        if (std::strncmp(argv[i], "--gtest_filter=", 15) == 0) filter = argv[i] + 15;
    auto matches = [&](const std::string &full) {
        if (filter == "*") return true;
        // supports "Suite.*", "*name*" and exact names, ':'-separated
        std::stringstream ss(filter);
        std::string pat;
        while (std::getline(ss, pat, ':')) {
            std::string core = pat;
            bool pre = !core.empty() && core.front() == '*', post = !core.empty() && core.back() == '*';
            if (pre) core.erase(0, 1);
            if (post && !core.empty()) core.pop_back();
            if (pre && post ? full.find(core) != std::string::npos
                : pre      ? full.size() >= core.size() && full.compare(full.size() - core.size(), core.size(), core) == 0
                : post     ? full.compare(0, core.size(), core) == 0
                           : full == core)
                return true;
        }
        return false;
    };
    auto &reg = ::testing::shim::registry();
    int ran = 0, failed = 0;
    std::vector<std::string> failed_names;
    std::cout << "[==========] Running tests from the gtest shim." << std::endl;
    for (auto &c: reg) {
        const std::string full = c.suite + "." + c.name;
        if (!matches(full)) continue;
        std::cout << "[ RUN      ] " << full << std::endl;
        ::testing::shim::current_failures() = 0;
        try {
            ::testing::Test *t = c.make();
            t->SetUp();
            t->TestBody();
            t->TearDown();
            delete t;
        } catch (const std::exception &e) {
            ::testing::shim::fail("<exception>", 0, std::string("uncaught exception: ") + e.what());
        } catch (...) {
            ::testing::shim::fail("<exception>", 0, "uncaught non-standard exception");
        }
        ran++;
        if (::testing::shim::current_failures()) {
            failed++;
            failed_names.push_back(full);
            std::cout << "[  FAILED  ] " << full << std::endl;
        } else {
            std::cout << "[       OK ] " << full << std::endl;
        }
    }
    std::cout << "[==========] " << ran << " tests ran." << std::endl;
    std::cout << "[  PASSED  ] " << (ran - failed) << " tests." << std::endl;
    if (failed) {
        std::cout << "[  FAILED  ] " << failed << " tests, listed below:" << std::endl;
        for (auto &n: failed_names) std::cout << "[  FAILED  ] " << n << std::endl;
    }
    return failed ? 1 : 0;
}
#endif
