/*
 * ORACLE build shim (test infrastructure): just enough of the googletest surface
 * (TEST, ASSERT_TRUE, InitGoogleTest, RUN_ALL_TESTS) to compile and run the reference's
 * ffiasm/c/alt_bn128_test.cpp unmodified.  googletest itself is downloaded by the reference's
 * build (ffiasm/tasksfile.js:7-17) and is not available offline.
 */
#ifndef ORACLE_SHIM_GTEST_H
#define ORACLE_SHIM_GTEST_H
#include <stdio.h>
#include <vector>

namespace testing {
struct Case { const char *suite; const char *name; void (*fn)(bool &); };
inline std::vector<Case> &registry() { static std::vector<Case> r; return r; }
struct Registrar { Registrar(const char *s, const char *n, void (*f)(bool &)) { registry().push_back({s, n, f}); } };
inline void InitGoogleTest(int *, char **) {}
}

#define TEST(suite, name)                                                              \
    static void suite##_##name##_body(bool &gtest_failed);                             \
    static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
    static void suite##_##name##_body(bool &gtest_failed)

#define ASSERT_TRUE(cond)                                                              \
    do { if (!(cond)) { printf("  ASSERT_TRUE failed: %s (line %d)\n", #cond, __LINE__); \
                        gtest_failed = true; return; } } while (0)

inline int RUN_ALL_TESTS()
{
    int failed = 0;
    for (auto &c : ::testing::registry()) {
        bool f = false;
        c.fn(f);
        printf("[%s] %s.%s\n", f ? "FAILED" : "    OK", c.suite, c.name);
        failed += f;
    }
    printf("%zu tests, %d failed\n", ::testing::registry().size(), failed);
    return failed ? 1 : 0;
}
#endif
