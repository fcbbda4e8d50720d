// Minimal test macros for the C++ host tests (test scaffolding; Boost.Test is not in this image).
#ifndef PB200_TEST_CHECK_HPP
#define PB200_TEST_CHECK_HPP
#include <cstdio>
#include <exception>
#include <iostream>
#include <sstream>

static int g_failures = 0, g_checks = 0;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) {                                                               \
            ++g_failures;                                                            \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);            \
        }                                                                            \
    } while (0)

#define CHECK_EQUAL(a, b)                                                            \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!((a) == (b))) {                                                         \
            ++g_failures;                                                            \
            std::ostringstream oss__;                                                \
            oss__ << (a) << " != " << (b);                                           \
            std::printf("FAILED %s:%d: %s == %s  [%s]\n", __FILE__, __LINE__, #a, #b, oss__.str().substr(0, 300).c_str()); \
        }                                                                            \
    } while (0)

#define CHECK_THROW(expr, exc)                                                       \
    do {                                                                             \
        ++g_checks;                                                                  \
        bool ok__ = false;                                                           \
        try {                                                                        \
            (void)(expr);                                                            \
        } catch (const exc &) {                                                      \
            ok__ = true;                                                             \
        } catch (const std::exception &e__) {                                        \
            std::printf("  (threw %s)\n", e__.what());                               \
        }                                                                            \
        if (!ok__) {                                                                 \
            ++g_failures;                                                            \
            std::printf("FAILED %s:%d: %s did not throw %s\n", __FILE__, __LINE__, #expr, #exc); \
        }                                                                            \
    } while (0)

#define CHECK_NO_THROW(expr)                                                         \
    do {                                                                             \
        ++g_checks;                                                                  \
        try {                                                                        \
            (void)(expr);                                                            \
        } catch (const std::exception &e__) {                                        \
            ++g_failures;                                                            \
            std::printf("FAILED %s:%d: %s threw %s\n", __FILE__, __LINE__, #expr, e__.what()); \
        }                                                                            \
    } while (0)

#define RUN_TEST(fn)                                                                 \
    do {                                                                             \
        const int before__ = g_failures;                                             \
        std::printf("[ RUN  ] %s\n", #fn);                                           \
        std::fflush(stdout);                                                         \
        try {                                                                        \
            fn();                                                                    \
        } catch (const std::exception &e__) {                                        \
            ++g_failures;                                                            \
            std::printf("FAILED %s: unexpected exception %s\n", #fn, e__.what());    \
        }                                                                            \
        std::printf("[ %s ] %s\n", g_failures == before__ ? " OK " : "FAIL", #fn);   \
    } while (0)

static inline int test_summary()
{
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures ? 1 : 0;
}
#endif
