// Minimal stand-in for <boost/test/included/unit_test.hpp> (Boost is not installed in this image).
//
// TEST INFRASTRUCTURE.  It exists so that the reference's own test sources
// (libs/bayesian/test/{belief_propagation,cpt,graph,matrix}.cpp) compile UNMODIFIED against the
// drop-in headers under include/bayesian/ -- the strongest drop-in proof available without Boost.
// Implements exactly what those files use: BOOST_TEST_MAIN, BOOST_AUTO_TEST_CASE, BOOST_CHECK,
// BOOST_CHECK_EQUAL, BOOST_CHECK_CLOSE (Boost's "strong" percent tolerance: the difference must be
// within tolerance of BOTH values; two exact zeros compare equal).
#ifndef BNB200_BOOST_TEST_SHIM_HPP
#define BNB200_BOOST_TEST_SHIM_HPP

#include <cmath>
#include <cstdio>
#include <exception>
#include <vector>

namespace boost_shim {

struct test_case {
    char const* name;
    void (*body)();
};

inline std::vector<test_case>& registry()
{
    static std::vector<test_case> all;
    return all;
}

inline int& failures()
{
    static int n = 0;
    return n;
}

inline int& checks()
{
    static int n = 0;
    return n;
}

struct registrar {
    registrar(char const* name, void (*body)()) { registry().push_back(test_case{name, body}); }
};

inline void report(bool ok, char const* file, int line, char const* what)
{
    ++checks();
    if (!ok) {
        ++failures();
        std::printf("%s(%d): error: check %s has failed\n", file, line, what);
    }
}

inline bool close_percent(double a, double b, double tol_percent)
{
    if (a == b) return true;
    double const d = std::fabs(a - b);
    double const frac = tol_percent / 100.0;
    return d <= frac * std::fabs(a) && d <= frac * std::fabs(b);
}

inline int run_all()
{
    for (test_case const& t : registry()) {
        int const before = failures();
        try {
            t.body();
        } catch (std::exception const& e) {
            ++failures();
            std::printf("%s: error: uncaught exception: %s\n", t.name, e.what());
        } catch (...) {
            ++failures();
            std::printf("%s: error: uncaught exception\n", t.name);
        }
        std::printf("[%s] %s\n", failures() == before ? "  ok  " : "FAILED", t.name);
    }
    if (failures() == 0)
        std::printf("\n*** No errors detected (%d test cases, %d checks)\n", (int)registry().size(), checks());
    else
        std::printf("\n*** %d failure(s) detected in %d test cases\n", failures(), (int)registry().size());
    return failures() == 0 ? 0 : 201;
}

} // namespace boost_shim

#define BOOST_AUTO_TEST_CASE(test_name)                                                          \
    static void test_name();                                                                     \
    static ::boost_shim::registrar test_name##_registrar(#test_name, &test_name);                \
    static void test_name()

#define BOOST_CHECK(expr) ::boost_shim::report(static_cast<bool>(expr), __FILE__, __LINE__, #expr)
#define BOOST_CHECK_EQUAL(a, b) ::boost_shim::report((a) == (b), __FILE__, __LINE__, #a " == " #b)
#define BOOST_CHECK_CLOSE(a, b, tol)                                                              \
    ::boost_shim::report(::boost_shim::close_percent((a), (b), (tol)), __FILE__, __LINE__,       \
                         "|" #a " - " #b "| within " #tol " %")

#define BOOST_CHECK_THROW(stmt, exception_type)                                                   \
    do {                                                                                          \
        bool caught__ = false;                                                                    \
        try { (void)(stmt); } catch (exception_type const&) { caught__ = true; }                  \
        ::boost_shim::report(caught__, __FILE__, __LINE__, #stmt " throws " #exception_type);     \
    } while (0)

#ifdef BOOST_TEST_MAIN
int main()
{
    return ::boost_shim::run_all();
}
#endif

#endif // BNB200_BOOST_TEST_SHIM_HPP
