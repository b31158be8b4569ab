// TEST INFRASTRUCTURE ONLY: stand-in for boost/test/included/unit_test.hpp -- enough of Boost.Test to run the
// reference's own test/testContactModel.cpp: auto-registered cases, CHECK / REQUIRE / CHECK_EQUAL / CHECK_SMALL /
// CHECK_CLOSE (tolerance in PERCENT, as Boost defines it) / FAIL, one line per failed assertion and a per-case summary:
//   CASE <name> checks=<n> failures=<m>
// The process exits with the number of failed cases.
#ifndef ESLAM_REF_SHIM_BOOST_TEST
#define ESLAM_REF_SHIM_BOOST_TEST
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <boost/bind.hpp>
namespace eslam_ref_shim { namespace ut {
struct Case { const char* name; void (*fn)(); };
inline std::vector<Case>& cases() { static std::vector<Case> c; return c; }
struct Reg { Reg(const char* n, void (*f)()) { Case c = {n, f}; cases().push_back(c); } };
struct Abort {};
inline int& checks() { static int n = 0; return n; }
inline int& failures() { static int n = 0; return n; }
inline void report(bool ok, const char* what, const char* file, int line, const std::string& detail = std::string()) {
    ++checks();
    if (!ok) { ++failures(); std::printf("  FAILED %s:%d %s %s\n", file, line, what, detail.c_str()); }
}
template <class A, class B> inline std::string show(const A& a, const B& b) { std::ostringstream s; s << "[" << a << " != " << b << "]"; return s.str(); }
inline bool close_pct(double a, double b, double pct) {
    const double d = std::fabs(a - b);
    if (d == 0) return true;                       // also inf == inf
    return d <= std::fabs(a) * pct / 100.0 && d <= std::fabs(b) * pct / 100.0;
}
}}
#define BOOST_AUTO_TEST_CASE(name) \
    static void name(); static eslam_ref_shim::ut::Reg reg_##name(#name, &name); static void name()
#define BOOST_CHECK(e) eslam_ref_shim::ut::report(!!(e), #e, __FILE__, __LINE__)
#define BOOST_REQUIRE(e) do { const bool ok_ = !!(e); eslam_ref_shim::ut::report(ok_, #e, __FILE__, __LINE__); if (!ok_) throw eslam_ref_shim::ut::Abort(); } while (0)
#define BOOST_CHECK_EQUAL(a, b) eslam_ref_shim::ut::report((a) == (b), #a " == " #b, __FILE__, __LINE__, eslam_ref_shim::ut::show(a, b))
#define BOOST_CHECK_SMALL(v, tol) eslam_ref_shim::ut::report(std::fabs(v) < (tol), "|" #v "| < " #tol, __FILE__, __LINE__, eslam_ref_shim::ut::show(v, tol))
#define BOOST_CHECK_CLOSE(a, b, pct) eslam_ref_shim::ut::report(eslam_ref_shim::ut::close_pct(a, b, pct), #a " ~ " #b, __FILE__, __LINE__, eslam_ref_shim::ut::show(a, b))
#define BOOST_FAIL(msg) do { eslam_ref_shim::ut::report(false, msg, __FILE__, __LINE__); throw eslam_ref_shim::ut::Abort(); } while (0)
int main() {
    using namespace eslam_ref_shim::ut;
    int failed_cases = 0;
    for (size_t i = 0; i < cases().size(); ++i) {
        checks() = 0; failures() = 0;
        try { cases()[i].fn(); } catch (Abort&) {} catch (std::exception& e) { report(false, e.what(), "exception", 0); }
        std::printf("CASE %s checks=%d failures=%d\n", cases()[i].name, checks(), failures());
        if (failures()) ++failed_cases;
    }
    return failed_cases;
}
#endif
