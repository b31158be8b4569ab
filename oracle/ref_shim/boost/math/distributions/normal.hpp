// TEST INFRASTRUCTURE ONLY: stand-in for boost/math/distributions/normal.hpp.  pdf and cdf follow boost's own
// expressions (exponent = -(x-mean)^2 / (2 sd^2); exp(exponent) / (sd sqrt(2 pi)); erfc(-(x-mean)/(sd sqrt 2)) / 2)
// on libm's exp / erfc.
#ifndef ESLAM_REF_SHIM_BOOST_MATH_NORMAL
#define ESLAM_REF_SHIM_BOOST_MATH_NORMAL
#include <cmath>
namespace boost { namespace math {
class normal {
    double m, s;
public:
    normal(double mean = 0, double sd = 1) : m(mean), s(sd) {}
    double mean() const { return m; }
    double standard_deviation() const { return s; }
};
inline double pdf(const normal& n, double x) {
    const double sd = n.standard_deviation();
    double exponent = x - n.mean();
    exponent *= -exponent;
    exponent /= 2 * sd * sd;
    double result = std::exp(exponent);
    result /= sd * std::sqrt(2 * M_PI);
    return result;
}
inline double cdf(const normal& n, double x) {
    const double diff = (x - n.mean()) / (n.standard_deviation() * std::sqrt(2.0));
    return std::erfc(-diff) / 2;
}
}}
#endif
