// TEST INFRASTRUCTURE ONLY: stand-in for boost/random/normal_distribution.hpp (Box-Muller, as boost < 1.56)
#ifndef ESLAM_REF_SHIM_BOOST_NORMAL_DIST
#define ESLAM_REF_SHIM_BOOST_NORMAL_DIST
#include <cmath>
#include "uniform_real.hpp"
namespace boost {
template <class T = double> struct normal_distribution {
    T mean, sigma;
    normal_distribution(T m = 0, T s = 1) : mean(m), sigma(s) {}
    enum { kind = eslam_ref_shim::DRAW_NORMAL };
    template <class E> T sample(E& e) const {
        uniform_real<T> u(0, 1);
        const T r1 = u.sample(e), r2 = u.sample(e);
        return mean + sigma * std::sqrt(-2.0 * std::log(1.0 - r2)) * std::cos(2.0 * M_PI * r1);
    }
};
}
#endif
