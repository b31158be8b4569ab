// TEST INFRASTRUCTURE ONLY: stand-in for boost/random/variate_generator.hpp (see linear_congruential.hpp for the hook)
#ifndef ESLAM_REF_SHIM_BOOST_VARIATE
#define ESLAM_REF_SHIM_BOOST_VARIATE
#include "linear_congruential.hpp"
namespace boost {
template <class Engine, class Dist> struct variate_generator;
template <class E, class Dist> struct variate_generator<E&, Dist> {
    E& eng; Dist dist;
    variate_generator(E& e, const Dist& d) : eng(e), dist(d) {}
    double operator()() {
        if (eslam_ref_shim::noise_hook()) return eslam_ref_shim::noise_hook()((int)Dist::kind, this);
        return dist.sample(eng);
    }
};
}
#endif
