// TEST INFRASTRUCTURE ONLY: stand-in for boost/random/uniform_real.hpp
#ifndef ESLAM_REF_SHIM_BOOST_UNIFORM_REAL
#define ESLAM_REF_SHIM_BOOST_UNIFORM_REAL
#include "linear_congruential.hpp"
namespace boost {
template <class T = double> struct uniform_real {
    T lo, hi;
    uniform_real(T a = 0, T b = 1) : lo(a), hi(b) {}
    enum { kind = eslam_ref_shim::DRAW_UNIFORM };
    template <class E> T sample(E& e) const { return lo + (hi - lo) * ((T)(e() - E::min()) / ((T)(E::max() - E::min()) + 1)); }
};
}
#endif
