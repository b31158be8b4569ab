// TEST INFRASTRUCTURE ONLY (oracle/ref_shim/README.md): stand-in for boost/random/linear_congruential.hpp.
// The engine keeps boost::minstd_rand's recurrence (x <- 48271 x mod 2^31-1), but the variate generators of this shim
// (variate_generator.hpp) take their numbers from a pluggable noise hook when one is installed: north_star compares the
// filters on IDENTICAL noise streams (Philox counters), so oracle/ref/ref_driver.cpp feeds the reference's own draw
// sites from the same counters the oracle and the GPU use.
#ifndef ESLAM_REF_SHIM_BOOST_LCG
#define ESLAM_REF_SHIM_BOOST_LCG
#include <stdint.h>
#include <cassert>            // boost/assert.hpp is part of what the real header pulls in
namespace boost {
struct minstd_rand {
    uint32_t x;
    explicit minstd_rand(unsigned long seed = 1u) { x = (uint32_t)(seed % 2147483647ul); if (x == 0) x = 1; }
    uint32_t operator()() { x = (uint32_t)(((uint64_t)x * 48271u) % 2147483647u); return x; }
    static uint32_t min() { return 1; }
    static uint32_t max() { return 2147483646u; }
};
}
namespace eslam_ref_shim {
enum { DRAW_UNIFORM = 0, DRAW_NORMAL = 1 };
// hook(kind, site) -> the next number of that kind; site identifies the generator object's address (the reference has
// three: PoseEstimator::rand_norm, ::rand_uni, and the local `rand` of resample_*).  NULL: the plain engine is used.
typedef double (*noise_hook_t)(int kind, const void* site);
inline noise_hook_t& noise_hook() { static noise_hook_t h = 0; return h; }
}
#endif
