// TEST INFRASTRUCTURE ONLY: boost::function -> std::function
#ifndef ESLAM_REF_SHIM_BOOST_FUNCTION
#define ESLAM_REF_SHIM_BOOST_FUNCTION
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>
#include <algorithm>
namespace boost {
template <class S> using function = std::function<S>;
using std::shared_ptr;
}
#endif
