// TEST INFRASTRUCTURE ONLY: boost::bind -> std::bind, placeholders _1.. in the global namespace as boost/bind.hpp has them
#ifndef ESLAM_REF_SHIM_BOOST_BIND
#define ESLAM_REF_SHIM_BOOST_BIND
#include <functional>
namespace boost { using std::bind; }
using namespace std::placeholders;
#endif
