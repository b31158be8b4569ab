// TEST INFRASTRUCTURE ONLY: stand-in for base/Time.hpp
#ifndef ESLAM_REF_SHIM_BASE_TIME
#define ESLAM_REF_SHIM_BASE_TIME
#include <stdint.h>
namespace base { struct Time { int64_t microseconds; Time() : microseconds(0) {} static Time now() { return Time(); } }; }
#endif
