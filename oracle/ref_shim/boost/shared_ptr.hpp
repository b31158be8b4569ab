// TEST INFRASTRUCTURE ONLY
#include "function.hpp"
