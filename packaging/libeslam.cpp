// libeslam.cpp -- the translation unit behind `-leslam` (the reference's pkg-config entry, src/eslam.pc.in:6-10, names that
// library): the class facade of include/eslam/*.hpp is header-only over the C ABI of libeslam_gpu.so, so libeslam.so carries
// one instantiation of every facade class (an ODR anchor and a link-time check that the headers are self-contained) and the
// version query; it pulls libeslam_gpu.so in as its dependency.
#include <eslam/EmbodiedSlamFilter.hpp>

namespace eslam {
// one out-of-line use of every class: a build of libeslam.so fails if a facade header stops compiling on its own
static void anchor() {
    odometry::Configuration oc; Configuration c;
    EmbodiedSlamFilter* f = 0; PoseEstimator* p = 0; ContactModel* m = 0; PoseDistribution* d = 0;
    (void)oc; (void)c; (void)f; (void)p; (void)m; (void)d;
}
}  // namespace eslam

extern "C" const char* eslam_version() { (void)&eslam::anchor; return "2.0.0-b200"; }
