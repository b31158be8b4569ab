// TEST INFRASTRUCTURE ONLY: stand-in for envire/tools/GaussianMixture.hpp (only named by PoseDistribution)
#ifndef ESLAM_REF_SHIM_ENVIRE_GMM
#define ESLAM_REF_SHIM_ENVIRE_GMM
#include <vector>
namespace envire {
template <class S, int D, class A> struct GaussianMixture {
    struct Parameter { S weight; typename A::Vector mean; typename A::Matrix cov; };
    std::vector<Parameter> params;
};
}
#endif
