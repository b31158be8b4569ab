// ParticleFilter.hpp -- eslam::ParticleFilter<PoseParticle> (src/ParticleFilter.hpp:18-178) over libeslam_gpu.so.
// The particle set lives on the device; xi_k is a lazily synchronised host mirror: getParticles() downloads it after a
// device step and, because the reference hands out a mutable reference (:150-153), re-uploads it before the next device
// operation.  Nothing here computes on the host: sums, normalisation, resampling and argmax are device calls.
#ifndef ESLAM_FACADE_PARTICLEFILTER_HPP
#define ESLAM_FACADE_PARTICLEFILTER_HPP
#include <stdexcept>
#include <string>
#include <vector>
#include "../eslam_gpu.h"
#include "PoseParticle.hpp"

namespace eslam {

inline void gpu_check(eslam_gpu_ctx* ctx, int status) {
    switch (status) {
        case ESLAM_OK: return;
        case ESLAM_ERR_NO_ENVIRONMENT: throw std::runtime_error("No environment attached.");                 // src/PoseEstimator.cpp:260
        case ESLAM_ERR_ZERO_VARIANCE: throw std::runtime_error("using a zero measurement variance leads to singularities");  // src/ContactModel.cpp:123
        default: throw std::runtime_error(std::string("eslam_gpu: ") + (ctx ? eslam_gpu_last_error(ctx) : "error") +
                                          " (status " + std::to_string(status) + ")");
    }
}

template <class _Particle> class ParticleFilter;        // only the PoseParticle instance exists on the device

template <>
class ParticleFilter<PoseParticle> {
public:
    typedef PoseParticle Particle;

    explicit ParticleFilter(const eslam_gpu_config& cfg, int device = 0) : gpu(NULL), hostValid(false), hostBorrowed(false) {
        if (eslam_gpu_create(&cfg, device, NULL, &gpu) != ESLAM_OK)
            throw std::runtime_error("eslam_gpu_create failed: no CUDA device (there is no CPU fallback)");
    }
    virtual ~ParticleFilter() { eslam_gpu_destroy(gpu); }

    double getWeightsSum() {                             // :34-39 (device sum; leaves the weights unchanged)
        flush(); double sum; gpu_check(gpu, eslam_gpu_weights_sum(gpu, &sum)); return sum;
    }
    double getWeightAvg() { return getWeightsSum() / size(); }                    // :41-44
    double normalizeWeights() { flush(); double ess; gpu_check(gpu, eslam_gpu_normalize_weights(gpu, &ess)); hostValid = false; return ess; }  // :46-70
    void resample() { flush(); gpu_check(gpu, eslam_gpu_resample(gpu)); hostValid = false; }                                    // :72-75
    size_t getBestParticleIndex() { flush(); uint64_t i; gpu_check(gpu, eslam_gpu_best_index(gpu, &i)); return (size_t)i; }     // :160-173
    size_t size() const { return (size_t)eslam_gpu_particle_count(gpu); }

    std::vector<Particle>& getParticles() {              // :150-153 mutable reference: re-uploaded before the next device op
        sync(); hostBorrowed = true; return xi_k;
    }
    const std::vector<Particle>& getParticles() const { const_cast<ParticleFilter*>(this)->sync(); return xi_k; }

    eslam_gpu_ctx* context() { return gpu; }

protected:
    void sync() {                                        // device -> host mirror
        if (hostValid) return;
        const size_t n = size();
        xi_k.resize(n);
        if (n == 0) { hostValid = true; return; }
        std::vector<double> x(n), y(n), th(n), z(n), zs(n), w(n), mp(n); std::vector<uint8_t> fl(n), nf(n);
        gpu_check(gpu, eslam_gpu_get_particles(gpu, x.data(), y.data(), th.data(), z.data(), zs.data(), w.data(), mp.data(), fl.data(), nf.data()));
        for (size_t i = 0; i < n; i++) {
            Particle& p = xi_k[i];
            p.position = base::Vector2d(x[i], y[i]); p.orientation = th[i]; p.zPos = z[i]; p.zSigma = zs[i];
            p.weight = w[i]; p.mprob = mp[i]; p.floating = fl[i] != 0; p.n_contacts = nf[i];
        }
        hostValid = true;
    }
    void flush() {                                       // host mirror -> device, if a mutable reference was handed out
        if (!hostBorrowed) return;
        const size_t n = xi_k.size();
        std::vector<double> x(n), y(n), th(n), z(n), zs(n), w(n);
        for (size_t i = 0; i < n; i++) {
            const Particle& p = xi_k[i];
            x[i] = p.position.x(); y[i] = p.position.y(); th[i] = p.orientation; z[i] = p.zPos; zs[i] = p.zSigma; w[i] = p.weight;
        }
        if (n) gpu_check(gpu, eslam_gpu_set_particles(gpu, n, x.data(), y.data(), th.data(), z.data(), zs.data(), w.data()));
        hostBorrowed = false;
    }
    void deviceChanged() { hostValid = false; hostBorrowed = false; }

    eslam_gpu_ctx* gpu;
    std::vector<Particle> xi_k;                          // :176 (host mirror)
    bool hostValid, hostBorrowed;
};

}  // namespace eslam
#endif
