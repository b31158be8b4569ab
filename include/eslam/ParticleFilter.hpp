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

    explicit ParticleFilter(const eslam_gpu_config& cfg, int device = 0) : gpu(NULL), gpuConfig(cfg), gpuDevice(device), hostValid(false), hostBorrowed(false) {
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
    void resample_stratified(size_t samples) { flush(); gpu_check(gpu, eslam_gpu_resample_n(gpu, samples, 0)); hostValid = false; }    // :85-108
    void resample_multinomial(size_t samples) { flush(); gpu_check(gpu, eslam_gpu_resample_n(gpu, samples, 1)); hostValid = false; }   // :120-147
    size_t getBestParticleIndex() { flush(); uint64_t i; gpu_check(gpu, eslam_gpu_best_index(gpu, &i)); return (size_t)i; }     // :160-173
    size_t size() const { return (size_t)eslam_gpu_particle_count(gpu); }

    std::vector<Particle>& getParticles() {              // :150-153 mutable reference: re-uploaded before the next device op
        sync(); hostBorrowed = true; return xi_k;
    }
    const std::vector<Particle>& getParticles() const { const_cast<ParticleFilter*>(this)->sync(); return xi_k; }

    eslam_gpu_ctx* context() { return gpu; }

protected:
    void sync() {                                        // device -> host mirror (staging arrays are kept between calls)
        if (hostValid) return;
        const size_t n = size();
        xi_k.resize(n);
        if (n == 0) { hostValid = true; return; }
        for (int f = 0; f < 7; f++) stage[f].resize(n);
        stageFlag[0].resize(n); stageFlag[1].resize(n);
        gpu_check(gpu, eslam_gpu_get_particles(gpu, stage[0].data(), stage[1].data(), stage[2].data(), stage[3].data(), stage[4].data(),
                                               stage[5].data(), stage[6].data(), stageFlag[0].data(), stageFlag[1].data()));
        for (size_t i = 0; i < n; i++) {
            Particle& p = xi_k[i];
            p.position = base::Vector2d(stage[0][i], stage[1][i]); p.orientation = stage[2][i]; p.zPos = stage[3][i]; p.zSigma = stage[4][i];
            p.weight = stage[5][i]; p.mprob = stage[6][i]; p.floating = stageFlag[0][i] != 0; p.n_contacts = stageFlag[1][i];
        }
        hostValid = true;
    }
    void flush() {                                       // host mirror -> device, if a mutable reference was handed out
        if (!hostBorrowed) return;
        const size_t n = xi_k.size();
        for (int f = 0; f < 6; f++) stage[f].resize(n);
        for (size_t i = 0; i < n; i++) {
            const Particle& p = xi_k[i];
            stage[0][i] = p.position.x(); stage[1][i] = p.position.y(); stage[2][i] = p.orientation; stage[3][i] = p.zPos;
            stage[4][i] = p.zSigma; stage[5][i] = p.weight;
        }
        if (n) gpu_check(gpu, eslam_gpu_set_particles(gpu, n, stage[0].data(), stage[1].data(), stage[2].data(), stage[3].data(),
                                                      stage[4].data(), stage[5].data()));
        hostBorrowed = false;
    }
    // a new context with another configuration (per-particle maps on / off): the particle set moves over
    void recreate(const eslam_gpu_config& cfg) {
        sync();
        eslam_gpu_ctx* fresh = NULL;
        if (eslam_gpu_create(&cfg, gpuDevice, NULL, &fresh) != ESLAM_OK) throw std::runtime_error("eslam_gpu_create failed");
        eslam_gpu_destroy(gpu);
        gpu = fresh; gpuConfig = cfg;
        hostBorrowed = !xi_k.empty();                    // re-uploaded before the next device operation
        flush();
    }
    void deviceChanged() { hostValid = false; hostBorrowed = false; }

    eslam_gpu_ctx* gpu;
    eslam_gpu_config gpuConfig; int gpuDevice;
    std::vector<Particle> xi_k;                          // :176 (host mirror)
    std::vector<double> stage[7]; std::vector<uint8_t> stageFlag[2];
    bool hostValid, hostBorrowed;
};

}  // namespace eslam
#endif
