// Configuration.hpp -- the reference's configuration structs with the reference's field names and defaults
// (liyangSKD/slam-eslam src/Configuration.hpp:12-213), so code that fills an eslam::Configuration compiles unchanged.
#ifndef ESLAM_FACADE_CONFIGURATION_HPP
#define ESLAM_FACADE_CONFIGURATION_HPP
#include <cmath>
#include <cstddef>
#include "compat.hpp"

namespace eslam {

struct UpdateThreshold {                        // src/Configuration.hpp:12-30
    UpdateThreshold() : distance(0), angle(0) {}
    UpdateThreshold(double distance, double angle) : distance(distance), angle(angle) {}
    bool test(double distance, double angle) { return distance > this->distance || angle > this->angle; }
    // the reference passes (angle, distance) into the (distance, angle) overload (:23-26); kept as is
    bool test(const base::Affine3d& pdelta) { return test(pdelta.rotationAngle(), pdelta.translation().norm()); }
    double distance, angle;
};

struct SurfaceHashConfig {                      // :32-49
    SurfaceHashConfig() : useHash(false), period(10), percentage(0.05), avgFactor(0.1), slopeBins(20), angularSteps(16) {}
    bool useHash; size_t period; double percentage, avgFactor; size_t slopeBins, angularSteps;
};

struct ContactModelConfiguration {              // :51-81
    ContactModelConfiguration()
        : useSlipUpdate(false), useShapeUpdate(true), minContacts(3), contactLikelihoodCorrection(0.33),
          contactPointRadius(0.01) {}
    bool useSlipUpdate, useShapeUpdate; size_t minContacts; double contactLikelihoodCorrection, contactPointRadius;
};

struct Configuration {                          // :83-213
    Configuration()
        : seed(42u), particleCount(250), minEffective(50), initialRotationError(0, 0, 0.1),
          initialTranslationError(0.1, 0.1, 1.0), measurementError(0.1), discountFactor(0.9), spreadThreshold(0.9),
          spreadTranslationFactor(0.1), spreadRotationFactor(0.05), slipFactor(0.05), maxYawDeviation(15 * M_PI / 180.0),
          measurementThreshold(0.1, 10 * M_PI / 180.0), mappingThreshold(0.02, 5 * M_PI / 180.0),
          mappingCameraThreshold(1.0, 30 * M_PI / 180.0), gridSize(20.0), gridResolution(0.05), gridThreshold(0.5),
          gridPatchThickness(0.1), gridGapSize(1.5), gridUseNegativeInformation(false), maxSensorRange(3.0),
          useVisualUpdate(false), logDebug(false), logParticlePeriod(100) {}
    unsigned long seed; size_t particleCount, minEffective;
    base::Vector3d initialRotationError, initialTranslationError;
    double measurementError, discountFactor, spreadThreshold, spreadTranslationFactor, spreadRotationFactor, slipFactor,
        maxYawDeviation;
    UpdateThreshold measurementThreshold, mappingThreshold, mappingCameraThreshold;
    double gridSize, gridResolution, gridThreshold, gridPatchThickness, gridGapSize;
    bool gridUseNegativeInformation; double maxSensorRange; bool useVisualUpdate;
    ContactModelConfiguration contactModel;
    bool logDebug; unsigned int logParticlePeriod;
};

}  // namespace eslam
#endif
