// TEST INFRASTRUCTURE ONLY: stand-in for terrain_estimator/TerrainConfiguration.hpp.  The slip path is inert in the
// reference; only the names ContactModel::matchTerrain mentions exist.
#ifndef ESLAM_REF_SHIM_TERRAIN
#define ESLAM_REF_SHIM_TERRAIN
#include <base/Eigen.hpp>
namespace terrain_estimator {
struct TerrainClassification {
    int wheel_idx;
    double p[3];
    TerrainClassification() : wheel_idx(0) { p[0] = p[1] = p[2] = 1.0 / 3; }
    static TerrainClassification fromRGB(const Eigen::Vector3d& c) { TerrainClassification t; t.p[0] = c.x(); t.p[1] = c.y(); t.p[2] = c.z(); return t; }
    Eigen::Vector3d toRGB() const { return Eigen::Vector3d(p[0], p[1], p[2]); }
    double jointProbability(const TerrainClassification& o) const { return p[0] * o.p[0] + p[1] * o.p[1] + p[2] * o.p[2]; }
};
}
#endif
