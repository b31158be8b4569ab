// SurfaceHash.hpp -- eslam::SurfaceHash (src/SurfaceHash.hpp:113-232) over libeslam_gpu.so: the slope-indexed table of candidate
// poses.  The table itself lives on the device of the filter that uses it (eslam_gpu_hash_create: built from the filter's map when
// PoseEstimator::init(n, hash) runs); this object carries the configuration and the surface parameters of a point set.
#ifndef ESLAM_FACADE_SURFACEHASH_HPP
#define ESLAM_FACADE_SURFACEHASH_HPP
#include <vector>
#include "Configuration.hpp"
#include "PoseParticle.hpp"

namespace eslam {

struct SurfaceParam {                                   // :52-111
    double slope_x, slope_y;
    SurfaceParam() : slope_x(0), slope_y(0) {}
    // least-squares plane z = slope_x x + slope_y y + c through the points (:60-110), by the normal equations
    void fromPoints(const std::vector<base::Vector3d>& points) {
        double x = 0, y = 0, z = 0, xx = 0, yy = 0, xy = 0, xz = 0, yz = 0;
        for (size_t i = 0; i < points.size(); i++) {
            const base::Vector3d& p = points[i];
            x += p.x(); y += p.y(); z += p.z(); xx += p.x() * p.x(); yy += p.y() * p.y(); xy += p.x() * p.y(); xz += p.x() * p.z(); yz += p.y() * p.z();
        }
        const double n = (double)points.size();
        const double det = xx * (yy * n - y * y) - xy * (xy * n - y * x) + x * (xy * y - yy * x);
        slope_x = (xz * (yy * n - y * y) - xy * (yz * n - y * z) + x * (yz * y - yy * z)) / det;
        slope_y = (xx * (yz * n - y * z) - xz * (xy * n - y * x) + x * (xy * z - yz * x)) / det;
    }
};

struct SurfaceHash {
    SurfaceHashConfig config;
    bool created;
    SurfaceHash() : created(false) {}
    void setConfiguration(const SurfaceHashConfig& c) { config = c; }                       // :122-125
    // create(gridTemplate) (:155-231): the table is built on the device from the map the filter holds; here only marked as wanted
    void create() { created = true; }
};

}  // namespace eslam
#endif
