// TEST INFRASTRUCTURE ONLY: stand-in for envire/maps/MLSGrid.hpp -- a dense multi-level surface grid with at most ONE
// surface patch per cell ({mean, stdev} as float, NaN mean = empty cell), which is the map definition of SURVEY.md §8c:
// toGrid = floor((x - offset) / scale) with the in-range test, get(position, patch, sigma) = the cell's patch if
// |patch.mean - mean| < sigma sqrt(patch.stdev^2 + stdev^2).
#ifndef ESLAM_REF_SHIM_ENVIRE_MLSGRID
#define ESLAM_REF_SHIM_ENVIRE_MLSGRID
#include <cmath>
#include <vector>
#include "../Core.hpp"
namespace envire {
class MultiLevelSurfaceGrid : public CartesianMap {
public:
    struct SurfacePatch {
        double mean, stdev, height;   // upstream: float; double here so that the query seed {z, sqrt(measVar)} is not rounded (the oracle's definition)
        long update_idx; bool horizontal;
        unsigned char color[3];
        SurfacePatch() : mean(0), stdev(0), height(0), update_idx(0), horizontal(true) { color[0] = color[1] = color[2] = 0; }
        SurfacePatch(double m, double s, double h = 0, bool hor = true) : mean(m), stdev(s), height(h), update_idx(0), horizontal(hor) { color[0] = color[1] = color[2] = 0; }
        void setHorizontal() { horizontal = true; }
        bool isHorizontal() const { return horizontal; }
        Eigen::Vector3d getColor() const { return Eigen::Vector3d(color[0] / 255.0, color[1] / 255.0, color[2] / 255.0); }
    };
    typedef const SurfacePatch* iterator;
    std::vector<float> cells;            // {mean, stdev} per cell, row major (y * width + x)
    std::vector<SurfacePatch> tmp;       // beginCell's view of one cell
    size_t width, height;
    double scalex, scaley, offsetx, offsety;
    MultiLevelSurfaceGrid(size_t w, size_t h, double sx, double sy, double ox = 0, double oy = 0)
        : cells(w * h * 2, std::nanf("")), tmp(1), width(w), height(h), scalex(sx), scaley(sy), offsetx(ox), offsety(oy) {}
    size_t getWidth() const { return width; } size_t getHeight() const { return height; }
    double getScaleX() const { return scalex; } double getScaleY() const { return scaley; }
    bool toGrid(double x, double y, size_t& m, size_t& n) const {
        const double qx = std::floor((x - offsetx) / scalex), qy = std::floor((y - offsety) / scaley);
        if (!(qx >= 0.0 && qx < (double)width)) return false;
        if (!(qy >= 0.0 && qy < (double)height)) return false;
        m = (size_t)qx; n = (size_t)qy; return true;
    }
    void fromGrid(size_t m, size_t n, double& x, double& y) const { x = (m + 0.5) * scalex + offsetx; y = (n + 0.5) * scaley + offsety; }
    iterator beginCell(size_t m, size_t n) {
        const float* c = &cells[2 * (n * width + m)];
        if (std::isnan(c[0]) || std::isnan(c[1])) return 0;
        tmp[0] = SurfacePatch(c[0], c[1]); return &tmp[0];
    }
    iterator endCell() { return 0; }
    // MLSGrid::get(position, patch, sigma_threshold): the patch of the cell that lies within sigma_threshold of the query
    bool get(const Eigen::Vector3d& position, SurfacePatch& patch, double sigma_threshold = 3.0) {
        size_t m, n;
        if (!toGrid(position.x(), position.y(), m, n)) return false;
        last_cell = (long)(n * width + m);
        const float* c = &cells[2 * (n * width + m)];
        if (std::isnan(c[0]) || std::isnan(c[1])) return false;
        const double interval = std::sqrt((double)patch.stdev * (double)patch.stdev + (double)c[1] * (double)c[1]) * sigma_threshold;
        if (!(std::fabs((double)patch.mean - (double)c[0]) < interval)) return false;
        patch.mean = c[0]; patch.stdev = c[1]; patch.height = 0; patch.horizontal = true;
        return true;
    }
    long last_cell;
};
typedef MultiLevelSurfaceGrid MLSGrid;
}
#endif
