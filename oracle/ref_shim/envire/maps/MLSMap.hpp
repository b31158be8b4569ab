// TEST INFRASTRUCTURE ONLY: stand-in for envire/maps/MLSMap.hpp: a list of MLS grids, newest first in getPatch
#ifndef ESLAM_REF_SHIM_ENVIRE_MLSMAP
#define ESLAM_REF_SHIM_ENVIRE_MLSMAP
#include "MLSGrid.hpp"
namespace envire {
class MLSMap : public CartesianMap {
public:
    typedef boost::intrusive_ptr<MLSMap> Ptr;
    std::vector<boost::intrusive_ptr<MLSGrid> > grids;
    bool getPatch(const Eigen::Vector3d& position, MLSGrid::SurfacePatch& patch, double sigma_threshold = 3.0) {
        for (size_t i = grids.size(); i-- > 0;)
            if (grids[i]->get(position, patch, sigma_threshold)) return true;
        return false;
    }
    MLSMap* cloneDeep() const {
        MLSMap* m = new MLSMap();
        for (size_t i = 0; i < grids.size(); ++i) m->grids.push_back(boost::intrusive_ptr<MLSGrid>(new MLSGrid(*grids[i])));
        return m;
    }
};
}
#endif
