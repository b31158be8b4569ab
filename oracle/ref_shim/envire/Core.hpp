// TEST INFRASTRUCTURE ONLY: stand-in for envire/Core.hpp -- the scene-graph calls eslam::GridAccess makes
// (src/PoseEstimator.hpp:36-95): one environment, frame nodes with a transform to the root, layers that can be detached.
#ifndef ESLAM_REF_SHIM_ENVIRE_CORE
#define ESLAM_REF_SHIM_ENVIRE_CORE
#include <list>
#include <base/Eigen.hpp>
#include <boost/intrusive_ptr.hpp>
namespace envire {
class Environment;
class CartesianMap;
class EnvironmentItem {
public:
    long ref_count; Environment* env;
    EnvironmentItem() : ref_count(0), env(0) {}
    EnvironmentItem(const EnvironmentItem&) : ref_count(0), env(0) {}
    virtual ~EnvironmentItem() {}
    bool isAttached() const { return env != 0; }
    Environment* getEnvironment() const { return env; }
    void detach() { env = 0; }
};
inline void intrusive_ptr_add_ref(EnvironmentItem* i) { ++i->ref_count; }
inline void intrusive_ptr_release(EnvironmentItem* i) { if (--i->ref_count == 0) delete i; }
class FrameNode : public EnvironmentItem {
public:
    Eigen::Affine3d toRoot;              // this frame -> root
    FrameNode* parent;
    std::list<CartesianMap*> maps;
    FrameNode() : parent(0) {}
    FrameNode* getParent() { return parent; }
    std::list<CartesianMap*> getMaps() { return maps; }
    Eigen::Affine3d relativeTransform(const FrameNode* to) const { return to->toRoot.inverse() * toRoot; }
};
class Layer : public EnvironmentItem {
public:
    std::list<Layer*> getParents() { return std::list<Layer*>(); }
};
class CartesianMap : public Layer {
public:
    FrameNode* frame;
    CartesianMap() : frame(0) {}
    FrameNode* getFrameNode() const { return frame; }
};
class Environment {
public:
    FrameNode root;
    Environment() { root.env = this; }
    FrameNode* getRootNode() { return &root; }
    Eigen::Affine3d relativeTransform(const FrameNode* from, const FrameNode* to) const { return from->relativeTransform(to); }
    std::list<Layer*> getChildren(Layer*) { return std::list<Layer*>(); }
    void setFrameNode(CartesianMap* m, FrameNode* fn) { m->frame = fn; m->env = this; }
    void attachItem(EnvironmentItem* i) { i->env = this; }
};
}
#endif
