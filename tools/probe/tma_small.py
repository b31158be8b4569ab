import sys; sys.path.insert(0,'/root/repo')
from slam_eslam_b200 import GpuPoseEstimator, abi, synthetic
g = GpuPoseEstimator(abi.default_config(min_effective=251))
contacts, step = synthetic.setup_filter(g, "C1")
g.project(step)
print(g.update(contacts).as_dict())
