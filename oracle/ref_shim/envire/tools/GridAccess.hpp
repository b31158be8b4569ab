// TEST INFRASTRUCTURE ONLY: envire/tools/GridAccess.hpp is included by src/PoseEstimator.cpp:3 but nothing of it is used
