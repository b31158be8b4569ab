"""Import shim: the package sources live in ``slam-eslam_b200/`` (the name the build contract asks for,
which is not a valid Python identifier).  ``import slam_eslam_b200`` resolves submodules from there."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__.append(_os.path.join(_os.path.dirname(_here), "slam-eslam_b200"))

from .abi import *  # noqa: F401,F403
from .filter import GpuPoseEstimator  # noqa: F401
