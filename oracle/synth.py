"""Input generator of the parity tests: lives in polychase_b200/synth.py (bench.py uses it too and must not
import the checker); re-exported here for the oracle modules and tests that say `oracle.synth`."""
from polychase_b200.synth import *  # noqa: F401,F403
from polychase_b200.synth import Clip, camera_path, homography, intrinsics, make_texture, plane_mesh, plane_scale, warp_frame  # noqa: F401
