from gm3d_b200.pointnet2_utils import *  # noqa: F401,F403
from gm3d_b200.pointnet2_utils import FurthestPointSampling, GatherOperation, furthest_point_sample, gather_operation  # noqa: F401
