"""Import shim: `from extensions.chamfer_dist import ChamferDistanceL1, ChamferDistanceL2` resolves to gm3d_b200."""
from gm3d_b200.chamfer import ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferFunction  # noqa: F401
