"""Import shim: `from pointnet2_ops import pointnet2_utils` resolves to gm3d_b200 (see gm3d_b200.install_shims)."""
from gm3d_b200 import pointnet2_utils  # noqa: F401
