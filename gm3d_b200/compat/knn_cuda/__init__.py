"""Import shim: `from knn_cuda import KNN` resolves to gm3d_b200.knn.KNN."""
from gm3d_b200.knn import KNN  # noqa: F401
