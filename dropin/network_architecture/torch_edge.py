"""Drop-in for .../network_architecture/torch_edge.py."""
from nextou_b200.graph import (pairwise_distance, xy_pairwise_distance, dense_knn_matrix, xy_dense_knn_matrix,  # noqa: F401
                               DenseDilated, DenseDilatedKnnGraph)
