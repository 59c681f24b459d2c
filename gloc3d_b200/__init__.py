"""gloc3d_b200 -- B200-native (sm_100a) implementation of GLoc3D's global-localization
query path: exhaustive exact top-k descriptor retrieval and correlative / branch-and-bound
scan-match verification, behind the reference's own interfaces.

The product is libgloc3d.so (hand-written CUDA, C ABI in include/gloc3d.h); this package
is the host-side mirror of the reference interface used by tests and bench.py.  There is
no CPU fallback anywhere in this package.
"""
from . import _lib  # noqa: F401
from ._lib import GlocError, KNN_AUTO, KNN_EXACT_SCAN, KNN_SHORTLIST  # noqa: F401
from .bev import BevProjector  # noqa: F401
from .descriptor import DescriptorExtractor, Encoder, NetVladHead  # noqa: F401
from .grid_store import StoredGrid, read_grid_file, write_grid_file  # noqa: F401
from .localize import LOC_FIRST_MATCH, LOC_VERIFY_ALL, Localizer  # noqa: F401
from .retrieval import InvKeyTree, KnnIndex, merge_topk_device  # noqa: F401
from .scan_matching import (CsmStore, FastCorrelativeScanMatcher2D,  # noqa: F401
                            FastCorrelativeScanMatcherOptions2D, MapLimits, ProbabilityGrid,
                            Rigid2d, grid_to_virtual_point_cloud, search_parameters)

__all__ = ["GlocError", "Localizer", "LOC_VERIFY_ALL", "LOC_FIRST_MATCH", "BevProjector", "InvKeyTree", "KnnIndex", "merge_topk_device", "CsmStore",
           "FastCorrelativeScanMatcher2D", "FastCorrelativeScanMatcherOptions2D", "MapLimits",
           "ProbabilityGrid", "Rigid2d", "grid_to_virtual_point_cloud", "search_parameters",
           "KNN_AUTO", "KNN_EXACT_SCAN", "KNN_SHORTLIST", "StoredGrid", "read_grid_file", "write_grid_file", "NetVladHead", "Encoder", "DescriptorExtractor"]
