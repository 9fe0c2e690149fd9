"""The reference keeps a second copy of the rasteriser in py_rmpe_server/py_rmpe_heatmapper_mask.py (same class, same
signatures); here both module paths resolve to the one implementation in py_rmpe_heatmapper.py."""
from .py_rmpe_heatmapper import Heatmapper, distances  # noqa: F401
