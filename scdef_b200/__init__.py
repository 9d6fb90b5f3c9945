"""scdef_b200 — B200-native training hot path of scDEF."""
from .anndata_compat import MiniAnnData
from .model import scDEF

__all__ = ["scDEF", "MiniAnnData"]
__version__ = "0.1.0"
