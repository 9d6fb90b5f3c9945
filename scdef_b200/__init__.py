"""scdef_b200 — B200-native training hot path of scDEF."""
from .anndata_compat import MiniAnnData
from .model import scDEF
from .sscdef import sscDEF
from .iscdef import iscDEF

__all__ = ["scDEF", "sscDEF", "iscDEF", "MiniAnnData"]
__version__ = "0.1.0"
