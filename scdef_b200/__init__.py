"""scdef_b200 — B200-native training hot path of scDEF."""
from .anndata_compat import MiniAnnData
from .model import scDEF
from .sscdef import sscDEF
from .iscdef import iscDEF
from . import tools as tl  # `scdef.tl`: device versions of `assign_confident` and `set_confident_signatures`

__all__ = ["scDEF", "sscDEF", "iscDEF", "MiniAnnData", "tl"]
__version__ = "0.1.0"
