from .gconv import GIN, GINConv, global_add_pool, global_mean_pool  # noqa: F401
from .rgconv import RGIN, RGCNConv  # noqa: F401
