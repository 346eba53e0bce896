from atomnas_b200.compress_utils import *  # noqa: F401,F403
from atomnas_b200.compress_utils import _copy, _find_only_one, _mask_along_dim, _scatter_by_bool  # noqa: F401
