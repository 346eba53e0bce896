from atomnas_b200.distributed import *  # noqa: F401,F403
from atomnas_b200.distributed import _allreduce, _get_env  # noqa: F401
