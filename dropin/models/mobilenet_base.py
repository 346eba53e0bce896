from atomnas_b200.mobilenet_base import *  # noqa: F401,F403
from atomnas_b200.mobilenet_base import _get_named_block_list, _make_divisible, active_fn_id  # noqa: F401
from atomnas_b200.fused_block import InvertedResidualChannelsFused, SqueezeAndExcitation  # noqa: F401
