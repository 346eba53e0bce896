from atomnas_b200.prune import *  # noqa: F401,F403
