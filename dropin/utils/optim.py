from atomnas_b200.optim import *  # noqa: F401,F403
