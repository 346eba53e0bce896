from atomnas_b200.rmsprop import *  # noqa: F401,F403
