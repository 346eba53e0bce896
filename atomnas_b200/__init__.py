"""atomnas_b200 -- B200-native (sm_100a) hot path for the AtomNAS atomic-block supernet.

Public surface:
  Trainer                      fused training step (forward, loss, backward, all-reduce, RMSprop, EMA)
  mobilenet_base / mobilenet_supernet / fused_block
                               nn.Module mirror of the reference's model zoo, executed by `engine`
  prune / rmsprop / optim / compress_utils / distributed
                               mirrors of the reference's utils on the fused kernels
  kernels / _lib               the C ABI (include/atomnas_b200.h) through ctypes
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    raise AttributeError(name)
