from atomnas_b200.mobilenet_supernet import MobileNetV2, get_block_wrapper  # noqa: F401
Model = MobileNetV2
