from atomnas_b200.mobilenet_supernet import MobileNetSearched  # noqa: F401
Model = MobileNetSearched
