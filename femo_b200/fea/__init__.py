from .fea_b200 import FEA  # noqa: F401
