from .shi_tomasi import ShiTomasiScore
from .akaze import AKAZE

__all__ = ["ShiTomasiScore", "AKAZE"]
