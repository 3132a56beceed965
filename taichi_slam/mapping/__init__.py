"""taichi_slam.mapping -> taichislam_b200.mapping (same names as the reference's mapping/__init__.py:1-6)."""
from taichislam_b200.mapping import *  # noqa: F401,F403
from taichislam_b200.mapping import __all__  # noqa: F401
