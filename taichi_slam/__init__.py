"""Namesake of the reference's top-level package so `from taichi_slam.mapping import *`
(scripts/taichislam_node.py:6, TaichiSLAM_demo.py) resolves to the B200 backend."""
