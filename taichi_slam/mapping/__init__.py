from taichislam_b200.mapping import *  # noqa: F401,F403
from taichislam_b200.mapping import __all__  # noqa: F401
