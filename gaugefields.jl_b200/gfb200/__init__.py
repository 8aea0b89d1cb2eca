"""gfb200 -- host-side mirror of the Gaugefields.jl API over libgfb200.so (B200, sm_100a).

Import path: add `<repo>/gaugefields.jl_b200` to sys.path (the directory name carries a dot, as the
reference's package name does), then `import gfb200`.
"""
from . import _lib
from .api import *  # noqa: F401,F403
from .api import md_step_size, pinned_empty  # noqa: F401

LIB_PATH = _lib.LIB_PATH
