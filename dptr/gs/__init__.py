"""``dptr.gs`` -> splatter_a_video_b200.gs (same names/signatures as the reference module of this name)."""
from splatter_a_video_b200.gs import *  # noqa: F401,F403
from splatter_a_video_b200.gs import __all__  # noqa: F401
