"""`import munk` shim: the reference package name, served by munk_b200 (drop-in)."""
from munk_b200 import *  # noqa: F401,F403
from munk_b200 import util, io, __version__  # noqa: F401
from munk_b200 import munk as _munk
import sys as _sys

_sys.modules.setdefault(__name__ + ".util", util)
_sys.modules.setdefault(__name__ + ".io", io)
_sys.modules.setdefault(__name__ + ".munk", _munk)
