"""`vpower` call surface (vpower.interp / vpower.spctrm) backed by libvpower_b200.so.

Put the directory that contains this package on sys.path (or use `vpower_b200_path()` from the repo's
`__graft_entry__`) and `import vpower.interp`, exactly as with the reference package.
"""
from . import interp, spctrm  # noqa: F401
from .interp import *  # noqa: F401,F403
from .spctrm import *  # noqa: F401,F403
