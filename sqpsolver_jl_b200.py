"""Import alias for the package directory ``sqpsolver.jl_b200/``.

The directory name mirrors the reference repository name (SqpSolver.jl) and
therefore contains a dot, which Python cannot import directly.  This module
turns itself into a package whose ``__path__`` is that directory, so
``import sqpsolver_jl_b200.host.sqp_trust_region`` works from the repo root.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sqpsolver.jl_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
