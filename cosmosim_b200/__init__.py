"""cosmosim_b200 -- B200-native per-timestep physics update behind cosmosim's Universe/Planet API.

    from cosmosim_b200 import Universe          # drop-in for the reference's README API
    u = Universe(); ...; u.simulate(steps=1000) # headless frames on the GPU

The reference's 3-D batch API (Object / State / Universe.run) lives in ``cosmosim_b200.core.state``.

``install_as_cosmosim()`` registers this package under the reference's import paths so the
reference's example scripts (``from cosmosim.core.universe import Universe``) run unmodified.
"""
from __future__ import annotations

import sys
import types

from . import _abi, _build  # noqa: F401
from .core.planet import AU, ME, MS, RE, RS, Planet
from .core.universe import Universe

__all__ = ["Universe", "Planet", "AU", "ME", "MS", "RE", "RS", "build", "install_as_cosmosim"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a) and return its path."""
    return _build.build(force=force, verbose=verbose)


def install_as_cosmosim(api: str = "old") -> None:
    """Alias this package as ``cosmosim`` (core.universe, core.planet, util.functions).

    api="old": ``cosmosim.core.universe`` is the interactive 2-D API the reference's README and
    example scenes use (reference universe_old.py); api="new": it is the 3-D batch API
    (Object / State / Universe.run of the reference's universe.py) that examples/testing/*.py use."""
    from .core import planet as _planet, state as _state, universe as _universe
    if api == "new":
        _universe_mod = _state
    elif api == "old":
        _universe_mod = _universe
    else:
        raise ValueError("api must be 'old' or 'new'")
    from .util import functions as _functions
    root = types.ModuleType("cosmosim")
    core = types.ModuleType("cosmosim.core")
    util = types.ModuleType("cosmosim.util")
    root.core, root.util = core, util
    core.universe, core.universe_old, core.planet = _universe_mod, _universe, _planet
    util.functions = _functions
    root.__path__, core.__path__, util.__path__ = [], [], []
    sys.modules.update({"cosmosim": root, "cosmosim.core": core, "cosmosim.util": util,
                        "cosmosim.core.universe": _universe_mod, "cosmosim.core.universe_old": _universe,
                        "cosmosim.core.planet": _planet, "cosmosim.util.functions": _functions})
