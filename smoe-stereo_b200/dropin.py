"""Make an unmodified reference checkout use this implementation.

The reference binds the names by value (``from core.corr import CorrBlock1D, ...`` in
core/SMoEStereo.py:7 and core/raft_stereo.py:6) and selects one by
``args.corr_implementation`` (core/SMoEStereo.py:211-221), and core/corr.py:5-8 does a bare
``import corr_sampler``.  ``install()`` therefore
  1. registers this package's ``corr_sampler`` module under that top-level name, and
  2. rebinds the four class names in every already-imported reference module that holds them.
No reference file is edited.
"""
from __future__ import annotations

import sys

from . import corr as _corr
from . import corr_sampler as _sampler

_NAMES = ("CorrBlock1D", "CorrBlockFast1D", "PytorchAlternateCorrBlock1D", "CorrSampler")
_TARGETS = ("core.corr", "core.raft_stereo", "core.SMoEStereo")
_saved: dict = {}
_ABSENT = object()


def install(modules=_TARGETS, output_ring: int = 0) -> list[str]:
    """Returns the list of 'module.name' bindings that were replaced.

    output_ring > 0 (opt-in): inference lookups are served from a ring of that many preallocated
    output buffers per block instead of a fresh tensor per call (corr.set_default_output_ring) --
    valid for the reference's update loops, which consume `corr` inside the iteration."""
    replaced = []
    _corr.set_default_output_ring(output_ring)
    if sys.modules.get("corr_sampler") is not _sampler:
        _saved[("sys.modules", "corr_sampler")] = sys.modules.get("corr_sampler")
        sys.modules["corr_sampler"] = _sampler
        replaced.append("sys.modules.corr_sampler")
    for mname in modules:
        mod = sys.modules.get(mname)
        if mod is None:
            continue
        for n in _NAMES:
            if hasattr(mod, n) and getattr(mod, n) is not getattr(_corr, n):
                _saved[(mname, n)] = getattr(mod, n)
                setattr(mod, n, getattr(_corr, n))
                replaced.append(f"{mname}.{n}")
        # core/corr.py:5-8 binds `corr_sampler` inside try/except: the global is absent when the
        # reference extension was never compiled; the reference's own CorrSampler needs it either way
        if (hasattr(mod, "corr_sampler") or mname == "core.corr") and \
                getattr(mod, "corr_sampler", None) is not _sampler:
            _saved[(mname, "corr_sampler")] = getattr(mod, "corr_sampler", _ABSENT)
            mod.corr_sampler = _sampler
            replaced.append(f"{mname}.corr_sampler")
    return replaced


def uninstall() -> None:
    _corr.set_default_output_ring(0)
    for (mname, n), old in list(_saved.items()):
        if mname == "sys.modules":
            if old is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = old
        elif mname in sys.modules:
            if old is _ABSENT:
                if hasattr(sys.modules[mname], n):
                    delattr(sys.modules[mname], n)
            else:
                setattr(sys.modules[mname], n, old)
        del _saved[(mname, n)]
