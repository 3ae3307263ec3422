"""Python surface of libriichi for the self-play hot path, backed by the CUDA environment.

Mirrors the submodule layout of libriichi's PyO3 module (lib.rs:152-157, py_helper.rs:5-18): `consts`,
`arena`, `dataset` (GameplayLoader, Grp), `stat` (Stat), `state` (PlayerState, ActionCandidate), `mjai` (Bot). `install()` registers
this package as `libriichi` (and `libriichi.<sub>`) in sys.modules the way the reference's `add_submodule` does, so
`from libriichi.arena import OneVsThree` / `from libriichi.mjai import Bot` in mortal/*.py resolve here.
"""
import sys

from . import arena, consts, dataset, mjai, stat, state  # noqa: F401

__profile__ = "release"
__version__ = "0.1.0-mortal_b200"


def install() -> None:
    mod = sys.modules[__name__]
    sys.modules.setdefault("libriichi", mod)
    for sub in ("arena", "consts", "dataset", "stat", "state", "mjai"):
        sys.modules.setdefault(f"libriichi.{sub}", getattr(mod, sub))
