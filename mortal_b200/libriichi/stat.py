"""libriichi.stat — Stat (stat.rs), host-side log statistics."""
from ..stat import Stat  # noqa: F401
