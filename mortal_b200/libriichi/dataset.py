"""libriichi.dataset — GameplayLoader (dataset/gameplay.rs) and Grp (dataset/grp.rs), backed by the device log replay."""
from ..dataset import Gameplay, GameplayLoader, Grp  # noqa: F401
