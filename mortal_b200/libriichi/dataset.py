"""libriichi.dataset — the GameplayLoader part (dataset/gameplay.rs), backed by the device log replay."""
from ..dataset import Gameplay, GameplayLoader  # noqa: F401
