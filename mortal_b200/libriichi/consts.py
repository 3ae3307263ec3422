"""libriichi.consts (consts.rs:5-52)."""
MAX_VERSION = 4
ACTION_SPACE = 37 + 1 + 3 + 1 + 1 + 1 + 1 + 1  # = 46 (consts.rs:7-15)
GRP_SIZE = 7


def obs_shape(version: int):
    return {1: (938, 34), 2: (942, 34), 3: (934, 34), 4: (1012, 34)}[version]


def oracle_obs_shape(version: int):
    return {1: (211, 34), 2: (217, 34), 3: (217, 34), 4: (217, 34)}[version]
