"""The shanten lookup tables are generated from first principles (tools/gen_shanten_tables.cc); this pins the generator."""
import gzip
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = "/root/reference/libriichi/src/algo/data"


@pytest.fixture(scope="module")
def generated():
    import build_tables

    return build_tables.generate_shanten_tables()


def nibbles(row5):
    out = []
    for b in row5:
        out += [b & 15, b >> 4]
    return out


def test_generated_rows_known_answers(generated):
    suhai = np.frombuffer(generated["shanten_suhai.bin"], dtype=np.uint8).reshape(-1, 5)
    jihai = np.frombuffer(generated["shanten_jihai.bin"], dtype=np.uint8).reshape(-1, 5)
    assert suhai.shape[0] == 1_940_777 and jihai.shape[0] == 78_032
    # empty suit: m melds cost 3m tiles, a pair 2 more (max nibble 14)
    assert nibbles(suhai[0]) == [0, 3, 6, 9, 12, 2, 5, 8, 11, 14]
    assert nibbles(jihai[0]) == [0, 3, 6, 9, 12, 2, 5, 8, 11, 14]
    idx = lambda counts: int(sum(c * 5 ** (len(counts) - 1 - i) for i, c in enumerate(counts)))
    # 123 456 789: three complete runs; a pair beside two runs borrows one tile, beside all three it needs two new ones
    assert nibbles(suhai[idx([1] * 9)]) == [0, 0, 0, 0, 3, 1, 1, 1, 2, 5]
    # 1112345678999 (the nine-gates shape): four melds + pair missing exactly one tile
    assert nibbles(suhai[idx([3, 1, 1, 1, 1, 1, 1, 1, 3])])[9] == 1
    # honours: a triplet is a meld, a pair is a pair, singles only save one tile each
    assert nibbles(jihai[idx([3, 2, 1, 0, 0, 0, 0])]) == [0, 0, 1, 3, 6, 0, 0, 2, 5, 8]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box): generator is pinned in the dev container")
def test_generated_tables_equal_reference_data(generated):
    for name in ("shanten_suhai.bin", "shanten_jihai.bin"):
        with gzip.open(os.path.join(REF, name + ".gz"), "rb") as f:
            assert f.read() == generated[name], name


def test_installed_tables_are_the_generated_ones(generated):
    for name in ("shanten_suhai.bin", "shanten_jihai.bin"):
        with open(os.path.join(ROOT, "mortal_b200", "data", name), "rb") as f:
            assert f.read() == generated[name], name
