"""Rule-logic parity of the product's step sources (host-emulated, single lane) against the oracle.

This is the GPU-less half of the parity story: the same mjx_step.cuh that the CUDA kernel compiles is
built with -DMJX_HOST_EMUL and driven by the shared counter-based test policies; every decision row
(table, step, seat, legal-mask bits, action) and every final score / rank / step count must be equal.
The `-m gpu` tests repeat this through the real kernels and the C ABI.
"""
import numpy as np
import pytest

import emul_lib as E
import oracle_lib as O


def sort_trace(t):
    # (table, step, seat, kan) is a unique key
    order = np.lexsort((t[:, 4], t[:, 2], t[:, 1], t[:, 0]))
    return t[order]


def first_diff(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if (a[i] != b[i]).any():
            return i, a[i], b[i]
    return n, None, None


@pytest.mark.parametrize("policy_kind", [1, 0])
@pytest.mark.parametrize("quick_eval", [True, False])
@pytest.mark.parametrize("shuffle_kind", [0, 1])
def test_selfplay_trace_parity(policy_kind, quick_eval, shuffle_kind):
    n = 48
    nonces = np.repeat(np.arange(10000, 10000 + n // 4, dtype=np.uint64), 4)  # OneVsThree seed layout
    keys = np.full(n, 0x2000, dtype=np.uint64)
    cap = 1 << 17
    ro = O.run_batch(nonces, keys, shuffle_kind=shuffle_kind, policy_kind=policy_kind, quick_eval=quick_eval,
                     trace_cap=cap)
    re = E.run(nonces, keys, shuffle_kind=shuffle_kind, policy_kind=policy_kind, quick_eval=quick_eval, trace_cap=cap)
    assert (re["errs"] == 0).all(), re["errs"]
    to, te = sort_trace(ro["trace"]), sort_trace(re["trace"])
    i, a, b = first_diff(to, te)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} emul {b}"
    assert len(to) == len(te)
    assert (ro["steps"] == re["steps"]).all()
    assert (ro["scores"] == re["scores"]).all()
    assert (ro["ranks"] == re["ranks"]).all()


def test_agari_guard_parity():
    """mortal.rs:319-336 + agent_helper.rs:262-368: with the rule-based agari guard on for every seat the decisions
    (incl. refused wins at all-last) and final scores still match the oracle."""
    n = 96
    nonces = np.arange(4000, 4000 + n, dtype=np.uint64)
    keys = np.full(n, 5, dtype=np.uint64)
    ro = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=True, trace_cap=1 << 18)
    re = E.run(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=True, trace_cap=1 << 18)
    assert (re["errs"] == 0).all(), re["errs"]
    rn = O.run_batch(nonces, keys, policy_kind=1, quick_eval=True, agari_guard=False)
    assert (rn["scores"] != ro["scores"]).any(), "guard never fired: the test does not cover it"
    to, te = sort_trace(ro["trace"]), sort_trace(re["trace"])
    i, a, b = first_diff(to, te)
    assert a is None, f"first divergence at sorted row {i}: oracle {a} emul {b}"
    assert (ro["scores"] == re["scores"]).all() and (ro["ranks"] == re["ranks"]).all()


def test_emul_shanten_matches_oracle_random_hands():
    rng = np.random.default_rng(0)
    n = 20000
    tiles = np.zeros((n, 34), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint8)
    deck = np.repeat(np.arange(34, dtype=np.uint8), 4)
    for i in range(n):
        k = rng.integers(0, 4) if i % 4 == 0 else 0
        cnt = 13 + (i & 1) - 3 * k
        pick = rng.permutation(deck)[:cnt]
        np.add.at(tiles[i], pick, 1)
        lens[i] = 4 - k
    out = np.zeros(n, dtype=np.int8)
    E.lib().emul_shanten(tiles.ctypes.data, lens.ctypes.data, out.ctypes.data, n)
    assert (out == O.shanten(tiles, lens)).all()


def test_emul_wall_matches_oracle():
    for kind in (0, 1):
        for kyoku, honba in ((0, 0), (3, 2), (7, 0), (11, 5)):
            a = np.zeros(136, dtype=np.uint8)
            b = np.zeros(136, dtype=np.uint8)
            E.lib().emul_make_wall(12345678901234567, 0xDEADBEEFCAFE, kyoku, honba, kind, a.ctypes.data)
            O.lib().orc_make_wall(12345678901234567, 0xDEADBEEFCAFE, kyoku, honba, kind, b.ctypes.data)
            assert (a == b).all()


def test_obs_encode_parity_emulated():
    """v4 observation rows 0..888 + legal masks of the product encoder (host-emulated) vs the oracle."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs()
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=6, min_rows=1500)


@pytest.mark.parametrize("version", [1, 2, 3])
def test_obs_encode_parity_emulated_legacy_versions(version):
    """obs versions 1-3 (consts.rs:20-28; 938 / 942 / 934 rows) of the product encoder (host-emulated) vs the oracle."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs(version=version)
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=4, min_rows=1000, version=version)


def test_sp_block_parity_emulated():
    """obs v4 rows 889..1011 (single-player tables): product DP (host-emulated) vs the oracle's memoised recursion.
    Same f32 operation order on both sides, so the comparison is exact."""
    from obs_check import check_obs_parity

    def make_env(nonces, keys):
        return E.EmulEnv(nonces, keys, enable_quick_eval=False)

    def fetch(env, first, prev):
        env.step(None if first else prev)
        rt, rs, m = env.rows()
        obs = env.encode_obs(sp=True)
        acts = env.policy_test(1)
        return rt, rs, m, obs, acts

    check_obs_parity(make_env, fetch, n=4, max_cycles=150, min_rows=400, sp=True, sp_tol=0.0)
    assert E.lib().emul_sp_overflows() == 0
