"""Vectorised synthetic inputs for BASELINE configs[2] (shanten + agari at 1M hands): shared by the GPU parity tests and
bench.py. Pure numpy, seeded; SURVEY.md §8(d) config 3 describes the distributions."""
import numpy as np

AGARI_IN_DTYPE = np.dtype([
    ("tehai", "u1", 34), ("chis", "u1", 4), ("pons", "u1", 4), ("minkans", "u1", 4), ("ankans", "u1", 4),
    ("n_chis", "u1"), ("n_pons", "u1"), ("n_minkans", "u1"), ("n_ankans", "u1"),
    ("bakaze", "u1"), ("jikaze", "u1"), ("winning_tile", "u1"), ("is_ron", "u1"),
    ("additional_hans", "u1"), ("doras", "u1"), ("is_oya", "u1"), ("pad", "u1"),
])
AGARI_OUT_DTYPE = np.dtype([("kind", "u1"), ("fu", "u1"), ("han", "u1"), ("yakuman", "u1"),
                            ("ron", "<i4"), ("tsumo_ko", "<i4"), ("tsumo_oya", "<i4")])


def random_hands(n, seed=0, chunk=1 << 17):
    """Seeded shuffles of the 136-tile multiset: hand i = the first 13 + (i & 1) tiles; every fourth hand is a melded
    variant (len_div3 = 4 - k, 3k tiles dropped, k in 0..3) -> (tiles uint8 [n, 34], len_div3 uint8 [n])."""
    rng = np.random.default_rng(seed)
    tiles = np.zeros((n, 34), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint8)
    deck = np.repeat(np.arange(34, dtype=np.uint8), 4)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        idx = np.arange(lo, hi)
        k = np.where(idx % 4 == 0, rng.integers(0, 4, size=m), 0)
        cnt = 13 + (idx & 1) - 3 * k
        perm = rng.permuted(np.broadcast_to(deck, (m, 136)), axis=1)[:, :14]
        keep = np.arange(14)[None, :] < cnt[:, None]
        flat = (np.arange(m)[:, None] * 34 + perm)[keep]
        tiles[lo:hi] = np.bincount(flat, minlength=m * 34).reshape(m, 34).astype(np.uint8)
        lens[lo:hi] = 4 - k
    return tiles, lens


def winning_hands(n, seed=1):
    """Constructed winning hands: 4 mentsu + pair (90 %), chiitoi (5 %), kokushi (5 %), at most four copies per tile; each
    mentsu is melded with probability 0.3 (chi / pon / minkan / ankan); random winning tile from the closed part, winds,
    ron / tsumo, situational han and dora counts -> structured array of mjx_agari_in records (include/mjx.h)."""
    rng = np.random.default_rng(seed)
    out = np.zeros(n, dtype=AGARI_IN_DTYPE)
    filled = 0
    yao = np.array([0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33])
    while filled < n:
        m = int((n - filled) * 1.6) + 1024
        q = np.zeros(m, dtype=AGARI_IN_DTYPE)
        kind = rng.integers(0, 20, size=m)
        closed = np.zeros((m, 34), dtype=np.int16)
        total = np.zeros((m, 34), dtype=np.int16)
        rows = np.arange(m)
        # --- standard hands
        for j in range(4):
            melded = rng.random(m) < 0.3
            is_run = rng.random(m) < 0.55
            run_start = rng.integers(0, 3, size=m) * 9 + rng.integers(0, 7, size=m)
            trip_tile = rng.integers(0, 34, size=m)
            r = rng.random(m)
            for d in range(3):
                np.add.at(total, (rows[is_run], run_start[is_run] + d), 1)
                sel = is_run & ~melded
                np.add.at(closed, (rows[sel], run_start[sel] + d), 1)
            is_minkan = ~is_run & melded & (r < 0.15)
            is_ankan = ~is_run & melded & (r >= 0.15) & (r < 0.3)
            is_pon = ~is_run & melded & (r >= 0.3)
            is_closed_trip = ~is_run & ~melded
            np.add.at(total, (rows[~is_run], trip_tile[~is_run]), 3)
            kan = is_minkan | is_ankan
            np.add.at(total, (rows[kan], trip_tile[kan]), 1)
            np.add.at(closed, (rows[is_closed_trip], trip_tile[is_closed_trip]), 3)
            for name, sel, val in (("chis", is_run & melded, run_start), ("pons", is_pon, trip_tile),
                                   ("minkans", is_minkan, trip_tile), ("ankans", is_ankan, trip_tile)):
                cnt = q["n_" + name]
                ii = np.nonzero(sel)[0]
                q[name][ii, cnt[ii]] = val[ii]
                cnt[ii] += 1
        pair = rng.integers(0, 34, size=m)
        np.add.at(total, (rows, pair), 2)
        np.add.at(closed, (rows, pair), 2)
        ok = (total <= 4).all(axis=1)
        # --- chiitoi
        chi = kind == 0
        if chi.any():
            ii = np.nonzero(chi)[0]
            pick = np.argsort(rng.random((len(ii), 34)), axis=1)[:, :7]
            c = np.zeros((len(ii), 34), dtype=np.int16)
            np.put_along_axis(c, pick, 2, axis=1)
            closed[ii] = c
            for name in ("chis", "pons", "minkans", "ankans"):
                q[name][ii] = 0
                q["n_" + name][ii] = 0
            ok[ii] = True
        # --- kokushi
        kok = kind == 1
        if kok.any():
            ii = np.nonzero(kok)[0]
            c = np.zeros((len(ii), 34), dtype=np.int16)
            c[:, yao] = 1
            c[np.arange(len(ii)), yao[rng.integers(0, 13, size=len(ii))]] += 1
            closed[ii] = c
            for name in ("chis", "pons", "minkans", "ankans"):
                q[name][ii] = 0
                q["n_" + name][ii] = 0
            ok[ii] = True
        q["tehai"] = closed.astype(np.uint8)
        # winning tile: a uniformly random tile kind present in the closed part
        present = closed > 0
        score = rng.random((m, 34)) * present
        q["winning_tile"] = score.argmax(axis=1)
        q["bakaze"] = 27 + rng.integers(0, 3, size=m)
        q["jikaze"] = 27 + rng.integers(0, 4, size=m)
        q["is_ron"] = rng.integers(0, 2, size=m)
        q["additional_hans"] = rng.integers(0, 4, size=m)
        q["doras"] = rng.integers(0, 5, size=m)
        q["is_oya"] = q["jikaze"] == 27
        good = q[ok]
        take = min(len(good), n - filled)
        out[filled:filled + take] = good[:take]
        filled += take
    return out
