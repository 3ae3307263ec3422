"""libriichi.dataset.GameplayLoader on the CUDA environment.

Mirror of dataset/gameplay.rs:24-45, 80-218: `GameplayLoader(version, *, oracle, player_names, excludes, trust_seed,
always_include_kan_select, augmented)`, `.load_gz_log_files(filenames) -> list[list[Gameplay]]`, `.load_log(text)`;
`Gameplay.take_obs / take_actions / take_masks / take_at_kyoku / take_dones / take_apply_gamma / take_at_turns /
take_shantens / take_player_id`. The logs are replayed on device (csrc/mjx_replay.cuh) and the observations come from the
same encoder kernels self-play uses. Differences, stated rather than hidden: `take_obs()` / `take_masks()` return ONE tensor
per Gameplay ([n_moves, C, 34] float32 / [n_moves, 46] bool, CUDA by default, `host=True` for numpy) instead of a list of
per-move arrays (`take_invisible_obs()` likewise: [n_moves, 217 | 211, 34]); logs must carry full information (no "?" tiles).
With `oracle=True` the hidden tiles come from the game seed when `trust_seed=True` (walls regenerated on device and checked
against the logged deal) and otherwise from the log plus a random fill of the unseen tiles, as dataset/invisible.rs does.
`Grp` (dataset/grp.rs:19-164) is plain host-side log arithmetic and is provided in Python.
"""
from __future__ import annotations

import gzip

import numpy as np

from . import dataset_codec
from .env import ReplayEnv


class Grp:
    """dataset/grp.rs:19-164: per-kyoku features [grand_kyoku, honba, kyotaku, scores / 10000] (float64 [n_kyoku, 7]),
    the final ranking and final scores of a game. Pure log arithmetic on the host."""

    def __init__(self, feature, rank_by_player, final_scores):
        self.feature, self.rank_by_player, self.final_scores = feature, rank_by_player, final_scores

    @staticmethod
    def load_events(events):
        info, rank, final_deltas, final_scores = [], None, [0, 0, 0, 0], [0, 0, 0, 0]
        for ev in reversed(events):  # grp.rs:96-152 walks the log backwards
            ty = ev["type"]
            if ty in ("hora", "ryukyoku"):
                if rank is None:
                    if ev.get("deltas") is None:
                        raise ValueError("invalid log: field `deltas` is required for Hora and Ryukyoku of AL")
                    final_deltas = [a + b for a, b in zip(final_deltas, ev["deltas"])]
            elif ty == "reach_accepted":
                if rank is None:
                    final_deltas[ev["actor"]] -= 1000
            elif ty == "start_kyoku":
                if rank is None:
                    final_scores = [a + b for a, b in zip(ev["scores"], final_deltas)]
                    order = sorted(range(4), key=lambda i: -final_scores[i])  # rankings.rs:8-22: stable by seat
                    total = sum(final_scores)
                    if total < 100_000:  # leftover riichi sticks go to the top (grp.rs:127-131)
                        final_scores[order[0]] += 100_000 - total
                    rank = [0, 0, 0, 0]
                    for r, pl in enumerate(order):
                        rank[pl] = r
                grand = {"E": ev["kyoku"] - 1, "S": 3 + ev["kyoku"]}.get(ev["bakaze"], 7 + ev["kyoku"])
                info.insert(0, [float(grand), float(ev["honba"]), float(ev["kyotaku"])] + [sc / 10000.0 for sc in ev["scores"]])
        if rank is None:
            raise ValueError("invalid log: no Hora or Ryukyoku after a StartKyoku")
        return Grp(np.array(info, dtype=np.float64).reshape(-1, 7), rank, final_scores)

    @staticmethod
    def load_log(raw_log: str):
        return Grp.load_events(dataset_codec.parse_log(raw_log))

    @staticmethod
    def load_gz_log_files(gzip_filenames):
        out = []
        for fn in gzip_filenames:
            with gzip.open(fn, "rt") as f:
                out.append(Grp.load_log(f.read()))
        return out

    def take_feature(self):
        return self.feature

    def take_rank_by_player(self):
        return list(self.rank_by_player)

    def take_final_scores(self):
        return list(self.final_scores)

    def __len__(self):
        return int(self.feature.shape[0])


class Gameplay:
    def __init__(self, player_id: int, player_name: str, obs, actions, masks, at_kyoku, apply_gamma, at_turns, shantens, grp=None,
                 invisible_obs=None):
        self.player_id, self.player_name = player_id, player_name
        self.grp = grp
        self._invisible = invisible_obs
        self._obs, self._masks = obs, masks
        self._actions, self._at_kyoku, self._apply_gamma = actions, at_kyoku, apply_gamma
        self._at_turns, self._shantens = at_turns, shantens
        # gameplay.rs:285-286: done wherever the next move belongs to a later kyoku, and at the last move
        self._dones = np.append(at_kyoku[1:] > at_kyoku[:-1], True) if len(at_kyoku) else np.zeros(0, dtype=bool)

    def take_obs(self, host: bool = False):
        return self._obs.cpu().numpy() if host else self._obs

    def take_masks(self, host: bool = False):
        return self._masks.cpu().numpy() if host else self._masks

    def take_invisible_obs(self, host: bool = False):
        """gameplay.rs:199-201: the invisible (oracle) observation of every move; only with GameplayLoader(oracle=True)"""
        if self._invisible is None:
            raise ValueError("the loader was created with oracle=False")
        return self._invisible.cpu().numpy() if host else self._invisible

    def take_grp(self):
        return self.grp

    def take_actions(self):
        return self._actions.tolist()

    def take_at_kyoku(self):
        return self._at_kyoku.tolist()

    def take_dones(self):
        return self._dones.tolist()

    def take_apply_gamma(self):
        return self._apply_gamma.tolist()

    def take_at_turns(self):
        return self._at_turns.tolist()

    def take_shantens(self):
        return self._shantens.tolist()

    def take_player_id(self):
        return self.player_id


class GameplayLoader:
    def __init__(self, version: int, *, oracle: bool = True, player_names=None, excludes=None, trust_seed: bool = False,
                 always_include_kan_select: bool = True, augmented: bool = False, device: int = 0, shuffle_kind: int = 0, rng=None):
        # gameplay.rs:80-113: same keywords and defaults (oracle = true, always_include_kan_select = true); `shuffle_kind` selects the
        # wall shuffle of the seeds (see mjx_env_create), `rng` the numpy Generator behind the random fill of unseen tiles
        if oracle and trust_seed and augmented:
            raise NotImplementedError("oracle + trust_seed + augmented: the regenerated walls would not match the augmented events "
                                      "(the reference mismatches them too, dataset/invisible.rs:51-66)")
        self.version, self.oracle, self.trust_seed = version, oracle, trust_seed
        self.shuffle_kind = shuffle_kind
        self.rng = rng if rng is not None else np.random.default_rng()
        self.player_names, self.excludes = list(player_names or []), list(excludes or [])
        self.always_include_kan_select, self.augmented = always_include_kan_select, augmented
        self.device = device

    def _players(self, names):  # gameplay.rs:166-176
        if self.player_names:
            return [i for i, nm in enumerate(names) if nm in set(self.player_names)]
        if self.excludes:
            return [i for i, nm in enumerate(names) if nm not in set(self.excludes)]
        return [0, 1, 2, 3]

    def load_log(self, raw_log: str):
        return self.load_logs([raw_log])[0]

    def load_gz_log_files(self, gzip_filenames):
        texts = []
        for fn in gzip_filenames:
            with gzip.open(fn, "rt") as f:
                texts.append(f.read())
        return self.load_logs(texts)

    def load_logs(self, texts):
        """list of log texts -> list (per log) of list (per selected player) of Gameplay; all logs replayed as one batch"""
        import torch

        games = [dataset_codec.parse_log(t) for t in texts]
        if self.augmented:  # gameplay.rs:126-128: manzu <-> pinzu on every event before anything else
            games = [dataset_codec.augment_events(ev) for ev in games]
        for ev in games:
            if not ev or ev[0].get("type") != "start_game" or len(ev) < 4:
                raise ValueError("empty or invalid game log")
        players = [self._players(ev[0].get("names", ["", "", "", ""])) for ev in games]
        walls = None
        if self.oracle and not self.trust_seed:  # dataset/invisible.rs:24-148: from the log, unseen tiles filled at random
            walls = [dataset_codec.reconstruct_walls(ev, self.rng) for ev in games]
        jobs = dataset_codec.build_jobs(games, players, walls)
        n_jobs = len(jobs["players"])
        out = [[] for _ in games]
        if n_jobs == 0:
            return out
        env = ReplayEnv(jobs, obs_version=self.version, always_include_kan_select=self.always_include_kan_select, device=self.device)
        chunks = []  # per step: (job ids, obs, masks, labels, meta[, invisible obs])
        try:
            if self.oracle and self.trust_seed:  # invisible.rs:35-66: the walls are those of the game's seed
                seeds = []
                for ev in games:
                    sd = ev[0].get("seed")
                    if sd is None:
                        raise ValueError("trust_seed=True needs start_game.seed in every log")
                    seeds.append((int(sd[0]), int(sd[1])))
                env.trust_seeds([seeds[g][0] for g in jobs["job_game"]], [seeds[g][1] for g in jobs["job_game"]], self.shuffle_kind)
            while True:
                env.replay_step()
                nr = env.num_rows()
                if nr:
                    obs = env.encode_obs()[:nr]
                    chunk = (env.row_table[:nr].long().clone(), obs.clone(), env.masks[:nr].clone(),
                             env.row_label[:nr].clone(), env.row_meta[:nr].clone())
                    if self.oracle:
                        chunk += (env.encode_invisible(self.version)[:nr].clone(),)
                    chunks.append(chunk)
                if env.num_live() == 0:
                    break
            res = env.results()
            sp_overflows = env.sp_overflows() if self.version == 4 else 0
        finally:
            env.close()
        if sp_overflows:
            raise RuntimeError(f"single-player state arena overflowed in {sp_overflows} replay step(s): observation rows 889-1011 "
                               "would be zero; load fewer logs per call")
        if (res["err"] != 0).any():
            bad = int(np.nonzero(res["err"])[0][0])
            raise RuntimeError(f"replay job {bad} (log {int(jobs['job_game'][bad])}, player {int(jobs['players'][bad])}) failed "
                               f"with mjx error code {int(res['err'][bad])}: the log is inconsistent or not full-information")
        if chunks:
            job = torch.cat([c[0] for c in chunks]); obs = torch.cat([c[1] for c in chunks]); masks = torch.cat([c[2] for c in chunks])
            label = torch.cat([c[3] for c in chunks]); meta = torch.cat([c[4] for c in chunks])
            order = torch.sort(job, stable=True).indices  # moves of a job stay in emission order
            inv = torch.cat([c[5] for c in chunks])[order] if self.oracle else None
            job, obs, masks, label, meta = job[order], obs[order], masks[order], label[order].cpu().numpy(), meta[order].cpu().numpy()
            counts = torch.bincount(job, minlength=n_jobs).cpu().numpy()
        else:
            counts = np.zeros(n_jobs, dtype=np.int64)
        grps = [Grp.load_events(ev) for ev in games]
        start = 0
        for j in range(n_jobs):
            n = int(counts[j]); sl = slice(start, start + n); start += n
            g, pid = int(jobs["job_game"][j]), int(jobs["players"][j])
            name = games[g][0].get("names", ["", "", "", ""])[pid]
            if n:
                gp = Gameplay(pid, name, obs[sl], label[sl], masks[sl], meta[sl, 0].copy(), meta[sl, 3].astype(bool),
                              meta[sl, 1].copy(), meta[sl, 2].astype(np.int8), grp=grps[g], invisible_obs=None if inv is None else inv[sl])
            else:
                z = np.zeros(0, dtype=np.uint8)
                gp = Gameplay(pid, name, torch.zeros((0, env.obs_rows, 34), device=obs.device if chunks else "cpu"), np.zeros(0, dtype=np.int64),
                              torch.zeros((0, 46), dtype=torch.bool), z, z.astype(bool), z, z.astype(np.int8), grp=grps[g])
            out[g].append(gp)
        return out
