"""libriichi.arena — OneVsThree / TwoVsTwo on the CUDA environment.

Call-compatible with arena/one_vs_three.rs:17-113 and arena/two_vs_two.rs:17-110 for the `py_vs_py` entry
(the one mortal/player.py:64-69,142-147 and mortal/one_vs_three.py:88-93 use). Seat / seed layout:
one_vs_three.rs:140-191 (game g = 4*s + r uses seed (seed_start[0] + s, seed_start[1]); the challenger
sits at absolute seat r). The step loop is BatchGame::run (game.rs:286-304) executed by mjx kernels;
engines are called once per cycle per agent like MortalBatchAgent::evaluate (mortal.rs:114-159).
"""
from __future__ import annotations

import time

import numpy as np

from ..engine import HostProtocolEngine
from ..env import BatchEnv


def _adapt(engine):
    if hasattr(engine, "react_device"):
        return engine
    return HostProtocolEngine(engine)


class _MetaRecorder:
    """Collects what agent/mortal.rs:161-186 gen_meta puts into a logged reaction: per step the rows' (table, seat, kan-select,
    action, legal mask, Q-values, shanten / furiten read back from the v4 observation) and the per-table log length right after
    every environment step; `finish()` groups them per game for mortal_b200.mjai_log.attach_meta."""

    def __init__(self, n_games: int, version: int):
        self.n, self.version = n_games, version
        self.bounds, self.rows, self.q = [], [], {}
        self.error = None  # the metadata is optional: a failure while recording must never take the game loop down

    def _guard(fn):
        def wrapped(self, *a, **k):
            if self.error is not None:
                return None
            try:
                return fn(self, *a, **k)
            except Exception as exc:
                self.error = exc
                return None
        return wrapped

    @_guard
    def add_bounds(self, log_len_dev):
        self.bounds.append(log_len_dev.cpu().numpy().copy())

    @_guard
    def add_agent(self, cycle, idx, q, eval_ns):
        self.q.setdefault(cycle, []).append((idx.cpu().numpy(), q.float().cpu().numpy(), int(eval_ns)))

    @_guard
    def add_rows(self, cycle, tbl, row_seat, actions, masks, obs):
        sh = fu = None
        if self.version == 4:  # v4 rows 861 (furiten) and 862-868 (shanten one-hot), obs_repr.rs
            sh = obs[:, 862:869, 0].argmax(1).cpu().numpy()
            fu = (obs[:, 861, 0] > 0).cpu().numpy()
        self.rows.append((cycle, tbl.cpu().numpy(), row_seat.cpu().numpy(), actions.cpu().numpy(), masks.cpu().numpy().astype(bool), sh, fu))

    def finish(self):
        from ..mjai_log import make_meta

        if self.error is not None:
            raise self.error
        decisions = [dict() for _ in range(self.n)]
        for cycle, tbl, rs, act, masks, sh, fu in self.rows:
            nr = len(tbl)
            q_rows = np.zeros((nr, 46), dtype=np.float32)
            batch, ns = np.zeros(nr, dtype=np.int64), np.zeros(nr, dtype=np.int64)
            for idx, q, eval_ns in self.q.get(cycle, []):
                q_rows[idx] = q
                batch[idx] = len(idx)
                ns[idx] = eval_ns
            pos = {(int(tbl[r]), int(rs[r] & 3), bool(rs[r] & 4)): r for r in range(nr)}
            for (t, seat, kan), r in pos.items():
                if kan:
                    continue
                common = dict(batch_size=int(batch[r]), eval_time_ns=int(ns[r]), shanten=None if sh is None else int(sh[r]),
                              at_furiten=None if fu is None else bool(fu[r]))
                kan_meta = None
                kr = pos.get((t, seat, True))
                if kr is not None and int(act[r]) == 42:
                    km = make_meta(int(act[kr]), masks[kr], q_rows[kr], **common)
                    kan_meta = {k: v for k, v in km.items() if not k.startswith("_") and v is not None}
                decisions[t].setdefault(cycle, {})[seat] = make_meta(int(act[r]), masks[r], q_rows[r], kan_select=kan_meta, **common)
        return np.array(self.bounds), decisions


class _Arena:
    SEATS_PER_SEED = 4

    def __init__(self, *, disable_progress_bar: bool = False, log_dir=None, shuffle_kind: int = 0, device: int = 0):
        self.disable_progress_bar = disable_progress_bar
        self.log_dir = log_dir
        self.shuffle_kind = shuffle_kind
        self.device = device
        self.last_stats = None
        self.record_decisions = False  # test hook: keep (table, step, seat, kan_select, action) of every row
        self.last_decisions = None
        self.log_dir = log_dir  # arena/one_vs_three.rs:26-34: gz mjai logs are written here when set
        self.log_meta = True    # attach the per-decision meta (q-values, mask bits, ...) to the agent events (mortal.rs:161-186)
        self.last_meta_error = None
        self.max_cycles = 0     # test hook: stop after this many BatchGame::run cycles (0 = play every table to the end)
        self.fast_forward_steps = 0  # bench hook: play this many batch steps with the counter-free test policy (kind 2) first
        self.cycle_hook = None       # bench hook: callable(cycle_index, env) at the top of every cycle
        self.env_factory = BatchEnv  # test hook: tests/emul_batch_env.py injects the host-emulated environment; the product has no CPU path
        self.last_decision_masks = None  # with record_decisions: the legal mask (46 bits) each recorded row was decided under

    def _challenger_seats(self, game_in_seed: int):
        raise NotImplementedError

    def _run(self, challenger, champion, seed_start, seed_count):
        import torch

        agents = [_adapt(challenger), _adapt(champion)]
        for a in agents:
            if getattr(a, "is_oracle", False):
                raise NotImplementedError("oracle (invisible) observations are out of this round's scope")
            if getattr(a, "version", 4) not in (1, 2, 3, 4):
                raise ValueError(f"unsupported obs version {a.version} (consts.rs:18 MAX_VERSION = 4)")
        versions = [int(getattr(a, "version", 4)) for a in agents]
        if versions[0] != versions[1]:
            raise NotImplementedError("challenger and champion must use the same obs version (one encoder pass per step)")
        qe = [bool(getattr(a, "enable_quick_eval", True)) for a in agents]
        if qe[0] != qe[1]:
            raise NotImplementedError("challenger and champion must agree on enable_quick_eval")
        per = self.GAMES_PER_SEED
        n = int(seed_count) * per
        nonces = np.repeat(np.arange(seed_start[0], seed_start[0] + int(seed_count), dtype=np.uint64), per)
        keys = np.full(n, seed_start[1], dtype=np.uint64)
        env = self.env_factory(nonces, keys, obs_version=versions[0], shuffle_kind=self.shuffle_kind, enable_quick_eval=qe[0],
                               device=self.device)
        dev = env.device
        # seat -> agent index table per game-in-seed
        is_challenger = torch.zeros((per, 4), dtype=torch.bool, device=dev)
        for g in range(per):
            for s in self._challenger_seats(g):
                is_challenger[g, s] = True
        meta_rec = None
        if self.log_dir is not None:
            env.enable_log()
            meta_rec = _MetaRecorder(n, versions[0]) if self.log_meta else None
        actions = torch.zeros(env.row_cap, dtype=torch.int64, device=dev)
        guards = [bool(getattr(a, "enable_rule_based_agari_guard", False)) for a in agents]
        q_all = None
        if any(guards):  # mortal.rs:319-336 needs the Q-values of the previous decision
            flags = np.zeros((n, 4), dtype=np.uint8)
            ic = is_challenger.cpu().numpy()
            for g in range(n):
                for seat in range(4):
                    flags[g, seat] = guards[0] if ic[g % per, seat] else guards[1]
            env.set_agari_guard(flags)
            q_all = torch.zeros((env.row_cap, 46), dtype=torch.float32, device=dev)
        # engines that only speak the reference protocol (react_batch over host arrays) get the observations through
        # mjx_env_encode_obs_host: pinned host buffers, D2H overlapped with the single-player kernels
        host_mode = all(isinstance(a, HostProtocolEngine) for a in agents)
        if host_mode:
            pin = (lambda t: t.pin_memory()) if dev.type == "cuda" else (lambda t: t)
            h_obs = pin(torch.empty((env.row_cap, env.obs_rows, 34), dtype=torch.float32))
            h_masks = pin(torch.empty((env.row_cap, 46), dtype=torch.bool))
            h_actions = pin(torch.zeros(env.row_cap, dtype=torch.int64))
            h_q = pin(torch.zeros((env.row_cap, 46), dtype=torch.float32)) if q_all is not None else None
            obs_np, masks_np = h_obs.numpy(), h_masks.numpy()
            ic_host = is_challenger.cpu().numpy()
        first = True
        cycles = 0
        recorded, recorded_masks = [], []
        mask_weights = (1 << torch.arange(46, dtype=torch.int64))

        def check_health():
            # game.rs:288,292: an error from any table aborts the whole batch at that cycle (`?`); so does a single-player arena
            # overflow, which would otherwise hand zeroed rows 889-1011 to the engines
            nr_, live_, n_err, sp_ovf = env.poll()
            if n_err:
                res_ = env.results()
                bad = int(np.nonzero(res_["err"])[0][0])
                env.close()
                raise RuntimeError(f"table {bad} (seed {int(nonces[bad])},{int(keys[bad])}) failed at cycle {cycles} with mjx error code "
                                   f"{int(res_['err'][bad])} (invalid action or inconsistent state; board.rs:527-532)")
            if sp_ovf:
                env.close()
                raise RuntimeError(f"single-player state arena overflowed at cycle {cycles}: observation rows 889-1011 would be zero; "
                                   "run fewer tables per environment")
            return nr_, live_

        skip_step = False
        if self.fast_forward_steps:
            env.step(None)
            for _ in range(int(self.fast_forward_steps)):
                env.policy_test(2, actions)
                env.step(actions)
            first, skip_step = False, True
        while True:
            if self.cycle_hook is not None:
                self.cycle_hook(cycles, env)
            if self.max_cycles and cycles >= self.max_cycles:
                break
            if not skip_step:
                env.step(None if first else actions, None if first else q_all)
            first = skip_step = False
            if meta_rec is not None:
                meta_rec.add_bounds(env.log_len)
            nr, n_live = check_health()
            if host_mode:
                if nr == 0 and n_live == 0:
                    break
                if nr > 0:
                    assert env.encode_obs_host(h_obs, h_masks) == nr
                if nr > 0:
                    tbl_h = env.row_table[:nr].cpu().numpy()
                    rs_h = env.row_seat[:nr].cpu().numpy()
                    chal_h = ic_host[tbl_h % per, rs_h & 3]
                    for idx, agent in ((np.nonzero(chal_h)[0], agents[0]), (np.nonzero(~chal_h)[0], agents[1])):
                        if idx.size == 0:
                            continue
                        t_eval = time.perf_counter_ns()
                        a, q = agent.react_host(obs_np, masks_np, idx)
                        if meta_rec is not None:
                            meta_rec.add_agent(cycles, torch.from_numpy(idx), torch.from_numpy(q).reshape(-1, 46), time.perf_counter_ns() - t_eval)
                        h_actions[torch.from_numpy(idx)] = torch.from_numpy(a)
                        if h_q is not None:
                            h_q[torch.from_numpy(idx)] = torch.from_numpy(q).reshape(-1, 46)
                    if meta_rec is not None:
                        meta_rec.add_rows(cycles, torch.from_numpy(tbl_h).long(), torch.from_numpy(rs_h), h_actions[:nr], h_masks[:nr], h_obs[:nr])
                    actions[:nr].copy_(h_actions[:nr], non_blocking=True)
                    if h_q is not None:
                        q_all[:nr].copy_(h_q[:nr], non_blocking=True)
                    if self.record_decisions:
                        recorded.append(torch.stack([torch.from_numpy(tbl_h).long(), env.row_step[:nr].cpu().long(),
                                                     torch.from_numpy(rs_h & 3).long(), torch.from_numpy((rs_h >> 2) & 1).long(),
                                                     h_actions[:nr].clone()], dim=1))
                        recorded_masks.append((h_masks[:nr].long() * mask_weights).sum(1))
                cycles += 1
                continue
            if nr == 0 and n_live == 0:
                break
            if nr > 0:
                obs = env.encode_obs()[:nr]
                masks = env.masks[:nr]
                tbl = env.row_table[:nr].long()
                seat = (env.row_seat[:nr] & 3).long()
                chal = is_challenger[tbl % per, seat]
                for idx, agent in ((chal.nonzero().squeeze(1), agents[0]), ((~chal).nonzero().squeeze(1), agents[1])):
                    if idx.numel() == 0:
                        continue
                    t_eval = time.perf_counter_ns()
                    a, q = agent.react_device(obs[idx], masks[idx])
                    actions[idx] = a.to(torch.int64)
                    if q_all is not None:
                        q_all[idx] = q.float()
                    if meta_rec is not None:
                        meta_rec.add_agent(cycles, idx, q, time.perf_counter_ns() - t_eval)
                if meta_rec is not None:
                    meta_rec.add_rows(cycles, tbl, env.row_seat[:nr], actions[:nr], masks, obs)
                if self.record_decisions:
                    recorded.append(torch.stack([tbl, env.row_step[:nr].long(), seat, (env.row_seat[:nr] >> 2).long() & 1,
                                                 actions[:nr]], dim=1).cpu())
                    recorded_masks.append((masks.long() * mask_weights.to(dev)).sum(1).cpu())
            cycles += 1
        res = env.results()
        if self.log_dir is not None:  # one_vs_three.rs:195-225: one {seed}_{key}_{split}.json.gz per game
            from .. import mjai_log

            words, lens = env.read_log()
            ic = is_challenger.cpu().numpy()
            agent_names = [str(getattr(a, "name", "NoName")) for a in agents]
            names = [[agent_names[0] if ic[g % per, seat] else agent_names[1] for seat in range(4)] for g in range(n)]
            seeds = [(int(nonces[g]), int(keys[g])) for g in range(n)]
            bounds = decisions = None
            if meta_rec is not None:
                try:
                    bounds, decisions = meta_rec.finish()
                except Exception as exc:  # the logs themselves must not depend on the optional metadata
                    self.last_meta_error = exc
                    bounds = decisions = None
            self.last_log_paths = mjai_log.write_logs(self.log_dir, words, lens, seeds, names, "abcd"[:per], bounds, decisions)
        self.last_stats = dict(cycles=cycles, table_steps=int(res["steps"].sum()), sp_overflows=env.sp_overflows())
        self.last_results = res
        if self.record_decisions:
            self.last_decisions = torch.cat(recorded).numpy() if recorded else np.zeros((0, 5), dtype=np.int64)
            self.last_decision_masks = torch.cat(recorded_masks).numpy() if recorded_masks else np.zeros(0, dtype=np.int64)
        env.close()
        if self.last_stats["sp_overflows"]:
            raise RuntimeError("single-player state arena overflowed during the run: observation rows 889-1011 were zero in "
                               f"{self.last_stats['sp_overflows']} step(s)")
        if (res["err"] != 0).any():
            bad = int(np.nonzero(res["err"])[0][0])
            raise RuntimeError(f"table {bad} failed with mjx error code {int(res['err'][bad])}")
        return res

    def ako_vs_py(self, *a, **k):
        raise NotImplementedError("akochan subprocess agents are out of scope (SURVEY.md §2.1 row 5)")

    py_vs_ako = ako_vs_py


class OneVsThree(_Arena):
    GAMES_PER_SEED = 4

    def _challenger_seats(self, g):
        return [g % 4]

    def py_vs_py(self, challenger, champion, seed_start, seed_count):
        res = self._run(challenger, champion, seed_start, seed_count)
        rankings = [0, 0, 0, 0]
        for i in range(res["ranks"].shape[0]):  # one_vs_three.rs:55-60
            rankings[int(res["ranks"][i, i % 4])] += 1
        return rankings


class TwoVsTwo(_Arena):
    GAMES_PER_SEED = 2

    def _challenger_seats(self, g):
        return [0, 2] if g % 2 == 0 else [1, 3]  # two_vs_two.rs:137-191

    def py_vs_py(self, challenger, champion, seed_start, seed_count):
        self._run(challenger, champion, seed_start, seed_count)
        return None
