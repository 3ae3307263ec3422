"""CPU stand-in with the surface of mortal_b200.BatchEnv that the libriichi.arena mirror uses — TEST INFRASTRUCTURE.
It wraps the single-lane host build of the product's device sources (tests/host_emul) and hands out CPU torch tensors, so the
arena's host-protocol loop (and engines written against libriichi, e.g. the reference's own mortal/engine.py) can be exercised
in a container without a GPU. Injected through `_Arena.env_factory` by tests only; the product never imports it."""
import numpy as np
import torch

import emul_lib as E


class EmulBatchEnv:
    def __init__(self, nonces, keys, *, obs_version=4, shuffle_kind=0, enable_quick_eval=True, device=0):
        self._e = E.EmulEnv(nonces, keys, shuffle_kind=shuffle_kind, enable_quick_eval=enable_quick_eval)
        self.L = self._e.L
        self.device = torch.device("cpu")
        self.n_tables = self._e.n_tables
        self.row_cap = self._e.row_cap
        self.obs_version, self.obs_rows = obs_version, {1: 938, 2: 942, 3: 934, 4: 1012}[obs_version]
        self.masks = torch.zeros((self.row_cap, 46), dtype=torch.bool)
        self.row_table = torch.zeros(self.row_cap, dtype=torch.int32)
        self.row_seat = torch.zeros(self.row_cap, dtype=torch.uint8)
        self.row_step = torch.zeros(self.row_cap, dtype=torch.int64)
        self._steps = 0
        self._h = True
        self._ovf0 = int(self.L.emul_sp_overflows())

    def close(self):
        if self._h:
            self._e.close()
            self._h = None

    def step(self, actions=None, q_values=None):
        assert q_values is None, "the emulation harness has no agari-guard input"
        self._e.step(None if actions is None else actions.numpy())
        self._steps += self._e.num_live()
        n = self._e.num_rows()
        rt, rs, m = self._e.rows()
        st = np.zeros(max(n, 1), dtype=np.uint32)
        self.L.emul_env_row_steps(self._e._h, st.ctypes.data)
        self.row_table[:n] = torch.from_numpy(rt)
        self.row_seat[:n] = torch.from_numpy(rs)
        self.masks[:n] = torch.from_numpy(m)
        self.row_step[:n] = torch.from_numpy(st[:n].astype(np.int64))

    def poll(self):
        errs = np.zeros(self.n_tables, dtype=np.int32)
        self.L.emul_env_errs(self._e._h, errs.ctypes.data)
        return self._e.num_rows(), self._e.num_live(), int((errs != 0).sum()), self.sp_overflows()

    def num_rows(self):
        return self._e.num_rows()

    def num_live(self):
        return self._e.num_live()

    def total_steps(self):
        return self._steps

    def sp_overflows(self):
        return int(self.L.emul_sp_overflows()) - self._ovf0

    def set_obs_version(self, version):
        self.obs_version, self.obs_rows = version, {1: 938, 2: 942, 3: 934, 4: 1012}[version]

    def set_quick_eval(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        self.L.emul_env_set_quick_eval(self._e._h, f.ctypes.data)

    def set_agari_guard(self, flags):
        assert flags is None or not np.asarray(flags).any()

    def encode_obs_host(self, h_obs, h_masks):
        n = self._e.num_rows()
        if n:
            h_obs[:n] = torch.from_numpy(self._e.encode_obs(sp=self.obs_version == 4, version=self.obs_version))
            h_masks[:n] = self.masks[:n]
        return n

    def encode_invisible(self, version=4):
        n = self._e.num_rows()
        out = torch.zeros((self.row_cap, 211 if version == 1 else 217, 34), dtype=torch.float32)
        if n:
            out[:n] = torch.from_numpy(self._e.encode_invisible(version))
        return out

    def encode_obs_host_begin(self, h_obs, h_masks):
        return self.encode_obs_host(h_obs, h_masks)

    def encode_obs_host_finish(self):
        pass

    def policy_test(self, kind, actions):
        actions.copy_(torch.from_numpy(self._e.policy_test(kind)))

    def results(self):
        n = self.n_tables
        scores = np.zeros((n, 4), dtype=np.int32); ranks = np.zeros((n, 4), dtype=np.uint8)
        steps = np.zeros(n, dtype=np.int32); err = np.zeros(n, dtype=np.int32); done = np.zeros(n, dtype=np.int32)
        self.L.emul_env_results(self._e._h, scores.ctypes.data, ranks.ctypes.data, steps.ctypes.data, err.ctypes.data, done.ctypes.data)
        return dict(scores=scores, ranks=ranks, steps=steps, err=err, done=done)
