"""libriichi.state.PlayerState / libriichi.mjai.Bot on the CUDA path (include/mjx.h mjx_state_*): the reference's state/test.rs
assertions (tests/state_cases.py), the single-seat observation against the oracle's PlayerState, and the Bot on the golden game —
the same bodies tests/test_emul_state.py runs on the host emulation, here through libmjx.so."""
import pytest

import state_cases as SC
import test_emul_state as TE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def PS():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from mortal_b200.libriichi import state

    state.set_backend(None)  # the default: the CUDA library
    assert type(state.get_backend()).__name__ == "_CudaBackend"
    return state.PlayerState


@pytest.mark.parametrize("case", SC.ALL_CASES, ids=lambda c: c.__name__)
def test_state_test_rs_cases_on_cuda(PS, case):
    case(PS)


def test_partial_information_obs_equals_oracle_player_state_on_cuda(PS):
    TE.test_partial_information_obs_equals_oracle_player_state(PS)


def test_mjai_bot_on_cuda(PS):
    TE.test_mjai_bot_plays_a_seat_of_a_logged_game(PS)
