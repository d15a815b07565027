"""CPU: the re-association used by the CUDA kernel (blocked exact sweep, (1+g/4) operands, fixed-point
panel sums, invariant 1'e, Gram corrections) reproduces the oracle's sequential chain."""
import numpy as np
import pytest

from blocked_model import BlockedModel
from common import make_problem, oracle_chain


@pytest.mark.parametrize("method,B,T", [(0, 64, 3), (1, 32, 2), (2, 64, 5)])
def test_blocked_arithmetic_equals_sequential_sweep(method, B, T):
    prob = make_problem(333, 150, 21 + method)
    ro = np.array([0, 30, 95, 150]) if method == 0 else None
    ch, S = oracle_chain(prob, method, 0.02, pi=0.1, est_pi=True, region_off=ro, v_e=1.0)
    m = BlockedModel(prob["codes"], prob["y"], method, 0.02, pi=0.1, est_pi=True, region_off=ro, v_e=1.0, B=B, T=T)
    for _ in range(15):
        log = ch.iteration(seed=11, chain=1)
        m.iteration(log)
        assert np.abs(m.beta - S.beta).max() <= 1e-10 * max(np.abs(S.beta).max(), 1e-300)
        assert abs(m.varE / ch.varE - 1) < 1e-11
        assert (m.delta == S.delta).all()
        assert np.allclose(m.e, ch.e, rtol=0, atol=1e-10 * np.abs(ch.e).max())
