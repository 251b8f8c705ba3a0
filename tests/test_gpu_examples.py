"""GPU: the reference's two examples (BASELINE.json configs[0] and [1]) through the public API, and the reference's own
GRAPE smoke tests (tests/test_core.py:563-602, :247-290: with max_control_norms = 1e-10 the optimised controls must
respect the bound)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_example_transmon_pi_optimises(capsys):
    from examples import transmon_pi
    res = transmon_pi.main(iteration_count=300, log_iteration_step=100)
    out = capsys.readouterr().out
    assert "iter" in out and "total error" in out and "grads_l2" in out          # the reference's log header
    assert res.best_error < 0.2 and res.best_iteration > 0                         # from 0.0488... flat start: error falls
    assert res.best_controls.shape == (11, 1) and res.best_final_states.shape == (1, 2, 1)


def test_example_transmon_pi_decoherence_optimises():
    from examples import transmon_pi_decoherence
    res = transmon_pi_decoherence.main(iteration_count=15, log_iteration_step=0)
    assert res.best_error < 0.75 and res.best_final_densities.shape == (1, 2, 2)


def test_reference_grape_smoke_bounds():
    import qoc_b200 as qoc
    from qoc_b200.models import MagnusPolicy
    from qoc_b200.standard import Adam, TargetStateInfidelity
    rng = np.random.default_rng(0)
    n, K = 4, 2
    h0 = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); h0 = (h0 + h0.conj().T) / 2
    hc = [rng.standard_normal((n, n)) for _ in range(K)]
    hc = [(h + h.T) / 2 for h in hc]
    ham = lambda c, t: h0 + c[0] * hc[0] + c[1] * hc[1]
    init = np.eye(n, dtype=complex)[:, :1].T[:, :, None]
    targ = np.eye(n, dtype=complex)[:, 1:2].T[:, :, None]
    mx = np.repeat(1e-10, K)
    for pol in (MagnusPolicy.M2, MagnusPolicy.M4, MagnusPolicy.M6):
        res = qoc.grape_schroedinger_discrete(K, 10, [TargetStateInfidelity(targ)], 1.0, ham, init, 10, iteration_count=5,
                                              log_iteration_step=0, magnus_policy=pol, max_control_norms=mx, optimizer=Adam())
        assert np.less_equal(np.abs(res.best_controls), mx + 1e-17).all()
