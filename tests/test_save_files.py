"""N2: the HDF5 save/restore contract of the reference, pinned with an in-memory h5py stand-in
(tests/fake_h5py.py).  Dataset names / shapes / dtypes follow qoc/models/schroedingermodels.py:58-110 (evolve),
:258-313 (grape) and qoc/models/lindbladmodels.py:60-99, :254-339.  CPU tests drive the program-state classes directly;
the GPU tests run the public programs end to end and check what landed in the file against the returned results."""
import numpy as np
import pytest

from tests import fake_h5py

GRAPE_COMMON = {"complex_controls", "control_count", "control_eval_count", "controls", "cost_eval_step", "cost_names", "error",
                "evolution_time", "grads", "initial_controls", "interpolation_policy", "iteration_count", "max_control_norms",
                "method", "optimizer", "program_type", "system_eval_count"}


def _schroedinger_state(path, save_intermediate, iteration_count=5, save_step=2, complex_controls=True):
    import qoc_b200.standard as std
    from qoc_b200.models import GrapeSchroedingerDiscreteState, InterpolationPolicy, MagnusPolicy
    init = np.eye(3, dtype=complex)[:2, :, None]
    costs = [std.TargetStateInfidelity(init), std.ForbidStates(init[:, None], 7)]
    ctl = np.zeros((4, 2), dtype=complex if complex_controls else float)
    return GrapeSchroedingerDiscreteState(complex_controls, 2, 4, 1, costs, 6.0, lambda c, t: None, None, ctl, init,
                                          InterpolationPolicy.LINEAR, iteration_count, 0, np.ones(2), MagnusPolicy.M4, 0.0,
                                          std.Adam(), path, save_intermediate, save_step, 7)


def test_grape_schroedinger_file_layout(monkeypatch):
    files = fake_h5py.install(monkeypatch)
    ps = _schroedinger_state("/x/run.h5", True)
    ps.log_and_save_initial()
    f = files["/x/run.h5"]
    assert set(f) == GRAPE_COMMON | {"final_states", "initial_states", "magnus_policy", "intermediate_states"}
    save_count = 3                                                       # iterations 0, 2, 4 (5 iterations, step 2)
    assert f["controls"].shape == (save_count, 4, 2) and f["controls"].dtype == np.complex128
    assert f["grads"].shape == (save_count, 4, 2) and f["grads"].dtype == np.complex128
    assert f["error"].shape == (save_count,) and np.all(f["error"] == np.finfo(np.float64).max)
    assert f["final_states"].shape == (save_count, 2, 3, 1) and f["final_states"].dtype == np.complex128
    assert f["intermediate_states"].shape == (save_count, 7, 2, 3, 1)
    assert f["initial_states"].shape == (2, 3, 1)
    assert [bytes(x) for x in f["cost_names"]] == [b"target_state_infidelity", b"forbid_states"]
    assert str(f["method"]) == "grape_schroedinger_discrete" and str(f["magnus_policy"]) == "magnus_m4"
    assert int(f["program_type"]) == 2 and int(f["system_eval_count"]) == 7 and float(f["evolution_time"]) == 6.0
    assert bool(f["complex_controls"]) is True and int(f["iteration_count"]) == 5
    # rows: iteration // save_iteration_step, also written on the final iteration; other iterations are skipped
    fin = np.arange(6, dtype=complex).reshape(2, 3, 1)
    traj = np.arange(7 * 6, dtype=complex).reshape(7, 2, 3, 1)
    for it in range(5):
        ps.log_and_save(np.full((4, 2), it + 1j), 0.5 / (it + 1), fin * (it + 1), np.full((4, 2), -it + 0j), it)
        ps.save_all_intermediate_states(it, traj + it)
    assert np.allclose(f["error"], [0.5, 0.5 / 3, 0.5 / 5])
    assert np.array_equal(f["controls"][1], np.full((4, 2), 2 + 1j)) and np.array_equal(f["grads"][2], np.full((4, 2), -4 + 0j))
    assert np.array_equal(f["final_states"][2], fin * 5) and np.array_equal(f["intermediate_states"][1], traj + 2)
    ps.log_and_save(np.zeros((4, 2)), 9.0, fin, np.zeros((4, 2)), 5)                      # beyond the last iteration: ignored
    assert np.allclose(f["error"], [0.5, 0.5 / 3, 0.5 / 5])


def test_grape_schroedinger_final_iteration_row_and_real_controls(monkeypatch):
    files = fake_h5py.install(monkeypatch)
    ps = _schroedinger_state("/x/r.h5", False, iteration_count=6, save_step=4, complex_controls=False)
    ps.log_and_save_initial()
    f = files["/x/r.h5"]
    assert "intermediate_states" not in f
    assert f["controls"].shape == (2, 4, 2) and f["controls"].dtype == np.float64     # 6 // 4 = 1, + 1 for the final iteration
    for it in range(6):
        ps.log_and_save(np.full((4, 2), float(it)), float(it), np.zeros((2, 3, 1)), np.zeros((4, 2)), it)
    assert np.allclose(f["error"], [0.0, 5.0])                                           # iteration 0 and the final one (5 // 4 = 1)


def test_evolve_files_and_lindblad_layout(monkeypatch):
    files = fake_h5py.install(monkeypatch)
    import qoc_b200.standard as std
    from qoc_b200.models import (EvolveLindbladDiscreteState, EvolveSchroedingerDiscreteState, GrapeLindbladDiscreteState,
                                 InterpolationPolicy, MagnusPolicy)
    init = np.eye(3, dtype=complex)[:2, :, None]
    ps = EvolveSchroedingerDiscreteState(4, 1, [std.TargetStateInfidelity(init)], 6.0, None, init, InterpolationPolicy.LINEAR,
                                         MagnusPolicy.M2, "/x/e.h5", True, 7)
    ps.save_initial(np.ones((4, 1)))
    f = files["/x/e.h5"]
    assert set(f) == {"controls", "cost_eval_step", "costs", "evolution_time", "initial_states", "interpolation_policy",
                      "intermediate_states", "magnus_policy", "method", "program_type", "system_eval_count"}
    assert f["intermediate_states"].shape == (7, 2, 3, 1) and str(f["method"]) == "evolve_schroedinger_discrete"
    traj = np.arange(7 * 6, dtype=complex).reshape(7, 2, 3, 1)
    ps.save_all_intermediate_states(0, traj)
    assert np.array_equal(f["intermediate_states"], traj)
    rho = np.stack([np.eye(3, dtype=complex) / 3] * 2)
    pl = EvolveLindbladDiscreteState(4, 1, [std.TargetDensityInfidelity(rho)], 6.0, None, rho, InterpolationPolicy.LINEAR, None,
                                     "/x/l.h5", True, 5)
    pl.save_initial(None)
    f = files["/x/l.h5"]
    assert f["intermediate_densities"].shape == (5, 2, 3, 3) and "initial_densities" in f
    dens = np.arange(5 * 18, dtype=complex).reshape(5, 2, 3, 3)
    pl.save_all_intermediate_densities(0, dens)
    assert np.array_equal(f["intermediate_densities"], dens)
    pg = GrapeLindbladDiscreteState(False, 1, 4, 1, [std.TargetDensityInfidelity(rho)], 6.0, None, None, np.zeros((4, 1)), rho,
                                    InterpolationPolicy.LINEAR, 3, None, 0, np.ones(1), 0.0, std.Adam(), "/x/g.h5", True, 1, 5)
    pg.log_and_save_initial()
    f = files["/x/g.h5"]
    assert set(f) == GRAPE_COMMON | {"final_densities", "initial_densities", "intermediate_densities"}
    assert f["final_densities"].shape == (3, 2, 3, 3) and f["intermediate_densities"].shape == (3, 5, 2, 3, 3)
    for it in range(3):
        pg.save_all_intermediate_densities(it, dens * (it + 1))
    assert np.array_equal(f["intermediate_densities"][2], dens * 3)


def test_save_path_without_h5py_is_loud():
    from qoc_b200.models import state
    if state.h5py is not None:
        pytest.skip("h5py is installed here")
    with pytest.raises(ImportError):
        _schroedinger_state("/x/run.h5", False)


@pytest.mark.gpu
def test_grape_programs_write_their_results(monkeypatch):
    """end to end on the GPU: what the public programs return is what they wrote, including every stored psi_j / rho_j."""
    files = fake_h5py.install(monkeypatch)
    import qoc_b200 as qoc
    import qoc_b200.standard as std
    from qoc_b200.models import MagnusPolicy
    from tests.problems import Problem
    p = Problem(4, 12, 1, 2, 4, complex_controls=True, seed=3)
    res = qoc.grape_schroedinger_discrete(1, p.M, p.costs(std), p.T, p.hamiltonian_numpy(), p.initial_states, p.N,
                                          complex_controls=True, initial_controls=p.controls, iteration_count=4,
                                          log_iteration_step=0, magnus_policy=MagnusPolicy.M4, save_file_path="/m/s.h5",
                                          save_intermediate_states=True, save_iteration_step=1)
    f = files["/m/s.h5"]
    best = int(np.argmin(f["error"]))
    assert f["error"][best] == res.best_error and np.array_equal(f["final_states"][best], res.best_final_states)
    assert np.array_equal(f["controls"][0], p.controls)
    traj = f["intermediate_states"]
    assert traj.shape == (4, p.N, 2, 4, 1)
    for it in range(4):
        assert np.allclose(traj[it, 0], p.initial_states) and np.array_equal(traj[it, -1], f["final_states"][it])
        assert np.abs(np.linalg.norm(traj[it, :, :, :, 0], axis=-1) - 1).max() < 1e-12
    ev = qoc.evolve_schroedinger_discrete(p.T, p.hamiltonian_numpy(), p.initial_states, p.N, controls=f["controls"][2],
                                          costs=p.costs(std), magnus_policy=MagnusPolicy.M4, save_file_path="/m/e.h5",
                                          save_intermediate_states=True)
    assert abs(ev.error - f["error"][2]) < 1e-14 and np.array_equal(files["/m/e.h5"]["intermediate_states"], traj[2])
    # Lindblad twin
    n = 2
    a = np.diag(np.sqrt(np.arange(1, n)), 1).astype(complex)
    h = lambda c, t: np.diag(np.arange(n) - 0.5).astype(complex) + c[0] * a + np.conjugate(c[0]) * a.T
    rho0 = np.zeros((1, n, n), dtype=complex); rho0[0, 0, 0] = 1
    targ = np.zeros((1, n, n), dtype=complex); targ[0, 1, 1] = 1
    res = qoc.grape_lindblad_discrete(1, 5, [std.TargetDensityInfidelity(targ)], 2.0, rho0, 4, complex_controls=True,
                                      hamiltonian=h, lindblad_data=lambda t: (np.array([0.05]), a[None]), iteration_count=3,
                                      log_iteration_step=0, save_file_path="/m/l.h5", save_intermediate_densities=True,
                                      save_iteration_step=1)
    f = files["/m/l.h5"]
    dens = f["intermediate_densities"]
    assert dens.shape == (3, 4, 1, n, n)
    for it in range(3):
        assert np.allclose(dens[it, 0], rho0) and np.allclose(dens[it, -1], f["final_densities"][it], atol=1e-14)
        assert np.abs(np.trace(dens[it, :, 0], axis1=-2, axis2=-1) - 1).max() < 1e-10
    assert f["error"][int(np.argmin(f["error"]))] == res.best_error
