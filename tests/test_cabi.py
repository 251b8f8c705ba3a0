"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/qocb200.h declares;
without a CUDA device every compute entry point fails loudly (no CPU fallback exists)."""
import ctypes
import os
import re

import numpy as np
import pytest

from qoc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda():
    import torch
    return torch.cuda.is_available()


def test_library_builds_and_exports_header_symbols():
    _lib.build_library()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "qocb200.h")).read()
    declared = sorted(set(re.findall(r"\b(qocb_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(declared) == sorted(_lib.EXPORTS)
    assert b"sm_100a" in lib.qocb_version()


def test_problem_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "qocb200.h")).read()
    body = header[header.index("typedef struct {"):header.index("} qocb_problem;")]
    fields = re.findall(r"(?:int32_t|double)\s+([a-z_0-9]+);", body)
    assert fields == [f[0] for f in _lib.Problem._fields_]


@pytest.mark.skipif(_cuda(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = _lib.load()
    pb = _lib.Problem(hilbert_size=4, state_count=1, control_count=1, control_eval_count=5, system_eval_count=5,
                      magnus_order=2, cost_eval_step=1, ensemble_count=1, device=0, store_tape=1,
                      chunks_per_member=0, channel_count=0, evolution_time=1.0)
    handle = ctypes.c_void_p()
    rc = lib.qocb_plan_create(ctypes.byref(pb), ctypes.byref(handle))
    assert rc != 0 and not handle
    assert b"no CUDA device" in lib.qocb_last_error(None)
    from qoc_b200.standard.functions import expm
    with pytest.raises(RuntimeError):
        expm(np.eye(2, dtype=complex))
    import qoc_b200 as qoc
    from qoc_b200.standard import TargetStateInfidelity
    init = np.array([[[1], [0]]], dtype=complex)
    with pytest.raises(RuntimeError):
        qoc.evolve_schroedinger_discrete(1.0, lambda c, t: np.eye(2), init, 5, costs=[TargetStateInfidelity(init)])


def test_bad_arguments_rejected():
    lib = _lib.load()
    handle = ctypes.c_void_p()
    pb = _lib.Problem(hilbert_size=4, state_count=1, control_count=1, control_eval_count=5, system_eval_count=5,
                      magnus_order=3, cost_eval_step=1, ensemble_count=1, device=0, store_tape=1,
                      chunks_per_member=0, channel_count=0, evolution_time=1.0)
    assert lib.qocb_plan_create(ctypes.byref(pb), ctypes.byref(handle)) == -1
    assert b"magnus_order" in lib.qocb_last_error(None)
    pb.magnus_order = 2
    pb.system_eval_count = 1
    assert lib.qocb_plan_create(ctypes.byref(pb), ctypes.byref(handle)) == -1
    assert lib.qocb_plan_create(None, ctypes.byref(handle)) == -1
