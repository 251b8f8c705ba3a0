"""A minimal stand-in for the parts of HIPS autograd the drop-in boundary touches (autograd is absent from this image and
from the GPU boxes): `autograd.extend.primitive / defvjp / vspace` and `autograd.core.make_vjp`, enough to trace ONE scalar
primitive call and pull a cotangent back through its registered VJP - the contract `make_autograd_primitive` and
`ans_jacobian` rely on (the vjp-maker runs at forward time, right after the primitive's evaluation).  Test infrastructure."""
import sys
import types

import numpy as np


class _Space(object):
    def __init__(self, value):
        self.shape = np.shape(value)
        self.dtype = np.asarray(value).dtype

    def standard_basis(self):
        if self.shape == ():
            yield 1.0
            return
        for idx in np.ndindex(*self.shape):
            e = np.zeros(self.shape, dtype=self.dtype)
            e[idx] = 1.0
            yield e


def install(monkeypatch):
    """put fake `autograd`, `autograd.extend`, `autograd.core` modules into sys.modules; returns the registry of VJPs."""
    registry, trace = {}, []

    def primitive(f):
        def wrapped(*args):
            ans = f(*args)
            trace.append((wrapped, ans, args))
            return ans
        wrapped.__wrapped__ = f
        return wrapped

    def defvjp(f, *makers):
        registry[f] = makers

    def make_vjp(fun, x):
        del trace[:]
        ans = fun(x)
        assert len(trace) == 1, "the fake tracer follows exactly one primitive call"
        prim, p_ans, p_args = trace[0]
        pull = registry[prim][0](p_ans, *p_args)          # vjp-maker runs at forward time
        return pull, ans

    ext = types.ModuleType("autograd.extend")
    ext.primitive, ext.defvjp, ext.vspace = primitive, defvjp, _Space
    core = types.ModuleType("autograd.core")
    core.make_vjp = make_vjp
    ag = types.ModuleType("autograd")
    ag.extend, ag.core = ext, core
    for name, mod in (("autograd", ag), ("autograd.extend", ext), ("autograd.core", core)):
        monkeypatch.setitem(sys.modules, name, mod)
    return registry
