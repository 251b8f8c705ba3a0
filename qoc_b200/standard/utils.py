"""
utils.py - file-path and JSON helpers (qoc/standard/utils/fileutil.py, jsonutil.py) and the HIPS-autograd
bridge: when `autograd` is importable the GPU evaluation is exposed as an autograd primitive with a VJP
(`make_autograd_primitive`); otherwise only the plain (value, gradient) call exists.
"""
import json
import os

import numpy as np


def generate_save_file_path(save_file_name, save_path):
    """`<save_path>/<5-digit running number>_<name>.h5`, numbering continues after the largest existing one."""
    os.makedirs(save_path, exist_ok=True)
    max_numeric_prefix = -1
    for file_name in os.listdir(save_path):
        if ("_{}.h5".format(save_file_name)) in file_name:
            try:
                max_numeric_prefix = max(int(file_name.split("_")[0]), max_numeric_prefix)
            except ValueError:
                pass
    return os.path.join(save_path, "{:05d}_{}.h5".format(max_numeric_prefix + 1, save_file_name))


class CustomJSONEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, (np.integer,)):
            return int(obj)
        if isinstance(obj, (np.floating,)):
            return float(obj)
        if isinstance(obj, complex):
            return (obj.real, obj.imag)
        return super().default(obj)


def make_autograd_primitive(value_and_grad):
    """Wrap `value_and_grad(controls) -> (value, grad_autograd_convention)` as a HIPS-autograd primitive with a VJP
    (`autograd.extend.primitive` + `defvjp`), the contract BASELINE.json's north_star names.  grad_autograd_convention is
    dE/dx - i dE/dy for complex controls (what autograd's make_vjp returns, qoc/core/schroedingerdiscrete.py:320-322) and
    the plain real gradient for real controls.

    The gradient is captured PER CALL: autograd invokes the vjp-maker right after the forward evaluation of the same call,
    so the maker takes the gradient that evaluation produced (matched by the controls' bytes; re-evaluated if something else
    ran in between) and closes over it.  Several forward calls followed by one backward pass therefore each keep their own
    gradient."""
    from autograd.extend import defvjp, primitive        # raises ImportError when autograd is absent
    pending = {}

    def _key(controls):
        a = np.ascontiguousarray(np.asarray(controls))
        return (a.shape, a.dtype.str, a.tobytes())

    @primitive
    def gpu_cost(controls):
        controls = np.asarray(controls)
        value, grad = value_and_grad(controls)
        pending[_key(controls)] = np.array(grad, copy=True)
        while len(pending) > 8:                          # forward-only calls never collect their entry
            pending.pop(next(iter(pending)))
        return value

    def vjp_maker(ans, controls):
        grad = pending.pop(_key(controls), None)
        if grad is None:
            grad = np.array(value_and_grad(np.asarray(controls))[1], copy=True)
        return lambda g: g * grad

    defvjp(gpu_cost, vjp_maker)
    return gpu_cost


def ans_jacobian(function, argnum=0):
    """drop-in for `qoc.standard.utils.autogradutil.ans_jacobian` (:10-31): returns a function giving
    (ans, jacobian of ans with respect to positional argument `argnum`) through HIPS autograd's `make_vjp`, one pull-back
    per basis vector of the output space (one for the scalar cost).  `function` may contain GPU-backed primitives made by
    `make_autograd_primitive`.  Needs `autograd` (ImportError otherwise); the product's own evaluation path calls the C
    ABI's value-and-gradient entry point directly and does not depend on it."""
    from autograd.core import make_vjp
    from autograd.extend import vspace

    def value_and_jacobian(*args, **kwargs):
        def of_one_argument(x):
            full = list(args)
            full[argnum] = x
            return function(*full, **kwargs)
        pullback, ans = make_vjp(of_one_argument, args[argnum])
        out_space = vspace(ans)
        rows = [pullback(basis_vector) for basis_vector in out_space.standard_basis()]
        return ans, np.stack(rows).reshape(out_space.shape + vspace(args[argnum]).shape)
    return value_and_jacobian


def autograd_available():
    try:
        import autograd.extend  # noqa: F401
        return True
    except ImportError:
        return False
