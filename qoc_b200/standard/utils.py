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
    """Wrap `value_and_grad(controls) -> (value, grad_autograd_convention)` as a HIPS-autograd primitive.
    grad_autograd_convention is dE/dx - i dE/dy for complex controls (what autograd's make_vjp returns,
    qoc/core/schroedingerdiscrete.py:320-322) and the plain real gradient for real controls."""
    from autograd.extend import defvjp, primitive        # raises ImportError when autograd is absent
    cache = {}

    @primitive
    def gpu_cost(controls):
        value, grad = value_and_grad(np.asarray(controls))
        cache["grad"] = grad
        return value

    defvjp(gpu_cost, lambda ans, controls: (lambda g: g * cache["grad"]))
    return gpu_cost


def ans_jacobian(value_and_grad, argnum=0):
    """drop-in for qoc.standard.utils.autogradutil.ans_jacobian on a GPU-backed evaluation: returns a function
    giving (value, jacobian) with the jacobian in autograd's convention."""
    def wrapped(*args):
        return value_and_grad(*args)
    return wrapped
