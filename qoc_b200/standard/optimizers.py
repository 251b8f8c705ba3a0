"""
optimizers.py - host optimisers with the reference's `run(function, iteration_count, initial_params, jacobian,
args)` protocol (qoc/standard/optimizers/adam.py:83-165, lbfgsb.py:21-49, sgd.py:25-59).  They stay on the
host by design: one cost+gradient evaluation per iteration is the GPU call.
"""
import numpy as np
from scipy.optimize import minimize

from qoc_b200.models.enums import OperationPolicy


class Adam(object):
    """Adam (https://arxiv.org/abs/1412.6980) with optional gradient scaling, clipping and
    exponential learning-rate decay, applied in that order as in the reference (adam.py:110-165)."""
    name = "adam"

    def __init__(self, beta_1=0.9, beta_2=0.999, clip_grads=None, epsilon=1e-8, learning_rate=1e-3,
                 learning_rate_decay=None, operation_policy=OperationPolicy.CPU, scale_grads=None):
        self.apply_scale_grads = scale_grads is not None
        self.apply_clip_grads = clip_grads is not None
        self.apply_learning_rate_decay = learning_rate_decay is not None
        self.beta_1, self.beta_2 = beta_1, beta_2
        self.clip_grads = clip_grads
        self.epsilon = epsilon
        self.gradient_moment = None
        self.gradient_square_moment = None
        self.initial_learning_rate = learning_rate
        self.iteration_count = 0
        self.learning_rate = learning_rate
        self.learning_rate_decay = learning_rate_decay
        self.scale_grads = scale_grads

    def __str__(self):
        return ("{}, beta_1: {}, beta_2: {}, epsilon: {}, lr0: {}, lr_decay: {}, clip_grads: {}, scale_grads: {}"
                "".format(self.name, self.beta_1, self.beta_2, self.epsilon, self.initial_learning_rate,
                          self.learning_rate_decay, self.clip_grads, self.scale_grads))

    def run(self, function, iteration_count, initial_params, jacobian, args=()):
        self.iteration_count = 0
        self.gradient_moment = np.zeros_like(initial_params)
        self.gradient_square_moment = np.zeros_like(initial_params)
        params = initial_params
        for _ in range(iteration_count):
            grads, terminate = jacobian(params, *args)
            if terminate:
                break
            params = self.update(grads, params)

    def update(self, grads, params):
        lr = self.initial_learning_rate
        if self.apply_learning_rate_decay:
            lr = lr * np.exp(-np.divide(self.iteration_count, self.learning_rate_decay))
        if self.apply_scale_grads:
            grads = grads / np.linalg.norm(grads) * self.scale_grads
        if self.apply_clip_grads:
            grads = np.clip(grads, -self.clip_grads, self.clip_grads)
        self.iteration_count += 1
        t = self.iteration_count
        self.gradient_moment = self.beta_1 * self.gradient_moment + (1 - self.beta_1) * grads
        self.gradient_square_moment = self.beta_2 * self.gradient_square_moment + (1 - self.beta_2) * np.square(grads)
        m_hat = self.gradient_moment / (1 - np.power(self.beta_1, t))
        v_hat = self.gradient_square_moment / (1 - np.power(self.beta_2, t))
        return params - lr * m_hat / (np.sqrt(v_hat) + self.epsilon)


class LBFGSB(object):
    """scipy's L-BFGS-B; like the reference it ignores the `terminate` flags (lbfgsb.py:36-49)."""

    def __str__(self):
        return "lbfgsb"

    def run(self, function, iteration_count, initial_params, jacobian, args=()):
        return minimize(lambda *a, **k: function(*a, **k)[0], initial_params, args=args, method="L-BFGS-B",
                        jac=lambda *a, **k: jacobian(*a, **k)[0], options={"maxiter": iteration_count})


class SGD(object):
    name = "sgd"

    def __init__(self, learning_rate=1e-3):
        self.learning_rate = learning_rate

    def __str__(self):
        return "{}, lr: {}".format(self.name, self.learning_rate)

    def run(self, function, iteration_count, initial_params, jacobian, args=()):
        params = initial_params
        for _ in range(iteration_count):
            grads, terminate = jacobian(params, *args)
            if terminate:
                break
            params = self.update(grads, params)

    def update(self, grads, params):
        return params - self.learning_rate * grads
