"""
common.py - control-array plumbing shared by the grape programs (host side of the drop-in boundary;
behaviour of qoc/core/common.py:8-30, :33-142, :146-198, :201-246).
"""
import numpy as np

_NORM_TOLERANCE = 1e-10


def clip_control_norms(controls, max_control_norms):
    """rescale, IN PLACE, every control point whose modulus exceeds its column's maximum norm back onto
    that norm (for real controls: clipping to [-max, max])."""
    for k, max_norm in enumerate(max_control_norms):
        column = controls[:, k]
        norms = np.abs(column)
        over = np.nonzero(max_norm < norms)
        column[over] = column[over] / norms[over] * max_norm


def _complexify(controls, complex_controls):
    return (controls - 1j * controls) / np.sqrt(2) if complex_controls else controls


def gen_controls_cos(complex_controls, control_count, control_eval_count, evolution_time,
                     max_control_norms, periods=10.):
    """half-amplitude cosine with `periods` periods over the pulse; exact zeros replaced by max/10."""
    b = 2 * np.pi / (control_eval_count / periods)
    controls = np.zeros((control_eval_count, control_count))
    for k in range(control_count):
        wave = max_control_norms[k] / 2 * np.cos(b * np.arange(control_eval_count))
        controls[:, k] = np.where(wave, wave, max_control_norms[k] * 1e-1)
    return _complexify(controls, complex_controls)


def gen_controls_white(complex_controls, control_count, control_eval_count, evolution_time,
                       max_control_norms, periods=10.):
    """white noise with standard deviation max/5."""
    controls = np.zeros((control_eval_count, control_count))
    for k in range(control_count):
        controls[:, k] = np.random.normal(0, max_control_norms[k] / 5.0, control_eval_count)
    return _complexify(controls, complex_controls)


def gen_controls_flat(complex_controls, control_count, control_eval_count, evolution_time,
                      max_control_norms, periods=10.):
    """flat line at a tenth of the maximum norm (the default initial guess)."""
    controls = np.zeros((control_eval_count, control_count))
    for k in range(control_count):
        controls[:, k] = np.repeat(max_control_norms[k] * 1e-1, control_eval_count)
    return _complexify(controls, complex_controls)


def initialize_controls(complex_controls, control_count, control_eval_count, evolution_time,
                        initial_controls, max_control_norms):
    """default max norms (ones) / default flat controls, and validation of user-supplied controls against the
    declared dtype and the max norms (ValueError, as the reference)."""
    if max_control_norms is None:
        max_control_norms = np.ones(control_count)
    if initial_controls is None:
        return (gen_controls_flat(complex_controls, control_count, control_eval_count, evolution_time,
                                  max_control_norms), max_control_norms)
    is_complex = np.iscomplexobj(initial_controls)
    if complex_controls and not is_complex:
        raise ValueError("The program expected that the initial_controls specified by the user conformed to "
                         "complex_controls, but the program found that the initial_controls were not complex "
                         "and complex_controls was set to True.")
    if not complex_controls and is_complex:
        raise ValueError("The program expected that the initial_controls specified by the user conformed to "
                         "complex_controls, but the program found that the initial_controls were complex "
                         "and complex_controls was set to False.")
    for step, row in enumerate(initial_controls):
        if not np.less_equal(np.abs(row), max_control_norms + _NORM_TOLERANCE).all():
            raise ValueError("The program expected that the initial_controls specified by the user conformed to "
                             "max_control_norms, but the program found a conflict at initial_controls[{}]={} and "
                             "max_control_norms={}.".format(step, row, max_control_norms))
    return initial_controls, max_control_norms


def slap_controls(complex_controls, controls, controls_shape):
    """optimiser format (flat float64; complex as [real..., imag...]) -> cost-function format.  For real
    controls the result is a reshape VIEW, so the in-place clip reaches the optimiser's array, as in the
    reference."""
    if complex_controls:
        real, imag = np.split(controls, 2)
        controls = real + 1j * imag
    return np.reshape(controls, controls_shape)


def strip_controls(complex_controls, controls):
    """cost-function format -> optimiser format."""
    flat = np.ravel(controls)
    if complex_controls:
        flat = np.hstack((np.real(flat), np.imag(flat)))
    return flat
