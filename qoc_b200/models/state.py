"""
state.py - program-state and result containers of the four public programs (mirror of
qoc/models/programstate.py, schroedingermodels.py, lindbladmodels.py, dummy.py).

Only the fields the hot-path seam reads (qoc/core/schroedingerdiscrete.py:371-388,
qoc/core/lindbladdiscrete.py:373-390) and the logging/saving behaviour are kept.  The HDF5 layout follows
qoc/models/schroedingermodels.py:258-313 / lindbladmodels.py:254-339 and needs `h5py` (import-guarded: it is
absent from the build image; asking for a save file without it raises ImportError).
"""
import numpy as np

from .enums import ProgramType

try:                                    # optional: on-disk contract of the reference
    import h5py
    from filelock import FileLock, Timeout
except ImportError:                     # pragma: no cover - h5py is not installed in the build image
    h5py = None
    FileLock = Timeout = None


class Dummy(object):
    """attribute bag used as the mutable reporter (qoc/models/dummy.py)."""
    pass


def _require_h5(path):
    if path is not None and h5py is None:
        raise ImportError("save_file_path was given but h5py/filelock are not installed")


class ProgramState(object):
    """fields shared by all programs (qoc/models/programstate.py:11-61)."""

    def __init__(self, control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian,
                 interpolation_policy, program_type, save_file_path, system_eval_count):
        _require_h5(save_file_path)
        self.control_eval_count = control_eval_count
        self.control_eval_times = np.linspace(0, evolution_time, control_eval_count)
        self.cost_eval_step = cost_eval_step
        self.costs = costs
        self.dt = evolution_time / (system_eval_count - 1)
        self.evolution_time = evolution_time
        self.final_system_eval_step = system_eval_count - 1
        self.hamiltonian = hamiltonian
        self.interpolation_policy = interpolation_policy
        self.program_type = program_type
        self.save_file_lock_path = "{}.lock".format(save_file_path)
        self.save_file_path = save_file_path
        self.step_cost_indices = [i for i, c in enumerate(costs) if c.requires_step_evaluation]
        self.step_costs = [costs[i] for i in self.step_cost_indices]
        self.system_eval_count = system_eval_count

    # -- HDF5 helpers -------------------------------------------------------------------------------
    def _write(self, mode, writer, what):
        try:
            with FileLock(self.save_file_lock_path):
                with h5py.File(self.save_file_path, mode) as f:
                    writer(f)
        except Timeout:
            print("Timeout while locking {} ({}).".format(self.save_file_lock_path, what))


class GrapeState(ProgramState):
    """fields shared by the grape programs (qoc/models/programstate.py:64-133)."""

    def __init__(self, complex_controls, control_count, control_eval_count, cost_eval_step, costs,
                 evolution_time, hamiltonian, impose_control_conditions, initial_controls,
                 interpolation_policy, iteration_count, log_iteration_step, max_control_norms, min_error,
                 optimizer, save_file_path, save_iteration_step, system_eval_count):
        super().__init__(control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian,
                         interpolation_policy, ProgramType.GRAPE, save_file_path, system_eval_count)
        self.complex_controls = complex_controls
        self.control_count = control_count
        self.controls_shape = (control_eval_count, control_count)
        self.final_iteration = iteration_count - 1
        self.impose_control_conditions = impose_control_conditions
        self.initial_controls = initial_controls
        self.iteration_count = iteration_count
        self.log_iteration_step = log_iteration_step
        self.max_control_norms = max_control_norms
        self.min_error = min_error
        self.optimizer = optimizer
        self.save_iteration_step = save_iteration_step
        self.should_log = log_iteration_step != 0
        self.should_save = (save_iteration_step != 0) and (save_file_path is not None)

    def _due(self, iteration, step):
        return np.mod(iteration, step) == 0 or iteration == self.final_iteration

    def _log_and_save(self, controls, error, finals, finals_key, grads, iteration):
        """stdout row `iter | total error | grads_l2` and the HDF5 row of this iteration
        (qoc/models/schroedingermodels.py:208-251)."""
        if iteration > self.final_iteration:
            return
        if self.should_log and self._due(iteration, self.log_iteration_step):
            print("{:^6d} | {:^1.8e} | {:^1.8e}".format(iteration, error, np.linalg.norm(grads)))
        if self.should_save and self._due(iteration, self.save_iteration_step):
            row = iteration // self.save_iteration_step

            def writer(f):
                f["controls"][row, ] = controls
                f["error"][row, ] = error
                f[finals_key][row, ] = finals
                f["grads"][row, ] = grads
            self._write("a", writer, "save after iteration {}".format(iteration))

    def _initial_header(self, f, finals_key, finals_shape):
        save_count, rem = divmod(self.iteration_count, self.save_iteration_step)
        if rem != 0:
            save_count += 1
        ctl_dtype = self.initial_controls.dtype
        f["complex_controls"] = self.complex_controls
        f["control_count"] = self.control_count
        f["control_eval_count"] = self.control_eval_count
        f["controls"] = np.zeros((save_count, self.control_eval_count, self.control_count), dtype=ctl_dtype)
        f["cost_eval_step"] = self.cost_eval_step
        f["cost_names"] = np.array([np.bytes_("{}".format(c)) for c in self.costs])
        f["error"] = np.repeat(np.finfo(np.float64).max, save_count)
        f["evolution_time"] = self.evolution_time
        f[finals_key] = np.zeros((save_count,) + tuple(finals_shape), dtype=np.complex128)
        f["grads"] = np.zeros((save_count, self.control_eval_count, self.control_count), dtype=ctl_dtype)
        f["initial_controls"] = self.initial_controls
        f["interpolation_policy"] = "{}".format(self.interpolation_policy)
        f["iteration_count"] = self.iteration_count
        f["max_control_norms"] = self.max_control_norms
        f["method"] = self.method
        f["optimizer"] = "{}".format(self.optimizer)
        f["program_type"] = self.program_type.value
        f["system_eval_count"] = self.system_eval_count
        return save_count

    def _print_header(self):
        if self.should_log:
            print("iter   |   total error  |    grads_l2   \n"
                  "=========================================")


# --- Schroedinger ----------------------------------------------------------------------------------
class EvolveSchroedingerDiscreteState(ProgramState):
    method = "evolve_schroedinger_discrete"

    def __init__(self, control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian, initial_states,
                 interpolation_policy, magnus_policy, save_file_path, save_intermediate_states_,
                 system_eval_count):
        super().__init__(control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian,
                         interpolation_policy, ProgramType.EVOLVE, save_file_path, system_eval_count)
        self.initial_states = initial_states
        self.magnus_policy = magnus_policy
        self.save_intermediate_states_ = save_file_path is not None and save_intermediate_states_

    def save_initial(self, controls):
        if self.save_file_path is None:
            return
        print("QOC is saving this evolution to {}.".format(self.save_file_path))

        def writer(f):
            f["controls"] = controls
            f["cost_eval_step"] = self.cost_eval_step
            f["costs"] = np.array([np.bytes_("{}".format(c)) for c in self.costs])
            f["evolution_time"] = self.evolution_time
            f["initial_states"] = self.initial_states
            f["interpolation_policy"] = "{}".format(self.interpolation_policy)
            if self.save_intermediate_states_:
                f["intermediate_states"] = np.zeros((self.system_eval_count,) + self.initial_states.shape,
                                                    dtype=np.complex128)
            f["magnus_policy"] = "{}".format(self.magnus_policy)
            f["method"] = self.method
            f["program_type"] = self.program_type.value
            f["system_eval_count"] = self.system_eval_count
        self._write("w", writer, "initial save")

    def save_all_intermediate_states(self, iteration, all_states):
        """one write of the whole [N][S][n][1] trajectory (the reference writes it step by step from inside
        its time loop, qoc/core/schroedingerdiscrete.py:395-402; the device keeps every psi_j anyway)."""
        if self.save_file_path is None:
            return

        def writer(f):
            f["intermediate_states"][...] = all_states.astype(np.complex128)
        self._write("a", writer, "intermediate states")


class EvolveSchroedingerResult(object):
    def __init__(self, error=None, final_states=None):
        self.error = error
        self.final_states = final_states


class GrapeSchroedingerDiscreteState(GrapeState):
    method = "grape_schroedinger_discrete"

    def __init__(self, complex_controls, control_count, control_eval_count, cost_eval_step, costs,
                 evolution_time, hamiltonian, impose_control_conditions, initial_controls, initial_states,
                 interpolation_policy, iteration_count, log_iteration_step, max_control_norms, magnus_policy,
                 min_error, optimizer, save_file_path, save_intermediate_states_, save_iteration_step,
                 system_eval_count):
        super().__init__(complex_controls, control_count, control_eval_count, cost_eval_step, costs,
                         evolution_time, hamiltonian, impose_control_conditions, initial_controls,
                         interpolation_policy, iteration_count, log_iteration_step, max_control_norms,
                         min_error, optimizer, save_file_path, save_iteration_step, system_eval_count)
        self.hilbert_size = initial_states[0].shape[0]
        self.initial_states = initial_states
        self.magnus_policy = magnus_policy
        self.save_intermediate_states_ = self.should_save and save_intermediate_states_

    def log_and_save(self, controls, error, final_states, grads, iteration):
        self._log_and_save(controls, error, final_states, "final_states", grads, iteration)

    def log_and_save_initial(self):
        if self.should_save:
            print("QOC is saving this optimization run to {}.".format(self.save_file_path))

            def writer(f):
                save_count = self._initial_header(f, "final_states",
                                                  (len(self.initial_states), self.hilbert_size, 1))
                f["initial_states"] = self.initial_states
                f["magnus_policy"] = "{}".format(self.magnus_policy)
                if self.save_intermediate_states_:
                    f["intermediate_states"] = np.zeros((save_count, self.system_eval_count)
                                                        + self.initial_states.shape, dtype=np.complex128)
            self._write("w", writer, "initial save")
        self._print_header()

    def save_all_intermediate_states(self, iteration, all_states):
        if iteration > self.final_iteration or not self.should_save:
            return
        if self._due(iteration, self.save_iteration_step):
            row = iteration // self.save_iteration_step

            def writer(f):
                f["intermediate_states"][row] = all_states.astype(np.complex128)
            self._write("a", writer, "intermediate states of iteration {}".format(iteration))


class GrapeSchroedingerResult(object):
    def __init__(self, best_controls=None, best_error=np.finfo(np.float64).max, best_final_states=None,
                 best_iteration=None):
        self.best_controls = best_controls
        self.best_error = best_error
        self.best_final_states = best_final_states
        self.best_iteration = best_iteration


# --- Lindblad ----------------------------------------------------------------------------------------
class EvolveLindbladDiscreteState(ProgramState):
    method = "evolve_lindblad_discrete"

    def __init__(self, control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian,
                 initial_densities, interpolation_policy, lindblad_data, save_file_path,
                 save_intermediate_densities_, system_eval_count):
        super().__init__(control_eval_count, cost_eval_step, costs, evolution_time, hamiltonian,
                         interpolation_policy, ProgramType.EVOLVE, save_file_path, system_eval_count)
        self.initial_densities = initial_densities
        self.lindblad_data = lindblad_data
        self.save_intermediate_densities_ = save_file_path is not None and save_intermediate_densities_

    def save_initial(self, controls):
        if self.save_file_path is None:
            return
        print("QOC is saving this evolution to {}.".format(self.save_file_path))

        def writer(f):
            f["controls"] = controls
            f["cost_eval_step"] = self.cost_eval_step
            f["costs"] = np.array([np.bytes_("{}".format(c)) for c in self.costs])
            f["evolution_time"] = self.evolution_time
            f["initial_densities"] = self.initial_densities
            f["interpolation_policy"] = "{}".format(self.interpolation_policy)
            if self.save_intermediate_densities_:
                f["intermediate_densities"] = np.zeros((self.system_eval_count,) + self.initial_densities.shape,
                                                       dtype=np.complex128)
            f["method"] = self.method
            f["program_type"] = self.program_type.value
            f["system_eval_count"] = self.system_eval_count
        self._write("w", writer, "initial save")


    def save_all_intermediate_densities(self, iteration, all_densities):
        """one write of the whole [N][D][n][n] trajectory (qoc/models/lindbladmodels.py:92-99 writes it step by step)."""
        if self.save_file_path is None:
            return

        def writer(f):
            f["intermediate_densities"][...] = all_densities.astype(np.complex128)
        self._write("a", writer, "intermediate densities")


class EvolveLindbladResult(object):
    def __init__(self, error=None, final_densities=None):
        self.error = error
        self.final_densities = final_densities


class GrapeLindbladDiscreteState(GrapeState):
    method = "grape_lindblad_discrete"

    def __init__(self, complex_controls, control_count, control_eval_count, cost_eval_step, costs,
                 evolution_time, hamiltonian, impose_control_conditions, initial_controls, initial_densities,
                 interpolation_policy, iteration_count, lindblad_data, log_iteration_step, max_control_norms,
                 min_error, optimizer, save_file_path, save_intermediate_densities_, save_iteration_step,
                 system_eval_count):
        super().__init__(complex_controls, control_count, control_eval_count, cost_eval_step, costs,
                         evolution_time, hamiltonian, impose_control_conditions, initial_controls,
                         interpolation_policy, iteration_count, log_iteration_step, max_control_norms,
                         min_error, optimizer, save_file_path, save_iteration_step, system_eval_count)
        self.hilbert_size = initial_densities[0].shape[0]
        self.initial_densities = initial_densities
        self.lindblad_data = lindblad_data
        self.save_intermediate_densities_ = self.should_save and save_intermediate_densities_

    def log_and_save(self, controls, error, final_densities, grads, iteration):
        self._log_and_save(controls, error, final_densities, "final_densities", grads, iteration)

    def log_and_save_initial(self):
        if self.should_save:
            print("QOC is saving this optimization run to {}.".format(self.save_file_path))

            def writer(f):
                save_count = self._initial_header(f, "final_densities", (len(self.initial_densities),
                                                                        self.hilbert_size, self.hilbert_size))
                f["initial_densities"] = self.initial_densities
                if self.save_intermediate_densities_:
                    f["intermediate_densities"] = np.zeros((save_count, self.system_eval_count)
                                                           + self.initial_densities.shape, dtype=np.complex128)
            self._write("w", writer, "initial save")
        self._print_header()


    def save_all_intermediate_densities(self, iteration, all_densities):
        """row `iteration // save_iteration_step` of the [save_count][N][D][n][n] dataset
        (qoc/models/lindbladmodels.py:315-334)."""
        if iteration > self.final_iteration or not self.should_save:
            return
        if self._due(iteration, self.save_iteration_step):
            row = iteration // self.save_iteration_step

            def writer(f):
                f["intermediate_densities"][row] = all_densities.astype(np.complex128)
            self._write("a", writer, "intermediate densities of iteration {}".format(iteration))


class GrapeLindbladResult(object):
    def __init__(self, best_controls=None, best_error=np.finfo(np.float64).max, best_final_densities=None,
                 best_iteration=None):
        self.best_controls = best_controls
        self.best_error = best_error
        self.best_final_densities = best_final_densities
        self.best_iteration = best_iteration
