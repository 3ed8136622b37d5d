"""IterLog: the iteration log of the reference's algorithms (src/utility/logging.jl:12-142) with its exact line formats,
and the verbosity levels of src/utility/defaults.jl:12-16.  Host-side; the objective sum(expectation_value(psi, H, envs)) is
evaluated only when the level that prints it is enabled, as in the reference (`@infov N` is lazy)."""
from __future__ import annotations
import sys
import time

VERBOSE_NONE, VERBOSE_WARN, VERBOSE_CONV, VERBOSE_ITER, VERBOSE_ALL = 0, 1, 2, 3, 4


def format_time(t):
    """logging.jl:105-109"""
    if t < 60:
        return f"{t:0.2f} sec"
    if t < 2600:
        return f"{t / 60:0.2f} min"
    return f"{t / 3600:0.2f} hr"


class IterLog:
    def __init__(self, name="", verbosity=VERBOSE_NONE, stream=None):
        self.name, self.verbosity = name, int(verbosity)
        self.stream = stream
        self.iter, self.error, self.objective = 0, float("nan"), None
        self.t_init = self.t_prev = self.t_last = time.time()
        self.state = "INIT"

    def enabled(self, level):
        """LoggingExtras.withlevel(; alg.verbosity): `@infov N` / `@warnv N` print when verbosity >= N"""
        return self.verbosity >= level

    def _set(self, state, it, error, objective):
        self.iter, self.error = int(it), float(error)
        if objective is not None:
            self.objective = float(objective.real if hasattr(objective, "real") else objective)
        if state == "INIT":
            self.t_init = self.t_prev = self.t_last = time.time()
        else:
            self.t_prev, self.t_last = self.t_last, time.time()
        self.state = state
        return self

    def _emit(self, prefix):
        print(f"{prefix}{self}", file=self.stream or sys.stderr)

    # logging.jl:33-103 + the `@infov` / `@warnv` call sites of dmrg.jl:27-49, vumps.jl:35-72
    def init(self, error, objective=None):
        self._set("INIT", 0, error, objective() if callable(objective) and self.enabled(VERBOSE_CONV) else None)
        if self.enabled(VERBOSE_CONV):
            self._emit("[ Info: ")

    def iteration(self, it, error, objective=None):
        if self.enabled(VERBOSE_ITER):
            self._set("ITER", it, error, objective() if callable(objective) else None)
            self._emit("[ Info: ")
        else:
            self._set("ITER", it, error, None)

    def finish(self, it, error, objective=None):
        if self.enabled(VERBOSE_CONV):
            self._set("CONV", it, error, objective() if callable(objective) else None)
            self._emit("[ Info: ")

    def cancel(self, it, error, objective=None):
        if self.enabled(VERBOSE_WARN):
            self._set("CANCEL", it, error, objective() if callable(objective) else None)
            self._emit("┌ Warning: ")

    def __str__(self):
        """Base.show(io, log) (logging.jl:112-142)"""
        obj = "" if self.objective is None else f"obj = {self.objective:+0.12e}\t"
        if self.state == "INIT":
            return f"{self.name} init:\t{obj}err = {self.error:0.4e}"
        if self.state == "CONV":
            return f"{self.name} conv {self.iter}:\t{obj}err = {self.error:0.10e}\ttime = {format_time(self.t_last - self.t_init)}"
        if self.state == "ITER":
            return f"{self.name} {self.iter:3d}:\t{obj}err = {self.error:0.10e}\ttime = {format_time(self.t_last - self.t_prev)}"
        return f"{self.name} cancel {self.iter}:\t{obj}err = {self.error:0.10e}\ttime = {format_time(self.t_last - self.t_init)}"
