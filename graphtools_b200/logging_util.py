"""Tiny task logger with the scope names the reference emits through `tasklogger`
(reference graphtools/graphs.py:873-885, :1197-1222, :1868-1915; base.py:242).  Verbose output is
plain prints of wall-clock per task so PHATE-style callers see the familiar lines."""
import contextlib
import time


class _TaskLogger:
    def __init__(self):
        self.level = 0
        self.indent = 0

    def set_level(self, level=1):
        if level is True:
            level = 1
        elif level is False or level is None:
            level = 0
        self.level = int(level)
        return self

    def _emit(self, msg, min_level):
        if self.level >= min_level:
            print("  " * self.indent + msg)

    @contextlib.contextmanager
    def log_task(self, name):
        self._emit("Calculating {}...".format(name), 1)
        self.indent += 1
        t0 = time.perf_counter()
        try:
            yield
        finally:
            self.indent -= 1
            self._emit("Calculated {} in {:.2f} seconds.".format(name, time.perf_counter() - t0), 1)

    def log_info(self, msg):
        self._emit(msg, 1)

    def log_debug(self, msg):
        self._emit(msg, 2)

    def log_warning(self, msg):
        self._emit(msg, 0)


logger = _TaskLogger()
