"""Minimal stand-in for the `tasklogger` package (not installed, no network).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference at /root/reference be
imported in the build container to generate golden fixtures.  Provides just the
calls the reference makes (reference graphtools/base.py:19-22, graphs.py:20,47).
"""
import contextlib


class _Logger:
    def __init__(self, name):
        self.name = name
        self.level = 0

    def set_level(self, level=1):
        self.level = level
        return self

    @contextlib.contextmanager
    def log_task(self, name):
        yield

    # backwards-compatible spellings used across tasklogger versions
    task = log_task

    def log_debug(self, msg):
        pass

    def log_info(self, msg):
        pass

    def log_warning(self, msg):
        pass

    def log_error(self, msg):
        pass

    debug = log_debug
    info = log_info
    warning = log_warning


_LOGGERS = {}


def get_tasklogger(name="TaskLogger"):
    if name not in _LOGGERS:
        _LOGGERS[name] = _Logger(name)
    return _LOGGERS[name]
