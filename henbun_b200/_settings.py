"""Configuration: the three numerics keys and the dtype key of the reference's ``henbunrc``
(Henbun/henbunrc:6-14, Henbun/_settings.py), with ``get_settings()`` / ``temp_settings()``.

Like the reference, ``jitter_level`` and ``clip_*`` are read when an objective is evaluated
(``with hb.settings.temp_settings(cfg): model.ELBO().compile()`` -- Expert_GPR.ipynb:224-226); the
compiled Optimizer captures them so later steps use the compile-time values.
"""
from __future__ import annotations

import configparser
import copy
import os
from collections import OrderedDict


class MutableNamedTuple(OrderedDict):
    """Nested attribute-style settings container (Henbun/_settings.py:66-87)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _parse(s):
    if not isinstance(s, str):
        raise ValueError
    if s in ("true", "True"):
        return True
    if s in ("false", "False"):
        return False
    if s in ("float64", "float32", "float16", "int64", "int32", "int16"):
        return s
    try:
        return int(s)
    except ValueError:
        pass
    try:
        return float(s)
    except ValueError:
        return s


_DEFAULTS = {
    "verbosity": {"tf_compile_verb": False, "hmc_verb": True, "optimisation_verb": False},
    "dtypes": {"float_type": "float32", "int_type": "int32"},
    "numerics": {"jitter_level": 1e-5, "clip_by_value": False, "clip_value_min": -50.0, "clip_value_max": 50.0},
    "profiling": {"dump_timeline": False, "dump_tensorboard": False},
}


def _load():
    cfg = MutableNamedTuple()
    for sec, kv in _DEFAULTS.items():
        cfg[sec] = MutableNamedTuple(kv)
    c = configparser.ConfigParser()
    # same search order as the reference (_settings.py:126-144): cwd, home, package directory
    for loc in (os.curdir, os.path.expanduser("~"), os.path.dirname(os.path.realpath(__file__))):
        for name in ("henbunrc", ".henbunrc"):
            if c.read(os.path.join(os.path.abspath(loc), name)):
                for sec in c.sections():
                    tgt = cfg.setdefault(sec, MutableNamedTuple())
                    for k, v in c[sec].items():
                        tgt[k] = _parse(v)
                return cfg
    return cfg


class _TempSettings:
    def __init__(self, manager, tmp):
        self._m, self._tmp = manager, tmp

    def __enter__(self):
        self._m.push(self._tmp)

    def __exit__(self, *exc):
        self._m.pop()


class SettingsManager:
    def __init__(self, cur):
        object.__setattr__(self, "_cur", cur)
        object.__setattr__(self, "_stack", [])

    def __getattr__(self, name):
        try:
            return self._cur[name]
        except KeyError:
            raise AttributeError("Unknown setting.")

    def push(self, s):
        self._stack.append(self._cur)
        object.__setattr__(self, "_cur", s)

    def pop(self):
        rem = self._cur
        object.__setattr__(self, "_cur", self._stack.pop())
        return rem

    def temp_settings(self, tmp):
        return _TempSettings(self, tmp)

    def get_settings(self):
        return copy.deepcopy(self._cur)


settings = SettingsManager(_load())
