"""Test-only shim that hosts the reference's pure-Python logic without Panda3D / gym.

NOT product code.  Only ``tools/make_golden.py`` imports this, and only inside the build
container where ``/root/reference`` exists; nothing under ``pgdrive_b200/`` may import it.
The recipe follows SURVEY.md Appendix D: no-op stub modules for the un-installed third-party
packages, numpy alias back-fills, and a package object whose ``__path__`` points at the
read-only reference so ``pgdrive/__init__.py`` (which needs gym) is skipped.
"""
import sys
import types

import numpy as np

REF_ROOT = "/root/reference/pgdrive"


class Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return Anything()

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return Anything()

    def __iter__(self):
        return iter(())

    def __or__(self, o):
        return self

    __ror__ = __or__

    def __bool__(self):
        return False

    def getWord(self):
        return 2


class Truthy(Anything):
    def __bool__(self):
        return True


class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Meta(k, (Anything, ), {})

    def __or__(cls, o):
        return cls

    __ror__ = __or__


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        c = _Meta(k, (Anything, ), {})
        setattr(self, k, c)
        return c


_INSTALLED = False


def install():
    global _INSTALLED
    if _INSTALLED:
        return
    for n, t in (("bool", bool), ("float", float), ("int", int)):
        if not hasattr(np, n):
            setattr(np, n, t)
    names = (
        "panda3d panda3d.core panda3d.bullet gym gym.spaces gym.envs gym.envs.registration seaborn pygame "
        "gltf direct direct.showbase direct.showbase.ShowBase direct.gui direct.gui.OnscreenImage "
        "direct.controls direct.controls.InputState simplepbr evdev"
    ).split()
    for name in names:
        sys.modules[name] = _Stub(name)
    pkg = types.ModuleType("pgdrive")
    pkg.__path__ = [REF_ROOT]
    sys.modules["pgdrive"] = pkg
    _INSTALLED = True


class FakeEngine(Truthy):
    """Stands in for BaseEngine.singleton while the reference's map / traffic code runs."""
    def __init__(self, seed, global_config):
        from pgdrive.utils.random_utils import get_np_random
        self.global_random_seed = seed
        self.global_config = global_config
        self.np_random = get_np_random(seed)
        self.worldNP = Truthy()
        self.physics_world = Truthy()
        self.pbr_worldNP = Truthy()
        self.current_map = None
        self.spawn_log = []
        self.policies = {}

    MAX_RAND_INT = 65536

    def generate_seed(self):
        return self.np_random.randint(0, self.MAX_RAND_INT)
