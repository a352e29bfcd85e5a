"""Import the UNMODIFIED reference modules from /root/reference (build container only).

Used by ``tests/golden/make_golden.py`` to generate golden vectors and by a few
``not gpu`` tests that are skipped when /root/reference is absent (the GPU box).
Three modules the reference imports are missing from this image and are stubbed;
none of them touches hot-path arithmetic:

* ``perlin_noise.PerlinNoise``  - only used by Environment.set_dynamics
  (environment.py:62-64, 85); the stub returns a smooth deterministic field so
  ``Environment()`` constructs.  Tests overwrite ``dynamics_speed/angle`` anyway.
* ``pyglet``                    - only needed to import graphics.PathToDraw.
* ``matplotlib(.pyplot)``       - imported but unused (robot.py:12).
"""
import math
import os
import sys
import types

REFERENCE_DIR = os.environ.get("RTD3_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "environment.py"))


class _StubPerlin:
    def __init__(self, octaves=1, seed=0):
        self.octaves = octaves
        self.seed = seed

    def __call__(self, p):
        x, y = p
        k = float(self.octaves)
        return 0.5 * math.sin(k * x * 1.7 + 0.3) * math.cos(k * y * 1.3 + 0.1)


def load_reference():
    """Returns (environment_module, robot_module, constants_module)."""
    if not reference_available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_DIR)
    for name in ("pyglet", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "perlin_noise" not in sys.modules:
        m = types.ModuleType("perlin_noise")
        m.PerlinNoise = _StubPerlin
        sys.modules["perlin_noise"] = m
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import constants  # noqa
    import configuration  # noqa
    import environment  # noqa
    import robot  # noqa
    return environment, robot, constants
