import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def rtlib():
    """The product library.  Built in-tree if stale; loading must work without a GPU."""
    from cs397raytracingsp22_b200 import _ffi
    return _ffi.load()


@pytest.fixture(scope="session")
def gpu(rtlib):
    """Skip-free guard for GPU tests: they FAIL (not skip) when no device is visible."""
    n = rtlib.rt_device_count()
    assert n > 0, "GPU test selected but no CUDA device is usable (there is no CPU fallback)"
    return n


# small variants of the five BASELINE configurations, sized so the oracle finishes in seconds
SMALL = {
    "c1": dict(width=96, height=96, spp=16),
    "c2": dict(width=96, height=96, spp=16),
    "c3": dict(width=96, height=96, spp=16),
    "c4": dict(width=160, height=90, spp=16, map_size=256),
    "c5": dict(width=160, height=90, spp=16, map_size=128, grid=6),
}


@pytest.fixture(scope="session")
def small_scenes():
    from cs397raytracingsp22_b200 import scenes
    cache = {}

    def get(name, **kw):
        key = (name, tuple(sorted(kw.items())))
        if key not in cache:
            args = dict(SMALL[name])
            args.update(kw)
            cache[key] = scenes.make_scene(name, **args)
        return cache[key]

    return get
