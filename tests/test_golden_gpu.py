"""GPU: replay the committed golden fixtures (tests/golden/engine_*.npz, made by tests/make_golden.py from the CPU
oracle) through the CUDA engine's C ABI.  Tolerance 1e-9 absolute on O(1) quantities after up to 800 steps."""
import os

import numpy as np
import pytest

import make_golden
from helpers import A, engine_factory

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["tully1_fssh", "spinboson_fssh", "spinboson_ehrenfest", "rpmd_harmonic32", "rpsh_morse3_16",
                                  "iesh_m30", "ehrenfest_na_m30", "rpsh_morse3_10", "langevin_harmonic8",
                                  "tully1_fssh_terminating", "iesh_m30_terminating"])
def test_engine_reproduces_golden(name):
    cs, T, _ = make_golden.cases()
    g = np.load(os.path.join(GOLDEN, f"engine_{name}.npz"))
    case = cs[name]
    case["r"], case["v"] = g["r0"], g["v0"]
    res = make_golden.run_case(engine_factory(), case, T, g["draws"], g["sdraw"], g["noise"] if "noise" in g.files else None)
    for k, val in res.items():
        scale = max(1.0, float(np.max(np.abs(g[k]))))
        assert np.max(np.abs(val - g[k])) < 1e-9 * scale, (name, k)
