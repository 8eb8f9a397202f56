import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def small():
    return dict(np.load(os.path.join(GOLDEN, "small.npz")))


@pytest.fixture(scope="session")
def golden_c1():
    return dict(np.load(os.path.join(GOLDEN, "c1.npz")))


@pytest.fixture(scope="session")
def golden_pose():
    return dict(np.load(os.path.join(GOLDEN, "pose.npz")))


@pytest.fixture(scope="session")
def golden_mid():
    return dict(np.load(os.path.join(GOLDEN, "mid.npz")))


# Long camera chains with the reference's inner solves converged (tests/golden/make_golden_tight.py): name ->
# (problem, PCG iteration cap of the converged engine solve)
CHAIN_PROBLEMS = {
    "chain": ((300, 6000, 30000, 33), 20000),        # 300 ring cameras, windows of 5 neighbours per point
    "chain1k": ((1000, 20000, 100000, 34), 100000),  # 1 000 cameras: inner solves of up to 28 k LSMR / 23 k PCG iterations
    "c4s": ("C4", 400000),                           # BASELINE configs[3] at 5 % of its points, all 1 778 cameras kept
}


def chain_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + "_tight.npz")))


@pytest.fixture(scope="session")
def golden_chain_tight():
    return chain_golden("chain")


def chain_problem(name="chain"):
    from meatmodeler_b200 import synth
    spec = CHAIN_PROBLEMS[name][0]
    if isinstance(spec, str):
        return synth.make_config(spec, hard=True, scale=0.05)
    nc, npts, nobs, seed = spec
    return synth.make_problem(nc, npts, nobs, seed=seed, hard=True)


def problem_x0(prob):
    """Parameter vector the reference packs (bundleAdjuster.py:172-176) for a synth.Problem."""
    from meatmodeler_b200 import bundleAdjuster as mm
    ext, K, pts, uv, fi, pi = prob.args()
    return np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
