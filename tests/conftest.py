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


@pytest.fixture(scope="session")
def golden_chain_tight():
    """300-camera chain, the unmodified reference with converged inner solves (tests/golden/make_golden_tight.py)."""
    return dict(np.load(os.path.join(GOLDEN, "chain_tight.npz")))


def chain_problem():
    """The problem of chain_tight.npz: 300 ring cameras, windows of 5 neighbours per point (a long camera chain)."""
    from meatmodeler_b200 import synth
    return synth.make_problem(300, 6000, 30000, seed=33, hard=True)


def problem_x0(prob):
    """Parameter vector the reference packs (bundleAdjuster.py:172-176) for a synth.Problem."""
    from meatmodeler_b200 import bundleAdjuster as mm
    ext, K, pts, uv, fi, pi = prob.args()
    return np.hstack((mm.frameParameters(ext), np.asarray(pts).reshape(-1)))
