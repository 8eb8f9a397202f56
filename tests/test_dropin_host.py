"""Host side of the drop-in module (meatmodeler_b200/bundleAdjuster.py) against outputs of the unmodified reference
(tests/golden/small.npz): the functions that keep running on the CPU in the reference too (O(Nc) packing / unpacking,
the sparsity pattern kept for callers), and the documented zero-change import route."""
import os
import subprocess
import sys

import numpy as np
import pytest

from meatmodeler_b200 import bundleAdjuster as mm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_parameters_vs_reference_golden(small):
    got = mm.frameParameters(small["ext"])
    np.testing.assert_allclose(got, small["frame_parameters"], rtol=0, atol=1e-15)
    assert np.all(got[:3] == 0)                       # identity rotation -> zero rvec (the reference's NaN -> 0 rule)
    # the parameter vector the reference packs (bundleAdjuster.py:172-176)
    np.testing.assert_array_equal(np.hstack((got, small["pts"].reshape(-1))), small["x0"])
    # (Nc,4,4) input, as processor.py passes after the first adjustment
    ext4 = np.zeros((len(small["ext"]), 4, 4))
    ext4[:, :3] = small["ext"][:, :3]
    ext4[:, 3, 3] = 1
    np.testing.assert_array_equal(mm.frameParameters(ext4), got)


def test_reformat_point_result_vs_reference_golden(small):
    """reformatPointResult on the reference's own solution vector reproduces the reference's outputs (the reference
    converts with cv2.Rodrigues, bundleAdjuster.py:153; the drop-in with its own batched Rodrigues formula)."""
    nc, npts = len(small["ext"]), len(small["pts"])
    pts, ext = mm.reformatPointResult(mm.SolveResult(x=small["ref_x"]), nc, npts)
    assert pts.shape == (npts, 3) and isinstance(ext, list) and len(ext) == nc and ext[0].shape == (4, 4)
    np.testing.assert_allclose(pts, small["adj_points"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(np.array(ext), small["adj_extrinsics"], rtol=0, atol=1e-12)


def test_sparsity_pattern_vs_reference_golden(small):
    nc, npts = len(small["ext"]), len(small["pts"])
    A = mm.pointAdjustmentSparsity(nc, npts, small["fi"], small["pi"]).tocsr()
    A.sort_indices()
    assert A.shape == (2 * len(small["fi"]), 6 * nc + 3 * npts) and A.dtype.kind == "i"
    np.testing.assert_array_equal(A.indices, small["pattern_indices"])
    np.testing.assert_array_equal(A.indptr, small["pattern_indptr"])
    assert A.data.min() == 1 and A.data.max() == 1


def test_solve_result_behaves_like_optimize_result():
    import copy
    import pickle
    r = mm.SolveResult(x=np.arange(3.0), cost=1.5)
    assert r.cost == 1.5 and r["cost"] == 1.5
    assert not hasattr(r, "nope") and getattr(r, "nope", 7) == 7
    with pytest.raises(AttributeError):
        r.nope
    assert copy.deepcopy(r).cost == 1.5 and pickle.loads(pickle.dumps(r)).cost == 1.5
    r.extra = 2
    assert r["extra"] == 2


def test_module_exposes_the_reference_names():
    for name in ("rotate", "project", "pointAdjustmentSparsity", "pointFun", "frameParameters", "reformatPointResult",
                 "adjustPoints", "reformatPoseResult", "poseFun", "adjustPose"):
        assert callable(getattr(mm, name)), name


@pytest.mark.skipif(not os.path.exists("/root/reference/processor.py"), reason="needs the reference checkout (build container)")
def test_documented_zero_change_route_imports_the_unmodified_processor():
    """with meatmodeler_b200/ ahead of the reference on sys.path, ``import processor`` loads the
    reference's processor.py unchanged and its ``import bundleAdjuster`` resolves to the drop-in.  (pyntcloud / lxml,
    which processor.py imports for its PLY export, are not installed in this image: stubbed.)"""
    code = (
        "import sys, types\n"
        "for m in ('pyntcloud', 'lxml'):\n"
        "    sys.modules[m] = types.ModuleType(m)\n"
        "sys.modules['pyntcloud'].PyntCloud = object\n"
        "sys.path.insert(0, '/root/reference')\n"
        f"sys.path.insert(0, {os.path.join(ROOT, 'meatmodeler_b200')!r})\n"
        "import processor\n"
        "assert processor.__file__.startswith('/root/reference'), processor.__file__\n"
        f"assert processor.bundleAdjuster.__file__.startswith({os.path.join(ROOT, 'meatmodeler_b200')!r}), processor.bundleAdjuster.__file__\n"
        "assert processor.bundleAdjuster.adjustPoints.__module__ == 'bundleAdjuster'\n"
        "print('ROUTE-OK')\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert out.returncode == 0 and "ROUTE-OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
