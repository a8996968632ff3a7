"""oracle/eigen_standin on its own: the stand-in for the dense-Eigen subset that lets the reference's mpc_osqp.cc
compile unmodified (tests/test_reference_source.py) is exercised by a small C++ probe and compared with NumPy /
SciPy - products, block l-values, comma initialisers, Map views, the 3x3 inverse, replicate / asDiagonal / clipping,
AngleAxis / Quaternion composition and rotation, the Pade matrix exponential in all its degree branches, and the
CSC arrays of sparseView() / triangularView<Upper>().  CPU only, needs g++ (no reference files)."""
import json
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse
from scipy.spatial.transform import Rotation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("standin") / "probe")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "oracle", "eigen_standin"),
                           os.path.join(ROOT, "tests", "standin", "standin_probe.cc"), "-o", exe])
    d = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    return {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in d.items()}


def test_dense_algebra_blocks_and_views(probe):
    A, B = probe["A"], probe["B"]
    assert np.allclose(probe["AB"], A @ B, rtol=0, atol=1e-15)
    assert np.allclose(probe["ABt_scaled"], 2 * (A @ B).T - (A @ B).T * 0.25, rtol=0, atol=1e-15)
    C = np.zeros((6, 6))
    C[1:3, 2:5] = A[3:5, 1:4]
    C[4:6, 0:2] += A[0:2, 0:2].T
    C[0:2, 0:2] = [[1, 2.5], [-3, 4]]
    assert np.array_equal(probe["C"], C)
    assert np.array_equal(probe["F"], [[0, 0, 0], [0.3, -1.2, 2.0]])
    raw = np.array([2, 0.1, 0.3, -0.2, 1.5, 0.4, 0.05, -0.6, 3.0])
    M3 = raw.reshape(3, 3).T                      # Matrix3d(const double*) reads column-major (the inertia quirk)
    assert np.array_equal(probe["M3"], M3)
    assert np.allclose(probe["M3inv"], np.linalg.inv(M3), rtol=1e-14, atol=1e-15)
    assert np.array_equal(probe["Mt"], [[1, 2], [3, 4], [5, 6.0]])   # column-major 2x3 view, transposed
    assert np.array_equal(probe["ov"].ravel(), [-1, 2, -3, 4])
    assert np.array_equal(probe["D"], np.diag([1, 2, 3, 1, 2, 3.0]))
    assert np.array_equal(probe["clip"].ravel(), [1.5, 2, 2.5])
    assert np.array_equal(probe["I23"], np.eye(2, 3))


def test_geometry_matches_scipy_rotation(probe):
    r, p, y = 0.31, -0.22, 1.05
    want = Rotation.from_euler("ZYX", [y, p, r]).as_matrix()          # Rz(yaw) Ry(pitch) Rx(roll), mpc_osqp.cc:283-291
    assert np.allclose(probe["Rzyx"], want, rtol=0, atol=2e-16 * 8)
    rot = Rotation.from_euler("XYZ", [r, p, y]).as_matrix()           # Rx Ry Rz, the foot rotation (mpc_osqp.cc:606-617)
    assert np.allclose(probe["rot_xyz_v"].ravel(), rot @ np.array([0.3, -1.2, 2.0]), rtol=0, atol=1e-15)


@pytest.mark.parametrize("k", ["E1", "E2", "E3"])
def test_pade_matrix_exponential_matches_scipy(probe, k):
    want = scipy.linalg.expm(probe[k])
    assert np.allclose(probe["exp" + k], want, rtol=1e-12, atol=1e-14 * np.abs(want).max())


def test_sparse_views_match_scipy_csc(probe):
    S = probe["S"]
    for key, M in (("S_csc", S), ("S_upper_csc", np.triu(S))):
        want = scipy.sparse.csc_matrix(M)
        want.sort_indices()
        got = probe[key]
        assert list(got["indptr"]) == list(want.indptr) and list(got["indices"]) == list(want.indices)
        assert np.array_equal(got["data"], want.data)
    assert 1e-300 in probe["S_csc"]["data"]       # sparseView() drops exact zeros only
