"""CPU: the numpy oracle against the golden vectors recorded from the reference (tests/golden/make_golden.py)."""
import numpy as np

from conftest import layers_from_golden
from oracle import siren_oracle as so

TOL_F = 1e-4 * np.sqrt(12.0)  # north_star: |df| <= 1e-4 of the bbox diagonal of [-1,1]^3
TOL_ANGLE = 0.1               # degrees


def test_value_and_gradient_match_reference(golden):
    L = layers_from_golden(golden, "sdf")
    y, J, _ = so.siren_jet(L, golden["x"].astype(np.float64), order=1)
    assert np.abs(y - golden["sdf/y"]).max() < 2e-6
    assert np.abs(y - golden["sdf/y"]).max() < TOL_F
    assert so.angle_deg(J[:, 0], golden["sdf/grad"]).max() < TOL_ANGLE
    assert np.abs(J[:, 0] - golden["sdf/grad"]).max() < 5e-5


def test_float32_jet_is_within_tolerance(golden):
    L = layers_from_golden(golden, "sdf")
    y, J, _ = so.siren_jet(L, golden["x"], order=1, dtype=np.float32)
    assert np.abs(y - golden["sdf/y"]).max() < 1e-5
    assert so.angle_deg(J[:, 0], golden["sdf/grad"]).max() < TOL_ANGLE


def test_quaternion_jacobian(golden):
    L = layers_from_golden(golden, "quat")
    y, J, _ = so.siren_jet(L, golden["x"].astype(np.float64), order=1)
    assert y.shape == (1024, 4) and J.shape == (1024, 4, 3)
    assert np.abs(y - golden["quat/y"]).max() < 2e-6
    assert np.abs(J - golden["quat/jac"]).max() < 5e-5


def test_hessian_matches_hessian3d_and_batched(golden):
    L = layers_from_golden(golden, "sdf")
    _, _, H = so.siren_jet(L, golden["sdf/hess_x"].astype(np.float64), order=2)
    ref = golden["sdf/hess3d"]
    rel = np.linalg.norm((H - ref).reshape(len(H), -1), axis=1) / np.linalg.norm(H.reshape(len(H), -1), axis=1)
    assert rel.max() < 1e-3
    assert np.allclose(golden["sdf/hess_batched"][0], ref)
    assert np.allclose(H, np.swapaxes(H, -1, -2))


def test_principal_curvatures(golden):
    L = layers_from_golden(golden, "sdf")
    x = golden["sdf/hess_x"][:64].astype(np.float64)
    _, J, H = so.siren_jet(L, x, order=2)
    k, d = so.principal_curvatures(J[:, 0], H[:, 0])
    kr, dr = golden["sdf/curv_kappa"], golden["sdf/curv_dirs"]
    assert np.abs(k - kr).max() < 2e-2 * max(1.0, np.abs(kr).max())
    cosang = np.abs((d * dr).sum(-1))
    sep = np.abs(kr[:, 1] - kr[:, 0]) > 1e-2 * np.abs(kr).max()  # away from umbilic points
    assert (cosang[sep] > np.cos(np.radians(0.5))).all()


def test_level_mlp(golden):
    L = [(golden[f"level/network.{i}.weight"], golden[f"level/network.{i}.bias"]) for i in (0, 2, 4)]
    lo = so.relu_mlp(L, golden["level/x"])
    assert np.abs(lo - golden["level/logits"]).max() < 1e-5
    lab, conf = so.level_with_confidence(lo)
    assert np.array_equal(lab, golden["level/logits"].argmax(-1))
    assert ((conf > 0) & (conf <= 1)).all()


def test_grid_generator_quirk(golden):
    g = so.gen_voxel_queries(8)
    assert np.array_equal(g, golden["grid8"])
    # the shear: y/x indices are fractional (sdf_meshing.py:57-58 true division on a LongTensor)
    assert not np.allclose(g[:, 1], np.round((g[:, 1] + 1) / (2 / 7)) * (2 / 7) - 1)
