import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def ensure_built():
    """Builds librt_b200.so / the oracle if they are missing (nvcc cross-compiles without a GPU)."""
    need = [os.path.join(ROOT, "ray_tracying_b200", "librt_b200.so"), os.path.join(ROOT, "oracle", "librt_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


ensure_built()


def golden_scene(name: str) -> dict:
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def golden_data(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def scene_file(tmp_path, scene: dict, name="scene.json") -> str:
    p = os.path.join(str(tmp_path), name)
    with open(p, "w") as f:
        json.dump(scene, f, separators=(",", ":"))
    return p


def with_resolution(scene: dict, w: int, h: int) -> dict:
    out = dict(scene)
    out["render"] = {"resolution_x": w, "resolution_y": h}
    return out


def lsb_agreement(a: np.ndarray, b: np.ndarray, tol: int = 1) -> float:
    """Fraction of PIXELS whose every channel differs by <= tol 8-bit steps."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=-1)
    return float((d <= tol).mean())


def linear_mismatch(a: np.ndarray, b: np.ndarray, spp: int = 1) -> float:
    """Largest |a - b| in units of the allowed bound for float images that took the same sample paths.
    The paths' colours agree to a few ulp (powf in the specular term is the only non-IEEE operation);
    the reference then adds its spp samples in float (up to spp * 2^-24 relative), the CUDA path adds them
    exactly (64-bit fixed point) -- hence a RELATIVE bound that grows with spp, plus 2e-7 absolute for
    channels near zero. <= 1 passes."""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    bound = (2e-6 + 1.2e-7 * spp) * np.maximum(np.abs(a), np.abs(b)) + 2e-7
    return float((np.abs(a - b) / bound).max()) if a.size else 0.0


def rmse(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


@pytest.fixture(scope="session")
def rt():
    import ray_tracying_b200
    return ray_tracying_b200


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    return oracle
