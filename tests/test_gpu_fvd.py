"""FVD parity of the product path against the reference path on identical seeds and pokes (BASELINE configs[4]) at a
size the GPU suite finishes in seconds; the full 1000-poke x 5-sample sweep is `python tests/fvd_parity.py`
(result committed as profiles/r01_fvd_parity.json)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c0,spatial", [(32, 64), (64, 64)])
def test_fvd_parity_small(c0, spatial):
    from fvd_parity import run_sweep
    r = run_sweep(n_pokes=24, n_samples=2, frames=10, spatial=spatial, c0=c0, hd=128, batch_pokes=12,
                  oracle_device="cuda", num_steps=[2, 1, 1] + [1] * 12, verbose=False)
    print(r)
    assert r["videos_per_set"] == 48
    assert r["max_abs_frames"] < 1e-3                       # north_star: per-frame max-abs, fp32 mode
    assert abs(r["fvd_ours_vs_ref"]) <= 1.0                 # +-1.0 FVD, identical seeds and pokes
    assert abs(r["delta_fvd"]) <= 1.0
