"""Host-side logic of the harnesses (no GPU): label rows round-trip through the reference's csv format, and the
visu.py update rule of harness/optimize.py drives the oracle loss down."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harness import make_dataset, optimize            # noqa: E402
from oracle import sq_oracle as O                      # noqa: E402
from oracle import ref_import                          # noqa: E402


def test_label_rows_round_trip(tmp_path):
    params = O.random_params(16, 4).numpy()
    rows = make_dataset.label_rows(params, [f"synth/{i:06d}.bmp" for i in range(16)])
    back = np.stack(make_dataset.parse_rows(rows))
    np.testing.assert_allclose(back, params, rtol=2e-7, atol=1e-7)
    if ref_import.available():                          # the reference's own parser, where the tree is mounted
        ref_import.load()                               # puts the reference's torch/ directory on sys.path (with stubs)
        import helpers                                  # noqa: E402  (reference module, bare name)
        f = tmp_path / "labels.csv"
        f.write_text("\n".join(rows) + "\n")
        ref = np.stack(helpers.parse_csv(str(f)))
        np.testing.assert_array_equal(ref, back)


def test_descent_reduces_oracle_loss():
    true = O.random_params(3, 9).double()
    pred = O.perturbed_params(O.random_params(3, 9), 2, sigma=0.05).double()
    crit = O.ExplicitLoss(16, "cpu")
    rec = []
    optimize.descend(crit, true, pred, 15, record=rec)
    assert rec[-1].item() < rec[0].item()
    np.testing.assert_allclose(pred[:, 8:].norm(dim=1).numpy(), 1.0, atol=1e-12)
