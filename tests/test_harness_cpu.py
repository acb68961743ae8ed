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


def test_package_input_generators_are_the_reference_distributions():
    """sq_recovery_b200/inputs.py (what bench.py, harness/ and tools/ draw their workloads from) yields the same rows
    as the oracle's restatement of randsq / randquat (visu.py:55-56, quaternion.py:139-145), bit for bit."""
    from sq_recovery_b200 import inputs
    for seed in (0, 7, 1003):
        a, b = inputs.random_params(9, seed), O.random_params(9, seed)
        assert torch.equal(a, b)
        assert torch.equal(inputs.perturbed_params(a, seed + 1), O.perturbed_params(b, seed + 1))
        assert torch.equal(inputs.random_params(4, seed, torch.float64), O.random_params(4, seed, torch.float64))
    if ref_import.available():
        _, rq = ref_import.load()
        np.random.seed(5)
        ref = rq.randquat()
        got = inputs.randquat(np.random.RandomState(5))
        np.testing.assert_array_equal(np.asarray(ref, dtype=np.float64).reshape(-1), got)
    dense = inputs.random_params(64, 3, size_range=inputs.DENSE_SIZE_RANGE)
    assert dense[:, :3].min() >= 0.5 and dense[:, :3].max() <= 1.0


def test_checkpoint_format_is_the_references(tmp_path):
    """harness/train_step.py writes the dict of torch/helpers.py:42-48; where the reference tree is mounted its own
    load_model reads it back."""
    from harness import train_step
    net = torch.nn.Linear(3, 2)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    net(torch.ones(1, 3)).sum().backward(); opt.step()
    hist = {"loss": [0.5, 0.4], "val_loss": [0.6, 0.45], "val_acc": [[0.1], [0.2]]}
    path = str(tmp_path / "model.pt")
    train_step.save_checkpoint(path, 1, net, opt, hist)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"} and ck["epoch"] == 1
    net2 = torch.nn.Linear(3, 2)
    opt2 = torch.optim.Adam(net2.parameters(), lr=1e-4)
    epoch, loss = train_step.load_checkpoint(path, net2, opt2, "cpu")
    assert epoch == 1 and loss == hist and torch.equal(net2.weight, net.weight)
    assert opt2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]
