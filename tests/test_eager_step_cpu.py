"""CPU: oracle/eager_step.py (the torch-eager comparator `bench.py --impl eager` times on the GPU) against
oracle/step_port.py (the checker pinned to the reference by tests/golden/*) from identical weights, inputs and
RNG seeds.  Both run on the CPU here, so they see the same conv library; they differ in where the loss math
runs (torch ops vs numpy + the C arithmetic contract): losses to 1e-4 relative, same bank rows."""
import numpy as np
import torch

import bench
from oracle import eager_step, model_port, step_port


def _clone(state):
    return {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in state.items()}


def test_eager_step_tracks_the_pinned_port():
    C, crop = 21, 65
    cfg = bench.make_cfg("tiny")
    cfg["net"]["num_classes"] = C
    base = model_port.init_state("resnet50", C, False, seed=4, peak=8.0)
    a = step_port.ReferenceStep(_clone(base), _clone(base), cfg, "resnet50")
    b = eager_step.EagerStep(_clone(base), _clone(base), cfg, "resnet50")
    for s in (a, b):
        s.student.dropout_p = s.teacher.dropout_p = 0.0
    for rnd in range(2):
        image_l, label_l, image_u = bench.synth_batch(50 + rnd, 2, 2, crop, C)
        out = []
        for s in (a, b):
            np.random.seed(11 + rnd)
            torch.manual_seed(12 + rnd)
            out.append(s.step(image_l, label_l, image_u, 40, 4000 + rnd, 100))
        for x, y in zip(*out):
            assert abs(x - y) <= 1e-4 * max(1.0, abs(x)), out
    assert out[0][2] > 0                                                  # contrastive branch ran
    for c in range(C):
        ra, rb = a.memobank[c][0], b.memobank[c][0].numpy()
        assert ra.shape == rb.shape and np.allclose(ra, rb, atol=1e-4), c   # same rows; step-2 weights differ by ulps
    k = "decoder.classifier.8.weight"
    assert (a.teacher.s[k] - b.teacher.s[k]).abs().max() <= 1e-6
    assert (a.student.s[k] - b.student.s[k]).abs().max() <= 1e-5


def test_eager_label_onehot_matches_golden_quirk():
    from oracle import port                     # port.label_onehot is pinned by golden/contra_c21_driver_onehot.npz
    g = np.random.default_rng(0)
    lab = g.integers(0, 21, (3, 9, 11))
    lab[g.random(lab.shape) < 0.1] = 255
    got = eager_step.label_onehot(torch.from_numpy(lab), 21).numpy()
    assert np.array_equal(got, port.label_onehot(lab, 21))
