"""Input pipeline (reference inference/utils.py:12-16).  tests/golden/preprocess_fractal.npz holds the
output of the REFERENCE's own NPZParser.preprocess on frames of its own fixture (inference/samples/fractal_sample.npz).
CPU: the oracle restatement reproduces it (fp32 rounding, 5e-7).  GPU: the CUDA kernel matches the reference output within
1e-6 absolute on [0,1] pixels, for the uint8 episode layout and for the reference's float [T,C,H,W] argument."""
import os

import numpy as np
import pytest
import torch

from helpers import ROOT

GOLD = os.path.join(ROOT, "tests", "golden", "preprocess_fractal.npz")


def test_oracle_reproduces_reference_preprocess():
    from oracle.preprocess_ref import aa_weights, preprocess_ref
    z = np.load(GOLD)
    assert np.abs(preprocess_ref(z["frames"], (64, 64)) - z["out64"]).max() < 5e-7
    assert np.abs(preprocess_ref(z["frames"][:1], (256, 256)) - z["out256"]).max() < 5e-7
    for n_in, n_out in ((320, 64), (256, 256), (64, 256), (7, 3)):          # weights: normalised, inside the input
        for lo, w in aa_weights(n_in, n_out):
            assert lo >= 0 and lo + len(w) <= n_in and abs(float(w.sum()) - 1.0) < 1e-6
    ident = preprocess_ref(z["frames"][:1, :64, :64], (64, 64))            # same size in and out: only the / 255
    assert np.array_equal(ident, (z["frames"][:1, :64, :64].astype(np.float32) / np.float32(255)).transpose(0, 3, 1, 2))


@pytest.mark.gpu
def test_cuda_preprocess_matches_reference_vectors(cuda):
    from ivideogpt_b200.data_io import NPZParser
    from oracle.preprocess_ref import preprocess_ref
    z = np.load(GOLD)
    frames = torch.from_numpy(z["frames"]).to(cuda)
    got64 = NPZParser(16, 64).preprocess(frames).cpu().numpy()
    assert got64.shape == z["out64"].shape and np.abs(got64 - z["out64"]).max() < 1e-6
    got256 = NPZParser(16, 256).preprocess(frames[:1]).cpu().numpy()
    assert np.abs(got256 - z["out256"]).max() < 1e-6
    # the reference's own argument: float [T,C,H,W] in 0..255
    as_ref = frames.float().permute(0, 3, 1, 2)
    again = NPZParser(16, 64).preprocess(as_ref).cpu().numpy()
    assert np.abs(again - z["out64"]).max() < 1e-6
    # up-scaling and odd sizes against the oracle
    small = z["frames"][:2, :37, :53]
    want = preprocess_ref(small, (64, 64))
    got = NPZParser(16, 64).preprocess(torch.from_numpy(np.ascontiguousarray(small)).to(cuda)).cpu().numpy()
    assert np.abs(got - want).max() < 1e-6


@pytest.mark.gpu
def test_parse_episode_file(cuda, tmp_path):
    from ivideogpt_b200.data_io import NPZParser
    z = np.load(GOLD)
    path = os.path.join(tmp_path, "ep.npz")
    np.savez(path, image=np.repeat(z["frames"], 8, axis=0).astype(np.int64), action=np.zeros((24, 4), np.float32))
    np.random.seed(0)
    frames, act = NPZParser(16, 64).parse(path, "some_dataset", load_action=True)
    assert frames.shape == (16, 3, 64, 64) and frames.is_cuda and act.shape == (16, 4)
    assert float(frames.min()) >= 0.0 and float(frames.max()) <= 1.0
    with pytest.raises(RuntimeError):
        NPZParser(16, 64).preprocess(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))     # CPU tensor: loud failure
