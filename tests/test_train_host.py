"""Host logic of the training step (pl-nerf_b200/train.py) that needs no GPU: the pixel draw
(the law of run_plnerf.py:1259-1280), the precrop window and the learning-rate decay (:1307-1315)."""
import numpy as np
import pytest
import torch

import plnerf_b200.train as T


def test_crop_window_matches_reference_linspace():
    for H, W, frac in ((800, 800, 0.5), (378, 504, 0.5), (401, 301, 0.3)):
        r0, c0, rows, cols = T.crop_window(H, W, frac)
        dH, dW = int(H // 2 * frac), int(W // 2 * frac)
        # the reference's coords: linspace(H//2 - dH, H//2 + dH - 1, 2*dH) x linspace(W//2 - dW, W//2 + dW - 1, 2*dW)
        rr = torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH).long()
        cc = torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW).long()
        assert (r0, rows) == (int(rr[0]), rr.numel()) and int(rr[-1]) == r0 + rows - 1
        assert (c0, cols) == (int(cc[0]), cc.numel()) and int(cc[-1]) == c0 + cols - 1


@pytest.mark.parametrize("frac", [None, 0.5])
def test_sample_pixels_distinct_in_window_and_reproducible(frac):
    H, W, N = 60, 80, 512
    gen = torch.Generator(device="cpu")
    gen.manual_seed(3)
    pix = T.sample_pixels(H, W, N, "cpu", gen, frac)
    assert pix.dtype == torch.int64 and pix.shape == (N,)
    assert torch.unique(pix).numel() == N                                   # replace=False
    r, c = pix // W, pix % W
    r0, c0, rows, cols = (0, 0, H, W) if frac is None else T.crop_window(H, W, frac)
    assert int(r.min()) >= r0 and int(r.max()) < r0 + rows and int(c.min()) >= c0 and int(c.max()) < c0 + cols
    gen.manual_seed(3)
    assert torch.equal(pix, T.sample_pixels(H, W, N, "cpu", gen, frac))
    assert not torch.equal(pix, T.sample_pixels(H, W, N, "cpu", gen, frac))    # the stream advances
    with pytest.raises(ValueError):
        T.sample_pixels(4, 4, 17, "cpu", gen)


def test_sample_pixels_is_uniform():
    """Every pixel of the window is equally likely: chi-square of 4000 draws of 8 pixels from a 6 x 8 window."""
    H, W, N, trials = 12, 16, 8, 4000
    gen = torch.Generator(device="cpu")
    gen.manual_seed(0)
    r0, c0, rows, cols = T.crop_window(H, W, 0.5)
    counts = np.zeros((H, W))
    for _ in range(trials):
        p = T.sample_pixels(H, W, N, "cpu", gen, 0.5).numpy()
        np.add.at(counts, (p // W, p % W), 1)
    win = counts[r0:r0 + rows, c0:c0 + cols]
    assert counts.sum() == win.sum() == trials * N
    exp = trials * N / win.size
    chi2 = ((win - exp) ** 2 / exp).sum()
    assert chi2 < 2.0 * win.size, chi2          # dof = 47; 2x dof is > 6 sigma


def test_decayed_lrate_formula():
    # run_plnerf.py:1307-1309 with blender_linear.txt's lrate_decay = 500
    for step in (0, 1, 1000, 250000, 500000):
        want = 5e-4 * (0.1 ** (step / (500 * 1000)))
        assert T.decayed_lrate(5e-4, 500, step) == want
    assert T.decayed_lrate(5e-4, 500, 500000) == pytest.approx(5e-5)


def test_train_step_needs_cuda_models():
    from plnerf_b200.run_nerf_helpers import NeRF
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        T.TrainStep(8, 8, np.eye(3), dict(network_fn=net, network_fine=net, N_samples=8, N_importance=8))
