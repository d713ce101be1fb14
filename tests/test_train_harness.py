"""Training harness (ddsp_pytorch_b200/train.py): the reference's on-disk dataset format
(ddsp/data.py:9-33), rank sharding of global batches, and -- on a GPU -- a short run that must reduce the loss."""
import numpy as np
import pytest
import torch


def test_dataset_format_and_collate(tmp_path):
    from ddsp_pytorch_b200 import train
    d = train.write_synthetic_dataset(tmp_path / "train", 10, sample_rate=16000, block_size=160, seconds=0.5, seed=3)
    assert sorted(p.name for p in d.iterdir()) == ["loudness.npy", "mfccs.npy", "pitchs.npy", "signals.npy"]   # data.py:13-16
    ds = train.NpyDataset(d)
    assert len(ds) == 10
    item = ds[4]
    assert item["sig"].shape == (8000,) and item["pitch"].shape == (50, 1) and item["loudness"].shape == (50, 1)
    assert item["mfcc"].shape == (50, 30)                                     # data.py:25 drops the last MFCC frame
    assert torch.equal(item["sig"], torch.from_numpy(np.load(d / "signals.npy")[4]))
    b = ds.batch(np.array([7, 2, 5]))
    assert b["sig"].shape == (3, 8000) and torch.equal(b["pitch"][0, :, 0], torch.from_numpy(np.load(d / "pitchs.npy")[2]))
    m, s = train.loudness_stats(ds, 5)
    from ddsp_pytorch_b200 import core
    l = torch.from_numpy(np.load(d / "loudness.npy"))
    mr, sr = core.mean_std_loudness([{"loudness": l[:5]}, {"loudness": l[5:]}])
    assert abs(m - mr) < 1e-6 and abs(s - sr) < 1e-6


def test_ranks_read_disjoint_slices_of_the_same_global_batches(tmp_path):
    from ddsp_pytorch_b200 import train
    d = train.write_synthetic_dataset(tmp_path / "t", 22, seconds=0.1, seed=1)
    ds = train.NpyDataset(d)
    loaders = [train.ShardedLoader(ds, 8, r, 4, "cpu", seed=5) for r in range(4)]
    assert all(len(l) == 2 for l in loaders)                                  # drop_last: 22 // 8
    for epoch in range(2):
        for step in range(2):
            parts = [l.indices(epoch, step) for l in loaders]
            assert all(len(p) == 2 for p in parts)
            assert len(set(np.concatenate(parts).tolist())) == 8              # disjoint, together one global batch
    assert not np.array_equal(loaders[0].indices(0, 0), loaders[0].indices(1, 0))   # reshuffled every epoch
    batches = list(loaders[1])
    assert len(batches) == 2 and batches[0]["sig"].shape == (2, 1600)
    with pytest.raises(ValueError):
        train.ShardedLoader(ds, 6, 0, 4, "cpu")                               # 6 voices do not split over 4 ranks


@pytest.mark.gpu
def test_short_training_run_reduces_the_loss_and_saves_a_loadable_state(tmp_path):
    from ddsp_pytorch_b200 import train
    data = train.write_synthetic_dataset(tmp_path / "data" / "train", 16, seconds=1.0, seed=0)
    args = train.parser().parse_args(["--data", str(tmp_path / "data"), "--root", str(tmp_path / "runs"), "--name", "t",
                                      "--steps", "40", "--batch", "8", "--scales", "1024", "512", "256", "128",
                                      "--log-every", "5", "--warmup", "2", "--lr", "1e-3"])
    rep = train.run(args)
    assert rep["steps"] == 40 and rep["world_size"] == 1 and rep["ms_per_step"] > 0
    assert rep["last_logged_loss"] < 0.8 * rep["first_logged_loss"], rep
    state = torch.load(tmp_path / "runs" / "t" / "state.pth")
    model = train.build_model("decoder", dict(hidden_size=512, n_harmonic=100, n_bands=65, sample_rate=16000,
                                              block_size=160, has_reverb=True))
    model.load_state_dict(state)
    assert (tmp_path / "runs" / "t" / "config.yaml").exists()


@pytest.mark.gpu
def test_training_curve_matches_the_stock_control_net(tmp_path):
    """Same data, seed and noise: 30 optimiser steps on this repo's control-net kernels and on the stock torch.nn
    layers (cuBLAS / cuDNN fp32) end at the same loss to within rounding-driven drift."""
    from ddsp_pytorch_b200 import train
    torch.backends.cudnn.allow_tf32 = False
    train.write_synthetic_dataset(tmp_path / "data" / "train", 16, seconds=1.0, seed=0)
    common = ["--data", str(tmp_path / "data"), "--root", str(tmp_path / "runs"), "--steps", "30", "--batch", "8",
              "--scales", "1024", "512", "256", "128", "--log-every", "5", "--warmup", "2", "--no-graph"]
    ours = train.run(train.parser().parse_args(common + ["--name", "k"]))
    stock = train.run(train.parser().parse_args(common + ["--name", "s", "--stock-control-net"]))
    assert ours["control_net"] == "kernels" and stock["control_net"] == "torch.nn"
    assert abs(ours["first_logged_loss"] - stock["first_logged_loss"]) < 1e-3 * stock["first_logged_loss"]
    assert abs(ours["last_logged_loss"] - stock["last_logged_loss"]) < 2e-2 * stock["last_logged_loss"], (ours, stock)
