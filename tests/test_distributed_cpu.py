"""world_size-2 gloo test of the N>1 path, on CPU: voices sharded over ranks, rank-local
compute (here the CPU oracle stands in for the kernels -- the collective logic is what is tested),
one all-reduce of the parameter gradients; the result must equal the single-process full batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(batch):
    g = torch.Generator().manual_seed(3)
    T, bs, H, NB, L = 6, 32, 5, 9, 40
    d = dict(amp=torch.randn(batch, T, 1, generator=g, dtype=torch.float64),
             dist=torch.randn(batch, T, H, generator=g, dtype=torch.float64),
             mag=torch.randn(batch, T, NB, generator=g, dtype=torch.float64),
             f0=(torch.rand(batch, T, 1, generator=g) * 300 + 100).double(),
             noise=(torch.rand(batch, T, bs, generator=g) * 2 - 1).double(),
             target=0.1 * torch.randn(batch, T * bs, generator=g, dtype=torch.float64))
    rp = {"noise": (torch.rand(L, 1, generator=g, dtype=torch.float64) * 2 - 1), "decay": torch.tensor(2.0).double(),
          "wet": torch.tensor(0.5).double(), "t": (torch.arange(L) / 2000.0).reshape(1, -1, 1).double()}
    return d, rp, dict(bs=bs, sr=2000, scales=[64, 32], overlap=0.75)


def _step(d, rp, cfg, lo, hi):
    from oracle import ddsp_oracle as orc
    return orc.synth_train_step(d["amp"][lo:hi], d["dist"][lo:hi], d["mag"][lo:hi], d["f0"][lo:hi],
                                d["noise"][lo:hi], d["target"][lo:hi], cfg["bs"], cfg["sr"], rp, cfg["scales"],
                                cfg["overlap"])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from ddsp_pytorch_b200.distributed import GradBucket, global_mean, shard_range
    d, rp, cfg = _inputs(4)
    lo, hi = shard_range(4, world, rank)
    loss, grads = _step(d, rp, cfg, lo, hi)
    param_grads = grads[3:]                       # reverb noise / decay / wet: shared parameters
    bucket = GradBucket([g.shape for g in param_grads], "cpu", torch.float64)
    reduced = bucket.all_reduce_mean(param_grads)
    gl = global_mean(loss)
    torch.save({"loss": gl, "grads": [g.clone() for g in reduced], "local": [g.clone() for g in grads[:3]],
                "range": (lo, hi)}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_equals_full_batch(tmp_path):
    world, port = 2, 29611 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    d, rp, cfg = _inputs(4)
    loss, grads = _step(d, rp, cfg, 0, 4)
    r = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(world)]
    for i in range(world):
        assert abs(float(r[i]["loss"]) - float(loss)) < 1e-12
        for got, ref in zip(r[i]["grads"], grads[3:]):
            assert torch.allclose(got, ref, rtol=1e-10, atol=1e-14)
        lo, hi = r[i]["range"]
        # per-voice gradients stay local; with the global-mean loss they are 1/world of the local-mean ones
        for got, ref in zip(r[i]["local"], grads[:3]):
            assert torch.allclose(got / world, ref[lo:hi], rtol=1e-10, atol=1e-14)


def test_shard_range_rejects_uneven_split():
    from ddsp_pytorch_b200.distributed import shard_range
    assert shard_range(64, 8, 3) == (24, 32)
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
