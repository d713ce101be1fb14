"""TorchScript export: the scripted module calls the ddsp_b200 custom ops, survives a
save/load round trip, reproduces the eager model and streams without phase or GRU discontinuities."""
import os

import pytest
import torch


def _model(seed=0, **kw):
    import ddsp_pytorch_b200  # noqa: F401
    from ddsp_pytorch_b200.models.decoder import DDSPDecoder
    torch.manual_seed(seed)
    cfg = dict(hidden_size=32, n_harmonic=16, n_bands=65, sample_rate=48000, block_size=512, has_reverb=True)
    cfg.update(kw)
    return DDSPDecoder(**cfg)


def test_script_save_load_on_cpu(tmp_path):
    from ddsp_pytorch_b200.export import export_torchscript
    path = os.path.join(tmp_path, "ddsp.ts")
    export_torchscript(_model(), path, mean_loudness=-30.0, std_loudness=12.0, realtime=True)
    loaded = torch.jit.load(path)
    graph = str(loaded.inlined_graph)
    for op in ("harmonic_controls_fwd", "harmonic_fwd", "noise_fwd"):
        assert "ddsp_b200::" + op in graph
    assert "fftconv_fwd" not in graph, "realtime export leaves the reverb to the host (patches/example.pd)"
    assert {"cache_gru", "phase"} <= {n for n, _ in loaded.named_buffers()}
    assert hasattr(loaded, "reset")


@pytest.mark.gpu
def test_offline_export_matches_eager_model(tmp_path):
    from ddsp_pytorch_b200.export import export_torchscript
    model = _model(seed=1).cuda().eval()
    with torch.no_grad():
        model.reverb.wet.fill_(0.5)
    path = os.path.join(tmp_path, "offline.ts")
    export_torchscript(model, path, realtime=False, device_noise=False)
    scripted = torch.jit.load(path).cuda()
    B, T, bs = 2, 6, 512
    g = torch.Generator().manual_seed(3)
    f0 = (torch.rand(B, T, 1, generator=g) * 300 + 100).cuda()
    ld = torch.randn(B, T, 1, generator=g).cuda()
    with torch.no_grad():
        torch.manual_seed(7)
        ref = model({"pitch": f0, "loudness": ld})["signal"]
        torch.manual_seed(7)
        got = scripted(f0.repeat_interleave(bs, 1), ld.repeat_interleave(bs, 1))
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 1e-5


@pytest.mark.gpu
def test_realtime_export_streams_continuously(tmp_path):
    """Two consecutive 1024-sample buffers through the realtime export == one 2048-sample pass of the
    model (GRU cache and oscillator phase carried); the reference restarted the phase every buffer."""
    from ddsp_pytorch_b200.export import export_torchscript
    model = _model(seed=2, has_reverb=False).cuda().eval()
    with torch.no_grad():                       # silence the (random) noise branch
        model.noise_proj.weight.zero_()
        model.noise_proj.bias.fill_(-60.0)
    path = os.path.join(tmp_path, "rt.ts")
    export_torchscript(model, path, realtime=True)
    rt = torch.jit.load(path).cuda()
    bs, T = 512, 4
    g = torch.Generator().manual_seed(5)
    f0 = (torch.rand(1, T, 1, generator=g) * 200 + 150).cuda()
    ld = torch.randn(1, T, 1, generator=g).cuda()
    pitch, loud = f0.repeat_interleave(bs, 1), ld.repeat_interleave(bs, 1)
    with torch.no_grad():
        whole = model({"pitch": f0, "loudness": ld})["harmonic_audio"]
        a = rt(pitch[:, :1024], loud[:, :1024])
        b = rt(pitch[:, 1024:], loud[:, 1024:])
    assert float((torch.cat([a, b], 1) - whole).abs().max()) < 2e-5
    rt.reset()
    with torch.no_grad():
        again = rt(pitch[:, :1024], loud[:, :1024])
    assert float((again - a).abs().max()) < 1e-6
