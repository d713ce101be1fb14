"""Builds tests/host/regfft_host_test.cu for the HOST (nvcc, no GPU) and runs it: the register-tiled FFT
stage functions of ddsp_pytorch_b200/csrc/regfft.cuh are __host__ __device__, so the exact code the
kernels run is checked here, thread by thread, against a float64 DFT for every size 64..4096."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_register_fft_stages_on_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "regfft_host"
    subprocess.run([nvcc, "-O2", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "ddsp_pytorch_b200", "csrc"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "host", "regfft_host_test.cu")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "N=4096 inverse" in out.stdout


def test_lane_packed_fft_stages_on_host(tmp_path):
    """pfft.cuh (two transforms in the two lanes of every value, used by the fused spectral loss): forward and
    inverse of every size against a float64 DFT, both lanes, results read from registers through slot_of_q."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = tmp_path / "pfft_host"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "ddsp_pytorch_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "host", "pfft_host_test.cu")], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "N=4096 inverse" in out.stdout
