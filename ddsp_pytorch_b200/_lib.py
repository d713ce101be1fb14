"""Loads the native libraries.  There is no fallback: if they are missing and cannot be built,
importing the package fails."""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
CORE_PATH = os.path.join(_PKG, "libddsp_b200.so")
TORCH_PATH = os.path.join(_PKG, "libddsp_b200_torch.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG), "include")

_lock = threading.Lock()
_state = {"loaded": False, "core": None}


def load():
    """Idempotent: dlopen libddsp_b200.so (C ABI) and register the ddsp_b200:: torch ops."""
    with _lock:
        if _state["loaded"]:
            return _state["core"]
        if not (os.path.exists(CORE_PATH) and os.path.exists(TORCH_PATH)):
            from . import build
            build.build_all()
        # RTLD_GLOBAL so the torch shim resolves the C ABI from the very same image
        _state["core"] = ctypes.CDLL(CORE_PATH, mode=ctypes.RTLD_GLOBAL)
        torch.ops.load_library(TORCH_PATH)
        if torch.ops.ddsp_b200.abi_version() != 1:
            raise ImportError("libddsp_b200_torch.so / libddsp_b200.so ABI mismatch; rebuild with "
                              "`python -m ddsp_pytorch_b200.build --force`")
        _state["loaded"] = True
        return _state["core"]


def core_library() -> ctypes.CDLL:
    """ctypes handle on the C ABI (tests call the kernels through it with raw device pointers)."""
    return load()


ops = None


def get_ops():
    load()
    return torch.ops.ddsp_b200
