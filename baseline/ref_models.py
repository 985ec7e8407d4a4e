"""Construct the UNMODIFIED reference models from baseline/_ref/core (installed by
tools/install_ref.sh) with seeded random-init weights -- test / benchmark infrastructure only.

Used by tests/test_gpu_e2e_reference.py (north_star tolerance 2: <= 1e-2 px end-point error after 32
GRU iterations with identical random-init weights) and by bench.py's cfg5 leg (end-to-end
frames/s with the new corr path swapped in).  Nothing here is imported by the product package.

Reference facts this relies on:
  * core/raft_stereo.py:86-103 RAFTStereo(args) constructs with random init; it reads args.proxy (:191).
  * core/SMoEStereo.py:87-124 SMoEStereo(args) hard-loads a VFM checkpoint with torch.load
    (core/depthanything_v2/dpt.py:312-317) and only *prints* missing keys
    (core/utils/utils.py:107-134), so stubbing torch.load to return {} gives a random-init model.
  * core/corr.py:5-8 does a bare `import corr_sampler` inside try/except at import time.
"""
from __future__ import annotations

import contextlib
import os
import sys
import warnings
from argparse import Namespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "core", "corr.py"))


def ref_sampler_available() -> bool:
    return os.path.isfile(os.path.join(HERE, "..", "oracle", "_ref", "corr_sampler_ref.so"))


def import_core(with_ref_sampler: bool = True):
    """Put baseline/_ref on sys.path and import the reference's core.corr.  When the reference's
    own CUDA sampler has been built (oracle/_ref), it is registered as `corr_sampler` FIRST so that
    core/corr.py binds it, exactly as a user who ran sampler/setup.py would have it."""
    if not available():
        raise RuntimeError("baseline/_ref/core not installed (run tools/install_ref.sh in the build container)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if with_ref_sampler and "core.corr" not in sys.modules and ref_sampler_available():
        root = os.path.join(HERE, "..")
        if root not in sys.path:
            sys.path.insert(0, root)
        from oracle.build_ref_sampler import load_module
        mod = load_module()
        if mod is not None and "corr_sampler" not in sys.modules:
            sys.modules["corr_sampler"] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import core.corr as ref_corr
    return ref_corr


def make_args(corr_implementation="reg", mixed_precision=False, model="raft", **over) -> Namespace:
    """The reference's own defaults (train_stereo_sceneflow.py:359-395, evaluate_stereo.py:1576-1613)."""
    a = dict(hidden_dims=[128] * 3, corr_implementation=corr_implementation, shared_backbone=False,
             corr_levels=4, corr_radius=4, n_downsample=2, context_norm="batch", slow_fast_gru=False,
             n_gru_layers=3, mixed_precision=mixed_precision, max_disp=192, proxy=False)
    if model == "smoe":
        a.update(context_norm="instance", max_disp=256, vfm_type="damv2", vfm_size="vitl", peft_type="smoe",
                 tunable_layers=list(range(12)), use_layer_selection=True, VFM_dims=[256] * 3, recon=False,
                 feature_maps=False)
    a.update(over)
    return Namespace(**a)


@contextlib.contextmanager
def _no_checkpoint_load():
    import torch
    real = torch.load
    torch.load = lambda *a, **k: {}
    try:
        yield
    finally:
        torch.load = real


def build_model(kind: str, args: Namespace, seed: int = 0, device=None):
    """kind: 'raft' -> core.raft_stereo.RAFTStereo, 'smoe' -> core.SMoEStereo.SMoEStereo; eval mode."""
    import io
    import torch
    import_core()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(seed)
        if kind == "raft":
            from core.raft_stereo import RAFTStereo
            m = RAFTStereo(args)
        elif kind == "smoe":
            from core.SMoEStereo import SMoEStereo
            with _no_checkpoint_load(), contextlib.redirect_stdout(io.StringIO()):
                m = SMoEStereo(args)
        else:
            raise ValueError(kind)
    m.eval()
    if device is not None:
        m.to(device)
    return m


def synthetic_pair(B, H, W, max_disp=48.0, seed=0, device=None):
    """A textured left image and a right image that is the left one shifted by a smooth disparity
    field (values 0..255 like the reference's loaders, dataloader/stereo_datasets)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    tex = torch.rand(B, 3, H // 4, W // 4, generator=g)
    left = F.interpolate(tex, size=(H, W), mode="bilinear", align_corners=False)
    left = left + 0.25 * torch.rand(B, 3, H, W, generator=g)
    ys = torch.linspace(0, 1, H).view(1, H, 1).expand(B, H, W)
    xs = torch.linspace(0, 1, W).view(1, 1, W).expand(B, H, W)
    disp = max_disp * (0.35 + 0.4 * ys + 0.25 * torch.sin(6.0 * xs))
    gx = (torch.arange(W).view(1, 1, W) + disp) / (W - 1) * 2 - 1
    gy = (torch.arange(H).view(1, H, 1).expand(B, H, W)) / (H - 1) * 2 - 1
    right = F.grid_sample(left, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="border",
                          align_corners=True)
    left, right = (left * 255.0 / 1.25).contiguous(), (right * 255.0 / 1.25).contiguous()
    if device is not None:
        left, right = left.to(device), right.to(device)
    return left, right
