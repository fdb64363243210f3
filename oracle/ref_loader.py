"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference adapter classes in place.

Only usable where ``/root/reference`` exists (the build container). It never travels
to the GPU box; what travels are the golden fixtures ``oracle/make_golden.py`` writes
under ``tests/golden/`` and the functional restatement in ``oracle/avmoe_oracle.py``.

The reference modules import a handful of third-party packages that are absent here
(timm, loralib, ipdb, torchlibrosa, h5py, librosa ...).  None of them is touched by the
two hot-path classes (``ExpertAdapter``/``MoEAdapter``), so they are replaced with
empty stub modules for the duration of the import.

Reference classes per variant:
  ave  -> AVMOE/AVE/nets/net_trans_v3.py:296-487
  avqa -> AVMOE/AVQA/net_grd_avst/net_avst_v2.py:215-399
  avvp -> AVMOE/AVVP/nets/mgn.py:39-224
  s4   -> AVMOE/AVS/avs_scripts/avs_s4/model/PVT_AVSModel_v2.py:90-318
  ms3  -> AVMOE/AVS/avs_scripts/avs_ms3/model/PVT_AVSModel_v2.py:282-501
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

def _reference_root() -> str:
    """/root/reference in the build container; on the GPU box the unmodified copies staged by oracle/stage_reference.py"""
    env = os.environ.get("AVMOE_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/AVMOE"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _reference_root()

_VARIANTS = {
    "ave": ("AVMOE/AVE", "nets.net_trans_v3"),
    "avqa": ("AVMOE/AVQA/net_grd_avst", "net_avst_v2"),
    "avvp": ("AVMOE/AVVP", "nets.mgn"),
    "s4": ("AVMOE/AVS/avs_scripts/avs_s4", "model.PVT_AVSModel_v2"),
    "ms3": ("AVMOE/AVS/avs_scripts/avs_ms3", "model.PVT_AVSModel_v2"),
}

# top-level module names the five task directories fight over
_COLLIDING = ("nets", "model", "utils", "esc_config", "htsat", "layers", "visual_net",
              "net_avst_v2", "models", "config", "torchvggish")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "AVMOE"))


class _Anything:
    """Attribute sink: any attribute access / call / subclassing succeeds."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]  # behaves as a pass-through decorator
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None or not getattr(mod, "__avmoe_stub__", False):
        if mod is not None:
            return mod  # a real install wins
        mod = types.ModuleType(name)
        mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        mod.__path__ = []  # behaves as a package
        mod.__avmoe_stub__ = True
        mod.__getattr__ = lambda attr, _n=name: _Anything()  # type: ignore[attr-defined]
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_stubs() -> None:
    import torch.nn as nn
    import transformers.activations  # noqa: F401  (must be real, import before stubbing)

    class _Identity(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, x, *a, **k):
            return x

    def _passthrough_decorator(fn=None, *a, **k):
        return fn

    _stub("ipdb", set_trace=lambda *a, **k: None)
    _stub("loralib")
    _stub("h5py")
    _stub("librosa")
    _stub("soundfile")
    _stub("timm", create_model=lambda *a, **k: _Identity())
    _stub("timm.models")
    _stub("timm.models.vision_transformer", Attention=_Identity, _cfg=lambda *a, **k: {})
    _stub("timm.models.layers", trunc_normal_=lambda t, *a, **k: t,
          to_2tuple=lambda x: (x, x) if not isinstance(x, tuple) else x, DropPath=_Identity)
    _stub("timm.models.registry", register_model=_passthrough_decorator)
    _stub("torchlibrosa", Spectrogram=_Identity, LogmelFilterBank=_Identity, SpecAugmentation=_Identity)
    _stub("torchlibrosa.stft", Spectrogram=_Identity, LogmelFilterBank=_Identity,
          STFT=_Identity, ISTFT=_Identity, magphase=lambda *a, **k: None)
    _stub("torchlibrosa.augmentation", SpecAugmentation=_Identity)
    try:
        import torchvision  # noqa: F401
    except Exception:
        _stub("torchvision")
        _stub("torchvision.models", resnet18=lambda *a, **k: _Identity())
        _stub("torchvision.transforms")


def load_variant(variant: str):
    """Return (ExpertAdapter, MoEAdapter) classes of the reference for ``variant``."""
    if variant not in _VARIANTS:
        raise KeyError(f"unknown variant {variant!r}; expected one of {sorted(_VARIANTS)}")
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    rel_dir, modname = _VARIANTS[variant]
    task_dir = os.path.join(REFERENCE_ROOT, rel_dir)
    _install_stubs()
    # purge names the task directories collide on, then import with the task dir first on sys.path
    for key in list(sys.modules):
        root = key.split(".")[0]
        if root in _COLLIDING and not getattr(sys.modules[key], "__avmoe_stub__", False):
            f = getattr(sys.modules[key], "__file__", "") or ""
            if f.startswith(REFERENCE_ROOT) or root in ("nets", "model", "net_avst_v2"):
                del sys.modules[key]
    sys.path.insert(0, task_dir)
    prev_dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    try:
        mod = importlib.import_module(modname)
    finally:
        sys.dont_write_bytecode = prev_dont_write
        sys.path.remove(task_dir)
    return mod.ExpertAdapter, mod.MoEAdapter
