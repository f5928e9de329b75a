"""TEST INFRASTRUCTURE — not product code.

Imports the *real* PQ3D reference decoder from /root/reference on CPU, in-process, so the
restatement in `oracle/restatement.py` can be validated against it and golden vectors can be
generated (`oracle/make_golden.py`).  The reference cannot be imported directly in this image
(fvcore, omegaconf, MinkowskiEngine, torch_scatter, hydra, accelerate are absent and there is no
network), so the few third-party names its hot-path files touch at import time are stubbed in
`sys.modules` (recipe from SURVEY.md §8c):

  * fvcore.common.registry.Registry   (used by modules/build.py:1-9, model/build.py:2,6)
  * omegaconf.OmegaConf.to_container  (used by common/type_utils.py:6-7)
  * MinkowskiEngine(+.MinkowskiPooling) (imported by model/query3d_unified.py:5,
    modules/heads/mask_head.py:3-4; only *used* on the online-voxel path)
  * namespace packages for `modules`, `modules.*`, `model`, `data`, `data.datasets` so the
    auto-import `__init__`s (which pull in MinkowskiEngine / torch_scatter / pointnet2) are skipped
  * modules.layers.pointnet.PointNetPP (imported by modules/vision/object_encoder.py:10)

/root/reference does not exist on the GPU box: nothing under `-m gpu`, `smoke()` or `bench.py`
may import this file.  `available()` says whether the reference tree is present.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("PQ3D_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "modules", "grounding", "query_encoder.py"))


class AttrDict(dict):
    """Stand-in for an omegaconf DictConfig: attribute access + .get()."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attr(v) for v in obj]
    return obj


def _to_plain(obj):
    if isinstance(obj, dict):
        return {k: _to_plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_to_plain(v) for v in obj]
    return obj


class _Registry:
    def __init__(self, name):
        self._name = name
        self._map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._map[o.__name__] = o
                return o
            return deco
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self._map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._map[name]

    def __contains__(self, name):
        return name in self._map


_LOADED = None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _ns_package(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's hot-path modules imported for real."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # --- third-party stubs -------------------------------------------------------------
    _stub("fvcore")
    _stub("fvcore.common")
    _stub("fvcore.common.registry", Registry=_Registry)

    class _OmegaConf:
        @staticmethod
        def to_container(cfg, resolve=True):
            return _to_plain(cfg)
    _stub("omegaconf", OmegaConf=_OmegaConf)

    me = _stub("MinkowskiEngine")
    mp = _stub("MinkowskiEngine.MinkowskiPooling", MinkowskiAvgPooling=object)
    me.MinkowskiPooling = mp

    # --- namespace packages: skip the auto-importing __init__.py files -------------------
    for pkg in ("modules", "modules/grounding", "modules/heads", "modules/layers", "modules/vision",
                "modules/language", "modules/third_party", "modules/third_party/mask3d",
                "model", "data", "data/datasets", "optim", "common"):
        _ns_package(pkg.replace("/", "."), os.path.join(REF_ROOT, pkg))
    _stub("modules.layers.pointnet", PointNetPP=object)

    ns = types.SimpleNamespace()
    ns.build = importlib.import_module("modules.build")
    ns.utils = importlib.import_module("modules.utils")
    ns.weights = importlib.import_module("modules.weights")
    ns.transformers = importlib.import_module("modules.layers.transformers")
    ns.query_encoder = importlib.import_module("modules.grounding.query_encoder")
    ns.mask_head = importlib.import_module("modules.heads.mask_head")
    ns.grounding_head = importlib.import_module("modules.heads.grounding_head")
    ns.object_encoder = importlib.import_module("modules.vision.object_encoder")
    ns.position_embedding = importlib.import_module("modules.third_party.mask3d.position_embedding")
    ns.model_build = importlib.import_module("model.build")
    ns.query3d_unified = importlib.import_module("model.query3d_unified")
    _LOADED = ns
    return ns
