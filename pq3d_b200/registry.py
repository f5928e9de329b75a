"""Plugging into the reference's own registries (modules/build.py:1-9, model/build.py:6).

The reference resolves `cfg.model.unified_encoder.name`, `cfg.model.mask_head.name`, ... through
fvcore `Registry` objects that auto-import every file in their package directory.  When the
reference is importable, `register_into_reference()` overrides the entries for the hot-path classes
with the B200 implementations, so an unmodified trainer (`trainer/build.py:98` -> `build_model(cfg)`)
builds them; INTEGRATION.md shows the 6-line plugin file a maintainer would drop into
`modules/grounding/` instead.
"""
from __future__ import annotations


def register_into_reference(override_model: bool = True) -> list:
    from modules import build as ref_build            # the reference's modules/build.py
    from . import mask_head, query3d_unified, query_encoder
    done = []

    def put(registry, cls):
        store = getattr(registry, "_obj_map", None)
        if store is None:
            store = getattr(registry, "_map")
        store[cls.__name__] = cls
        done.append(f"{registry._name}:{cls.__name__}")

    put(ref_build.GROUNDING_REGISTRY, query_encoder.QueryMaskEncoder)
    put(ref_build.GROUNDING_REGISTRY, query_encoder.QueryEncoder)
    put(ref_build.HEADS_REGISTRY, mask_head.MaskHeadSegLevel)
    put(ref_build.HEADS_REGISTRY, query3d_unified.GroundHead)
    put(ref_build.VISION_REGISTRY, query3d_unified.ObjectEncoder)
    if override_model:
        from model import build as model_build
        put(model_build.MODEL_REGISTRY, query3d_unified.Query3DUnified)
    return done
