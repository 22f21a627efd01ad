"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference sources.

This module makes the reference's hot-path files importable in the build
container straight from ``/root/reference`` (read-only), by pre-seeding
``sys.modules`` with tiny stand-ins for the third-party packages that are not
installed here (``omegaconf``, ``bpy``, ``skimage``, ``gpytoolbox``).  Nothing is
copied: the reference code that runs is the reference's own.

It is used by exactly two things:

* ``oracle/make_golden.py`` -- generates the committed fixtures under
  ``tests/golden/`` (``/root/reference`` does not exist on the GPU box, so the
  vectors travel, the reference does not);
* ``tests/test_oracle_vs_reference.py`` -- re-validates the restatement in
  ``oracle/field_oracle.py`` against the live reference when it is present
  (skipped otherwise).

The product package ``sculptmate_b200`` never imports anything under ``oracle/``.

Reference entry points this shim exposes (all paths relative to /root/reference):
  TripoSR/tsr/models/nerf_renderer.py:15-91   TriplaneNeRFRenderer.query_triplane
  TripoSR/tsr/models/network_utils.py:35-124  NeRFMLP
  TripoSR/tsr/models/isosurface.py:18-54      MarchingCubeHelper
  TripoSR/tsr/utils.py:152-252                chunk_batch / scale_tensor / get_activation
  StableFast/sf3d/models/isosurface.py:24-229 MarchingTetrahedraHelper
  StableFast/sf3d/models/network.py:148-208   MaterialMLP
"""
from __future__ import annotations

import dataclasses
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SCULPTMATE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "TripoSR", "tsr"))


# --------------------------------------------------------------------------
# omegaconf stand-in: just enough for BaseModule.__init__ -> parse_structured
# (TripoSR/tsr/utils.py:16-18,44-59) and TSR.from_pretrained (system.py:51-66).
# --------------------------------------------------------------------------
class _DictConfig(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _to_cfg(obj):
    if isinstance(obj, dict):
        return _DictConfig({k: _to_cfg(v) for k, v in obj.items()})
    return obj


class _OmegaConf:
    @staticmethod
    def structured(cls_or_obj):
        out = _DictConfig()
        if dataclasses.is_dataclass(cls_or_obj):
            for f in dataclasses.fields(cls_or_obj):
                if f.default is not dataclasses.MISSING:
                    out[f.name] = f.default
                elif f.default_factory is not dataclasses.MISSING:  # type: ignore[misc]
                    out[f.name] = f.default_factory()  # type: ignore[misc]
        return out

    @staticmethod
    def merge(*cfgs):
        out = _DictConfig()
        for c in cfgs:
            if c is None:
                continue
            for k, v in dict(c).items():
                out[k] = _to_cfg(v)
        return out

    @staticmethod
    def load(path):
        import yaml

        with open(path) as f:
            return _to_cfg(yaml.safe_load(f))

    @staticmethod
    def resolve(cfg):
        # the only interpolation in the reference configs is
        # ``${tokenizer.num_channels}`` (TripoSR/checkpoints/config.yaml:14)
        def walk(node, root):
            for k, v in list(node.items()):
                if isinstance(v, dict):
                    walk(v, root)
                elif isinstance(v, str) and v.startswith("${") and v.endswith("}"):
                    cur = root
                    for part in v[2:-1].split("."):
                        cur = cur[part]
                    node[k] = cur

        walk(cfg, cfg)

    @staticmethod
    def create(obj=None):
        return _to_cfg(obj or {})


def _install_stubs() -> None:
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")
        m.DictConfig = _DictConfig
        m.OmegaConf = _OmegaConf
        sys.modules["omegaconf"] = m
    if "bpy" not in sys.modules:
        sys.modules["bpy"] = types.ModuleType("bpy")
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        ms = types.ModuleType("skimage.measure")

        def _absent(*a, **k):
            raise ModuleNotFoundError(
                "scikit-image is not installed in this environment; patch "
                "skimage.measure.marching_cubes with the in-repo MC oracle"
            )

        ms.marching_cubes = _absent
        sk.measure = ms
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = ms
    if "gpytoolbox" not in sys.modules:
        sys.modules["gpytoolbox"] = types.ModuleType("gpytoolbox")


def load_triposr():
    """Return the reference's ``tsr`` package (unmodified, from /root/reference)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    import PIL.Image  # noqa: F401  (tsr/utils.py uses PIL.Image after bare import)

    p = os.path.join(REFERENCE_ROOT, "TripoSR")
    if p not in sys.path:
        sys.path.insert(0, p)
    import tsr.models.isosurface as iso
    import tsr.models.nerf_renderer as rend
    import tsr.models.network_utils as net
    import tsr.utils as utils

    return types.SimpleNamespace(isosurface=iso, nerf_renderer=rend, network_utils=net, utils=utils)


def load_sf3d():
    """Return the reference's ``sf3d`` hot-path modules (unmodified)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    import PIL.Image  # noqa: F401

    p = os.path.join(REFERENCE_ROOT, "StableFast")
    if p not in sys.path:
        sys.path.insert(0, p)
    import sf3d.models.isosurface as iso
    import sf3d.models.network as net

    return types.SimpleNamespace(isosurface=iso, network=net)


def load_sf3d_system():
    """The reference's own ``sf3d.system`` module (for SF3D.query_triplane and
    SF3D.triplane_to_meshes, system.py:141-198).  Its image->triplane imports need
    third-party versions that are not installed here (transformers API drift, open_clip),
    so those OUT-OF-SCOPE sub-modules are pre-seeded with empty stand-ins exposing only the
    names system.py imports; the two methods that run are the reference's, unmodified."""
    ref = load_sf3d()
    names = {
        "sf3d.models.tokenizers.image": ["DINOV2SingleImageTokenizer"],
        "sf3d.models.tokenizers.triplane": ["TriplaneLearnablePositionalEmbedding"],
        "sf3d.models.transformers.backbone": ["TwoStreamInterleaveTransformer"],
        "sf3d.models.global_estimator.multi_head_estimator": ["MultiHeadEstimator"],
        "sf3d.models.image_estimator.clip_based_estimator": ["ClipBasedHeadEstimator"],
        "sf3d.models.camera": ["LinearCameraEmbedder"],
        "sf3d.texture_baker.baker": ["TextureBaker"],
    }
    for mod, attrs in names.items():
        if mod not in sys.modules:
            parent = mod.rsplit(".", 1)[0]
            if parent not in sys.modules:
                pm = types.ModuleType(parent)
                pm.__path__ = []
                sys.modules[parent] = pm
            m = types.ModuleType(mod)
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[mod] = m
    import sf3d.system as system

    ref.system = system
    return ref


TRIPOSR_DECODER_CFG = dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu")
TRIPOSR_RENDERER_CFG = dict(
    radius=0.87,
    feature_reduction="concat",
    density_activation="exp",
    density_bias=-1.0,
    num_samples_per_ray=128,
)


def make_reference_decoder(seed: int):
    """NeRFMLP with the reference's default init (config.yaml:25-30) under a seed."""
    import torch

    ref = load_triposr()
    torch.manual_seed(seed)
    return ref.network_utils.NeRFMLP(dict(TRIPOSR_DECODER_CFG))


def make_reference_renderer(chunk_size: int = 8192):
    ref = load_triposr()
    r = ref.nerf_renderer.TriplaneNeRFRenderer(dict(TRIPOSR_RENDERER_CFG))
    r.set_chunk_size(chunk_size)
    return r
