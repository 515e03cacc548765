"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box; nothing in
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this.  It is used by
``oracle/make_golden.py`` (fixture generation) and by CPU tests that are skipped when
the reference tree is absent.

Recipe: SURVEY.md Appendix C.
"""
from __future__ import annotations

import ast
import importlib.abc
import importlib.machinery
import importlib.util
import os
import re
import sys
import types
from unittest import mock

REF_ROOT = "/root/reference/MViT"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "slowfast/models/attention.py"))


_ATT = None


def load_attention():
    """Block-level oracle: attention.py + common.py only (zero third-party deps)."""
    global _ATT
    if _ATT is not None:
        return _ATT
    if "slowfast.models.video_model_builder" in sys.modules:  # full model already imported
        _ATT = sys.modules["slowfast.models.attention"]
        return _ATT
    R = os.path.join(REF_ROOT, "slowfast")
    for n in ("slowfast", "slowfast.models"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m

    def load(name, path):
        s = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(s)
        sys.modules[name] = m
        s.loader.exec_module(m)
        return m

    load("slowfast.models.common", R + "/models/common.py")
    _ATT = load("slowfast.models.attention", R + "/models/attention.py")
    return _ATT


_MISSING = ["fvcore", "detectron2", "pytorchvideo", "iopath", "simplejson", "fairscale", "matplotlib",
            "yacs", "av", "decord", "moviepy", "librosa", "soundfile", "submitit"]


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__name__ = spec.name
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, m):
        pass


class _Node(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v


def load_full_model(yaml_rel: str = "configs/Kinetics/MVITv2_S_16x4.yaml", overrides: dict = None):
    """Build the reference ``MViT`` from its own YAML with stubbed third-party imports.
    Returns (model, cfg).  Must run in a process that has NOT called load_attention()."""
    import torch
    import yaml

    for n in list(sys.modules):
        if n == "slowfast" or n.startswith("slowfast."):
            del sys.modules[n]
    global _ATT
    _ATT = None
    sys.meta_path.insert(0, _Finder())
    import fvcore.common.registry as reg

    class Registry:
        def __init__(self, n):
            self.d = {}

        def register(self, obj=None):
            if obj is None:
                def deco(o):
                    self.d[o.__name__] = o
                    return o
                return deco
            self.d[obj.__name__] = obj

        def get(self, n):
            return self.d[n]

    reg.Registry = Registry
    import detectron2.layers as d2l
    import pytorchvideo.layers.batch_norm as pbn
    import pytorchvideo.layers.swish as psw

    for n in ("NaiveSyncBatchNorm1d", "NaiveSyncBatchNorm3d"):
        setattr(pbn, n, type(n, (torch.nn.Module,), {}))
    psw.Swish = type("Swish", (torch.nn.Module,), {})
    d2l.ROIAlign = type("ROIAlign", (torch.nn.Module,), {})
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import slowfast.models.video_model_builder as vmb

    C = _Node()
    for line in open(os.path.join(REF_ROOT, "slowfast/config/defaults.py")):
        m = re.match(r"^_C\.([A-Z_0-9\.]+) = (.*)$", line.rstrip("\n"))
        if not m:
            continue
        *path, leaf = m.group(1).split(".")
        node = C
        for q in path:
            node = node.setdefault(q, _Node())
        if m.group(2).startswith("CfgNode"):
            node.setdefault(leaf, _Node())
        else:
            try:
                node[leaf] = ast.literal_eval(m.group(2))
            except Exception:
                pass

    def merge(n, d):
        for k, v in d.items():
            if isinstance(v, dict):
                merge(n.setdefault(k, _Node()), v)
            else:
                n[k] = list(ast.literal_eval(v)) if isinstance(v, str) and v.startswith("(") else v

    merge(C, yaml.safe_load(open(os.path.join(REF_ROOT, yaml_rel))))
    C.NUM_GPUS = 0
    for dotted, v in (overrides or {}).items():  # e.g. {"DATA.TRAIN_CROP_SIZE_RECT": [128, 96]}
        *path, leaf = dotted.split(".")
        node = C
        for q in path:
            node = node.setdefault(q, _Node())
        node[leaf] = v
    model = vmb.MViT(C)
    return model, C
