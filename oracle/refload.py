"""Import the LIVE reference (``/root/reference``) with its unused dependencies
stubbed.  Only usable in the build container; used by
``oracle/make_golden.py`` to pin the restatement.  Never imported at test/bench
time on the GPU box."""
import os
import sys
import types

REF = os.environ.get("DAD_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "distillanydepth"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def install_stubs():
    import torch.nn as nn

    class ModelMixin(nn.Module):
        pass

    class ConfigMixin:
        pass

    def register_to_config(f):
        return f

    for name in ("diffusers", "diffusers.models", "diffusers.models.modeling_utils",
                 "diffusers.configuration_utils", "timm", "timm.models",
                 "timm.models.vision_transformer", "omegaconf", "matplotlib", "matplotlib.pyplot",
                 "detectron2", "detectron2.utils", "detectron2.utils.comm", "detectron2.engine",
                 "yapf", "yapf.yapflib", "yapf.yapflib.yapf_api", "addict", "xformers_absent"):
        if name not in sys.modules or name.startswith(("diffusers", "timm", "detectron2", "yapf")):
            _stub(name)
    _stub("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _stub("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _stub("timm.models.vision_transformer", vit_large_patch16_224=None, vit_large_patch14_224=None)
    _stub("omegaconf", OmegaConf=object)
    _stub("matplotlib", cm=types.SimpleNamespace(), use=lambda *a, **k: None)
    _stub("matplotlib.pyplot")
    _stub("detectron2.utils", comm=sys.modules["detectron2.utils.comm"])
    _stub("detectron2.engine", launch=lambda *a, **k: None)
    _stub("yapf.yapflib.yapf_api", FormatCode=lambda *a, **k: None)
    _stub("addict", Dict=dict)
    for p in (REF, os.path.join(REF, "tools")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_models():
    """Returns (DepthAnythingV2, DepthAnything) reference classes."""
    install_stubs()
    from distillanydepth.depth_anything_v2.dpt import DepthAnythingV2
    from distillanydepth.modeling.archs.dam.dam import DepthAnything
    return DepthAnythingV2, DepthAnything


def load_losses():
    """Returns the reference tools/train_distillation.py module."""
    install_stubs()
    import importlib
    return importlib.import_module("train_distillation")
