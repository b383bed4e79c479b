"""Checkpoint I/O next to the hot path: read / write the reference's on-disk formats and map
between its two key layouts, so real DAv2 / Distill-Any-Depth weights drop into the native models.

Reference behaviour mirrored here:
  * ``create_teacher_model`` (``tools/train_distillation.py:743-793``): ``.safetensors`` via ``load_file``, anything else via
    ``torch.load`` (unwrapping a ``state_dict`` entry), ``pretrained.`` -> ``backbone.`` for the teacher class;
  * ``vit_large`` (``modeling/backbones/vit/ViT_DINO.py:1372-1388``): ``blocks.N`` -> ``blocks.0.N`` when the flat layout does
    not fit the chunked module;
  * ``tools/convert_checkpoint.py:7-28``: the same prefix rewrite as an offline file conversion;
  * ``save_file(student_model.state_dict(), ...)`` (``tools/train_distillation.py:1612-1615``).
Unlike the reference's silent ``strict=False`` fallback, a key that still does not fit after remapping is an error
unless ``strict=False`` is asked for.
"""
import re

import torch

from .dam import DepthAnything, student_to_teacher_keys

_CHUNKED_BLOCK = re.compile(r"^(pretrained|backbone)\.blocks\.0\.(\d+)\.")


def read_state_dict(path, device="cpu"):
    """``.safetensors`` -> ``safetensors.torch.load_file``; otherwise ``torch.load`` (``state_dict`` entry unwrapped)."""
    if str(path).endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(str(path), device=str(device))
    sd = torch.load(str(path), map_location=device, weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    return sd


def to_student_layout(sd):
    """Any of the reference's layouts -> ``pretrained.blocks.N.*`` (DepthAnythingV2, ``dpt.py:196``)."""
    out = {}
    for k, v in sd.items():
        k = _CHUNKED_BLOCK.sub(lambda m: f"{m.group(1)}.blocks.{m.group(2)}.", k)  # blocks.0.N -> blocks.N
        if k.startswith("backbone."):
            k = "pretrained." + k[len("backbone."):]
        out[k] = v
    return out


def to_teacher_layout(sd):
    """Any of the reference's layouts -> ``backbone.blocks.0.N.*`` (DepthAnything, ``dam.py:333-360``)."""
    return student_to_teacher_keys(to_student_layout(sd))


def remap_for(model, sd):
    """State dict in the key layout ``model`` expects."""
    return to_teacher_layout(sd) if isinstance(model, DepthAnything) else to_student_layout(sd)


def load_checkpoint(model, path, strict=True, device="cpu"):
    """Load ``path`` into ``model`` (either native class), remapping the key layout as needed.  Tensors are cast to
    the parameter dtype by ``load_state_dict``; the packed bf16 operands are rebuilt lazily at the next forward."""
    sd = remap_for(model, read_state_dict(path, device))
    return model.load_state_dict(sd, strict=strict)


def save_checkpoint(model, path):
    """``save_file(model.state_dict(), path)`` for ``.safetensors``, ``torch.save`` otherwise."""
    sd = {k: v.detach().to("cpu").contiguous() for k, v in model.state_dict().items()}
    if str(path).endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file(sd, str(path))
    else:
        torch.save(sd, str(path))


def convert_checkpoint(input_path, output_path, layout="teacher"):
    """Offline conversion (``tools/convert_checkpoint.py``): ``layout='teacher'`` writes ``backbone.blocks.0.N.*`` keys,
    ``'student'`` writes ``pretrained.blocks.N.*``."""
    sd = read_state_dict(input_path)
    sd = to_teacher_layout(sd) if layout == "teacher" else to_student_layout(sd)
    sd = {k: v.contiguous() for k, v in sd.items()}
    if str(output_path).endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file(sd, str(output_path))
    else:
        torch.save(sd, str(output_path))
    return sorted(sd)
