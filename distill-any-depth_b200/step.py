"""Forward half of the reference's distillation training step (``tools/train_distillation.py:1503-1560``):
student forward on the global and the local image, teacher forward on the local image, then the five
loss terms and their lambda-weighted sum.  Everything runs on the caller's CUDA stream through the C ABI;
nothing synchronises the host.  Under ``torch.distributed`` (one process per GPU, images sharded by rank)
each loss is finished from its (numerator, denominator) partials by ONE all-reduce (``dist.finish_losses``),
so the value equals the single-process full-batch loss.

``distillation_train_step`` is the whole step including the backward / optimiser half (``:1556-1575``, SURVEY.md 8f
N1): the student forwards are differentiable (``precision="fp32"``, or ``"bf16"`` with ``bf16_backward=True``), the
losses are the public autograd-aware functions, and ``batch_loss.backward()`` fills ``.grad`` of the student.
"""
import torch

import torch.distributed as tdist

from . import losses
from .dist import finish_losses, shard_loss_weights, allreduce_gradients

# scripts/train_test.sh:21-25
DEFAULT_LAMBDAS = dict(sc=0.5, lg=0.5, feat=1.0, grad=0.2, hdn=0.8)


def distillation_step_partials(student_model, teacher_model, global_image, local_image, normalization="hybrid",
                               use_hdn_loss=True, hdn_level=3, hdn_variant="dr", dedup_student=False):
    """The three forwards and the per-rank (numerator, denominator) partials of the five loss terms:
    ``{name: (kind, float64[2])}``.  No collective: capturable in a CUDA graph on any number of ranks."""
    if use_hdn_loss and hdn_variant != "dr":
        # the reference passes mask_valid_list=None for any other variant and crashes (:1547)
        raise NotImplementedError("train() only wires hdn_variant='dr' (tools/train_distillation.py:1547)")
    with torch.no_grad():
        student_global_disp, _ = student_model(global_image)
        if dedup_student and global_image is local_image:
            student_local_disp, student_local_features = student_global_disp, _
        else:
            student_local_disp, student_local_features = student_model(local_image)
        teacher_local_disp, teacher_local_features = teacher_model(local_image)
        parts = {}
        _, p, _, _ = losses._distill(student_local_disp, teacher_local_disp, normalization, want_partials=True)
        parts["sc_loss"] = ("distill", p)
        _, p, _, _ = losses._distill(student_global_disp, student_local_disp, normalization, want_partials=True)
        parts["lg_loss"] = ("distill", p)
        parts["feat_loss"] = ("feat", losses._feat(student_local_features, teacher_local_features, want_partials=True)[1])
        parts["grad_loss"] = ("grad", losses._grad(student_local_disp, want_partials=True)[1])
        if use_hdn_loss:
            parts["hdn_loss"] = ("hdn", losses.hdn_loss_dr(student_local_disp, teacher_local_disp, None, hdn_level,
                                                           want_partials=True)[1])
    return parts


def combine_step_losses(out, lambdas=None, use_hdn_loss=True):
    """``batch_loss`` = the lambda-weighted sum (``:1556-1560``) of finished loss scalars; fills ``hdn_loss`` / ``batch_loss``."""
    lam = dict(DEFAULT_LAMBDAS)
    lam.update(lambdas or {})
    if not use_hdn_loss:
        out["hdn_loss"] = torch.zeros((), device=out["sc_loss"].device)
    batch = (lam["sc"] * out["sc_loss"] + lam["lg"] * out["lg_loss"] + lam["feat"] * out["feat_loss"]
             + lam["grad"] * out["grad_loss"])
    if use_hdn_loss:
        batch = batch + lam["hdn"] * out["hdn_loss"]
    out["batch_loss"] = batch
    return out


def distillation_step_losses(student_model, teacher_model, global_image, local_image, normalization="hybrid",
                             lambdas=None, use_hdn_loss=True, hdn_level=3, hdn_variant="dr", dedup_student=False):
    """Returns ``dict(sc_loss, lg_loss, feat_loss, grad_loss, hdn_loss, batch_loss)`` of fp32 device scalars.

    ``dedup_student``: the reference runs the student twice, on ``global_image`` and ``local_image``
    (``:1509-1510``) even when both are the same tensor (NYU path, ``:1480-1484``); with ``True`` and identical
    inputs the second, bit-identical forward is elided.
    """
    parts = distillation_step_partials(student_model, teacher_model, global_image, local_image, normalization,
                                       use_hdn_loss, hdn_level, hdn_variant, dedup_student)
    with torch.no_grad():
        out = finish_losses(parts)  # one all-reduce of the stacked partials when torch.distributed is initialised
        return combine_step_losses(out, lambdas, use_hdn_loss)


def distillation_train_step(student_model, teacher_model, global_image, local_image, optimizer=None, normalization="hybrid",
                            lambdas=None, use_hdn_loss=True, hdn_level=3, grad_clip=None, data_parallel=None):
    """One update of the reference loop (``tools/train_distillation.py:1503-1575``): teacher forward without
    gradients, the two student forwards with gradients, the five losses, ``batch_loss.backward()`` and - when an
    ``optimizer`` is given - ``optimizer.step()`` / ``zero_grad()``.  Returns the dict of detached loss scalars.

    ``data_parallel`` (default: on when ``torch.distributed`` is initialised with more than one rank): the images are this
    rank's shard of the global batch.  Every local loss is scaled by its shard weight ``(den_r + eps) / (den + eps)``
    (``dist.shard_loss_weights``: one all-reduce of five denominators) before ``backward()``, then the gradients are
    SUM-all-reduced in flat fp32 buckets (``dist.allreduce_gradients``) - exactly the gradient of the single-process
    full-batch loss the reference computes.  The student backward is ONE library call that fills every ``.grad`` at once,
    so the collective runs after it rather than bucket by bucket under it; at ViT-B size it moves 390 MB (~1 ms over
    NVLink against a ~57 ms step).  The returned scalars are the full-batch values on every rank."""
    lam = dict(DEFAULT_LAMBDAS)
    lam.update(lambdas or {})
    if not torch.is_grad_enabled():
        raise RuntimeError("distillation_train_step needs grad mode")
    if getattr(student_model, "precision", "fp32") == "bf16" and not getattr(student_model, "bf16_backward", True):
        raise RuntimeError("the student runs the inference-only bf16 forward: set student.bf16_backward = True "
                           "(or student.precision = 'fp32')")
    with torch.no_grad():
        teacher_local_disp, teacher_local_features = teacher_model(local_image)
    student_global_disp, _ = student_model(global_image)
    student_local_disp, student_local_features = student_model(local_image)
    out = {}
    out["sc_loss"] = losses.distillation_loss(student_local_disp, teacher_local_disp, normalization)
    out["lg_loss"] = losses.distillation_loss(student_global_disp, student_local_disp, normalization)
    out["feat_loss"] = losses.feature_distillation_loss(student_local_features, teacher_local_features)
    out["grad_loss"] = losses.gradient_preservation_loss(student_local_disp)
    batch = (lam["sc"] * out["sc_loss"] + lam["lg"] * out["lg_loss"] + lam["feat"] * out["feat_loss"]
             + lam["grad"] * out["grad_loss"])
    if use_hdn_loss:
        ctx = losses.get_contexts_dr(hdn_level, teacher_local_disp, None)
        out["hdn_loss"] = losses.compute_hdn_loss(losses.SSILoss(), student_local_disp, teacher_local_disp, ctx)
        batch = batch + lam["hdn"] * out["hdn_loss"]
    else:
        out["hdn_loss"] = torch.zeros((), device=local_image.device)
    out["batch_loss"] = batch
    if data_parallel is None:
        data_parallel = tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
    if data_parallel:
        dev = student_local_disp.device
        sd = student_local_disp

        def den(v):
            return torch.tensor([0.0, float(v)], dtype=torch.float64, device=dev)
        with torch.no_grad():   # denominators of the five ratios on this shard (the HDN one is data dependent)
            parts = {"sc_loss": ("distill", den(sd.numel())), "lg_loss": ("distill", den(sd.numel())),
                     "feat_loss": ("feat", den(student_local_features.shape[0] * min(student_local_features.shape[-1],
                                                                                      teacher_local_features.shape[-1]))),
                     "grad_loss": ("grad", den(sd.numel()))}
            if use_hdn_loss:
                parts["hdn_loss"] = ("hdn", losses.hdn_loss_dr(sd.detach(), teacher_local_disp, None, hdn_level,
                                                               want_partials=True)[1])
            w = shard_loss_weights(parts)
        out = {k: (v * w[k] if k in w else v) for k, v in out.items() if k != "batch_loss"}
        batch = (lam["sc"] * out["sc_loss"] + lam["lg"] * out["lg_loss"] + lam["feat"] * out["feat_loss"]
                 + lam["grad"] * out["grad_loss"])
        if use_hdn_loss:
            batch = batch + lam["hdn"] * out["hdn_loss"]
        out["batch_loss"] = batch
    batch.backward()
    if data_parallel:
        allreduce_gradients(student_model.parameters())
        with torch.no_grad():   # the weighted local terms sum to the full-batch loss values: one small all-reduce for the report
            names = sorted(out)
            vec = torch.stack([out[n].detach().to(torch.float64) for n in names])
            tdist.all_reduce(vec, op=tdist.ReduceOp.SUM)
            out = {n: vec[i].to(torch.float32) for i, n in enumerate(names)}
    if optimizer is not None:
        if grad_clip is not None:
            torch.nn.utils.clip_grad_norm_(student_model.parameters(), grad_clip)
        optimizer.step()
        optimizer.zero_grad(set_to_none=True)
    return {k: v.detach() for k, v in out.items()}
