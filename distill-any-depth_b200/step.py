"""Forward half of the reference's distillation training step (``tools/train_distillation.py:1503-1560``):
student forward on the global and the local image, teacher forward on the local image, then the five
loss terms and their lambda-weighted sum.  Everything runs on the caller's CUDA stream through the C ABI;
nothing synchronises the host.  Under ``torch.distributed`` (one process per GPU, images sharded by rank)
each loss is finished from its (numerator, denominator) partials by ONE all-reduce (``dist.finish_losses``),
so the value equals the single-process full-batch loss.

The backward / optimiser half (``:1561-1573``) is outside this path.
"""
import torch

from . import losses
from .dist import finish_losses

# scripts/train_test.sh:21-25
DEFAULT_LAMBDAS = dict(sc=0.5, lg=0.5, feat=1.0, grad=0.2, hdn=0.8)


def distillation_step_losses(student_model, teacher_model, global_image, local_image, normalization="hybrid",
                             lambdas=None, use_hdn_loss=True, hdn_level=3, hdn_variant="dr", dedup_student=False):
    """Returns ``dict(sc_loss, lg_loss, feat_loss, grad_loss, hdn_loss, batch_loss)`` of fp32 device scalars.

    ``dedup_student``: the reference runs the student twice, on ``global_image`` and ``local_image``
    (``:1509-1510``) even when both are the same tensor (NYU path, ``:1480-1484``); with ``True`` and identical
    inputs the second, bit-identical forward is elided.
    """
    lam = dict(DEFAULT_LAMBDAS)
    lam.update(lambdas or {})
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
        out = finish_losses(parts)  # one all-reduce of the stacked partials when torch.distributed is initialised
        if not use_hdn_loss:
            out["hdn_loss"] = torch.zeros((), device=local_image.device)
        batch = (lam["sc"] * out["sc_loss"] + lam["lg"] * out["lg_loss"] + lam["feat"] * out["feat_loss"]
                 + lam["grad"] * out["grad_loss"])
        if use_hdn_loss:
            batch = batch + lam["hdn"] * out["hdn_loss"]
        out["batch_loss"] = batch
    return out
