"""B200-native (sm_100a) drop-in for the Distill-Any-Depth hot path: DepthAnythingV2 / DepthAnything
forward and the distillation losses, behind the reference's Python signatures, computed by
hand-written CUDA in lib/libdad_b200.so (C ABI: include/dad_b200.h)."""
from . import synthetic  # noqa: F401
from .dpt import DepthAnythingV2, DPTHead  # noqa: F401
from .dam import DepthAnything  # noqa: F401
from .losses import (masked_shift_and_scale, masked_l1_loss, SSILoss, get_contexts_dr, get_contexts_dp,  # noqa: F401
                     get_contexts_ds, compute_hdn_loss, hdn_loss_dr, ssi_hdn_dr, gradient_preservation_loss,
                     feature_distillation_loss, distillation_loss, global_normalize, hybrid_normalize,
                     local_normalize, normalize_depth)
from . import dist  # noqa: F401
from . import checkpoint  # noqa: F401
from . import preprocess  # noqa: F401
from .graph import capture  # noqa: F401
from .step import distillation_step_losses, distillation_train_step  # noqa: F401
