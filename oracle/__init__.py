"""CPU oracle for the Distill-Any-Depth hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain PyTorch-CPU fp32 restatement of the reference's
DepthAnythingV2 / DepthAnything forward and of its distillation losses.  It is
the checker the CUDA path is compared against; it is never the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under
``distill-any-depth_b200/`` imports it.

Pinning: the reference ships no golden vectors for this path.
The restatement is pinned instead against the *live* reference imported from
``/root/reference`` in the build container (``oracle/make_golden.py``), whose
outputs on seeded inputs are committed under ``tests/golden/`` together with
the generating script; ``tests/test_oracle_golden.py`` replays them anywhere.
"""
from .model import (VIT_CONFIGS, vit_intermediate_layers, dpt_head_forward,
                    depth_anything_forward, interpolate_pos_encoding)
from .losses import (masked_shift_and_scale, masked_l1_loss, SSILoss,
                     get_contexts_dr, get_contexts_dp, get_contexts_ds,
                     compute_hdn_loss, gradient_preservation_loss,
                     feature_distillation_loss, distillation_loss,
                     global_normalize, hybrid_normalize, normalize_depth)
