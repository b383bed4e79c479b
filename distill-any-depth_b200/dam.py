"""Drop-in teacher class ``DepthAnything`` (reference
``distillanydepth/modeling/archs/dam/dam.py:307-419``): the same network as ``DepthAnythingV2``
 under the ``backbone.*`` / ``blocks.0.N`` key layout, LayerScale initialised to 1e-5
(``ViT_DINO.py:587``)."""
from .dpt import DinoV2Params, DPTHead, INTERMEDIATE_LAYER_IDX, _NativeDepthModel


def teacher_to_student_keys(sd):
    """``backbone.blocks.0.N.*`` / ``backbone.*`` (teacher, ViT_DINO.py:592) -> ``pretrained.blocks.N.*`` /
    ``pretrained.*`` (student, dpt.py:196) for every key of a state dict; other keys pass through."""
    out = {}
    for k, v in sd.items():
        if k.startswith("backbone.blocks.0."):
            k = "pretrained.blocks." + k[len("backbone.blocks.0."):]
        elif k.startswith("backbone."):
            k = "pretrained." + k[len("backbone."):]
        out[k] = v
    return out


def student_to_teacher_keys(sd):
    """Inverse of :func:`teacher_to_student_keys` (what ``tools/convert_checkpoint.py:7-28`` does to a DAv2 file)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("pretrained.blocks."):
            k = "backbone.blocks.0." + k[len("pretrained.blocks."):]
        elif k.startswith("pretrained."):
            k = "backbone." + k[len("pretrained."):]
        out[k] = v
    return out


class DepthAnything(_NativeDepthModel):
    _encoder_attr = "backbone"

    def __init__(self, encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024], head_out_channels=1,
                 wo_relu_1_2_channel=False, use_bn=False, use_clstoken=False, use_registers=False, max_depth=1.0,
                 mode="disparity", num_depth_regressor_anchor=512, depth_normalize=(0.1, 150),
                 pretrain_type="dinov2", del_mask_token=True):
        super().__init__()
        assert encoder in ["vits", "vitb", "vitl", "vitg"]
        if use_registers or pretrain_type != "dinov2" or head_out_channels != 1 or wo_relu_1_2_channel:
            raise NotImplementedError("registers / non-dinov2 backbones / multi-channel heads are outside the hot "
                                      "path (no caller enables them, SURVEY.md 2.1)")
        if encoder != "vitl":
            # dam.py:361-365: 'vitb' selects the windowed ViT (unreachable from any caller), others raise
            raise NotImplementedError(f"DepthAnything(encoder={encoder!r}) is outside the hot path; "
                                      "use DepthAnythingV2 for vits / vitb")
        self.pretrain_type, self.mode = pretrain_type, mode
        self.min_depth, self.max_depth = depth_normalize
        self.num_depth_regressor_anchor = num_depth_regressor_anchor
        self.wo_relu_1_2_channel = wo_relu_1_2_channel
        self.intermediate_layer_idx = dict(INTERMEDIATE_LAYER_IDX)
        self.backbone_name = encoder
        self.backbone = DinoV2Params(encoder, init_values=1e-5, chunked=True, mask_token=True)
        self.depth_head = DPTHead(self.backbone.embed_dim, features, use_bn, out_channels=out_channels,
                                  use_clstoken=use_clstoken)
        self._init_native(encoder, features, out_channels)

    def _student_key(self, k):
        return next(iter(teacher_to_student_keys({k: None})))

    def forward(self, x):
        """-> (depth [B,1,H,W], features[3][0])  (dam.py:396-419; the identity-size interpolate at :412 is exact)."""
        return self._run(x)
