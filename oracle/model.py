"""Oracle: functional fp32 restatement of the reference model forward.

Everything takes a flat ``state_dict`` in the *student* key layout
(``pretrained.*`` / ``depth_head.*``; reference
``distillanydepth/depth_anything_v2/dpt.py:187-225``).  The teacher class
(``distillanydepth/modeling/archs/dam/dam.py:307-419``) computes the same
function with ``backbone.blocks.0.N`` keys; use :func:`teacher_to_student_keys`.

TEST INFRASTRUCTURE: see oracle/__init__.py.
"""
import math
import torch
import torch.nn.functional as F

# reference: depth_anything_v2/dinov2.py:339-378 (dims), dpt.py:198-203 (taps)
VIT_CONFIGS = {
    "vits": dict(embed_dim=384, depth=12, num_heads=6, taps=[2, 5, 8, 11]),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12, taps=[2, 5, 8, 11]),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16, taps=[4, 11, 17, 23]),
    "vitg": dict(embed_dim=1536, depth=40, num_heads=24, taps=[9, 19, 29, 39]),  # dinov2.py:381-395, SwiGLU FFN
}
PATCH = 14
LN_EPS = 1e-6  # dinov2.py:95


def teacher_to_student_keys(sd):
    """backbone.blocks.0.N.* -> pretrained.blocks.N.* (SURVEY.md F6)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("backbone.blocks.0."):
            k = "pretrained.blocks." + k[len("backbone.blocks.0."):]
        elif k.startswith("backbone."):
            k = "pretrained." + k[len("backbone."):]
        out[k] = v
    return out


def interpolate_pos_encoding(pos_embed, H, W):
    """dinov2.py:179-210.  ``pos_embed`` is [1, 1+37*37, D]; returns [1, 1+ph*pw, D].

    NB the reference passes (w, h) = (x.shape[2], x.shape[3]) = (H, W) (dinov2.py:213):
    its "w" is the image height.  We keep that order.
    """
    N = pos_embed.shape[1] - 1
    npatch = (H // PATCH) * (W // PATCH)
    if npatch == N and H == W:
        return pos_embed
    pe = pos_embed.float()
    cls_pe, patch_pe = pe[:, 0], pe[:, 1:]
    dim = pe.shape[-1]
    w0, h0 = H // PATCH + 0.1, W // PATCH + 0.1
    sq = math.sqrt(N)
    sx, sy = float(w0) / sq, float(h0) / sq
    patch_pe = F.interpolate(
        patch_pe.reshape(1, int(sq), int(sq), dim).permute(0, 3, 1, 2),
        scale_factor=(sx, sy), mode="bicubic", antialias=False)
    assert int(w0) == patch_pe.shape[-2] and int(h0) == patch_pe.shape[-1]
    patch_pe = patch_pe.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pe.unsqueeze(0), patch_pe), dim=1)


def _attention(x, sd, p, num_heads):
    """dinov2_layers/attention.py:49-62 (naive path; xFormers absent)."""
    B, N, C = x.shape
    hd = C // num_heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def _block(x, sd, p, num_heads):
    """dinov2_layers/block.py:82-107 (eval branch), layer_scale.py:27-28, mlp.py:35-41."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    x = x + _attention(h, sd, p + "attn.", num_heads) * sd[p + "ls1.gamma"]
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    if p + "mlp.w12.weight" in sd:  # SwiGLUFFN.forward, dinov2_layers/swiglu_ffn.py:30-34 (ViT-g, dinov2.py:410)
        x1, x2 = F.linear(h, sd[p + "mlp.w12.weight"], sd[p + "mlp.w12.bias"]).chunk(2, dim=-1)
        h = F.linear(F.silu(x1) * x2, sd[p + "mlp.w3.weight"], sd[p + "mlp.w3.bias"])
    else:
        h = F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
        h = F.gelu(h)  # exact erf (nn.GELU default, dinov2.py:61)
        h = F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h * sd[p + "ls2.gamma"]


def prepare_tokens(x, sd):
    """patch_embed.py:69-82 + dinov2.py:212-219."""
    B, _, H, W = x.shape
    assert H % PATCH == 0 and W % PATCH == 0
    t = F.conv2d(x, sd["pretrained.patch_embed.proj.weight"],
                 sd["pretrained.patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["pretrained.cls_token"].expand(B, -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd["pretrained.pos_embed"], H, W)


def vit_intermediate_layers(x, sd, encoder, return_all=False):
    """dinov2.py:271-281 + :297-321 with norm=True, return_class_token=True.

    Returns a list of 4 (patch_tokens [B,N-1,D], cls [B,D]); with
    ``return_all`` also the raw residual stream after every block.
    """
    cfg = VIT_CONFIGS[encoder]
    t = prepare_tokens(x, sd)
    raw, taps = [t], []
    for i in range(cfg["depth"]):
        t = _block(t, sd, f"pretrained.blocks.{i}.", cfg["num_heads"])
        if return_all:
            raw.append(t)
        if i in cfg["taps"]:
            taps.append(t)
    D = t.shape[-1]
    outs = [F.layer_norm(o, (D,), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], LN_EPS)
            for o in taps]
    res = [(o[:, 1:], o[:, 0]) for o in outs]
    return (res, raw) if return_all else res


def _rcu(x, sd, p):
    """util/blocks.py:67-80: conv2(relu(conv1(relu(x)))) + x (non-inplace ReLU); BatchNorm after each conv when the
    state dict holds ``bn1`` / ``bn2`` (use_bn=True)."""
    def bn(t, q):  # use_bn=True (blocks.py:49-51, 69-75), eval mode: running statistics, eps = 1e-5
        if p + q + ".weight" not in sd:
            return t
        return F.batch_norm(t, sd[p + q + ".running_mean"], sd[p + q + ".running_var"], sd[p + q + ".weight"],
                            sd[p + q + ".bias"], training=False, eps=1e-5)

    out = F.relu(x)
    out = bn(F.conv2d(out, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1), "bn1")
    out = F.relu(out)
    out = bn(F.conv2d(out, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1), "bn2")
    return out + x


def _fusion(sd, p, x0, x1=None, size=None):
    """util/blocks.py:129-146."""
    out = x0
    if x1 is not None:
        out = out + _rcu(x1, sd, p + "resConfUnit1.")
    out = _rcu(out, sd, p + "resConfUnit2.")
    if size is None:
        out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[p + "out_conv.weight"], sd[p + "out_conv.bias"])


def dpt_head_forward(feats, sd, ph, pw, return_intermediates=False):
    """dpt.py:150-184; the use_clstoken readout (:153-156) is applied when the state dict holds
    ``readout_projects``.  Returns the map BEFORE the final
    head ReLUs are all applied the reference way: conv3x3 -> ReLU -> conv1x1 -> ReLU."""
    h = "depth_head."
    outs = []
    for i, (x, cls) in enumerate(feats):
        B, _, D = x.shape
        if h + "readout_projects.0.0.weight" in sd:
            readout = cls.unsqueeze(1).expand_as(x)
            x = F.gelu(F.linear(torch.cat((x, readout), -1), sd[h + f"readout_projects.{i}.0.weight"],
                                sd[h + f"readout_projects.{i}.0.bias"]))
        x = x.permute(0, 2, 1).contiguous().reshape(B, D, ph, pw)
        x = F.conv2d(x, sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"])
        if i == 0:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            x = F.conv_transpose2d(x, sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            x = F.conv2d(x, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        outs.append(x)
    s = h + "scratch."
    l_rn = [F.conv2d(outs[i], sd[s + f"layer{i + 1}_rn.weight"], None, padding=1) for i in range(4)]
    p4 = _fusion(sd, s + "refinenet4.", l_rn[3], None, size=l_rn[2].shape[2:])
    p3 = _fusion(sd, s + "refinenet3.", p4, l_rn[2], size=l_rn[1].shape[2:])
    p2 = _fusion(sd, s + "refinenet2.", p3, l_rn[1], size=l_rn[0].shape[2:])
    p1 = _fusion(sd, s + "refinenet1.", p2, l_rn[0], size=None)
    o1 = F.conv2d(p1, sd[s + "output_conv1.weight"], sd[s + "output_conv1.bias"], padding=1)
    up = F.interpolate(o1, (ph * PATCH, pw * PATCH), mode="bilinear", align_corners=True)
    o2 = F.relu(F.conv2d(up, sd[s + "output_conv2.0.weight"], sd[s + "output_conv2.0.bias"], padding=1))
    pre = F.conv2d(o2, sd[s + "output_conv2.2.weight"], sd[s + "output_conv2.2.bias"])
    if return_intermediates:
        return pre, dict(reassemble=outs, layer_rn=l_rn, paths=[p1, p2, p3, p4], out1=o1)
    return pre


def depth_anything_forward(x, sd, encoder, return_pre_relu=False):
    """DepthAnythingV2.forward (dpt.py:211-225) == DepthAnything.forward
    (dam.py:396-419; the identity-size interpolate at :412 is exact).
    Returns (depth [B,1,H,W], feat [B,N-1,D])."""
    B, _, H, W = x.shape
    ph, pw = H // PATCH, W // PATCH
    feats = vit_intermediate_layers(x, sd, encoder)
    pre = dpt_head_forward(feats, sd, ph, pw)
    depth = F.relu(pre)  # student: ReLU (dpt.py:146) then F.relu (:222); teacher: F.relu (dam.py:415)
    if return_pre_relu:
        return depth, feats[3][0], pre
    return depth, feats[3][0]
