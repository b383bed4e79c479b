"""Deterministic synthetic weights and inputs.

numpy PCG64 streams so the same seed gives the same tensors on any box,
independently of the torch build.  The key layout and shapes are the
reference's (student layout, ``distillanydepth/depth_anything_v2/dpt.py:187-209``;
239 tensors for vitb).  LayerScale gammas are O(1) as in trained checkpoints
 and the last bias is positive so the depth map is live.
"""
import math
import numpy as np
import torch

ENCODERS = {
    "vits": dict(embed_dim=384, depth=12, num_heads=6, taps=[2, 5, 8, 11]),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12, taps=[2, 5, 8, 11]),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16, taps=[4, 11, 17, 23]),
    "vitg": dict(embed_dim=1536, depth=40, num_heads=24, taps=[9, 19, 29, 39], ffn_hidden=4096),  # SwiGLU, dinov2.py:381-395
}
MODEL_PRESETS = {  # tools/train_distillation.py:713-730, :802-808; BASELINE.json configs
    "vits": dict(encoder="vits", features=64, out_channels=[48, 96, 192, 384]),
    "vitb": dict(encoder="vitb", features=128, out_channels=[96, 192, 384, 768]),
    "vitl": dict(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]),
    "vitg": dict(encoder="vitg", features=384, out_channels=[1536, 1536, 1536, 1536]),  # Depth-Anything-V2 giant config
}


def param_shapes(encoder, features, out_channels, use_clstoken=False, use_bn=False):
    """Ordered {key: shape} of the student state dict (``use_clstoken`` appends the readout projections, dpt.py:116-122;
    ``use_bn`` the BatchNorm parameters and buffers of the residual conv units, util/blocks.py:49-51)."""
    cfg = ENCODERS[encoder]
    D, L = cfg["embed_dim"], cfg["depth"]
    s = {}
    p = "pretrained."
    s[p + "cls_token"] = (1, 1, D)
    s[p + "pos_embed"] = (1, 1370, D)
    s[p + "mask_token"] = (1, D)
    s[p + "patch_embed.proj.weight"] = (D, 3, 14, 14)
    s[p + "patch_embed.proj.bias"] = (D,)
    for i in range(L):
        b = f"{p}blocks.{i}."
        s[b + "norm1.weight"] = (D,); s[b + "norm1.bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D); s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D); s[b + "attn.proj.bias"] = (D,)
        s[b + "ls1.gamma"] = (D,)
        s[b + "norm2.weight"] = (D,); s[b + "norm2.bias"] = (D,)
        if "ffn_hidden" in cfg:  # SwiGLUFFNFused (swiglu_ffn.py:44-63)
            Hd = cfg["ffn_hidden"]
            s[b + "mlp.w12.weight"] = (2 * Hd, D); s[b + "mlp.w12.bias"] = (2 * Hd,)
            s[b + "mlp.w3.weight"] = (D, Hd); s[b + "mlp.w3.bias"] = (D,)
        else:
            s[b + "mlp.fc1.weight"] = (4 * D, D); s[b + "mlp.fc1.bias"] = (4 * D,)
            s[b + "mlp.fc2.weight"] = (D, 4 * D); s[b + "mlp.fc2.bias"] = (D,)
        s[b + "ls2.gamma"] = (D,)
    s[p + "norm.weight"] = (D,); s[p + "norm.bias"] = (D,)
    h = "depth_head."
    oc, Fe = list(out_channels), features
    for i in range(4):
        s[h + f"projects.{i}.weight"] = (oc[i], D, 1, 1); s[h + f"projects.{i}.bias"] = (oc[i],)
    s[h + "resize_layers.0.weight"] = (oc[0], oc[0], 4, 4); s[h + "resize_layers.0.bias"] = (oc[0],)
    s[h + "resize_layers.1.weight"] = (oc[1], oc[1], 2, 2); s[h + "resize_layers.1.bias"] = (oc[1],)
    s[h + "resize_layers.3.weight"] = (oc[3], oc[3], 3, 3); s[h + "resize_layers.3.bias"] = (oc[3],)
    sc = h + "scratch."
    for i in range(4):
        s[sc + f"layer{i + 1}_rn.weight"] = (Fe, oc[i], 3, 3)
    for r in (1, 2, 3, 4):
        q = sc + f"refinenet{r}."
        s[q + "out_conv.weight"] = (Fe, Fe, 1, 1); s[q + "out_conv.bias"] = (Fe,)
        for u in (1, 2):
            for c in (1, 2):
                s[q + f"resConfUnit{u}.conv{c}.weight"] = (Fe, Fe, 3, 3)
                s[q + f"resConfUnit{u}.conv{c}.bias"] = (Fe,)
    s[sc + "output_conv1.weight"] = (Fe // 2, Fe, 3, 3); s[sc + "output_conv1.bias"] = (Fe // 2,)
    s[sc + "output_conv2.0.weight"] = (32, Fe // 2, 3, 3); s[sc + "output_conv2.0.bias"] = (32,)
    s[sc + "output_conv2.2.weight"] = (1, 32, 1, 1); s[sc + "output_conv2.2.bias"] = (1,)
    if use_bn:
        for r in (1, 2, 3, 4):
            for u in (1, 2):
                for c in (1, 2):
                    q = sc + f"refinenet{r}.resConfUnit{u}.bn{c}."
                    for leaf in ("weight", "bias", "running_mean", "running_var"):
                        s[q + leaf] = (Fe,)
                    s[q + "num_batches_tracked"] = ()
    if use_clstoken:  # appended last so the other tensors keep the values they have without the option
        for i in range(4):
            s[h + f"readout_projects.{i}.0.weight"] = (D, 2 * D); s[h + f"readout_projects.{i}.0.bias"] = (D,)
    return s


def make_state_dict(encoder="vits", features=64, out_channels=(48, 96, 192, 384), seed=0,
                    head_bias=0.25, head_gain=6.0, use_clstoken=False, use_bn=False):
    """Random-init weights: Linear N(0, 0.02) (dinov2.py:331-336), conv
    U(+-1/sqrt(fan_in)) (PyTorch default), small random biases so every bias path
    is exercised, LayerNorm/LayerScale around 1, final bias ``head_bias`` > 0."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for k, shp in param_shapes(encoder, features, out_channels, use_clstoken, use_bn).items():
        n = int(np.prod(shp))
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[k] = torch.tensor(100, dtype=torch.int64)
            continue
        if ".bn" in k:  # BatchNorm of a trained net: scale near 1, shifts / running means of the activations' order
            v = {"weight": lambda: 1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32),
                 "bias": lambda: 0.05 * rng.standard_normal(n, dtype=np.float32),
                 "running_mean": lambda: 0.1 * rng.standard_normal(n, dtype=np.float32),
                 "running_var": lambda: rng.uniform(0.5, 1.5, n).astype(np.float32)}[leaf]()
        elif k.endswith("gamma"):
            v = 1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)
        elif ".norm" in k and leaf == "weight":
            v = 1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)
        elif ".norm" in k and leaf == "bias":
            v = 0.05 * rng.standard_normal(n, dtype=np.float32)
        elif k.startswith("pretrained.") and leaf == "weight" and len(shp) == 2:
            v = 0.02 * rng.standard_normal(n, dtype=np.float32)
        elif leaf in ("cls_token", "pos_embed", "mask_token"):
            v = 0.02 * rng.standard_normal(n, dtype=np.float32)
        elif leaf == "weight" and len(shp) == 2:  # head Linear layers (readout): PyTorch default U(+-1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(shp[1])
            v = rng.uniform(-bound, bound, n).astype(np.float32)
        elif leaf == "weight":  # convs
            fan_in = shp[1] * shp[2] * shp[3]
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            v = rng.uniform(-bound, bound, n).astype(np.float32)
        else:  # biases
            v = 0.02 * rng.standard_normal(n, dtype=np.float32)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v.reshape(shp)))
    sd["depth_head.scratch.output_conv2.2.bias"] = torch.full((1,), float(head_bias))
    sd["depth_head.scratch.output_conv2.2.weight"] *= float(head_gain)  # signal std ~0.04 around the bias
    return sd


def make_images(B, H, W, seed=1234, scale=1.0):
    """x ~ N(0,1) [B,3,H,W] (ImageNet-normalised statistics); scale=255 with
    ``uniform`` gives the raw-range case of SURVEY.md A.10."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((B, 3, H, W), dtype=np.float32) * scale)


def make_depth_pair(B, H, W, seed=7):
    """pred, gt ~ U[0,1) and mask = rand > 0.5 (demo.py:165-167)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pred = torch.from_numpy(rng.random((B, 1, H, W), dtype=np.float32))
    gt = torch.from_numpy(rng.random((B, 1, H, W), dtype=np.float32))
    mask = torch.from_numpy(rng.random((B, 1, H, W), dtype=np.float32) > 0.5)
    return pred, gt, mask


def make_features(B, N, D, seed=11):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((B, N, D), dtype=np.float32))
