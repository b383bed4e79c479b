"""Drop-in ``DepthAnythingV2`` (reference ``distillanydepth/depth_anything_v2/dpt.py:187-262``).

The class keeps the reference's constructor signature, attribute names, parameter names / shapes
(so ``load_state_dict(strict=True)`` of reference checkpoints and ``state_dict()`` saves keep
working) and the ``(depth, features)`` tuple return, but ``forward`` runs the whole network inside
``libdad_b200.so`` (tcgen05 GEMM / implicit-GEMM conv engine, fused attention, fused output head).
The sub-modules below are parameter containers only; none of them has a forward of its own.

Training: in ``precision = "fp32"`` the outputs are differentiable with respect to the
parameters - ``loss.backward()`` (tools/train_distillation.py:1556-1575) runs ``dad_backward`` behind a
``torch.autograd.Function`` and fills ``.grad`` of every parameter the reference's autograd would reach.  In
``precision = "bf16"`` the forward is inference-only and its outputs are detached.

Options: ``encoder="vitg"`` (SwiGLU FFN, dinov2.py:381-395) and ``use_clstoken=True`` (readout
projection, dpt.py:116-122, 153-156) are part of the native forward AND backward; ``use_bn=True`` (util/blocks.py:49-51)
runs forward-only in ``eval()`` mode, with the running statistics folded into the convolutions when the weights are
synchronised.
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

ENCODERS = {  # dinov2.py:339-378
    "vits": dict(embed_dim=384, depth=12, num_heads=6),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16),
    # vit_giant2 (dinov2.py:381-395) with ffn_layer="swiglufused" (dinov2.py:410): forward-only here
    "vitg": dict(embed_dim=1536, depth=40, num_heads=24, ffn="swiglu"),
}
INTERMEDIATE_LAYER_IDX = {  # dpt.py:198-203
    "vits": [2, 5, 8, 11], "vitb": [2, 5, 8, 11], "vitl": [4, 11, 17, 23], "vitg": [9, 19, 29, 39],
}
_MODES = {"bf16": 0, "fp32": 1}


class _Params(nn.Module):
    """Plain parameter container."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the forward pass runs inside libdad_b200.so")


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class _LayerScale(_Params):
    def __init__(self, dim, init_values):
        super().__init__()
        self.gamma = nn.Parameter(init_values * torch.ones(dim))


def swiglu_hidden(dim, mlp_ratio=4):
    """SwiGLUFFNFused's hidden width (swiglu_ffn.py:55-56): 2/3 of mlp_ratio * dim, rounded up to a multiple of 8."""
    return (int(int(dim * mlp_ratio) * 2 / 3) + 7) // 8 * 8


def _make_block(dim, init_values, ffn="mlp"):
    b = _Params()
    b.norm1 = nn.LayerNorm(dim, eps=1e-6)
    b.attn = _Params()
    b.attn.qkv = nn.Linear(dim, 3 * dim, bias=True)
    b.attn.proj = nn.Linear(dim, dim, bias=True)
    b.ls1 = _LayerScale(dim, init_values)
    b.norm2 = nn.LayerNorm(dim, eps=1e-6)
    b.mlp = _Params()
    if ffn == "swiglu":  # SwiGLUFFN (swiglu_ffn.py:13-34): w3(silu(x1) * x2) with x1, x2 = w12(x).chunk(2)
        hid = swiglu_hidden(dim)
        b.mlp.w12 = nn.Linear(dim, 2 * hid, bias=True)
        b.mlp.w3 = nn.Linear(hid, dim, bias=True)
    else:
        b.mlp.fc1 = nn.Linear(dim, 4 * dim, bias=True)
        b.mlp.fc2 = nn.Linear(4 * dim, dim, bias=True)
    b.ls2 = _LayerScale(dim, init_values)
    return b


class DinoV2Params(_Params):
    """Parameter layout of ``DinoVisionTransformer`` (dinov2.py:44-177) for patch 14, img_size 518."""

    def __init__(self, encoder, init_values=1.0, chunked=False, mask_token=True):
        super().__init__()
        if encoder not in ENCODERS:
            raise KeyError(encoder)  # as model_zoo[model_name] in dinov2.py:406
        cfg = ENCODERS[encoder]
        D = cfg["embed_dim"]
        self.embed_dim = self.num_features = D
        self.n_blocks = cfg["depth"]
        self.num_heads = cfg["num_heads"]
        self.patch_size = 14
        self.num_register_tokens = 0
        self.chunked_blocks = chunked
        self.patch_embed = _Params()
        self.patch_embed.proj = nn.Conv2d(3, D, kernel_size=14, stride=14)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, 37 * 37 + 1, D))
        if mask_token:
            self.mask_token = nn.Parameter(torch.zeros(1, D))
        blocks = [_make_block(D, init_values, cfg.get("ffn", "mlp")) for _ in range(cfg["depth"])]
        # teacher layout (ViT_DINO.py:592): one BlockChunk holding all blocks -> keys blocks.0.N.*
        self.blocks = nn.ModuleList([nn.ModuleList(blocks)]) if chunked else nn.ModuleList(blocks)
        self.norm = nn.LayerNorm(D, eps=1e-6)
        _trunc_normal_(self.pos_embed)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):  # init_weights_vit_timm, dinov2.py:331-336
                _trunc_normal_(m.weight)
                nn.init.zeros_(m.bias)


class DPTHead(_Params):
    """Parameter layout of the reference ``DPTHead`` (dpt.py:71-148, util/blocks.py:4-148)."""

    def __init__(self, in_channels, features=256, use_bn=False, out_channels=(256, 512, 1024, 1024),
                 use_clstoken=False):
        super().__init__()
        oc = list(out_channels)
        self.use_bn = bool(use_bn)
        self.use_clstoken = use_clstoken
        self.projects = nn.ModuleList([nn.Conv2d(in_channels, c, 1) for c in oc])
        self.resize_layers = nn.ModuleList([
            nn.ConvTranspose2d(oc[0], oc[0], kernel_size=4, stride=4),
            nn.ConvTranspose2d(oc[1], oc[1], kernel_size=2, stride=2),
            nn.Identity(),
            nn.Conv2d(oc[3], oc[3], kernel_size=3, stride=2, padding=1)])
        if use_clstoken:  # dpt.py:116-122; forward-only here
            self.readout_projects = nn.ModuleList(
                [nn.Sequential(nn.Linear(2 * in_channels, in_channels), nn.GELU()) for _ in oc])
        sc = _Params()
        for i in range(4):
            setattr(sc, f"layer{i + 1}_rn", nn.Conv2d(oc[i], features, 3, padding=1, bias=False))
        sc.stem_transpose = None
        for r in (1, 2, 3, 4):
            fb = _Params()
            fb.out_conv = nn.Conv2d(features, features, 1)
            for u in (1, 2):
                rcu = _Params()
                rcu.conv1 = nn.Conv2d(features, features, 3, padding=1)
                rcu.conv2 = nn.Conv2d(features, features, 3, padding=1)
                if use_bn:  # blocks.py:49-51: BatchNorm after each conv; eval-mode statistics are folded into the conv
                    rcu.bn1 = nn.BatchNorm2d(features)
                    rcu.bn2 = nn.BatchNorm2d(features)
                setattr(fb, f"resConfUnit{u}", rcu)
            setattr(sc, f"refinenet{r}", fb)
        sc.output_conv1 = nn.Conv2d(features, features // 2, 3, padding=1)
        sc.output_conv2 = nn.Sequential(
            nn.Conv2d(features // 2, 32, 3, padding=1), nn.ReLU(True), nn.Conv2d(32, 1, 1), nn.ReLU(True),
            nn.Identity())
        self.scratch = sc


class _NativeDepthModel(nn.Module):
    """Shared machinery: owns the ``dad_model`` handle, syncs parameters into it, runs forward."""

    _encoder_attr = "pretrained"  # attribute holding the ViT ("backbone" for the teacher class)

    def _init_native(self, encoder, features, out_channels):
        self._desc = dict(encoder=encoder, features=int(features), out_channels=[int(c) for c in out_channels])
        self._handle = None
        self._sig = None
        self._ws = None
        self._prepared = set()
        self.precision = "bf16"  # "bf16" (tensor cores) or "fp32" (verification mode)
        # precision "bf16": False = inference forward (outputs detached); True = differentiable forward with a bf16
        # activation tape and the tensor-core backward.  precision "fp32" is always differentiable under grad mode.
        self.bf16_backward = False

    # -- key mapping: our state-dict key -> student-layout key understood by the library
    def _student_key(self, k):
        return k

    def _ensure_handle(self):
        if self._handle is not None:
            return
        lib = _lib.load()
        cfg = ENCODERS[self._desc["encoder"]]
        d = _lib.ModelDesc()
        d.embed_dim, d.depth, d.num_heads = cfg["embed_dim"], cfg["depth"], cfg["num_heads"]
        d.taps = (ctypes.c_int * 4)(*INTERMEDIATE_LAYER_IDX[self._desc["encoder"]])
        d.features = self._desc["features"]
        d.out_channels = (ctypes.c_int * 4)(*self._desc["out_channels"])
        h = ctypes.c_void_p()
        _lib.check(lib.dad_model_create(ctypes.byref(d), ctypes.byref(h)), "dad_model_create")
        self._handle = h

    def _sync_weights(self, device):
        params = list(self.named_parameters())
        use_bn = self.depth_head.use_bn
        buffers = list(self.named_buffers()) if use_bn else []   # BatchNorm running statistics
        sig = tuple((p.data_ptr(), p._version) for _, p in params + buffers)
        if sig == self._sig:
            return
        # upload only what changed since the last call (an optimiser step changes every trained parameter; a manual edit
        # of one tensor re-uploads one tensor); any BatchNorm change re-folds the affected convolutions
        prev = self._sig_by_key if getattr(self, "_sig_by_key", None) is not None else {}
        now = {k: (p.data_ptr(), p._version) for k, p in params + buffers}
        bn_changed = use_bn and any(prev.get(k) != now[k] for k, _ in buffers)
        lib = _lib.load()
        st = _lib.stream_ptr()
        folded = self._fold_batchnorm() if use_bn else {}
        for k, p in params:
            if p.device != device:
                raise RuntimeError(f"parameter {k} is on {p.device}, input on {device}: call model.to(device)")
            if use_bn and (".bn1." in k or ".bn2." in k):
                bn_changed = bn_changed or prev.get(k) != now[k]
                continue   # lives on inside the folded convolution
        for k, p in params:
            if use_bn and (".bn1." in k or ".bn2." in k):
                continue
            if prev.get(k) == now[k] and not (bn_changed and k in folded):
                continue
            t = folded.get(k, p).detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            _lib.check(lib.dad_model_set_weight(self._handle, self._student_key(k).encode(), _lib.ptr(t), t.numel(), st),
                       f"set_weight({k})")
        torch.cuda.current_stream(device).synchronize()  # temporaries above may be freed after this
        self._sig = sig
        self._sig_by_key = now
        self._prepared = set()

    @torch.no_grad()
    def _fold_batchnorm(self):
        """use_bn=True (util/blocks.py:49-51, 67-75), eval mode: ``bn(conv(x)) = conv'(x)`` with
        ``W' = W * a[co]``, ``b' = (b - running_mean) * a + beta``, ``a = gamma / sqrt(running_var + eps)``.
        Returns {conv parameter key: folded fp32 tensor}; the library never sees the BatchNorm layers."""
        out = {}
        for name, rcu in self.depth_head.named_modules():
            if not hasattr(rcu, "bn1"):
                continue
            for conv, bn, c in ((rcu.conv1, rcu.bn1, "conv1"), (rcu.conv2, rcu.bn2, "conv2")):
                a = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
                out[f"depth_head.{name}.{c}.weight"] = (conv.weight.double() * a[:, None, None, None]).float()
                out[f"depth_head.{name}.{c}.bias"] = ((conv.bias.double() - bn.running_mean.double()) * a
                                                      + bn.bias.double()).float()
        return out

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.load().dad_model_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # parameters the reference forward never touches (autograd leaves their .grad = None): the mask token
    # (del_mask_token / masks=None, dinov2.py:212-218) and refinenet4's first RCU (no skip input, blocks.py:137-141)
    _UNUSED = ("mask_token", "refinenet4.resConfUnit1.")

    def _run(self, x, captures=None):
        if not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected an image batch of shape [B, 3, H, W]")
        if not x.is_cuda:
            raise RuntimeError("the B200 forward path runs on CUDA tensors only (no CPU fallback)")
        B, _, H, W = x.shape
        # same failure mode as PatchEmbed.forward (patch_embed.py:73-74)
        assert H % 14 == 0, f"Input image height {H} is not a multiple of patch height 14"
        assert W % 14 == 0, f"Input image width {W} is not a multiple of patch width: 14"
        if self.precision not in _MODES:
            raise ValueError("precision must be 'bf16' or 'fp32'")
        mode = _MODES[self.precision]
        if self.depth_head.use_bn and self.training:
            raise NotImplementedError("use_bn=True runs with the running statistics only (call model.eval()): batch-statistics "
                                      "BatchNorm is outside the hot path (no caller enables use_bn)")
        if (mode == 1 or self.bf16_backward) and captures is None and torch.is_grad_enabled():
            live = [(k, p) for k, p in self.named_parameters()
                    if p.requires_grad and not any(u in k for u in self._UNUSED)]
            if live and self.depth_head.use_bn:
                raise NotImplementedError("use_bn=True is forward-only here (folded running statistics): run under "
                                          "torch.no_grad(); the training backward covers the graph without BatchNorm")
            if live:
                return _TrainForward.apply(self, x, mode, tuple(k for k, _ in live), *[p for _, p in live])
        return self._run_native(x, mode, captures)

    def _train_begin(self, x, mode):
        """dad_forward_train: same outputs as the inference forward, activations kept on a tape tensor."""
        B, _, H, W = x.shape
        lib = _lib.load()
        with torch.cuda.device(x.device):
            self._ensure_handle()
            self._sync_weights(x.device)
            st = _lib.stream_ptr()
            for md in {mode, 1}:   # the bf16 backward still runs a few small contractions on the fp32 pack
                if (md, H, W) not in self._prepared:
                    _lib.check(lib.dad_model_prepare(self._handle, md, H, W, st), "dad_model_prepare")
                    self._prepared.add((md, H, W))
            need = int(lib.dad_train_workspace_bytes(self._handle, B, H, W, mode))
            if need == 0:
                _lib.check(-1, "dad_train_workspace_bytes")
            tape = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
            off = (-tape.data_ptr()) % 1024
            xin = x.detach()
            if xin.dtype != torch.float32 or not xin.is_contiguous():
                xin = xin.float().contiguous()
            D = ENCODERS[self._desc["encoder"]]["embed_dim"]
            depth = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
            feat = torch.empty(B, (H // 14) * (W // 14), D, dtype=torch.float32, device=x.device)
            _lib.check(lib.dad_forward_train(self._handle, _lib.ptr(xin), B, H, W, mode, _lib.ptr(depth), _lib.ptr(feat),
                                             ctypes.c_void_p(tape.data_ptr() + off), tape.numel() - off, st),
                       "dad_forward_train")
        return depth, feat, tape

    def _train_backward(self, tape, shape, mode, names, gdepth, gfeat):
        """dad_backward: returns one fp32 gradient tensor per name (accumulated by the library from zero)."""
        B, H, W = shape
        lib = _lib.load()
        own = dict(self.named_parameters())
        grads = []
        with torch.cuda.device(tape.device):
            st = _lib.stream_ptr()
            registered = []
            try:
                for k in names:
                    g = torch.zeros(own[k].shape, dtype=torch.float32, device=tape.device)
                    grads.append(g)
                    sk = self._student_key(k).encode()
                    _lib.check(lib.dad_model_set_grad(self._handle, sk, _lib.ptr(g), g.numel()), f"set_grad({k})")
                    registered.append(sk)
                off = (-tape.data_ptr()) % 1024
                gd = gdepth.float().contiguous()
                gf = None if gfeat is None else gfeat.float().contiguous()
                _lib.check(lib.dad_backward(self._handle, B, H, W, mode, _lib.ptr(gd), _lib.ptr(gf),
                                            ctypes.c_void_p(tape.data_ptr() + off), tape.numel() - off, st), "dad_backward")
            finally:
                for sk in registered:
                    lib.dad_model_set_grad(self._handle, sk, None, 0)
        return grads

    def _run_native(self, x, mode, captures=None):
        B, _, H, W = x.shape
        lib = _lib.load()
        with torch.cuda.device(x.device):
            self._ensure_handle()
            self._sync_weights(x.device)
            st = _lib.stream_ptr()
            if (mode, H, W) not in self._prepared:
                _lib.check(lib.dad_model_prepare(self._handle, mode, H, W, st), "dad_model_prepare")
                self._prepared.add((mode, H, W))
            need = int(lib.dad_forward_workspace_bytes(self._handle, B, H, W, mode))
            need += int(os.environ.get("DAD_WS_EXTRA_MB", "0")) << 20  # debugging aid
            if need == 0:
                _lib.check(-1, "dad_forward_workspace_bytes")
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = None
                self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
            off = (-self._ws.data_ptr()) % 1024
            xin = x.detach()
            if xin.dtype != torch.float32 or not xin.is_contiguous():
                xin = xin.float().contiguous()
            D = ENCODERS[self._desc["encoder"]]["embed_dim"]
            depth = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
            feat = torch.empty(B, (H // 14) * (W // 14), D, dtype=torch.float32, device=x.device)
            for name, buf in (captures or {}).items():
                _lib.check(lib.dad_model_debug_capture(self._handle, name.encode(), _lib.ptr(buf), buf.numel()))
            try:
                _lib.check(lib.dad_forward(self._handle, _lib.ptr(xin), B, H, W, mode, _lib.ptr(depth), _lib.ptr(feat),
                                           ctypes.c_void_p(self._ws.data_ptr() + off), self._ws.numel() - off, st),
                           "dad_forward")
            finally:
                for name in (captures or {}):
                    lib.dad_model_debug_capture(self._handle, name.encode(), None, 0)
        return depth, feat


class _TrainForward(torch.autograd.Function):
    """Differentiable forward: forward = dad_forward_train (activation tape), backward = dad_backward."""

    @staticmethod
    def forward(ctx, model, x, mode, names, *params):
        depth, feat, tape = model._train_begin(x, mode)
        ctx.model, ctx.tape, ctx.names, ctx.mode = model, tape, names, mode
        ctx.shape = (x.shape[0], x.shape[2], x.shape[3])
        ctx.versions = tuple(p._version for p in params)
        ctx.params = params
        return depth, feat

    @staticmethod
    def backward(ctx, gdepth, gfeat):
        if ctx.tape is None:
            raise RuntimeError("backward through the native forward a second time: the activation tape was freed")
        if tuple(p._version for p in ctx.params) != ctx.versions:
            raise RuntimeError("a parameter was modified between the native forward and its backward")
        if gdepth is None:
            gdepth = torch.zeros(ctx.shape[0], 1, ctx.shape[1], ctx.shape[2], device=ctx.tape.device)
        grads = ctx.model._train_backward(ctx.tape, ctx.shape, ctx.mode, ctx.names, gdepth, gfeat)
        ctx.tape = None
        return (None, None, None, None, *grads)


class DepthAnythingV2(_NativeDepthModel):
    def __init__(self, encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024], use_bn=False,
                 use_clstoken=False):
        super().__init__()
        self.intermediate_layer_idx = dict(INTERMEDIATE_LAYER_IDX)
        self.encoder = encoder
        self.pretrained = DinoV2Params(encoder, init_values=1.0, chunked=False)  # DINOv2(), dinov2.py:398-415
        self.depth_head = DPTHead(self.pretrained.embed_dim, features, use_bn, out_channels=out_channels,
                                  use_clstoken=use_clstoken)
        self._init_native(encoder, features, out_channels)

    def forward(self, x):
        """-> (depth [B,1,H,W], features[3][0] = last-tap patch tokens [B,(H/14)(W/14),D])  (dpt.py:211-225)"""
        return self._run(x)

    @torch.no_grad()
    def infer_image(self, raw_image, input_size=518):
        """dpt.py:227-235 with its evident intent (upstream indexes the tuple and raises):
        BGR uint8 image -> relative depth ``np.ndarray [h, w]`` at the raw resolution.  Pre- and post-processing
        run on the GPU (preprocess.py): the only host<->device traffic is the uint8 image and the final map."""
        from . import preprocess
        image, (h, w) = self.image2tensor(raw_image, input_size)
        depth, _ = self.forward(image)
        depth = preprocess.resize_depth(depth, (h, w))[0, 0]
        return depth.cpu().numpy()

    def image2tensor(self, raw_image, input_size=518):
        """dpt.py:237-262: keep-aspect lower-bound resize to a multiple of 14 (cv2 INTER_CUBIC on the float64
        image / 255), ImageNet normalisation, HWC->CHW; computed by one CUDA kernel from the uint8 image."""
        from . import preprocess
        device = next(self.parameters()).device
        return preprocess.image_to_tensor(raw_image, input_size, device=device, bgr=True)
