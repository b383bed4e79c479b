"""CPU: host-side mirror of the reference interface (parameter layout, option handling, errors) and the
data-parallel loss finishing (world_size-2 gloo) - no GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from distill_any_depth_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_student_state_dict_layout_matches_reference_shapes():
    import distill_any_depth_b200 as d
    for preset in ("vits", "vitb"):
        kw = synthetic.MODEL_PRESETS[preset]
        m = d.DepthAnythingV2(**kw)
        want = synthetic.param_shapes(**kw)
        got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert got == {k: tuple(s) for k, s in want.items()}
        assert m.load_state_dict(synthetic.make_state_dict(seed=0, **kw), strict=True)
    assert len(d.DepthAnythingV2(**synthetic.MODEL_PRESETS["vitb"]).state_dict()) == 239
    m = d.DepthAnythingV2(**synthetic.MODEL_PRESETS["vits"])
    assert m.encoder == "vits" and m.pretrained.embed_dim == 384 and m.pretrained.n_blocks == 12
    assert m.pretrained.num_heads == 6 and m.pretrained.patch_size == 14
    assert m.intermediate_layer_idx["vitl"] == [4, 11, 17, 23]


def test_teacher_key_layout_and_mapping():
    import distill_any_depth_b200 as d
    from oracle.make_golden import student_to_teacher_keys
    kw = synthetic.MODEL_PRESETS["vitl"]
    t = d.DepthAnything(**kw)
    keys = list(t.state_dict())
    assert len(keys) == 407 and "backbone.blocks.0.23.mlp.fc2.weight" in keys and "backbone.mask_token" in keys
    want = student_to_teacher_keys({k: None for k in synthetic.param_shapes(**kw)})
    assert set(keys) == set(want)
    assert all(t._student_key(k) in synthetic.param_shapes(**kw) for k in keys)
    assert float(t.backbone.blocks[0][0].ls1.gamma.detach()[0]) == pytest.approx(1e-5)  # ViT_DINO.py:587


def test_options_outside_the_hot_path_raise():
    import distill_any_depth_b200 as d
    kw = synthetic.MODEL_PRESETS["vits"]
    with pytest.raises(KeyError):
        d.DepthAnythingV2(encoder="vitx")
    with pytest.raises(NotImplementedError):
        d.DepthAnything(encoder="vitb")
    with pytest.raises(NotImplementedError):
        d.DepthAnything(use_registers=True)
    with pytest.raises(RuntimeError):  # CPU tensor: no fallback
        d.get_contexts_ds(3, torch.ones(1, 1, 4, 4, dtype=torch.bool))


def test_option_key_layouts_match_the_reference():
    """use_clstoken adds depth_head.readout_projects.{i}.0.{weight,bias} (dpt.py:116-122); vitg swaps the
    Mlp for SwiGLUFFNFused (mlp.w12 / mlp.w3, hidden 4096 = 2/3 * 4 * 1536 rounded up to 8; swiglu_ffn.py:44-63)."""
    import distill_any_depth_b200 as d
    from distill_any_depth_b200.dpt import swiglu_hidden
    kw = synthetic.MODEL_PRESETS["vits"]
    m = d.DepthAnythingV2(use_clstoken=True, **kw)
    want = synthetic.param_shapes(use_clstoken=True, **kw)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want and got["depth_head.readout_projects.3.0.weight"] == (384, 768)
    assert m.depth_head.use_clstoken is True
    assert swiglu_hidden(1536) == 4096 and swiglu_hidden(384) == 1024
    with torch.device("meta"):  # 1.26e9 parameters: shapes only
        g = d.DepthAnythingV2(**synthetic.MODEL_PRESETS["vitg"])
    got = {k: tuple(v.shape) for k, v in g.state_dict().items()}
    assert got == synthetic.param_shapes(**synthetic.MODEL_PRESETS["vitg"])
    assert got["pretrained.blocks.39.mlp.w12.weight"] == (8192, 1536) and "pretrained.blocks.0.mlp.fc1.weight" not in got
    assert g.pretrained.n_blocks == 40 and g.pretrained.num_heads == 24 and g.intermediate_layer_idx["vitg"] == [9, 19, 29, 39]


def test_use_bn_folding_reproduces_the_batchnorm_graph():
    """use_bn=True (util/blocks.py:49-51): the front-end folds the eval-mode BatchNorm into the conv it follows before the
    weights reach the library.  Host arithmetic only, so it is checked here: the oracle on the folded state dict (no bn
    keys) equals the oracle on the BatchNorm graph, which is pinned to the live reference (golden vits_bn_70x98)."""
    import distill_any_depth_b200 as d
    import oracle
    kw = dict(synthetic.MODEL_PRESETS["vits"], use_bn=True)
    sd = synthetic.make_state_dict(seed=8, **kw)
    m = d.DepthAnythingV2(**kw)
    m.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == synthetic.param_shapes(**kw)
    folded = m._fold_batchnorm()
    assert len(folded) == 4 * 2 * 2 * 2  # 4 fusion blocks x 2 units x 2 convs x (weight, bias)
    plain = {k: folded.get(k, v) for k, v in sd.items() if ".bn1." not in k and ".bn2." not in k}
    x = synthetic.make_images(1, 70, 98, seed=1240)
    with torch.no_grad():
        want, _ = oracle.depth_anything_forward(x, sd, "vits")
        got, _ = oracle.depth_anything_forward(x, plain, "vits")
    assert (got - want).abs().max().item() <= 2e-6 * want.abs().max().item()
    # batch-statistics mode is not offered: the check runs before any device work
    m.train()
    with pytest.raises((NotImplementedError, RuntimeError)):
        m(x)


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: the product functions fail loudly instead of silently computing on the host."""
    import distill_any_depth_b200 as d
    m = d.DepthAnythingV2(**synthetic.MODEL_PRESETS["vits"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 28, 28))
    x = torch.rand(1, 1, 8, 8)
    for fn in (lambda: d.SSILoss()(x, x, x > 0.5), lambda: d.gradient_preservation_loss(x),
               lambda: d.distillation_loss(x, x, "hybrid"), lambda: d.get_contexts_dr(3, x, None),
               lambda: d.masked_shift_and_scale(x, x, x > 0.5),
               lambda: d.feature_distillation_loss(torch.rand(1, 4, 8), torch.rand(1, 4, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()


def test_product_package_does_not_import_the_oracle():
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); import distill_any_depth_b200; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)
    for root, _, files in os.walk(os.path.join(ROOT, "distill-any-depth_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                assert "import oracle" not in open(os.path.join(root, f)).read(), f


def test_shard_range_partitions_the_batch():
    from distill_any_depth_b200.dist import shard_range
    for n in (1, 7, 8, 128, 130):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _partials_cpu(pred, gt, fs, ft):
    """(numerator, denominator) of each loss for one shard, computed with the oracle's formulas."""
    ssi = oracle.SSILoss()
    full = torch.ones_like(gt, dtype=torch.bool)
    dense = ssi(pred, gt, full, dense=True)
    ctx = oracle.get_contexts_dr(3, gt, None)
    K = ctx.shape[0]
    rep = lambda x: x.unsqueeze(0).expand(K, *x.shape).reshape(-1, *x.shape[-3:])
    per = ssi(rep(pred), rep(gt), ctx.reshape(-1, *ctx.shape[-3:]), dense=True).reshape(*ctx.shape).sum(0)
    times = ctx.sum(0)
    valid = times != 0
    per = torch.where(valid, per / times.clamp(min=1), per)
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]).view(1, 1, 3, 3)
    gx = torch.nn.functional.conv2d(pred, kx, padding=1)
    gy = torch.nn.functional.conv2d(pred, kx.transpose(2, 3), padding=1)
    g = torch.exp(-torch.sqrt(gx ** 2 + gy ** 2 + 1e-6))
    sn = torch.nn.functional.normalize(fs, dim=1)
    tn = torch.nn.functional.normalize(ft, dim=1)
    cos = torch.nn.functional.cosine_similarity(sn, tn, dim=1)
    t64 = lambda a, b: torch.tensor([float(a), float(b)], dtype=torch.float64)
    return {"ssi": ("ssi", t64(dense.double().sum(), full.sum())),
            "hdn": ("hdn", t64(per.double().sum(), valid.sum())),
            "grad": ("grad", t64(g.double().sum(), g.numel())),
            "feat": ("feat", t64(cos.double().sum(), cos.numel()))}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from distill_any_depth_b200.dist import finish_losses, shard_batch
    pred, gt, _ = synthetic.make_depth_pair(6, 40, 56, seed=3)
    fs, ft = synthetic.make_features(6, 20, 32, seed=4), synthetic.make_features(6, 20, 32, seed=5)
    sl = lambda x: shard_batch(x, rank, world)
    out = finish_losses(_partials_cpu(sl(pred), sl(gt), sl(fs), sl(ft)))
    if rank == 0:
        q.put({k: float(v) for k, v in out.items()})
    dist.destroy_process_group()


def test_two_rank_losses_equal_single_process_full_batch():
    """8e: sharded partials + ONE all-reduce == the reference computed on the concatenated batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pred, gt, _ = synthetic.make_depth_pair(6, 40, 56, seed=3)
    fs, ft = synthetic.make_features(6, 20, 32, seed=4), synthetic.make_features(6, 20, 32, seed=5)
    ssi = oracle.SSILoss()
    ref = {"ssi": ssi(pred, gt, torch.ones_like(gt, dtype=torch.bool)),
           "hdn": oracle.compute_hdn_loss(ssi, pred, gt, oracle.get_contexts_dr(3, gt, None)),
           "grad": oracle.gradient_preservation_loss(pred),
           "feat": oracle.feature_distillation_loss(fs, ft)}
    for k, v in ref.items():
        assert abs(got[k] - float(v)) <= 1e-5 * max(abs(float(v)), 1e-6), (k, got[k], float(v))


def test_checkpoint_io_round_trips_both_key_layouts(tmp_path):
    """8f N3: safetensors / torch checkpoints in either reference key layout load into either native class
    (tools/train_distillation.py:743-793, ViT_DINO.py:1372-1388, tools/convert_checkpoint.py)."""
    import distill_any_depth_b200 as d
    from distill_any_depth_b200 import checkpoint as ck
    kw = synthetic.MODEL_PRESETS["vitl"]
    sd = synthetic.make_state_dict(seed=4, **kw)
    from safetensors.torch import save_file, load_file
    flat = str(tmp_path / "dav2_vitl.safetensors")           # DAv2 release layout: pretrained.blocks.N.*
    save_file({k: v.contiguous() for k, v in sd.items()}, flat)
    t = d.DepthAnything(**kw)
    res = ck.load_checkpoint(t, flat, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(t.backbone.blocks[0][5].attn.qkv.weight, sd["pretrained.blocks.5.attn.qkv.weight"])
    assert torch.equal(t.backbone.blocks[0][0].norm1.bias, sd["pretrained.blocks.0.norm1.bias"])  # block 0 is not "chunk 0"
    # the reference's half-converted layout (backbone.* but flat blocks, what convert_checkpoint.py writes) also loads
    half = {("backbone" + k[len("pretrained"):] if k.startswith("pretrained.") else k): v for k, v in sd.items()}
    half_path = str(tmp_path / "half.pth")
    torch.save({"state_dict": half}, half_path)
    t2 = d.DepthAnything(**kw)
    ck.load_checkpoint(t2, half_path, strict=True)
    assert all(torch.equal(a, b) for a, b in zip(t.state_dict().values(), t2.state_dict().values()))
    # teacher save -> student load (blocks.0.N -> blocks.N), bit-identical tensors
    out = str(tmp_path / "teacher.safetensors")
    ck.save_checkpoint(t, out)
    assert "backbone.blocks.0.23.mlp.fc2.weight" in load_file(out)
    s = d.DepthAnythingV2(**kw)
    ck.load_checkpoint(s, out, strict=True)
    for k, v in sd.items():
        assert torch.equal(s.state_dict()[k], v), k
    keys = ck.convert_checkpoint(flat, str(tmp_path / "conv.safetensors"), layout="teacher")
    assert "backbone.blocks.0.0.ls1.gamma" in keys and not any(k.startswith("pretrained.") for k in keys)
    with pytest.raises(RuntimeError):  # a foreign key is an error, not a silent strict=False fallback
        bad = dict(sd)
        bad["pretrained.blocks.0.attn.extra"] = torch.zeros(1)
        save_file({k: v.contiguous() for k, v in bad.items()}, str(tmp_path / "bad.safetensors"))
        ck.load_checkpoint(d.DepthAnythingV2(**kw), str(tmp_path / "bad.safetensors"), strict=True)


def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from distill_any_depth_b200.dist import shard_loss_weights, allreduce_gradients, shard_batch
    net, x, gt, mask = _toy_problem()
    sl = lambda t: shard_batch(t, rank, world)
    pred = net(sl(x))
    ssi = oracle.SSILoss()(pred, sl(gt), sl(mask))
    grad = oracle.gradient_preservation_loss(pred)
    t64 = lambda a, b: torch.tensor([float(a), float(b)], dtype=torch.float64)
    w = shard_loss_weights({"ssi": ("ssi", t64(0, sl(mask).sum())), "grad": ("grad", t64(0, pred.numel()))})
    (w["ssi"] * ssi + 0.2 * w["grad"] * grad).backward()
    unused = torch.nn.Parameter(torch.zeros(3))   # no gradient on any rank (the reference's mask_token): must stay None
    n = allreduce_gradients(list(net.parameters()) + [unused], bucket_bytes=64)   # tiny buckets: several collectives
    if rank == 0:
        q.put(([p.grad.clone() for p in net.parameters()], n, unused.grad is None))
    dist.destroy_process_group()


def _toy_problem():
    g = torch.Generator().manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 3, padding=1))
    for p in net.parameters():
        p.data = torch.randn(p.shape, generator=g) * 0.3
    x = torch.randn(5, 3, 24, 32, generator=g)      # 5 images over 2 ranks: uneven shards (3 + 2)
    gt = torch.rand(5, 1, 24, 32, generator=g) + 0.1
    mask = torch.rand(5, 1, 24, 32, generator=g) > 0.3   # per-shard valid counts differ
    return net, x, gt, mask


def test_two_rank_training_gradient_equals_single_process_full_batch():
    """N1 x 8e: shard-weighted local losses + SUM all-reduce of the gradients == the gradient of the full-batch loss."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, n_coll, unused_stays_none = q.get(timeout=120)
    assert unused_stays_none
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_coll >= 2
    net, x, gt, mask = _toy_problem()
    pred = net(x)
    (oracle.SSILoss()(pred, gt, mask) + 0.2 * oracle.gradient_preservation_loss(pred)).backward()
    for a, p in zip(got, net.parameters()):
        assert float((a - p.grad).abs().max()) <= 1e-5 * float(p.grad.abs().max()) + 1e-8


def test_colormap_lut_restates_matplotlib_spectral():
    """The 256-entry Spectral table (host logic of colorize_depth_maps): end points and nodes are the ColorBrewer colours,
    channels are piecewise linear between them, the reversed map is the mirror image; values that published matplotlib
    returns for Spectral (cm.Spectral(0.0), (0.5), (1.0)) are pinned as known answers."""
    import numpy as np
    from distill_any_depth_b200 import preprocess
    lut = preprocess.colormap_lut("Spectral")
    assert lut.shape == (256, 3) and lut.min() >= 0 and lut.max() <= 1
    np.testing.assert_allclose(lut[0], np.array([158, 1, 66]) / 255.0, atol=1e-15)
    np.testing.assert_allclose(lut[255], np.array([94, 79, 162]) / 255.0, atol=1e-15)
    # matplotlib: cm.Spectral(0.5) -> index 128 -> between the 6th node (255,255,191) at 127.5 and the 7th (230,245,152) at 153
    t = (128 - 127.5) / 25.5
    np.testing.assert_allclose(lut[128], (np.array([255, 255, 191]) + t * (np.array([230, 245, 152]) - np.array([255, 255, 191]))) / 255.0,
                               atol=1e-12)
    np.testing.assert_allclose(preprocess.colormap_lut("Spectral_r"), lut[::-1], atol=1e-12)
    d2 = np.diff(lut, n=2, axis=0)           # piecewise linear: second differences vanish away from the 9 interior nodes
    assert (np.abs(d2).max(axis=1) > 1e-9).sum() <= 18
    import oracle.preprocess as P
    img = P.colorize_depth_maps(np.array([[[0.0, 0.5, 1.0], [1.0, 0.5, 0.0]]], dtype=np.float32), 0, 1, lut=lut)
    np.testing.assert_allclose(img[0, :, 0, :].T, lut[[0, 128, 255]], atol=0)
    np.testing.assert_allclose(img[0, :, 1, :].T, lut[[255, 128, 0]], atol=0)
    # the uint8 image truncates the FLOAT64 product lut * 255 (254 / 255 * 255 = 253.99999999999997 -> 253): the device
    # therefore indexes a uint8 table computed here in float64, never its own fp32 product
    assert (lut[255] * 255).astype(np.uint8).tolist() == [94, 79, 162]
