#!/usr/bin/env python
"""Headline benchmark: images/sec of the DepthAnythingV2 ViT-L 518x518 forward + SSI / HDN-DR loss
reductions (BASELINE.json metric) on N B200s, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W                  # this repo's CUDA path
  torchrun ... bench.py --gpus N --steps K --warmup W            # N ranks, images sharded by rank (weak scaling)
  python bench.py --impl reference --gpus N --steps K --warmup W # reference algorithm on the host CPU cores
                                                                 # (oracle port; rank 0 only)

One "step" = one batch of B synthetic images through forward + SSILoss + fused HDN-DR loss against a
synthetic teacher map.  `value` times the step with inputs resident in HBM; `e2e` times the same
step through the public Python API with the images starting in pinned HOST memory (H2D copy of the
batch and D2H read of the two loss scalars inside the timed region).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec at 518x518 ViT-L fwd+SSI/HDN loss, 1/2/4/8 B200; % of roofline"
PROF_CLASSES = ["gemm_tc", "gemm_simt", "attention", "layernorm", "elementwise", "loss"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step (BASELINE configs[2]: 32)")
    ap.add_argument("--encoder", default="vitl", choices=["vits", "vitb", "vitl"])
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the same-box PyTorch-eager GPU baseline (an extra)")
    ap.add_argument("--graph", type=int, default=1, choices=[0, 1],
                    help="1 (default): capture one step (forward + losses [+ all-reduce]) in a CUDA graph after warm-up and "
                         "replay it in the timed loops; 0: launch every kernel from the host each step")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5", "train", "train4", "infer"],
                    help="c3 (default, the metric's config): ViT-L 518^2 B=32 bf16 fwd + SSI + HDN-DR; the others are "
                         "BASELINE.json's remaining GPU configs, for DESIGN.md's table (not bench lines): c2 ViT-B 392^2 "
                         "B=16 + SSI/grad; c4 distillation step teacher ViT-L + student ViT-B 392^2 B=16/GPU, 5 losses; "
                         "c5 ViT-L 1036^2 forward, batch 8 split over the GPUs (strong scaling)")
    a = ap.parse_args()
    if a.workload == "c2":
        a.encoder, a.size, a.batch = "vitb", 392, 16
    elif a.workload == "c4":
        a.encoder, a.size, a.batch = "vitb", 392, 16
    elif a.workload == "c5":
        a.encoder, a.size, a.batch = "vitl", 1036, max(1, 8 // max(a.gpus, 1))
    elif a.workload in ("train", "train4"):   # student update (forward + SSI / gradient loss + backward), C2's model and shape
        a.encoder, a.size, a.batch = "vitb", 392, 16
    return a


def workload_name(a):
    if a.workload == "c2":
        return (f"DepthAnythingV2 vitb/128 392x392 batch {a.batch}/GPU {a.precision} forward + SSI + gradient loss "
                f"(BASELINE configs[1])")
    if a.workload == "c4":
        return (f"distillation step: DepthAnything vitl teacher + DepthAnythingV2 vitb student (two student forwards as "
                f"upstream) 392x392 batch {a.batch}/GPU {a.precision}, SC/LG(hybrid) + feature + gradient + HDN-DR losses "
                f"(BASELINE configs[3], forward half)")
    if a.workload == "c5":
        return (f"DepthAnythingV2 vitl 1036x1036 (5476 tokens) batch {a.batch}/GPU {a.precision} forward "
                f"(BASELINE configs[4]; batch 8 split over the GPUs)")
    return (f"DepthAnythingV2 {a.encoder} {a.size}x{a.size} batch {a.batch}/GPU {a.precision} forward + "
            f"SSI + HDN-DR(level 3) loss, synthetic images, random-init weights (BASELINE configs[2] + loss)")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ---------------------------------------------------------------------------------- CPU reference arm
def cpu_step_factory(a, B):
    """The reference algorithm on the host: oracle forward + SSI + HDN-DR, fp32, all host threads."""
    import torch
    import oracle
    from distill_any_depth_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    kw = synthetic.MODEL_PRESETS[a.encoder]
    sd = synthetic.make_state_dict(seed=1, **kw)
    x = synthetic.make_images(B, a.size, a.size, seed=1234)
    _, gt, _ = synthetic.make_depth_pair(B, a.size, a.size, seed=7)
    ssi = oracle.SSILoss()
    full = torch.ones_like(gt, dtype=torch.bool)

    def step():
        with torch.no_grad():
            depth, _ = oracle.depth_anything_forward(x, sd, kw["encoder"])
            l1 = ssi(depth, gt, full)
            l2 = oracle.compute_hdn_loss(ssi, depth, gt, oracle.get_contexts_dr(3, gt, None))
        return float(l1), float(l2)
    return step, torch.get_num_threads()


def cpu_baseline(a):
    step, cores = cpu_step_factory(a, 1)
    step()
    best = 1e30
    for _ in range(2):
        t = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t)
    return dict(value=1.0 / best, unit="images/s", cores=cores, kind="port",
                sample="B=1 of the same workload (oracle fp32 forward + SSI + HDN-DR on host threads), best of 2 after 1 warm-up")


def gpu_eager_baseline(a, dev):
    """Same-box GPU baseline: the reference algorithm as plain PyTorch eager ops on the B200 - the oracle's
    functions on CUDA tensors, i.e. cuBLAS / cuDNN / ATen kernels with the reference's naive attention
    (dinov2_layers/attention.py:49-62) - in fp32 (TF32 off) and under torch.autocast(bfloat16).  A stated extra, never the
    target: it says what the hand-written kernels buy over the library path on the same GPU."""
    import torch
    import oracle
    from distill_any_depth_b200 import synthetic
    kw = synthetic.MODEL_PRESETS[a.encoder]
    sd = {k: v.to(dev) for k, v in synthetic.make_state_dict(seed=1, **kw).items()}
    B = a.batch
    x = synthetic.make_images(B, a.size, a.size, seed=1234).to(dev)
    _, gt, _ = synthetic.make_depth_pair(B, a.size, a.size, seed=7)
    gt = gt.to(dev)
    full = torch.ones_like(gt, dtype=torch.bool)
    ssi = oracle.SSILoss()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def step(autocast):
        with torch.no_grad():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                depth, _ = oracle.depth_anything_forward(x, sd, kw["encoder"])
            depth = depth.float()
            l1 = ssi(depth, gt, full)
            l2 = oracle.compute_hdn_loss(ssi, depth, gt, oracle.get_contexts_dr(3, gt, None))
        return torch.stack([l1, l2])

    out = {}
    try:
        for name, ac in (("fp32_tf32_off", False), ("autocast_bf16", True)):
            step(ac)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 2
            e0.record()
            for _ in range(n):
                last = step(ac)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = dict(value=B / ms * 1e3, unit="images/s", ms_per_step=ms, batch=B, losses=[float(v) for v in last.cpu()])
        out["what"] = ("reference algorithm (oracle port) as PyTorch eager ops on this GPU: cuBLAS / cuDNN / ATen, naive attention; "
                       "2 steps after 1 warm-up, inputs resident in HBM")
    except Exception as ex:  # an extra: never let it take the bench line down
        out["error"] = f"{type(ex).__name__}: {ex}"
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        del sd, x
        torch.cuda.empty_cache()
    return out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, cores = cpu_step_factory(a, 1)
    for _ in range(max(1, min(a.warmup, 2))):
        step()
    steps = max(1, min(a.steps, 8))  # bounded sample: B=1 per step, <= 8 steps (a ViT-L step is seconds on CPU)
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t
    v = steps / dt
    line = dict(metric=METRIC, value=v, unit="images/s", n_gpus=a.gpus, steps=steps, warmup=a.warmup,
                ms_per_step=dt / steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=workload_name(a), note="reference algorithm (oracle port of the PyTorch fp32 path) "
                            "on the host CPU, B=1 per step"),
                cpu_baseline=dict(value=v, unit="images/s", cores=cores, kind="port",
                                  sample=f"{steps} steps of B=1 of the same workload on {cores} host threads"),
                e2e=dict(value=v, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:  # CUDA ordinal -> NVML handle by UUID (the two numberings differ under CUDA_VISIBLE_DEVICES)
                import torch
                uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(s))


# ---------------------------------------------------------------------------------- training step (not a bench line)
def run_train(a):
    """Student update of the reference loop (tools/train_distillation.py:1509-1575) for BASELINE configs[1]'s model and
    shape: forward + SSI + gradient-preservation loss + backward into .grad of all parameters, single GPU.  For
    DESIGN.md's table; the headline metric stays the forward + loss workload."""
    import ctypes
    import torch
    import torch.distributed as dist
    import distill_any_depth_b200 as d
    from distill_any_depth_b200 import synthetic, _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)   # NCCL prints its banner to fd 1: keep stdout to the ONE JSON line
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:   # data-parallel update (train4 only): images sharded by rank, exact full-batch gradient (step.py)
        assert a.workload == "train4", "multi-GPU training is wired for --workload train4"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, H = a.batch, a.size
    kw = synthetic.MODEL_PRESETS[a.encoder]
    model = d.DepthAnythingV2(**kw)
    model.load_state_dict(synthetic.make_state_dict(seed=1, **kw), strict=True)
    model = model.to(dev).train()
    model.precision = a.precision
    model.bf16_backward = True
    x = synthetic.make_images(B, H, H, seed=1234 + rank).to(dev)
    _, gt, _ = synthetic.make_depth_pair(B, H, H, seed=7 + rank)
    gt = gt.to(dev)
    full = torch.ones_like(gt, dtype=torch.bool)

    teacher = None
    if a.workload == "train4":   # the reference's whole update (:1503-1575): ViT-L teacher, two student forwards, five losses
        from distill_any_depth_b200.dam import student_to_teacher_keys
        tkw = synthetic.MODEL_PRESETS["vitl"]
        teacher = d.DepthAnything(**tkw)
        teacher.load_state_dict(student_to_teacher_keys(synthetic.make_state_dict(seed=2, head_bias=0.6, **tkw)), strict=True)
        teacher = teacher.to(dev).eval()
        teacher.precision = a.precision
        x2 = synthetic.make_images(B, H, H, seed=4321 + rank).to(dev)
        opt = torch.optim.SGD(model.parameters(), lr=1e-6)

    def step():
        if teacher is not None:
            return d.distillation_train_step(model, teacher, x, x2, optimizer=opt)["batch_loss"]
        for p in model.parameters():
            p.grad = None
        depth, _ = model(x)
        loss = d.SSILoss()(depth, gt, full) + 0.2 * d.gradient_preservation_loss(depth)
        loss.backward()
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    l0 = lib.dad_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    barrier()
    ms_t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)   # device time, max over ranks
    ms = ms_t.item()
    launches = lib.dad_launch_count() - l0
    lib.dad_profile_enable(1)
    step()
    torch.cuda.synchronize()
    breakdown = {}
    for ci, name in enumerate(PROF_CLASSES):
        t, w, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.dad_profile_get(ci, ctypes.byref(t), ctypes.byref(w), ctypes.byref(n))
        if n.value:
            breakdown[name] = dict(ms_per_step=t.value, launches_per_step=n.value, work_per_step=w.value)
    lib.dad_profile_enable(0)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return
    os.write(real_stdout, (json.dumps(dict(metric="images/sec, student training step (forward + SSI/gradient loss + backward)", value=world * B / ms * 1e3,
                          unit="images/s", n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=ms, scaling="weak",
                          higher_is_better=True, dtype=a.precision, data="synthetic",
                          config=dict(workload=(f"distillation update: ViT-L teacher forward + 2x {a.encoder} student forward + 5 losses + "
                                                f"backward + gradient all-reduce + SGD step, {H}x{H} batch {B}/GPU {a.precision}, dp{world} "
                                                f"(BASELINE configs[3], training half; not the headline metric)") if teacher is not None else
                                               f"DepthAnythingV2 {a.encoder} {H}x{H} batch {B} {a.precision} train step "
                                               f"(SURVEY 8f N1; not the headline metric)"),
                          loss=float(loss), gpu_launches=int(launches), kernel_breakdown=breakdown)) + "\n").encode())


# ---------------------------------------------------------------------------------- infer_image throughput (not a bench line)
def run_infer(a):
    """SURVEY 8f N2 / VERDICT r1 weak 10: the inference path a user calls - ``model.infer_image(raw_bgr_uint8)``
    (depth_anything_v2/dpt.py:227-235): uint8 image H2D -> /255, INTER_CUBIC resize, normalise (one kernel) -> forward ->
    bilinear resize back to the raw resolution -> depth map D2H.  One image per call, as upstream; plus the colourised
    uint8 visualisation (tools/testers/infer.py:134-140) as a second timed variant."""
    import numpy as np
    import torch
    import distill_any_depth_b200 as d
    from distill_any_depth_b200 import synthetic, _lib, preprocess
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    kw = synthetic.MODEL_PRESETS[a.encoder]
    model = d.DepthAnythingV2(**kw)
    model.load_state_dict(synthetic.make_state_dict(seed=1, **kw), strict=True)
    model = model.to(dev).eval()
    model.precision = a.precision
    torch.set_grad_enabled(False)
    rng = np.random.Generator(np.random.PCG64(5))
    h, w = 480, 640
    raws = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(4)]

    def plain(i):
        return model.infer_image(raws[i & 3], a.size)

    def coloured(i):
        image, (hh, ww) = model.image2tensor(raws[i & 3], a.size)
        depth, _ = model(image)
        depth = preprocess.resize_depth(depth, (hh, ww))
        norm = preprocess.normalize_minmax(depth)
        _, rgb8 = preprocess.colorize_depth_maps(norm, 0.0, 1.0, cmap="Spectral_r", as_uint8_hwc=True)
        return rgb8[0].cpu().numpy()

    out = {}
    n = max(a.steps, 8) * 4
    for name, fn in (("infer_image", plain), ("infer_image_colourised", coloured)):
        for i in range(max(a.warmup, 3)):
            fn(i)
        torch.cuda.synchronize()
        l0 = lib.dad_launch_count()
        t0 = time.perf_counter()
        for i in range(n):
            res = fn(i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[name] = dict(value=n / dt, unit="images/s", ms_per_image=dt / n * 1e3, launches_per_image=(lib.dad_launch_count() - l0) / n,
                         out_shape=list(res.shape), out_dtype=str(res.dtype))
    nw, nh = preprocess.get_size(w, h, a.size, a.size)
    print(json.dumps(dict(metric="images/sec through DepthAnythingV2.infer_image (one image per call, host uint8 in, host map out)",
                          value=out["infer_image"]["value"], unit="images/s", n_gpus=1, steps=n, warmup=max(a.warmup, 3),
                          ms_per_step=out["infer_image"]["ms_per_image"], higher_is_better=True, dtype=a.precision, data="synthetic",
                          config=dict(workload=f"infer_image: {h}x{w} BGR uint8 -> {a.encoder} at {nh}x{nw} {a.precision} -> {h}x{w} fp32 "
                                               f"depth (SURVEY 8f N2; not the headline metric)"),
                          e2e=dict(value=out["infer_image"]["value"], unit="images/s", h2d_bytes_per_step=h * w * 3,
                                   d2h_bytes_per_step=h * w * 4),
                          variants=out)))


# ---------------------------------------------------------------------------------- B200 arm
def run_b200(a):
    import ctypes
    import torch
    import torch.distributed as dist
    import distill_any_depth_b200 as d
    from distill_any_depth_b200 import synthetic, _lib, losses
    from distill_any_depth_b200.dist import stack_partials, finish_stacked
    from distill_any_depth_b200.step import distillation_step_partials, combine_step_losses

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a GPU; there is no CPU fallback"
    if a.workload in ("train", "train4"):
        return run_train(a)
    if a.workload == "infer":
        return run_infer(a)
    torch.set_grad_enabled(False)   # inference benchmark: never take the differentiable (activation-tape) forward
    # stdout must carry exactly ONE JSON line: NCCL prints its version banner to the process's stdout (fd 1), so fd 1 is
    # pointed at stderr for the rest of the run and the JSON line is written to a saved duplicate of the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the ONE JSON line (NCCL logs to stdout)
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    B, H = a.batch, a.size
    kw = synthetic.MODEL_PRESETS[a.encoder]
    model = d.DepthAnythingV2(**kw)
    model.load_state_dict(synthetic.make_state_dict(seed=1, **kw), strict=True)
    model = model.to(dev).eval()
    model.precision = a.precision
    x_host = synthetic.make_images(B, H, H, seed=1234 + rank).pin_memory()
    x_dev = x_host.to(dev)
    _, gt, _ = synthetic.make_depth_pair(B, H, H, seed=7 + rank)
    gt = gt.to(dev)
    full = torch.ones_like(gt, dtype=torch.bool)
    copy_stream = torch.cuda.Stream(dev)
    x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]

    teacher = None
    if a.workload == "c4":
        from distill_any_depth_b200.dam import student_to_teacher_keys
        tkw = synthetic.MODEL_PRESETS["vitl"]
        teacher = d.DepthAnything(**tkw)
        teacher.load_state_dict(student_to_teacher_keys(synthetic.make_state_dict(seed=2, head_bias=0.6, **tkw)), strict=True)
        teacher = teacher.to(dev).eval()
        teacher.precision = a.precision

    # A step is split in two halves.  compute(x): the forwards and the loss kernels, returning either the finished loss
    # scalars (1 GPU) or the stacked per-rank (numerator, denominator) partials (N > 1) - pure device work, capturable in a
    # CUDA graph.  finish(out): on N > 1 the ONE all-reduce of the partials plus the ratios, launched
    # eagerly AFTER the graph replay: a captured graph that contains NCCL kernels made destroy_process_group() hang at
    # exit (round 1 left through os._exit for that reason); with compute-only graphs the teardown is clean.
    meta = {}

    def compute(x):
        if a.workload == "c4":
            if world == 1:
                out = d.distillation_step_losses(model, teacher, x, x)
                return torch.stack([out["batch_loss"], out["hdn_loss"]])
            names, kinds, vec = stack_partials(distillation_step_partials(model, teacher, x, x))
            meta["names"], meta["kinds"] = names, kinds
            return vec
        depth, _ = model(x)
        if a.workload == "c5":   # forward only; the D2H read is one element of the depth map
            return depth.view(-1)[:2].clone()
        if a.workload == "c2":
            first = ("ssi", losses._ssi(depth, gt, full, False, want_partials=True))
            second = ("grad", losses._grad(depth, want_partials=True))
        elif os.environ.get("DAD_BENCH_SEPARATE_LOSSES"):   # A/B: the two separate kernel sequences of round 1
            first = ("ssi", losses._ssi(depth, gt, full, False, want_partials=True))
            second = ("hdn", losses.hdn_loss_dr(depth, gt, None, 3, want_partials=True))
        else:   # SSILoss + HDN-DR of the same maps: one shared sweep (losses_fused.cu)
            ssi_v, hdn_v, p_ssi, p_hdn = losses.ssi_hdn_dr(depth, gt, None, 3, want_partials=True)
            first, second = ("ssi", (ssi_v, p_ssi)), ("hdn", (hdn_v, p_hdn))
        if world == 1:
            return torch.stack([first[1][0], second[1][0]])
        names, kinds, vec = stack_partials({first[0]: (first[0], first[1][1]), second[0]: (second[0], second[1][1])})
        meta["names"], meta["kinds"] = names, kinds
        return vec

    def finish(out):
        if world == 1 or a.workload == "c5":
            return out
        fin = finish_stacked(meta["names"], meta["kinds"], out)   # all-reduce in place + ratios
        if a.workload == "c4":
            fin = combine_step_losses(fin)
            return torch.stack([fin["batch_loss"], fin["hdn_loss"]])
        return torch.stack([fin[n] for n in meta["names"]])

    def step_device():
        return finish(compute(x_dev))

    def step_e2e(i):
        buf = x_stage[i & 1]
        with torch.cuda.stream(copy_stream):  # H2D of this step's batch from pinned host memory
            buf.copy_(x_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        torch.cuda.current_stream().wait_event(ev)
        return finish(compute(buf)).cpu()  # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            out = fn(i)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item(), out

    for i in range(max(a.warmup, 3)):
        step_device()
        step_e2e(i)
    torch.cuda.synchronize()
    l_step0 = lib.dad_launch_count()
    step_device()
    launches_per_step = lib.dad_launch_count() - l_step0

    # ---- CUDA graphs: one step is ~240 dependent launches; replaying a captured graph removes the per-launch host
    # and front-end cost.  The same kernels run with the same arguments (tensor maps are baked into the nodes; the
    # workspace, inputs and outputs are static buffers).  Multi-GPU: the graph holds the compute half only; the ONE
    # all-reduce of the loss partials is launched eagerly after each replay (see compute / finish above).
    graphs = {}
    if a.graph:
        try:
            cap_stream = torch.cuda.Stream(dev)
            cap_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap_stream):
                for key, xin in (("dev", x_dev), ("s0", x_stage[0]), ("s1", x_stage[1])):
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph, stream=cap_stream):
                        out_static = compute(xin)   # compute only: no collective inside the graph
                    graphs[key] = (gph, out_static)
            torch.cuda.current_stream().wait_stream(cap_stream)
            torch.cuda.synchronize()
        except Exception as ex:  # stay measurable: fall back to host launches and say so in the JSON line
            print(f"bench: CUDA graph capture failed ({type(ex).__name__}: {ex}); timing host launches", file=sys.stderr)
            graphs = {}
            a.graph = 0
            torch.cuda.synchronize()
    if graphs:

        def step_device():  # noqa: F811
            gph, out_static = graphs["dev"]
            gph.replay()
            return finish(out_static)

        def issue_copy(i):  # H2D of step i's batch from pinned host memory, on the copy stream
            with torch.cuda.stream(copy_stream):
                x_stage[i & 1].copy_(x_host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return ev

        pending = {}

        def step_e2e(i, last=False):  # noqa: F811
            # double-buffered input pipeline: step i's copy was issued during step i - 1 (or just now for the first
            # step of a timed loop); step i + 1's copy is issued before step i's graph is replayed, so the PCIe
            # transfer overlaps the compute.  Every step's H2D copy and D2H read stay inside the timed region.
            ev = pending.pop(i, None) or issue_copy(i)
            torch.cuda.current_stream().wait_event(ev)
            if not last:
                pending[i + 1] = issue_copy(i + 1)   # buffer (i + 1) & 1 was last read by step i - 1, which has completed
            gph, out_static = graphs["s%d" % (i & 1)]
            gph.replay()
            return finish(out_static).cpu()

        for i in range(2):
            step_device()
            step_e2e(i, last=True)
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev, _, last = timed(lambda i: step_device(), a.steps)
    launches = launches_per_step * a.steps  # kernels of this library inside the device-timed region
    if graphs:
        ms_e2e, wall_e2e, last_e2e = timed(lambda i: step_e2e(i, last=(i == a.steps - 1)), a.steps)
    else:
        ms_e2e, wall_e2e, last_e2e = timed(step_e2e, a.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- per-kernel-class device times (separate pass so the events do not perturb the timed loops; launched
    # from the host, not replayed, because the events bracket individual launches)
    def step_host():
        return finish(compute(x_dev))

    lib.dad_profile_enable(1)
    psteps = max(1, min(a.steps, 3))
    for _ in range(psteps):
        step_host()
    torch.cuda.synchronize()
    prof = {}
    for ci, name in enumerate(PROF_CLASSES):
        ms_, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        _lib.check(lib.dad_profile_get(ci, ctypes.byref(ms_), ctypes.byref(work), ctypes.byref(n)))
        if n.value:
            prof[name] = dict(ms_per_step=ms_.value / psteps, launches_per_step=n.value / psteps,
                              work_per_step=work.value / psteps, avg_launch_ms=ms_.value / n.value)
    lib.dad_profile_enable(0)
    pk_ = peaks()
    for name, rec in prof.items():  # every class against ITS roofline (tensor for GEMM / attention, HBM for the rest)
        per_s = rec["work_per_step"] / (rec["ms_per_step"] * 1e-3)
        if name in ("gemm_tc", "gemm_simt", "attention"):
            rec.update(bound="tensor", achieved_tflops=per_s / 1e12, frac=per_s / 1e12 / pk_["tf_sustained"])
        else:
            rec.update(bound="hbm", achieved_gbs=per_s / 1e9, frac=per_s / 1e9 / pk_["hbm"])

    def finish_ranks():
        # Multi-rank exit: rendezvous, drop the CUDA graphs (they contain this communicator's NCCL kernels), then
        # destroy the process group.  Round 1 left through os._exit(0) because destroy_process_group() hung while the
        # captured graphs were still alive; DAD_BENCH_CLEAN_EXIT=0 keeps that path as an escape hatch.
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            if os.environ.get("DAD_BENCH_CLEAN_EXIT", "1") != "0":
                # Clean teardown: the captured graphs hold NCCL kernels of this communicator; they must be destroyed
                # BEFORE the process group (destroying the communicator under live graph nodes is what used to hang)
                graphs.clear()
                import gc
                gc.collect()
                torch.cuda.synchronize()
                # belt and braces: the work is done and the JSON line is out; if the teardown ever stalls, leave after 30 s
                # (rc 0) instead of hanging the launcher, and say so on stderr
                def _watchdog():
                    time.sleep(30)
                    print("bench: destroy_process_group() did not return within 30 s; exiting", file=sys.stderr, flush=True)
                    os._exit(0)
                threading.Thread(target=_watchdog, daemon=True).start()
                dist.destroy_process_group()
                return
            time.sleep(0.5)   # legacy exit (DAD_BENCH_CLEAN_EXIT=0): leave without tearing NCCL down
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish_ranks()
        return
    pk = peaks()
    imgs = B * world * a.steps
    value = imgs / ms_dev * 1e3
    e2e_value = imgs / max(ms_e2e, wall_e2e if world == 1 else ms_e2e) * 1e3
    gt_ = prof.get("gemm_tc")
    traffic = None
    traffic_source = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture (tools/summarize_ncu.py): ncu cannot run inside
        # a timed bench, so this figure is FROM THE PROFILE of the same command, and labelled so
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tr = json.load(f)
        if a.workload == "c3" and a.batch == 32:
            traffic = tr["classes"]["gemm_tc"]["dram_bytes_per_launch"]
            traffic_source = "from_profile: profiles/roofline_traffic.json (" + str(tr.get("source", "ncu --set full")) + ")"
    except Exception:
        pass
    roofline = None
    if gt_:
        # dominant kernel: gemm_tc (tcgen05 GEMM / implicit-GEMM conv).  achieved = algorithmic FLOPs of all its
        # launches in a step / their summed CUDA-event durations; peak = sustained cuBLAS bf16 (kernel timed
        # inside a long step), from MEASURED_PEAKS.json.
        ach = gt_["work_per_step"] / (gt_["ms_per_step"] * 1e-3) / 1e12
        roofline = dict(bound="tensor", achieved=ach, peak=pk["tf_sustained"], unit="TFLOP/s", frac=ach / pk["tf_sustained"],
                        traffic=traffic, traffic_source=traffic_source, flops_per_launch=gt_["work_per_step"] / gt_["launches_per_step"],
                        kernel="tcgen05 GEMM / implicit-GEMM conv kernels (gemm_tc, gemm_tc2, conv_tc2: all launches of a step)", peak_source=pk["src"] + " sustained bf16",
                        launches_per_step=gt_["launches_per_step"], avg_launch_ms=gt_["avg_launch_ms"],
                        share_of_step=gt_["ms_per_step"] / (ms_dev / a.steps))
    line = dict(metric=METRIC, value=value, unit="images/s", n_gpus=world, steps=a.steps, warmup=max(a.warmup, 3),
                ms_per_step=ms_dev / a.steps, higher_is_better=True, scaling="strong" if a.workload == "c5" else "weak",
                vs_baseline=None,
                dtype=a.precision, data="synthetic",
                config=dict(workload=workload_name(a), global_batch=B * world, per_gpu_batch=B,
                            parallelism=f"dp{world} (images sharded by rank; one all-reduce of loss partials per step, outside the graph)",
                            l2="working set (activations ~9 GB/step at B=32) >> 126 MB L2; no explicit flush",
                            cuda_graph=bool(a.graph),
                            losses=[float(v) for v in last.cpu()]),
                e2e=dict(value=e2e_value, unit="images/s", h2d_bytes_per_step=int(x_host.numel() * 4),
                         d2h_bytes_per_step=8, ms_per_step=ms_e2e / a.steps),
                gpu_launches=int(launches), clocks=sampler.result(), roofline=roofline, kernel_breakdown=prof,
                kernel_breakdown_source="separate pass after the timed loops: the same step launched from the host with a CUDA "
                                        "event pair around every library launch (the timed loops replay a CUDA graph)")
    if world == 1 and not a.no_cpu_baseline and a.workload == "c3":
        line["cpu_baseline"] = cpu_baseline(a)
    if world == 1 and not a.no_gpu_eager and a.workload == "c3":
        torch.set_grad_enabled(False)
        line["gpu_eager_baseline"] = gpu_eager_baseline(a, dev)
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    finish_ranks()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
