#!/usr/bin/env python
"""Power / clock telemetry per kernel class (run on the GPU box):  python tools/power_probe.py [seconds per class]

The headline step runs under `sw_power_cap` at ~1500 of 1965 MHz (bench.py's `clocks`), so what limits it is energy per unit
of work, class by class.  This probe loops ONE representative launch of each class at the benchmark's shapes (ViT-L 518^2,
batch 32) for a few seconds while sampling NVML power draw, SM clock and throttle reasons, and prints one JSON line per class:
achieved TFLOP/s (or GB/s), median watts, median SM MHz.  A class that holds the maximum clock is not what the cap bites on.
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Sampler(threading.Thread):
    def __init__(self, handle, nv):
        super().__init__(daemon=True)
        self.h, self.nv, self.stop, self.w, self.mhz = handle, nv, False, [], []

    def run(self):
        while not self.stop:
            try:
                self.w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                pass
            time.sleep(0.02)


def main():
    import torch
    import pynvml
    from distill_any_depth_b200 import _lib as L
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    lib = L.load()
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(0).uuid)).encode())
    limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    st = L.stream_ptr()
    B, T, D, heads = 32, 1370, 1024, 16
    M = B * T

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).bfloat16()

    cases = []
    # encoder GEMMs (2-CTA kernel): qkv (bias -> bf16) and fc1 (bias + GELU -> bf16)
    for name, N, K, act in (("gemm qkv 43840x3072x1024", 3 * D, D, 0), ("gemm fc1+gelu 43840x4096x1024", 4 * D, D, 1),
                            ("gemm fc2 43840x1024x4096 (bias->bf16)", D, 4 * D, 0)):
        A, W, bias = rnd(M, K), rnd(N, K, scale=0.03), torch.zeros(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        cases.append((name, 2.0 * M * N * K, "TFLOP/s", 1e12, (A, W, bias, out),
                      lambda A=A, W=W, bias=bias, out=out, N=N, K=K, act=act: L.check(lib.dad_gemm_ex(
                          L.ptr(A), L.ptr(W), L.ptr(bias), None, None, 0, L.ptr(out), 1, act, M, N, K, 0, st))))
    # fused attention
    qkv, att = rnd(M, 3 * D), torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    cases.append(("attention 32x16 heads x 1370^2", 4.0 * B * heads * T * T * 64, "TFLOP/s", 1e12, (qkv, att),
                  lambda: L.check(lib.dad_attention(L.ptr(qkv), L.ptr(att), B, T, heads, 0, st))))
    # HBM streaming reference: device-to-device copy of 1 GiB
    src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    cases.append(("copy 1 GiB (read + write)", 2.0 * (1 << 30), "GB/s", 1e9, (src, dst), lambda: dst.copy_(src)))

    print(json.dumps(dict(device=torch.cuda.get_device_name(0), power_limit_w=limit, seconds_per_class=secs)))
    for name, work, unit, div, _keep, fn in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s = Sampler(h, pynvml)
        s.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n, t0 = 0, time.perf_counter()
        e0.record()
        while time.perf_counter() - t0 < secs:
            for _ in range(8):
                fn()
            n += 8
            torch.cuda.synchronize()   # keep the queue short so the loop ends on time
        e1.record()
        torch.cuda.synchronize()
        s.stop = True
        s.join(timeout=1)
        ms = e0.elapsed_time(e1) / n
        w, mhz = sorted(s.w[len(s.w) // 4:]), sorted(s.mhz[len(s.mhz) // 4:])   # drop the ramp-up quarter
        print(json.dumps(dict(kernel=name, ms=round(ms, 4), achieved=round(work / (ms * 1e-3) / div, 1), unit=unit,
                              watts=w[len(w) // 2] if w else None, sm_mhz=mhz[len(mhz) // 2] if mhz else None,
                              launches=n)), flush=True)


if __name__ == "__main__":
    main()
