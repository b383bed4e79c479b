#!/usr/bin/env python
"""Turn the raw ncu outputs of tools/profile_round.sh (gpurun_out/) into the small tracked summaries under
profiles/:   python tools/summarize_ncu.py <tag>

  profiles/ncu_launches_<tag>.json   one steady-state step: per kernel name -> launches, total us, share of step
  profiles/ncu_full_<tag>.json       per kernel name (and GEMM shape class) -> DRAM bytes per launch, tensor-pipe %,
                                     issue-active %, duration; `traffic` for bench.py's roofline comes from here
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"void |dad::|<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def read_csv_after_header(path, first_col="ID"):
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == first_col)
    return rows[hi], rows[hi + 1:]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    hdr, rows = read_csv_after_header(path)
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(short(r[ik]), float(r[iv].replace(",", "")) / 1e3, r[ig]) for r in rows if len(r) > iv]
    # one steady-state step = the launches between the last two patch_im2col kernels of the device-timed loop;
    # use the LAST complete period in the list (the profile pass of bench.py)
    starts = [i for i, s in enumerate(seq) if s[0].startswith("patch_im2col")]
    lo, hi = starts[-2], starts[-1]
    step = seq[lo:hi]
    agg = OrderedDict()
    for n, us, _ in step:
        a = agg.setdefault(n, dict(launches=0, us=0.0))
        a["launches"] += 1
        a["us"] += us
    tot = sum(a["us"] for a in agg.values())
    for a in agg.values():
        a["share"] = a["us"] / tot
        a["us"] = round(a["us"], 1)
    out = dict(tag=tag, note="ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache serialised launches: "
               "compare SHARES, not absolutes", launches_in_step=len(step), total_us=round(tot, 1),
               total_launches_in_run=len(seq),
               kernels=OrderedDict(sorted(agg.items(), key=lambda kv: -kv[1]["us"])))
    with open(os.path.join(ROOT, "profiles", f"ncu_launches_{tag}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"step: {len(step)} launches, {tot / 1e3:.2f} ms under ncu")
    for n, a in list(out["kernels"].items())[:25]:
        print(f"  {a['share'] * 100:5.1f}%  {a['us'] / 1e3:8.3f} ms  x{a['launches']:3d}  {n}")
    return step


WANT = OrderedDict([
    ("duration_us", "gpu__time_duration.sum"),
    ("dram_read_bytes", "dram__bytes_read.sum"),
    ("dram_write_bytes", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("sm_mhz", "sm__cycles_elapsed.avg.per_second"),
])
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
              "ms ": 1e3}


def full(tag):
    path = os.path.join(ROOT, "gpurun_out", f"full_raw_{tag}.csv")
    if not os.path.exists(path) or os.path.getsize(path) < 1000:
        print("no full capture for", tag)
        return
    rows = list(csv.reader(open(path, newline="")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    ik, ig = hdr.index("Kernel Name"), hdr.index("Grid Size")
    groups = defaultdict(list)
    for r in data:
        if len(r) <= ik:
            continue
        rec = {}
        for key, metric in WANT.items():
            if metric in hdr:
                j = hdr.index(metric)
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    continue
                u = units[j]
                if key.endswith("_bytes") or key == "duration_us":
                    v *= UNIT_SCALE.get(u, 1.0)
                if key == "sm_mhz":
                    v *= {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
                rec[key] = v
        rec["grid"] = r[ig]
        groups[short(r[ik])].append(rec)
    out = OrderedDict()
    for n, recs in sorted(groups.items(), key=lambda kv: -sum(x.get("duration_us", 0) for x in kv[1])):
        m = dict(launches=len(recs))
        for key in list(WANT) :
            vals = [x[key] for x in recs if key in x]
            if vals:
                m[key] = round(sum(vals) / len(vals), 3)
        m["dram_bytes_per_launch"] = round(m.get("dram_read_bytes", 0) + m.get("dram_write_bytes", 0))
        m["total_us"] = round(sum(x.get("duration_us", 0) for x in recs), 1)
        out[n] = m
    res = dict(tag=tag, note="ncu --set full --clock-control none, one steady-state forward+loss of "
               "`python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (ViT-L 518^2 B=32 bf16); averages per launch",
               kernels=out)
    with open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}.json"), "w") as f:
        json.dump(res, f, indent=1)
    for n, m in list(out.items())[:30]:
        print(f"  {m['total_us'] / 1e3:8.3f} ms x{m['launches']:3d} dram {m['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch "
              f"tensor {m.get('tensor_pct', 0):5.1f}% issue {m.get('issue_pct', 0):5.1f}% dram {m.get('dram_pct', 0):5.1f}%  {n}")


CLASSES = [("gemm_tc", ("gemm_tc", "conv_tc2")), ("attention", ("attention",)), ("layernorm", ("layernorm",)),
           ("elementwise", ("bilinear", "im2col", "patch_im2col", "head1x1")),
           ("loss", ("sel_", "mad_", "final_", "minmax", "init_minmax", "scale_", "ratio_", "sobel", "featcos", "l1_", "hyb_",
                     "set_den", "contexts_", "fz_"))]


def klass(name):
    for c, prefixes in CLASSES:
        if any(name.startswith(p) for p in prefixes):
            return c
    return None


def traffic(tag):
    """DRAM bytes per launch of each bench.py kernel class over one step = sum over the class's kernels of
    (launches in a step, from the launch list) x (average dram bytes per launch, from the --set full capture)."""
    lp = os.path.join(ROOT, "profiles", f"ncu_launches_{tag}.json")
    fp = os.path.join(ROOT, "profiles", f"ncu_full_{tag}.json")
    if not (os.path.exists(lp) and os.path.exists(fp)):
        return
    L, F = json.load(open(lp))["kernels"], json.load(open(fp))["kernels"]
    out = {}
    for c, _ in CLASSES:
        n = b = 0
        missing = []
        for k, a in L.items():
            if klass(k) != c:
                continue
            if k not in F:
                missing.append(k)
                continue
            n += a["launches"]
            b += a["launches"] * F[k]["dram_bytes_per_launch"]
        if n:
            out[c] = dict(launches_per_step=n, dram_bytes_per_step=b, dram_bytes_per_launch=b / n, not_captured=missing)
    res = dict(tag=tag, source=f"profiles/ncu_full_{tag}.json x profiles/ncu_launches_{tag}.json",
               note="dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the launches of one step",
               classes=out)
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    for c, v in out.items():
        print(f"  traffic {c:12s} {v['dram_bytes_per_launch'] / 1e6:9.1f} MB/launch x {v['launches_per_step']}")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    launches(tag)
    full(tag)
    traffic(tag)
