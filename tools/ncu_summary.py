#!/usr/bin/env python
"""Turn ncu output brought back in gpurun_out/ into the small, reviewable summaries kept under profiles/.

    # launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv ...)
    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/launches_r2_summary.md

    # one full capture: first export the pages here (no GPU needed), then summarise
    ncu -i gpurun_out/full_gemm_tc_kernel.ncu-rep --page raw --csv > /tmp/raw.csv
    ncu -i gpurun_out/full_gemm_tc_kernel.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_summary.py raw /tmp/raw.csv > profiles/ncu_r2_gemm_raw.csv          # metric,unit,value per kernel
    python tools/ncu_summary.py source /tmp/src.csv > profiles/ncu_r2_gemm_hotspots.md   # top stall-sampled SASS

Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes (the absolute
numbers of bench.py come from CUDA events without a profiler attached).
"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

# metrics worth keeping from a `--set full` raw page (substring match on the column name)
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "gpu__dram_throughput.avg.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "lts__average_t_sector_hit_rate", "sm__throughput.avg.pct", "sm__pipe_tensor_cycles_active.avg.pct",
        "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__cluster", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts", "smsp__warp_issue_stalled")


def _rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]          # ncu banner lines
    return list(csv.reader(lines))


def short_kernel(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)        # drop the argument list
    name = re.sub(r"\((int|bool)\)", "", name)
    return name


def launches(path: str) -> None:
    rows = _rows(path)
    head = rows[0]
    ik, im, iv = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    iu = head.index("Metric Unit") if "Metric Unit" in head else None
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu] if iu is not None else "ns"
        ms = v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3,
                  "second": 1e3}.get(unit, 1e-6)
        k = short_kernel(r[ik])
        agg[k][0] += 1
        agg[k][1] += ms
        total += ms
        n += 1
    ours = sum(v[1] for k, v in agg.items() if k.startswith("v3d::"))
    print(f"launches captured: {n}; total kernel time {total:.1f} ms; v3d:: kernels {100 * ours / max(total, 1e-9):.1f}% of it\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k[:100]}` | {cnt} | {ms:.2f} | {100 * ms / max(total, 1e-9):.1f}% |")


def raw(path: str) -> None:
    rows = _rows(path)
    head, units = rows[0], rows[1]
    ik = head.index("Kernel Name")
    keep = [i for i, h in enumerate(head) if any(s in h for s in KEEP)]
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "unit", "value"])
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        w.writerow(["kernel", "", short_kernel(r[ik])])
        seen = OrderedDict()
        for i in keep:
            name = head[i].split(".", 2)[-1] if head[i].count(".") > 3 and head[i].split(".")[1].startswith("Triage") else head[i]
            if i < len(r) and r[i] != "" and name not in seen:
                seen[name] = (units[i] if i < len(units) else "", r[i])
        for name, (u, v) in sorted(seen.items()):
            w.writerow([name, u, v])


def source(path: str, top: int = 25) -> None:
    rows = _rows(path)
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")      # skip the "Kernel Name" banner rows
    if hi > 0 and rows[0] and rows[0][0] == "Kernel Name":
        print(f"kernel: `{short_kernel(rows[0][1])}`\n")
    head = rows[hi]

    def col(*subs):
        for i, h in enumerate(head):
            if all(s.lower() in h.lower() for s in subs):
                return i
        return None

    isrc = col("source")
    isamp = col("warp stall sampling", "all") or col("sampling")
    iexec = col("instructions executed") or col("inst", "executed")
    if isrc is None or isamp is None:
        raise SystemExit(f"unexpected source-page columns: {head[:12]}")
    reasons = [(i, h) for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
    data, agg = [], defaultdict(int)
    seen = set()
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x") or r[0] in seen:       # a second launch of the kernel repeats the listing
            continue
        seen.add(r[0])
        try:
            n = int(float(r[isamp] or 0))
        except (ValueError, IndexError):
            continue
        why = []
        for i, h in reasons:
            try:
                v = int(float(r[i] or 0))
            except (ValueError, IndexError):
                v = 0
            agg[h] += v
            if v:
                why.append((v, h))
        data.append((n, r[isrc].strip(), r[iexec] if iexec is not None else "", sorted(why, reverse=True)[:2]))
    total = sum(d[0] for d in data)
    print(f"warp-stall samples: {total}\n")
    print("stall reasons over the whole kernel: " +
          ", ".join(f"{h[6:]} {100 * v / max(total, 1):.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v) + "\n")
    print("| samples | share | executed | SASS | main reasons |\n|---|---|---|---|---|")
    for n, src, ex, why in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"| {n} | {100 * n / max(total, 1):.1f}% | {ex} | `{src}` | {', '.join(f'{h[6:]} {v}' for v, h in why)} |")


def traffic(path: str) -> None:
    """`roofline.traffic` of bench.py: DRAM bytes (read + write) per launch of every captured kernel of a `--set full`
    raw page, stamped with the digest of the kernel sources the capture was taken from (bench.py drops the object when
    the sources have changed since).  python tools/ncu_summary.py traffic /tmp/raw.csv > profiles/traffic_r2.json"""
    import json
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench

    rows = _rows(path)
    head, units = rows[0], rows[1]
    ik = head.index("Kernel Name")
    col = {h: i for i, h in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, name):
        i = col[name]
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)

    out = {"csrc_digest": bench.csrc_digest(), "source": "ncu --set full --clock-control none (tools/microbench.py "
           "ncu_set: every kernel family at its V3D_512 top-level shape, second launch of each)", "kernels": {}}
    seen = {}
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        k = short_kernel(r[ik])
        seen[k] = seen.get(k, 0) + 1
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        out["kernels"][f"{k} #{seen[k]}"] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
                                            "duration_" + units[col["gpu__time_duration.sum"]]: dur}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("launches", "raw", "source", "traffic"):
        raise SystemExit(__doc__)
    {"launches": launches, "raw": raw, "source": source, "traffic": traffic}[sys.argv[1]](sys.argv[2])
