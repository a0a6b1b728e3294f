#!/usr/bin/env python
"""BASELINE.json configs[4]: sampler-step x frame-count sweep of the hot path on one GPU (or N with torchrun-less
image-parallel runs left to bench.py itself).  Runs bench.py once per (T, S), collects the JSON lines, prints a table
and writes gpurun_out/sweep.json (copy the file to profiles/ to have it judged).

    python tools/sweep.py                       # S in {10, 25, 50} x T in {14, 18, 25}
    python tools/sweep.py --frames 18 24 --edm-steps 25 --extra --shard views      # any bench.py flags after --extra
"""
import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="+", default=[14, 18, 25])
    ap.add_argument("--edm-steps", type=int, nargs="+", default=[10, 25, 50])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--extra", nargs=argparse.REMAINDER, default=[])
    args = ap.parse_args()
    rows = []
    for T in args.frames:
        for S in args.edm_steps:
            cmd = [sys.executable, str(ROOT / "bench.py"), "--frames", str(T), "--edm-steps", str(S), "--steps",
                   str(args.steps), "--warmup", str(args.warmup), "--no-cpu-baseline", *args.extra]
            t0 = time.time()
            res = subprocess.run(cmd, capture_output=True, text=True)
            line = next((l for l in reversed(res.stdout.splitlines()) if l.startswith("{")), None)
            if res.returncode != 0 or line is None:
                rows.append({"frames": T, "edm_steps": S, "error": (res.stderr or res.stdout)[-500:]})
                print(f"T={T} S={S}: FAILED", file=sys.stderr)
                continue
            d = json.loads(line)
            rows.append({"frames": T, "edm_steps": S, "value": d["value"], "unit": d["unit"],
                         "ms_per_image": d["ms_per_step"], "e2e": d["e2e"]["value"],
                         "gemm_roofline_frac": d["roofline"]["frac"], "model_frac": d["roofline"]["model"]["frac"],
                         "clocks": d.get("clocks"), "wall_s": round(time.time() - t0, 1)})
            print(f"T={T:2d} S={S:2d}: {d['value']:.2f} {d['unit']}  {d['ms_per_step']:.0f} ms/image  "
                  f"gemm {100 * d['roofline']['frac']:.1f}%  model {100 * d['roofline']['model']['frac']:.1f}%")
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "sweep.json").write_text(json.dumps(rows, indent=1))
    print("wrote", out / "sweep.json")


if __name__ == "__main__":
    main()
