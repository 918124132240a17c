#!/usr/bin/env python
"""profiles/r2_scaling.md from the committed bench lines profiles/r2_bench_{1,2,4,8}gpu.json (weak headline + `strong` object)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(n):
    p = os.path.join(ROOT, "profiles", f"r2_bench_{n}gpu.json")
    if not os.path.exists(p):
        return None
    return json.loads([l for l in open(p) if l.startswith("{")][-1])


def main():
    d1 = line(1)
    rows = ["# Round 2 scaling (bench.py --steps 20 --warmup 5; weak = 19 960 hexahedra per GPU, strong = 19 960 in total; one 8 x B200 box)", "",
            "| GPUs | data plane | weak DOF*steps/s | ms/step | pressure its/step | weak efficiency | strong DOF*steps/s | strong ms/step | strong efficiency | parity_n | Arnoldi s/iteration (weak) |",
            "|---|---|---|---|---|---|---|---|---|---|---|"]
    for n in (1, 2, 4, 8):
        d = line(n)
        if d is None:
            continue
        st = d.get("strong") or {}
        weff = d["value"] / (n * d1["value"])
        seff = (st["value"] / (n * d1["value"])) if st else (1.0 if n == 1 else None)
        rows.append(f"| {n} | {d.get('data_plane', '-')} | {d['value']:.4g} | {d['ms_per_step']:.2f} | {d['config']['pres_iters_per_step']:.1f} | {weff:.3f} | "
                    f"{(st.get('value') or d['value']):.4g} | {(st.get('ms_per_step') or d['ms_per_step']):.2f} | {seff:.3f} | "
                    f"{d.get('parity_n') if d.get('parity_n') is not None else '-'} | {d.get('arnoldi', {}).get('wall_s_per_iteration', float('nan')):.2f} |")
    rows += ["", "Per pressure iteration (sampled, ms): " + "; ".join(
        f"N={n}: " + ", ".join(f"{k} {v['avg_ms']:.3f}" for k, v in line(n)["roofline"]["kernels"].items() if k.startswith("pcg") or k == "dssum")
        for n in (1, 2, 4, 8) if line(n))]
    with open(os.path.join(ROOT, "profiles", "r2_scaling.md"), "w") as f:
        f.write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
