"""Markdown tables of the degree/size sweep (scripts/sweep.py output) for BASELINE.md section 6.2.
usage: python scripts/make_sweep_tables.py profiles/sweep_r1_final.jsonl"""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
for quad, title in (("gll", "GLL collocation"), ("gauss", "QGauss(p+1)")):
    print(f"**{title}** — merged CG GDoF·it/s (vmult fraction of the measured HBM peak)\n")
    print("| p | ~1 M | ~4 M | ~16 M | ~64 M | ~200 M DoFs |\n|---|---|---|---|---|---|")
    for p in range(2, 9):
        rs = sorted([r for r in rows if r["p"] == p and r["quad"] == quad], key=lambda r: r["dofs"])
        print(f"| {p} | " + " | ".join(f"{r['cg_gdofs']:.1f} ({r['vmult_frac']:.2f})" for r in rs) + " |")
    print()
