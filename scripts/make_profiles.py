"""Condenses gpurun_out/ ncu captures + bench line into profiles/ (tracked).  Run in the build container after a gpurun call."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
bench_file = sys.argv[2] if len(sys.argv) > 2 else "r2_d_bench.json"
os.makedirs(P, exist_ok=True)
rows = [r for r in csv.reader(open(f"{G}/launches_{tag}.csv")) if len(r) > 5]
hdr = None; agg = collections.OrderedDict(); out = []
for r in rows:
    if r[0] == "ID": hdr = r; continue
    if not hdr: continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except ValueError: continue
    name = d["Kernel Name"].split("(")[0].replace("dcs::", "").replace("<unnamed>::", "").replace("void ", "")
    out.append((d["ID"], name, d["Grid Size"], d["Block Size"], v / 1e3))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v / 1e3
tot = sum(a[1] for a in agg.values())
with open(f"{P}/{tag}_launches.csv", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 900: python bench.py --steps 5 --warmup 3 --lm-iters 1 --no-cpu (B200)\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\nid,kernel,grid,block,us\n")
    for o in out: f.write("%s,%s,%s,%s,%.2f\n" % (o[0], o[1], o[2].replace(",", " "), o[3].replace(",", " "), o[4]))
summary = [f"| {k} | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |" for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
raw = subprocess.run(["ncu", "-i", f"{G}/prof_{tag}_final.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB"), ("dram__bytes_write.sum", "MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"), ("lts__t_sector_hit_rate.pct", "%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "inst")]
lines = []; seen = set(); traffic = None
for r in rr[2:]:
    d = dict(zip(h, r)); name = d["Kernel Name"].split("(")[0].replace("void ", "")
    if name in seen: continue
    seen.add(name)
    lines.append("| " + name + " | " + d.get("Grid Size", "") + " x " + d.get("Block Size", "") + " | " + " | ".join(d.get(w, "")[:10] for w, _ in want) + " |")
    if name == "k_linearize": traffic = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e6
json.dump({"kernel": "k_linearize", "dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": 552000000.0,
           "source": f"profiles/{tag}_kernels.md (ncu --set full of the shipped build: dram__bytes_read.sum + dram__bytes_write.sum, 1M poses / 4M edges)"},
          open(f"{P}/linearize_traffic.json", "w"), indent=1)
bench = json.loads(open(f"{G}/{bench_file}").read().strip().splitlines()[-1])
open(f"{P}/{tag}_bench.json", "w").write(json.dumps(bench) + "\n")
hdr_cols = " | ".join(w.split(".")[0].replace("gpu__", "").replace("sm__", "").replace("smsp__", "").replace("launch__", "") + " (" + u + ")" for w, u in want)
extra = open(f"{P}/{tag}_notes.md").read() if os.path.exists(f"{P}/{tag}_notes.md") else ""
lm = bench.get("lm") or {}
cb = bench.get("cpu_baseline") or {}
md = f"""# Round {tag[1:]} profiles (B200, sm_100a, CUDA 12.9, driver 580)

All numbers from `gpurun` boxes (one B200 unless stated), clocks untouched (`--clock-control none`; SM 1965 MHz, no throttle reasons).
Bench line, launch list and the `--set full` capture are from the SAME build (the one this file is committed with).
`compute-sanitizer` is closed on this pool (refused by gpurun); memory safety rests on the parity suite and on the
`-DDCS_CHECK` bounds-assert build (`tests/test_gpu_parity.py::test_bounds_asserts_build_on_the_odd_cases`).

## 1. Bench line (`python bench.py --steps 20 --warmup 5`, not under a profiler) — `profiles/{tag}_bench.json`

* `value` = {bench['value']:.4g} edges/s ({bench['ms_per_step'] * 1e3:.1f} us per step = fused eval+assembly launch + its fold kernel, 1 M poses / 4 M edges, inputs resident)
* `roofline.frac` = {bench['roofline']['frac']:.3f} of the measured HBM peak ({bench['roofline']['peak']} GB/s); algorithmic bytes 108 E + 120 N = 552 MB per launch, DRAM traffic {traffic / 1e6:.0f} MB (ncu, section 3)
* with the linear-solver setup (`k_expand`, once per LM iteration): {bench['config']['ms_per_step_with_solver_setup'] * 1e3:.1f} us
* `e2e` = {bench['e2e']['value']:.4g} edges/s ({bench['e2e']['ms_per_step']:.2f} ms per C-ABI call: page-locked host poses in (24 MB H2D), launch, scalar result out)
* `cpu_baseline` = {cb.get('value', 0):.4g} edges/s (oracle port, {cb.get('cores')} host threads); 1 thread: {(cb.get('threads_1') or {}).get('value', 0):.4g} edges/s
* full DCS-LM solve, 1 M poses, pcg_rel_tol 1e-12: {lm.get('lm_iterations')} LM iterations in {lm.get('seconds', 0):.1f} s = {lm.get('lm_iters_per_sec', 0):.2f} LM iterations/s; {lm.get('pcg_iterations')} PCG iterations ({(lm.get('pcg_iterations_per_step') or {}).get('min')}-{(lm.get('pcg_iterations_per_step') or {}).get('max')} per step) at {lm.get('us_per_pcg_iteration', 0):.1f} us; cost {lm.get('initial_cost', 0):.6g} -> {lm.get('final_cost', 0):.6g}; largest true residual |(H+L)w-g|/|g| over the 50 linear solves {lm.get('max_true_residual', 0):.2e}
* `dcs_create` (pattern build + uploads): {bench['config']['create_s']:.2f} s

## 2. Launch list — `profiles/{tag}_launches.csv`

`ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 5 --warmup 3 --lm-iters 1 --no-cpu`

| kernel | launches | total us | avg us | share |
|---|---|---|---|---|
""" + "\n".join(summary) + f"""

The bench step proper is `k_linearize` (+ `k_fold_tasks`), one pair per step.  The LM measurement is dominated by
`k_spmv` (one per PCG iteration), then the chain-preconditioned vector kernel `k_pcg_chain` and `k_pcg_direction`.

## 3. `ncu --set full` of the hot kernels — `scripts/prof_kernels.py` (1 M poses / 4 M edges)

| kernel | grid x block | {hdr_cols} |
|---|---|""" + "---|" * len(want) + "\n" + "\n".join(lines) + "\n\n" + extra
open(f"{P}/{tag}_kernels.md", "w").write(md)
print(md[:3000])
