"""Print the essentials of a bench.py JSON line (tools only): python tools/bench_summary.py FILE"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][0])
print(f"N={d['n_gpus']} value {d['value']:.4e} tests/s  {d['ms_per_step']:.3f} ms/step   e2e {d['e2e']['value']:.3e} "
      f"({d['e2e']['ms_per_step']:.2f} ms)  launches {d['gpu_launches']}  clocks {d.get('clocks')}")
r = d["roofline"]
print(f"roofline frac {r['frac']:.3f} achieved {r['achieved']:.0f} GB/s launch_ms {r['launch_ms']:.3f} other {r['other_kernels_ms_per_step']:.3f}")
for k in r.get("kernels", []):
    print(f"   {k['name']:12s} {k['ms']:.3f} ms frac {k['frac']:.3f}")
if "cpu_baseline" in d:
    print("cpu", d["cpu_baseline"])
if "parity" in d:
    print("parity", {k: (v if not isinstance(v, dict) else v["max_rel_err"]) for k, v in d["parity"].items()})
if "multi_gpu_check" in d:
    print("check", d["multi_gpu_check"])
for c in d.get("configs", []):
    extra = f" finish {c['finish_ms']:.3f} other {c['other_kernels_ms']:.3f} frac {c['roofline_frac']:.3f} parity {c['parity']['ok']}" if "parity" in c else ""
    print(f"   {c['name']:18s} {c['value']:.3e} tests/s  {c['ms_per_pass']:.3f} ms  kernel {c['kernel_ms']:.3f}{extra}")
for w in d.get("derand_wall", []):
    print("   wall", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in w.items() if k not in ("genes", "host_cores")})
