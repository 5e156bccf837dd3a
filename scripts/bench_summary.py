"""One-screen summary of bench.py JSON lines: python scripts/bench_summary.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    e2e = d.get("e2e") or {}
    print(f"{path}: {d['ms_per_step']:.3f} ms/step  {d['value'] / 1e6:.3f} M{d['unit']}  e2e {e2e.get('value', 0) / 1e6:.3f}  "
          f"launches {d.get('gpu_launches')}  n_gpus {d.get('n_gpus')}  mode {d.get('mlp_mode')}")
    for k in d.get("kernels", [])[:14]:
        print(f"    {k['kernel']:28s} {k['ms_per_step']:7.3f} ms  x{k.get('launches_per_step', 0):g}  frac {k.get('frac', 0):.3f}")
