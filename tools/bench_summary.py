"""Prints a compact per-kernel table from a bench.py JSON line (file path argument)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f} tok/s  {d['ms_per_step']:.2f} ms/step  launches {d['gpu_launches']}")
for k in d["kernels"]:
    print(f"{k['name']:28s} n={k['launches']:4d} ms={k['ms']:8.3f} avg_us={k['avg_us']:8.2f} tops={str(k.get('tops')):>8s} gbs={k['gbs']:8.1f} share={k['share']:.3f}")
