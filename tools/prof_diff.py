"""Compare two bench.py --profile-json files: per-category and per-entry-point times (ms)."""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2]))
print("ms/step", a["ms_per_step"], b["ms_per_step"])
ea, eb = a["by_entry_point"], b["by_entry_point"]
def cats(e):
    c = {}
    for k, v in e.items():
        n = k.split("[")[0] + ("[fc]" if "[fc]" in k else "")
        c[n] = c.get(n, 0.0) + v["ms"]
    return c
ca, cb = cats(ea), cats(eb)
for k in sorted(set(ca) | set(cb), key=lambda k: -cb.get(k, 0)):
    print(f"{k:28s} {ca.get(k, 0):8.2f} {cb.get(k, 0):8.2f}")
if len(sys.argv) > 3:
    for k in sorted(set(ea) | set(eb), key=lambda k: -eb.get(k, {}).get("ms", 0)):
        if sys.argv[3] in k:
            print(f"{k:52s} {ea.get(k, {}).get('ms', 0):7.2f} {eb.get(k, {}).get('ms', 0):7.2f}")
