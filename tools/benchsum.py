import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "roofline frac", d["roofline"]["frac"] if d.get("roofline") else None)
for k,v in d["kernels"].items(): print(f"  {k:28s} {v['us_per_image']:.3f} us/img  share {v['share']:.3f}")
for k in ("hamming","cpu_baseline","clocks"):
    if k in d: print(k, d[k])
