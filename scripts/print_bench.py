import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])   # NCCL may print its version line first
r=d["roofline"]
print(d["value"], d["e2e"]["value"], d["clocks"], r.get("kernel_ms_per_step"), r.get("executed_frac"))
