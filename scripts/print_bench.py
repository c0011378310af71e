import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(d["value"], d["e2e"]["value"], d["clocks"], r["kernel_ms_per_step"], r.get("executed_frac"))
