import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],1), {k:round(v["ms_per_step"],1) if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
