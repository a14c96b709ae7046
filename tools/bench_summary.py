import sys,json
d=json.loads(sys.stdin.read()); r=d["roofline"]
print("captions/s %.0f  ms/step %.3f  fwd us/step %.1f  bwd us/step %.1f  e2e %.0f launches %d" % (d["value"], d["ms_per_step"], r["us_per_step"], r["bwd_us_per_step"], d["e2e"]["value"], d["gpu_launches"]))
