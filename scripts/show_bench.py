import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    print(f, d["ms_per_step"], d["gpu_launches"], d.get("error"))
    for k in d["roofline"]["kernels"]:
        print("   %-60s %.4f ms x %.0f frac %.3f" % (k["kernel"][:60], k["launch_ms"], k["launches_per_step"], k["frac"]))
