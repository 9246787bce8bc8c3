import json, sys
for f in sys.argv[1:]:
    try:
        txt=open(f).read().strip().splitlines()
        d=json.loads(txt[-1]); r=d["roofline"]
    except Exception as e:
        print(f, "ERR", e, txt[-3:] if txt else ""); continue
    print(f, "n=%d value %.0f ms/step %.4f e2e %.0f launches %d"%(d["n_gpus"],d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], (d.get("cpu_baseline") or {}).get("value"))
    print("  dom", r["kernel"], "%.0f GB/s frac %.3f"%(r["achieved"], r["frac"]), "whole %.3f"%r["whole_step"]["frac"], "serial ms %.4f"%r["serialised_ms_per_step"])
    for k in r["kernels"]: print("    %-40s n=%d ms/launch=%.4f GB/s=%.0f"%(k["name"],k["launches"],k["ms"]/k["launches"], k["bytes"]/(k["ms"]/k["launches"]*1e-3)/1e9 if k["ms"] else 0))
