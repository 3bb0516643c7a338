import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.strip()); continue
    print(d["config"], "cold %.1f min %.1f warm %.1f us  rk4 %.1f us/step  GB/s cold %.0f" % (d["rhs_us_cold_mean"], d["rhs_us_cold_min"], d["rhs_us_warm"], d["rk4_us_per_step"], d["GBs_cold"]))
