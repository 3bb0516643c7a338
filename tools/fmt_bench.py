import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception:
        print(l.strip()); continue
    if "skipped" in d:
        print("%-8s %-14s skipped: %s" % (d["config"], d.get("mode", ""), d["skipped"][:90])); continue
    print("%-8s %-14s cold %.1f min %.1f warm %.1f us  rk4 %s us/step  GB/s cold %.0f  parity %s" % (
        d["config"], d.get("mode", ""), d["rhs_us_cold_mean"], d["rhs_us_cold_min"], d["rhs_us_warm"],
        ("%.1f" % d["rk4_us_per_step"]) if "rk4_us_per_step" in d else "-", d["GBs_cold"], d.get("parity", "-")))
