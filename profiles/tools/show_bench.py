"""python profiles/tools/show_bench.py FILE...: the few numbers of a bench.py JSON line one looks at first."""
import json, sys
for f in sys.argv[1:]:
    b = json.load(open(f))
    print(f, "ms/step %.4f" % b["ms_per_step"], b.get("step_ms"), "e2e ms", (b.get("e2e") or {}).get("ms_per_step"),
          "whole-path frac", round((b.get("job") or {}).get("whole_path_frac_of_hbm_peak", 0), 4),
          "roofline", round((b.get("roofline") or {}).get("frac", 0), 4))
    print("   ", b.get("entry_point_ms"))
