import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    k = {a: round(b * 1e3, 1) for a, b in d.get("kernels_ms", {}).items()}
    print(d["config"]["scene"][:6], d.get("impl", "ours"), round(d["value"], 1), "Mrays/s", round(d["ms_per_step"], 3), "ms | e2e",
          round(d["e2e"]["value"], 1), d["scene_stats"], k)
