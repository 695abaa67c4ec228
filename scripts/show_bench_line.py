import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "n_gpus", d["n_gpus"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, json.dumps(v)[:int(sys.argv[2]) if len(sys.argv) > 2 else 300])
if "rank_balance" in d: print(d["rank_balance"]); print(d["per_rank"])
