import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(d["config"]["workload"][:30], round(d["value"]), round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["sample"])
