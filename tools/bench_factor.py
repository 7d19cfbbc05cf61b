import json, sys
d = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "", round(d["value"], 3), "factor", round(d["roofline"]["stages_ms"]["factor"], 3), "dx", d["check"]["dx_norm"])
