"""Group the per-launch timings dumped by `bench.py --dump-launches` by layer shape."""
import json, sys
from collections import defaultdict
d = json.load(open(sys.argv[1]))
g = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in d:
    if r["kind"] == "tapgemm":
        key = ("tapgemm", r["n"], r["oh"], r["ow"], r["rows"], r["taps"], r["chunks"], r["phases"], r["bn"], tuple(r["tile"]))
    else:
        key = ("wgrad", r["n"], r["oh"], r["ow"], r["m"], r["ncols"], r["taps"], r["bn"])
    e = g[key]
    e[0] += 1; e[1] += r["ms"]; e[2] += r["alg_gflop"]; e[3] += r.get("exec_gflop", 0.0)
tot = sum(e[1] for e in g.values())
print("total tensor-kernel ms %.2f" % tot)
for k, e in sorted(g.items(), key=lambda kv: -kv[1][1]):
    print("%-70s x%-3d %7.2f ms (%4.1f%%)  alg %6.1f TF/s  exec %6.1f TF/s" % (str(k), e[0], e[1], 100 * e[1] / tot, e[2] / e[1], e[3] / e[1] if e[3] else 0))
