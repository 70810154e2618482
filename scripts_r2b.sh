cd /root/repo
O=gpurun_out/r2w
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench$i.json 2> $O/bench.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench$i.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms", round(d["ms_per_step"],3), "launches", d.get("gpu_launches"), d["e2e"]["last_losses"])
PY
done
