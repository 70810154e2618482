cd /root/repo
O=gpurun_out/r2y
mkdir -p $O
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_trainer_gpu.py tests/test_train_script_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
for ov in 0 1 0 1; do
MUNIT_OVERLAP_UPDATES=$ov timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$ov.json 2> $O/bench_$ov.err; echo "overlap_updates=$ov rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_$ov.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d["e2e"]["last_losses"])
PY
done
