cd /root/repo
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests/test_simt_gpu.py tests/test_networks_gpu.py tests/test_heads_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/tests.log
MUNIT_NORM_STREAM=0 timeout 900 python -m pytest tests/test_simt_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > $O/tests0.log 2>&1; echo "tests(stream=0) rc=$?"; tail -1 $O/tests0.log
for st in 0 1; do
echo "== MUNIT_NORM_STREAM=$st"
MUNIT_NORM_STREAM=$st timeout 300 python tools/bench_norm.py $O/bench_norm_$st.json 2>&1 | grep -v Warn | sed 's/stats=[0-9.]* finalize=[0-9.]* apply=[0-9.]* fwd(3)=[0-9.]* //'
done
for st in 0 1 0 1; do
MUNIT_NORM_STREAM=$st timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$st.json 2> $O/bench.err; echo "stream=$st rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_$st.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms", round(d["ms_per_step"],3), d["e2e"]["last_losses"])
h=d["roofline_hbm"]; print("   hbm", round(h["achieved"]), round(h["frac"],3), round(h["kernel_ms_per_step"],2), {k: round(v["ms"],2) for k,v in h["per_kernel"].items()})
PY
done
