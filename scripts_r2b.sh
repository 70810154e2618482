mkdir -p gpurun_out/r2b
cd /root/repo
for f in tests/test_simt_gpu.py tests/test_networks_gpu.py tests/test_conv_gpu.py tests/test_heads_gpu.py tests/test_trainer_gpu.py tests/test_engine_gpu.py tests/test_train_script_gpu.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2b/$b.log 2>&1
  echo "$b exit=$?"; tail -3 gpurun_out/r2b/$b.log
done
grep -h "^block" gpurun_out/r2b/test_networks_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["last_losses"], d["roofline"]["achieved"], d["roofline"]["kernel_ms_per_step"])
PY
