cd /root/repo
mkdir -p gpurun_out/r2m
for pr in 0 1 0 1; do
MUNIT_PAIR=$pr python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m/bench_pair$pr.json 2> gpurun_out/r2m/bench.err; echo "pair=$pr rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2m/bench_pair$pr.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms", round(d["ms_per_step"],3), d["e2e"]["last_losses"], "tensor", round(d["roofline"]["achieved"]), round(d["roofline"]["kernel_ms_per_step"],2), "wgrad", round(d["roofline"]["wgrad"]["kernel_ms_per_step"],2))
h=d["roofline_hbm"]; print("   hbm", round(h["achieved"]), round(h["frac"],3), round(h["kernel_ms_per_step"],2), {k: round(v["ms"],2) for k,v in h["per_kernel"].items()})
PY
done
MUNIT_PAIR=1 timeout 900 python -m pytest tests/test_networks_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
