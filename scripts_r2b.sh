mkdir -p gpurun_out/r2g
cd /root/repo
echo "== default"; python tools/bench_norm.py 2>&1 | tail -8 | cut -c1-110
echo "== apply U2"; MUNIT_LIB=/root/repo/munit_b200/csrc/variants/lib_au2.so python tools/bench_norm.py 2>&1 | tail -8 | cut -c1-60
echo "== apply U8"; MUNIT_LIB=/root/repo/munit_b200/csrc/variants/lib_au8.so python tools/bench_norm.py 2>&1 | tail -8 | cut -c1-60
for lb in 1 0 1 0; do
MUNIT_NORM_LASTBLOCK=$lb python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g/bench_lb$lb.json 2> gpurun_out/r2g/bench.err; echo "lastblock=$lb rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2g/bench_lb$lb.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d["gpu_launches"]/10)
h=d["roofline_hbm"]; print("hbm", round(h["achieved"]), round(h["frac"],3), round(h["kernel_ms_per_step"],2), {k: round(v["ms"],2) for k,v in h["per_kernel"].items()})
PY
done
