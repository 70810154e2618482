cd /root/repo
O=gpurun_out/r3a
mkdir -p $O
N=8
T0=$(date +%s); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_dp$N.json 2> $O/bench_dp$N.err; echo "bench dp$N rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.loads(open("$O/bench_dp$N.json").read().strip().splitlines()[-1])
print("value", round(d["value"],3), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],3), d["engine"]["grad_exchange"][:100], d["clocks"])
for k in ("forward_reuse","global_batch_64","hd","infer"):
    if k in d:
        e=d[k]; print(k, {kk: (round(v,3) if isinstance(v,float) else v) for kk,v in e.items() if kk in ("ms_per_step","steps_per_s","pairs_per_s","value","error","n_gpus")})
PY
T0=$(date +%s); timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_1.json 2>/dev/null; echo "1 gpu wall $(( $(date +%s) - T0 )) s"
python - <<PY
import json
d=json.loads(open("$O/bench_1.json").read().strip().splitlines()[-1])
print("1 GPU value", round(d["value"],3), "ms", round(d["ms_per_step"],3))
PY
