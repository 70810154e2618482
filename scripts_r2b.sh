mkdir -p gpurun_out/r2i
cd /root/repo
timeout 1500 python -m pytest tests/test_inference_gpu.py tests/test_simt_gpu.py tests/test_train_script_gpu.py -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/r2i/tests.log 2>&1; echo "tests exit=$?"; tail -3 gpurun_out/r2i/tests.log
grep -h "inference scripts\|Error\|assert " gpurun_out/r2i/tests.log | cut -c1-700 | head
timeout 600 python tools/bench_loader.py 256 16 gpurun_out/r2i/loader.json 2>&1 | grep -v Warning | tail -4
