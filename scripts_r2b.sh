cd /root/repo
O=gpurun_out/r2x
mkdir -p $O
for cfg in "0 0" "1 0" "0 6" "1 6" "0 4" "0 3" "0 8" "0 16"; do set -- $cfg
echo "== WG_DBG=$1 KSPLIT=$2"
MUNIT_WG_DBG=$1 KSPLIT=$2 python tools/bench_conv.py res3x3 down2_4x4s2 dec4_5x5 2>&1 | grep -v Warn | sed "s/'fwd': ([0-9., ]*), 'dgrad': ([0-9., ]*), //"
done
