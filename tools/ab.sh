for cfg in "4 4" "16 8" "16 4" "8 8" "16 16" "32 8" "4 4"; do set -- $cfg
  MUNIT_RSPLIT=$1 MUNIT_ASPLIT=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('AB rsplit=$1 asplit=$2 ms/step %.2f' % d['ms_per_step'])"
done
