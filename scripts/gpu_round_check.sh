#!/bin/bash
# The round's GPU check: every -m gpu test, smoke(), then the default bench line and the reference arm.  Output under gpurun_out/.
set -u
O=gpurun_out
TAG=${1:-r2c}
python -m pytest tests -m gpu -q -s > $O/${TAG}_gpu_tests_full.txt 2>&1; echo "gpu tests rc=$?"
grep -h "passed\|failed\|^FAILED" $O/${TAG}_gpu_tests_full.txt | tail -8
python __graft_entry__.py --smoke > $O/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/${TAG}_smoke.txt
python bench.py ${BENCH_ARGS:-} > $O/${TAG}_bench_c3_n1.json 2> $O/${TAG}_bench_c3_n1.err; echo "bench rc=$?"
python - <<P
import json
d = json.loads(open('$O/${TAG}_bench_c3_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('by_role'), d['cpu_baseline'], d['clocks'])
print({k: round(v['frac'], 3) for k, v in d['hbm_kernels'].items()}, {k: round(v['share_of_step'], 4) for k, v in d['hbm_kernels'].items()})
print(d['message_rows']['rescore'], d['host'])
P
if [ -n "${WITH_C5:-}" ]; then
    python bench.py --config c5 --no-cpu-baseline > $O/${TAG}_bench_c5_n1.json 2> $O/${TAG}_bench_c5_n1.err; echo "bench c5 rc=$?"
    python - <<P
import json
d = json.loads(open('$O/${TAG}_bench_c5_n1.json').read().strip().splitlines()[-1])
print('c5', d['value'], d['e2e'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('by_role'))
P
fi
