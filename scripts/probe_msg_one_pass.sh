#!/bin/bash
# One-pass message rows (A_hi . B_hi + spike compensation of both dropped terms) against the default two-pass rows:
# hostile-data gates, C3 parity at full size, and the bench A/B.  Output under gpurun_out/ (copied to profiles/ by hand).
set -u
O=gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -k "spike" > $O/p1_kernels.txt 2>&1; echo "kernels rc=$?"
python -m pytest tests/test_gpu_gates.py -s -q > $O/p1_gates_default.txt 2>&1; echo "gates default rc=$?"
MLBP_MSG_PASSES=1 python -m pytest tests/test_gpu_gates.py tests/test_gpu_engine.py -s -q > $O/p1_gates_one_pass.txt 2>&1; echo "gates+engine one-pass rc=$?"
python scripts/c3_parity_check.py --n 12 > $O/p1_c3_parity_default.json 2> $O/p1_c3_parity_default.err; echo "parity default rc=$?"
python scripts/c3_parity_check.py --n 12 --msg-passes 1 > $O/p1_c3_parity_one_pass.json 2> $O/p1_c3_parity_one_pass.err; echo "parity one-pass rc=$?"
python bench.py --warmup 6 --steps 3 --no-cpu-baseline --no-e2e > $O/p1_bench_default.json 2> $O/p1_bench_default.err; echo "bench default rc=$?"
python bench.py --warmup 6 --steps 3 --no-cpu-baseline --no-e2e --msg-passes 1 > $O/p1_bench_one_pass.json 2> $O/p1_bench_one_pass.err; echo "bench one-pass rc=$?"
grep -h "worst\|passed\|failed" $O/p1_gates_default.txt $O/p1_gates_one_pass.txt | tail -30
cat $O/p1_c3_parity_default.json $O/p1_c3_parity_one_pass.json
python - <<'P'
import json
for n in ('default', 'one_pass'):
    try:
        d = json.loads(open('gpurun_out/p1_bench_%s.json' % n).read().strip().splitlines()[-1])
        print(n, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['by_passes'], d.get('message_rows', {}).get('rescore'))
    except Exception as e:
        print(n, 'failed', e)
P
