#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU):  bash scripts/profile_round.sh <tag>
#   1. the plain command must exit 0 first;  2. launch list (device time + DRAM bytes of every launch);
#   3. `--set full` of the K4 CTA-pair launches of the timed step (message rows, then the gradient rows at its end);
#   4. `--set full` of four resident-K3 launches.  Summaries: scripts/ncu_summary.py / `ncu -i ... --page raw --csv`.
TAG=${1:-r1f}
OUT=gpurun_out
CMD="python bench.py --sentences 128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p $OUT
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed" >> $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_ncu_launches_128sent.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_split_f16_pair -s ${GEMM_SKIP:-58} -c ${GEMM_COUNT:-16} -f -o $OUT/${TAG}_gemm $CMD \
    > $OUT/${TAG}_ncu_gemm.log 2>&1
[ -n "$SKIP_K3" ] || ncu --set full --clock-control none --import-source on -k regex:var_to_factor_resident -s 9 -c 4 -f -o $OUT/${TAG}_k3 $CMD \
    > $OUT/${TAG}_ncu_k3.log 2>&1
ls -la $OUT | grep ${TAG}
