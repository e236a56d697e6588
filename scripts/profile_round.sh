#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU):  bash scripts/profile_round.sh <tag>
#   1. the plain command must exit 0 first;
#   2. launch list (device time + DRAM bytes of every launch of one small step; cold-cache, serialised: compare SHARES);
#   3. `--set full` of the K4 CTA-pair launches (message rows, then the gradient rows at the step's end), of four resident-K3
#      launches and of one launch each of K1, K2, K5, K5b, K6a, K4b.
# Summaries: scripts/ncu_summary.py (launch list) and scripts/ncu_full_summary.sh (`ncu -i ... --page raw --csv`).
TAG=${1:-r2}
OUT=gpurun_out
# six warm-up SGD steps first: theta has then drifted into the regime of the long bench (peaked beliefs, spikes compensated)
CMD="python bench.py --sentences 128 --steps 1 --warmup ${WARM:-6} --no-cpu-baseline --no-e2e"
mkdir -p $OUT
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed" >> $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_ncu_launches_128sent.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
full() {   # name, kernel regex, skip, count -> <tag>_<name>.ncu-rep + the key counters of every profiled launch as CSV
    ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $OUT/${TAG}_$1 $CMD > $OUT/${TAG}_ncu_$1.log 2>&1
    tail -3 $OUT/${TAG}_ncu_$1.log > $OUT/${TAG}_ncu_$1.tail; mv $OUT/${TAG}_ncu_$1.tail $OUT/${TAG}_ncu_$1.log
}
# the last launches of the program are the profiling pass of bench.py: its final 48 pair-kernel launches cover the last levels'
# message rows (gated: every second launch returns at once) and the gradient rows (the last 24 launches)
NPAIR=$(( $(grep -c gemm_split_f16_pair $OUT/${TAG}_ncu_launches_128sent.csv) / 3 ))
full gemm gemm_split_f16_pair ${GEMM_SKIP:-$(( NPAIR > 24 ? NPAIR - 24 : 0 ))} ${GEMM_COUNT:-24}
NK3=$(( $(grep -c var_to_factor_resident $OUT/${TAG}_ncu_launches_128sent.csv) / 3 ))
[ -n "$SKIP_K3" ] || full k3 var_to_factor_resident $(( NK3 > 6 ? NK3 - 6 : 0 )) 6
if [ -z "$SKIP_SMALL" ]; then
    full k1 unary_products_kernel ${WARM:-6} 1
    full k2 build_pairwise_tables_kernel ${WARM:-6} 1
    full k5 marginals_kernel ${WARM:-6} 1
    full k5b rescore_kernel ${WARM:-6} 1
    full k6a pair_expectations_kernel ${WARM:-6} 1
    full k4a spike_scan_kernel 200 2
    NSP=$(( $(grep -c spike_correct_kernel $OUT/${TAG}_ncu_launches_128sent.csv) / 3 ))
    full k4b spike_correct_kernel $(( NSP > 4 ? NSP - 4 : 0 )) 4
fi
# what travels back is limited to 64 MiB: summarise every report on the box, keep only small reports
bash scripts/ncu_full_summary.sh ${TAG} $OUT > $OUT/${TAG}_summary.log 2>&1
ncu -i $OUT/${TAG}_k3.ncu-rep --page source --csv > $OUT/${TAG}_k3_source.csv 2>/dev/null
find $OUT -name "${TAG}_*.ncu-rep" -size +6M -delete
gzip -f $OUT/${TAG}_k3_source.csv $OUT/${TAG}_ncu_launches_128sent.csv 2>/dev/null
du -sh $OUT; ls -la $OUT | grep ${TAG}
