#!/bin/bash
# Key counters of every `ncu --set full` report of a round, one CSV per report (run where ncu is installed; no GPU needed):
#   bash scripts/ncu_full_summary.sh <tag> [outdir]    reads gpurun_out/<tag>_*.ncu-rep, writes <outdir>/<tag>_ncu_full_<name>.csv
TAG=${1:-r2}
DST=${2:-profiles}
PAT='Kernel Name|gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem_per_block|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum |lts__throughput.avg.pct|l1tex__throughput.avg.pct|smsp__cycles_active.avg|sm__cycles_elapsed.avg |smsp__inst_executed.sum |launch__occupancy_limit|sm__cycles_active.avg |gpc__cycles_elapsed.max'
for f in gpurun_out/${TAG}_*.ncu-rep; do
    n=$(basename $f .ncu-rep); n=${n#${TAG}_}
    ncu -i $f --page raw --csv 2>/dev/null | python3 -c "
import csv, sys, re
rows = list(csv.reader(sys.stdin))
if len(rows) < 3: sys.exit(0)
hdr, units = rows[0], rows[1]
pat = re.compile(r'$PAT')
keep = [i for i, h in enumerate(hdr) if pat.search(h)]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] + (' [' + units[i] + ']' if units[i] else '') for i in keep])
for r in rows[2:]:
    w.writerow([r[i] for i in keep])
" > $DST/${TAG}_ncu_full_${n}.csv
    echo $DST/${TAG}_ncu_full_${n}.csv $(wc -l < $DST/${TAG}_ncu_full_${n}.csv)
done
