for d in 0 4 7; do echo "== dbg $d"; MLBP_K3_DBG=$d timeout 200 python scripts/k3_probe.py --reps 3 2>&1 | grep -v "^{" ; done
