timeout 200 python scripts/k3_probe.py --reps 3 2>&1 | grep -v '^{"k3'
