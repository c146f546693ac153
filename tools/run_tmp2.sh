timeout 120 python bench.py --no-cpu-baseline --no-train-block --steps 5 2>&1 | tail -12 | cut -c1-300
