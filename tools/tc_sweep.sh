for d in 0 1500 3000 4500 6000; do
  GMP_TC_PHASE_DELAY=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('delay $d', d['ms_per_step'], d['phases_ms_per_step']['edge_feature'])"
done
