for vp in 8 16 24 32; do
  python bench.py --config c5 --steps 3 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=$vp 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1M   voices_pieces $vp', d['value'], d['ms_per_step'])"
  python bench.py --config c5 --voices 131072 --steps 5 --e2e-steps 0 --no-cpu-baseline --plan-opt voices_pieces=$vp 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('131k voices_pieces $vp', d['value'], d['ms_per_step'])"
done
