for kb in 12 24; do for pdl in 1; do echo "== item_kb=$kb"; FF_RANGE_ITEM_KB=$kb timeout 200 python tools/bench_variants.py --only C2:None,C3:None,C2/16:None,C2/8:None --reps 7 2>&1 | grep -o '"config": "C[0-9/]*".*"whole_path_ms": [0-9.]*' | sed 's/"frames.*"stream_kernel_ms"/ kernel_ms/; s/"stream_kernel_gbs.*"stream_kernel_frac_of_measured_peak"/frac/'; FF_RANGE_ITEM_KB=$kb timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-pageable --no-head --legs c3_strong --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench C2 ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'C3 ms', round(d['c3_strong']['ms_per_step'],4))"; done; done
