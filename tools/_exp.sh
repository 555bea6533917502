run() { timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-pageable --no-head --legs c3_strong --e2e-steps 1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C2 ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'C3 ms', round(d['c3_strong']['ms_per_step'],4), 'C3 kernel', round(d['c3_strong']['range_kernel_ms_slowest_rank'],4))"; }
echo "== default"; run
for st in 3 6; do echo "== stages=$st"; FF_COUNT12_STAGES=$st run; done
for ct in 1 3; do echo "== ctas=$ct stages 6"; FF_COUNT12_CTAS=$ct FF_COUNT12_STAGES=6 run; done
echo "== item 24KB x3"; FF_RANGE_ITEM_KB=24 run
