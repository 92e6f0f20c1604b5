timeout 900 python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -3
for W in newref_600x50kb newref_600x250kb; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-test --workload $W 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'], d['ms_per_step'], d['phases_ms'], d['roofline']['frac'])"
done
python tools/profile_k5.py newref_600x50kb 0 2>&1 | tail -1 | cut -c1-700
