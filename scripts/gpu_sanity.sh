#!/bin/bash
# what the driver runs at round end on one GPU: pytest -m gpu, smoke, a short bench
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])"
