# round 2: pageable caller buffers staged through page-locked memory by one host thread per pipeline slot — parity at full size, then the two e2e legs
set -x
python -m pytest tests/test_gpu_full_size.py tests/test_gpu_trace.py tests/test_gpu_boundary.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload trace --steps 10 --no-cpu-baseline --no-secondary 2>/dev/null > gpurun_out/r2o_trace_staged.json
ECHO_B200_STAGE_PAGEABLE=0 python bench.py --workload trace --steps 10 --no-cpu-baseline --no-secondary 2>/dev/null > gpurun_out/r2o_trace_unstaged.json
python - <<'PY'
import json
for k in ('staged', 'unstaged'):
    d = json.load(open(f'gpurun_out/r2o_trace_{k}.json'))
    print(k, 'value', round(d['value']), 'e2e pinned', round(d['e2e']['value']), 'e2e pageable', round(d['e2e_pageable']['value']), round(d['e2e_pageable']['ms_per_step'], 1), 'ms')
PY
