#!/bin/bash
# Round-end validation and evidence refresh on the GPU box (run through gpurun from the repository root).  Every command
# runs under a hard limit: a persistent kernel that is never asked to leave would otherwise keep the box until gpurun's own.
R=${ROUND:-r2}
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 500 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --report > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1_kernel_table.txt
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null
timeout -s KILL 300 python bench.py --config conv64 --steps 3 --warmup 3 > gpurun_out/${R}_bench_conv64.json 2>/dev/null
( echo "== tools/fc_call_latency.py"; PNN_FC_STAMPS=1 timeout -s KILL 100 python tools/fc_call_latency.py 2>&1 | tail -6
  echo "== tools/hm_latency.py"; timeout -s KILL 100 python tools/hm_latency.py 2>&1 | tail -6
  echo "== tools/hm_latency.py --no-fused"; timeout -s KILL 100 python tools/hm_latency.py --no-fused 2>&1 | tail -6
  echo "== tools/ref_backend_latency.py"; timeout -s KILL 200 python tools/ref_backend_latency.py 2>&1 | tail -8 ) > gpurun_out/${R}_inloop_latency.txt 2>&1
timeout -s KILL 1500 python hm/config4.py --out gpurun_out/${R}_config3_hm_substitution_1080p.json 2>&1 | tail -30
timeout -s KILL 400 python bench.py --config hm > gpurun_out/${R}_bench_hm.json 2>/dev/null
( VARIANT=switch ORDER="0 1" VAR=PNN_HM_PREFETCH_QUADRANT bash tools/hm_prefetch_ab.sh ) > gpurun_out/${R}_hm_prefetch_quadrant_ab_switch_1080p.txt 2>&1
python - <<PY
import json
d = json.load(open('gpurun_out/${R}_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])
PY
