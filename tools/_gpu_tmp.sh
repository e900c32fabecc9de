cd $GRAFT_REPO_ROOT
R=r2
timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout -s KILL 500 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 --report > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1_kernel_table.txt
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>/dev/null
python - <<PY
import json
d = json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
r = json.load(open('gpurun_out/r2_bench_reference_arm.json'))
print('reference arm', r['value'], r['ms_per_step'], r['ms_full_workload_at_this_rate'], r['cpu_baseline']['cores'])
PY
timeout -s KILL 400 python bench.py --config hm > gpurun_out/${R}_bench_hm.json 2> gpurun_out/${R}_bench_hm.err; tail -c 1200 gpurun_out/${R}_bench_hm.json; tail -2 gpurun_out/${R}_bench_hm.err
