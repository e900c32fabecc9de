#!/bin/bash
# Round-end validation and evidence refresh on the GPU box (run through gpurun from the repository root).
timeout -k 10 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -k 10 500 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -3
timeout -k 10 400 python bench.py --steps 5 --warmup 3 --report > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1_table.txt
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -o gpurun_out/r1_gemm_fc_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e > /dev/null 2>&1
timeout -k 10 300 ncu --set full --clock-control none -k 'regex:gemm_tc|conv_first|merger_mma' -s 21 -c 22 -o gpurun_out/r1_conv16_full -f python tools/profile_net.py 16 32768 2 > /dev/null 2>&1
timeout -k 10 200 python tools/config_sweeps.py > gpurun_out/r1_configs_0_and_2.txt 2>&1
timeout -k 10 400 python hm/run_hm.py --qps 22,32 --frozen-graphs > gpurun_out/hm_final4.json 2> gpurun_out/hm_final4.err
timeout -k 10 300 python tools/hm_latency.py --cpu > gpurun_out/r1_config3_cpu_projection.txt 2>&1
python - <<'PY'
import json
d = json.load(open('gpurun_out/r1_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])
for l in open('gpurun_out/hm_final4.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['qp'], round(d['encoder_wall_s'], 2), d['decoder_hash_ok'], d['recon_enc_equals_dec'], [x.split(', ')[-1] for x in d['pnn_encoder']])
PY
