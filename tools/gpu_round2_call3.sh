cd $GRAFT_REPO_ROOT
summ() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('backend','ref_threads','qp','encoder_wall_s','encoder_total_time_s','decoder_wall_s','decoder_total_time_s','bytes','decoder_hash_ok','recon_enc_equals_dec')}); print(d['pnn_encoder']); print(d['pnn_decoder'][-2:])
    else: print(l.strip())
"; }
timeout -s KILL 100 python hm/run_hm.py --backend direct --qps 32 2>&1 | summ
timeout -s KILL 100 python hm/run_hm.py --backend cuda --qps 32 2>&1 | summ
PNN_HM_WARM_UP=0 timeout -s KILL 100 python hm/run_hm.py --backend direct --qps 32 2>&1 | summ
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1_mid.json 2> gpurun_out/r2_bench_n1_mid.err; tail -c 3000 gpurun_out/r2_bench_n1_mid.json; tail -5 gpurun_out/r2_bench_n1_mid.err
