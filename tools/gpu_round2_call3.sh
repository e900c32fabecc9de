cd $GRAFT_REPO_ROOT
nproc; lscpu | grep "Model name"
timeout 600 python hm/run_hm.py --qps 32,22 2>&1 | tee gpurun_out/r2_hm_cuda_first.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('backend','qp','encoder_wall_s','encoder_total_time_s','decoder_wall_s','decoder_total_time_s','bytes','decoder_hash_ok','recon_enc_equals_dec')}); print(d['pnn_encoder']); print(d['pnn_decoder'][-2:])
    else: print(l.strip())
"
timeout 900 python hm/run_hm.py --backend cpu --qps 32 2>&1 | tee gpurun_out/r2_hm_cpu_first.json | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('backend','qp','host_cores','encoder_wall_s','encoder_total_time_s','decoder_wall_s','decoder_total_time_s','bytes','decoder_hash_ok','recon_enc_equals_dec')}); print(d['pnn_encoder'])
    else: print(l.strip())
"
timeout 300 python hm/run_hm.py --variant regular --qps 32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('qp','encoder_wall_s','encoder_total_time_s','decoder_wall_s','bytes')})
"
