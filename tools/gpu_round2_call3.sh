cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
summ() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('backend','ref_threads','qp','encoder_wall_s','encoder_total_time_s','decoder_wall_s','decoder_total_time_s','bytes','decoder_hash_ok','recon_enc_equals_dec')}); print(d['pnn_encoder']); print(d['pnn_decoder'][-2:])
    else: print(l.strip())
"; }
PNN_TIMING=1 timeout 600 python hm/run_hm.py --backend direct --qps 32 --width 832 --height 480 2>&1 | summ
timeout 600 python hm/run_hm.py --backend cuda --qps 32 --width 832 --height 480 2>&1 | summ
timeout 600 python hm/run_hm.py --backend cpu --qps 32 --width 832 --height 480 2>&1 | summ
timeout 600 python hm/run_hm.py --backend cpu --ref-threads 8,16 --qps 32 --width 832 --height 480 2>&1 | summ
timeout 600 python hm/run_hm.py --backend cpu --ref-threads 4,16 --qps 32 --width 832 --height 480 2>&1 | summ
