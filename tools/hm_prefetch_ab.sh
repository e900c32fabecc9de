#!/bin/bash
# A/B of the fast-pass prefetch (pnn_predict_hm_begin) with ONE executable on ONE box: the direct-binding HM encoder with
# PNN_HM_PREFETCH=0 and =1 (twice each), 1080p, after a small warm-up encode (first use of a box pays for paging the
# libraries in).  The bitstreams must be identical: the md5 of each is printed.
QPS=${QPS:-32}
VARIANT=${VARIANT:-substitution}
SIZE=${SIZE:---width 1920 --height 1080}
timeout -s KILL 200 python hm/run_hm.py --variant $VARIANT --backend direct --width 416 --height 240 --qps 32 > /dev/null 2>&1
for P in ${ORDER:-0 1 0 1}; do
  echo "== ${VAR:-PNN_HM_PREFETCH}=$P"
  rm -rf /tmp/hm_ab_$P; 
  env ${VAR:-PNN_HM_PREFETCH}=$P timeout -s KILL 600 python hm/run_hm.py --variant $VARIANT --backend direct --qps $QPS $SIZE --keep /tmp/hm_ab_$P 2>&1 | python -c "
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        print(line[:300]); continue
    r = json.loads(line)
    if 'error' in r:
        print(r); continue
    print(r['qp'], 'enc wall %.2f s' % r['encoder_wall_s'], 'HM total', r['encoder_total_time_s'], 'bytes', r['bytes'], 'hash', r['decoder_hash_ok'], 'recon', r['recon_enc_equals_dec'])
    for l in r['pnn_encoder']: print('   ', l)
"
  md5sum /tmp/hm_ab_$P/str_*.bin
done
