"""Rate-distortion comparison of two HM runs (reference comparing_rate_distortion.py:491-561, hevc/performance.py).

    python hm/run_hm.py --variant regular      --qps 22,27,32,37 > regular.json
    python hm/run_hm.py --variant substitution --qps 22,27,32,37 > substitution.json
    python hm/rd_compare.py substitution.json regular.json 1080 1920

Prints the Bjontegaard metric (per-cent bitrate of the SECOND run relative to the first; positive = the first run saves
bitrate) and the two curves.  With the seeded random-init nets of this repository the number says nothing about
compression (the pretrained HM nets are not shipped); the script is the harness a user with the real weights runs.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import rd   # noqa: E402


def load(path):
    return [json.loads(line) for line in open(path) if line.startswith('{')]


def main():
    a, b = load(sys.argv[1]), load(sys.argv[2])
    height, width = int(sys.argv[3]), int(sys.argv[4])
    out = {
        'bjontegaard_percent': rd.compare_runs(a, b, height, width),
        'curve_0': [(r['qp'], rd.rate_of(r['bytes'], height, width), float(r['y_psnr_kbps'][1])) for r in sorted(a, key=lambda r: r['qp'])],
        'curve_1': [(r['qp'], rd.rate_of(r['bytes'], height, width), float(r['y_psnr_kbps'][1])) for r in sorted(b, key=lambda r: r['qp'])],
        'hash_ok': all(r.get('decoder_hash_ok') for r in a + b),
    }
    print(json.dumps(out))


if __name__ == '__main__':
    main()
