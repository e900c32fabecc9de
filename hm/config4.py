"""BASELINE.json configs[3] measured honestly, in one run on one box: the reference's HM-16.15 substitution codec (unmodified
sources + the bindings of hm/) built against libpnn_cuda and, beside it, against the CPU baseline backend, plus stock HM.

    python hm/config4.py [--qps 22,27,32,37] [--out profiles/r2_config3_hm.json]

Per QP: wall time of encoder and decoder (process start to exit), HM's own `Total Time`, PNN call statistics, decoder picture
hash, encoder / decoder reconstruction identity.  Summary: wall ratio CPU build / GPU build (the CPU backend with all host
threads -- what the reference's SessionOptions() default means -- and with the fastest thread setting of
tools/ref_backend_latency.py), and the Bjontegaard rate difference between the GPU build and the CPU build
(<pkg>/rd.py, pinned to the reference's compute_bjontegaard; reference comparing_rate_distortion.py:491-561).
The second part repeats the rate-distortion comparison on a real luminance image with the two pretrained nets the reference
ships in the 4x4 / 8x8 slots, where the encoder does select the neural-network mode.
"""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from context_adaptive_neural_network_based_prediction_b200 import rd


def run(backend, qps, extra=(), timeout=1500, env=None):
    cmd = [sys.executable, os.path.join(ROOT, 'hm', 'run_hm.py'), '--qps', qps] + list(extra)
    cmd += ['--variant', 'regular'] if backend == 'regular' else ['--backend', backend]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})))
    rows = [json.loads(l) for l in p.stdout.splitlines() if l.startswith('{')]
    for r in rows:
        if 'error' in r:
            raise RuntimeError('%s failed: %s' % (backend, r))
    return rows


def brief(rows):
    keys = ('qp', 'encoder_wall_s', 'encoder_total_time_s', 'decoder_wall_s', 'decoder_total_time_s', 'bytes', 'bitstream_md5', 'y_psnr_kbps',
            'decoder_hash_ok', 'recon_enc_equals_dec', 'pnn_encoder', 'pnn_decoder')
    return [{k: r.get(k) for k in keys} for r in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--qps', default='22,27,32,37')
    ap.add_argument('--out', default='')
    ap.add_argument('--skip-real', action='store_true')
    ap.add_argument('--best-threads', default='8,16')
    args = ap.parse_args()
    out = {'config': 'configs[3]: HM-16.15 substitution encoder + decoder, first-frame intra, synthetic 1920x1080 4:0:0, QP ' + args.qps,
           'host_cores': os.cpu_count()}
    # a small encode first: the first CUDA process on a fresh box pays the driver's cold start (seconds), which belongs to no arm
    run('direct', '32', ['--width', '416', '--height', '240'])
    runs = {}
    # gpu_direct: the direct binding, neural-network requests posted at the start of the fast pass (pnn_predict_hm_begin);
    # gpu_direct_no_prefetch: the same executable with PNN_HM_PREFETCH=0; gpu_seam: unmodified codec sources (link seam)
    for name, backend, extra in (('regular', 'regular', ()), ('gpu_direct', 'direct', ()), ('gpu_direct_no_prefetch', 'direct', ()),
                                 ('gpu_seam', 'cuda', ()),
                                 ('cpu_all_threads', 'cpu', ()), ('cpu_best_threads', 'cpu', ('--ref-threads', args.best_threads))):
        runs[name] = run(backend, args.qps, extra, env={'PNN_HM_PREFETCH': '0'} if name == 'gpu_direct_no_prefetch' else None)
        out[name] = brief(runs[name])
        print(name, [(r['qp'], round(r['encoder_wall_s'], 2)) for r in runs[name]], file=sys.stderr, flush=True)
    summary = {}
    for gpu in ('gpu_direct', 'gpu_direct_no_prefetch', 'gpu_seam'):
        for cpu in ('cpu_all_threads', 'cpu_best_threads'):
            summary['encoder_wall_ratio_%s_over_%s' % (cpu, gpu)] = {
                str(a['qp']): b['encoder_wall_s'] / a['encoder_wall_s'] for a, b in zip(runs[gpu], runs[cpu])}
        summary['bjontegaard_percent_%s_vs_cpu' % gpu] = rd.compare_runs(runs[gpu], runs['cpu_best_threads'], 1080, 1920) \
            if len(runs[gpu]) >= 4 else None
        summary['identical_bitstream_sizes_%s_vs_cpu' % gpu] = all(a['bytes'] == b['bytes'] for a, b in zip(runs[gpu], runs['cpu_best_threads']))
        summary['identical_bitstreams_md5_%s_vs_gpu_seam' % gpu] = all(a['bitstream_md5'] == b['bitstream_md5'] for a, b in zip(runs[gpu], runs['gpu_seam']))
        summary['identical_bitstreams_md5_%s_vs_cpu' % gpu] = all(a['bitstream_md5'] == b['bitstream_md5'] for a, b in zip(runs[gpu], runs['cpu_best_threads']))
    summary['all_hash_ok'] = all(r['decoder_hash_ok'] and r['recon_enc_equals_dec'] for k in runs if k != 'regular' for r in runs[k])
    out['summary'] = summary
    if not args.skip_real:
        real = {}
        image = os.path.join(ROOT, 'tests', 'golden', 'cliff_luma.npy')
        for name, backend in (('regular', 'regular'), ('gpu_direct', 'direct'), ('cpu', 'cpu')):
            extra = ['--image', image] + ([] if backend == 'regular' else ['--trained-small-nets', '--ref-threads', args.best_threads])
            rows = run(backend, args.qps, extra)
            real[name] = brief(rows)
        real['bjontegaard_percent_gpu_vs_cpu'] = rd.compare_runs(real['gpu_direct'], real['cpu'], 160, 240)
        real['bjontegaard_percent_gpu_vs_regular'] = rd.compare_runs(real['gpu_direct'], real['regular'], 160, 240)
        real['bjontegaard_percent_cpu_vs_regular'] = rd.compare_runs(real['cpu'], real['regular'], 160, 240)
        real['note'] = ('160x240 luminance crop of the reference image sets/pseudo_data/rgb_cliff.jpg; widths 4 and 8 use the pretrained '
                        'CONV-4 / CONV-8 checkpoints the reference ships (tests/golden/), widths 16-64 seeded random init')
        out['real_image_trained_small_nets'] = real
    text = json.dumps(out, indent=1)
    if args.out:
        open(args.out, 'w').write(text + '\n')
    print(json.dumps(out['summary'], indent=1))
    if 'real_image_trained_small_nets' in out:
        print(json.dumps({k: v for k, v in out['real_image_trained_small_nets'].items() if k.startswith('bjon')}, indent=1))


if __name__ == '__main__':
    main()
