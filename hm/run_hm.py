"""BASELINE.json configs[3]: HM-16.15 substitution encoder + decoder, first-frame intra of a synthetic frame.

Runs the executables built by hm/build_hm.sh (the reference's codec, unmodified, linked to libpnn_cuda through
hm/shim/), with seeded random-init PNN weights (the pretrained HM weights are not shipped with the reference),
and prints one JSON object per QP: wall times, HM's own "Total Time", PNN call statistics per block width,
decoder picture-hash status and whether encoder and decoder reconstructions are byte-identical.
"""
import argparse, hashlib, json, os, pickle, re, subprocess, sys, tempfile, time
import numpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from context_adaptive_neural_network_based_prediction_b200 import weights

ap = argparse.ArgumentParser()
ap.add_argument('--width', type=int, default=1920)
ap.add_argument('--height', type=int, default=1080)
ap.add_argument('--qps', default='22,27,32,37')
ap.add_argument('--variant', default='substitution')
ap.add_argument('--seed', type=int, default=0)
ap.add_argument('--backend', default='cuda', choices=('cuda', 'direct', 'cpu', 'null'),
                help='cuda: libpnn_cuda through the link seam hm/shim (only Session::Run replaced); direct: libpnn_cuda through the patched hooks of hm/direct (gather and epilogue in the library too); cpu: the *_cpu executables = same codec objects linked against '
                     'oracle/_ref/libpnn_ref.so (libtorch-CPU stand-in for the TensorFlow-CPU build, the baseline leg); '
                     'null: the same with PNN_REF_NULL=1 (predictions are zeros at once: the codec\'s own time)')
ap.add_argument('--ref-threads', default='', help='cpu backend: intra-op threads "fc,conv" (default: all cores for both)')
ap.add_argument('--keep', default='', help='directory that receives the bitstreams and reconstructions (for rd_compare)')
ap.add_argument('--image', default='', help='.npy uint8 luminance image instead of the synthetic frame')
ap.add_argument('--trained-small-nets', action='store_true',
                help='widths 4 and 8 use the two pretrained checkpoints the reference ships (CONV-4, CONV-8, tests/golden/)')
ap.add_argument('--frozen-graphs', action='store_true',
                help='the paths file names frozen graphs (binary GraphDef `graph_output.pbtxt`, as the reference HM set-up has them) '
                     'instead of PNNW files: libpnn_cuda reads their constants itself')
args = ap.parse_args()

build = os.path.join(ROOT, 'hm', '_build')
suffix = '' if args.backend == 'cuda' or args.variant == 'regular' else ('_direct' if args.backend == 'direct' else '_cpu')
enc = os.path.join(build, 'TAppEncoderStatic_' + args.variant + suffix)
dec = os.path.join(build, 'TAppDecoderStatic_' + args.variant + suffix)
if args.backend == 'null':
    os.environ['PNN_REF_NULL'] = '1'
if args.ref_threads:
    os.environ['PNN_REF_THREADS_FC'], os.environ['PNN_REF_THREADS_CONV'] = args.ref_threads.split(',')
cfg = os.path.join(build, 'intra_main_rext.cfg')
tmp = tempfile.mkdtemp(prefix='pnn_hm_')
if args.keep:
    os.makedirs(args.keep, exist_ok=True)
    tmp = args.keep
# weights: FC-4, FC-8, CONV-16, CONV-32, CONV-64 (hevc/hm_common/paths_to_graphs_output/pair.txt format)
lines = []
for w, is_fc in ((4, True), (8, True), (16, False), (32, False), (64, False)):
    path = os.path.join(tmp, 'net_%d.pnnw' % w)
    if args.trained_small_nets and w <= 8:
        path = os.path.join(ROOT, 'tests', 'golden', 'conv%d_single.pnnw' % w)
    else:
        weights.save_flat(path, w, is_fc, weights.init_weights(w, is_fc, seed=w))
    if args.frozen_graphs:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        from test_weights_export_cpu import make_graph          # GraphDef encoder of the tests
        width_f, is_fc_f, wts = weights.load_flat(path)
        os.makedirs(os.path.join(tmp, 'graph_%d' % w), exist_ok=True)
        path = os.path.join(tmp, 'graph_%d' % w, 'graph_output.pbtxt')
        make_graph(path, wts, width_f, bool(is_fc_f))
    lines += ['%d,0,0,%s' % (w, path), '%d,1,0,%s' % (w, path)]
paths_file = os.path.join(tmp, 'paths.txt')
open(paths_file, 'w').write('\n'.join(lines) + '\n')
mean_file = os.path.join(tmp, 'mean_training.pkl')
pickle.dump(bench.MEAN, open(mean_file, 'wb'), protocol=2)
frame = bench.synthetic_image(args.height, args.width, args.seed)
if args.image:
    frame = numpy.ascontiguousarray(numpy.load(args.image), dtype=numpy.uint8)
    args.height, args.width = frame.shape
yuv = os.path.join(tmp, 'in.yuv')
frame.tofile(yuv)
extra = [] if args.variant == 'regular' else ['--PathToAdditionalDirectory=' + tmp, '--PathToMeanTraining=' + mean_file, '--PathToFilePathsToGraphsOutput=' + paths_file]

for qp in [int(q) for q in args.qps.split(',')]:
    bit, rec_e, rec_d = (os.path.join(tmp, n % qp) for n in ('str_%d.bin', 'rec_enc_%d.yuv', 'rec_dec_%d.yuv'))
    stats = os.path.join(tmp, 'stats_%d.txt' % qp)
    env = dict(os.environ, PNN_HM_STATS=stats)
    cmd = [enc, '-c', cfg, '-i', yuv, '-b', bit, '-o', rec_e, '-wdt', str(args.width), '-hgt', str(args.height),
           '--InputBitDepth=8', '--InputChromaFormat=400', '--FramesToBeEncoded=1', '--QP=%d' % qp] + extra
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=None if os.environ.get('PNN_TIMING') else subprocess.PIPE, text=True, env=env)
    t_enc = time.time() - t0
    if p.returncode != 0:
        print(json.dumps({'qp': qp, 'error': 'encoder failed', 'stderr': (p.stderr or '')[-800:], 'stdout': p.stdout[-400:]}))
        continue
    enc_total = re.findall(r'Total Time:\s+([0-9.]+) sec', p.stdout)
    bits = re.findall(r'Bytes written to file:\s+(\d+)', p.stdout)
    psnr = re.findall(r'\s+1\s+a\s+([0-9.]+)\s+([0-9.]+)', p.stdout)
    enc_stats = open(stats).read().strip().split('\n') if os.path.exists(stats) else []
    stats_d = os.path.join(tmp, 'stats_dec_%d.txt' % qp)
    t0 = time.time()
    d = subprocess.run([dec, '-b', bit, '-o', rec_d, '-d', '8'] + extra, stdout=subprocess.PIPE, stderr=None if os.environ.get('PNN_TIMING') else subprocess.PIPE, text=True,
                       env=dict(os.environ, PNN_HM_STATS=stats_d))
    t_dec = time.time() - t0
    dec_total = re.findall(r'Total Time:\s+([0-9.]+) sec', d.stdout)
    same = os.path.exists(rec_e) and os.path.exists(rec_d) and open(rec_e, 'rb').read() == open(rec_d, 'rb').read()
    print(json.dumps({
        'config': 'configs[3]: HM-16.15 %s, first-frame intra, %s %dx%d 4:0:0, intra_main_rext.cfg' % (args.variant, os.path.basename(args.image) if args.image else 'synthetic', args.width, args.height),
        'backend': args.backend, 'host_cores': os.cpu_count(), 'ref_threads': args.ref_threads,
        'trained_small_nets': args.trained_small_nets, 'qp': qp, 'encoder_wall_s': t_enc, 'encoder_total_time_s': float(enc_total[0]) if enc_total else None,
        'decoder_wall_s': t_dec, 'decoder_total_time_s': float(dec_total[0]) if dec_total else None,
        'bytes': int(bits[0]) if bits else None, 'bitstream_md5': hashlib.md5(open(bit, 'rb').read()).hexdigest() if os.path.exists(bit) else None, 'y_psnr_kbps': psnr[0] if psnr else None,
        'decoder_rc': d.returncode, 'decoder_hash_ok': '(OK)' in d.stdout and 'ERROR' not in d.stdout,
        'recon_enc_equals_dec': same, 'pnn_encoder': enc_stats,
        'pnn_decoder': open(stats_d).read().strip().split('\n') if os.path.exists(stats_d) else [],
    }), flush=True)
