"""CPU tests of the image-sharded offline evaluation: sharding, the single gather (gloo, world_size 2), the reduction."""
import os
import socket

import numpy
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from context_adaptive_neural_network_based_prediction_b200 import offline
from oracle import epilogue


def test_round_robin_sharding_covers_every_image_once():
    for n, world in ((100, 1), (100, 2), (100, 8), (24, 5)):
        shards = [offline.shard_image_indices(n, r, world) for r in range(world)]
        assert sorted(i for s in shards for i in s) == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_grid_blocks_counts():
    """320x480 images -> 79x119, 39x59, 19x29, 9x14 blocks; Kodak-shaped 512x768, W=8 -> 63x95."""
    for w, count in ((4, 79 * 119), (8, 39 * 59), (16, 19 * 29), (32, 9 * 14)):
        rows, cols = offline.grid_blocks(320, 480, w)
        assert len(rows) == count and rows.min() == w and cols.min() == w
        assert rows.max() + w <= 320 and cols.max() + w <= 480
    assert len(offline.grid_blocks(512, 768, 8)[0]) == 63 * 95
    idx, rows, cols = offline.blocks_of_images(3, 64, 96, 16)
    assert len(idx) == 3 * 3 * 5 and idx.tolist() == sorted(idx.tolist())


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = numpy.random.default_rng(rank)
    psnrs = torch.from_numpy(rng.uniform(10., 50., 37))
    wins = (psnrs > 30.).to(torch.uint8)
    g_psnr, g_win = offline.gather_statistics(psnrs, wins, rank, world)
    if rank == 0:
        numpy.save(os.path.join(out_dir, 'psnr.npy'), g_psnr.numpy())
        numpy.save(os.path.join(out_dir, 'win.npy'), g_win.numpy())
    else:
        assert g_psnr is None and g_win is None
    dist.barrier()
    dist.destroy_process_group()


def _worker_ragged(rank, world, port, out_dir):
    """Ranks with different block counts (100 images over 8 ranks: 13 or 12 images each)."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    counts = [41, 29]
    rng = numpy.random.default_rng(10 + rank)
    psnrs = torch.from_numpy(rng.uniform(10., 50., counts[rank]))
    wins = (psnrs > 25.).to(torch.uint8)
    g_psnr, g_win = offline.gather_statistics(psnrs, wins, rank, world, counts=counts)
    if rank == 0:
        stats = offline.reduce_statistics_device(g_psnr, g_win)
        numpy.save(os.path.join(out_dir, 'ragged.npy'), stats['psnrs_pnn'])
        numpy.save(os.path.join(out_dir, 'ragged_freq.npy'), numpy.array([stats['frequency_win_pnn']]))
    dist.barrier()
    dist.destroy_process_group()


def test_single_gather_with_different_block_counts(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker_ragged, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = numpy.concatenate([numpy.random.default_rng(10 + r).uniform(10., 50., c) for r, c in ((0, 41), (1, 29))])
    numpy.testing.assert_array_equal(numpy.load(str(tmp_path / 'ragged.npy')), want)
    assert numpy.load(str(tmp_path / 'ragged_freq.npy'))[0] == float((want > 25.).sum()) / want.size


def test_single_gather_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g_psnr = numpy.load(str(tmp_path / 'psnr.npy'))
    g_win = numpy.load(str(tmp_path / 'win.npy'))
    assert g_psnr.shape == (2, 37) and g_win.dtype == numpy.uint8
    for rank in range(2):
        want = numpy.random.default_rng(rank).uniform(10., 50., 37)
        numpy.testing.assert_array_equal(g_psnr[rank], want)            # float64 moved bit-exactly
        numpy.testing.assert_array_equal(g_win[rank], (want > 30.).astype(numpy.uint8))
    stats = offline.reduce_statistics(g_psnr, g_win)
    assert abs(stats['mean_psnr_pnn'] - g_psnr.mean()) < 1e-12
    assert stats['frequency_win_pnn'] == float((g_psnr > 30.).sum()) / g_psnr.size


def test_reduce_statistics_matches_reference_formula():
    """reference comparing_pnn_ipfcns_hevc_best_mode.py:39-88."""
    rng = numpy.random.default_rng(0)
    targets = rng.integers(0, 256, (20, 8, 8)).astype(numpy.uint8)
    preds = numpy.clip(targets.astype(int) + rng.integers(-6, 7, targets.shape), 0, 255).astype(numpy.uint8)
    base = rng.uniform(28., 36., 20)
    psnrs, freq = epilogue.performance_vs_baseline(targets, preds, base)
    stats = offline.reduce_statistics(psnrs, (psnrs - base > 0.).astype(numpy.uint8))
    assert stats['frequency_win_pnn'] == freq and abs(stats['mean_psnr_pnn'] - psnrs.mean()) < 1e-12


def test_reduce_statistics_device_matches_numpy():
    """The torch (device-side) reduction used by rank 0 in the multi-rank bench gives the reference's keys and values."""
    import torch
    rng = numpy.random.default_rng(3)
    psnrs = torch.from_numpy(rng.uniform(5., 45., (3, 777)))
    wins = torch.from_numpy((rng.uniform(size=(3, 777)) > 0.6).astype(numpy.uint8))
    a = offline.reduce_statistics_device(psnrs, wins)
    b = offline.reduce_statistics(psnrs.numpy(), wins.numpy())
    assert abs(a['mean_psnr_pnn'] - b['mean_psnr_pnn']) < 1e-12
    assert a['frequency_win_pnn'] == b['frequency_win_pnn']
    numpy.testing.assert_array_equal(a['psnrs_pnn'], b['psnrs_pnn'])


def test_driver_loop_finds_the_model_with_the_longest_training(tmp_path):
    """`predict_masks` (comparing_pnn_ipfcns_hevc_best_mode.py:386-433): directory tags, skipped directories, latest model."""
    tr, val = offline.masks_training_and_validation(8)
    assert tr == ((0, 0), (0, 8), (8, 0), (8, 8), ()) and val == ((0, 0), (0, 8), (8, 0), (8, 8))
    assert offline.find_model(str(tmp_path / 'masks_tr_0_0')) is None              # no such directory
    d = tmp_path / 'masks_tr_random'
    d.mkdir()
    assert offline.find_model(str(d)) is None                                      # no model in it
    for name in ('model_1000.pnnw', 'model_30000.pnnw', 'model_x.pnnw', 'model_200000.ckpt.meta', 'notes.txt'):
        (d / name).write_bytes(b'')
    assert offline.find_model(str(d)) == ('pnnw', str(d / 'model_30000.pnnw'), 30000)   # a .meta without its .index is no model
    (d / 'model_200000.ckpt.index').write_bytes(b'')
    assert offline.find_model(str(d)) == ('checkpoint', str(d / 'model_200000.ckpt'), 200000)
    (d / 'model_200000.pnnw').write_bytes(b'')
    assert offline.find_model(str(d))[0] == 'pnnw'                                 # the exported form of the same model wins
