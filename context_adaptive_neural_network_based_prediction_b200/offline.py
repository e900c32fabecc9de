"""Offline batched evaluation sharded by image.

Host-side mirror of the PNN half of the reference's `predict_mask`
(comparing_pnn_ipfcns_hevc_best_mode.py:162-322): for every image of the rank's shard, every block
is predicted, its PSNR and win flag against a supplied baseline are computed, and ONE gather brings the
per-block statistics to rank 0, which reduces them to the reference's dictionary keys
(`psnrs_pnn`, `mean_psnr_pnn`, `frequency_win_pnn`, comparing_pnn_...py:264-322).

The blocks of different images are independent, so the shards need no data-path collective; the gather
moves <= tens of MB and is latency-bound.
"""
import os

import numpy


def shard_image_indices(nb_images, rank, world_size):
    """Round-robin assignment of images to ranks (image i goes to rank i % world_size)."""
    return list(range(rank, nb_images, world_size))


def grid_blocks(height, width_image, width_target):
    """All W-aligned target blocks whose context anchor (row - W, col - W) lies inside the image."""
    rows, cols = numpy.meshgrid(numpy.arange(width_target, height - width_target + 1, width_target),
                                numpy.arange(width_target, width_image - width_target + 1, width_target), indexing='ij')
    return rows.ravel().astype(numpy.int32), cols.ravel().astype(numpy.int32)


def blocks_of_images(nb_images, height, width_image, width_target):
    """(image_index, rows, cols) int32 arrays for all grid blocks of `nb_images` images, image-major."""
    rows, cols = grid_blocks(height, width_image, width_target)
    idx = numpy.repeat(numpy.arange(nb_images, dtype=numpy.int32), len(rows))
    return idx, numpy.tile(rows, nb_images), numpy.tile(cols, nb_images)


def evaluate_blocks(engine, images_uint8, width_target, is_fully_connected, rows, cols, image_index=None, masks=(0, 0)):
    """PNN versus the best HEVC intra mode on the given blocks: the body of the reference's `predict_mask`
    (comparing_pnn_ipfcns_hevc_best_mode.py:220-322) in two library calls, with the reference's dictionary keys."""
    pnn = engine.predict_image_blocks(width_target, is_fully_connected, images_uint8, rows, cols, image_index, masks=masks,
                                      want_float=False)
    hevc = engine.hevc_best_mode(width_target, images_uint8, rows, cols, image_index, masks=masks, want_predictions=False)
    psnrs_pnn, psnrs_hevc = pnn['psnrs'], hevc['psnrs_hevc_best_mode']
    n = max(1, len(psnrs_pnn))
    return {
        'psnrs_pnn': psnrs_pnn,
        'indices_hevc_best_mode': hevc['indices_hevc_best_mode'],
        'psnrs_hevc_best_mode': psnrs_hevc,
        'mean_psnr_pnn': float(numpy.mean(psnrs_pnn)) if len(psnrs_pnn) else float('nan'),
        'mean_psnr_hevc_best_mode': float(numpy.mean(psnrs_hevc)) if len(psnrs_hevc) else float('nan'),
        # comparing_pnn_ipfcns_hevc_best_mode.py:87
        'frequency_win_pnn': float(numpy.count_nonzero(psnrs_pnn - psnrs_hevc > 0.)) / n,
        'predictions_pnn_uint8': pnn['predictions_uint8'],
    }


def masks_training_and_validation(width_target):
    """The five training maskings and the four test maskings of the reference's driver
    (comparing_pnn_ipfcns_hevc_best_mode.py:501-513); `()` is the model trained with random masks."""
    w = width_target
    return ((0, 0), (0, w), (w, 0), (w, w), ()), ((0, 0), (0, w), (w, 0), (w, w))


def find_model(path_to_directory_load_save):
    """The model with the longest training in a `masks_tr_*` directory, or None
    (comparing_pnn_ipfcns_hevc_best_mode.py:418-433: `model_<iterations>.ckpt`, chosen through its `.ckpt.meta` file).

    Accepted forms, in this order: an exported flat binary `model_<iterations>.pnnw`, the TensorFlow V2 checkpoint itself
    (`model_<iterations>.ckpt.index` + `.data-00000-of-00001`, read without TensorFlow by `weights.read_tf_v2_bundle`).
    Returns (kind, path-or-prefix, iterations)."""
    if not os.path.isdir(path_to_directory_load_save):
        return None
    found = {}
    for name in os.listdir(path_to_directory_load_save):
        for kind, extension in (('pnnw', '.pnnw'), ('checkpoint', '.ckpt.index'), ('checkpoint', '.ckpt.meta')):
            if name.startswith('model_') and name.endswith(extension):
                try:
                    iterations = int(name[len('model_'):-len(extension)])
                except ValueError:
                    continue      # (tools/tools.py:175-183: a name whose middle is not an integer is ignored)
                if kind == 'checkpoint' and not os.path.exists(os.path.join(path_to_directory_load_save, 'model_%d.ckpt.index' % iterations)):
                    continue
                found.setdefault(iterations, set()).add(kind)
    if not found:
        return None
    iterations = max(found)
    if 'pnnw' in found[iterations]:
        return 'pnnw', os.path.join(path_to_directory_load_save, 'model_%d.pnnw' % iterations), iterations
    return 'checkpoint', os.path.join(path_to_directory_load_save, 'model_%d.ckpt' % iterations), iterations


def predict_masks(engine, images_uint8, width_target, is_fully_connected, rows, cols, path_to_directory_coeffs_load_save,
                  image_index=None, tuples_width_height_masks_tr=None, tuples_width_height_masks_val=None,
                  path_to_directory_coeffs_vis=None):
    """The driver loop of the reference's offline comparison (`predict_masks`,
    comparing_pnn_ipfcns_hevc_best_mode.py:324-452): every PNN model trained with one masking (sub-directories
    `masks_tr_<w>_<h>` and `masks_tr_random` of `path_to_directory_coeffs_load_save`; a missing directory or one without a model
    is skipped, the model with the longest training is taken) is evaluated under every test masking, against the best HEVC
    mode on the same blocks.  `rows` / `cols` are the top-left pixels of the TARGET blocks (reference: `row_1sts + W`).
    Returns {tag_masks_tr: {tag_masks_val: dictionary of `evaluate_blocks`}}; with `path_to_directory_coeffs_vis` each
    dictionary's scalars are also written as `<vis>/<tag_tr>/<tag_val>/dictionary_performance.pkl` (protocol 2, the
    reference's file name; IPFCN-S, a Caffe model, is out of scope)."""
    import pickle
    import tempfile

    from . import weights as weights_module
    default_tr, default_val = masks_training_and_validation(width_target)
    tuples_tr = default_tr if tuples_width_height_masks_tr is None else tuples_width_height_masks_tr
    tuples_val = default_val if tuples_width_height_masks_val is None else tuples_width_height_masks_val
    results = {}
    for masks_tr in tuples_tr:
        tag_masks_tr = 'masks_tr_{0}_{1}'.format(masks_tr[0], masks_tr[1]) if masks_tr else 'masks_tr_random'
        model = find_model(os.path.join(path_to_directory_coeffs_load_save, tag_masks_tr))
        if model is None:
            continue
        kind, path, _ = model
        if kind == 'checkpoint':
            with tempfile.TemporaryDirectory(prefix='pnn_export_') as tmp:
                flat = os.path.join(tmp, 'model.pnnw')
                weights_module.export_checkpoint(path, width_target, is_fully_connected, flat)
                engine.load_net(flat)
        else:
            engine.load_net(path)
        results[tag_masks_tr] = {}
        for masks_val in tuples_val:
            tag_masks_val = 'masks_val_{0}_{1}'.format(masks_val[0], masks_val[1])
            out = evaluate_blocks(engine, images_uint8, width_target, is_fully_connected, rows, cols, image_index,
                                  masks=tuple(masks_val))
            results[tag_masks_tr][tag_masks_val] = out
            if path_to_directory_coeffs_vis is not None:
                path_to_directory_vis = os.path.join(path_to_directory_coeffs_vis, tag_masks_tr, tag_masks_val)
                os.makedirs(path_to_directory_vis, exist_ok=True)
                scalars = {key: out[key] for key in ('mean_psnr_pnn', 'mean_psnr_hevc_best_mode', 'frequency_win_pnn')}
                with open(os.path.join(path_to_directory_vis, 'dictionary_performance.pkl'), 'wb') as file:
                    pickle.dump(scalars, file, protocol=2)
    return results


def gather_statistics(psnrs_local, wins_local, rank, world_size, group=None, counts=None):
    """ONE gather of the per-block statistics to rank 0.

    `psnrs_local` float64 and `wins_local` uint8 are torch tensors (CUDA for NCCL, CPU for gloo).  They travel as bytes in
    one message of 9 bytes per block (8 of the float64 PSNR, 1 of the win flag), so a single collective is issued and both
    arrive bit-exactly.  `counts` (blocks per rank) allows ranks to hold different numbers of blocks, as happens when 100
    images are sharded over 8 ranks: messages are padded to the largest count and cut back on rank 0.
    Returns (psnrs, wins) on rank 0 -- tensors [world, n] when every rank holds n blocks, else lists of per-rank tensors --
    and (None, None) elsewhere.
    """
    import torch
    import torch.distributed as dist
    n = psnrs_local.numel()
    n_max = n if counts is None else max(counts)
    packed = torch.zeros(9 * n_max, dtype=torch.uint8, device=psnrs_local.device)
    packed[:8 * n] = psnrs_local.to(torch.float64).contiguous().view(torch.uint8)
    packed[8 * n_max:8 * n_max + n] = wins_local.to(torch.uint8)
    if world_size == 1:
        out = [packed]
    else:
        out = [torch.empty_like(packed) for _ in range(world_size)] if rank == 0 else None
        dist.gather(packed, out, dst=0, group=group)
    if rank != 0:
        return None, None
    sizes = [n] * world_size if counts is None else list(counts)
    psnrs = [o[:8 * c].view(torch.float64) for o, c in zip(out, sizes)]
    wins = [o[8 * n_max:8 * n_max + c] for o, c in zip(out, sizes)]
    if counts is None:
        return torch.stack(psnrs), torch.stack(wins)
    return psnrs, wins


def reduce_statistics_device(psnrs, wins, pinned=None):
    """`reduce_statistics` for the gathered torch tensors of rank 0 while they are still on the GPU: the two scalars are
    reduced on the device, the per-block PSNRs come back through one copy (into `pinned`, a pinned float64 tensor of the
    same number of elements, when given)."""
    import torch
    if isinstance(psnrs, (list, tuple)):                     # ranks with different block counts
        psnrs, wins = torch.cat(list(psnrs)), torch.cat(list(wins))
    flat = psnrs.reshape(-1)
    mean = flat.mean() if flat.numel() else torch.tensor(float('nan'))
    won = (wins.reshape(-1) != 0).sum()
    if pinned is not None:
        host = pinned[:flat.numel()]
        host.copy_(flat, non_blocking=True)
    else:
        host = flat.cpu()
    mean_host, won_host = float(mean.item()), int(won.item())          # .item() also orders the copy above
    if psnrs.is_cuda:
        torch.cuda.current_stream(psnrs.device).synchronize()
    return {
        'psnrs_pnn': host.numpy(),
        'mean_psnr_pnn': mean_host,
        'frequency_win_pnn': float(won_host) / max(1, flat.numel()),
    }


def reduce_statistics(psnrs, wins):
    """The reference's result keys (comparing_pnn_ipfcns_hevc_best_mode.py:264-322) from the gathered arrays."""
    psnrs = numpy.asarray(psnrs, dtype=numpy.float64).ravel()
    wins = numpy.asarray(wins).ravel()
    return {
        'psnrs_pnn': psnrs,
        'mean_psnr_pnn': float(psnrs.mean()) if psnrs.size else float('nan'),
        'frequency_win_pnn': float(numpy.count_nonzero(wins)) / max(1, wins.size),
    }
