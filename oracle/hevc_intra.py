"""HEVC intra prediction (35 modes) and best-mode selection on the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates reference hevc/intraprediction/c++/source/extracted_hevc_intraprediction.cpp:3-421 (the HM-16.15 functions
xPredIntraAng / xPredIntraPlanar / predIntraGetPredValDC / xDCPredFiltering on an UNFILTERED intra pattern) and
hevc/intraprediction/intraprediction.py:8-292 (pattern extraction, best mode = highest prediction PSNR over the 35 modes).
"""
import numpy

ANG_TABLE = (0, 2, 5, 9, 13, 17, 21, 26, 32)
INV_ANG_TABLE = (0, 4096, 1638, 910, 630, 482, 390, 315, 256)


def extract_intra_pattern(image_u8, width, row_ref, col_ref, mask_w, mask_h):
    """reference intraprediction.py:8-89.  (row_ref, col_ref) is the pixel above-left of the target block.

    Returns (first row [wp], first column [hp]) with wp = 2W+1-mask_w, hp = 2W+1-mask_h, additionally cut at the image
    border (the C-ABI treats pixels outside the image like masked ones).
    """
    hp = min(2 * width + 1 - mask_h, image_u8.shape[0] - row_ref)
    wp = min(2 * width + 1 - mask_w, image_u8.shape[1] - col_ref)
    return image_u8[row_ref, col_ref:col_ref + wp].astype(numpy.int64), image_u8[row_ref:row_ref + hp, col_ref].astype(numpy.int64)


def full_references(first_row, first_col, width):
    """reference extracted_hevc_intraprediction.cpp:34-84: pad the missing top-right / bottom-left part with the last pixel.

    Returns ref_above[0 .. 2W] and ref_left[0 .. 2W], index 0 = the above-left corner pixel.
    """
    n = 2 * width + 1
    above = numpy.empty(n, dtype=numpy.int64)
    left = numpy.empty(n, dtype=numpy.int64)
    above[:len(first_row)] = first_row
    above[len(first_row):] = first_row[-1]
    left[:len(first_col)] = first_col
    left[len(first_col):] = first_col[-1]
    return above, left


def predict_mode(above, left, width, mode):
    """One mode, int array [W, W] (reference extracted_hevc_intraprediction.cpp:86-421, bit depth 8, luma)."""
    w = width
    pred = numpy.zeros((w, w), dtype=numpy.int64)
    edge_filter = w <= 16
    if mode == 0:                                            # planar, :324-383
        shift = int(numpy.log2(w))
        top, lft = above[1:w + 2], left[1:w + 2]
        bottom_left, top_right = lft[w], top[w]
        for y in range(w):
            for x in range(w):
                hor = (lft[y] << shift) + w + (x + 1) * (top_right - lft[y])
                ver = (top[x] << shift) + (y + 1) * (bottom_left - top[x])
                pred[y, x] = (hor + ver) >> (shift + 1)
        return pred
    if mode == 1:                                            # DC, :286-322 and :385-421
        dc = (int(above[1:w + 1].sum()) + int(left[1:w + 1].sum()) + w) // (2 * w)
        pred[:, :] = dc
        if edge_filter:
            pred[0, 0] = (above[1] + left[1] + 2 * dc + 2) >> 2
            for x in range(1, w):
                pred[0, x] = (above[x + 1] + 3 * dc + 2) >> 2
            for y in range(1, w):
                pred[y, 0] = (left[y + 1] + 3 * dc + 2) >> 2
        return pred
    is_ver = mode >= 18                                      # angular, :161-283
    ang_mode = mode - 26 if is_ver else -(mode - 10)
    abs_mode = abs(ang_mode)
    angle = (-1 if ang_mode < 0 else 1) * ANG_TABLE[abs_mode]
    inv_angle = INV_ANG_TABLE[abs_mode]
    main_src, side_src = (above, left) if is_ver else (left, above)
    off = w - 1
    ref_main = numpy.zeros(3 * w + 2, dtype=numpy.int64)      # ref_main[off + i] holds refMain[i], i in [-(w-1), 2w]
    if angle < 0:
        ref_main[off:off + w + 1] = main_src[:w + 1]
        inv_sum = 128
        k = -1
        while k > ((w * angle) >> 5):
            inv_sum += inv_angle
            ref_main[off + k] = side_src[inv_sum >> 8]
            k -= 1
    else:
        ref_main[off:off + 2 * w + 1] = main_src
    tmp = numpy.zeros((w, w), dtype=numpy.int64)
    if angle == 0:
        for y in range(w):
            tmp[y, :] = ref_main[off + 1:off + w + 1]
        if edge_filter:
            for y in range(w):
                tmp[y, 0] = min(max(tmp[y, 0] + ((side_src[y + 1] - side_src[0]) >> 1), 0), 255)
    else:
        for y in range(w):
            delta = (y + 1) * angle
            d_int, d_fract = delta >> 5, delta & 31
            for x in range(w):
                if d_fract:
                    tmp[y, x] = ((32 - d_fract) * ref_main[off + x + d_int + 1] + d_fract * ref_main[off + x + d_int + 2] + 16) >> 5
                else:
                    tmp[y, x] = ref_main[off + x + d_int + 1]
    return tmp if is_ver else tmp.T.copy()


def predict_all_modes(first_row, first_col, width):
    above, left = full_references(first_row, first_col, width)
    return numpy.stack([predict_mode(above, left, width, m) for m in range(35)]).astype(numpy.uint8)


def best_mode(first_row, first_col, target_u8):
    """reference intraprediction.py:226-292: first mode with the strictly highest PSNR (= strictly lowest SSE).

    Returns (index, psnr float64, prediction uint8 [W, W]).
    """
    width = target_u8.shape[0]
    preds = predict_all_modes(first_row, first_col, width)
    sse = ((preds.astype(numpy.int64) - target_u8.astype(numpy.int64)[None]) ** 2).reshape(35, -1).sum(axis=1)
    idx = int(numpy.argmin(sse))                              # argmin returns the first minimum: same tie rule
    mse = sse[idx] / float(width * width)
    return idx, 10. * numpy.log10(255. ** 2 / (mse + 1.e-6)), preds[idx]


def best_modes_of_blocks(images_u8, img_idx, rows, cols, width, mask_w, mask_h):
    """(rows[i], cols[i]) = top-left pixel of target block i; the pattern anchor is (row - 1, col - 1)
    (reference comparing_pnn_ipfcns_hevc_best_mode.py:234-235)."""
    n = len(rows)
    idx = numpy.zeros(n, dtype=numpy.uint8)
    psnrs = numpy.zeros(n)
    preds = numpy.zeros((n, width, width), dtype=numpy.uint8)
    for i in range(n):
        img = images_u8[img_idx[i]]
        fr, fc = extract_intra_pattern(img, width, rows[i] - 1, cols[i] - 1, mask_w, mask_h)
        idx[i], psnrs[i], preds[i] = best_mode(fr, fc, img[rows[i]:rows[i] + width, cols[i]:cols[i] + width])
    return idx, psnrs, preds
