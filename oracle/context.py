"""Context gather / mask / mean subtraction on the CPU (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two reference twins are restated:
  * the numpy path of the offline evaluation, reference sets/common.py:13-110 (slicing)
    and sets/common.py:351-475 (mean subtraction, masks, FC flattening);
  * the HM path, reference hevc/hm_common/c++/source_common/extraction_context.cpp:3-208.
"""
import numpy


def extract_portions(image_u8, width, row_1st, col_1st):
    """reference sets/common.py:99-109.  image_u8 [H,Wimg]; (row_1st, col_1st) = top-left of the context.

    Returns above [W,3W], left [2W,W], target [W,W] (uint8).  Unlike the
    reference (which raises when the 3W x 3W window leaves the image,
    sets/common.py:88-91) pixels outside the image come back as 0 together with
    a validity mask, because the C-ABI entry point masks them.
    """
    h, w = image_u8.shape
    above = numpy.zeros((width, 3 * width), dtype=numpy.uint8)
    left = numpy.zeros((2 * width, width), dtype=numpy.uint8)
    va = numpy.zeros((width, 3 * width), dtype=bool)
    vl = numpy.zeros((2 * width, width), dtype=bool)
    r1 = min(row_1st + 3 * width, h)
    c1 = min(col_1st + 3 * width, w)
    a = image_u8[row_1st:row_1st + width, col_1st:c1]
    above[:a.shape[0], :a.shape[1]] = a
    va[:a.shape[0], :a.shape[1]] = True
    l = image_u8[row_1st + width:r1, col_1st:col_1st + width]
    left[:l.shape[0], :l.shape[1]] = l
    vl[:l.shape[0], :l.shape[1]] = True
    target = image_u8[row_1st + width:row_1st + 2 * width, col_1st + width:col_1st + 2 * width]
    return above, left, target, va, vl


def preprocess(above_u8, left_u8, mean, mask_w, mask_h, valid_above=None, valid_left=None):
    """reference sets/common.py:454-461 for ONE block.

    float32(x) - mean (mean is a Python float; numpy keeps float32 for array-scalar
    arithmetic, i.e. the subtraction is done in float32 with float32(mean)), then the
    right `mask_w` columns of above and the bottom `mask_h` rows of left are zeroed.
    """
    width = above_u8.shape[0]
    if mask_w < 0 or mask_w > width or mask_w % 4:
        raise ValueError('mask_w does not belong to {0, 4, ..., W}')   # sets/common.py:444-445
    if mask_h < 0 or mask_h > width or mask_h % 4:
        raise ValueError('mask_h does not belong to {0, 4, ..., W}')   # sets/common.py:446-447
    m = numpy.float32(mean)
    above = above_u8.astype(numpy.float32) - m
    left = left_u8.astype(numpy.float32) - m
    above[:, 3 * width - mask_w:] = 0.
    left[2 * width - mask_h:, :] = 0.
    if valid_above is not None:
        above[~valid_above] = 0.
    if valid_left is not None:
        left[~valid_left] = 0.
    return above, left


def gather_image_blocks(images_u8, img_idx, rows, cols, width, mean, mask_w, mask_h):
    """Batched gather used as the checker of `pnn_predict_image_blocks`.

    images_u8 [n_img,H,Wimg]; (rows[i], cols[i]) = top-left pixel of TARGET block i
    (so the context's first pixel is (rows[i]-W, cols[i]-W), reference
    sets/common.py:99-109).  Returns above [N,W,3W,1], left [N,2W,W,1] float32,
    flat [N,5W^2] (reference sets/common.py:467-472) and targets [N,W,W] uint8.
    """
    n = len(rows)
    above = numpy.zeros((n, width, 3 * width, 1), dtype=numpy.float32)
    left = numpy.zeros((n, 2 * width, width, 1), dtype=numpy.float32)
    targets = numpy.zeros((n, width, width), dtype=numpy.uint8)
    for i in range(n):
        a, l, t, va, vl = extract_portions(images_u8[img_idx[i]], width, rows[i] - width, cols[i] - width)
        pa, pl = preprocess(a, l, mean, mask_w, mask_h, va, vl)
        above[i, :, :, 0] = pa
        left[i, :, :, 0] = pl
        targets[i] = t
    flat = numpy.concatenate([above.reshape(n, -1), left.reshape(n, -1)], axis=1)
    return above, left, flat, targets


def extract_context_portions_hm(plane_i32, stride, origin, flags, n_avail, unit_w, unit_h,
                                above_units, left_units, width, mean):
    """Restatement of reference extraction_context.cpp:3-208 (same control flow).

    plane_i32: flat int32 buffer (HM reconstruction), `origin` = flat index of the
    top-left pixel of the current TB, flags[left_units+above_units+1] ordered
    bottom-left -> top-left, above-left, above -> above-right
    (reference TComPattern.cpp:260-280).  Returns (status, above [W*3W], left [2W*W]).
    """
    cw = 3 * width
    above = numpy.zeros(width * cw, dtype=numpy.float32)
    left = numpy.zeros(2 * width * width, dtype=numpy.float32)
    m = numpy.float32(mean)
    if n_avail <= 0:                                   # extraction_context.cpp:43-47
        return -1, above, left
    total = above_units + left_units + 1
    if n_avail == total:                               # extraction_context.cpp:56-90
        p = origin - width * stride - width
        for i in range(width):
            above[i * cw:(i + 1) * cw] = plane_i32[p:p + cw].astype(numpy.float32) - m
            p += stride
        p = origin - width
        for i in range(2 * width):
            left[i * width:(i + 1) * width] = plane_i32[p:p + width].astype(numpy.float32) - m
            p += stride
        return 0, above, left
    # partially available: zero fill, then the above-left W x W block unconditionally (119-127)
    p = origin - width * stride - width
    for i in range(width):
        above[i * cw:i * cw + width] = plane_i32[p:p + width].astype(numpy.float32) - m
        p += stride
    if not flags[left_units]:                          # extraction_context.cpp:133-138
        return -1, above, left
    for i in range(above_units):                       # extraction_context.cpp:149-166
        if flags[left_units + 1 + i]:
            dst = width + i * unit_w
            p = origin - width * stride + i * unit_w
            for j in range(width):
                above[dst:dst + unit_w] = plane_i32[p:p + unit_w].astype(numpy.float32) - m
                dst += cw
                p += stride
    # left + below-left: neither the source nor the destination pointer advances on an
    # unavailable unit (extraction_context.cpp:189-205)
    p = origin - width
    dst = 0
    for i in range(left_units):
        if flags[left_units - 1 - i]:
            for _ in range(unit_h):
                left[dst:dst + width] = plane_i32[p:p + width].astype(numpy.float32) - m
                dst += width
                p += stride
    return 0, above, left
