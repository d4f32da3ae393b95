"""Host-side geometry of the detection path (pure Python, no device work).

Mirrors, argument for argument, the arithmetic Ultralytics 8.3.176 does on the host around the
kernels (the reference reaches it through ``model(frame)``, ``/root/reference/detect.py:541``):
``LetterBox.__call__`` sizes and pads, ``ops.scale_boxes`` gain/pad, feature-map shapes per stride.
Python ``round`` (half to even) and true division are part of the specification.
"""

from __future__ import annotations


def letterbox_geometry(shape_hw, new_shape=(640, 640), auto=False, scale_fill=False, scaleup=True,
                       center=True, stride=32):
    """``ultralytics/data/augment.py::LetterBox.__call__`` geometry for one (h, w) source shape."""
    h, w = int(shape_hw[0]), int(shape_hw[1])
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    elif scale_fill:
        dw, dh = 0.0, 0.0
        new_w, new_h = new_shape[1], new_shape[0]
    if center:
        dw /= 2
        dh /= 2
    top, bottom = (int(round(dh - 0.1)) if center else 0), int(round(dh + 0.1))
    left, right = (int(round(dw - 0.1)) if center else 0), int(round(dw + 0.1))
    return dict(new_w=new_w, new_h=new_h, top=top, bottom=bottom, left=left, right=right,
                out_h=new_h + top + bottom, out_w=new_w + left + right, ratio=r)


def referenced_rows(src_h, new_h):
    """Source rows the vertical pass of ``cv2.resize(INTER_LINEAR)`` reads for src_h -> new_h, as
    (row0, row_step, n_rows).  cv2 maps output row y to ``fy = (y + 0.5) * (src_h / new_h) - 0.5``; for an
    odd integer scale k this is the integer ``k*y + (k-1)/2`` exactly, the fractional weight is 0 and the
    second tap contributes nothing (coefficients 2048 / 0): one row in k is read.  Every other scale
    reads (nearly) all rows, reported as the full range."""
    src_h, new_h = int(src_h), int(new_h)
    if new_h > 0 and src_h % new_h == 0 and (src_h // new_h) % 2 == 1 and src_h // new_h >= 3:
        k = src_h // new_h
        return (k - 1) // 2, k, new_h
    return 0, 1, src_h


def slice_boxes(image_h, image_w, slice_h=640, slice_w=640, overlap_h=0.2, overlap_w=0.2):
    """Windows of SAHI-style sliced prediction as ``sahi.slicing.get_slice_bboxes`` lays them out (the reference
    calls ``get_sliced_prediction(frame, slice_height=640, slice_width=640, overlap_*_ratio=0.2)``,
    ``pipe.py:183-194``): row-major scan with steps of ``slice - int(overlap * slice)``; a window that would
    cross the right/bottom edge is shifted back inside, so every window has the same size
    ``(min(slice_h, image_h), min(slice_w, image_w))``.  Returns ``[[x_min, y_min, x_max, y_max], ...]``.
    sahi is not installed in this container: restated from its published algorithm (parity unpinned)."""
    image_h, image_w, slice_h, slice_w = int(image_h), int(image_w), int(slice_h), int(slice_w)
    y_overlap, x_overlap = int(overlap_h * slice_h), int(overlap_w * slice_w)
    if slice_h <= y_overlap or slice_w <= x_overlap:
        raise ValueError("overlap must be smaller than the slice")
    out = []
    y_max = y_min = 0
    while y_max < image_h:
        x_min = x_max = 0
        y_max = y_min + slice_h
        while x_max < image_w:
            x_max = x_min + slice_w
            if y_max > image_h or x_max > image_w:
                xmax, ymax = min(image_w, x_max), min(image_h, y_max)
                out.append([max(0, xmax - slice_w), max(0, ymax - slice_h), xmax, ymax])
            else:
                out.append([x_min, y_min, x_max, y_max])
            x_min = x_max - x_overlap
        y_min = y_max - y_overlap
    return out


def scale_boxes_params(img1_shape, img0_shape, ratio_pad=None):
    """``ops.scale_boxes`` gain and (pad_x, pad_y) for letterboxed shape img1 -> source shape img0."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
               round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return gain, pad


def level_shapes(in_h, in_w, strides=(8, 16, 32)):
    """Detect-head feature-map (h, w) per stride for a letterboxed input (in_h, in_w)."""
    return [(in_h // int(s), in_w // int(s)) for s in strides]


def num_anchors(in_h, in_w, strides=(8, 16, 32)):
    return sum(h * w for h, w in level_shapes(in_h, in_w, strides))


def class_mask_words(classes, nc):
    """Allow-list of class ids -> little-endian uint32 bit words (nc bits)."""
    words = [0] * ((nc + 31) // 32)
    for c in classes:
        c = int(c)
        if 0 <= c < nc:
            words[c >> 5] |= 1 << (c & 31)
    return words


def shard_range(n_items, rank, world_size):
    """Static block partition of frame indices: rank g of G takes [g*N/G, (g+1)*N/G)."""
    return (rank * n_items) // world_size, ((rank + 1) * n_items) // world_size
