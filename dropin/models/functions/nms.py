from planerecnet_b200.postprocess import mask_nms, matrix_nms as _matrix_nms, point_nms  # noqa: F401


def matrix_nms(cate_labels, seg_masks, sum_masks, cate_scores, sigma=2.0, kernel="gaussian"):
    """Reference signature (nms.py:15): seg_masks [n, h, w] bool."""
    n = len(cate_labels)
    if n == 0:
        return []
    return _matrix_nms(cate_labels, seg_masks.reshape(n, -1).float(), sum_masks, cate_scores, sigma, kernel)
