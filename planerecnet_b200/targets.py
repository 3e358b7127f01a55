"""Ground-truth assignment of the SOLOv2-style heads (models/functions/losses.py:200-275), device-resident.

The reference sends every image's plane masks to the host, rescales them by 1/4 with cv2 (bilinear, uint8) and sends
them back (losses.py:243-247).  For binary masks and an exact 1/4 scale that resize has a closed form — the output pixel is
1 iff at least two of the four centre pixels of its 4x4 block are set (cv2's fixed-point bilinear with weights 1/2, 1/2
rounds 0.5 up; checked against cv2 in tests/test_targets_cpu.py) — so the whole assignment runs where the masks live, with
one host round trip per FPN level for the handful of integer cell coordinates.  Plain tensor indexing only (no arithmetic
hot path); device-agnostic, which is what lets the CPU suite compare it with the oracle bit for bit.

Device independence: the reference's `(c / extent) // (1 / S)` sits exactly on grid-cell boundaries for round inputs, and
torch's CUDA kernels divide by Python scalars through a reciprocal (last-bit differences), so an earlier version of this
module resolved one synthetic case differently on CUDA than on the CPU.  The moments are now accumulated exactly (float64)
and every division uses device-tensor divisors (IEEE quotients on both devices).  tests/test_loss_kernels_gpu.py asserts on the
B200 that the CUDA assignment equals the CPU one and the unmodified reference's (tests/golden/loss_golden.pt), bit for bit.
`assign_targets_batch` does a whole batch with a single host round trip."""
import torch


_SCALARS = {}


def _dev_scalar(value, dtype, dev):
    """A 0-dim device tensor holding `value`, cached: creating one costs a synchronous host-to-device copy, and the target
    assignment needs the same dozen divisors every step."""
    key = (float(value), dtype, str(dev))
    t = _SCALARS.get(key)
    if t is None:
        t = _SCALARS[key] = torch.tensor(float(value), dtype=dtype, device=dev)
    return t


def quarter_masks(masks):
    """[n, H, W] {0,1} (any integer / bool dtype), H and W multiples of 4 -> [n, H/4, W/4] uint8, identical to
    cv2.resize(..., fx=fy=0.25, INTER_LINEAR) of the uint8 masks."""
    n, H, W = masks.shape
    if H % 4 or W % 4:
        raise ValueError("quarter_masks needs H and W to be multiples of 4 (got %dx%d)" % (H, W))
    m = masks.to(torch.uint8).reshape(n, H // 4, 4, W // 4, 4)
    centre = m[:, :, 1:3, :, 1:3].sum(dim=(2, 4))
    return (centre >= 2).to(torch.uint8)


@torch.no_grad()
def assign_targets(gt, feat_hw, num_grids, scale_ranges, num_classes=2, sigma=0.2):
    """losses.py:200-275 for one image.  gt: {'boxes' [n,4] xyxy, 'classes' [n], 'masks' [n,H,W]} on any device.
    Returns per level (ins_label uint8 [m, fh, fw], cate_label int64 [S, S], ins_ind bool [S*S], grid_order list[int])."""
    boxes, labels, masks = gt["boxes"], gt["classes"], gt["masks"]
    dev = masks.device
    fh, fw = feat_hw
    up_h, up_w = fh * 4, fw * 4
    areas = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    out = []
    for (lo, hi), S in zip(scale_ranges, num_grids):
        hit = ((areas >= lo) & (areas <= hi)).nonzero().flatten()
        cate = torch.full((S, S), num_classes, dtype=torch.int64, device=dev)
        ind = torch.zeros(S * S, dtype=torch.bool, device=dev)
        if hit.numel() == 0:
            out.append((torch.zeros(0, fh, fw, dtype=torch.uint8, device=dev), cate, ind, []))
            continue
        bx, lb, mk = boxes[hit], labels[hit], masks[hit]
        half_w = 0.5 * (bx[:, 2] - bx[:, 0]) * sigma
        half_h = 0.5 * (bx[:, 3] - bx[:, 1]) * sigma
        # centre of mass (funcs.py:213-224).  The moments are sums of up to 3e5 pixel coordinates: accumulated in float64
        # they are exact and therefore identical on every device, whereas the reference's float32 sums depend on the
        # reduction order (its own CPU and CUDA runs can put a centre on different sides of a grid-cell boundary).
        ys = torch.arange(mk.shape[1], dtype=torch.float64, device=dev)
        xs = torch.arange(mk.shape[2], dtype=torch.float64, device=dev)
        mk64 = mk.to(torch.float64)
        m00 = mk64.sum(-1).sum(-1).clamp(min=1e-6)
        cw = ((mk64 * xs).sum(-1).sum(-1) / m00).to(torch.float32)
        ch = ((mk64 * ys[:, None]).sum(-1).sum(-1) / m00).to(torch.float32)
        nonempty = mk.sum(-1).sum(-1) > 0

        def cell(v, extent):                       # int((v / extent) // (1 / S)) of the reference, for all instances at once
            # divisors as device tensors of v's dtype: with Python-scalar divisors torch's CUDA kernels take a
            # multiply-by-reciprocal shortcut that may differ from the IEEE quotient in the last bit — enough to move a
            # centre that sits exactly on a grid-cell boundary (0.5 // 0.025) to the other cell than the CPU run
            e = _dev_scalar(extent, v.dtype, dev)
            pitch = _dev_scalar(1.0 / S, v.dtype, dev)
            return torch.floor_divide(v / e, pitch).to(torch.int64)

        cx, cy = cell(cw, up_w), cell(ch, up_h)
        top = torch.maximum(cell(ch - half_h, up_h).clamp(min=0), cy - 1)
        down = torch.minimum(cell(ch + half_h, up_h).clamp(max=S - 1), cy + 1)
        left = torch.maximum(cx - 1, cell(cw - half_w, up_w).clamp(min=0))
        right = torch.minimum(cell(cw + half_w, up_w).clamp(max=S - 1), cx + 1)
        # the one host round trip of this level: a few integers per instance
        rows = torch.stack([top, down, left, right, nonempty.to(torch.int64), lb.to(torch.int64)], 1).tolist()
        src, order = [], []
        for k, (t, d, l, r, ok, lab) in enumerate(rows):
            if not ok:
                continue
            cate[t:d + 1, l:r + 1] = lab
            for i in range(t, d + 1):
                for j in range(l, r + 1):
                    src.append(k)
                    order.append(i * S + j)
        if order:
            small = quarter_masks(mk)
            canvas = torch.zeros(len(order), fh, fw, dtype=torch.uint8, device=dev)
            canvas[:, :small.shape[1], :small.shape[2]] = small[torch.tensor(src, device=dev)]
            ind[torch.tensor(order, device=dev)] = True
            ins = canvas
        else:
            ins = torch.zeros(0, fh, fw, dtype=torch.uint8, device=dev)
        out.append((ins, cate, ind, order))
    return out


@torch.no_grad()
def assign_targets_batch(gts, feat_hw, num_grids, scale_ranges, num_classes=2, sigma=0.2):
    """assign_targets for every image of a batch with ONE host round trip (the per-image version makes one per image and FPN
    level: 32 for a batch of 8) and a handful of batched tensor ops: the per-instance scalars of all images (box scale, centre
    of mass, cell ranges for every level) are computed together, fetched once, and the label maps / cell lists are assembled
    on the host and uploaded in two copies.  Returns [assign_targets(gt, ...) for gt in gts], element for element."""
    import numpy as np
    dev = gts[0]["masks"].device
    fh, fw = feat_hw
    up_h, up_w = fh * 4, fw * 4
    n_img = [int(g["boxes"].shape[0]) for g in gts]
    boxes = torch.cat([g["boxes"] for g in gts])
    labels = torch.cat([g["classes"] for g in gts]).to(torch.int64)
    areas = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    half_w = 0.5 * (boxes[:, 2] - boxes[:, 0]) * sigma
    half_h = 0.5 * (boxes[:, 3] - boxes[:, 1]) * sigma
    cws, chs, nes = [], [], []
    for g in gts:                                      # exact float64 moments per image (see assign_targets)
        mk = g["masks"]
        ys = torch.arange(mk.shape[1], dtype=torch.float64, device=dev)
        xs = torch.arange(mk.shape[2], dtype=torch.float64, device=dev)
        mk64 = mk.to(torch.float64)
        m00 = mk64.sum(-1).sum(-1).clamp(min=1e-6)
        cws.append(((mk64 * xs).sum(-1).sum(-1) / m00).to(torch.float32))
        chs.append(((mk64 * ys[:, None]).sum(-1).sum(-1) / m00).to(torch.float32))
        nes.append(mk.sum(-1).sum(-1) > 0)
    cw, ch, nonempty = torch.cat(cws), torch.cat(chs), torch.cat(nes)
    rows = []
    for (lo, hi), S in zip(scale_ranges, num_grids):
        def cell(v, extent):
            e = _dev_scalar(extent, v.dtype, dev)
            pitch = _dev_scalar(1.0 / S, v.dtype, dev)
            return torch.floor_divide(v / e, pitch).to(torch.int64)

        hit = (areas >= lo) & (areas <= hi)
        cx, cy = cell(cw, up_w), cell(ch, up_h)
        top = torch.maximum(cell(ch - half_h, up_h).clamp(min=0), cy - 1)
        down = torch.minimum(cell(ch + half_h, up_h).clamp(max=S - 1), cy + 1)
        left = torch.maximum(cx - 1, cell(cw - half_w, up_w).clamp(min=0))
        right = torch.minimum(cell(cw + half_w, up_w).clamp(max=S - 1), cx + 1)
        rows.append(torch.stack([hit.to(torch.int64), top, down, left, right, nonempty.to(torch.int64), labels], 1))
    table = torch.stack(rows).tolist()                 # the one host round trip: [levels][instances][7]
    total_cells = sum(S * S for S in num_grids)
    cate_np = np.full((len(gts), total_cells), num_classes, dtype=np.int64)
    ind_np = np.zeros((len(gts), total_cells), dtype=np.bool_)
    plan, src_all = [], []
    start = 0
    for b, n in enumerate(n_img):
        off = 0
        per_level = []
        for lvl, S in enumerate(num_grids):
            cate = cate_np[b, off:off + S * S].reshape(S, S)
            src, order = [], []
            any_hit = False
            for k in range(n):
                h, t, d, l, r, ok, lab = table[lvl][start + k]
                if not h:
                    continue
                any_hit = True
                if not ok:
                    continue
                cate[t:d + 1, l:r + 1] = lab
                for i in range(t, d + 1):
                    for j in range(l, r + 1):
                        src.append(k)
                        order.append(i * S + j)
            if order:
                ind_np[b, off + np.asarray(order)] = True
            per_level.append((off, S, order, len(src_all), len(src), any_hit))
            src_all += src
            off += S * S
        plan.append(per_level)
        start += n
    cate_t = torch.from_numpy(cate_np).to(dev)
    ind_t = torch.from_numpy(ind_np).to(dev)
    src_t = torch.tensor(src_all, dtype=torch.int64, device=dev) if src_all else None
    # dense batch form for the dice / lava node (losses._InsLava): the positive rows of image b — level after level, the order
    # of the per-level lists — are rows [0, n_b) of ONE [B, n, fh * fw] target tensor; gidx[b, r] is the row's cell in the
    # level-concatenated kernel map (level offset + cell).  The per-level tensors handed out below are views of it.
    n_b = [sum(sn for (_o, _S, _ord, _s0, sn, _a) in plan[b]) for b in range(len(gts))]
    n_rows = max(16, (max(n_b) + 15) // 16 * 16)
    dense_tgt = torch.zeros(len(gts), n_rows, fh, fw, dtype=torch.uint8, device=dev)
    gidx_np = np.zeros((len(gts), n_rows), dtype=np.int64)
    valid_np = np.zeros((len(gts), n_rows), dtype=np.bool_)
    out = TargetsBatch()
    for b, g in enumerate(gts):
        if n_b[b]:
            small = quarter_masks(g["masks"])
            s_lo = plan[b][0][3]
            dense_tgt[b, :n_b[b], :small.shape[1], :small.shape[2]] = small[src_t[s_lo:s_lo + n_b[b]]]
            valid_np[b, :n_b[b]] = True
        res, row = [], 0
        for (off, S, order, s0, sn, _any) in plan[b]:
            cate = cate_t[b, off:off + S * S].view(S, S)
            ind = ind_t[b, off:off + S * S]
            if sn:
                gidx_np[b, row:row + sn] = off + np.asarray(order, dtype=np.int64)
            res.append((dense_tgt[b, row:row + sn], cate, ind, order))
            row += sn
        out.append(res)
    out.dense = dict(tgt=dense_tgt.view(len(gts), n_rows, fh * fw), gidx=torch.from_numpy(gidx_np).to(dev, non_blocking=True),
                     valid=torch.from_numpy(valid_np).to(dev, non_blocking=True), n_b=n_b, n=n_rows, cate=cate_t,
                     level_off=[p_[0] for p_ in plan[0]] if plan else [], num_grids=list(num_grids))
    return out


class TargetsBatch(list):
    """[per image][per level] (ins_label, cate_label, ins_ind, grid_order) like a list of assign_targets results, plus `.dense`: the
    same targets in the batch layout the loss node consumes without a per-(image, level) loop."""
    dense = None
