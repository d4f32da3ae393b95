"""tracker association costs (detect.py:557 -> supervision ByteTrack iou_distance / fuse_score)
against the numpy restatement in oracle/assoc.py (supervision itself is not installed: parity unpinned)."""
import numpy as np
import pytest
import torch

import manual_yolo_b200 as m
from manual_yolo_b200 import handoff
from oracle import assoc as oassoc

pytestmark = pytest.mark.gpu


def test_iou_cost_matrix_bit_exact_and_assignment(cuda_dev):
    g = torch.Generator().manual_seed(2)
    B, T, max_det = 3, 40, 300
    rows = torch.zeros((B, max_det, 6))
    xy = torch.rand((B, max_det, 2), generator=g) * 1500
    wh = torch.rand((B, max_det, 2), generator=g) * 120 + 10
    rows[..., :2], rows[..., 2:4] = xy, xy + wh
    rows[..., 4] = torch.rand((B, max_det), generator=g)
    rows[..., 5] = torch.randint(0, 64, (B, max_det), generator=g).float()
    dcount = torch.tensor([300, 0, 37], dtype=torch.int32)
    tcount = torch.tensor([40, 5, 12], dtype=torch.int32)
    # tracks = jittered copies of some detections (so that real matches exist) + degenerate / identical boxes
    tracks = rows[:, :T, :4].clone() + torch.randn((B, T, 4), generator=g) * 3
    tracks[0, 0] = rows[0, 0, :4]                       # identical box: iou 1
    tracks[0, 1] = torch.tensor([5., 5., 5., 5.])       # zero area, disjoint
    rows[2, 3, :4] = torch.tensor([7., 7., 7., 7.]); tracks[2, 3] = torch.tensor([7., 7., 7., 7.])   # 0/0 -> NaN
    det = m.Detections(rows.to(cuda_dev), torch.zeros((B, max_det), dtype=torch.int32, device=cuda_dev), dcount.to(cuda_dev))
    for fuse in (False, True):
        cost = m.iou_cost_matrix(tracks.to(cuda_dev), tcount.to(cuda_dev), det, fuse_score=fuse, pad_cost=2.0).cpu().numpy()
        for b in range(B):
            nt, nd = int(tcount[b]), int(dcount[b])
            ref = oassoc.iou_cost_ref(tracks[b, :nt].numpy(), rows[b, :nd].numpy(), fuse_score=fuse)
            got = cost[b, :nt, :nd]
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])          # bit-exact fp32
            assert (cost[b, nt:, :] == 2.0).all() and (cost[b, :, nd:] == 2.0).all()
    cost = m.iou_cost_matrix(tracks.to(cuda_dev), tcount.to(cuda_dev), det).cpu().numpy()
    matches, ut, ud = handoff.associate(cost[0, :40, :300], thresh=0.8)
    assert len(matches) >= 30 and all(int(r) == int(c) for r, c in matches)           # jittered copies find their source
    assert 1 in ut and len(ud) == 300 - len(matches)
    mt, ut2, ud2 = handoff.associate(cost[1, :5, :0], thresh=0.8)
    assert len(mt) == 0 and ut2 == [0, 1, 2, 3, 4] and ud2 == []
    with pytest.raises(ValueError):
        m.iou_cost_matrix(tracks, tcount, det)                                          # CPU tensors are refused
