import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import manual_yolo_b200 as m
from manual_yolo_b200 import synth, geometry, api, _lib
dev = torch.device('cuda:0')
B=64
head, _ = synth.synth_head_from_labels(B, 64, seed=0)
head = head.to(dev)
lv = geometry.level_shapes(640, 640)
lib = ctypes.CDLL('/root/repo/scratch/libppdbg.so')
ws = m.Workspace(B, 1024, 300, dev)
dbg = torch.zeros((B,12), dtype=torch.int64, device=dev)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
for it in range(4):
    cands = m.decode_and_filter(head, conf_thres=0.25, level_hw=lv, cap=1024, defer_boxes=True)
    levels, n_levels, *_ = api._head_levels(head, (8,16,32), None, lv)
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.dbg_postprocess_small(levels, n_levels, ctypes.c_void_p(cands.rows.data_ptr()), ctypes.c_void_p(cands.anchor.data_ptr()), ctypes.c_void_p(cands.count.data_ptr()), B, 1024, 30000, ctypes.c_double(0.45), ctypes.c_float(7680.0), 0, 300, None, ctypes.c_void_p(ws.det.rows.data_ptr()), ctypes.c_void_p(ws.det.anchor.data_ptr()), ctypes.c_void_p(ws.det.count.data_ptr()), None, 0, None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.c_void_p(dbg.data_ptr()))
    e1.record(); torch.cuda.synchronize()
    print('rc', rc, 'us', e0.elapsed_time(e1)*1e3)
d = dbg.cpu()
print('phases (cycles since start): load, dfl, sort, boxes, nms, out, end | n')
for b in [0,1,2,3,int(d[:,11].argmax())]:
    print(b, d[b,:8].tolist(), int(d[b,11]))
print('mean', d[:, :8].float().mean(0).tolist())
