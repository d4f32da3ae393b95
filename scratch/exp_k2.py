import sys, torch, statistics
sys.path.insert(0, '/root/repo')
import manual_yolo_b200 as m
from manual_yolo_b200 import synth, geometry
dev = torch.device('cuda:0')
head, _ = synth.synth_head_from_labels(64, 64, seed=0)
head = head.to(dev)
lv = geometry.level_shapes(640, 640)
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)
for conf in (0.25, 0.9999):
    c = m.decode_and_filter(head, conf_thres=conf, level_hw=lv)
    ts = []
    for i in range(13):
        flush.zero_()
        c.count.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m.decode_and_filter(head, conf_thres=conf, level_hw=lv, out=c); e1.record()
        torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    print('conf', conf, 'cands', int(c.count.sum()), 'us', statistics.median(ts), min(ts))
