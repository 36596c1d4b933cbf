#!/usr/bin/env python
"""2+ GPU check (torchrun): the NMS kernels' replicated outputs (posted NVLink stores into symmetric memory) give every
rank exactly what pack + NCCL all_gather gives.  Prints PEER_GATHER_OK on every rank."""
import os
import sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, yolo  # noqa: E402
from heltondetection_b200 import dist as hd_dist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
B, MD = 8, 300
heads, _ = synth.yolo_heads(B, 640, 80, 20, 4321 + rank)
heads = [h.to(dev) for h in heads]
# reference: local post-process + NCCL all-gather of the packed records
pp = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, max_det=MD, device=dev)
det, cnt, _ = pp(heads)
g = hd_dist.DetectionGather(B, MD, dev)
slot = g(det, cnt)
ref_det, ref_cnt = g.result(slot)
ref_det, ref_cnt = ref_det.clone(), ref_cnt.clone()
# fused: kernels write straight into every rank's gather buffer
peer = hd_dist.PeerDetectionBuffers(B, MD, dev)
pp2 = yolo.YoloPostprocessor(conf_thres=0.25, iou_thres=0.45, max_det=MD, device=dev)
for s in (0, 1, 0):
    replay, d, c, _ = pp2.graph(heads, peer=peer, slot=s)
    replay()
    peer.barrier()
    torch.cuda.synchronize()
    gd, gc = peer.gathered(s)
    assert torch.equal(gc, ref_cnt), (rank, s, gc.tolist()[:16], ref_cnt.tolist()[:16])
    m = torch.arange(MD, device=dev)[None, :] < ref_cnt[:, None]
    assert torch.equal(gd[m], ref_det[m]), (rank, s)
print(rank, "PEER_GATHER_OK", int(ref_cnt.sum()), "detections gathered from", world, "ranks", flush=True)
dist.barrier()
dist.destroy_process_group()
