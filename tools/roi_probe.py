"""cfg3 multi-level RoIAlign timing: gather kernels (mode 1) vs staged-row TMA ring kernel (mode 2)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, roi

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

B, img = (int(sys.argv[1]) if len(sys.argv) > 1 else 16), 832
obj, dlt, bases, _ = synth.rpn_heads(B, img, G=20, seed=1237)
feats = [f.cuda().contiguous(memory_format=torch.channels_last) for f in synth.fpn_features(B, img, 256, 1237)]
obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (img, img), n_pre_nms=12000, n_post_nms=2000, min_size=16)
rois, cnt, sc, idx = pr(obj, dlt)
scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32]
fbytes = sum(f.numel() * 4 for f in feats); obytes = rois.shape[0] * 256 * 49 * 4
res = {}
for sr in (2, 1):
    for mode in (1, 2):
        roi.set_mode(mode)
        t = timeit(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, sr, False))
        out, lv = roi.multilevel_roi_align(feats, rois, 7, scales, sr, False)
        res[(sr, mode)] = out
        print(f"sr={sr} mode {mode} ({'gather' if mode == 1 else 'staged rows'}): {t:8.1f} us  {(fbytes + obytes) / t / 1e3:7.1f} GB/s algorithmic", flush=True)
    print("   identical:", torch.equal(res[(sr, 1)], res[(sr, 2)]))
for dbg, name in ((1, "no tile store"), (2, "no compute"), (3, "no store, no compute (row streaming + sync only)"), (8, "nothing staged (direct loads in the persistent kernel)")):
    roi.set_mode(2 | (dbg << 4))
    t = timeit(lambda: roi.multilevel_roi_align(feats, rois, 7, scales, 2, False))
    print(f"   staged-row kernel, {name}: {t:8.1f} us", flush=True)
roi.set_mode(0)
