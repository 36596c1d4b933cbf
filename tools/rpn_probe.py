"""cfg3 RPN stage timing: one CTA per image (mode 1) vs cluster of 8 CTAs per image (mode 2), plus the cluster kernel's phase clocks."""
import sys, os, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heltondetection_b200 import synth, rpn, _lib

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
obj, dlt, bases, _ = synth.rpn_heads(B, 832, G=20, seed=1237)
obj, dlt = [o.cuda() for o in obj], [d.cuda() for d in dlt]
print("clusters resident at once (size 8/4/2/1):", [_lib.lib().hd_rpn_cluster_capacity(c) for c in (8, 4, 2, 1)])
res = {}
for mode in (1, 8, 4, 0):
    cl = {8: 8, 4: 4, 22: 2, 11: 1}.get(mode, 0)
    _lib.lib().hd_rpn_set_cluster_size(cl)
    rpn.set_mode(mode if mode in (0, 1) else 2)
    pr = rpn.RpnProposals(bases, (4, 8, 16, 32), (832, 832), n_pre_nms=12000, n_post_nms=2000, min_size=16)
    t = timeit(lambda: pr(obj, dlt))
    td = timeit(lambda: pr.decode(obj, dlt))
    rois, cnt, sc, idx = pr(obj, dlt)
    res[mode] = (rois.clone(), cnt.clone(), sc.clone(), idx.clone())
    print(f"mode {'single-CTA' if mode == 1 else 'auto' if mode in (0, 2) else 'cluster x%d' % cl}: decode+select+nms {t:8.1f} us  (decode alone {td:6.1f} us)  B={B}  kept {cnt.tolist()[:4]}...", flush=True)
    a = (C.c_longlong * 16)()
    _lib.check(_lib.lib().hd_debug_phases(1, a))
    v = list(a)
    names = ["start", "select", "compaction", "sort+merge+gather", "grid", "adjacency", "resolve", "outputs"] if mode != 1 else None
    if names:
        for i in range(1, 8):
            if v[i] and v[i - 1]: print(f"     {names[i]:20s} +{(v[i] - v[i - 1]) / 1.9e3:7.1f} us")
print("identical outputs:", all(all(torch.equal(x, y) for x, y in zip(res[1], res[m])) for m in res))
_lib.lib().hd_rpn_set_cluster_size(0)
rpn.set_mode(0)
