import sys; sys.path.insert(0,'/root/repo')
import torch, oracle
from heltondetection_b200 import synth, yolo
heads,_=synth.yolo_heads(2,640,80,20,77)
pred=oracle.yolo.decode_box(heads)
ref,ridx=oracle.yolo.non_max_suppression(pred,0.001,0.6,return_index=True,multi_label=True)
pp=yolo.YoloPostprocessor(conf_thres=0.001,iou_thres=0.6,multi_label=True,dense_read=True)
det,cnt,idx=pp([h.cuda() for h in heads])
for b in range(2):
    n=int(cnt[b]); g=idx[b,:n].cpu(); r=ridx[b]
    print('image',b,'n',n,'ref n',r.numel())
    m=min(n,r.numel()); d=(g[:m]!=r[:m]).nonzero().flatten()
    print(' first diffs at', d[:10].tolist())
    for p in d[:3].tolist():
        print('  pos',p,'got',int(g[p]),float(det[b,p,4]),'ref',int(r[p]),float(ref[b][p,4]))
    print(' set diff', len(set(g.tolist())-set(r.tolist())), len(set(r.tolist())-set(g.tolist())))
# candidate counts
x,ai=oracle.yolo.filter_candidates_multi_label(pred[0],0.001)
print('oracle candidates img0', x.shape[0])
pp2=yolo.YoloPostprocessor(conf_thres=0.001,iou_thres=0.6,multi_label=True,one_call=False)
pp2([h.cuda() for h in heads]); print('gpu candidate counts', pp2._buf.count.tolist(), 'cap', pp2._buf.cap)
