"""CPU, world_size 2 over gloo: image sharding + the padded all-gather reproduce the single-process result.
The per-rank compute in this test is the oracle (test infrastructure) -- what is under test is the host
logic of heltondetection_b200/dist.py."""
import os
import socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_slice_partitions_exactly():
    from heltondetection_b200.dist import shard_slice
    for n in (0, 1, 7, 8, 256, 257):
        for w in (1, 2, 3, 8):
            sl = [shard_slice(n, r, w) for r in range(w)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1
    assert shard_slice(256, 3, 8) == slice(96, 128)


def test_pack_unpack_roundtrip():
    from heltondetection_b200.dist import pack_records, unpack_records
    det = torch.randn((5, 7, 6))
    cnt = torch.tensor([0, 7, 3, 1, 2], dtype=torch.int32)
    d2, c2 = unpack_records(pack_records(det, cnt), 7)
    assert torch.equal(d2, det) and torch.equal(c2, cnt)


def _padded_oracle(heads, max_det):
    import oracle
    out = oracle.yolo.non_max_suppression(oracle.yolo.decode_box(heads), 0.25, 0.45, max_det=max_det)
    det = torch.zeros((len(out), max_det, 6))
    cnt = torch.zeros((len(out),), dtype=torch.int32)
    for b, o in enumerate(out):
        det[b, : o.shape[0]] = o
        cnt[b] = o.shape[0]
    return det, cnt


def _worker(rank, world, port, n_images, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from heltondetection_b200 import synth
        from heltondetection_b200.dist import shard_slice, DetectionGather, gather_ragged
        torch.set_num_threads(1)
        heads, _ = synth.yolo_heads(n_images, 128, 4, 3, 99)
        sl = shard_slice(n_images, rank, world)
        det, cnt = _padded_oracle([h[sl] for h in heads], 20)
        if n_images % world == 0:
            gd, gc = DetectionGather(sl.stop - sl.start, 20, torch.device("cpu"))(det, cnt)
        else:
            gd, gc = gather_ragged(det, cnt, n_images)
        if rank == 0:
            torch.save((gd.clone(), gc.clone()), out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [6, 5])
def test_two_rank_gather_equals_single_process(n_images, tmp_path):
    from heltondetection_b200 import synth
    out_path = str(tmp_path / "gathered.pt")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    gd, gc = torch.load(out_path)
    heads, _ = synth.yolo_heads(n_images, 128, 4, 3, 99)
    det, cnt = _padded_oracle(heads, 20)
    assert torch.equal(gc, cnt) and torch.equal(gd, det)
    assert int(cnt.sum()) > 0
