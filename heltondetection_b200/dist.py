"""Image sharding + the one collective of the path: an all-gather of the padded per-image detections
(SURVEY.md 8e).  Every stage is per-image independent, so ranks never exchange data until the final
fixed-size record {int32 count, float32 [max_det, 6]} per image is gathered over NCCL (NVLink 5 / NVSwitch).
The host logic is backend-agnostic so it is tested on CPU with gloo, world_size 2.
"""
import torch
import torch.distributed as dist


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE pinned host buffers are allocated (first touch
    then places them on that node), so that zero-copy reads and H2D copies of 8 ranks do not all cross the socket interconnect.
    Best effort: returns the node id (sysfs), "nvml:<first>-<last cpu>" when only NVML knows the GPU's CPU set, or None when the
    topology cannot be read (then nothing is changed)."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:  # noqa: BLE001
        pass
    # sysfs hides the node (containers / VMs report -1): ask NVML for the GPU's ideal CPU set instead
    try:
        import pynvml
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0) & cpus
        if len(allowed) >= 4 and len(allowed) < len(os.sched_getaffinity(0)):   # (never squeeze a rank onto a handful of CPUs)
            os.sched_setaffinity(0, allowed)
            return f"nvml:{min(allowed)}-{max(allowed)}"
    except Exception:  # noqa: BLE001
        pass
    return None


def shard_slice(n_images, rank, world):
    """Contiguous image slice of `rank`: the first (n % world) ranks take one extra image."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def pack_records(det, count, out=None):
    """det [B,max_det,6] f32 + count [B] i32 -> records [B, 1 + max_det*6] f32 (count bits in column 0)."""
    B, md = det.shape[0], det.shape[1]
    if out is None:
        out = torch.empty((B, 1 + md * 6), dtype=torch.float32, device=det.device)
    out[:, 1:].copy_(det.reshape(B, md * 6))
    out[:, 0].view(torch.int32).copy_(count.to(torch.int32))
    return out


def unpack_records(rec, max_det):
    count = rec[:, 0].contiguous().view(torch.int32)
    return rec[:, 1:].reshape(rec.shape[0], max_det, 6), count


class DetectionGather:
    """Preallocated all-gather of the padded detections of equally sized shards.

    On CUDA the collective runs on a side stream with double-buffered records, so it overlaps the next batch's
    kernels (the payload is ~1.8 MB per 256 images: latency bound, not bandwidth bound).  __call__ returns the
    buffer slot; result(slot) makes the current stream wait for that gather; finish() joins the side stream."""

    def __init__(self, B, max_det, device, group=None):
        self.group, self.max_det = group, max_det
        self.world = dist.get_world_size(group)
        self.cuda = torch.device(device).type == "cuda"
        nbuf = 2 if self.cuda else 1
        self.rec = [torch.empty((B, 1 + max_det * 6), dtype=torch.float32, device=device) for _ in range(nbuf)]
        self.out = [torch.empty((self.world * B, 1 + max_det * 6), dtype=torch.float32, device=device) for _ in range(nbuf)]
        self.k = 0
        if self.cuda:
            self.side = torch.cuda.Stream(device)
            self.ready = [torch.cuda.Event() for _ in range(nbuf)]
            self.done = [torch.cuda.Event() for _ in range(nbuf)]

    def __call__(self, det, count):
        if not self.cuda:
            pack_records(det, count, self.rec[0])
            chunks = list(self.out[0].chunk(self.world, 0))
            dist.all_gather(chunks, self.rec[0], group=self.group)
            return unpack_records(self.out[0], self.max_det)
        i = self.k & 1
        self.k += 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.done[i])          # slot i is free again (no-op the first two times)
        pack_records(det, count, self.rec[i])
        self.ready[i].record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready[i])
            dist.all_gather_into_tensor(self.out[i], self.rec[i], group=self.group)
            self.done[i].record(self.side)
        return i

    def result(self, slot):
        torch.cuda.current_stream().wait_event(self.done[slot])
        return unpack_records(self.out[slot], self.max_det)

    def finish(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.side)


def gather_ragged(det, count, n_images, group=None):
    """All-gather for uneven shards (n_images % world != 0): pads every shard to the largest one,
    gathers, and strips the padding so the result is in global image order."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_slice(n_images, r, world) for r in range(world)]
    bmax = max(s.stop - s.start for s in sizes)
    md = det.shape[1]
    rec = torch.zeros((bmax, 1 + md * 6), dtype=torch.float32, device=det.device)
    pack_records(det, count, rec[: det.shape[0]])
    outs = [torch.empty_like(rec) for _ in range(world)]
    dist.all_gather(outs, rec, group=group)
    full = torch.cat([o[: s.stop - s.start] for o, s in zip(outs, sizes)], 0)
    return unpack_records(full, md)


class PeerDetectionBuffers:
    """Gather buffers in symmetric (peer-mapped) memory: det [slots, world*B, max_det, 6] + count [slots, world*B] on
    every rank.  The NMS kernels of rank r write image b of their shard to row r*B+b of EVERY rank's buffer, so after a
    step (and a cross-GPU barrier) each rank holds the gathered detections without pack kernels or an NCCL collective.
    `slots` buffers alternate between steps so a consumer can still read step k while step k+1 is being written."""

    def __init__(self, B, max_det, device, group=None, slots=2):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.B, self.max_det, self.slots = B, max_det, slots
        if self.world - 1 > _lib.HD_MAX_REPLICAS:
            raise RuntimeError("too many peers for hd_replicas")
        n = self.world * B
        self.det = symm.empty((slots, n, max_det, 6), dtype=torch.float32, device=device)
        self.count = symm.empty((slots, n), dtype=torch.int32, device=device)
        self.det.zero_()
        self.count.zero_()
        self.h_det = symm.rendezvous(self.det, self.group.group_name)
        self.h_cnt = symm.rendezvous(self.count, self.group.group_name)
        self._rep = []
        for s in range(slots):
            r = _lib.Replicas()
            k = 0
            for p in range(self.world):
                if p == self.rank:
                    continue
                r.det[k] = self.h_det.buffer_ptrs[p] + ((s * n + self.rank * B) * max_det * 6) * 4
                r.count[k] = self.h_cnt.buffer_ptrs[p] + (s * n + self.rank * B) * 4
                k += 1
            r.n = k
            self._rep.append(r)

    def local(self, slot):
        """this rank's slice of its own gather buffer: (det [B,max_det,6], count [B])"""
        lo = self.rank * self.B
        return self.det[slot, lo:lo + self.B], self.count[slot, lo:lo + self.B]

    def replicas(self, slot):
        import os
        if os.environ.get("HD_PIPE_NO_REPLICA"):    # developer ablation switch: local outputs only
            return None
        return self._rep[slot]

    def gathered(self, slot):
        """(det [world*B,max_det,6], count [world*B]) -- complete once every rank's step has finished (see barrier())"""
        return self.det[slot], self.count[slot]

    def barrier(self, channel=0):
        """cross-GPU barrier on the current stream: afterwards the stores of all ranks' previous kernels have landed.
        Barriers that may be in flight at the same time (pipelined steps on different streams) must use different channels."""
        self.h_det.barrier(channel=channel)
