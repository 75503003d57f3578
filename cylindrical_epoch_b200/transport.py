"""torch.distributed plumbing for the x-neighbour exchange (host logic).

`TorchRing.sendrecv` is the host-side equivalent of the two MPI_SENDRECVs the reference issues
per halo (boundary.F90:528,541): send one buffer to the left neighbour and one to the right,
receive one from each.  It works on CPU tensors over gloo (tests) and on CUDA tensors over
NCCL; `make_callback` adapts it to the C-ABI's CALLBACK transport by wrapping the raw device
pointers the library passes.  (The default multi-GPU path is the library's own dlopen()ed
NCCL transport; this one exists so that the exchange protocol can be driven and tested from
the host, and for launchers that already own a process group.)
"""
import ctypes as C

import torch
import torch.distributed as dist


def neighbours(rank, nranks, periodic):
    """Cartesian-communicator neighbours of an x-slab (mpi_routines.F90:186-227): -1 = MPI_PROC_NULL."""
    left = rank - 1 if rank - 1 >= 0 else (nranks - 1 if periodic else -1)
    right = rank + 1 if rank + 1 < nranks else (0 if periodic else -1)
    return left, right


class TorchRing:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)

    def sendrecv(self, left, right, send_left, recv_left, send_right, recv_right):
        """Tensors may be None (no neighbour / nothing to move).  A left-going message always
        lands in the receiver's `recv_right`, a right-going one in `recv_left`; tags keep the
        two apart when left == right (2 ranks, periodic ring)."""
        TAG_LEFTGOING, TAG_RIGHTGOING = 11, 12
        if left == self.rank or right == self.rank:   # periodic wrap onto myself
            if send_left is not None and recv_right is not None:
                recv_right.copy_(send_left)
            if send_right is not None and recv_left is not None:
                recv_left.copy_(send_right)
            return
        ops = []
        if left >= 0 and send_left is not None and send_left.numel():
            ops.append(dist.P2POp(dist.isend, send_left, left, self.group, TAG_LEFTGOING))
        if right >= 0 and send_right is not None and send_right.numel():
            ops.append(dist.P2POp(dist.isend, send_right, right, self.group, TAG_RIGHTGOING))
        if right >= 0 and recv_right is not None and recv_right.numel():
            ops.append(dist.P2POp(dist.irecv, recv_right, right, self.group, TAG_LEFTGOING))
        if left >= 0 and recv_left is not None and recv_left.numel():
            ops.append(dist.P2POp(dist.irecv, recv_left, left, self.group, TAG_RIGHTGOING))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def exchange_counts_then_payload(self, left, right, payload_left, payload_right, width=7):
        """partlist_sendrecv (partlist.F90:822-876): counts first, then count*width doubles.
        Returns (from_left, from_right) float64 tensors of shape (n, width)."""
        dev = payload_left.device
        cnt_s = torch.tensor([payload_left.shape[0], payload_right.shape[0]], dtype=torch.int64, device=dev)
        cnt_r = torch.zeros(2, dtype=torch.int64, device=dev)
        self.sendrecv(left, right, cnt_s[0:1], cnt_r[0:1], cnt_s[1:2], cnt_r[1:2])
        n_from_left = int(cnt_r[0]) if left >= 0 else 0
        n_from_right = int(cnt_r[1]) if right >= 0 else 0
        rl = torch.empty((n_from_left, width), dtype=torch.float64, device=dev)
        rr = torch.empty((n_from_right, width), dtype=torch.float64, device=dev)
        self.sendrecv(left, right, payload_left.contiguous(), rl, payload_right.contiguous(), rr)
        return rl, rr


class _DevBuf:
    """raw device pointer -> __cuda_array_interface__ so torch can wrap it without a copy"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _wrap(ptr, nbytes):
    if not ptr or not nbytes:
        return None
    return torch.as_tensor(_DevBuf(ptr, nbytes), device="cuda")


def make_callback(ring):
    """Python function with the cylgpu_sendrecv_fn signature (see include/cylgpu.h)."""

    def cb(user, left, right, sl, sl_b, rl, rl_b, sr, sr_b, rr, rr_b, stream):
        try:
            ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                ring.sendrecv(left, right, _wrap(sl, sl_b), _wrap(rl, rl_b), _wrap(sr, sr_b), _wrap(rr, rr_b))
            return 0
        except Exception as e:   # noqa: BLE001
            print("cylgpu sendrecv callback failed:", e)
            return 1

    return cb
