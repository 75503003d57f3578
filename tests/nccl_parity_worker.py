"""Launched by test_gpu_nccl.py under torch.distributed.run with 2+ ranks, one GPU each:
the library's own NCCL transport (dlopen()ed libnccl, unique id broadcast by the launcher)
against the oracle run with the same number of x-slabs."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import decks  # noqa: E402
from cylindrical_epoch_b200 import _lib  # noqa: E402
from cylindrical_epoch_b200.constants import FIELD_NAMES, TRANSPORT_CALLBACK, TRANSPORT_NCCL  # noqa: E402
from parity import TOL, TOL_HOT, by_weight, J_FLOOR  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    L = _lib.load()
    def transport_kw():
        if mode == "nccl":   # an ncclUniqueId is good for ONE communicator: a fresh one per handle
            buf = ctypes.create_string_buffer(128)
            if rank == 0:
                assert L.cylgpu_nccl_unique_id(buf) == 0, L.cylgpu_last_error()
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            return dict(transport=TRANSPORT_NCCL, nccl_unique_id=bytes(t.cpu().numpy().tobytes()))
        from cylindrical_epoch_b200.transport import TorchRing, make_callback
        return dict(transport=TRANSPORT_CALLBACK, sendrecv=make_callback(TorchRing()))

    failures = []
    for deckname in ("lwfa", "thermal", "window"):
        kw = transport_kw()
        d = {"lwfa": lambda: decks.lwfa(nx=96, ny=24, n_mode=2, ppc_e=4),
             "thermal": lambda: decks.thermal(nx=64, ny=24, ppc=6),
             # C3's shape in small: laser + plasma + moving window (columns inserted on the last rank, dropped on the first)
             "window": lambda: decks.lwfa(nx=64, ny=24, n_mode=2, ppc_e=4, ppc_p=1, window=True, t_centre=30e-15)}[deckname]()
        tol = TOL_HOT if deckname == "thermal" else (TOL if deckname == "lwfa" else 1e-9)
        w = decks.make_oracle(d, nranks=world)        # every process steps the whole oracle world
        s = decks.make_slab(d, rank=rank, nranks=world, device=local, **kw)
        if mode == "nccl" and os.environ.get("CYLGPU_P2P", "0") in ("1", "particles") and deckname != "thermal":
            # (the thermal ring of two ranks has left == right: mailboxes as well; open decks: one link per rank)
            info = s.transport_info()
            if not (info[1] or info[2]):
                failures.append(f"{deckname}: peer-memory mailboxes not mapped on rank {rank}: {info}")
        decks.copy_state(w, s, rank)
        s.rng_set_state(*w.rng_state(rank))           # the loader's stream continues in the product (window columns)
        w.call("init_half_step")
        s.init_half_step()
        for _ in range(30 if deckname == "window" else 12):
            w.call("step")
            s.step_once()
        st, ref = s.stats(), w.stats(rank)
        if (st.n_sent_left, st.n_sent_right, st.n_removed, st.n_recv) != (
                ref["sent_left"], ref["sent_right"], ref["removed"], ref["received"]):
            failures.append(f"{deckname}: migration counts differ on rank {rank}")
        qnc = sum(abs(sp.charge) * sp.density for sp in d.species) * 2.99792458e8
        for name in FIELD_NAMES:
            a = w.field(rank, name)
            den = max(np.abs(w.field(k, name)).max() for k in range(world))
            if name.startswith("j"):
                den = max(den, J_FLOOR * qnc)
            err = np.abs(s.download_field(name) - a).max()
            if den > 0 and err / den > tol:
                failures.append(f"{deckname}: {name} rel err {err / den:.3e} on rank {rank}")
        for isp in range(len(d.species)):
            got, a = by_weight(s.download_particles(isp)), by_weight(w.particles(rank, isp))
            if got.shape != a.shape:
                failures.append(f"{deckname}: species {isp} particle count {got.shape[0]} != {a.shape[0]} on rank {rank}")
            elif a.size and np.abs(got[:, :6] - a[:, :6]).max() > tol * np.abs(a[:, :3]).max() + tol * np.abs(a[:, 3:6]).max():
                failures.append(f"{deckname}: species {isp} particle phase space differs on rank {rank}")
        s.close()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    for f in failures:
        print("FAIL", f)
    if rank == 0:
        print("NCCL_PARITY_OK" if int(flag) == 0 else "NCCL_PARITY_FAILED", mode, "world", world)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 0 else 1)


if __name__ == "__main__":
    main()
