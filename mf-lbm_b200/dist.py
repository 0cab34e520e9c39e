"""torch.distributed plumbing shared by bench.py and the multi-rank tests: one process per GPU / per z-slab.

Only rendezvous, scalar reductions and the broadcast of the 128-byte NCCL id go through torch.distributed (backend
"nccl" on the GPU box, "gloo" in CPU tests); the halo exchange itself is issued by libmflbm.so on its own NCCL
communicator (mflbm_create with use_nccl=1), which replaces the reference's MPI_CART_CREATE ring (MP/Mpi_misc.F90:19-38).
"""
import os


class Ranks:
    def __init__(self, backend=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
                kw["device_id"] = self.device
            dist.init_process_group(backend, **kw)
            self.dist = dist
        self.backend = backend

    def _tensor(self, values, dtype):
        import torch
        return torch.tensor(values, dtype=dtype, device=self.device if self.device is not None else "cpu")

    def allreduce(self, x, op="sum"):
        """float64 scalar reduction over the slabs (pore counts, max time over ranks)"""
        if self.dist is None:
            return float(x)
        import torch
        t = self._tensor([float(x)], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
        return float(t.item())

    def allgather(self, x):
        """one float64 per rank -> list over ranks (per-rank fluid-node counts and step times)"""
        if self.dist is None:
            return [float(x)]
        import torch
        t = self._tensor([0.0] * self.world, torch.float64)
        t[self.rank] = float(x)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu().tolist()]

    def broadcast_bytes(self, payload, nbytes, src=0):
        """rank src's bytes object to every rank (the NCCL unique id of the slab ring)"""
        if self.dist is None:
            return payload
        import torch
        t = self._tensor(list(payload) if self.rank == src else [0] * nbytes, torch.uint8)
        self.dist.broadcast(t, src)
        return bytes(t.cpu().tolist())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None


def slab_partition(nz_global, npz):
    """Equal z slabs like the reference requires (MP/Mpi_misc.F90:64-73 with mod(nzGlobal,npz)==0):
    returns [(k_first, k_last)] in global 1-based planes for idz = 0..npz-1."""
    if nz_global % npz:
        raise ValueError("nzGlobal must be a multiple of npz (template-simulation_control.txt:105-106)")
    nz = nz_global // npz
    return [(r * nz + 1, (r + 1) * nz) for r in range(npz)]
