"""Host-side description of the t-slab decomposition -- the same arithmetic the
library applies internally (csrc/api.cu make_geom / exchange_halo_buffers), exposed so that callers that keep
rank-local host arrays (one process per GPU) know which global time-slices they own, and so that the ring
logic can be tested without a GPU."""
import numpy as np


class SlabDecomposition:
    def __init__(self, nt, world, rank):
        if nt % world != 0:
            raise ValueError("NT must be divisible by the number of GPUs (t-slab decomposition)")
        if world > 1 and nt // world < 2:
            raise ValueError("each t-slab needs at least 2 time-slices")
        self.nt, self.world, self.rank = int(nt), int(world), int(rank)
        self.tloc = nt // world

    def t_range(self):
        return self.rank * self.tloc, (self.rank + 1) * self.tloc

    def neighbours(self):
        """(previous, next) rank on the periodic t ring."""
        return (self.rank - 1) % self.world, (self.rank + 1) % self.world

    def exchange_halo(self, local, dist):
        """local: array (4, tloc, ...).  Returns (up, dn): the next rank's first slice and the previous rank's
        last slice, exchanged with the same send/recv pairing the library posts to NCCL."""
        import torch

        prev, nxt = self.neighbours()
        first = torch.from_numpy(np.ascontiguousarray(local[:, 0]).view(np.float64).copy())
        last = torch.from_numpy(np.ascontiguousarray(local[:, -1]).view(np.float64).copy())
        up = torch.empty_like(first)
        dn = torch.empty_like(last)
        if self.world == 1:
            up.copy_(first)
            dn.copy_(last)
        else:
            reqs = [dist.isend(first, prev, tag=1), dist.isend(last, nxt, tag=2), dist.irecv(up, nxt, tag=1), dist.irecv(dn, prev, tag=2)]
            for r in reqs:
                r.wait()
        shape = local[:, 0].shape
        return up.numpy().view(np.complex128).reshape(shape), dn.numpy().view(np.complex128).reshape(shape)

    @staticmethod
    def local_plaquette_sum(ext):
        """sum over the owned sites of Re tr P for ext = [t-1 halo, owned slices..., t+1 halo] (host layout
        (4, tloc+2, NZ, NY, NX, 3, 3) with (column,row) matrix axes).  Spatial directions wrap periodically."""
        m = np.swapaxes(ext, -1, -2)
        axis = {0: 3, 1: 2, 2: 1}  # mu -> array axis of m[mu] (t, z, y, x, i, j)
        tot = 0.0

        def shift(a, mu):
            if mu == 3:
                return a[2:]  # t+1 for the owned slices 1..tloc
            return np.roll(a, -1, axis=axis[mu])[1:-1]

        for mu in range(4):
            for nu in range(mu + 1, 4):
                a = m[mu][1:-1] @ shift(m[nu], mu)
                b = m[nu][1:-1] @ shift(m[mu], nu)
                tot += float(np.sum(a * b.conj()).real)
        return tot

    def ordered_sum(self, value, dist):
        """all-gather one scalar per rank and add in rank order on every rank (deterministic)."""
        import torch

        if self.world == 1:
            return float(value)
        mine = torch.tensor([value], dtype=torch.float64)
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(self.world)]
        dist.all_gather(parts, mine)
        tot = 0.0
        for p in parts:
            tot += float(p.item())
        return tot
