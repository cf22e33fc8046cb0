"""k-slab halo layer: replaces DAGlobalToLocal/DALocalToLocal between ranks on this path.

One process per GPU; torch.distributed is the plumbing (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The C library calls back into `TorchHalo.exchange(scalar_ids)` whenever the
reference would refresh ghosts (Source/momentum.c:2293, rhs.c:251,290,690,746, les.c:254-267,675,
1026,1320 ...).  k is the slowest index, so the G ghost planes of one scalar are one contiguous
block; the requested scalars are gathered into one packed buffer per neighbour, i.e. each exchange
is 2 sends + 2 receives regardless of the number of fields.
"""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist


class _CudaBlob:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class TorchHalo:
    def __init__(self, rank, world, periodic_k, device):
        self.rank, self.world, self.per, self.device = rank, world, bool(periodic_k), torch.device(device)
        self.lo = rank - 1 if rank > 0 else (world - 1 if self.per else None)
        self.hi = rank + 1 if rank < world - 1 else (0 if self.per else None)
        self.pool = None
        self.nexchanges = 0
        self.bytes_sent = 0

    def attach(self, ctx):
        self.ctx = ctx
        n = ctx.nscalars * ctx.scalar_len
        base = ctx.scalar_ptr(0)
        if self.device.type == "cuda":
            ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            flat = torch.as_tensor(_CudaBlob(base, n), device=self.device)
        else:
            arr = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_double)), shape=(n,))
            flat = torch.from_numpy(arr)
        G, plane = ctx.G, ctx.plane
        self.pool = flat.view(ctx.nscalars, ctx.nzt, plane)
        self.G = G
        ctx.set_halo_callback(self.exchange)

    def _tail_view(self, sid):
        """A scalar outside the main pool (allocated on first use: Ucont_rm1, Adv1-3, the clark gradient planes)."""
        ctx = self.ctx
        ptr = ctx.scalar_ptr(sid)
        if not ptr:
            raise RuntimeError("halo exchange of scalar %d before its allocation" % sid)
        if self.device.type == "cuda":
            flat = torch.as_tensor(_CudaBlob(ptr, ctx.scalar_len), device=self.device)
        else:
            flat = torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(ctx.scalar_len,)))
        return flat.view(ctx.nzt, ctx.plane)

    def _exchange_tail(self, ids):
        """Same exchange for ids that include scalars outside the pool: gathered view by view."""
        G = self.G
        views = [self.pool[i] if i < self.pool.shape[0] else self._tail_view(i) for i in ids]
        nzt = views[0].shape[0]
        nlo, nhi = self.ctx.halo_layers()
        ops, recvs = [], []
        if self.hi is not None:
            sb = torch.stack([v[nzt - G - nlo:nzt - G] for v in views]).contiguous()
            ops.append(dist.P2POp(dist.isend, sb, self.hi)); self.bytes_sent += sb.numel() * 8
        if self.lo is not None:
            sb2 = torch.stack([v[G:G + nhi] for v in views]).contiguous()
            ops.append(dist.P2POp(dist.isend, sb2, self.lo)); self.bytes_sent += sb2.numel() * 8
        if self.lo is not None:
            rb = torch.empty((len(ids), nlo, views[0].shape[1]), dtype=views[0].dtype, device=views[0].device)
            ops.append(dist.P2POp(dist.irecv, rb, self.lo)); recvs.append((rb, slice(G - nlo, G)))
        if self.hi is not None:
            rb2 = torch.empty((len(ids), nhi, views[0].shape[1]), dtype=views[0].dtype, device=views[0].device)
            ops.append(dist.P2POp(dist.irecv, rb2, self.hi)); recvs.append((rb2, slice(nzt - G, nzt - G + nhi)))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for rb, sl in recvs:
            for q, v in enumerate(views):
                v[sl] = rb[q]
        self.nexchanges += 1
        return 0

    def exchange(self, ids):
        G, p = self.G, self.pool
        if max(ids) >= p.shape[0]:
            return self._exchange_tail(list(ids))
        nzt = p.shape[1]
        nlo, nhi = self.ctx.halo_layers()          # ghost planes to fill below / above the slab (<= G)
        idx = torch.as_tensor(ids, device=p.device, dtype=torch.long)
        ops, recvs = [], []
        # order matters when lo == hi (2 ranks, periodic): first send pairs with the peer's first recv
        if self.hi is not None:                    # my top nlo owned planes -> the hi neighbour's low ghosts
            sb = p[idx, nzt - G - nlo:nzt - G].contiguous()
            ops.append(dist.P2POp(dist.isend, sb, self.hi))
            self.bytes_sent += sb.numel() * 8
        if self.lo is not None:                    # my bottom nhi owned planes -> the lo neighbour's high ghosts
            sb2 = p[idx, G:G + nhi].contiguous()
            ops.append(dist.P2POp(dist.isend, sb2, self.lo))
            self.bytes_sent += sb2.numel() * 8
        if self.lo is not None:
            rb = torch.empty((len(ids), nlo, p.shape[2]), dtype=p.dtype, device=p.device)
            ops.append(dist.P2POp(dist.irecv, rb, self.lo))
            recvs.append((rb, slice(G - nlo, G)))
        if self.hi is not None:
            rb2 = torch.empty((len(ids), nhi, p.shape[2]), dtype=p.dtype, device=p.device)
            ops.append(dist.P2POp(dist.irecv, rb2, self.hi))
            recvs.append((rb2, slice(nzt - G, nzt - G + nhi)))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for rb, sl in recvs:
            p[idx, sl] = rb
        self.nexchanges += 1
        return 0
