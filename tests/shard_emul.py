"""Torch-CPU emulation of the csrc/shard.cu kernels -- TEST INFRASTRUCTURE.  It lets the world_size-2
gloo tests run the host protocol of reviews4rec_b200/sharded.py (buffer layouts, counts, scaling,
collectives) without a GPU, and serves as the checker of the real kernels in the GPU tests."""
import torch


class CpuKernels:
    @staticmethod
    def mark(idx, V, flags):
        assert int(idx.min()) >= 0 and int(idx.max()) < V
        flags[idx.reshape(-1)] = 1

    @staticmethod
    def plan(flags, V, P, cap, req):
        for o in range(P):
            local = torch.nonzero(flags[o::P]).reshape(-1)
            n = local.numel()
            req[o, 0] = n
            req[o, 1:1 + n] = local                   # the device appends in arbitrary order; any order is valid
        flags.zero_()

    @staticmethod
    def bucket(ids, R, P, cap, req, pos):
        req.zero_()
        for i, v in enumerate(ids.reshape(-1).tolist()):
            assert 0 <= v < R
            o = v % P
            k = int(req[o, 0])
            req[o, 1 + k] = v // P
            req[o, 0] = k + 1
            pos[i] = o * cap + k

    @staticmethod
    def serve(shard, rreq, P, cap, out):
        flat = shard.reshape(shard.shape[0], -1)
        o = out.reshape(P, cap, -1)
        for q in range(P):
            n = int(rreq[q, 0])
            assert 0 <= n <= cap
            o[q, :n] = flat[rreq[q, 1:1 + n]]

    @staticmethod
    def place(rows, req, P, cap, cache, V):
        r = rows.reshape(P, cap, -1)
        c = cache.reshape(cache.shape[0], -1)
        for q in range(P):
            n = int(req[q, 0])
            ids = req[q, 1:1 + n] * P + q
            assert n == 0 or (int(ids.min()) >= 0 and int(ids.max()) < V)
            c[ids] = r[q, :n]

    @staticmethod
    def gather(table, pos, out):
        out.copy_(table.reshape(table.shape[0], -1)[pos].reshape(out.shape))

    @staticmethod
    def scatter_unique(gout, pos, send):
        send.index_add_(0, pos, gout.reshape(pos.numel(), -1))

    @staticmethod
    def scatter_owner(grads, rreq, P, cap, gtable, scale):
        g = gtable.reshape(gtable.shape[0], -1)
        for q in range(P):
            n = int(rreq[q, 0])
            g.index_add_(0, rreq[q, 1:1 + n], grads[q * cap:q * cap + n] * scale)
