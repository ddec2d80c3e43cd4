"""Test helper: the collectives of pgrc_b200.matcher.TorchComm between several matcher contexts that live in ONE process
(one thread per "rank", all on one GPU or on the CPU model).  Lets the GPU tests drive the product's own sharded
orchestration (run_plan_sharded / run_plan_routed / merge_accumulators) without a second GPU."""
import threading

import torch


class LocalWorld:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world, timeout=300)
        self.slots = [None] * world
        self.result = None

    def comm(self, rank: int) -> "LocalComm":
        return LocalComm(self, rank)

    def run(self, fn):
        """fn(rank, comm) on one thread per rank; returns the list of results, re-raises the first exception."""
        out, err = [None] * self.world, []

        def body(r):
            try:
                out[r] = fn(r, self.comm(r))
            except BaseException as e:  # noqa: BLE001 - must release the other ranks
                err.append(e)
                self.barrier.abort()

        th = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            real = [e for e in err if not isinstance(e, threading.BrokenBarrierError)]
            raise (real or err)[0]
        return out


class LocalComm:
    def __init__(self, w: LocalWorld, rank: int):
        self.w, self.rank, self.world = w, rank, w.world

    def _sync(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.w.barrier.wait()

    def all_reduce(self, t, op: str):
        w = self.w
        w.slots[self.rank] = t
        self._sync()
        if self.rank == 0:
            st = torch.stack([x for x in w.slots])
            w.result = {"min": lambda: st.min(dim=0).values, "max": lambda: st.max(dim=0).values,
                        "sum": lambda: st.sum(dim=0).to(t.dtype)}[op]()
        self._sync()
        t.copy_(w.result)
        self._sync()

    def exchange_counts(self, counts, device):
        w = self.w
        w.slots[self.rank] = list(counts)
        self._sync()
        res = [int(w.slots[s][self.rank]) for s in range(self.world)]
        self._sync()
        return res

    def sibling(self):
        return self

    def all_gather_bytes(self, blob, device):
        w = self.w
        w.slots[self.rank] = bytes(blob)
        self._sync()
        res = list(w.slots)
        self._sync()
        return res

    def all_to_all_async(self, recv, send):
        self.all_to_all(recv, send)

        class _Done:
            def wait(self):
                pass
        return _Done()

    def all_to_all(self, recv, send):
        w = self.w
        w.slots[self.rank] = send
        self._sync()
        for s in range(self.world):
            if recv[s] is not None:
                recv[s].copy_(w.slots[s][self.rank])
        self._sync()
