"""One process per GPU: row partitioning, comm schedule and halo exchange (replaces the reference's MPI layer,
code/mpi_funcs.hpp + SpmvKernel::{init,finalize}_halo_exchange, classes_structs.hpp:857-995).

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests); the pack
kernel and the SpMV kernels are the library's own.  Per SpMV (comm_halos = 1):

    comm stream : pack (one launch for all peers) -> isend/irecv per neighbour, receiving IN PLACE at the tail of x
    main stream : interior chunks (no halo column)  ... wait for the exchange ... boundary chunks

The reference does begin -> finish -> execute with no overlap (main.cpp:464-468).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import call, vp


def seg_rows_equal(n_rows_total: int, world: int) -> np.ndarray:
    """seg-rows work_sharing_arr for a matrix whose last row is non-empty (mpi_funcs.hpp:446-465)."""
    per = n_rows_total // world
    wsa = np.arange(world + 1, dtype=np.int64) * per
    wsa[world] = n_rows_total
    return wsa.astype(np.int32)


def seg_nnz_from_row_counts(counts, world: int) -> np.ndarray:
    """seg-nnz work_sharing_arr (mpi_funcs.hpp:466-493,602-606) from the per-row element counts of a row-sorted COO instead of
    its row array — the same walk (count nnz // world elements, close the segment after the row holding the NEXT element, restart
    counting after that element), done on row boundaries so that no rank needs the whole COO."""
    counts = np.asarray(counts, np.int64)
    n_rows, nnz = len(counts), int(counts.sum())
    if n_rows < world:
        raise ValueError("seg_work_sharing_arr ERROR: total_mtx->n_rows < comm_size.")
    ptr = np.concatenate(([0], np.cumsum(counts)))
    last_row = int(np.max(np.nonzero(counts)[0]))
    per = nnz // world
    wsa = np.zeros(world + 1, np.int64)
    g, seg = 0, 1
    while True:
        g += per              # element index at which local == per
        if g >= nnz:
            break
        if seg <= world:
            wsa[seg] = int(np.searchsorted(ptr, g, side="right") - 1) + 1  # row of element g, + 1
        seg += 1
        g += 1                # that element is skipped by the `continue`
    wsa[world] = last_row + 1
    if wsa[world - 1] == wsa[world]:
        wsa[1:world] -= 1
    if np.any(np.diff(wsa) < 0):
        raise ValueError("seg_work_sharing_arr ERROR: flaw in work_sharing_arr, work_sharing_arr[i] < work_sharing_arr[i-1].")
    return wsa.astype(np.int32)


def bind_to_gpu_numa_node(device_index: int):
    """Restricts this process's CPU affinity to the cores of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards are first-touched next to the GPU's PCIe root (8 ranks staging 268 MB per SpMV through one socket's memory is what
    bounds the host-buffer path otherwise).  Returns {"node", "cpus", "previous"} or None when the topology is not exposed."""
    import os
    import subprocess
    try:
        bdf = None
        try:
            import torch
            pr = torch.cuda.get_device_properties(device_index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except Exception:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().lower()
            bdf = out[-12:] if len(out) >= 12 else None
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        previous = os.sched_getaffinity(0)
        allowed = previous & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed), "previous": previous, "bdf": bdf}
    except Exception:
        return None


def comm_schedule(need_lists, rank: int, world: int, group=None):
    """collect_comm_idxs (mpi_funcs.hpp:117-172): every rank tells every owner which owner-local x indices it needs.
    need_lists[p] = indices this rank needs from owner p.  Returns send_lists[q] = indices rank q needs from us."""
    import torch.distributed as dist
    gathered = [None] * world
    dist.all_gather_object(gathered, [np.asarray(a, np.int32) for a in need_lists], group=group)
    return [np.asarray(gathered[q][rank], np.int32) for q in range(world)]


class HaloExchange:
    """begin/finish_communicate_halo_elements (mpi_funcs.hpp:16-66) on torch tensors (CPU/gloo or CUDA/nccl).
    sendbuf holds the packed elements for all peers back to back (send_ptr); the halo of peer p is received in place
    at x[n_local + recv_cumsum[p] : n_local + recv_cumsum[p+1]]."""

    def __init__(self, rank, world, n_local, recv_cumsum, send_ptr, group=None):
        self.rank, self.world, self.n_local, self.group = rank, world, int(n_local), group
        self.recv_cumsum = [int(v) for v in recv_cumsum]
        self.send_ptr = [int(v) for v in send_ptr]
        self.non_zero_senders = [p for p in range(world) if self.recv_cumsum[p + 1] > self.recv_cumsum[p]]
        self.non_zero_receivers = [p for p in range(world) if self.send_ptr[p + 1] > self.send_ptr[p]]

    def begin(self, x, sendbuf):
        import torch.distributed as dist
        ops = []
        for p in self.non_zero_senders:
            ops.append(dist.P2POp(dist.irecv, x[self.n_local + self.recv_cumsum[p]: self.n_local + self.recv_cumsum[p + 1]], p, self.group))
        for p in self.non_zero_receivers:
            ops.append(dist.P2POp(dist.isend, sendbuf[self.send_ptr[p]: self.send_ptr[p + 1]], p, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs):
        for r in reqs:
            r.wait()


class HaloPlan:
    """Device-side collect_local_needed_heri + comm schedule for one rank."""

    def __init__(self, scs, wsa, rank, world, group=None):
        from .engine import _hp
        self.scs, self.rank, self.world = scs, rank, world
        wsa = np.ascontiguousarray(wsa, np.int32)
        h = vp()
        call("uspmv_halo_plan_create", scs.h, _hp(wsa), int(rank), int(world), C.byref(h))
        self.h = h
        cum = np.zeros(world + 1, np.int32)
        nh = C.c_long(0)
        call("uspmv_halo_plan_counts", self.h, _hp(cum), C.byref(nh))
        self.recv_cumsum, self.n_halo = cum, int(nh.value)
        flat = np.zeros(max(self.n_halo, 1), np.int32)
        ptr = np.zeros(world + 1, np.int32)
        call("uspmv_halo_plan_need", self.h, _hp(flat), _hp(ptr))
        self.need_lists = [flat[ptr[p]:ptr[p + 1]].copy() for p in range(world)]
        self.send_lists = None
        self.send_ptr = None

    def set_send(self, send_lists):
        from .engine import _hp
        self.send_lists = send_lists
        ptr = np.cumsum([0] + [len(a) for a in send_lists]).astype(np.int32)
        flat = np.ascontiguousarray(np.concatenate(send_lists) if ptr[-1] else np.zeros(1, np.int32), np.int32)
        call("uspmv_halo_plan_set_send", self.h, _hp(flat), _hp(ptr))
        self.send_ptr = ptr
        self.n_send = int(ptr[-1])

    def pack(self, x, sendbuf, bvs=1, vec_length=0, layout=capi.COLWISE):
        from .engine import _dp, _stream
        vt = {8: capi.F64, 4: capi.F32, 2: capi.F16}[x.element_size()]
        call("uspmv_halo_pack", self.h, _dp(x), _dp(sendbuf), vt, int(bvs), int(vec_length), int(layout), _stream())

    def __del__(self):
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_halo_destroy(self.h)
            self.h = None


class _DevArray:
    """Zero-copy torch view of library-owned device memory (the P2P arena's x)."""

    def __init__(self, ptr, n, np_dtype):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": np.dtype(np_dtype).str, "data": (int(ptr), False), "version": 3}


class P2PHalo:
    """NVLink peer-to-peer exchange (uspmv_p2p_*): arena with n_buf buffers (each bvs vectors of vec_length elements) + epoch
    flags, IPC handles all-gathered over torch.distributed at setup; no collective call inside the SpMV loop."""

    def __init__(self, plan: HaloPlan, vt: int, vec_length: int, rank: int, world: int, group=None, bvs: int = 1,
                 layout: int = capi.COLWISE, n_buf: int = 1):
        import torch
        import torch.distributed as dist
        from .engine import NP_OF
        self.plan, self.bvs, self.layout, self.n_buf, self.vec_length = plan, int(bvs), int(layout), int(n_buf), int(vec_length)
        self.group = group
        h = vp()
        xptr = (vp * n_buf)()
        handle = (C.c_ubyte * 64)()
        call("uspmv_p2p_create_ex", plan.h, int(vt), int(vec_length), int(bvs), int(layout), int(n_buf), C.byref(h), handle, xptr)
        self.h = h
        es = {capi.F64: 8, capi.F32: 4, capi.F16: 2}[vt]
        x_len = vec_length * bvs
        x_bytes = (x_len * es + 255) // 256 * 256
        info = {"handle": bytes(handle), "x_bytes": x_bytes, "n_local": plan.scs.n_rows, "recv_cumsum": [int(v) for v in plan.recv_cumsum],
                "vec_length": int(vec_length), "n_buf": int(n_buf), "bvs": int(bvs), "layout": int(layout)}
        infos = [None] * world
        dist.all_gather_object(infos, info, group=group)
        if any((i["n_buf"], i["bvs"], i["layout"]) != (n_buf, bvs, layout) for i in infos):
            raise ValueError("every rank must create the P2P arena with the same n_buf / block_vec_size / layout")
        handles = b"".join(i["handle"] for i in infos)
        peer_x_bytes = (C.c_long * world)(*[i["x_bytes"] for i in infos])
        peer_base = (C.c_long * world)(*[i["n_local"] + i["recv_cumsum"][rank] for i in infos])
        peer_ld = (C.c_long * world)(*[i["vec_length"] for i in infos])
        call("uspmv_p2p_connect_ex", self.h, handles, peer_x_bytes, peer_base, peer_ld)
        dev = f"cuda:{plan.scs.ctx.device}"
        self.bufs = [torch.as_tensor(_DevArray(xptr[b], x_len, NP_OF[vt]), device=dev) for b in range(n_buf)]
        self.x = self.bufs[0]
        dist.barrier(group=group)

    def spmv(self, scs, y, main_stream, comm_stream):
        call("uspmv_p2p_spmv", self.h, scs.h, vp(y.data_ptr()), vp(main_stream.cuda_stream), vp(comm_stream.cuda_stream))

    def spmv_buf(self, scs, x_buf, y_buf, main_stream, comm_stream):
        """x = buffer x_buf, y -> buffer y_buf (rows < n_rows): one step of the device-resident solve loop."""
        call("uspmv_p2p_spmv_buf", self.h, scs.h, int(x_buf), int(y_buf), None, vp(main_stream.cuda_stream), vp(comm_stream.cuda_stream))

    def spmmv(self, scs, Y, main_stream, comm_stream, x_buf=0):
        call("uspmv_p2p_spmmv", self.h, scs.h, int(x_buf), vp(Y.data_ptr()), vp(main_stream.cuda_stream), vp(comm_stream.cuda_stream))

    def status(self):
        err, ep = C.c_int(0), C.c_long(0)
        call("uspmv_p2p_status", self.h, C.byref(err), C.byref(ep))
        return int(err.value), int(ep.value)

    def sync(self):
        """Device sync; raises when a bounded flag wait of any step timed out (a neighbour never pushed / acknowledged)."""
        call("uspmv_p2p_sync", self.h)

    def close(self):
        """COLLECTIVE teardown (every rank of the group must call it): the arena is CUDA-IPC-exported and the neighbours' last
        kernels still write acknowledgements into it after this rank's last step has completed, so: device sync -> barrier ->
        close the imported arenas -> barrier -> free.  Raises if the arena's error word was raised."""
        if not getattr(self, "h", None):
            return
        import torch.distributed as dist
        err = None
        try:
            self.sync()
        except capi.UspmvError as e:
            err = e
        if dist.is_initialized():
            dist.barrier(group=self.group)
        capi.lib.uspmv_p2p_disconnect(self.h)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        self.bufs, self.x = [], None
        capi.lib.uspmv_p2p_destroy(self.h)
        self.h = None
        if err is not None:
            raise err

    def __del__(self):
        # not collective: only safe once every rank is idle (close() is the supported path)
        if getattr(self, "h", None) and capi is not None:
            capi.lib.uspmv_p2p_destroy(self.h)
            self.h = None


class DistributedSpmv:
    """bench.py's N > 1 runner (weak scaling): every rank owns an n^3 z-slab of an n x n x (n*world) stencil grid.
    Harness order (main.cpp:1104-1128,1271-1308): slab -> convert_to_scs -> halo discovery -> permute_scs_cols."""

    kernel_name = "k_scs32_stream"

    def __init__(self, ctx, points, n, C_, sigma, vt, rank, world, overlap=True, group=None, halo="p2p", strong=False, bvs=1,
                 layout="colwise", n_buf=1):
        """strong=False: weak scaling, an n^3 slab per rank (grid n x n x n*world); strong=True: ONE n^3 grid cut into
        `world` z-slabs (BASELINE.json config 5: 512^3 27-point over 2/4/8 GPUs).
        bvs > 1: SpMMV over a block vector of bvs right-hand sides in `layout`, all vectors of a neighbour's halo rows exchanged in
        one push (the reference's bulkvec mode).  n_buf = 2: two x buffers for the device-resident solve loop (solve())."""
        import torch
        import torch.distributed as dist
        from . import engine as eng
        self.ctx, self.rank, self.world, self.overlap = ctx, rank, world, overlap
        self.bvs, self.layout = int(bvs), eng.LAYOUT[layout] if isinstance(layout, str) else int(layout)
        if self.bvs > 1 and halo != "p2p":
            raise ValueError("block vectors are exchanged over the P2P arena only")
        if self.bvs > 1 and n_buf != 1:
            raise ValueError("the two-buffer solve loop is SpMV only")
        nz_total = n if strong else n * world
        if strong and n % world:
            raise ValueError("strong scaling needs n divisible by the number of ranks")
        wsa = seg_rows_equal(n * n * nz_total, world)
        self.grid, self.row0, self.row1, self.group = (points, n, n, nz_total), int(wsa[rank]), int(wsa[rank + 1]), group
        mtx = eng.MtxData.stencil(points, n, n, nz_total, int(wsa[rank]), int(wsa[rank + 1]), ctx=ctx)
        self.scs = eng.convert_to_scs(mtx, C_, sigma, vt)
        del mtx
        self.plan = HaloPlan(self.scs, wsa, rank, world)
        eng.permute_scs_cols(self.scs)
        self.plan.set_send(comm_schedule(self.plan.need_lists, rank, world, group))
        ni, nb = C.c_long(0), C.c_long(0)
        call("uspmv_scs_split_chunks", self.scs.h, C.byref(ni), C.byref(nb))
        self.n_interior_chunks, self.n_boundary_chunks = int(ni.value), int(nb.value)
        s = self.scs
        self.nnz, self.n_elements, self.n_chunks = s.nnz, s.n_elements, s.n_chunks
        self.n_rows, self.n_rows_padded, self.n_cols_local, self.n_halo = s.n_rows, s.n_rows_padded, s.n_rows, self.plan.n_halo
        dt = eng.torch_dtype(s.vt)
        dev = f"cuda:{ctx.device}"
        # vector length n_local + max(scs_padding, halo_count) (main.cpp:1405-1420)
        x_len = s.n_rows + max(s.n_rows_padded - s.n_rows, self.n_halo)
        self.vec_length = x_len
        self.halo = halo
        self.p2p = None
        if halo == "p2p":
            self.p2p = P2PHalo(self.plan, s.vt, x_len, rank, world, group, bvs=self.bvs, layout=self.layout, n_buf=n_buf)
            # True / False: fused one-launch step / exchange first; 0, 1, 2: the library's modes (uspmv_p2p_set_overlap)
            call("uspmv_p2p_set_overlap", self.p2p.h, 2 if overlap is True else (0 if overlap is False else int(overlap)))
            self.x = self.p2p.x
            self.x.fill_(5.0)
        else:
            self.x = torch.full((x_len,), 5.0, dtype=dt, device=dev)
        # column-major Y uses the same leading dimension as X (one vec_length for both, like the harness)
        self.y = torch.zeros((x_len if self.layout == capi.COLWISE else s.n_rows_padded) * self.bvs if self.bvs > 1 else s.n_rows_padded,
                             dtype=dt, device=dev)
        self.sendbuf = torch.zeros(max(self.plan.n_send, 1), dtype=dt, device=dev)
        self.ex = HaloExchange(rank, world, s.n_rows, self.plan.recv_cumsum, self.plan.send_ptr, group)
        self.comm_stream = torch.cuda.Stream(device=dev)
        self.e2e_h2d_bytes = s.n_rows * self.x.element_size() * self.bvs
        self.e2e_d2h_bytes = self.y.numel() * self.y.element_size()
        self._torch, self._eng = torch, eng
        if self.bvs > 1:
            self.kernel_name = "k_scs32_stream_mmv"
        dist.barrier()

    def step(self):
        torch, eng = self._torch, self._eng
        main = torch.cuda.current_stream()
        if self.bvs > 1:
            self.p2p.spmmv(self.scs, self.y, main, self.comm_stream)
            return
        if self.p2p is not None:
            self.p2p.spmv(self.scs, self.y, main, self.comm_stream)
            return
        if not self.overlap:
            self.plan.pack(self.x, self.sendbuf)
            HaloExchange.finish(self.ex.begin(self.x, self.sendbuf))
            eng.spmv(self.scs, self.x, self.y)
            return
        self.comm_stream.wait_stream(main)  # x (and last step's boundary SpMV, which reads the halo) are done
        with torch.cuda.stream(self.comm_stream):
            self.plan.pack(self.x, self.sendbuf)
            reqs = self.ex.begin(self.x, self.sendbuf)
        call("uspmv_spmv_part", self.scs.h, 1, vp(self.x.data_ptr()), vp(self.y.data_ptr()), vp(main.cuda_stream))
        with torch.cuda.stream(self.comm_stream):
            HaloExchange.finish(reqs)
        main.wait_stream(self.comm_stream)
        call("uspmv_spmv_part", self.scs.h, 2, vp(self.x.data_ptr()), vp(self.y.data_ptr()), vp(main.cuda_stream))

    def _perm_tensor(self):
        from .validate import device_int_tensor
        return device_int_tensor(self.scs.device_arrays()["old_to_new"].value, self.scs.n_rows, self.x.device).long()

    def set_x(self, xs_of_v, poison_halo=False):
        """x_v (original local row order, one tensor per block-vector column) -> the permuted / laid-out device vector.
        poison_halo: NaN in every halo slot, so that a step whose exchange does not deliver cannot produce a finite y."""
        perm, nl, ld, bvs = self._perm_tensor(), self.scs.n_rows, self.vec_length, self.bvs
        rowwise = bvs > 1 and self.layout == capi.ROWWISE
        nan = float("nan")
        for v in range(bvs):
            xs = xs_of_v(v).to(self.x.dtype)
            if rowwise:
                self.x[: nl * bvs].view(nl, bvs)[perm, v] = xs
            else:
                self.x[v * ld: v * ld + nl][perm] = xs
                if poison_halo:
                    self.x[v * ld + nl: (v + 1) * ld] = nan
        if rowwise and poison_halo:
            self.x[nl * bvs:] = nan

    def get_y(self, v=0):
        """Column v of the result in the original local row order."""
        perm, npad, ld, bvs = self._perm_tensor(), self.scs.n_rows_padded, self.vec_length, self.bvs
        if bvs > 1 and self.layout == capi.ROWWISE:
            return self.y[: npad * bvs].view(npad, bvs)[perm, v]
        return self.y[v * ld: v * ld + npad][perm] if bvs > 1 else self.y[:npad][perm]

    def validate(self, steps: int = 2):
        """Checked steps (COLLECTIVE): x = f(global row), NaN-poisoned halo, `steps` distributed steps (epochs, acks and buffer
        reuse), y of every rank against the stencil formula.  Returns this rank's max relative error |y - ref| / sum|a||x|
        (inf if y holds a NaN or the arena's error word was raised); x is restored to the timing default 5.0 afterwards."""
        import torch
        import torch.distributed as dist
        from . import validate as V
        pts, nx, ny, nz = self.grid
        refs = [V.stencil_product(pts, nx, ny, nz, self.row0, self.row1, v, self.x.device, self.x.dtype) for v in range(self.bvs)]
        self.set_x(lambda v: refs[v][0], poison_halo=True)
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)  # no neighbour pushes into x before it has been poisoned
        ep0 = self.p2p.status()[1] if self.p2p is not None else 0
        for _ in range(steps):
            self.y.fill_(float("nan"))
            self.step()
        torch.cuda.synchronize()
        worst = 0.0
        for v in range(self.bvs):
            worst = max(worst, V.max_rel_err(self.get_y(v), refs[v][1], refs[v][2]))
        self._refs = refs  # the host-buffer path (time_e2e) is checked against the same products
        if self.p2p is not None:
            err, ep = self.p2p.status()
            if err != 0 or ep != ep0 + steps:
                worst = float("inf")
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for b in (self.p2p.bufs if self.p2p is not None else [self.x]):
            b.fill_(5.0)
        self.y.zero_()
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)
        return worst

    def close(self):
        """COLLECTIVE teardown of the P2P arena (see P2PHalo.close)."""
        if self.p2p is not None:
            self.x = None
            self.p2p.close()
            self.p2p = None

    def solve(self, revisions: int, first_buf: int = 0):
        """Solve mode (main.cpp:528-631): `revisions` x { halo exchange ; SpMV ; swap }, device resident: step k reads buffer
        k & 1 of the arena and writes its y (the next x, already in permuted order) into the other buffer, so the swap is free.
        Returns the tensor that holds the final vector (first n_rows entries, permuted order)."""
        if self.p2p is None or self.p2p.n_buf != 2:
            raise ValueError("solve() needs halo='p2p' and n_buf=2")
        main = self._torch.cuda.current_stream()
        b = first_buf
        for _ in range(revisions):
            self.p2p.spmv_buf(self.scs, b, b ^ 1, main, self.comm_stream)
            b ^= 1
        return self.p2p.bufs[b]

    def time_kernel(self, steps):
        torch, eng = self._torch, self._eng
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            if self.bvs > 1:
                eng.spmmv(self.scs, self.x, self.y, self.bvs, self.vec_length, self.layout)
            else:
                eng.spmv(self.scs, self.x, self.y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def time_e2e(self, steps, barrier, pipelined=None):
        """Host-buffer step: H2D of this rank's x slab, halo exchange + SpMV, D2H of y.  pipelined (needs the two-buffer P2P arena,
        single vector): uspmv_p2p_spmv_host_submit / _wait, two calls in flight, each still copying its own x in and its own y out."""
        import time
        torch = self._torch
        n = self.scs.n_rows
        if pipelined is None:
            pipelined = self.p2p is not None and self.p2p.n_buf == 2 and self.bvs == 1
        self.e2e_pipelined = bool(pipelined)
        refs = getattr(self, "_refs", None)
        from . import validate as V
        if pipelined:
            xh = [torch.full((n,), 5.0, dtype=self.x.dtype).pin_memory() for _ in range(2)]
            yh = [torch.zeros(self.scs.n_rows_padded, dtype=self.y.dtype).pin_memory() for _ in range(2)]
            self._e2e_k = getattr(self, "_e2e_k", 0)
            if refs is not None:  # host x = the validation vector (permuted): the y that comes back is checked below
                perm = self._perm_tensor()
                xp = torch.zeros(n, dtype=self.x.dtype, device=self.x.device)
                xp[perm] = refs[0][0].to(self.x.dtype)
                for b in xh:
                    b.copy_(xp)
                del xp

            def run(k):
                for _ in range(k):
                    sl = self._e2e_k & 1
                    call("uspmv_p2p_spmv_host_wait", self.p2p.h, sl)
                    call("uspmv_p2p_spmv_host_submit", self.p2p.h, self.scs.h, vp(xh[sl].data_ptr()), vp(yh[sl].data_ptr()), sl)
                    self._e2e_k += 1
                for sl in range(2):
                    call("uspmv_p2p_spmv_host_wait", self.p2p.h, sl)
            run(2)
            barrier()
            t0 = time.perf_counter()
            run(steps)
            barrier()
            dt = (time.perf_counter() - t0) / steps
            self.e2e_y = yh[(self._e2e_k - 1) & 1]
            if refs is not None:
                self.e2e_max_rel_err = max(V.max_rel_err(b.to(self.x.device)[: self.scs.n_rows_padded][perm], refs[0][1], refs[0][2]) for b in yh)
            return dt
        rowwise = self.bvs > 1 and self.layout == capi.ROWWISE
        n_in = n * self.bvs if (self.bvs == 1 or rowwise) else n  # column-major: the local part of every vector
        xh = torch.full((n_in,), 5.0, dtype=self.x.dtype).pin_memory()
        yh = torch.zeros(self.y.numel(), dtype=self.y.dtype).pin_memory()
        if refs is not None:  # row-major: the validation block vector; column-major: x_0 in every column (one host vector)
            self.set_x((lambda v: refs[v][0]) if rowwise or self.bvs == 1 else (lambda v: refs[0][0]))
            xh.copy_(self.x[:n_in])

        def one():
            if self.bvs > 1 and not rowwise:
                for v in range(self.bvs):
                    self.x[v * self.vec_length: v * self.vec_length + n].copy_(xh, non_blocking=True)
            else:
                self.x[:n_in].copy_(xh, non_blocking=True)
            self.step()
            yh.copy_(self.y, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        barrier()
        dt = (time.perf_counter() - t0) / steps
        self.e2e_y = yh
        if refs is not None:
            self.y.copy_(yh)
            self.e2e_max_rel_err = max(V.max_rel_err(self.get_y(v), refs[v if (rowwise or self.bvs == 1) else 0][1],
                                                     refs[v if (rowwise or self.bvs == 1) else 0][2]) for v in range(self.bvs))
        return dt


class DistributedApSpmv:
    """Row-partitioned adaptive-precision SpMV (BASELINE.json config 4: ap[...] with seg_nnz partitioning).  The reference refuses
    AP under MPI (utilities.hpp:1445-1451); defined here as: contiguous row blocks from work_sharing_arr, per-rank
    partition_precisions, the first part sigma-sorted and the others built on its permutation, ONE halo numbering over all parts
    (uspmv_halo_plan_create_multi), x in the original local row order with the halo at its tail, y in the first part's row order.

    local_coo = (n_local, n_global_cols, I_local, J_global, V) of this rank's rows [wsa[rank], wsa[rank+1])."""

    kernel_name = "k_scs32_stream_ap"

    def __init__(self, ctx, wsa, local_coo, mode, t1, t2, C_, sigma, rank, world, group=None, keep_coo=False):
        import torch
        import torch.distributed as dist
        from . import engine as eng
        self.ctx, self.rank, self.world, self.mode = ctx, rank, world, mode
        if isinstance(local_coo, eng.MtxData):  # already on the device (uspmv_coo_powerlaw)
            n_loc, n_glob, I, J, V = local_coo.n_rows, local_coo.n_cols, None, None, None
        else:
            n_loc, n_glob, I, J, V = local_coo
        mtx = local_coo if isinstance(local_coo, eng.MtxData) else eng.MtxData.from_host(n_loc, n_glob, I, J, V, ctx=ctx)
        self.nnz = mtx.nnz
        self.group, self.row0, self.t1, self.t2 = group, int(wsa[rank]), float(t1), float(t2)
        self._coo = mtx if keep_coo else None   # validate() forms the reference product from the COO triplets
        coos = eng.partition_precisions(mtx, mode, t1, t2)
        del mtx
        self.used = [k for k in range(3) if coos[k] is not None]
        vts = ("dp", "sp", "hp")
        self.parts = [None] * 3
        first = self.used[0]
        self.parts[first] = eng.convert_to_scs(coos[first], C_, sigma, vts[first])
        self.old_to_new = self.parts[first].export().old_to_new
        for k in self.used[1:]:
            self.parts[k] = eng.convert_to_scs(coos[k], C_, sigma, vts[k], fixed_permutation=self.old_to_new)
        del coos
        # one halo plan over all parts; x is NOT permuted for AP
        arr = (vp * len(self.used))(*[self.parts[k].h for k in self.used])
        from .engine import _hp
        wsa = np.ascontiguousarray(wsa, np.int32)
        h = vp()
        call("uspmv_halo_plan_create_multi", arr, len(self.used), _hp(wsa), int(rank), int(world), 0, C.byref(h))
        plan = HaloPlan.__new__(HaloPlan)
        plan.scs, plan.rank, plan.world, plan.h = self.parts[first], rank, world, h
        cum = np.zeros(world + 1, np.int32)
        nh = C.c_long(0)
        call("uspmv_halo_plan_counts", h, _hp(cum), C.byref(nh))
        plan.recv_cumsum, plan.n_halo = cum, int(nh.value)
        flat = np.zeros(max(plan.n_halo, 1), np.int32)
        ptr = np.zeros(world + 1, np.int32)
        call("uspmv_halo_plan_need", h, _hp(flat), _hp(ptr))
        plan.need_lists = [flat[ptr[p]:ptr[p + 1]].copy() for p in range(world)]
        plan.send_lists = plan.send_ptr = None
        self.plan = plan
        plan.set_send(comm_schedule(plan.need_lists, rank, world, group))
        s0 = self.parts[first]
        self.n_rows, self.n_rows_padded, self.n_halo = s0.n_rows, s0.n_rows_padded, plan.n_halo
        self.n_elements = [self.parts[k].n_elements if self.parts[k] is not None else 0 for k in range(3)]
        self.n_chunks = s0.n_chunks
        xvt = capi.F32 if mode == "ap[sp_hp]" else capi.F64
        self.vec_length = s0.n_rows + max(s0.n_rows_padded - s0.n_rows, plan.n_halo)
        self.p2p = P2PHalo(plan, xvt, self.vec_length, rank, world, group)
        self.x = self.p2p.x
        self.x.fill_(1.0)
        self.y = torch.zeros(s0.n_rows_padded, dtype=self.x.dtype, device=f"cuda:{ctx.device}")
        self.comm_stream = torch.cuda.Stream(device=f"cuda:{ctx.device}")
        self._torch, self._eng = torch, eng
        dist.barrier(group=group)

    def validate(self, steps: int = 2):
        """Checked steps (COLLECTIVE; needs keep_coo=True): x = f(global row) with a NaN-poisoned halo, y against the product formed
        from the COO triplets with the values as partition_precisions stores them (dp / fp32 / fp16 by |v| against t1, t2).
        Returns this rank's max |y - ref| / sum|a||x| (inf on NaN or a raised error word); x is restored to 1.0."""
        import torch
        import torch.distributed as dist
        from . import validate as V
        if self._coo is None:
            raise ValueError("validate() needs keep_coo=True")
        dev = self.x.device
        nl = self.n_rows
        ref, scale = V.coo_ap_product(self._coo, self.mode, self.t1, self.t2, self.x.dtype, dev)
        g = torch.arange(self.row0, self.row0 + nl, device=dev, dtype=torch.float64)
        self.x[:nl] = V.x_of(g).to(self.x.dtype)
        self.x[nl:] = float("nan")
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)
        ep0 = self.p2p.status()[1]
        for _ in range(steps):
            self.y.fill_(float("nan"))
            self.step()
        torch.cuda.synchronize()
        o2n = V.device_int_tensor(self.parts[self.used[0]].device_arrays()["old_to_new"].value, nl, dev).long()
        worst = V.max_rel_err(self.y[o2n], ref, scale)
        err, ep = self.p2p.status()
        if err != 0 or ep != ep0 + steps:
            worst = float("inf")
        if dist.is_initialized():
            dist.barrier(group=self.group)
        self.x.fill_(1.0)
        self.y.zero_()
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)
        return worst

    def close(self):
        """COLLECTIVE teardown of the P2P arena (see P2PHalo.close)."""
        if self.p2p is not None:
            self.x = None
            self.p2p.close()
            self.p2p = None

    def algorithmic_bytes(self):
        vs = (8, 4, 2)
        xs = self.x.element_size()
        return (sum(self.n_elements[k] * (vs[k] + 4) + 8 * self.n_chunks for k in self.used) + xs * (self.n_rows + self.n_halo)
                + xs * self.n_rows_padded)

    def step(self):
        main = self._torch.cuda.current_stream()
        P = self.parts
        h = lambda q: q.h if q is not None else None
        call("uspmv_p2p_ap_spmv", self.p2p.h, self._eng.AP_MODE[self.mode], h(P[0]), h(P[1]), h(P[2]), vp(self.y.data_ptr()),
             vp(main.cuda_stream), vp(self.comm_stream.cuda_stream))

    def time_kernel(self, steps):
        torch, eng = self._torch, self._eng
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            eng.ap_spmv(self.mode, self.parts[0], self.parts[1], self.parts[2], self.x, self.y)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps


class BandedApSpmv:
    """One-GPU runner of the column-banded execution plan (uspmv_banded_*) for matrices whose x does not fit the L2 (BASELINE config 4
    at 2^25 rows): same interface as DistributedApSpmv (step / validate / time_kernel / close) so that bench.py measures both plans with
    one protocol.  mode None: plain dp."""

    kernel_name = "k_scs32_stream_ap (one launch per column band)"
    p2p = None

    def __init__(self, ctx, mtx, mode, t1, t2, C_, sigma, n_bands=0, algorithmic_bytes=None):
        import torch
        from . import engine as eng
        self.ctx, self.mode, self.t1, self.t2 = ctx, mode, float(t1), float(t2)
        self._coo = mtx
        self.nnz = mtx.nnz
        self.plan = eng.BandedPlan(mtx, C_, sigma, "dp", ap=mode, t1=t1, t2=t2, n_bands=n_bands)
        p = self.plan
        self.n_rows, self.n_rows_padded, self.n_halo, self.n_cols_local = p.n_rows, p.n_rows_padded, 0, p.n_cols
        self.n_elements, self.n_chunks = [p.n_elements], p.n_rows_padded // C_
        dt = torch.float32 if mode == "ap[sp_hp]" else torch.float64
        dev = f"cuda:{ctx.device}"
        self.x = torch.full((p.n_cols,), 1.0, dtype=dt, device=dev)
        self.y = torch.zeros(p.n_rows_padded, dtype=dt, device=dev)
        self._alg_bytes = algorithmic_bytes
        self._torch = torch

    def step(self):
        self.plan.spmv(self.x, self.y)

    def describe(self):
        return {"n_bands": self.plan.n_bands, "band_width": self.plan.band_width, "stored_elements_all_bands": int(self.plan.n_elements)}

    def algorithmic_bytes(self):
        if self._alg_bytes is not None:
            return self._alg_bytes
        raise ValueError("pass the un-banded structures' algorithmic bytes (the roofline denominator is the reference format's)")

    def validate(self, steps: int = 2):
        import torch
        from . import validate as V
        dev = self.x.device
        ref, scale = V.coo_ap_product(self._coo, self.mode, self.t1, self.t2, self.x.dtype, dev)
        self.x.copy_(V.x_of(torch.arange(self.plan.n_cols, device=dev, dtype=torch.float64)).to(self.x.dtype))
        for _ in range(steps):
            self.y.fill_(float("nan"))
            self.step()
        torch.cuda.synchronize()
        o2n = torch.from_numpy(self.plan.old_to_new.astype(np.int64)).to(dev)
        worst = V.max_rel_err(self.y[o2n], ref, scale)
        self.x.fill_(1.0)
        self.y.zero_()
        torch.cuda.synchronize()
        return worst

    def time_kernel(self, steps):
        torch = self._torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def close(self):
        self.plan = None
