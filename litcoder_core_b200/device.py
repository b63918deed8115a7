"""Device layer: thin, typed wrappers around the C ABI (liblitridge.so).

PyTorch is used for plumbing only -- it owns the device allocations (caching allocator), the
current CUDA stream and the timing events.  Every arithmetic operation on the hot path is one
of our own sm_100a kernels (or cuSOLVER syevd, the one library call, timed separately).  No torch
compute op is used here, and there is no CPU fallback: constructing `DeviceOps` without a CUDA
device or without the built library raises.

The nested-CV engine (engine.py) talks only to this interface, which is what lets the host logic
be tested on CPU against a NumPy stand-in that lives under tests/ (never shipped).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import queue
import threading
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .engine import SolverAccuracyError

_vp = C.c_void_p


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Mat:
    """Row-major fp32 device matrix [rows][cols] with pitch `ld`; `lo` is the optional second
    plane of a 3xTF32 split pair (then `hi` holds rna_tf32(x) and hi + lo == x up to 2^-22)."""

    __slots__ = ("hi", "lo", "rows", "cols", "_ld")

    def __init__(self, hi, lo, rows: int, cols: int, ld: Optional[int] = None):
        self.hi, self.lo, self.rows, self.cols, self._ld = hi, lo, rows, cols, ld

    @property
    def ld(self) -> int:
        return self._ld if self._ld is not None else self.hi.shape[1]

    @property
    def is_split(self) -> bool:
        return self.lo is not None

    def __repr__(self):
        return f"Mat({self.rows}x{self.cols}, ld={self.ld}, split={self.is_split})"


class MatF16:
    """fp16 split pair of a matrix (lit_split_f16): hi = fp16(s x), lo = fp16(s x - hi) with one power-of-two
    scale per group of `rows_per_group` rows; `inv_scale[g]` = 1/s.  Pitch `ld` in fp16 elements."""

    __slots__ = ("hi", "lo", "rows", "cols", "ld", "rows_per_group", "inv_scale")

    def __init__(self, hi, lo, rows: int, cols: int, ld: int, rows_per_group: int, inv_scale):
        self.hi, self.lo, self.rows, self.cols, self.ld = hi, lo, rows, cols, ld
        self.rows_per_group, self.inv_scale = rows_per_group, inv_scale


class SeriesStack:
    """Alpha stack of a GEMM-only fold in its compact form: `n_cheb` ordinary groups of `rows_pad` rows (the
    Chebyshev solutions, group i -> alpha slot_cheb[i]) followed by `n_tiles` series tiles (lit_series_stack) that
    serve the alphas slot_series with the 4-term combinations `coef` (n_series x 4, float64)."""

    __slots__ = ("mat", "n_cheb", "rows_pad", "n_tiles", "slot_cheb", "slot_series", "coef")

    def __init__(self, mat, n_cheb, rows_pad, n_tiles, slot_cheb, slot_series, coef):
        self.mat, self.n_cheb, self.rows_pad, self.n_tiles = mat, n_cheb, rows_pad, n_tiles
        self.slot_cheb, self.slot_series, self.coef = slot_cheb, slot_series, coef


class Partials:
    """Per-tile partial sums written by the fused correlation epilogue.  inv_row / inv_tile: the operand
    scales to undo when the GEMM ran on fp16 split pairs (None for the 3xTF32 form).  series / stack: the
    14-sum partials of the series tiles and the SeriesStack they belong to (compact stacks only)."""

    __slots__ = ("dot", "ssq", "n_tiles", "ld", "inv_row", "inv_tile", "series", "stack")

    def __init__(self, dot, ssq, n_tiles: int, ld: int, inv_row=None, inv_tile=None, series=None, stack=None):
        self.dot, self.ssq, self.n_tiles, self.ld = dot, ssq, n_tiles, ld
        self.inv_row, self.inv_tile, self.series, self.stack = inv_row, inv_tile, series, stack


class _EigTicket:
    __slots__ = ("issued", "done", "error")

    def __init__(self):
        self.issued = threading.Event()  # set once the worker has issued the call and recorded `done`
        self.done = None  # CUDA event on the side stream
        self.error = None


class _EigWorker(threading.Thread):
    """Issues (host-blocking) cuSOLVER calls on its own side stream.  A decomposition of a few thousand
    rows is latency-bound (thousands of small kernels), so several of them in flight on separate streams
    overlap with each other and with the tensor-core GEMMs."""

    def __init__(self, ops: "DeviceOps", jobs: "queue.Queue", stream):
        super().__init__(name="litridge-eig", daemon=True)
        self.ops, self.jobs, self.stream = ops, jobs, stream

    def run(self):
        ops, t = self.ops, self.ops.torch
        t.cuda.set_device(ops.device)
        while True:
            job = self.jobs.get()
            if job is None:
                return
            G, lam, ready, ticket = job
            try:
                with t.cuda.stream(self.stream):
                    self.stream.wait_event(ready)
                    ops.syevd(G, lam=lam)
                    ticket.done = t.cuda.Event()
                    ticket.done.record()
            except BaseException as e:  # surfaced by wait()
                ticket.error = e
            with ops._eig_lock:
                ops.eig_pending -= 1
            ticket.issued.set()


class DeviceOps:
    TILE_N = 256  # accumulator columns per tile of the fused-correlation GEMM (row padding of the stacked design)
    PART_N = 128  # columns reduced into one partial sum by the fused epilogue (half a tile)

    def __init__(self, device_index: Optional[int] = None, gemm_variant: int = _lib.GEMM_AUTO):
        import torch

        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("litcoder_core_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if device_index is None:
            device_index = torch.cuda.current_device()
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        cc = torch.cuda.get_device_capability(self.device)
        if cc[0] != 10:
            raise RuntimeError(f"litcoder_core_b200 kernels are built for sm_100a only; device has sm_{cc[0]}{cc[1]}")
        self.gemm_variant = gemm_variant
        self.reset_counters()
        self._eig_ws: Dict[Tuple[int, int], tuple] = {}
        self._bh_ws = None
        # concurrent eigendecompositions: more than one in flight was measured to be SLOWER on one GPU
        # (3.3 s per config-2 fit with 1 worker, 3.7 s with 2 or 3: they contend for the SMs the GEMMs leave)
        self.eig_workers = int(os.environ.get("LIT_EIG_WORKERS", "1"))
        self._eig_workers: List[object] = []
        self._copy_stream = None
        self._eig_jobs = None
        self._eig_lock = threading.Lock()
        self.eig_pending = 0  # decompositions queued or running
        self.overlap_sms = int(os.environ.get("LIT_GEMM_OVERLAP_SMS", "100"))  # GEMM grid while eigs are in flight
        self._overlap_always = os.environ.get("LIT_GEMM_OVERLAP_ALWAYS", "0") == "1"  # development knob
        self._cur_sm_limit = 0
        self.set_gemm_sm_limit(0)
        self._eig_infos: List[object] = []

    # ------------------------------------------------------------------ bookkeeping
    def reset_counters(self):
        self.launches = 0  # kernels of ours launched (cuSOLVER internals not counted)
        self.gemm_flops = 0.0  # algorithmic 2*M*N*K of the executed GEMMs (x3 tensor-core MMAs each)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.store_gemms = 0  # launches of the store-epilogue GEMM (both operand formats)
        self.store_gemms_f16 = 0  # ... of which on fp16 split pairs (lit_gemm_f16x3_nt)
        self.gemm_flops_f16 = 0.0  # algorithmic flops of those
        self.compact_stacks = 0  # fused-GEMM launches on a compact (series) alpha stack
        self._corr_log: List[tuple] = []  # (start, stop, flops) of every fused prediction+correlation GEMM
        self._staging: List[object] = []
        self._timed: Dict[str, List[tuple]] = {}
        self._solver_checks: List[tuple] = []  # (probe residuals, pivot flags) of the direct inner solves
        self.last_solver_residual = 0.0

    @property
    def stream(self) -> int:
        return self.torch.cuda.current_stream(self.device).cuda_stream

    @contextlib.contextmanager
    def timed(self, name: str):
        """Bracket a region with CUDA events on the current stream (resolved lazily by timings())."""
        t = self.torch
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        try:
            yield
        finally:
            e1.record()
            self._timed.setdefault(name, []).append((e0, e1))

    def timings(self) -> Dict[str, float]:
        """Milliseconds per category (synchronises the device)."""
        self.torch.cuda.synchronize(self.device)
        return {k: float(sum(a.elapsed_time(b) for a, b in v)) for k, v in self._timed.items()}

    def corr_launches(self):
        """(milliseconds, algorithmic flops) of every fused prediction+correlation GEMM launch (synchronises)."""
        self.torch.cuda.synchronize(self.device)
        return [(float(a.elapsed_time(b)), fl) for a, b, fl in self._corr_log]

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ memory
    def _empty2d(self, rows: int, ld: int):
        return self.torch.empty((max(rows, 1), ld), dtype=self.torch.float32, device=self.device)

    def empty(self, rows: int, cols: int, split: bool = False, ld: Optional[int] = None) -> Mat:
        ld = round_up(max(cols, 1), 32) if ld is None else ld
        hi = self._empty2d(rows, ld)
        lo = self._empty2d(rows, ld) if split else None
        return Mat(hi, lo, rows, cols)

    def zeros(self, rows: int, cols: int) -> Mat:
        m = self.empty(rows, cols)
        check(self.lib.lit_fill_f32(_vp(m.hi.data_ptr()), m.hi.numel(), 0.0, _vp(self.stream)), "fill")
        self.launches += 1
        return m

    def vec(self, n: int, dtype: str = "f32"):
        t = self.torch
        dt = {"f32": t.float32, "f64": t.float64, "i32": t.int32, "u8": t.uint8}[dtype]
        return t.empty((max(n, 1),), dtype=dt, device=self.device)

    def upload_vector(self, arr, dtype: str):
        t = self.torch
        npdt = {"f32": np.float32, "f64": np.float64, "i32": np.int32}[dtype]
        a = np.ascontiguousarray(np.asarray(arr, dtype=npdt))
        out = self.vec(a.size, dtype)
        if a.size:
            # through a page-locked staging block and an ASYNCHRONOUS copy: a pageable source would make the copy
            # drain the stream (the host would stop running ahead of the device at every small upload).  The
            # caching host allocator keeps the block alive until the copy has executed.
            host = t.empty((a.size,), dtype=out.dtype, pin_memory=True)
            host.numpy()[...] = a.reshape(-1)
            out[: a.size].copy_(host, non_blocking=True)
            self.h2d_bytes += a.nbytes
        return out

    def stage_indices(self, arrays) -> list:
        """Row-index vectors (int32) for a whole fit: one pinned staging buffer, ONE asynchronous H2D copy on
        the current stream; returns a device view per input array (each padded to a 16-byte boundary)."""
        t = self.torch
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (len(a) + 3) // 4 * 4
        host = t.empty((max(total, 4),), dtype=t.int32, pin_memory=True)
        hv = host.numpy()
        for a, o in zip(arrays, offs):
            a = np.asarray(a, dtype=np.int64)
            if len(a) and (a.min() < 0 or a.max() > 2 ** 31 - 1):
                raise ValueError("row index out of int32 range")
            hv[o:o + len(a)] = a
        dev = t.empty((max(total, 4),), dtype=t.int32, device=self.device)
        dev.copy_(host, non_blocking=True)
        self._staging.append(host)  # keep the pinned buffer alive until the counters are reset
        self.h2d_bytes += total * 4
        return [dev[o:o + max(len(a), 1)] for a, o in zip(arrays, offs)]

    def upload_index(self, idx) -> "object":
        return self.upload_vector(np.asarray(idx, dtype=np.int64), "i32")

    def upload_matrix(self, host: np.ndarray, col_start: int = 0, col_stop: Optional[int] = None,
                      chunk_bytes: int = 1 << 29) -> Mat:
        """H2D of host[:, col_start:col_stop] into an fp32 device matrix (fp64 input is converted on
        the device, mirroring torch.tensor(..., dtype=float32) at nested_cv.py:99-100).  The column
        block is copied with a pitched memcpy straight from the caller's array -- no host repack."""
        host = np.asarray(host)
        if host.ndim != 2:
            raise ValueError("expected a 2-D array")
        if host.dtype not in (np.float32, np.float64):
            host = host.astype(np.float32)
        if not host.flags.c_contiguous:
            host = np.ascontiguousarray(host)
        n, c_all = host.shape
        col_stop = c_all if col_stop is None else col_stop
        cols = col_stop - col_start
        out = self.empty(n, cols)
        if n == 0 or cols == 0:
            return out
        item = host.dtype.itemsize
        self.h2d_bytes += n * cols * item
        src0 = host.ctypes.data + col_start * item
        s = _vp(self.stream)
        if host.dtype == np.float32:
            check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr()), out.ld * 4, _vp(src0), c_all * 4, cols * 4, n, 1, s),
                  "memcpy_2d(H2D)")
            return out
        # float64: stage row chunks through a device buffer and convert
        rows_per = max(1, min(n, chunk_bytes // max(cols * 8, 1)))
        stage = self.torch.empty((rows_per * cols,), dtype=self.torch.float64, device=self.device)
        tmp32 = self.torch.empty((rows_per * cols,), dtype=self.torch.float32, device=self.device)
        for r0 in range(0, n, rows_per):
            nr = min(rows_per, n - r0)
            check(self.lib.lit_memcpy_2d(_vp(stage.data_ptr()), cols * 8, _vp(src0 + r0 * c_all * 8), c_all * 8,
                                         cols * 8, nr, 1, s), "memcpy_2d(H2D)")
            check(self.lib.lit_convert_f64_to_f32(_vp(stage.data_ptr()), _vp(tmp32.data_ptr()), nr * cols, s), "convert")
            check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr() + r0 * out.ld * 4), out.ld * 4, _vp(tmp32.data_ptr()),
                                         cols * 4, cols * 4, nr, 3, s), "memcpy_2d(D2D)")
            self.launches += 1
        return out

    BG_UPLOAD_MIN_BYTES = 256 << 20  # pageable float32 blocks at least this large are uploaded by a helper thread
    BG_CHUNK_BYTES = 64 << 20
    # staging threads per process: half the host cores, shared between the ranks of this node (torchrun's
    # LOCAL_WORLD_SIZE); 4 threads staged 25 GB/s and 8 threads 38 GB/s on the 16-core bench host
    BG_COPY_THREADS = max(2, min(8, (os.cpu_count() or 4) // (2 * max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)))))

    def upload_matrix_bg(self, host, col_start: int = 0, col_stop: Optional[int] = None):
        """upload_matrix for the responses: returns (Mat, ticket) at once; the copy runs on the copy stream behind
        the work already queued on the current stream and must be awaited with wait_copy(ticket).
        A helper thread feeds the copy engine 64 MB at a time, never more than two chunks ahead:
        * a large PAGEABLE float32 array (what a drop-in caller passes) would make cudaMemcpy2DAsync block the
          calling thread at ~10 GB/s, so it is staged through three page-locked buffers filled by BG_COPY_THREADS threads;
        * a PAGE-LOCKED array is copied in place, but still chunk by chunk: one 3.6 GB cudaMemcpy2DAsync occupies
          the host-to-device copy engine for 65 ms, and the first small upload of the design side (fold indices,
          alphas) queued behind it -- the whole design phase then ran after the responses instead of beside them.
        Either way the upload overlaps the design side of the fit."""
        t = self.torch
        arr = host if isinstance(host, np.ndarray) else None
        big = (arr is not None and arr.ndim == 2 and arr.dtype == np.float32 and arr.flags.c_contiguous
               and arr.shape[0] * ((arr.shape[1] if col_stop is None else col_stop) - col_start) * 4 >= self.BG_UPLOAD_MIN_BYTES)
        kind = self.lib.lit_host_pointer_kind(_vp(arr.ctypes.data)) if big else 2
        if kind not in (0, 1):
            with self.copy_stream() as ticket:
                out = self.upload_matrix(np.asarray(host), col_start, col_stop)
            return out, ticket
        n, c_all = arr.shape
        col_stop = c_all if col_stop is None else col_stop
        cols = col_stop - col_start
        out = self.empty(n, cols)
        if self._copy_stream is None:
            self._copy_stream = t.cuda.Stream(device=self.device)
        ready = t.cuda.Event()
        ready.record()
        self._copy_stream.wait_event(ready)
        ticket = _EigTicket()
        self.h2d_bytes += n * cols * 4
        if n == 0 or cols == 0:
            ticket.issued.set()
            return out, ticket
        threading.Thread(target=self._bg_upload, name="litridge-h2d", daemon=True,
                         args=(arr, col_start, cols, out, ticket, kind == 1)).start()
        return out, ticket

    def _bg_upload(self, arr, col_start: int, cols: int, out: Mat, ticket, page_locked: bool = False) -> None:
        from concurrent.futures import ThreadPoolExecutor

        t = self.torch
        try:
            t.cuda.set_device(self.device)
            if not page_locked and getattr(self, "_bg_bufs", None) is None:
                self._bg_bufs = [t.empty((self.BG_CHUNK_BYTES // 4,), dtype=t.float32, pin_memory=True) for _ in range(3)]
                self._bg_pool = ThreadPoolExecutor(self.BG_COPY_THREADS, thread_name_prefix="litridge-stage")
            stream = self._copy_stream
            n = arr.shape[0]
            rows_per = max(1, (self.BG_CHUNK_BYTES // 4) // cols)
            free = [None] * 3
            for k, r0 in enumerate(range(0, n, rows_per)):
                b = k % len(free)
                if free[b] is not None:
                    free[b].synchronize()  # chunk k - 3 is through: its staging buffer is free, the queue stays short
                nr = min(rows_per, n - r0)
                if page_locked:
                    src, pitch = arr.ctypes.data + (r0 * arr.shape[1] + col_start) * 4, arr.shape[1] * 4
                else:
                    view = self._bg_bufs[b].numpy()[: nr * cols].reshape(nr, cols)
                    step = -(-nr // self.BG_COPY_THREADS)
                    futs = [self._bg_pool.submit(np.copyto, view[s0:min(s0 + step, nr)],
                                                 arr[r0 + s0:r0 + min(s0 + step, nr), col_start:col_start + cols])
                            for s0 in range(0, nr, step)]
                    for f in futs:
                        f.result()
                    src, pitch = view.ctypes.data, cols * 4
                check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr() + r0 * out.ld * 4), out.ld * 4, _vp(src), pitch,
                                             cols * 4, nr, 1, _vp(stream.cuda_stream)), "memcpy_2d(H2D)")
                free[b] = t.cuda.Event()
                free[b].record(stream)
            ticket.done = t.cuda.Event()
            ticket.done.record(stream)
        except BaseException as e:  # surfaced by wait_copy
            ticket.error = e
        ticket.issued.set()

    @contextlib.contextmanager
    def copy_stream(self):
        """Run the enclosed uploads on a dedicated copy stream (ordered after the work already queued on the
        current stream); yields a ticket whose `done` event the consumer stream must wait on (wait_copy)."""
        t = self.torch
        if self._copy_stream is None:
            self._copy_stream = t.cuda.Stream(device=self.device)
        ready = t.cuda.Event()
        ready.record()
        self._copy_stream.wait_event(ready)
        ticket = _EigTicket()
        with t.cuda.stream(self._copy_stream):
            yield ticket
            ticket.done = t.cuda.Event()
            ticket.done.record()
        ticket.issued.set()

    @contextlib.contextmanager
    def side_stream(self):
        """Run the enclosed work on a second compute stream (ordered after what is already queued on the current
        one); yields a ticket for wait_copy.  Used to overlap the two serial panel chains of a fit (inner solves,
        outer inverses): each is a sequence of short launches that leaves most of the GPU idle."""
        t = self.torch
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = t.cuda.Stream(device=self.device)
        ready = t.cuda.Event()
        ready.record()
        self._side_stream.wait_event(ready)
        ticket = _EigTicket()
        with t.cuda.stream(self._side_stream):
            yield ticket
            ticket.done = t.cuda.Event()
            ticket.done.record()
        ticket.issued.set()

    def adopt(self, tensors) -> None:
        """Tell the caching allocator that tensors allocated on another stream are now used on the current one."""
        cur = self.torch.cuda.current_stream(self.device)
        for x in tensors:
            if self.torch.is_tensor(x):
                x.record_stream(cur)

    def wait_copy(self, ticket) -> None:
        if ticket is None:
            return
        ticket.issued.wait()  # a background upload may still be issuing its copies
        if ticket.error is not None:
            raise ticket.error
        if ticket.done is not None:
            self.torch.cuda.current_stream(self.device).wait_event(ticket.done)

    def wrap(self, tensor) -> Mat:
        """Adopt a resident torch CUDA float32 matrix (no copy) -- inputs already in HBM."""
        t = self.torch
        if tensor.dtype != t.float32 or tensor.dim() != 2 or tensor.device != self.device:
            raise ValueError("wrap() expects a 2-D float32 CUDA tensor on this device")
        if tensor.stride(1) != 1:
            raise ValueError("wrap() expects a row-major tensor")
        rows, cols = tensor.shape
        ld = tensor.stride(0) if rows > 1 else max(cols, 1)
        if ld % 4 or tensor.data_ptr() % 16:
            # TMA and the float4 kernels need a 16-byte aligned base and pitch: re-pitch on the device
            out = self.empty(rows, cols)
            check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr()), out.ld * 4, _vp(tensor.data_ptr()), ld * 4, cols * 4,
                                         rows, 3, _vp(self.stream)), "memcpy_2d(D2D)")
            return out
        view = tensor.as_strided((rows, ld), (ld, 1)) if ld != cols else tensor
        return Mat(view, None, rows, cols)

    def wrap_view(self, tensor, c0: int, c1: int) -> Mat:
        """Columns [c0, c1) of a resident row-major float32 CUDA matrix as a Mat (no copy)."""
        if c0 % 4 or tensor.stride(0) % 4 or tensor.data_ptr() % 16:
            return self.wrap(tensor[:, c0:c1].contiguous())
        full = self.wrap(tensor)
        return Mat(tensor[:, c0:c1], None, full.rows, c1 - c0, ld=full.ld)

    def col_view(self, m: Mat, c0: int, c1: int) -> Mat:
        """Columns [c0, c1) of a matrix or split pair as a view (c0 on a multiple of 4 floats: TMA base alignment)."""
        if c0 % 4:
            raise ValueError("col_view: the first column must be a multiple of 4")
        return Mat(m.hi[:, c0:c1], m.lo[:, c0:c1] if m.lo is not None else None, m.rows, c1 - c0, ld=m.ld)

    def raw(self, x):
        """Underlying torch tensor of a Mat (first plane) or device vector, for in-place collectives."""
        return x.hi if isinstance(x, Mat) else x

    def planes(self, m: Mat) -> list:
        """The torch tensors behind a Mat (one, or two for a split pair), for in-place collectives."""
        return [m.hi] + ([m.lo] if m.lo is not None else [])

    def download(self, t) -> np.ndarray:
        self.d2h_bytes += t.numel() * t.element_size()
        return t.detach().cpu().numpy()

    def download_matrix_start(self, m: Mat):
        """download_matrix in two halves: the copy into page-locked memory is queued now (current stream); the
        returned callable waits for it and hands out the ndarray.  The caller keeps `m` alive until then and does
        host-only work in between (the metrics dictionary of a fit is ~10 ms of list building)."""
        out = self.torch.empty((m.rows, m.cols), dtype=self.torch.float32, pin_memory=True).numpy()
        done = None
        if m.rows and m.cols:
            check(self.lib.lit_memcpy_2d(_vp(out.ctypes.data), m.cols * 4, _vp(m.hi.data_ptr()), m.ld * 4, m.cols * 4,
                                         m.rows, 2, _vp(self.stream)), "memcpy_2d(D2H)")
            done = self.torch.cuda.Event()
            done.record()
            self.d2h_bytes += out.nbytes

        def finish(keep=m):
            if done is not None:
                done.synchronize()
            return out
        return finish

    def download_matrix(self, m: Mat) -> np.ndarray:
        """D2H of the logical [rows][cols] block (hi + lo for split pairs is NOT applied here)."""
        if m.rows * m.cols >= (1 << 22):
            # large results land in page-locked memory (PCIe rate instead of the pageable staging rate); the
            # returned ndarray owns that buffer
            out = self.torch.empty((m.rows, m.cols), dtype=self.torch.float32, pin_memory=True).numpy()
        else:
            out = np.empty((m.rows, m.cols), dtype=np.float32)
        if m.rows and m.cols:
            check(self.lib.lit_memcpy_2d(_vp(out.ctypes.data), m.cols * 4, _vp(m.hi.data_ptr()), m.ld * 4, m.cols * 4,
                                         m.rows, 2, _vp(self.stream)), "memcpy_2d(D2H)")
            self.torch.cuda.current_stream(self.device).synchronize()
            self.d2h_bytes += out.nbytes
        return out

    # ------------------------------------------------------------------ layout kernels
    def gather_rows_T_split(self, src: Mat, idx, n: int, split: bool = True) -> Mat:
        """(cols x n) transpose of the gathered rows; split=False: one fp32 plane (enough for an operand that is
        re-split into fp16 pairs by its GEMM: 4 instead of 8 bytes written, 4 instead of 8 read twice)."""
        out = self.empty(src.cols, n, split=split)
        check(self.lib.lit_gather_rows_transpose_split(_vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr()), n, src.cols,
                                                       _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr() if split else 0),
                                                       out.ld, _vp(self.stream)), "gather_rows_transpose_split")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ fp16 pairs written by the producer
    def col_reduce(self, src: Mat, idx, n: int, sumsq: bool = True, absmax: bool = False):
        """(sum of squares, |max|) per column over the gathered rows (idx None: rows 0..n-1); device f32 vectors
        (None for the one not asked for).  lit_gather_col_reduce."""
        ss = self.vec(src.cols) if sumsq else None
        am = self.vec(src.cols) if absmax else None
        check(self.lib.lit_gather_col_reduce(_vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr() if idx is not None else 0),
                                             n, src.cols, _vp(ss.data_ptr() if sumsq else 0),
                                             _vp(am.data_ptr() if absmax else 0), _vp(self.stream)), "gather_col_reduce")
        self.launches += 1
        return ss, am

    def row_absmax(self, src: Mat):
        out = self.vec(src.rows)
        check(self.lib.lit_row_absmax(_vp(src.hi.data_ptr()), src.ld, src.rows, src.cols, _vp(out.data_ptr()),
                                      _vp(self.stream)), "row_absmax")
        self.launches += 1
        return out

    def f16_bound_scales(self, n: int, absmax=None, row_sumsq=None, col_sumsq=None):
        """(scale, inv_scale) device vectors of n power-of-two scales that put the bound
        absmax + sqrt(row_sumsq * max(col_sumsq)) into [2^14, 2^15) (lit_f16_bound_scales)."""
        scale, inv = self.vec(n), self.vec(n)
        check(self.lib.lit_f16_bound_scales(
            _vp(absmax.data_ptr() if absmax is not None else 0), _vp(row_sumsq.data_ptr() if row_sumsq is not None else 0),
            _vp(col_sumsq.data_ptr() if col_sumsq is not None else 0), col_sumsq.numel() if col_sumsq is not None else 0,
            n, _vp(scale.data_ptr()), _vp(inv.data_ptr()), _vp(self.stream)), "f16_bound_scales")
        self.launches += 1
        return scale, inv

    def gather_rows_T_f16(self, src: Mat, idx, n: int, scales) -> MatF16:
        """(cols x n) transpose of the gathered rows, written directly as the fp16 pair of scales[0][c] * value
        (scales = f16_bound_scales of the column maxima of |src|): no fp32 intermediate, no lit_split_f16 pass."""
        t = self.torch
        ld = round_up(max(n, 1), 64)
        hi = t.empty((max(src.cols, 1), ld), dtype=t.float16, device=self.device)
        lo = t.empty((max(src.cols, 1), ld), dtype=t.float16, device=self.device)
        check(self.lib.lit_gather_rows_transpose_f16(_vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr()), n, src.cols,
                                                     _vp(scales[0].data_ptr()), _vp(hi.data_ptr()), _vp(lo.data_ptr()),
                                                     ld, _vp(self.stream)), "gather_rows_transpose_f16")
        self.launches += 1
        return MatF16(hi, lo, src.cols, n, ld, 1, scales[1])

    def gather_rows(self, src: Mat, idx, n: int, rows_out: Optional[int] = None, split: bool = False) -> Mat:
        rows_out = n if rows_out is None else rows_out
        out = self.empty(rows_out, src.cols, split=split)
        check(self.lib.lit_gather_rows_f32(_vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr() if idx is not None else 0),
                                           n, src.cols, _vp(out.hi.data_ptr()),
                                           _vp(out.lo.data_ptr() if split else 0), out.ld, rows_out, _vp(self.stream)),
              "gather_rows")
        self.launches += 1
        return out

    def transpose(self, src: Mat, split: bool = False) -> Mat:
        out = self.empty(src.cols, src.rows, split=split)
        check(self.lib.lit_transpose_f32(_vp(src.hi.data_ptr()), src.rows, src.cols, src.ld, _vp(out.hi.data_ptr()),
                                         _vp(out.lo.data_ptr() if split else 0), out.ld, _vp(self.stream)), "transpose")
        self.launches += 1
        return out

    def split(self, src: Mat) -> Mat:
        """Split pair of an fp32 matrix (new planes; src is left untouched)."""
        out = self.empty(src.rows, src.cols, split=True, ld=src.ld)
        check(self.lib.lit_split_tf32(_vp(src.hi.data_ptr()), src.rows, src.cols, src.ld, _vp(out.hi.data_ptr()),
                                      _vp(out.lo.data_ptr()), out.ld, _vp(self.stream)), "split_tf32")
        self.launches += 1
        return out

    def copy(self, src: Mat) -> Mat:
        """Device-to-device copy of an fp32 matrix (same pitch)."""
        out = self.empty(src.rows, src.cols, ld=src.ld)
        check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr()), out.ld * 4, _vp(src.hi.data_ptr()), src.ld * 4,
                                     src.cols * 4, src.rows, 3, _vp(self.stream)), "memcpy_2d(D2D)")
        return out

    def axpy(self, a: float, x: Mat, y: Mat) -> None:
        check(self.lib.lit_axpy_f32(a, _vp(x.hi.data_ptr()), _vp(x.lo.data_ptr() if x.is_split else 0), x.ld,
                                    _vp(y.hi.data_ptr()), y.ld, x.rows, x.cols, _vp(self.stream)), "axpy")
        self.launches += 1

    # ------------------------------------------------------------------ statistics
    def col_stats(self, src: Mat, idx, n: int, ddof: int):
        mean, std = self.vec(src.cols), self.vec(src.cols)
        scratch = self.vec(2 * src.cols, "f64")
        check(self.lib.lit_col_stats(_vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr() if idx is not None else 0), n,
                                     src.cols, ddof, _vp(mean.data_ptr()), _vp(std.data_ptr()),
                                     _vp(scratch.data_ptr()), _vp(self.stream)), "col_stats")
        self.launches += 2
        return mean, std

    def gather_normalize(self, src: Mat, idx, n: int, mean, std, mode: int, eps: float,
                         rows_out: Optional[int] = None, split: bool = False, out: Optional[Mat] = None) -> Mat:
        rows_out = n if rows_out is None else rows_out
        if out is None:
            out = self.empty(rows_out, src.cols, split=split)
        elif out.rows != rows_out or out.cols != src.cols or out.is_split != split:
            raise ValueError("gather_normalize: destination does not match")
        check(self.lib.lit_gather_normalize_rows(
            _vp(src.hi.data_ptr()), src.ld, _vp(idx.data_ptr() if idx is not None else 0), n, src.cols,
            _vp(mean.data_ptr()), _vp(std.data_ptr() if std is not None else 0), mode, eps, _vp(out.hi.data_ptr()),
            _vp(out.lo.data_ptr() if split else 0), out.ld, rows_out, _vp(self.stream)), "gather_normalize_rows")
        self.launches += 1
        return out

    # ------------------------------------------------------------------ GEMMs
    def set_gemm_sm_limit(self, n_sms: int) -> None:
        """Restrict the persistent GEMM grids to n_sms SMs (0 = all); see lit_gemm_set_sm_limit."""
        check(self.lib.lit_gemm_set_sm_limit(int(n_sms)), "gemm_set_sm_limit")
        self._cur_sm_limit = int(n_sms)

    def _apply_sm_limit(self) -> None:
        """While eigendecompositions are queued or running, leave SMs free for them: the GEMMs are
        persistent kernels that otherwise occupy every SM (all registers) until they finish, which would
        serialise the two.  Measured on config 2 (1 GPU): 4.00 s per fit with 148 SMs, 3.1 s with 100."""
        want = self.overlap_sms if ((self.eig_pending > 0 or self._overlap_always) and self.overlap_sms > 0) else 0
        if want != self._cur_sm_limit:
            self.set_gemm_sm_limit(want)

    def gemm(self, A: Mat, B: Mat, alpha: float = 1.0, Cin: Optional[Mat] = None, beta: float = 0.0,
             split_out: bool = False, out: Optional[Mat] = None, ld_out: Optional[int] = None,
             precision: str = "tf32x3", pair_out=None):
        """out[M,N] = alpha * A[M,K] @ B[N,K]^T + beta * Cin (A, B split pairs).
        precision "f16x3": the operands are re-split into scaled fp16 pairs (one power-of-two scale per row of A
        and of B, lit_split_f16; an operand that already is a MatF16 is taken as it is) and multiplied by
        lit_gemm_f16x3_nt, whose epilogue undoes the scales: the same product accuracy at twice the tensor-core
        rate; pays for the large voxel-side products.
        pair_out = (scale, inv_scale) of f16_bound_scales (f16x3 only): the result is written ONLY as the fp16 pair
        of scale[row] * value, the A operand of the next fp16-pair GEMM (returns a MatF16)."""
        if precision not in ("tf32x3", "f16x3"):
            raise ValueError(f"gemm: unknown precision {precision!r}")
        if precision == "tf32x3" and not (isinstance(A, Mat) and isinstance(B, Mat) and A.is_split and B.is_split):
            raise ValueError("GEMM operands must be 3xTF32 split pairs")
        if pair_out is not None and (precision != "f16x3" or split_out or out is not None):
            raise ValueError("gemm: pair_out needs precision='f16x3' and no other output option")
        if A.cols != B.cols:
            raise ValueError(f"GEMM K mismatch: {A} x {B}")
        M, N, K = A.rows, B.rows, A.cols
        t = self.torch
        if pair_out is not None:
            ldh = round_up(max(N, 1), 64)
            out = MatF16(t.empty((max(M, 1), ldh), dtype=t.float16, device=self.device),
                         t.empty((max(M, 1), ldh), dtype=t.float16, device=self.device), M, N, ldh, 1, pair_out[1])
        elif out is None:
            out = self.empty(M, N, split=split_out, ld=ld_out)
        if M == 0 or N == 0:
            return out  # an empty voxel shard: nothing to launch
        if precision == "f16x3":
            A = A if isinstance(A, MatF16) else self.split_f16(A, 1)
            B = B if isinstance(B, MatF16) else self.split_f16(B, 1)
            if A.rows_per_group != 1 or B.rows_per_group != 1:
                raise ValueError("gemm: fp16-pair operands need one scale per row")
        self._apply_sm_limit()
        with self.timed("gemm_f16" if isinstance(A, MatF16) else "gemm"):
            if pair_out is not None:
                variant = self.gemm_variant if self.gemm_variant in (_lib.GEMM_AUTO, _lib.GEMM_1CTA_N256,
                                                                     _lib.GEMM_2CTA_N256) else _lib.GEMM_AUTO
                check(self.lib.lit_gemm_f16x3_nt_pairout(
                    _vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(B.hi.data_ptr()), _vp(B.lo.data_ptr()), B.ld,
                    M, N, K, alpha, _vp(Cin.hi.data_ptr() if Cin is not None else 0), Cin.ld if Cin is not None else 0,
                    beta, _vp(0), 0, _vp(A.inv_scale.data_ptr()), _vp(B.inv_scale.data_ptr()),
                    _vp(pair_out[0].data_ptr()), _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr()), out.ld, variant,
                    _vp(self.stream)), "gemm_f16x3_nt_pairout")
            elif isinstance(A, MatF16):
                variant = self.gemm_variant if self.gemm_variant in (_lib.GEMM_AUTO, _lib.GEMM_1CTA_N256,
                                                                     _lib.GEMM_2CTA_N256) else _lib.GEMM_AUTO
                check(self.lib.lit_gemm_f16x3_nt(
                    _vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(B.hi.data_ptr()), _vp(B.lo.data_ptr()), B.ld,
                    M, N, K, alpha, _vp(Cin.hi.data_ptr() if Cin is not None else 0), Cin.ld if Cin is not None else 0,
                    beta, _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr() if out.is_split else 0), out.ld,
                    _vp(A.inv_scale.data_ptr()), _vp(B.inv_scale.data_ptr()), variant, _vp(self.stream)),
                    "gemm_f16x3_nt")
            else:
                self._gemm_call(A, B, M, N, K, alpha, Cin, beta, out)
        self.launches += 1
        self.store_gemms += 1
        self.gemm_flops += 2.0 * M * N * K
        if isinstance(A, MatF16):
            self.store_gemms_f16 += 1
            self.gemm_flops_f16 += 2.0 * M * N * K
        return out

    def _gemm_call(self, A, B, M, N, K, alpha, Cin, beta, out):
        check(self.lib.lit_gemm_tf32x3_nt(
            _vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(B.hi.data_ptr()), _vp(B.lo.data_ptr()), B.ld, M, N, K,
            alpha, _vp(Cin.hi.data_ptr() if Cin is not None else 0), Cin.ld if Cin is not None else 0, beta,
            _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr() if out.is_split else 0), out.ld, self.gemm_variant,
            _vp(self.stream)), "gemm_tf32x3_nt")

    def split_f16(self, src: Mat, rows_per_group: int = 1) -> MatF16:
        """fp16 split pair of `src` (fp32 or 3xTF32 split pair) with one scale per `rows_per_group` rows."""
        t = self.torch
        rows, cols = src.rows, src.cols
        ld = round_up(max(cols, 1), 64)
        n_groups = -(-max(rows, 1) // rows_per_group)
        hi = t.empty((max(rows, 1), ld), dtype=t.float16, device=self.device)
        lo = t.empty((max(rows, 1), ld), dtype=t.float16, device=self.device)
        inv = self.vec(n_groups)
        scratch = self.vec(2 * n_groups)
        with self.timed("split_f16"):
            check(self.lib.lit_split_f16(
                _vp(src.hi.data_ptr()), _vp(src.lo.data_ptr() if src.is_split else 0), src.ld, rows, cols, rows_per_group,
                _vp(hi.data_ptr()), _vp(lo.data_ptr()), ld, _vp(inv.data_ptr()), _vp(scratch.data_ptr()),
                _vp(self.stream)), "split_f16")
        self.launches += 3
        return MatF16(hi, lo, rows, cols, ld, rows_per_group, inv)

    def gemm_corr(self, A: Mat, B, n_groups: int, rows_per_group: int, Yz: Mat,
                  precision: str = "tf32x3") -> Partials:
        """Fused prediction + per-voxel reduction (lit_gemm_tf32x3_nt_corr / lit_gemm_f16x3_nt_corr /
        lit_gemm_corr_series).  B is the alpha stack: a split Mat of n_groups * rows_per_group rows, or a
        SeriesStack (compact form; n_groups is then the number of alphas it serves).
        precision "f16x3": the operands are first re-split into scaled fp16 pairs (one scale per voxel row of A,
        one per 256-row tile of B); corr_finalize undoes the scales."""
        if precision not in ("tf32x3", "f16x3"):
            raise ValueError(f"gemm_corr: unknown precision {precision!r}")
        stack = B if isinstance(B, SeriesStack) else None
        if stack is not None:
            if stack.rows_pad != rows_per_group or stack.n_cheb + len(stack.slot_series) != n_groups:
                raise ValueError("gemm_corr: SeriesStack does not match the alpha count / row padding")
            B, n_plain, n_st = stack.mat, stack.n_cheb, stack.n_tiles
        else:
            n_plain, n_st = n_groups, 0
        if rows_per_group % self.TILE_N or B.rows != n_plain * rows_per_group + n_st * self.TILE_N \
                or Yz.rows != rows_per_group:
            raise ValueError("gemm_corr: stacked design / response rows must be padded to the tile size")
        if Yz.cols != A.rows or A.cols != B.cols:
            raise ValueError("gemm_corr: shape mismatch")
        M, K = A.rows, A.cols
        n_tiles = n_plain * rows_per_group // self.PART_N
        ld = round_up(M, 32)
        t = self.torch
        dot = t.empty((max(n_tiles, 1), ld), dtype=t.float32, device=self.device)
        ssq = t.empty((max(n_tiles, 1), ld), dtype=t.float32, device=self.device)
        series = t.empty((2 * n_st * 14, ld), dtype=t.float32, device=self.device) if n_st else None
        variant = self.gemm_variant if self.gemm_variant in (_lib.GEMM_AUTO, _lib.GEMM_1CTA_N256, _lib.GEMM_2CTA_N256) \
            else _lib.GEMM_AUTO
        flops = 2.0 * M * B.rows * K
        inv_row = inv_tile = None
        if isinstance(A, MatF16) and (precision != "f16x3" or A.rows_per_group != 1):
            raise ValueError("gemm_corr: an fp16-pair A operand needs precision='f16x3' and one scale per row")
        if precision == "f16x3":
            A = A if isinstance(A, MatF16) else self.split_f16(A, 1)
            B = self.split_f16(B, self.TILE_N)
            inv_row, inv_tile = A.inv_scale, B.inv_scale
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        self._timed.setdefault("gemm_corr", []).append((e0, e1))
        self._corr_log.append((e0, e1, flops))
        self._apply_sm_limit()
        e0.record()
        check(self.lib.lit_gemm_corr_series(
            int(precision == "f16x3"), _vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(B.hi.data_ptr()),
            _vp(B.lo.data_ptr()), B.ld, M, n_plain, rows_per_group, n_st, K, _vp(Yz.hi.data_ptr()), Yz.ld,
            _vp(dot.data_ptr()), _vp(ssq.data_ptr()), _vp(series.data_ptr() if n_st else 0), ld, variant,
            _vp(self.stream)), "gemm_corr_series")
        e1.record()
        self.launches += 1
        self.compact_stacks += int(stack is not None)
        self.gemm_flops += flops
        return Partials(dot, ssq, n_tiles, ld, inv_row, inv_tile, series, stack)

    # ------------------------------------------------------------------ eigendecomposition
    def syevd(self, G: Mat, lam=None):
        """In place: rows of G become the eigenvectors; returns ascending eigenvalues (device f32)."""
        n = G.rows
        if G.cols != n:
            raise ValueError("syevd needs a square matrix")
        key = (n, threading.get_ident())
        if key not in self._eig_ws:
            dev_b, host_b = C.c_size_t(0), C.c_size_t(0)
            check(self.lib.lit_syevd_workspace(n, 0, 1, C.byref(dev_b), C.byref(host_b)), "syevd_workspace")
            work = self.torch.empty((max(dev_b.value, 16),), dtype=self.torch.uint8, device=self.device)
            work_h = (C.c_uint8 * max(host_b.value, 16))()
            for k in [k for k in self._eig_ws if k[1] == key[1]]:
                del self._eig_ws[k]  # keep only the latest size per calling thread
            self._eig_ws[key] = (work, dev_b.value, work_h, host_b.value)
        work, dev_b, work_h, host_b = self._eig_ws[key]
        info = self.vec(1, "i32")
        self._eig_infos.append(info)
        lam = self.vec(n) if lam is None else lam
        with self.timed("eig"):
            check(self.lib.lit_syevd(_vp(G.hi.data_ptr()), n, G.ld, 0, 1, _vp(lam.data_ptr()), _vp(work.data_ptr()),
                                     dev_b, C.cast(work_h, _vp), host_b, _vp(info.data_ptr()), _vp(self.stream)),
                  "syevd")
        return lam

    def check_eig(self) -> None:
        """Raise if any syevd since the last check reported failure (synchronises)."""
        infos, self._eig_infos = self._eig_infos, []
        self.torch.cuda.synchronize(self.device)
        bad = [int(i.item()) for i in infos if int(i.item()) != 0]
        if bad:
            raise _lib.LitRidgeError(f"cuSOLVER syevd did not converge (info = {bad})")

    def syevd_async(self, G: Mat):
        """Queue syevd(G) behind everything already queued on the current stream, to run on the side
        stream.  Returns (lam, ticket); G and lam may be read on the current stream after wait(ticket).

        cuSOLVER's syevd blocks its calling host thread for the whole decomposition (measured: host call
        42.7 ms of 45.0 ms device time at n = 3072), so the calls are issued by a worker thread; the
        caller keeps feeding the main stream meanwhile."""
        t = self.torch
        if not self._eig_workers:
            self._eig_jobs = queue.Queue()
            for _ in range(self.eig_workers):
                w = _EigWorker(self, self._eig_jobs, t.cuda.Stream(device=self.device))
                w.start()
                self._eig_workers.append(w)
        lam = self.vec(G.rows)
        ready = t.cuda.Event()
        ready.record()
        ticket = _EigTicket()
        with self._eig_lock:
            self.eig_pending += 1
        self._eig_jobs.put((G, lam, ready, ticket))
        return lam, ticket

    def wait(self, ticket) -> None:
        """Make the current stream wait for an eigendecomposition queued with syevd_async."""
        ticket.issued.wait()
        if ticket.error is not None:
            raise ticket.error
        self.torch.cuda.current_stream(self.device).wait_event(ticket.done)

    def close(self) -> None:
        for w in self._eig_workers:
            self._eig_jobs.put(None)
        for w in self._eig_workers:
            w.join(timeout=10)
        self._eig_workers = []

    # ------------------------------------------------------------------ eigendecomposition-free inner solver
    def lambda_max(self, G: Mat, steps: int = 96):
        """Device scalar (1-element f64 tensor) with lambda_max of the symmetric PSD matrix G (Lanczos + Sturm
        bisection on the device; lit_lanczos_lambda_max).  Reading it on the host synchronises."""
        n = G.rows
        steps = min(steps, n)
        vec_scratch = self.vec(3 * n)
        scal_scratch = self.vec(2 * steps + 4, "f64")
        out = self.vec(1, "f64")
        check(self.lib.lit_lanczos_lambda_max(_vp(G.hi.data_ptr()), G.ld, n, steps, _vp(vec_scratch.data_ptr()),
                                              _vp(scal_scratch.data_ptr()), _vp(0), _vp(out.data_ptr()),
                                              _vp(self.stream)), "lanczos_lambda_max")
        self.launches += 2 * steps + 2
        return out

    def lambda_max_batched(self, mats, steps: int = 96):
        """lambda_max of several symmetric matrices of ONE size and pitch in lock step (one launch per Lanczos step
        for all of them; lit_lanczos_lambda_max_batched).  Returns a device f64 vector with one value per matrix."""
        if not mats:
            return self.vec(0, "f64")
        n, ld = mats[0].rows, mats[0].ld
        if any(m.rows != n or m.cols != n or m.ld != ld for m in mats):
            raise ValueError("lambda_max_batched: the matrices must share one size and pitch")
        steps = min(steps, n)
        nb = len(mats)
        vec_scratch = self.vec(nb * 3 * n)
        scal_scratch = self.vec(nb * (2 * steps + 4), "f64")
        out = self.vec(nb, "f64")
        ptrs = (C.c_void_p * nb)(*[m.hi.data_ptr() for m in mats])
        check(self.lib.lit_lanczos_lambda_max_batched(C.cast(ptrs, _vp), nb, ld, n, steps, _vp(vec_scratch.data_ptr()),
                                                      _vp(scal_scratch.data_ptr()), _vp(out.data_ptr()),
                                                      _vp(self.stream)), "lanczos_lambda_max_batched")
        self.launches += (2 * steps + 2) * -(-nb // 32)
        return out

    @staticmethod
    def chebyshev_plan_interval(lo: float, hi: float, tol: float = 1e-6):
        """Scalar schedule of the Chebyshev iteration for a symmetric operator with spectrum in [lo, hi], lo > 0:
        list of (c1, c2) per step (d = c1 d + c2 r), after Saad, Iterative Methods, alg. 12.1."""
        if not (0.0 < lo <= hi):
            raise ValueError(f"chebyshev_plan_interval: need 0 < lo <= hi (got {lo}, {hi})")
        theta = 0.5 * (hi + lo)
        delta = max(0.5 * (hi - lo), 1e-9 * theta)
        sigma1 = theta / delta
        kappa = hi / lo
        rate = (np.sqrt(kappa) - 1.0) / (np.sqrt(kappa) + 1.0)
        n_steps = max(2, int(np.ceil(np.log(2.0 / tol) / -np.log(rate)))) if rate > 0 else 2
        rho = 1.0 / sigma1
        plan = [(0.0, 1.0 / theta)]
        for _ in range(1, n_steps):
            rho_new = 1.0 / (2.0 * sigma1 - rho)
            plan.append((rho_new * rho, 2.0 * rho_new / delta))
            rho = rho_new
        return plan

    @staticmethod
    def chebyshev_plan(lam_max: float, a2: float, tol: float = 1e-6, safety: float = 1.02):
        """Schedule for (G + a2 I) with spec(G) in [0, safety * lam_max]."""
        return DeviceOps.chebyshev_plan_interval(a2, a2 + safety * lam_max, tol)

    @staticmethod
    def solver_partition(lam_max: float, a2_list, series_ratio: float = 60.0):
        """Which alphas are solved by Chebyshev iteration and which by the Neumann series (a^2 >= ratio * lam_max)."""
        series = [j for j, a2 in enumerate(a2_list) if a2 >= series_ratio * lam_max]
        cheb = [j for j in range(len(a2_list)) if j not in series]
        return cheb, series

    def solver_block_rows(self, n_rows: int, lam_max: float, a2_list, series_ratio: float = 60.0) -> int:
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        return (len(cheb) + (3 if series else 0)) * n_rows

    def _view_rows(self, m: Mat, r0: int, rows: int) -> Mat:
        return Mat(m.hi[r0:r0 + rows], m.lo[r0:r0 + rows] if m.lo is not None else None, rows, m.cols, ld=m.ld)

    def lbo_prepare(self, Pv: Mat, Vt: Mat, lam_o, a2_min: float, steps: int = 48, lanczos: bool = True) -> dict:
        """Leave-block-out form of an inner fold whose validation rows R are exactly the rows removed from the
        outer training set (G_in = G_o - X_R^T X_R) once the outer Gram's eigendecomposition G_o = V diag(lam_o) V^T
        is known:
            X_R (G_in + a^2 I)^-1 = (I - H_a)^-1 X_R M_a,   M_a = V diag(1/(lam_o + a^2)) V^T,  H_a = X_R M_a X_R^T
        (Woodbury on the downdate), i.e. an |R| x |R| system with spectrum in (0, 1] instead of a p x p one with
        condition number (lam_max + a^2)/a^2.  Queues B = X_R V (|R| x k), E = B diag(1/(lam_o + a2_min)),
        H = E B^T for the smallest alpha and (unless the caller batches it: lanczos=False) the Lanczos estimate of
        lambda_max(H); nothing is read back here.
        Pv: split pair of X_R (|R| x p); Vt: split pair of the eigenvectors as rows (k x p)."""
        B = self.gemm(Pv, Vt, split_out=True)
        E = self._lbo_scaled(B, lam_o, a2_min)
        H = self.gemm(E, B, split_out=True)
        return {"B": B, "E": E, "H": H, "a2": float(a2_min), "hmax_dev": self.lambda_max(H, steps) if lanczos else None}

    def _lbo_scaled(self, B: Mat, lam_o, a2: float) -> Mat:
        """B diag(1 / (lam_o + a2)) as a split pair (columns of non-positive eigenvalues -- numerically null
        directions, where B vanishes -- are dropped)."""
        alpha_v = self.vec(B.rows)
        check(self.lib.lit_fill_f32(_vp(alpha_v.data_ptr()), alpha_v.numel(), float(np.sqrt(np.float64(a2))),
                                    _vp(self.stream)), "fill")
        self.launches += 1
        return self.scale_rows_by_alpha(B, lam_o, alpha_v, False, 0.0)

    @staticmethod
    def lbo_bounds(h0: float, a2_0: float, a2: float, lam_top: float):
        """Spectral interval [lo, 1] of I - H_a from the Lanczos estimate h0 of lambda_max(H_a0) at the smallest
        alpha.  A Ritz value never exceeds the eigenvalue it approaches, so 15 % of the gap 1 - h0 is given back as
        a margin; H_a <= H_a0 (lam_top + a2_0)/(lam_top + a2) for a2 >= a2_0 (the ratio of the two diagonal scalings
        is largest at the top eigenvalue), and I - H_a >= a2/(a2 + lam_top) always (G_R <= G_o)."""
        hi_h0 = 1.0 - 0.85 * min(max(1.0 - h0, 0.0), 1.0)
        hi_h = hi_h0 * (1.001 * lam_top + a2_0) / (1.001 * lam_top + a2)
        lo_rigorous = a2 / (a2 + 1.001 * lam_top)
        return max(1.0 - hi_h, lo_rigorous), 1.0

    def _lbo_solve(self, block: Mat, lbo: dict, cheb, a2_list, n_rows: int, tol: float = 2e-7) -> None:
        """Rows [i * n_rows, (i+1) * n_rows) of `block` <- J (I - H_a)^-1 B D_a V^T for the i-th Chebyshev alpha
        (J = column centring over the validation rows).  The |R| x |R| systems are solved by Chebyshev iteration on
        the transposed unknown x = Z^T (k x |R|), so that every step is one NT GEMM r = t + d H; a handful of steps
        each (the spectrum of I - H_a lies in [1 - lambda_max(H_a), 1])."""
        prep, Vs, lam_o = lbo["prep"], lbo["V"], lbo["lam"]
        B = prep["B"]
        k = B.cols
        s = _vp(self.stream)
        x, dvec, t, r = (self.empty(k, n_rows) for _ in range(4))
        dsp = self.empty(k, n_rows, split=True)
        ld = x.ld
        for i, j in enumerate(cheb):
            a2 = float(a2_list[j])
            if abs(a2 - prep["a2"]) <= 1e-12 * a2:
                E, H = prep["E"], prep["H"]
            else:
                E = self._lbo_scaled(B, lam_o, a2)
                H = self.gemm(E, B, split_out=True)
            Ef = self.zeros(n_rows, k)
            self.axpy(1.0, E, Ef)  # hi + lo
            Et = self.transpose(Ef)  # right-hand side, transposed: (k x |R|)
            if Et.ld != ld:
                raise ValueError("lbo_solve: pitch mismatch")
            lo, hi = self.lbo_bounds(lbo["h0"], prep["a2"], a2, lbo["lam_top"])
            plan = self.chebyshev_plan_interval(lo, hi, tol)
            for step, (c1, c2) in enumerate(plan):
                src = Et if step == 0 else r
                check(self.lib.lit_cheb_update(_vp(dvec.hi.data_ptr()), _vp(src.hi.data_ptr()), _vp(x.hi.data_ptr()),
                                               _vp(t.hi.data_ptr()), _vp(dsp.hi.data_ptr()), _vp(dsp.lo.data_ptr()), ld,
                                               k, n_rows, c1, c2, 1.0, int(step == 0), s), "cheb_update")
                self.launches += 1
                if step + 1 < len(plan):
                    self.gemm(dsp, H, alpha=1.0, Cin=t, beta=1.0, out=r)  # r = t + d H  (operator I - H, t = r - d)
            Z = self.transpose(x, split=True)  # (|R| x k)
            S = self.gemm(Z, Vs)  # back to the feature basis: Z V^T  (|R| x p)
            mean, _ = self.col_stats(S, None, n_rows, ddof=0)
            self.gather_normalize(S, None, n_rows, mean, None, 2, 0.0, out=self._view_rows(block, i * n_rows, n_rows))

    def solve_blocks(self, Gs: Mat, Pc: Mat, n_rows: int, lam_max: float, a2_list, series_ratio: float = 60.0,
                     lbo: Optional[dict] = None) -> Mat:
        """The expensive, alpha-specific part of P_c (G + a^2 I)^-1 as ONE compact fp32 matrix
        [(n_cheb + 3) * n_rows][p]: the solutions of the small alphas followed by P_c G^q, q = 1..3
        (the shared powers of the Neumann series).  This is what a rank broadcasts for a fold it owns
        (130 MB at config 2 instead of the 755 MB alpha stack).  Gs: split pair of G (p x p); Pc: fp32 (n_rows x p).
        The small alphas are solved by Chebyshev iteration on (G + a^2 I), or -- with `lbo` (lbo_prepare's dict plus
        "V", "lam", "lam_top", "h0") -- through the leave-block-out identity on the outer eigendecomposition."""
        p = Gs.rows
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        n_q = 3 if series else 0
        block = self.zeros((len(cheb) + n_q) * n_rows, p)
        s = _vp(self.stream)
        ld = block.ld
        if cheb and lbo is not None:
            self._lbo_solve(block, lbo, cheb, a2_list, n_rows)
        elif cheb:
            # The systems of all Chebyshev alphas advance together: one update launch per still-active system
            # (own scalars) and ONE stacked GEMM r = t - d G per step over the rows of the active systems.
            # cheb is in alpha order = descending step count, so the active systems are always a row prefix.
            plans = [self.chebyshev_plan(lam_max, float(a2_list[j])) for j in cheb]
            order = sorted(range(len(cheb)), key=lambda i: -len(plans[i]))  # work-buffer slot -> system
            nc = len(cheb)
            d, t, r = (self.empty(nc * n_rows, p, ld=ld) for _ in range(3))
            dsp = self.empty(nc * n_rows, p, split=True, ld=ld)
            if Pc.ld != ld:
                raise ValueError("solve_blocks: Pc must share the block's pitch")
            blk = n_rows * ld * 4  # bytes per system in every buffer
            for k in range(len(plans[order[0]])):
                n_active = sum(1 for i in order if k < len(plans[i]))  # a prefix of the slots
                for slot in range(n_active):
                    i = order[slot]
                    c1, c2 = plans[i][k]
                    src = Pc.hi.data_ptr() if k == 0 else r.hi.data_ptr() + slot * blk
                    check(self.lib.lit_cheb_update(_vp(d.hi.data_ptr() + slot * blk), _vp(src),
                                                   _vp(block.hi.data_ptr() + i * blk), _vp(t.hi.data_ptr() + slot * blk),
                                                   _vp(dsp.hi.data_ptr() + slot * blk), _vp(dsp.lo.data_ptr() + slot * blk),
                                                   ld, n_rows, p, c1, c2, float(a2_list[cheb[i]]), int(k == 0), s),
                          "cheb_update")
                    self.launches += 1
                # systems whose plan ends at this step do not need their residual any more
                n_need = sum(1 for i in order if k + 1 < len(plans[i]))
                if n_need:
                    rows = n_need * n_rows
                    self.gemm(self._view_rows(dsp, 0, rows), Gs, alpha=-1.0, Cin=self._view_rows(t, 0, rows), beta=1.0,
                              out=self._view_rows(r, 0, rows))  # r = t - d G
        if series:
            Q = self.split(Pc)
            for q in range(3):
                Q = self.gemm(Q, Gs, split_out=True, ld_out=ld)
                self.axpy(1.0, Q, self._view_rows(block, (len(cheb) + q) * n_rows, n_rows))
        return block

    # ------------------------------------------------------------------ batched direct inner solver
    SOLVER_TOL = 2e-4  # a-posteriori probe residual above which a direct solve is rejected (check_solver)
    SPD_MAX_BATCH = 128

    def solve_blocks_many(self, jobs, series_ratio: float = 60.0) -> list:
        """The compact solution blocks of SEVERAL GEMM-only folds at once (layout of solve_blocks).  jobs: dicts with
        G (fp32 Mat p x p), Pc (fp32 Mat n_rows x p), n_rows, lam_max, a2 (list).  The small-alpha systems of all
        jobs -- up to 128 per launch -- go through the batched blocked Cholesky solver (lit_spd_solve_batched: every
        panel / trailing update is ONE batched tcgen05 GEMM for all systems), followed by one batched product per
        job and a probe-vector residual check that is read back by check_solver()."""
        t = self.torch
        blocks, parts = [], []
        for job in jobs:
            cheb, series = self.solver_partition(job["lam_max"], job["a2"], series_ratio)
            n_rows, p = job["n_rows"], job["G"].rows
            block = self.zeros((len(cheb) + (3 if series else 0)) * n_rows, p)
            if job["Pc"].ld != block.ld or job["G"].ld != block.ld:
                raise ValueError("solve_blocks_many: G, Pc and the block must share one pitch")
            blocks.append(block)
            parts.append((cheb, series))
        s = _vp(self.stream)
        # ---- small alphas: chunks of whole jobs, at most SPD_MAX_BATCH systems each
        i0 = 0
        while i0 < len(jobs):
            i1, nsys = i0, 0
            while i1 < len(jobs) and (i1 == i0 or nsys + len(parts[i1][0]) <= self.SPD_MAX_BATCH) \
                    and jobs[i1]["G"].rows == jobs[i0]["G"].rows:
                nsys += len(parts[i1][0])
                i1 += 1
            if nsys > self.SPD_MAX_BATCH:
                raise ValueError("solve_blocks_many: more than 128 solved alphas in one fold")
            if nsys:
                self._spd_chunk(jobs[i0:i1], blocks[i0:i1], [c for c, _ in parts[i0:i1]], nsys, s)
            i0 = i1
        # ---- shared powers of the Neumann series
        for job, block, (cheb, series) in zip(jobs, blocks, parts):
            if series:
                Gs = self.split(job["G"])
                Q = self.split(job["Pc"])
                for q in range(3):
                    Q = self.gemm(Q, Gs, split_out=True, ld_out=block.ld)
                    self.axpy(1.0, Q, self._view_rows(block, (len(cheb) + q) * job["n_rows"], job["n_rows"]))
        return blocks

    def _spd_chunk(self, jobs, blocks, chebs, nsys: int, s) -> None:
        t = self.torch
        p = jobs[0]["G"].rows
        mp = max(j["n_rows"] for j in jobs)
        fF, fS, fD, ldw, rows = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_long(0), C.c_long(0)
        check(self.lib.lit_spd_solve_workspace(nsys, p, mp, C.byref(fF), C.byref(fS), C.byref(fD), C.byref(ldw),
                                               C.byref(rows)), "spd_solve_workspace")
        ldw, rows = ldw.value, rows.value
        f32 = dict(dtype=t.float32, device=self.device)
        F = t.empty((fF.value,), **f32)
        S_hi, S_lo = t.empty((fS.value,), **f32), t.empty((fS.value,), **f32)
        Dg_hi, Dg_lo = t.empty((fD.value,), **f32), t.empty((fD.value,), **f32)
        info = t.empty((nsys,), dtype=t.int32, device=self.device)
        Gp, Rp, mh, a2h = (C.c_void_p * nsys)(), (C.c_void_p * nsys)(), (C.c_int * nsys)(), (C.c_float * nsys)()
        k = 0
        for job, cheb in zip(jobs, chebs):
            for j in cheb:
                Gp[k], Rp[k], mh[k], a2h[k] = job["G"].hi.data_ptr(), job["Pc"].hi.data_ptr(), job["n_rows"], job["a2"][j]
                k += 1
        ldg, ldr = jobs[0]["G"].ld, jobs[0]["Pc"].ld
        with self.timed("spd_solve"):
            check(self.lib.lit_spd_solve_batched(nsys, p, mp, C.cast(Gp, _vp), ldg, C.cast(Rp, _vp), ldr,
                                                 C.cast(mh, _vp), C.cast(a2h, _vp), _vp(F.data_ptr()),
                                                 _vp(S_hi.data_ptr()), _vp(S_lo.data_ptr()), _vp(Dg_hi.data_ptr()),
                                                 _vp(Dg_lo.data_ptr()), _vp(info.data_ptr()), s), "spd_solve_batched")
            steps = ldw // 128
            self.launches += 1 + 4 * steps
            self.gemm_flops += sum(2.0 * nsys * (rows - 128 * (j + 1)) * 128 * (128 + max(ldw - 128 * (j + 1), 0))
                                   for j in range(steps))
            sys_stride = rows * ldw
            k = 0
            for job, block, cheb in zip(jobs, blocks, chebs):
                nc, n_rows = len(cheb), job["n_rows"]
                if not nc:
                    continue
                y0 = (k * sys_stride + ldw * ldw) * 4  # rows [ldw, ldw + mp): Y = R L^-T
                w0 = (k * sys_stride + (ldw + mp) * ldw) * 4  # rows [ldw + mp, rows): W = L^-T
                check(self.lib.lit_gemm_tf32x3_nt_batched(
                    _vp(S_hi.data_ptr() + y0), _vp(S_lo.data_ptr() + y0), ldw, sys_stride,
                    _vp(S_hi.data_ptr() + w0), _vp(S_lo.data_ptr() + w0), ldw, sys_stride, n_rows, p, p, 1.0, _vp(0), 0, 0,
                    0.0, _vp(block.hi.data_ptr()), _vp(0), block.ld, n_rows * block.ld, nc, 1, s), "gemm_nt_batched")
                self.launches += 1
                self.gemm_flops += 1.0 * nc * n_rows * p * p  # upper-triangular W: half of 2 m n^2
                scratch, rel = t.empty((nc * p,), **f32), t.empty((2 * nc,), dtype=t.float64, device=self.device)
                check(self.lib.lit_spd_probe_residual(
                    nc, p, _vp(C.addressof(Gp) + 8 * k), ldg, _vp(C.addressof(Rp) + 8 * k), ldr,
                    _vp(C.addressof(mh) + 4 * k), _vp(C.addressof(a2h) + 4 * k), _vp(block.hi.data_ptr()), block.ld,
                    n_rows * block.ld, _vp(scratch.data_ptr()), _vp(rel.data_ptr()), s), "spd_probe_residual")
                self.launches += 2
                self._solver_checks.append((rel, info[k:k + nc]))
                k += nc

    # ------------------------------------------------------------------ eigendecomposition-free outer fit
    GROUP_TILE = 256  # rows per tile of the grouped GEMM (one CTA pair)

    def outer_inverses(self, G: Mat, lam_max: float, a2_list, series_ratio: float = 60.0):
        """(G + a^2 I)^-1 for every alpha of the grid: see outer_inverses_many."""
        return self.outer_inverses_many([G], [lam_max], [a2_list], series_ratio)[0]

    def outer_inverse_systems(self, lam_maxs, a2_lists, series_ratio: float = 60.0) -> list:
        """(Gram index, alpha slot) of every inverse that needs a Cholesky solve, in the order outer_inverses_many
        enumerates them (several ranks deal these out: `owned`)."""
        return [(i, j) for i, (lm, a2) in enumerate(zip(lam_maxs, a2_lists))
                for j in self.solver_partition(lm, a2, series_ratio)[0]]

    def inverse_slot(self, inv, j: int) -> list:
        """The two planes of alpha slot j of an inverse stack (for in-place collectives)."""
        return [inv[0][j], inv[1][j]]

    def outer_inverses_many(self, Gs, lam_maxs, a2_lists, series_ratio: float = 60.0, owned=None) -> list:
        """For each Gram G (fp32 Mat p x p; one per outer fold): (G + a^2 I)^-1 for every alpha of the grid as ONE stack
        of split pairs [n_alphas][p][ld] (the B operands of gemm_grouped).  The alphas below the series threshold -- of
        ALL Grams together -- go through the batched Cholesky solver (elimination of [G + a^2 I; I] gives W = L^-T,
        then W W^T); the others are 4-term Neumann polynomials in G, G^2, G^3.  Replaces the per-unique-alpha
        `Vh.T @ diag(S / (S^2 + a^2))` of ridge_regression.py:56-61 (no SVD / syevd).
        owned: indices into outer_inverse_systems() to solve here (None = all); the other slots are left unwritten."""
        t = self.torch
        p, ld = Gs[0].rows, Gs[0].ld
        if any(G.rows != p or G.ld != ld for G in Gs):
            raise ValueError("outer_inverses_many: the Grams must share one size and pitch")
        f32 = dict(dtype=t.float32, device=self.device)
        s = _vp(self.stream)
        parts = [self.solver_partition(lm, a2, series_ratio) for lm, a2 in zip(lam_maxs, a2_lists)]
        out = [(t.empty((len(a2), p, ld), **f32), t.empty((len(a2), p, ld), **f32), len(a2), p, ld) for a2 in a2_lists]
        systems = [(i, j) for i, (cheb, _) in enumerate(parts) for j in cheb]
        if owned is not None:  # several ranks: this rank's share; the caller broadcasts the slots afterwards
            systems = [sj for k, sj in enumerate(systems) if k in owned]
        for c0 in range(0, len(systems), self.SPD_MAX_BATCH):
            chunk = systems[c0:c0 + self.SPD_MAX_BATCH]
            nsys = len(chunk)
            fF, fS, fD, ldw, rows = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_long(0), C.c_long(0)
            check(self.lib.lit_spd_solve_workspace(nsys, p, 0, C.byref(fF), C.byref(fS), C.byref(fD), C.byref(ldw),
                                                   C.byref(rows)), "spd_solve_workspace")
            ldw, rows = ldw.value, rows.value
            F = t.empty((fF.value,), **f32)
            S_hi, S_lo = t.empty((fS.value,), **f32), t.empty((fS.value,), **f32)
            Dg_hi, Dg_lo = t.empty((fD.value,), **f32), t.empty((fD.value,), **f32)
            info = t.empty((nsys,), dtype=t.int32, device=self.device)
            Gp, Rp, mh, a2h = (C.c_void_p * nsys)(), (C.c_void_p * nsys)(), (C.c_int * nsys)(), (C.c_float * nsys)()
            for k, (i, j) in enumerate(chunk):
                Gp[k], Rp[k], mh[k], a2h[k] = Gs[i].hi.data_ptr(), Gs[i].hi.data_ptr(), 0, a2_lists[i][j]
            with self.timed("spd_solve"):
                check(self.lib.lit_spd_solve_batched(nsys, p, 0, C.cast(Gp, _vp), ld, C.cast(Rp, _vp), ld,
                                                     C.cast(mh, _vp), C.cast(a2h, _vp), _vp(F.data_ptr()),
                                                     _vp(S_hi.data_ptr()), _vp(S_lo.data_ptr()), _vp(Dg_hi.data_ptr()),
                                                     _vp(Dg_lo.data_ptr()), _vp(info.data_ptr()), s), "spd_solve_batched")
                self.launches += 1 + 4 * (ldw // 128)
                sys_stride = rows * ldw
                k = 0
                while k < nsys:  # runs of consecutive alpha slots of one Gram share a batched launch
                    k1 = k + 1
                    while k1 < nsys and chunk[k1][0] == chunk[k][0] and chunk[k1][1] == chunk[k1 - 1][1] + 1:
                        k1 += 1
                    i, j0 = chunk[k]
                    w0 = (k * sys_stride + ldw * ldw) * 4  # rows [ldw, 2 ldw) of system k: W = L^-T
                    check(self.lib.lit_gemm_tf32x3_nt_batched(
                        _vp(S_hi.data_ptr() + w0), _vp(S_lo.data_ptr() + w0), ldw, sys_stride,
                        _vp(S_hi.data_ptr() + w0), _vp(S_lo.data_ptr() + w0), ldw, sys_stride, p, p, p, 1.0, _vp(0), 0, 0,
                        0.0, _vp(out[i][0].data_ptr() + j0 * p * ld * 4), _vp(out[i][1].data_ptr() + j0 * p * ld * 4), ld,
                        p * ld, k1 - k, 1, s), "gemm_nt_batched")
                    self.launches += 1
                    self.gemm_flops += 1.0 * (k1 - k) * p * p * p
                    k = k1
            self._solver_checks.append((None, info))
        eye = None
        for i, (G, (cheb, series)) in enumerate(zip(Gs, parts)):
            if not series:
                continue
            Gsp = self.split(G)
            G2 = self.gemm(Gsp, Gsp, split_out=True, ld_out=ld)
            G3 = self.gemm(G2, Gsp, split_out=True, ld_out=ld)
            if eye is None:
                eye = t.eye(p, ld, **f32)  # plumbing: the q = 0 term of the series
            hi = (C.c_void_p * 4)(eye.data_ptr(), Gsp.hi.data_ptr(), G2.hi.data_ptr(), G3.hi.data_ptr())
            lo = (C.c_void_p * 4)(0, Gsp.lo.data_ptr(), G2.lo.data_ptr(), G3.lo.data_ptr())
            coef = np.array([[(-1.0) ** q / float(a2_lists[i][j]) ** (q + 1) for q in range(4)] for j in series],
                            dtype=np.float64)
            d_coef = self.upload_vector(coef.reshape(-1), "f64")
            d_slots = self.upload_vector(np.asarray(series), "i32")
            check(self.lib.lit_poly_combine(C.cast(hi, _vp), C.cast(lo, _vp), 4, ld, p, p, p, _vp(d_coef.data_ptr()),
                                            _vp(d_slots.data_ptr()), len(series), _vp(out[i][0].data_ptr()),
                                            _vp(out[i][1].data_ptr()), ld, s), "poly_combine")
            self.launches += 1
        return out

    def group_plan(self, idx, n_vox: int, n_groups: int):
        """Counting sort of the voxels by alpha index (lit_group_plan): (pos, perm, tile_group, rows_cap)."""
        tile = self.GROUP_TILE
        cap = (-(-max(n_vox, 1) // tile) + n_groups) * tile
        pos, perm, tg = self.vec(n_vox, "i32"), self.vec(cap, "i32"), self.vec(cap // tile, "i32")
        check(self.lib.lit_group_plan(_vp(idx.data_ptr()), n_vox, n_groups, tile, cap, _vp(pos.data_ptr()),
                                      _vp(perm.data_ptr()), _vp(tg.data_ptr()), _vp(self.stream)), "group_plan")
        self.launches += 1
        return pos, perm, tg, cap

    def gemm_grouped(self, A: Mat, inv, tile_group, split_out: bool = True) -> Mat:
        """out[rows of tile t] = A[rows of tile t] @ inv[tile_group[t]]^T (lit_gemm_tf32x3_nt_grouped)."""
        inv_hi, inv_lo, nA, p, ld = inv
        if A.cols != p or not A.is_split or A.rows % self.GROUP_TILE:
            raise ValueError("gemm_grouped: A must be a split pair of tile-padded rows with p columns")
        out = self.empty(A.rows, p, split=split_out)
        self._apply_sm_limit()
        with self.timed("gemm"):
            check(self.lib.lit_gemm_tf32x3_nt_grouped(
                _vp(A.hi.data_ptr()), _vp(A.lo.data_ptr()), A.ld, _vp(inv_hi.data_ptr()), _vp(inv_lo.data_ptr()), ld,
                p * ld, nA, A.rows, p, p, _vp(tile_group.data_ptr()), _vp(out.hi.data_ptr()),
                _vp(out.lo.data_ptr() if split_out else 0), out.ld, _vp(self.stream)), "gemm_grouped")
        self.launches += 1
        self.store_gemms += 1
        self.gemm_flops += 2.0 * A.rows * p * p
        return out

    def check_solver(self) -> None:
        """Raise SolverAccuracyError if a direct inner solve since the last check was not positive definite or
        failed its a-posteriori probe (synchronises; called once per fit next to check_eig)."""
        checks, self._solver_checks = self._solver_checks, []
        if not checks:
            return
        self.torch.cuda.synchronize(self.device)
        worst, bad = 0.0, 0
        for rel, info in checks:
            bad += int((info.cpu().numpy() != 0).sum())
            if rel is None:
                continue
            nd = rel.cpu().numpy().reshape(-1, 2)
            with np.errstate(divide="ignore", invalid="ignore"):
                r = np.where(nd[:, 1] > 0, np.sqrt(nd[:, 0] / nd[:, 1]), np.where(nd[:, 0] > 0, np.inf, 0.0))
            worst = max(worst, float(np.max(r)) if np.isfinite(r).all() else float("inf"))
        self.last_solver_residual = worst
        if bad or not (worst <= self.SOLVER_TOL):
            raise SolverAccuracyError(f"direct inner solve rejected: {bad} systems not positive definite, worst probe "
                                      f"residual {worst:.3e} (tolerance {self.SOLVER_TOL:.1e})")

    SERIES_MIN_ALPHAS = 5  # the compact stack pays once more alphas ride the series than it has terms (4)

    def assemble_series_stack(self, block: Mat, Pc: Mat, n_rows: int, rows_pad: int, lam_max: float, a2_list,
                              series_ratio: float = 60.0) -> SeriesStack:
        """Compact alpha stack from the block of solve_blocks: the Chebyshev solutions as ordinary groups, then the
        series tiles holding Q_q = P_c G^q / lam_max^q (q = 0..3) interleaved per 64 time points; the alphas of the
        series become 4-term combinations with coef[a][q] = (-1)^q lam_max^q / a2^(q+1)."""
        p = block.cols
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        n_tiles = -(-n_rows // 64)
        out = self.empty(len(cheb) * rows_pad + n_tiles * self.TILE_N, p, split=True)
        s = _vp(self.stream)
        for i in range(len(cheb)):
            off = i * rows_pad * out.ld * 4
            check(self.lib.lit_gather_rows_f32(_vp(block.hi.data_ptr() + i * n_rows * block.ld * 4), block.ld, _vp(0),
                                               n_rows, p, _vp(out.hi.data_ptr() + off), _vp(out.lo.data_ptr() + off),
                                               out.ld, rows_pad, s), "gather_rows")
            self.launches += 1
        if Pc.ld != block.ld:
            raise ValueError("assemble_series_stack: Pc must share the block's pitch")
        base = block.hi.data_ptr() + len(cheb) * n_rows * block.ld * 4
        hi = (C.c_void_p * 4)(Pc.hi.data_ptr(), *[base + q * n_rows * block.ld * 4 for q in range(3)])
        scale = (C.c_double * 4)(*[float(lam_max) ** -q for q in range(4)])
        off = len(cheb) * rows_pad * out.ld * 4
        check(self.lib.lit_series_stack(C.cast(hi, _vp), _vp(0), block.ld, n_rows, p, C.cast(scale, _vp), n_tiles,
                                        _vp(out.hi.data_ptr() + off), _vp(out.lo.data_ptr() + off), out.ld, s),
              "series_stack")
        self.launches += 1
        coef = np.array([[(-1.0) ** q * float(lam_max) ** q / float(a2_list[j]) ** (q + 1) for q in range(4)]
                         for j in series], dtype=np.float64)
        return SeriesStack(out, len(cheb), rows_pad, n_tiles, np.asarray(cheb, dtype=np.int32),
                           np.asarray(series, dtype=np.int32), coef)

    def assemble_stack(self, block: Mat, Pc: Mat, n_rows: int, rows_pad: int, lam_max: float, a2_list,
                       series_ratio: float = 60.0, series_moments: bool = False):
        if series_moments and len(self.solver_partition(lam_max, a2_list, series_ratio)[1]) >= self.SERIES_MIN_ALPHAS:
            return self.assemble_series_stack(block, Pc, n_rows, rows_pad, lam_max, a2_list, series_ratio)
        return self._assemble_full_stack(block, Pc, n_rows, rows_pad, lam_max, a2_list, series_ratio)

    def _assemble_full_stack(self, block: Mat, Pc: Mat, n_rows: int, rows_pad: int, lam_max: float, a2_list,
                             series_ratio: float = 60.0) -> Mat:
        """Alpha-stacked M_a = P_c (G + a^2 I)^-1 as a split pair [len(a2_list) * rows_pad][p] (pad rows zero) from
        the compact block of solve_blocks: copies of the Chebyshev solutions, Neumann-series combinations
        sum_q (-1)^q a^-2(q+1) P_c G^q for the large alphas."""
        p = block.cols
        cheb, series = self.solver_partition(lam_max, a2_list, series_ratio)
        out = self.empty(len(a2_list) * rows_pad, p, split=True)
        s = _vp(self.stream)
        for i, j in enumerate(cheb):
            off = j * rows_pad * out.ld * 4
            check(self.lib.lit_gather_rows_f32(_vp(block.hi.data_ptr() + i * n_rows * block.ld * 4), block.ld, _vp(0),
                                               n_rows, p, _vp(out.hi.data_ptr() + off), _vp(out.lo.data_ptr() + off),
                                               out.ld, rows_pad, s), "gather_rows")
            self.launches += 1
        if series:
            if Pc.ld != block.ld:
                raise ValueError("assemble_stack: Pc must share the block's pitch")
            base = block.hi.data_ptr() + len(cheb) * n_rows * block.ld * 4
            hi = (C.c_void_p * 4)(Pc.hi.data_ptr(), *[base + q * n_rows * block.ld * 4 for q in range(3)])
            coef = np.zeros((len(series), 4), dtype=np.float64)
            for g, j in enumerate(series):
                a2 = float(a2_list[j])
                coef[g] = [(-1.0) ** q / a2 ** (q + 1) for q in range(4)]
            d_coef = self.upload_vector(coef.reshape(-1), "f64")
            d_slots = self.upload_vector(np.asarray(series), "i32")
            check(self.lib.lit_poly_combine(C.cast(hi, _vp), _vp(0), 4, block.ld, n_rows, rows_pad, p,
                                            _vp(d_coef.data_ptr()), _vp(d_slots.data_ptr()), len(series),
                                            _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr()), out.ld, s), "poly_combine")
            self.launches += 1
        return out

    def inverse_stack(self, Gs: Mat, Pc: Mat, n_rows: int, rows_pad: int, lam_max: float, a2_list,
                      series_ratio: float = 60.0) -> Mat:
        """solve_blocks + assemble_stack on one rank."""
        block = self.solve_blocks(Gs, Pc, n_rows, lam_max, a2_list, series_ratio)
        return self.assemble_stack(block, Pc, n_rows, rows_pad, lam_max, a2_list, series_ratio)

    # ------------------------------------------------------------------ ridge kernels
    def build_alpha_stack(self, L: Mat, n_rows: int, rows_pad: int, lam, alphas_dev, n_alphas: int, normalpha: bool,
                          singcutoff: float) -> Mat:
        k = L.cols
        out = self.empty(n_alphas * rows_pad, k, split=True)
        col_mean = self.vec(k)
        scratch = self.vec(2 * k, "f64")
        check(self.lib.lit_build_alpha_stack(
            _vp(L.hi.data_ptr()), L.ld, n_rows, rows_pad, k, _vp(lam.data_ptr()), _vp(alphas_dev.data_ptr()), n_alphas,
            int(normalpha), singcutoff, _vp(col_mean.data_ptr()), _vp(scratch.data_ptr()), _vp(out.hi.data_ptr()),
            _vp(out.lo.data_ptr()), out.ld, _vp(self.stream)), "build_alpha_stack")
        self.launches += 3
        return out

    def scale_rows_by_alpha(self, Z: Mat, lam, alpha_v, normalpha: bool, singcutoff: float) -> Mat:
        out = self.empty(Z.rows, Z.cols, split=True, ld=Z.ld)
        check(self.lib.lit_scale_rows_by_alpha(
            _vp(Z.hi.data_ptr()), _vp(Z.lo.data_ptr() if Z.is_split else 0), Z.ld, Z.rows, Z.cols, _vp(lam.data_ptr()),
            _vp(alpha_v.data_ptr()), int(normalpha), singcutoff, _vp(out.hi.data_ptr()), _vp(out.lo.data_ptr()), out.ld,
            _vp(self.stream)), "scale_rows_by_alpha")
        self.launches += 1
        return out

    def corr_finalize(self, parts: Partials, tiles_per_group: int, n_groups: int, n_vox: int, n_rows: int, eps: float,
                      corr: Mat, accumulate: bool, metric: int = 0, resp_std=None) -> None:
        """Scores of all n_groups alphas into corr (rows = alpha slots).  A compact stack (parts.stack) is finalised
        in two launches: its ordinary groups through their slot map, the series alphas from the 14-sum partials."""
        ir, it = parts.inv_row, parts.inv_tile
        st = parts.stack
        rs = _vp(resp_std.data_ptr() if resp_std is not None else 0)
        n_plain = n_groups if st is None else st.n_cheb
        slots = None
        if st is not None and n_plain:
            slots = self.upload_vector(np.asarray(st.slot_cheb), "i32")
        if n_plain:
            check(self.lib.lit_corr_finalize_scaled(
                _vp(parts.dot.data_ptr()), _vp(parts.ssq.data_ptr()), parts.ld, tiles_per_group, n_plain, n_vox, n_rows,
                eps, int(accumulate), metric, rs, _vp(ir.data_ptr() if ir is not None else 0),
                _vp(it.data_ptr() if it is not None else 0), _vp(slots.data_ptr() if slots is not None else 0),
                _vp(corr.hi.data_ptr()), corr.ld, _vp(self.stream)), "corr_finalize")
            self.launches += 1
        if st is not None and st.n_tiles:
            coef = self.upload_vector(np.asarray(st.coef, dtype=np.float64).reshape(-1), "f64")
            sl = self.upload_vector(np.asarray(st.slot_series), "i32")
            tile0 = n_plain * tiles_per_group // 2  # first series tile among the stack's 256-row tiles
            it_s = it.data_ptr() + 4 * tile0 if it is not None else 0
            check(self.lib.lit_corr_finalize_series(
                _vp(parts.series.data_ptr()), parts.ld, 2 * st.n_tiles, n_vox, n_rows, eps, int(accumulate), metric, rs,
                _vp(ir.data_ptr() if ir is not None else 0), _vp(it_s), _vp(coef.data_ptr()), _vp(sl.data_ptr()),
                len(st.slot_series), _vp(corr.hi.data_ptr()), corr.ld, _vp(self.stream)), "corr_finalize_series")
            self.launches += 1

    def argmax_alpha(self, corr_sum: Mat, n_folds: int, alphas_dev, want_sums: bool):
        n_alphas, n_vox = corr_sum.rows, corr_sum.cols
        best, alpha_v = self.vec(n_vox, "i32"), self.vec(n_vox)
        sums = self.vec(n_alphas, "f64") if want_sums else None
        check(self.lib.lit_argmax_alpha(_vp(corr_sum.hi.data_ptr()), corr_sum.ld, n_alphas, n_vox, n_folds,
                                        _vp(alphas_dev.data_ptr()), _vp(best.data_ptr()), _vp(alpha_v.data_ptr()),
                                        _vp(sums.data_ptr() if want_sums else 0), _vp(self.stream)), "argmax_alpha")
        self.launches += 1
        return best, alpha_v, sums

    # ------------------------------------------------------------------ test statistics
    def pearson_finalize(self, parts: Partials, n_vox: int, n_samples: int, p_round_f32: bool):
        r, p = self.vec(n_vox), self.vec(n_vox, "f64")
        check(self.lib.lit_pearson_finalize(_vp(parts.dot.data_ptr()), _vp(parts.ssq.data_ptr()), parts.ld,
                                            parts.n_tiles, n_vox, n_samples, int(p_round_f32), _vp(r.data_ptr()),
                                            _vp(p.data_ptr()), _vp(self.stream)), "pearson_finalize")
        self.launches += 1
        return r, p

    def bh_fdr(self, p, n: int, alpha: float):
        need = C.c_size_t(0)
        check(self.lib.lit_bh_workspace(n, C.byref(need)), "bh_workspace")
        if self._bh_ws is None or self._bh_ws.numel() < need.value:
            self._bh_ws = self.torch.empty((need.value,), dtype=self.torch.uint8, device=self.device)
        reject, padj, count = self.vec(n, "u8"), self.vec(n, "f64"), self.vec(1, "i32")
        check(self.lib.lit_bh_fdr(_vp(p.data_ptr()), n, alpha, _vp(reject.data_ptr()), _vp(padj.data_ptr()),
                                  _vp(count.data_ptr()), _vp(self._bh_ws.data_ptr()), self._bh_ws.numel(),
                                  _vp(self.stream)), "bh_fdr")
        n_pad = max(2048, 1 << (n - 1).bit_length())
        stages = sum(1 + max(0, (k.bit_length() - 1) - 11) for k in
                     (1 << s for s in range(12, n_pad.bit_length())))
        self.launches += 3 + stages
        return reject, padj, count

    def fisher(self, p_stack, n_folds: int, n_vox: int, p_round_f32: bool):
        out = self.vec(n_vox, "f64")
        check(self.lib.lit_fisher_combine(_vp(p_stack.data_ptr()), p_stack.shape[1], n_folds, n_vox, int(p_round_f32),
                                          _vp(out.data_ptr()), _vp(self.stream)), "fisher_combine")
        self.launches += 1
        return out

    def stack_vectors(self, vecs, n: int, dtype: str = "f64"):
        """[len(vecs)][n] device array from device vectors (D2D copies; plumbing)."""
        t = self.torch
        dt = {"f32": t.float32, "f64": t.float64}[dtype]
        out = t.empty((len(vecs), max(n, 1)), dtype=dt, device=self.device)
        item = 8 if dtype == "f64" else 4
        for i, v in enumerate(vecs):
            check(self.lib.lit_memcpy_2d(_vp(out.data_ptr() + i * out.shape[1] * item), n * item, _vp(v.data_ptr()),
                                         n * item, n * item, 1, 3, _vp(self.stream)), "memcpy_2d(D2D)")
        return out

    # ------------------------------------------------------------------ feature construction
    def fir_make_delayed(self, stim: np.ndarray, delays, circpad: bool) -> np.ndarray:
        """Host array in, host float64 array out (FIR_expander.py:24-43 on the device)."""
        t = self.torch
        stim = np.ascontiguousarray(stim)
        if stim.dtype not in (np.float32, np.float64):
            stim = stim.astype(np.float64)
        nt, ndim = stim.shape
        nd = len(delays)
        out_h = np.empty((nt, nd * ndim), dtype=np.float64)
        if out_h.size == 0:
            return out_h
        d_stim = t.from_numpy(stim).to(self.device)
        d_del = self.upload_vector(np.asarray(delays, dtype=np.int64), "i32")
        d_out = t.empty((nt, nd * ndim), dtype=t.float64, device=self.device)
        check(self.lib.lit_fir_make_delayed(_vp(d_stim.data_ptr()), 0 if stim.dtype == np.float32 else 1, nt, ndim,
                                            ndim, _vp(d_del.data_ptr()), nd, int(bool(circpad)),
                                            _vp(d_out.data_ptr()), nd * ndim, _vp(self.stream)), "fir_make_delayed")
        self.launches += 1
        t.from_numpy(out_h).copy_(d_out)
        return out_h

    def row_view(self, m: Mat, r0: int, rows: int) -> Mat:
        """Rows [r0, r0 + rows) of a Mat as a Mat sharing its storage."""
        if r0 < 0 or rows < 0 or r0 + rows > m.rows:
            raise ValueError("row_view: rows outside the matrix")
        return self._view_rows(m, r0, rows)

    def upload_into(self, host: np.ndarray, out: Mat) -> None:
        """H2D of a 2-D float32 / float64 host block into the rows of `out` (float64 is converted on the device)."""
        host = np.asarray(host)
        if host.shape != (out.rows, out.cols):
            raise ValueError("upload_into: shape mismatch")
        if out.rows == 0 or out.cols == 0:
            return
        tmp = self.upload_matrix(host)
        check(self.lib.lit_memcpy_2d(_vp(out.hi.data_ptr()), out.ld * 4, _vp(tmp.hi.data_ptr()), tmp.ld * 4,
                                     out.cols * 4, out.rows, 3, _vp(self.stream)), "memcpy_2d(D2D)")

    def fir_zscore_rows(self, stim: np.ndarray, delays, circpad: bool, row_start: int, row_stop: int, zscore: bool,
                        out: Mat) -> None:
        """Rows [row_start, row_stop) of FIR.make_delayed(stim, delays), z-scored per column over those rows
        (population std; zero-std columns centred only; NaN -> 0) when `zscore`, written as fp32 into `out`
        (lit_fir_zscore_rows: trainer.py:203-209,236-239 fused; the float64 delayed matrix is never formed)."""
        stim, d_stim, dt = self._feature_in(stim)
        nt, ndim = stim.shape
        nd = len(delays)
        if out.rows != row_stop - row_start or out.cols != nd * ndim:
            raise ValueError("fir_zscore_rows: destination does not match")
        d_del = self.upload_vector(np.asarray(delays, dtype=np.int64), "i32")
        check(self.lib.lit_fir_zscore_rows(_vp(d_stim.data_ptr()), dt, nt, ndim, ndim, _vp(d_del.data_ptr()), nd,
                                           int(bool(circpad)), row_start, row_stop, int(bool(zscore)),
                                           _vp(out.hi.data_ptr()), out.ld, _vp(self.stream)), "fir_zscore_rows")
        self.launches += 1

    def as_tensor(self, m: Mat):
        """The logical [rows][cols] block of a Mat as a torch CUDA tensor (a view; no copy)."""
        return m.hi[: m.rows, : m.cols]

    def _feature_in(self, data: np.ndarray):
        data = np.ascontiguousarray(data)
        if data.dtype not in (np.float32, np.float64):
            data = data.astype(np.float64)
        self.h2d_bytes += data.nbytes
        return data, self.torch.from_numpy(data).to(self.device), 0 if data.dtype == np.float32 else 1

    def _feature_out(self, d_out) -> np.ndarray:
        out_h = np.empty(tuple(d_out.shape), dtype=np.float64)
        self.torch.from_numpy(out_h).copy_(d_out)
        self.d2h_bytes += out_h.nbytes
        return out_h

    def resample(self, kind: str, data: np.ndarray, data_times: np.ndarray, tr_times: np.ndarray, window: float,
                 cutoff: float, flag_a: bool, flag_b: bool, lo: Optional[np.ndarray], hi: Optional[np.ndarray]) -> np.ndarray:
        """Weighted resampling onto the TR grid.  kind "lanczos": flag_a = rectify; kind "sinc": flag_a = causal,
        flag_b = renorm.  Host arrays in, host float64 array out."""
        t = self.torch
        n_s, ndim = data.shape
        n_tr = len(tr_times)
        width = (2 if (kind == "lanczos" and flag_a) else 1) * ndim
        if n_tr * width == 0:
            return np.empty((n_tr, width), dtype=np.float64)
        if n_s == 0:
            return np.zeros((n_tr, width), dtype=np.float64)
        data, d_data, dt = self._feature_in(data)
        d_dt = self.upload_vector(data_times, "f64")
        d_tr = self.upload_vector(tr_times, "f64")
        d_lo = self.upload_vector(lo, "i32") if lo is not None else None
        d_hi = self.upload_vector(hi, "i32") if hi is not None else None
        d_out = t.empty((n_tr, width), dtype=t.float64, device=self.device)
        common = (_vp(d_data.data_ptr()), dt, n_s, ndim, ndim, _vp(d_dt.data_ptr()), _vp(d_tr.data_ptr()), n_tr,
                  float(window), float(cutoff))
        tail = (_vp(d_lo.data_ptr() if d_lo is not None else 0), _vp(d_hi.data_ptr() if d_hi is not None else 0),
                _vp(d_out.data_ptr()), width, _vp(self.stream))
        if kind == "lanczos":
            check(self.lib.lit_lanczos_downsample(*common, int(bool(flag_a)), *tail), "lanczos_downsample")
        elif kind == "sinc":
            check(self.lib.lit_sinc_downsample(*common, int(bool(flag_a)), int(bool(flag_b)), *tail), "sinc_downsample")
        else:
            raise ValueError(f"unknown resampling kind {kind}")
        self.launches += 1
        return self._feature_out(d_out)

    def lanczos_downsample(self, data, data_times, tr_times, window, cutoff, rectify, lo, hi) -> np.ndarray:
        return self.resample("lanczos", data, data_times, tr_times, window, cutoff, rectify, False, lo, hi)

    def csr_rows_apply(self, data: np.ndarray, row_ptr: np.ndarray, col_idx: np.ndarray,
                       weights: Optional[np.ndarray], mean: bool) -> np.ndarray:
        """out[r] = (mean of | sum of) weights[e] * data[col_idx[e]] over the entries of row r (float64)."""
        t = self.torch
        n_out, ndim = len(row_ptr) - 1, data.shape[1]
        if n_out * ndim == 0:
            return np.zeros((n_out, ndim), dtype=np.float64)
        data, d_data, dt = self._feature_in(data)
        d_rp = self.upload_vector(row_ptr, "i32")
        d_ci = self.upload_vector(col_idx, "i32")
        d_w = self.upload_vector(weights, "f64") if weights is not None else None
        d_out = t.empty((n_out, ndim), dtype=t.float64, device=self.device)
        check(self.lib.lit_csr_rows_apply(_vp(d_data.data_ptr()), dt, ndim, ndim, _vp(d_rp.data_ptr()),
                                          _vp(d_ci.data_ptr()), _vp(d_w.data_ptr() if d_w is not None else 0), n_out,
                                          int(bool(mean)), _vp(d_out.data_ptr()), ndim, _vp(self.stream)), "csr_rows_apply")
        self.launches += 1
        return self._feature_out(d_out)

    def gabor_downsample(self, data: np.ndarray, data_times: np.ndarray, tr_times: np.ndarray, freqs: np.ndarray,
                         sigma: float) -> np.ndarray:
        t = self.torch
        n_s, ndim = data.shape
        n_tr, n_f = len(tr_times), len(freqs)
        if n_tr * ndim * n_f == 0:
            return np.zeros((n_tr, ndim * n_f), dtype=np.float64)
        data, d_data, dt = self._feature_in(data)
        d_dt = self.upload_vector(data_times, "f64")
        d_tr = self.upload_vector(tr_times, "f64")
        d_f = self.upload_vector(freqs, "f64")
        d_out = t.empty((n_tr, ndim * n_f), dtype=t.float64, device=self.device)
        check(self.lib.lit_gabor_downsample(_vp(d_data.data_ptr()), dt, n_s, ndim, ndim, _vp(d_dt.data_ptr()),
                                            _vp(d_tr.data_ptr()), n_tr, _vp(d_f.data_ptr()), n_f, float(sigma),
                                            _vp(d_out.data_ptr()), ndim * n_f, _vp(self.stream)), "gabor_downsample")
        self.launches += 1
        return self._feature_out(d_out)


_default_ops: Optional[DeviceOps] = None


def default_ops() -> DeviceOps:
    """Process-wide DeviceOps on the current CUDA device (LOCAL_RANK-aware callers set the device first)."""
    global _default_ops
    import torch

    if _default_ops is None or _default_ops.device.index != torch.cuda.current_device():
        _default_ops = DeviceOps()
    return _default_ops
