"""Device-side engine: torch tensors as device buffers, libcpb200 kernels as the compute.

One ``Engine`` per process/GPU.  It owns ``nstreams`` CUDA streams, each with its own
libcpb200 handle (scratch workspace is per handle), so that independent layer problems
can overlap: the LASSO search is a single-CTA latency-bound kernel and the Cholesky
panels are small, while the Gram/trailing-update kernels fill the chip -- running
several layers concurrently on separate streams keeps the SMs busy.

Every method enqueues work on the *current* torch stream unless stated; nothing here
touches the CPU oracle, and there is no fallback: without a CUDA device or without
libcpb200.so construction fails.
"""
from __future__ import annotations

import os
import threading

import numpy as np
import torch

from . import _cabi

RAND_R_MAX = 2147483647
MAX_PROBES = 64
# Least squares from tensor-core (3xTF32) statistics: the ~4e-7 relative error of that Gram is amplified by the
# conditioning of the system (profiles/r2_conditioning.md maps it: 5e-6 of relative weight error at a pivot ratio of
# 0.2, 2e-4 at 0.01, 5e-3 at 1e-4), so every such solve is followed by ONE step of iterative refinement against the
# same factor with the residual taken from the data (engine.ls_refine) -- which squares the error -- and is accepted
# only while the smallest Cholesky pivot keeps at least LS_RATIO_MIN of its original diagonal entry (1 - R^2 of the
# most collinear column); below that the layer is re-solved from exact-product fp64 statistics.
LS_RATIO_MIN = float(os.environ.get("CPB200_LS_RATIO_MIN", "0.005"))
LS_REFINE = os.environ.get("CPB200_LS_REFINE", "1") == "1"
# Prediction X W' of the refinement residual: "fp64" (SIMT fp64 GEMM), "tc" (tensor cores, through cp_gram on the
# transposed patches) or "auto" (tensor cores for N >= 20000 only).  With the first-generation 3xTF32 kernel "tc" was
# a loss inside the 13-layer pipeline (59.6 vs 48.5 ms per step, profiles/r2_summary.md: one CTA per tile held a whole
# SM's shared memory for ~70 us and the latency-bound chains of the other layers queued behind them).  The second-
# generation kernel (gram_tc2.cu) is a ~100 us persistent launch: 32.1 vs 36.1 ms per step (call 19) -- the FP64 pipe
# is the step's scarce resource and this takes 2NK'n flop per layer off it.  Default "tc" (in the tensor-core mode).
LS_RESID = os.environ.get("CPB200_LS_RESID", "tc")
LS_RESID_TC_MIN_N = 20000
# Bulk products of the Cholesky solve on the tensor cores when the statistics came from there (cp_ls_tensor_cores)
LS_TC = os.environ.get("CPB200_LS_TC", "1") == "1"
# Full Gram of a layer enqueued after its channel search, on a lowest-priority stream (select_channels_async)
DEFER_FULL_GRAM = os.environ.get("CPB200_DEFER_GRAM", "1") == "1"

_PRIO_HIGHEST = -5  # cudaDeviceGetStreamPriorityRange on B200: [0, -5]; out-of-range values are clamped by the runtime
_LAYOUTS = {"nchw": 0, "nhwc": 1}
GRAM_FP64, GRAM_3XTF32 = 0, 1


class LassoResult:
    """Device-resident outputs of one channel selection (cp_lasso_select)."""

    __slots__ = ("idxs", "coef", "scalars", "probe_log", "seeds")

    def __init__(self, idxs, coef, scalars, probe_log, seeds):
        self.idxs, self.coef, self.scalars, self.probe_log, self.seeds = idxs, coef, scalars, probe_log, seeds


def gpu_numa_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (from sysfs), or None when the platform does not say."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


class numa_local:
    """Context: page-locked host buffers allocated inside are placed on the NUMA node of ``device_index`` (first
    touch under a temporary CPU affinity; a GPU reading host memory of the other socket pays the inter-socket
    link on every PCIe read -- round-1 measurement: 0.75 e2e scaling efficiency at 8 GPUs from one node's memory)."""

    def __init__(self, device_index):
        self.cpus = gpu_numa_cpus(device_index)
        self.old = None

    def __enter__(self):
        if self.cpus:
            try:
                self.old = os.sched_getaffinity(0)
                allowed = self.cpus & self.old
                if allowed:
                    os.sched_setaffinity(0, allowed)
                else:
                    self.old = None
            except Exception:
                self.old = None
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except Exception:
                pass
        return False


class Engine:
    def __init__(self, device=None, nstreams: int = 1, gram_mode: int = None):
        if not torch.cuda.is_available():
            raise RuntimeError("cpb200: no CUDA device visible; the solver has no CPU path")
        self.ffi, self.lib = _cabi.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if gram_mode is None:  # tensor cores by default; CPB200_GRAM=fp64 selects the exact-product DFMA mode
            gram_mode = GRAM_FP64 if os.environ.get("CPB200_GRAM", "tc").lower() == "fp64" else GRAM_3XTF32
        self.gram_mode = gram_mode
        self._handles = []
        self.streams = []
        with torch.cuda.device(self.device):
            for i in range(max(1, nstreams)):
                hp = self.ffi.new("cp_handle_t*")
                _cabi.check(self.lib.cp_create(hp, self.device.index))
                self._handles.append(hp[0])
                # The pipeline hands its most expensive problems to the first slots.  CPB200_STREAM_PRIORITY:
                #   "1"      (default) two levels: the first half of the slots high
                #   "graded" one level per slot from the highest down (B200: -5 .. 0)        "0"  none
                # measured on the 13-layer step (profiles/r2_summary.md): 48.3 / 48.9 / 49.0 ms -- priorities only order
                # CTAs that are not yet resident, and the step is bound by FP64 throughput, not by the order
                pol = os.environ.get("CPB200_STREAM_PRIORITY", "1")
                if pol == "graded":
                    prio = min(0, _PRIO_HIGHEST + i)
                else:
                    prio = -2 if (i < nstreams // 2 and pol == "1") else -1  # 0 is left to the deferred-Gram stream
                self.streams.append(torch.cuda.Stream(self.device, priority=prio) if nstreams > 1 else None)
        self._tls = threading.local()   # current (handle, stream) slot of each host thread (use_slot)
        self._lock = threading.Lock()
        self.launches = 0  # libcpb200 calls issued (each launches >= 1 kernel)
        self._pinned = {}   # (key, shape, dtype) -> page-locked host buffer, allocated once (cudaHostAlloc is slow)
        self._staging = {}  # (key, shape) -> device staging buffer for maps that are cheaper to DMA whole
        self._xfer = None   # (zero-copy gather stream, DMA stream) of the host-resident input path
        self._aux = None    # (handle, lowest-priority stream) of the deferred full Grams (select_channels_async)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        for h in self._handles:
            self.lib.cp_destroy(h)
        self._handles = []
        if self._aux is not None:
            self.lib.cp_destroy(self._aux[0])
            self._aux = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def pinned(self, key, shape, dtype):
        """Engine-owned page-locked host buffer, reused across calls (contents valid until the next call that
        asks for the same key)."""
        k = (key, tuple(shape), dtype)
        t = self._pinned.get(k)
        if t is None:
            with numa_local(self.device.index if self.device.index is not None else 0):
                t = self._pinned[k] = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
        return t

    def _ring(self):
        """Rotating index for small pinned staging buffers: a buffer is rewritten by the host only 256 calls
        after the asynchronous copy that read it was enqueued (every step synchronises on each layer's search)."""
        with self._lock:
            self._seq = (getattr(self, "_seq", -1) + 1) % 256
            return self._seq

    def staging(self, key, shape):
        k = (key, tuple(shape))
        t = self._staging.get(k)
        if t is None:
            t = self._staging[k] = torch.empty(tuple(shape), dtype=torch.float32, device=self.device)
        return t

    def xfer_streams(self):
        if self._xfer is None:
            self._xfer = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        return self._xfer

    def use_slot(self, i: int):
        """Select which (handle, stream) pair subsequent calls OF THE CALLING THREAD use (the selection is
        thread-local: pruner.prune_layers issues the reconstructions of different layers from worker threads)."""
        self._tls.cur = i % len(self._handles)
        return self.streams[self._tls.cur]

    def slot_lock(self, i: int):
        """Lock of slot i's (handle, stream) pair: a handle's scratch is reused by every call in stream order, so two
        host threads must not interleave their launches on one slot."""
        with self._lock:
            if not hasattr(self, "_slot_locks"):
                self._slot_locks = [threading.Lock() for _ in self._handles]
        return self._slot_locks[i % len(self._handles)]

    @property
    def _cur(self):
        return getattr(self._tls, "cur", 0)

    @property
    def h(self):
        over = getattr(self._tls, "handle", None)
        return over if over is not None else self._handles[self._cur]

    def aux_slot(self):
        """(handle, stream) for work that may run in the shadow of the channel searches: the lowest stream priority, its
        own handle (scratch), created on first use."""
        if self._aux is None:
            with torch.cuda.device(self.device):
                hp = self.ffi.new("cp_handle_t*")
                _cabi.check(self.lib.cp_create(hp, self.device.index))
                self._aux = (hp[0], torch.cuda.Stream(self.device, priority=0))
        return self._aux

    def _s(self):
        return self.ffi.cast("void*", torch.cuda.current_stream(self.device).cuda_stream)

    def _p(self, t, ctype):
        if t is None:
            return self.ffi.NULL
        if not t.is_cuda:
            # pinned (page-locked) host memory is mapped into the device address space under UVA: kernels may
            # read it in place over PCIe -- used by the sparse gathers, which touch a small part of each map
            assert t.is_pinned(), "host tensors must be pinned to be read by a kernel"
        else:
            assert t.device == self.device, "tensor must live on %s" % self.device
        return self.ffi.cast(ctype, t.data_ptr())

    def _call(self, rc):
        self.launches += 1
        _cabi.check(rc)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ kernels
    def patch_gather(self, fmap, randx, randy, B, P, k, pad, stride, relu=True, layout="nchw", out=None):
        """fmap: (nbatch*B, c, H, W) [nchw] or (nbatch*B, H, W, c) [nhwc] fp32 on device;
        randx/randy: (nbatch, P) int32 on device.  Returns X (nbatch*P*B, c*k*k) fp32."""
        assert fmap.dtype == torch.float32 and fmap.is_contiguous()
        nimg = fmap.shape[0]
        assert nimg % B == 0
        nbatch = nimg // B
        if layout == "nchw":
            c, H, W = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        else:
            H, W, c = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        assert randx.dtype == torch.int32 and randx.numel() == nbatch * P and randx.is_contiguous()
        assert randy.dtype == torch.int32 and randy.numel() == nbatch * P and randy.is_contiguous()
        rows, K = nbatch * P * B, c * k * k
        if out is None:
            out = self.empty(rows, K, dtype=torch.float32)
        assert out.shape == (rows, K) and out.dtype == torch.float32 and out.stride(1) == 1
        self._call(self.lib.cp_patch_gather(self.h, self._p(fmap, "const float*"), nbatch, B, c, H, W,
                                            _LAYOUTS[layout], self._p(randx, "const int32_t*"),
                                            self._p(randy, "const int32_t*"), P, k, pad, stride, int(bool(relu)),
                                            self._p(out, "float*"), out.stride(0), self._s()))
        return out

    def point_gather(self, fmap, randx, randy, B, P, layout="nchw", out=None):
        assert fmap.dtype == torch.float32 and fmap.is_contiguous()
        nimg = fmap.shape[0]
        nbatch = nimg // B
        if layout == "nchw":
            n, H, W = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        else:
            H, W, n = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        rows = nbatch * P * B
        if out is None:
            out = self.empty(rows, n, dtype=torch.float32)
        self._call(self.lib.cp_point_gather(self.h, self._p(fmap, "const float*"), nbatch, B, n, H, W,
                                            _LAYOUTS[layout], self._p(randx, "const int32_t*"),
                                            self._p(randy, "const int32_t*"), P, self._p(out, "float*"),
                                            out.stride(0), self._s()))
        return out

    def gram(self, X, Y=None, y_bias=None, rows=None, want_G=True, want_B=True, want_sums=True, want_yy=False,
             mode=None):
        """Sufficient statistics of (X, Y) (optionally over gathered rows).
        Returns dict with G (K,K), B (K,n), sx (K), sy (n), yy (1) as fp64 device tensors."""
        assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1
        N, K = X.shape
        out = {}
        n = 0
        y_dt = 0
        if Y is not None:
            assert Y.dim() == 2 and Y.shape[0] == N and Y.stride(1) == 1
            assert Y.dtype in (torch.float32, torch.float64)
            y_dt = 0 if Y.dtype == torch.float32 else 1
            n = Y.shape[1]
        if rows is not None:
            assert rows.dtype == torch.int32 and rows.is_contiguous()
        G = self.empty(K, K) if want_G else None
        Bxy = self.empty(K, n) if (want_B and Y is not None) else None
        sx = self.empty(K) if want_sums else None
        sy = self.empty(n) if (want_sums and Y is not None) else None
        yy = self.empty(1) if (want_yy and Y is not None) else None
        self._call(self.lib.cp_gram(self.h, self._p(X, "const float*"), N, K, X.stride(0),
                                    self._p(Y, "const void*"), y_dt, n, Y.stride(0) if Y is not None else 0,
                                    self._p(y_bias, "const float*"), self._p(rows, "const int32_t*"),
                                    rows.numel() if rows is not None else 0, self._p(G, "double*"),
                                    self._p(Bxy, "double*"), self._p(sx, "double*"), self._p(sy, "double*"),
                                    self._p(yy, "double*"), self.gram_mode if mode is None else mode, self._s()))
        out.update(G=G, B=Bxy, sx=sx, sy=sy, yy=yy, N=N, K=K, n=n, mode=self.gram_mode if mode is None else mode)
        return out

    def gemm_tc_split(self, A, B, C=None, alpha=1.0, beta=0.0, lower=False, b_nc=False):
        """C = alpha A B' + beta C on the tensor cores in 22-bit split precision (cp_gemm_tc_split): A (M, R), B (Nn, R)
        -- or (R, Nn) with b_nc -- C (M, Nn), fp64 device tensors with unit inner stride."""
        assert A.dtype == B.dtype == torch.float64 and A.stride(1) == 1 and B.stride(1) == 1
        M, R = A.shape
        Nn = B.shape[1] if b_nc else B.shape[0]
        assert (B.shape[0] if b_nc else B.shape[1]) == R
        if C is None:
            assert beta == 0.0
            C = self.empty(M, Nn)
        assert C.shape == (M, Nn) and C.dtype == torch.float64 and C.stride(1) == 1
        self._call(self.lib.cp_gemm_tc_split(self.h, M, Nn, R, float(alpha), self._p(A, "const double*"), A.stride(0),
                                             self._p(B, "const double*"), B.stride(0), float(beta), self._p(C, "double*"),
                                             C.stride(0), (1 if lower else 0) | (2 if b_nc else 0), self._s()))
        return C

    def ls_tensor_cores(self, enable):
        """Bulk products of the following ls_solve / ls_factor / ls_resolve calls on the tensor cores (cp_ls_tensor_cores)."""
        self._call(self.lib.cp_ls_tensor_cores(self.h, 1 if enable else 0))

    def gram_profile(self, enable=True):
        """CUDA events around the tensor-core GEMM launch of every following gram() on this engine's handle."""
        self._call(self.lib.cp_gram_profile(self.h, 1 if enable else 0))

    def gram_kernel_ms(self):
        """Elapsed time (ms) of the tensor-core GEMM launch of the last gram() (needs gram_profile(True))."""
        ms = self.ffi.new("float*")
        self._call(self.lib.cp_gram_kernel_ms(self.h, ms))
        return float(ms[0])

    def lasso_build(self, gs, gw, W2m, c, k2, S):
        """gs: gram over sampled rows (with yy); gw: gram of W2 viewed (n, K); W2m: (n, K) fp32."""
        n = W2m.shape[0]
        ldq = c + (c & 1)  # even leading dimension: 16-byte aligned rows for the selection kernel
        Q = self.empty(c, ldq)
        qv = self.empty(c)
        yn2 = self.empty(1)
        self._call(self.lib.cp_lasso_build(self.h, self._p(gs["G"], "const double*"), self._p(gs["B"], "const double*"),
                                           self._p(gs["sx"], "const double*"), self._p(gs["sy"], "const double*"),
                                           self._p(gs["yy"], "const double*"), self._p(gw["G"], "const double*"),
                                           self._p(gw["sx"], "const double*"), self._p(W2m, "const float*"), c, k2, n,
                                           S, self._p(Q, "double*"), ldq, self._p(qv, "double*"),
                                           self._p(yn2, "double*"), self._s()))
        return Q[:, :c], qv, yn2

    def lasso_select(self, Q, qv, yn2, m, rank, lbound, rbound, right0, seeds, tol=1e-4, max_iter=1000):
        c = Q.shape[0]
        if Q.stride(1) != 1 or Q.stride(0) % 2 or Q.data_ptr() % 16:
            Qp = torch.zeros(c, c + (c & 1), dtype=torch.float64, device=self.device)
            Qp[:, :c] = Q
            Q = Qp[:, :c]
        if isinstance(seeds, torch.Tensor) and seeds.is_cuda:
            assert seeds.dtype == torch.int32 and seeds.is_contiguous()
            seeds_d = seeds
        else:
            # staged through an engine-owned pinned buffer: a pageable H2D copy would make torch synchronise
            # the stream, i.e. block the host until everything queued on this layer's stream has run
            sh = np.asarray(seeds, dtype=np.int64).astype(np.int32)
            pin = self.pinned(("seeds", self._ring()), sh.shape, torch.int32)
            pin.numpy()[...] = sh
            seeds_d = pin.to(self.device, non_blocking=True)
        maxp = seeds_d.numel()
        idxs = self.empty(c, dtype=torch.uint8)
        coef = self.empty(c)
        scalars = self.empty(4)
        plog = torch.zeros(maxp, 4, dtype=torch.float64, device=self.device)
        self._call(self.lib.cp_lasso_select(self.h, self._p(Q, "const double*"), Q.stride(0),
                                            self._p(qv, "const double*"),
                                            self._p(yn2, "const double*"), c, float(m), int(rank), float(lbound),
                                            float(rbound), float(right0), float(tol), int(max_iter),
                                            self._p(seeds_d, "const uint32_t*"), maxp, self._p(idxs, "uint8_t*"),
                                            self._p(coef, "double*"), self._p(scalars, "double*"),
                                            self._p(plog, "double*"), self._s()))
        return LassoResult(idxs, coef, scalars, plog, seeds_d)

    def ls_solve(self, g, sel_cols):
        """Centred normal-equation LS on columns ``sel_cols`` (int32 device tensor, ascending; None = all).
        Returns (W (n, Ksel) fp64, b (n,) fp64, info (1,) int32, stat (1,) fp64) on device; stat is the smallest
        pivot / original-diagonal ratio of the Cholesky (conditioning signal, see include/cpb200.h)."""
        Ks = sel_cols.numel() if sel_cols is not None else g["K"]
        n = g["n"]
        W = self.empty(n, Ks)
        b = self.empty(n)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        stat = self.empty(1)
        self.ls_tensor_cores(LS_TC and g.get("mode", GRAM_FP64) != GRAM_FP64)
        self._call(self.lib.cp_ls_solve(self.h, self._p(g["G"], "const double*"), self._p(g["B"], "const double*"),
                                        self._p(g["sx"], "const double*"), self._p(g["sy"], "const double*"),
                                        g["N"], g["K"], n, self._p(sel_cols, "const int32_t*"), Ks,
                                        self._p(W, "double*"), self._p(b, "double*"), self._p(info, "int32_t*"),
                                        self._p(stat, "double*"), self._s()))
        return W, b, info, stat

    def ls_solve_dual(self, X, Y, y_bias, sel_cols):
        N, K = X.shape
        n = Y.shape[1]
        Ks = sel_cols.numel() if sel_cols is not None else K
        W = self.empty(n, Ks)
        b = self.empty(n)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        stat = self.empty(1)
        self._call(self.lib.cp_ls_solve_dual(self.h, self._p(X, "const float*"), N, K, X.stride(0),
                                             self._p(Y, "const void*"), 0 if Y.dtype == torch.float32 else 1, n,
                                             Y.stride(0), self._p(y_bias, "const float*"),
                                             self._p(sel_cols, "const int32_t*"), Ks, self._p(W, "double*"),
                                             self._p(b, "double*"), self._p(info, "int32_t*"),
                                             self._p(stat, "double*"), self._s()))
        return W, b, info, stat

    def ls_factor(self, g, sel_cols=None):
        """Keeps the Cholesky factor of the centred Gram (columns ``sel_cols``) on the CURRENT handle for
        subsequent ls_resolve calls.  Returns (info (1,) int32, stat (1,) fp64)."""
        Ks = sel_cols.numel() if sel_cols is not None else g["K"]
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        stat = self.empty(1)
        self.ls_tensor_cores(LS_TC and g.get("mode", GRAM_FP64) != GRAM_FP64)
        self._call(self.lib.cp_ls_factor(self.h, self._p(g["G"], "const double*"), self._p(g["sx"], "const double*"),
                                         g["N"], g["K"], self._p(sel_cols, "const int32_t*"), Ks,
                                         self._p(info, "int32_t*"), self._p(stat, "double*"), self._s()))
        return info, stat

    def ls_resolve(self, Bxy, sx, sy, sel_cols=None, Ks=None, accumulate_into=None):
        """Solve against the factor kept on the CURRENT handle (ls_factor, or the last ls_solve): Bxy (K, n) = X'U,
        sy (n,) = 1'U.  Returns (W (n, Ksel), b (n,)); with accumulate_into=(W, b) the solution is ADDED to those."""
        n = Bxy.shape[1]
        Ks = (sel_cols.numel() if sel_cols is not None else Bxy.shape[0]) if Ks is None else Ks
        if accumulate_into is None:
            W, b, acc = self.empty(n, Ks), self.empty(n), 0
        else:
            (W, b), acc = accumulate_into, 1
            assert W.shape == (n, Ks) and W.is_contiguous() and b.shape == (n,)
        self._call(self.lib.cp_ls_resolve(self.h, self._p(Bxy, "const double*"), self._p(sx, "const double*"),
                                          self._p(sy, "const double*"), n, self._p(sel_cols, "const int32_t*"),
                                          self._p(W, "double*"), self._p(b, "double*"), acc, self._s()))
        return W, b

    def ls_residual(self, X, Y, y_bias, sel_cols, W, b, mode=None):
        """fp32 residual (N, n) of the fit (W, b) on columns sel_cols, computed from the data (cp_ls_residual): prediction
        in fp64 (mode 0) or on the tensor cores (mode 1, default = the engine's Gram mode)."""
        N, K = X.shape
        n = Y.shape[1]
        Ks = sel_cols.numel() if sel_cols is not None else K
        R = self.empty(N, n, dtype=torch.float32)
        self._call(self.lib.cp_ls_residual(self.h, self._p(X, "const float*"), N, K, X.stride(0), self._p(Y, "const void*"),
                                           0 if Y.dtype == torch.float32 else 1, n, Y.stride(0),
                                           self._p(y_bias, "const float*"), self._p(sel_cols, "const int32_t*"), Ks,
                                           self._p(W, "const double*"), self._p(b, "const double*"), self._p(R, "float*"),
                                           R.stride(0), self.gram_mode if mode is None else mode, self._s()))
        return R

    def ls_refine(self, g, X, Y, y_bias, sel_cols, W, b):
        """One step of iterative refinement of (W, b) against the factor the last ls_solve left on this handle:
        residual from the data (exact fp64), its cross products with X on the tensor cores, forward/backward
        substitution, correction added in place.  Removes the error tensor-core statistics put into the solution."""
        tc = LS_RESID == "tc" or (LS_RESID == "auto" and X.shape[0] >= LS_RESID_TC_MIN_N)
        R = self.ls_residual(X, Y, y_bias, sel_cols, W, b, mode=g["mode"] if tc else GRAM_FP64)
        gr = self.gram(X, R, want_G=False, mode=g["mode"])
        self.ls_resolve(gr["B"], g["sx"], gr["sy"], sel_cols, accumulate_into=(W, b))
        return W, b

    # ------------------------------------------------------------------ data-form LASSO (benchmark kernel)
    def lasso_dataform_build(self, X, W2m, Y, y_bias, samples, c, k2):
        """Materialises the reference's design matrix Z ((S*n) x c, column major fp32) and target y (lib/decompose.py:
        428-437).  Returns (Z as a (c, m) tensor whose rows are the columns of Z, y (m,) fp64)."""
        N, K = X.shape
        n = W2m.shape[0]
        S = samples.numel()
        m = S * n
        Z = self.empty(c, m, dtype=torch.float32)
        y = self.empty(m)
        self._call(self.lib.cp_lasso_dataform_build(self.h, self._p(X, "const float*"), N, K, X.stride(0),
                                                    self._p(W2m, "const float*"), n, c, k2,
                                                    self._p(samples, "const int32_t*"), S, self._p(Y, "const void*"),
                                                    0 if Y.dtype == torch.float32 else 1, Y.stride(0),
                                                    self._p(y_bias, "const float*"), self._p(Z, "float*"), Z.stride(0),
                                                    self._p(y, "double*"), self._s()))
        return Z, y

    def lasso_cd_dataform(self, Z, y, alpha, seed, w=None, tol=1e-4, max_iter=1000):
        """One Lasso.fit in data form (cp_lasso_cd_dataform).  Returns (w (c,), scalars [n_iter, gap, tol, sweeps, checks])."""
        c, m = Z.shape
        if w is None:
            w = torch.zeros(c, dtype=torch.float64, device=self.device)
        out = torch.zeros(8, dtype=torch.float64, device=self.device)
        self._call(self.lib.cp_lasso_cd_dataform(self.h, self._p(Z, "const float*"), Z.stride(0), self._p(y, "const double*"),
                                                 m, c, float(alpha), float(tol), int(max_iter), int(seed) & 0xffffffff,
                                                 self._p(w, "double*"), self._p(out, "double*"), self._s()))
        return w, out

    def dataform_benchmark(self, X, W2m, Y, y_bias, samples, shape, alpha, seed=12345):
        """Times one cold-start data-form fit at ``alpha`` (profiles/sweep_config5.py): bytes of Z streamed per second."""
        Z, y = self.lasso_dataform_build(X, W2m, Y, y_bias, samples, shape.c, shape.k * shape.k)
        self.lasso_cd_dataform(Z, y, alpha, seed)  # warm-up
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        w, out = self.lasso_cd_dataform(Z, y, alpha, seed)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        o = out.cpu().numpy()
        sweeps, checks = int(o[3]), int(o[4])
        nbytes = 4.0 * Z.shape[0] * Z.shape[1] * (sweeps + checks)  # 4 m c per sweep (+ per gap check)
        return {"m": int(Z.shape[1]), "c": int(Z.shape[0]), "alpha": alpha, "n_iter": int(o[0]), "sweeps": sweeps,
                "gap_checks": checks, "ms": ms, "stream_gbs": nbytes / (ms / 1e3) / 1e9, "nnz": int((w != 0).sum().item()),
                "algorithmic_bytes": nbytes}

    # ------------------------------------------------------------------ dense fp64 blocks (3C companions)
    def gemm(self, A, B, a_mc=False, b_nc=False, alpha=1.0, beta=0.0, out=None):
        """C[m, nn] = alpha * sum_r a(m, r) b(nn, r) + beta * C   (cp_gemm_f64; fp64 device tensors, unit inner stride)
        a_mc: A is (R, M) else (M, R);  b_nc: B is (R, Nn) else (Nn, R)."""
        assert A.dtype == torch.float64 and B.dtype == torch.float64 and A.dim() == 2 and B.dim() == 2
        assert (A.stride(1) == 1 or A.shape[1] == 1) and (B.stride(1) == 1 or B.shape[1] == 1)
        R, M = (A.shape[0], A.shape[1]) if a_mc else (A.shape[1], A.shape[0])
        Rb, Nn = (B.shape[0], B.shape[1]) if b_nc else (B.shape[1], B.shape[0])
        assert R == Rb, "inner dimensions differ: %d vs %d" % (R, Rb)
        if out is None:
            assert beta == 0.0
            out = self.empty(M, Nn)
        assert out.shape == (M, Nn) and out.dtype == torch.float64 and (out.stride(1) == 1 or Nn == 1)
        self._call(self.lib.cp_gemm_f64(self.h, int(a_mc), int(b_nc), M, Nn, R, float(alpha), self._p(A, "const double*"),
                                        A.stride(0), self._p(B, "const double*"), B.stride(0), float(beta),
                                        self._p(out, "double*"), out.stride(0), self._s()))
        return out

    def mm(self, A, B):
        """A @ B"""
        return self.gemm(A, B, a_mc=False, b_nc=True)

    def mm_nt(self, A, B):
        """A @ B.T"""
        return self.gemm(A, B, a_mc=False, b_nc=False)

    def mm_tn(self, A, B):
        """A.T @ B"""
        return self.gemm(A, B, a_mc=True, b_nc=True)

    def svd(self, F, max_sweeps=40):
        """Thin SVD of a small dense fp64 matrix (m, n) by one-sided Jacobi (cp_svd_jacobi).  Returns (U (m, r),
        s (r,) descending, Vh (r, n)), r = min(m, n) -- the convention of scipy.linalg.svd(full_matrices=False)."""
        assert F.dtype == torch.float64 and F.dim() == 2
        m, n = F.shape
        if m < n:  # orthogonalise the columns of the taller orientation
            U, s, Vh = self.svd(F.T.contiguous(), max_sweeps)
            return Vh.T.contiguous(), s, U.T.contiguous()
        Ft = F.T.contiguous()
        Wt = self.empty(n, n)
        sigma = self.empty(n)
        sweeps = self.ffi.new("int32_t*")
        tol = float(np.sqrt(m)) * 2.220446049250313e-16
        self._call(self.lib.cp_svd_jacobi(self.h, self._p(Ft, "double*"), m, n, Ft.stride(0), self._p(Wt, "double*"),
                                          Wt.stride(0), self._p(sigma, "double*"), 1, tol, max_sweeps, sweeps,
                                          self._s()))
        if sweeps[0] >= max_sweeps:
            raise np.linalg.LinAlgError("Jacobi SVD did not converge in %d sweeps" % max_sweeps)
        order = torch.argsort(sigma, descending=True, stable=True)
        return Ft[order].T.contiguous(), sigma[order].contiguous(), Wt[order].contiguous()

    def solve_relu(self, RUraw, bias, Z, lam, want_mean=False):
        """U = solve_relu(RUraw + bias, Z, lam) (lib/decompose.py:51-59), optional column means of U."""
        N, n = RUraw.shape
        U = self.empty(N, n)
        mean = self.empty(n) if want_mean else None
        self._call(self.lib.cp_solve_relu(self.h, self._p(RUraw, "const double*"), RUraw.stride(0),
                                          self._p(bias, "const double*"), self._p(Z, "const double*"), Z.stride(0),
                                          float(lam), self._p(U, "double*"), U.stride(0), N, n, self._p(mean, "double*"),
                                          self._s()))
        return (U, mean) if want_mean else U

    def colstats(self, X, scale=1.0, centre=False):
        """(scale * column sums, X - that) of an fp64 matrix (cp_colstats_f64); centre with scale = 1/N."""
        N, n = X.shape
        assert X.dtype == torch.float64 and X.stride(1) == 1
        cs = self.empty(n)
        Xc = self.empty(N, n) if centre else None
        self._call(self.lib.cp_colstats_f64(self.h, self._p(X, "const double*"), X.stride(0), N, n, float(scale),
                                            self._p(cs, "double*"), self._p(Xc, "double*"), n, self._s()))
        return (cs, Xc) if centre else cs

    # ------------------------------------------------------------------ composite: one layer problem
    def select_channels_async(self, X, W2m, Y, y_bias, samples, c, k2, rank, rank_tol, right0, seeds):
        """Everything of decompose.dictionary up to (and including) the alpha search, enqueued
        without host synchronisation.  Returns (g_full, LassoResult)."""
        n = W2m.shape[0]
        S = samples.numel()
        # The search needs only the channel-space statistics (Q from the sampled rows and W2); the full Gram feeds the
        # reconstruction.  In the pipeline (one stream per layer) it is therefore enqueued AFTER the search, on a
        # lowest-priority stream with its own handle: the ~0.5 ms full-GPU launch no longer delays the start of this
        # and of every later layer's 8 ms single-SM search, it runs in their shadow.
        defer = DEFER_FULL_GRAM and self.streams[0] is not None
        if defer:
            cur = torch.cuda.current_stream(self.device)
            ev_x = torch.cuda.Event()
            ev_x.record(cur)  # X and Y are complete here (recorded BEFORE the search is enqueued on this stream)
        g_full = None if defer else self.gram(X, Y, y_bias=y_bias)
        g_s = self.gram(X, Y, y_bias=y_bias, rows=samples, want_yy=True, mode=GRAM_FP64)
        g_w = self.gram(W2m, None, want_B=False, mode=GRAM_FP64)
        Q, qv, yn2 = self.lasso_build(g_s, g_w, W2m, c, k2, S)
        lbound, rbound = window(rank, rank_tol)
        res = self.lasso_select(Q, qv, yn2, float(S) * n, rank, lbound, rbound, right0, seeds)
        if defer:
            ah, astream = self.aux_slot()
            self._tls.handle = ah
            try:
                with torch.cuda.stream(astream):
                    astream.wait_event(ev_x)
                    g_full = self.gram(X, Y, y_bias=y_bias)
                    ready = torch.cuda.Event()
                    ready.record(astream)
            finally:
                self._tls.handle = None
            for t in (X, Y, y_bias):
                if t is not None:
                    t.record_stream(astream)
            for key in ("G", "B", "sx", "sy"):
                if g_full.get(key) is not None:
                    g_full[key].record_stream(cur)
            g_full["ready"] = ready
        return g_full, res

    def _cols_device(self, idxs_host, k2, K):
        sel = np.flatnonzero(idxs_host)
        cols = (sel[:, None] * k2 + np.arange(k2)[None, :]).reshape(-1).astype(np.int32)
        pin = self.pinned(("cols", self._ring()), (K,), torch.int32)  # no pageable (synchronising) copy
        pin.numpy()[:cols.size] = cols
        return pin[:cols.size].to(self.device, non_blocking=True)

    def reconstruct_async(self, g_full, X, Y, y_bias, idxs_host, k2):
        """LS on the surviving channels (device outputs; no host sync).  Returns (W, b, info, stat)."""
        if g_full.get("ready") is not None:  # full Gram enqueued on the deferred-Gram stream
            torch.cuda.current_stream(self.device).wait_event(g_full["ready"])
        cols_d = self._cols_device(idxs_host, k2, g_full["K"])
        if g_full["N"] - 1 >= cols_d.numel():
            W, b, info, stat = self.ls_solve(g_full, cols_d)
            if g_full["mode"] != GRAM_FP64 and LS_REFINE:
                # statistics from the 3xTF32 Gram carry ~4e-7 relative error, which the conditioning of a wide layer
                # amplifies to ~4e-5 in W and more in b (measured at conv4_x, N=5000): one refinement step against
                # the same factor, with the residual taken from the data, restores fp64-level accuracy
                self.ls_refine(g_full, X, Y, y_bias, cols_d, W, b)
            return W, b, info, stat
        return self.ls_solve_dual(X, Y, y_bias, cols_d)

    def reconstruct_exact_async(self, X, Y, y_bias, idxs_host, k2):
        """The same solve from exact-product fp64 statistics (the slow path of the conditioning policy)."""
        g = self.gram(X, Y, y_bias=y_bias, mode=GRAM_FP64)
        return self.reconstruct_async(g, X, Y, y_bias, idxs_host, k2)

    def reconstruct_truncated(self, X, Y, y_bias, idxs_host, k2):
        """Minimum-norm least squares with the reference's rank cut-off, for systems the Cholesky flags as numerically
        rank deficient: LinearRegression.fit -> scipy.linalg.lstsq(Xc, Yc, cond=1e-6) (gelsd, sklearn _base.py:752)
        drops singular values below 1e-6 sigma_max.  Here: one-sided Jacobi SVD of the centred selected columns
        (cp_svd_jacobi works on X itself, so small singular values are resolved like gelsd resolves them), then
        W = V diag(1/s) U' Yc over the kept ones.  Slow path (O(10) sweeps over N x K'), rare."""
        cols = self._cols_device(idxs_host, k2, X.shape[1]).long()
        N = X.shape[0]
        Xs = X[:, cols].to(torch.float64).contiguous()
        xm, Xc = self.colstats(Xs, 1.0 / N, centre=True)
        Yd = Y.to(torch.float64)
        if y_bias is not None:
            Yd = Yd - y_bias.to(torch.float64)[None, :]
        ym, Yc = self.colstats(Yd.contiguous(), 1.0 / N, centre=True)
        U, s, Vh = self.svd(Xc)
        keep = s > 1e-6 * s[0]
        inv = torch.where(keep, 1.0 / torch.where(keep, s, torch.ones_like(s)), torch.zeros_like(s))
        UtY = self.mm_tn(U, Yc)                                   # (r, n)
        W = self.mm_tn((inv[:, None] * UtY).contiguous(), Vh)    # (n, K') = (V diag(1/s) U' Yc)'
        b = ym - self.mm(W, xm[:, None].contiguous())[:, 0]
        return W, b, int(keep.sum().item())

    @staticmethod
    def ls_verdict(info, stat, mode, dual=False):
        """'ok' | 'redo' (tensor-core statistics too inaccurate for this conditioning: re-solve in fp64) |
        'singular' (a pivot fell below sklearn's rank cut-off even with exact statistics)."""
        if mode == GRAM_FP64 or dual:
            return "singular" if info else "ok"
        if info or not (stat >= LS_RATIO_MIN):
            return "redo"
        return "ok"


def window(rank, rank_tol):
    """Acceptance window of the alpha search, reference lib/decompose.py:492-501."""
    lbound = rank
    if rank_tol >= 1:
        rbound = rank + rank_tol
    else:
        rbound = rank + rank_tol * rank
        if rank_tol == .2:
            lbound = rank + 0.1 * rank
            rbound = rank + 0.2 * rank
    return lbound, rbound


_ENGINE = None


def get_engine(**kw) -> Engine:
    """Process-wide engine on the current CUDA device."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = Engine(**kw)
    return _ENGINE


def reset_engine():
    global _ENGINE
    if _ENGINE is not None:
        _ENGINE.close()
    _ENGINE = None
