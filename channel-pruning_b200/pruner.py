"""Layer-problem driver: one network's independent pruning problems on one or more B200s.

Within a GPU, problems are pipelined over the engine's streams in two phases (the only host
synchronisation points): phase 1 enqueues gather -> Gram statistics -> LASSO search for every
problem; phase 2, once the kept-channel masks are known on the host, enqueues the
least-squares reconstructions.  Across GPUs (one process per GPU, torch.distributed/NCCL) the
problems are assigned by longest-processing-time-first on an analytic cost and the packed
results are re-assembled on every rank with ONE all_gather (static upper-bound payload size, so
no size exchange is needed).  There is no data-path collective: layer problems are independent
 -- in the reference's *sequential* R3 mode (lib/net.py:1386,1698) they are
coupled and this sharding does not apply (use lib.net.Net.R3 on one GPU).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .engine import Engine, window


# ---------------------------------------------------------------------------- assignment
# host threads that issue the reconstructions of phase 2 (CPB200_PHASE2_THREADS=1: the single-threaded polling loop)
PHASE2_THREADS = int(os.environ.get("CPB200_PHASE2_THREADS", "6"))
_POOLS = {}


def _phase2_pool(n):
    import concurrent.futures

    p = _POOLS.get(n)
    if p is None:
        p = _POOLS[n] = concurrent.futures.ThreadPoolExecutor(max_workers=n, thread_name_prefix="cpb200-phase2")
    return p


def assign_layers(costs, world_size):
    """LPT: returns owner[i] for each problem; deterministic (ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


# ---------------------------------------------------------------------------- packing
def slot_size(c, n, k2, rank, rank_tol):
    """Upper bound (in float64 words) of one packed layer result."""
    _, rbound = window(rank, rank_tol)
    cmax = c if rank == c else min(c, int(np.floor(rbound)))
    return 4 + c + n + n * cmax * k2


def pack_result(buf, offset, idxs, W, b, alpha, nprobe, c, n, k2, eng=None, slot=0):
    """buf: 1-D float64 tensor (any device).  Layout: [c', alpha, nprobe, 0, idxs(c), b(n), W(n*c'*k2)].
    With ``eng`` the host-side scalars/mask are staged through an engine-owned PINNED buffer and every copy is
    asynchronous on the current stream (a pageable copy would make torch synchronise the stream: the pattern
    engine.lasso_select documents as a serialising bug); W / b may live on the device or in pinned host memory."""
    cp = int(idxs.sum())
    if eng is not None:
        head = eng.pinned(("pack", slot), (4 + c,), torch.float64)
        hn = head.numpy()
        hn[0], hn[1], hn[2], hn[3] = cp, alpha, nprobe, 0.0
        hn[4:] = idxs
        buf[offset:offset + 4 + c].copy_(head, non_blocking=True)
        buf[offset + 4 + c:offset + 4 + c + n].copy_(b.reshape(-1), non_blocking=True)
        buf[offset + 4 + c + n:offset + 4 + c + n + n * cp * k2].copy_(W.reshape(-1), non_blocking=True)
        return
    head = torch.tensor([cp, alpha, nprobe, 0.0], dtype=torch.float64)
    buf[offset:offset + 4] = head.to(buf.device)
    buf[offset + 4:offset + 4 + c] = torch.as_tensor(idxs.astype(np.float64)).to(buf.device)
    buf[offset + 4 + c:offset + 4 + c + n] = b.reshape(-1).to(buf.device, torch.float64)
    buf[offset + 4 + c + n:offset + 4 + c + n + n * cp * k2] = W.reshape(-1).to(buf.device, torch.float64)


def unpack_result(buf, offset, c, n, k2):
    head = buf[offset:offset + 4].cpu().numpy()
    cp = int(head[0])
    idxs = buf[offset + 4:offset + 4 + c].cpu().numpy() != 0
    b = buf[offset + 4 + c:offset + 4 + c + n].cpu().numpy()
    k = int(round(np.sqrt(k2)))
    W = buf[offset + 4 + c + n:offset + 4 + c + n + n * cp * k2].cpu().numpy().reshape(n, cp, k, k)
    return dict(idxs=idxs, W=W, b=b, alpha=float(head[1]), nprobe=int(head[2]))


def allgather_results(local_buf, world_size, group=None):
    """ONE collective: every rank contributes a buffer of identical (upper-bound) length."""
    import torch.distributed as dist

    out = torch.empty(world_size * local_buf.numel(), dtype=local_buf.dtype, device=local_buf.device)
    dist.all_gather_into_tensor(out, local_buf, group=group)
    return out.view(world_size, -1)


# ---------------------------------------------------------------------------- single-GPU pipeline
class LayerResult:
    __slots__ = ("idxs", "W", "b", "alpha", "nprobe", "probes", "info")


def prune_layers(eng: Engine, shapes, datas, right0=1e-3, rank_tol=.1, from_host=False, to_host=False, trace=None):
    """Runs the layer problems ``shapes[i]`` / ``datas[i]`` (see synth.make_problem_device) on
    ``eng``.  from_host: feature maps are taken from pinned host memory (datas[i]['fmap_host'])
    and copied in the pipeline; to_host: results are copied back to pinned host memory.
    trace: optional dict; filled with {layer name: [(label, timing event), ...]} plus '_t0' (device timeline
    of the step: profiles/e2e_breakdown.py prints it).
    Batch mode: the problems are INDEPENDENT -- every alpha search starts from ``right0`` and the seeds come with the
    problem (datas[i]['seeds']), so the call neither reads nor writes ``cfgs.alpha`` and draws nothing from numpy's
    global RNG.  The reference's sequential walk carries alpha from layer to layer (lib/decompose.py:491,627) and draws
    one seed per probe; that behaviour lives in lib.decompose.dictionary / Net.R3, which run the layers one after the
    other.  Masks and alphas of the two modes are therefore not comparable layer by layer.
    Every reconstruction is verified before it is returned (Cholesky status + pivot ratio, see the end of
    _prune_layers_ordered): r.info['verdict'] is 'ok', 'redo->ok' (re-solved from exact-product statistics) or
    'truncated' (rank deficient: gelsd's minimum-norm solution).
    Returns a list of LayerResult (W, b as device fp64 tensors unless to_host)."""
    nslots = len(eng.streams)
    main = torch.cuda.current_stream(eng.device)
    # longest problems first: their (latency-bound) LASSO searches start early and the short ones fill the GPU
    order = sorted(range(len(shapes)), key=lambda i: (-shapes[i].cost(), i))
    inv = {orig: pos for pos, orig in enumerate(order)}
    shapes_o = [shapes[i] for i in order]
    datas_o = [datas[i] for i in order]
    if trace is not None:
        trace["_t0"] = torch.cuda.Event(enable_timing=True)
        trace["_t0"].record()
    res_o = _prune_layers_ordered(eng, shapes_o, datas_o, right0, rank_tol, from_host, to_host, main, trace)
    return [res_o[inv[i]] for i in range(len(shapes))]


def _zero_copy_seconds(s):
    """Model of the in-place gather over PCIe: it is bound by the number of read requests (one per 128-byte
    line touched: the k rows of a window are W*4 bytes apart), ~4.5e8 lines/s measured (profiles/r1c_summary.md)."""
    row = s.W * 4 if hasattr(s, "W") else 4 * 64
    lines = min(s.k, -(-((s.k - 1) * row + s.k * 4) // 128) + 1) if s.k > 1 else 1
    return s.N * s.c * lines / 4.5e8


def h2d_plan(shapes, datas, from_host):
    """Per layer: 'zc' (gather kernel reads the windows in place from pinned host memory) or 'dma' (copy engine
    moves the whole map at full PCIe bandwidth, gather from HBM).  DMA pays off when the windows cover most of
    the map (small spatial maps: conv5_x).  CPB200_DMA_MAX_MB caps the size of a map that may be staged.
    Tried and measured worse (profiles/r2_summary.md): also staging one conv4_x map (802 MB) on the copy engine next to
    the reader -- 86.1 instead of 78.7 ms per step; the two paths share the link (the reader slows down 1.7x while a
    copy is in flight), so moving work between them buys nothing unless it removes bytes."""
    if from_host == "zc":
        return ["zc"] * len(shapes)
    if from_host == "copy":
        return ["dma"] * len(shapes)
    cap = float(os.environ.get("CPB200_DMA_MAX_MB", "300")) * 1e6
    ratio = float(os.environ.get("CPB200_DMA_RATIO", "0.8"))
    plan = []
    for s, d in zip(shapes, datas):
        nbytes = d["fmap_host"].numel() * 4
        t_dma = nbytes / 50e9 + 1e-4
        plan.append("dma" if (nbytes <= cap and t_dma < ratio * _zero_copy_seconds(s)) else "zc")
    return plan


def _mark(trace, name, label):
    if trace is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        trace.setdefault(name, []).append((label, e))


def _prune_layers_ordered(eng, shapes, datas, right0, rank_tol, from_host, to_host, main, trace=None):
    # ---- host-resident inputs: PCIe is the shared resource, so the transfers are issued in the (longest-first)
    # layer order on two dedicated streams -- zero-copy gathers on one, whole-map DMAs on the copy engine --
    # and each layer stream waits only for its own transfer.
    pre = [None] * len(shapes)
    if from_host and len(eng.streams) > 1 and eng.streams[0] is not None:
        plan = h2d_plan(shapes, datas, from_host)
        zc_stream, dma_stream = eng.xfer_streams()
        zc_stream.wait_stream(main)
        dma_stream.wait_stream(main)
        dma_order = sorted((i for i in range(len(shapes)) if plan[i] == "dma"),
                           key=lambda i: (datas[i]["fmap_host"].numel(), i))
        for i in dma_order:
            with torch.cuda.stream(dma_stream):
                st = eng.staging(("fmap", i), datas[i]["fmap_host"].shape)
                st.copy_(datas[i]["fmap_host"], non_blocking=True)
                _mark(trace, shapes[i].name, "dma_done")
                ev = torch.cuda.Event()
                ev.record()
            pre[i] = ("dma", st, ev)
        for i, (s, d) in enumerate(zip(shapes, datas)):
            if plan[i] != "zc":
                continue
            eng.use_slot(i)
            with torch.cuda.stream(zc_stream):
                X = eng.patch_gather(d["fmap_host"], d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True)
                _mark(trace, s.name, "zc_done")
                ev = torch.cuda.Event()
                ev.record()
            pre[i] = ("zc", X, ev)
    phase1 = []
    for i, (s, d) in enumerate(zip(shapes, datas)):
        stream = eng.use_slot(i)
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with ctx:
            if stream is not None:
                stream.wait_stream(main)
            X = None
            layout = "nchw"
            if pre[i] is not None:
                kind, obj, ev = pre[i]
                stream.wait_event(ev)
                if kind == "zc":
                    X = obj
                else:
                    fmap = obj
            elif from_host == "copy":
                fmap = eng.staging(("fmap", i), d["fmap_host"].shape)
                fmap.copy_(d["fmap_host"], non_blocking=True)
            elif from_host:
                # zero-copy: the gather kernel reads the sampled windows straight out of pinned host memory;
                # only the touched 32-byte sectors cross PCIe (1-40 % of a feature map), not the whole map
                fmap = d["fmap_host"]
            else:
                fmap = d["fmap"]
                layout = d.get("layout", "nchw")  # host copies keep the reference's NCHW blob order
            if X is None:
                X = eng.patch_gather(fmap, d["randx"], d["randy"], s.B, s.P, s.k, s.pad, s.stride, relu=True,
                                     layout=layout)
            W2m = d["W2"].reshape(s.n, s.K)
            if s.rank == s.c:
                g_full = eng.gram(X, d["feats"], y_bias=d["b2"])
                res = None
                host = None
            else:
                g_full, res = eng.select_channels_async(X, W2m, d["feats"], d["b2"], d["samples"], s.c, s.k * s.k,
                                                        s.rank, rank_tol, right0, d["seeds"])
                host = (eng.pinned(("scal", i), (4,), torch.float64), eng.pinned(("idxs", i), (s.c,), torch.uint8))
                host[0].copy_(res.scalars, non_blocking=True)
                host[1].copy_(res.idxs, non_blocking=True)
            _mark(trace, s.name, "select_done")
            ev = torch.cuda.Event()
            ev.record()
        phase1.append((X, g_full, res, host, ev))
    out = [None] * len(shapes)
    # phase 2 in COMPLETION order of the searches (which layer finishes first depends on sizes and, with
    # host-resident inputs, on the transfer order): poll the events, reconstruct whichever is ready.
    # A reconstruction of a wide layer is ~390 kernel launches (~1.3 ms of host time inside libcpb200 calls, which
    # release the GIL): issued from one thread, the five c = 512 layers of VGG-16 queued behind one another on the
    # HOST (step timeline, call 23: the last one was not even issued until 7.5 ms after its search had finished).
    # Worker threads issue them side by side; the engine's current (handle, stream) slot is thread-local.
    checks = []

    def reconstruct_layer(i):
        s, d = shapes[i], datas[i]
        X, g_full, res, host, ev = phase1[i]
        if res is not None:
            ev.synchronize()
        torch.cuda.set_device(eng.device)
        stream = eng.use_slot(i)
        r = LayerResult()
        if res is None:
            r.idxs = np.ones(s.c, dtype=bool)
            r.alpha, r.nprobe = 1e-4, 0  # lib/decompose.py:386 argument default survives (:627)
        else:
            scal = host[0].numpy()
            if int(scal[2]) != 0:
                raise RuntimeError("layer %s: alpha search hit the probe cap" % s.name)
            r.idxs = host[1].numpy().astype(bool)
            r.alpha, r.nprobe = float(scal[0]), int(scal[1])
        ctx = torch.cuda.stream(stream) if stream is not None else _null()
        with eng.slot_lock(i), ctx:  # layers that share a slot (more layers than streams) are issued one after the other
            W, b, info, stat = eng.reconstruct_async(g_full, X, d["feats"], d["b2"], r.idxs, s.k * s.k)
            chk = (eng.pinned(("lsinfo", i), (1,), torch.int32), eng.pinned(("lsstat", i), (1,), torch.float64))
            chk[0].copy_(info, non_blocking=True)
            chk[1].copy_(stat, non_blocking=True)
            r.W, r.b = _maybe_to_host(eng, i, W, b, to_host)  # eager: the copy overlaps the other layers' solves
            ev_ls = torch.cuda.Event()
            ev_ls.record()
            _mark(trace, s.name, "ls_done")
        r.probes = res
        r.info = {"mode": g_full["mode"], "dual": not (g_full["N"] - 1 >= int(r.idxs.sum()) * s.k * s.k)}
        out[i] = r
        return (i, chk, ev_ls)

    nworkers = min(PHASE2_THREADS, len(shapes)) if eng.streams[0] is not None else 1
    pool = _phase2_pool(nworkers) if nworkers > 1 else None
    futures = []
    pending = list(reversed(range(len(shapes))))
    while pending:
        ready = [i for i in pending if phase1[i][2] is None or phase1[i][4].query()]
        if not ready:
            time.sleep(2e-5)
            continue
        i = ready[0]
        pending.remove(i)
        if pool is not None:
            futures.append(pool.submit(reconstruct_layer, i))
        else:
            checks.append(reconstruct_layer(i))
    checks.extend(f.result() for f in futures)
    # ---- every reconstruction is CHECKED before it is handed out: Cholesky status + conditioning signal.
    # Tensor-core statistics are accepted only for well-conditioned systems (engine.LS_RATIO_MIN); otherwise the
    # layer is re-solved from exact-product fp64 statistics; a system that is rank deficient by sklearn's cut-off
    # gets the truncated minimum-norm solution the reference's gelsd returns.
    for i, chk, ev_ls in checks:
        ev_ls.synchronize()
        s, d, r = shapes[i], datas[i], out[i]
        fail, ratio = int(chk[0][0]), float(chk[1][0])
        verdict = eng.ls_verdict(fail, ratio, r.info["mode"], r.info["dual"])
        r.info.update(pivot_ratio=ratio, verdict=verdict)
        if verdict == "redo":
            stream = eng.use_slot(i)
            ctx = torch.cuda.stream(stream) if stream is not None else _null()
            with ctx:
                X = phase1[i][0]
                W, b, info, stat = eng.reconstruct_exact_async(X, d["feats"], d["b2"], r.idxs, s.k * s.k)
                fail, ratio = int(info.cpu()[0]), float(stat.cpu()[0])
                r.W, r.b = _maybe_to_host(eng, i, W, b, to_host)
            verdict = "singular" if fail else "ok"
            r.info.update(pivot_ratio_exact=ratio, verdict="redo->" + verdict)
        if verdict == "singular":  # gelsd's truncated minimum-norm solution (slow path, see Engine.reconstruct_truncated)
            stream = eng.use_slot(i)
            ctx = torch.cuda.stream(stream) if stream is not None else _null()
            with ctx:
                W, b, kept = eng.reconstruct_truncated(phase1[i][0], d["feats"], d["b2"], r.idxs, s.k * s.k)
                r.W, r.b = _maybe_to_host(eng, i, W, b, to_host)
            r.info.update(verdict="truncated", rank=kept)
    for st in eng.streams:
        if st is not None:
            main.wait_stream(st)
    return out


def _maybe_to_host(eng, i, W, b, to_host):
    if not to_host:
        return W, b
    # engine-owned pinned buffers: valid until the next prune_layers call on this engine
    Wh = eng.pinned(("W", i), W.shape, torch.float64)
    bh = eng.pinned(("b", i), b.shape, torch.float64)
    Wh.copy_(W, non_blocking=True)
    bh.copy_(b, non_blocking=True)
    return Wh, bh


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


# ---------------------------------------------------------------------------- multi-GPU
def prune_network_sharded(eng: Engine, shapes, make_data, rank, world_size, right0=1e-3, rank_tol=.1, group=None):
    """Each rank solves its LPT share and all ranks end with every layer's result.
    make_data(i) -> data dict for problem i (only called for owned problems)."""
    owner = assign_layers([s.cost() for s in shapes], world_size)
    mine = [i for i, o in enumerate(owner) if o == rank]
    datas = [make_data(i) for i in mine]
    res = prune_layers(eng, [shapes[i] for i in mine], datas, right0=right0, rank_tol=rank_tol)
    sizes = [slot_size(s.c, s.n, s.k * s.k, s.rank, rank_tol) for s in shapes]
    # static layout: rank r's buffer holds its problems in index order; pad to the largest rank buffer
    per_rank = [sum(sizes[i] for i in range(len(shapes)) if owner[i] == r) for r in range(world_size)]
    buf = torch.zeros(max(per_rank), dtype=torch.float64, device=eng.device)
    off = 0
    for j, i in enumerate(mine):
        s = shapes[i]
        pack_result(buf, off, res[j].idxs, res[j].W, res[j].b, res[j].alpha, res[j].nprobe, s.c, s.n, s.k * s.k,
                    eng=eng, slot=j)
        off += sizes[i]
    if world_size > 1:
        allbuf = allgather_results(buf, world_size, group)
    else:
        allbuf = buf.view(1, -1)
    return owner, sizes, allbuf


def unpack_network(shapes, owner, sizes, allbuf):
    offs = [0] * allbuf.shape[0]
    out = [None] * len(shapes)
    for i, s in enumerate(shapes):
        r = owner[i]
        out[i] = unpack_result(allbuf[r], offs[r], s.c, s.n, s.k * s.k)
        offs[r] += sizes[i]
    return out
