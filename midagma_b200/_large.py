"""Multi-CTA engine behind ``DagmaLinear`` for d > 64 and for the logistic loss.

One inner iteration of ``DagmaLinear.minimize`` (src/dagma/linear.py:224-276) is a fixed
sequence of kernel launches from libdagma_b200.so working on a 19-double device state
block (mu, s, lr, bias-correction powers, iteration counter, feasibility latch):

    fused slogdet+inverse  ->  score GEMM(s)  [-> all-reduce when rows are sharded]
    ->  fused Gobj/Adam/step/mask/feasibility

captured once as a CUDA graph and replayed; the host only synchronises at the
convergence checkpoints (every ``checkpoint`` iterations), exactly where the reference
evaluates the objective (linear.py:279-331).  Infeasible inverses latch ``halted`` on
the device; the (rare) undo / halve-lr / redo logic of linear.py:230-241 then runs on
the host with the same kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

ST_DOUBLES = 19
(F_MU, F_S, F_LR, F_LAM, F_B1, F_B2, F_P1H, F_P1L, F_P2H, F_P2L, F_LAD, F_H, F_MIN, F_SCORE, F_L1, F_LOSS,
 F_GSCALE) = range(17)
I_IT, I_HALTED, I_INFO = 0, 1, 2          # int32 slots inside doubles 17..18


def gemm(A, B, C, *, trans_a=False, alpha=1.0, beta=0.0, epilogue=0, ws=None):
    """C = alpha * op(A) @ B + beta * C on device (FP64 DMMA kernel); 2-D contiguous tensors."""
    K, N = B.shape
    M = A.shape[1] if trans_a else A.shape[0]
    assert (A.shape[0] if trans_a else A.shape[1]) == K and C.shape == (M, N)
    assert A.is_contiguous() and B.is_contiguous() and C.is_contiguous()
    _lib.check(_lib.load().dagma_gemm_f64(
        _lib.stream_ptr(), int(trans_a), M, N, K, float(alpha), A.data_ptr(), A.shape[1], B.data_ptr(), N,
        float(beta), C.data_ptr(), N, int(epilogue), ws.data_ptr() if ws is not None else None,
        ws.numel() * 8 if ws is not None else 0), "dagma_gemm_f64")
    return C


def capture_graph(fn) -> "torch.cuda.CUDAGraph":
    """Capture ``fn()`` into a CUDA graph on a side stream -- ``torch.cuda.graph`` without its prologue (device
    synchronise, ``gc.collect``, ``torch.cuda.empty_cache``).  The iteration works on pre-allocated buffers, so nothing
    has to be released for the capture, and emptying the allocator's cache costs 150 ms per capture at d = 2000
    (measured: 0.77 s of a 4.9 s reduced-schedule C5 fit, whose five stages re-capture because ``s`` is a launch
    argument of the inverse)."""
    g = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    cur.wait_stream(side)
    return g


class LargeLinearEngine:
    def __init__(self, model, group=None):
        _lib.require_device()
        self.lib = _lib.load()
        self.loss_type = model.loss_type
        self.d, self.n = model.d, model.n
        self.dev = model._device
        self.cov = model._cov_dev.contiguous()
        self.X = model._X_dev.contiguous() if self.loss_type == "logistic" else None
        self.mask_exc, self.mask_inc = model._mask_exc, model._mask_inc
        self.group = group                       # torch.distributed group when rows of X are sharded
        # the per-iteration all-reduce of the d x d partial gradient is captured INSIDE the iteration's CUDA graph when
        # the backend is NCCL (graph-capturable): the sharded iteration then costs one graph launch like the
        # single-GPU one instead of ~6 kernel launches + one NCCL launch from the host (DAGMA_GRAPH_NCCL=0: eager)
        self._graph_collectives = False
        if group is not None:
            import os
            import torch.distributed as dist
            self._graph_collectives = (dist.get_backend(group) == "nccl"
                                       and os.environ.get("DAGMA_GRAPH_NCCL", "1") != "0")
        self.n_total = getattr(model, "_n_total", self.n)
        d = self.d
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.W = torch.zeros(d, d, **f64)
        self.Minv = torch.empty(d, d, **f64)
        self.T = torch.empty(d, d, **f64)
        self.m = torch.zeros(d, d, **f64)
        self.v = torch.zeros(d, d, **f64)
        self.state = torch.zeros(ST_DOUBLES, **f64)
        self.state_host = torch.zeros(ST_DOUBLES, dtype=torch.float64).pin_memory()
        ws_bytes = self.lib.dagma_large_workspace_bytes(d)
        self.ws = torch.empty(ws_bytes // 8 + 8, **f64)
        self.kws = torch.empty(148 * d * d + 8, **f64) if self.loss_type == "logistic" else None
        if self.X is not None:
            self.R = torch.empty(self.n, d, **f64)
            self.partial = torch.empty(1024, **f64)
        self.obj_ws = torch.zeros(self.lib.dagma_linear_objective_workspace_bytes() // 8 + 1, **f64)
        self._graph = None
        self._side = None
        self._model_cov_ptr = model._cov_dev.data_ptr()
        self.launches_per_iter = None
        self._trek_setup(getattr(model, "_trek_plan", None))
        # d <= 128, no trek regulariser: every iteration up to the next checkpoint is ONE persistent kernel,
        # csrc/lin_iter.cu (DAGMA_LIN_FUSED=0: the launch sequence below, replayed as a graph).  Rows sharded over the
        # GPUs of one box: the same kernel on every GPU, the d x d partial products summed INSIDE it over NVLink peer
        # memory (DAGMA_LIN_PEER=0: launch sequence + NCCL all-reduce)
        import os
        logistic = int(self.loss_type == "logistic")
        self._peer = None
        self.one_kernel = (self.trek is None and os.environ.get("DAGMA_LIN_FUSED", "1") != "0"
                           and bool(self.lib.dagma_linear_iter_supported(logistic, self.n if logistic else 0, d)))
        if group is not None:
            self.one_kernel = self._peer_setup(group, self.one_kernel and logistic == 1
                                               and os.environ.get("DAGMA_LIN_PEER", "1") != "0")
        if self.one_kernel:
            self.iter_ws = torch.empty(self.lib.dagma_linear_iter_workspace_doubles(
                logistic, self.n if logistic else 0, d), **f64)
            self.iter_sync = torch.zeros(4, dtype=torch.int32, device=self.dev)

    # ------------------------------------------------------------------ peer memory of the row-sharded fused iteration
    def _peer_setup(self, group, want: bool) -> bool:
        """Exchange buffers of the row-sharded persistent kernel (``_peer.PeerExchange``); every rank takes the same
        decision: the peer path is used only when every rank supports its shape and every mapping succeeded."""
        from ._peer import PeerExchange
        import torch.distributed as dist
        nbytes = self.lib.dagma_linear_iter_exchange_bytes(self.d, dist.get_world_size(group))
        self._peer = PeerExchange.create(group, nbytes, self.dev, want)
        return self._peer is not None

    def close(self):
        """Release the peer mappings (collective: every rank of the group calls it) and free the own buffer."""
        if self._peer is None:
            return
        pr, self._peer = self._peer, None
        self.one_kernel = False
        pr.close()

    # ------------------------------------------------------------------ trek regulariser
    def _trek_setup(self, plan):
        """PST penalty of any series / aggregation (src/notreks/notreks.py:454-619) through ``_pst.PstEngine``:
        closed-form adjoints instead of autograd, d pst / d W = 2 W o trek_GT^T -- for ``seq="inv"`` one more fused
        inverse and four DMMA GEMMs per iteration, GEMM chains for the other series."""
        self.trek = plan
        self.tcc = None
        if plan is None:
            return
        if plan["kind"] == "tcc":
            # spectral TCC as the reference's dispatch calls it (notreks.py:699-707); with notreks.FORWARD_TCC_CONFIG the
            # regulariser's own version / method; a log-det TCC regulariser has no fused form here
            from . import notreks
            from ._tcc import SpectralTcc
            kw = notreks._tcc_call_args(plan["reg"], notreks.FORWARD_TCC_CONFIG)
            if kw.pop("cycle_penalty", "spectral") != "spectral":
                raise NotImplementedError("a log-det TCC regulariser inside minimize: use trek_cycle_coupling_value_gradW")
            kw.pop("s", None)
            self.tcc = SpectralTcc(self.d, plan["I"], **kw)
            self.trek_GT = torch.empty(self.d, self.d, dtype=torch.float64, device=self.dev)
            return
        from ._pst import PstEngine
        self.pst = PstEngine(self.d, plan["I"], plan["seq"], plan["agg"], eps_inv=plan["eps_inv"], K_log=plan["K_log"])
        self.trek_GT = self.pst.GT

    def _trek_grad(self):
        """trek_GT (consumed by the update kernel) at the current W; no host synchronisation for PST and for TCC with
        the power method (graph-captured); the converged Perron iteration of the eig methods polls the host."""
        if self.tcc is not None:
            self.tcc.compute(self.W, out=self.trek_GT, fold_w=False)
            return
        self.pst.grad(self.W)

    def _trek_value(self) -> float:
        if self.trek is None:
            return 0.0
        if self.tcc is not None:
            self.tcc.compute(self.W, out=self.tcc.grad, fold_w=True)
            return float(self.tcc.value().item())
        return float(self.pst.value(self.W).item())

    def stale(self, model) -> bool:
        return model._cov_dev.data_ptr() != self._model_cov_ptr or model.d != self.d

    # ------------------------------------------------------------------ state block
    def _sptr(self, field: int) -> int:
        return self.state.data_ptr() + 8 * field

    def _iptr(self, slot: int) -> int:
        return self.state.data_ptr() + 8 * 17 + 4 * slot

    def _push(self, **kw):
        self.state_host.copy_(self.state)
        ints = self.state_host[17:].view(torch.int32)
        for k, val in kw.items():
            if k in ("it", "halted", "info"):
                ints[{"it": I_IT, "halted": I_HALTED, "info": I_INFO}[k]] = int(val)
            else:
                self.state_host[k] = float(val)
        self.state.copy_(self.state_host, non_blocking=True)

    def _pull(self):
        self.state_host.copy_(self.state)
        ints = self.state_host[17:].view(torch.int32)
        if int(ints[I_INFO]) == 99:
            raise _lib.DagmaB200Error("dagma_linear_iter_f64: a grid barrier timed out (the grid was not co-resident)")
        return self.state_host, int(ints[I_IT]), int(ints[I_HALTED]), int(ints[I_INFO])

    # ------------------------------------------------------------------ launch sequences
    def _inverse(self, s: float, W=None, want_grad=None):
        W = self.W if W is None else W
        _lib.check(self.lib.dagma_logdet_inv_ws_f64(
            _lib.stream_ptr(), self.d, float(s), W.data_ptr(), self.d, 1, self._sptr(F_LAD), self._sptr(F_H),
            self.Minv.data_ptr(), want_grad.data_ptr() if want_grad is not None else None, self.d,
            self._sptr(F_MIN), self._iptr(I_INFO), self.ws.data_ptr(), self.ws.numel() * 8), "dagma_logdet_inv_ws_f64")

    def _score_T(self, W=None, sigmoid=True):
        """l2: T = cov @ W.  logistic: T = X^T sigmoid(X W) (partial over the local rows, then all-reduced)."""
        W = self.W if W is None else W
        if self.loss_type == "l2":
            gemm(self.cov, W, self.T)
        else:
            gemm(self.X, W, self.R, epilogue=1 if sigmoid else 0)
            if sigmoid:
                gemm(self.X, self.R, self.T, trans_a=True, ws=self.kws)
                if self.group is not None:
                    torch.distributed.all_reduce(self.T, group=self.group)

    def _trek_opt(self) -> bool:
        return self.trek is not None and self.trek["mode"] == "opt"

    def _update(self):
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        opt = self._trek_opt()
        _lib.check(self.lib.dagma_linear_update_ex_f64(
            _lib.stream_ptr(), self.d, self.state.data_ptr(), self.W.data_ptr(), self.Minv.data_ptr(),
            self.T.data_ptr(), self.cov.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), ptr(self.mask_exc),
            ptr(self.mask_inc), self.trek_GT.data_ptr() if opt else None, self.trek["weight"] if opt else 0.0),
            "dagma_linear_update_ex_f64")

    def _inverse_and_score(self, s: float):
        """l2: the inverse of sI - W o W and T = cov @ W as one call (one persistent kernel for d > 256: the GEMM
        tiles fill the time the engines would spend waiting for the pivot chain of the inverse)."""
        _lib.check(self.lib.dagma_logdet_inv_gemm_ws_f64(
            _lib.stream_ptr(), self.d, float(s), self.W.data_ptr(), self.d, 1, self._sptr(F_LAD), self._sptr(F_H),
            self.Minv.data_ptr(), None, self.d, self._sptr(F_MIN), self._iptr(I_INFO), self.ws.data_ptr(),
            self.ws.numel() * 8, self.cov.data_ptr(), self.W.data_ptr(), self.T.data_ptr()),
            "dagma_logdet_inv_gemm_ws_f64")

    def _gradient_pieces(self, s: float):
        if self.loss_type == "l2" and self.d > 256:
            self._inverse_and_score(s)           # both saturate the GPU: back to back
        else:
            # the inverse of sI - W o W (a few CTAs at these sizes, a serial pivot chain) and the score products are
            # independent: fork the score onto a side stream, join before the update (captured as parallel graph
            # branches; the multi-GPU all-reduce of the partial gradient rides on the side stream as well)
            if self._side is None:
                self._side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._score_T()
            self._inverse(s)
            cur.wait_stream(self._side)
        if self._trek_opt():
            self._trek_grad()

    def _iteration(self, s: float):
        self._gradient_pieces(s)
        self._update()

    def _fused_iterations(self, n: int):
        """n inner iterations in one launch (stops early, like the sequence, once `halted` is latched)."""
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        logistic = int(self.loss_type == "logistic")
        if self._peer is not None:
            _lib.check(self.lib.dagma_linear_iter_sharded_f64(
                _lib.stream_ptr(), self.n, self.d, int(n), self.state.data_ptr(), self.W.data_ptr(), self.m.data_ptr(),
                self.v.data_ptr(), self.Minv.data_ptr(), self.T.data_ptr(), self.cov.data_ptr(), ptr(self.X),
                ptr(self.mask_exc), ptr(self.mask_inc), self.iter_ws.data_ptr(), self.iter_sync.data_ptr(),
                self._peer.rank, self._peer.world, self._peer.ptrs), "dagma_linear_iter_sharded_f64")
            return
        _lib.check(self.lib.dagma_linear_iter_f64(
            _lib.stream_ptr(), logistic, self.n if logistic else 0, self.d, int(n), self.state.data_ptr(),
            self.W.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.Minv.data_ptr(), self.T.data_ptr(),
            self.cov.data_ptr(), ptr(self.X), ptr(self.mask_exc), ptr(self.mask_inc), self.iter_ws.data_ptr(),
            self.iter_sync.data_ptr()), "dagma_linear_iter_f64")

    def _replay(self, s: float, n: int):
        if self.one_kernel:
            if n > 0:
                self._fused_iterations(n)
            return
        if self.group is not None and not self._graph_collectives:   # NCCL outside a graph: launch eagerly
            for _ in range(n):
                self._iteration(s)
            return
        if self.tcc is not None and self.tcc.method != "power" and self._trek_opt():
            for _ in range(n):                   # the converged Perron iteration polls the host: not capturable
                self._iteration(s)
            return
        if self._graph is None or self._graph_s != s:
            self._iteration(s)                   # warm-up on the current stream (also sets kernel attributes)
            torch.cuda.synchronize()
            # the warm-up advanced the state: it is re-initialised by the caller before counting
            self._graph_s = s
            self._restore_after_warmup()
            self._graph = capture_graph(lambda: self._iteration(s))
        for _ in range(n):
            self._graph.replay()

    PROBE = 64       # graph path: iterations between two looks at the `halted` latch

    def _replay_probed(self, s: float, n: int):
        """``_replay`` of n iterations.  The persistent kernel leaves its loop by itself once an inverse is infeasible;
        graph replays cannot, and every replay after the latch is a full inverse + score product that changes nothing
        (up to checkpoint - 1 of them).  So the graph path looks at the latch every ``PROBE`` iterations -- one 80-byte
        read-back per 64 iterations (at d = 2000: every 92 ms) -- and stops issuing replays once it is set."""
        if self.one_kernel or n <= self.PROBE:
            self._replay(s, n)
            return
        done = 0
        while done < n:
            k = min(self.PROBE, n - done)
            self._replay(s, k)
            done += k
            if done < n and self._pull()[2]:
                break

    def _snapshot(self):
        self._snap = (self.W.clone(), self.m.clone(), self.v.clone(), self.state.clone())

    def _restore_after_warmup(self):
        W, m, v, st = self._snap
        self.W.copy_(W)
        self.m.copy_(m)
        self.v.copy_(v)
        self.state.copy_(st)

    def _apply_dir(self, sign: float):
        _lib.check(self.lib.dagma_linear_apply_dir_f64(_lib.stream_ptr(), self.d, self.state.data_ptr(),
                                                       self.W.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                                       float(sign)), "dagma_linear_apply_dir_f64")

    # ------------------------------------------------------------------ objective (linear.py:118-135)
    def _objective(self, mu: float, s: float, lambda1: float):
        self._inverse(s)
        if self.loss_type == "l2":
            self._score_T()
            self._objective_sums(self.W, True)
            st, *_ = self._pull()
            score = float(st[F_SCORE])
        else:
            score = self._logistic_loss(self.W)
            self._objective_sums(self.W, False)
            st, *_ = self._pull()
        h, l1 = float(st[F_H]), float(st[F_L1])
        obj = mu * (score + lambda1 * l1) + h
        self.last_trek_val = self._trek_value()                  # linear.py:122-133
        if self._trek_opt():
            obj = obj + self.trek["weight"] * self.last_trek_val
        return obj, score, h

    def _objective_sums(self, W, l2: bool):
        """score_acc = 1/2 sum (I - W) o (cov - T) (l2 only) and l1_acc = sum |W| into the state block."""
        _lib.check(self.lib.dagma_linear_objective_ws_f64(
            _lib.stream_ptr(), self.d, self.state.data_ptr(), W.data_ptr(), self.T.data_ptr() if l2 else None,
            self.cov.data_ptr() if l2 else None, int(l2), self.obj_ws.data_ptr(), self.obj_ws.numel() * 8),
            "dagma_linear_objective_ws_f64")

    def _logistic_loss(self, W) -> float:
        gemm(self.X, W, self.R)
        _lib.check(self.lib.dagma_logistic_loss_f64(_lib.stream_ptr(), self.n, self.d, self.X.data_ptr(),
                                                    self.R.data_ptr(), 1.0 / self.n_total, self.partial.data_ptr(), 592,
                                                    self._sptr(F_LOSS)), "dagma_logistic_loss_f64")
        if self.group is not None:
            torch.distributed.all_reduce(self.state[F_LOSS:F_LOSS + 1], group=self.group)
        st, *_ = self._pull()
        return float(st[F_LOSS])

    # ------------------------------------------------------------------ public pieces
    def score(self, W: torch.Tensor):
        """``_score`` (linear.py:70-94): (loss, gradient) at W."""
        W = W.contiguous()
        G = self.cov.clone()
        if self.loss_type == "l2":
            gemm(self.cov, W, G, alpha=1.0, beta=-1.0)                 # cov W - cov = -cov (I - W)
            gemm(self.cov, W, self.T)
            self._objective_sums(W, True)
            st, *_ = self._pull()
            return float(st[F_SCORE]), G
        loss = self._logistic_loss(W)
        gemm(self.X, W, self.R, epilogue=1)
        if self.group is None:
            gemm(self.X, self.R, G, trans_a=True, alpha=1.0 / self.n_total, beta=-1.0, ws=self.kws)
        else:
            gemm(self.X, self.R, self.T, trans_a=True, ws=self.kws)
            torch.distributed.all_reduce(self.T, group=self.group)
            G = self.T / self.n_total - self.cov
        return loss, G

    def _verbose_iteration(self, s, mu, lambda1):
        """The checkpoint iteration with the reference's gradient diagnostics (linear.py:262-273): the same launch
        sequence as every other iteration, plus device-side norms of the pieces it already computed."""
        self._gradient_pieces(s)
        _, _, halted, info = self._pull()
        if halted or info:
            self._update()                       # latches `halted`; the caller's back-tracking path takes over
            return None
        W = self.W
        sgn = torch.sign(W)
        gscale = 1.0 if self.loss_type == "l2" else 1.0 / self.n_total
        G_score = mu * (gscale * self.T - self.cov)
        G_h = 2.0 * W * (self.Minv.T + 1e-16)
        G_l1 = (mu * lambda1) * sgn
        G_inc = (-2.0 * mu * lambda1) * sgn * self.mask_inc.to(torch.float64) if self.mask_inc is not None else None
        Gobj = G_score + G_l1 + G_h
        if G_inc is not None:
            Gobj = Gobj + G_inc
        G_trek = None
        if self._trek_opt():
            G_trek = self.trek["weight"] * 2.0 * W * self.trek_GT.T
            Gobj = Gobj + G_trek
        zero = torch.zeros((), dtype=torch.float64, device=self.dev)
        norms = torch.stack([torch.linalg.norm(Gobj), torch.linalg.norm(G_score), torch.linalg.norm(G_h),
                             torch.linalg.norm(G_l1), torch.linalg.norm(G_inc) if G_inc is not None else zero,
                             torch.linalg.norm(G_trek) if G_trek is not None else zero])
        self._update()
        st, it, _, _ = self._pull()
        c1 = 1.0 / ((1.0 - float(st[F_P1H])) - float(st[F_P1L]))
        c2 = 1.0 / ((1.0 - float(st[F_P2H])) - float(st[F_P2L]))
        step = torch.linalg.norm((self.m * c1) / (torch.sqrt(self.v * c2) + 1e-8))
        Wa = self.W.abs()
        nz = Wa[Wa > 0]
        wst = torch.stack([torch.linalg.norm(self.W), Wa.sum(), Wa.max(), nz.min() if nz.numel() else zero, step])
        vals = torch.cat([norms, wst]).cpu().tolist()
        keys = ("grad_raw_norm", "grad_score_norm", "grad_dag_norm", "grad_l1_norm", "grad_inc_norm", "grad_trek_norm",
                "w_norm", "w_abs_sum", "max_abs_w", "min_abs_w_nonzero", "grad_step_norm")
        return dict(zip(keys, vals))

    def minimize(self, W_np: np.ndarray, mu, max_iter, s, lr, tol, beta_1, beta_2, lambda1, checkpoint, log=None,
                 telemetry=None):
        """linear.py:165-333 on device; returns (status, iterations done).  ``W_np`` is updated in place.
        ``telemetry(diag, it, obj, score, h, trek_val, lr)`` is called at every checkpoint when given."""
        self.W.copy_(torch.from_numpy(np.ascontiguousarray(W_np)))
        self.m.zero_()
        self.v.zero_()
        self.state.zero_()
        gscale = 1.0 if self.loss_type == "l2" else 1.0 / self.n_total
        self.state_host.zero_()
        for f, val in ((F_MU, mu), (F_S, s), (F_LR, lr), (F_LAM, lambda1), (F_B1, beta_1), (F_B2, beta_2),
                       (F_P1H, 1.0), (F_P2H, 1.0), (F_GSCALE, gscale)):
            self.state_host[f] = float(val)
        self.state.copy_(self.state_host)
        self._snapshot()
        status, it_done, obj_prev = _lib.ST_OK, 0, 1e16
        max_iter = int(max_iter)
        while it_done < max_iter:
            chunk_end = min((it_done // checkpoint + 1) * checkpoint, max_iter)
            diag = None
            if telemetry is None:
                self._replay_probed(s, chunk_end - it_done)
            else:                                # the last iteration of the chunk also reports its gradient norms
                if chunk_end - it_done > 1:
                    self._replay(s, chunk_end - it_done - 1)
                _, _, halted0, _ = self._pull()
                if not halted0:
                    diag = self._verbose_iteration(s, mu, lambda1)
            st, it_dev, halted, info = self._pull()
            if halted:
                it_done = it_dev
                if it_done == 0 or s <= 0.9:                      # linear.py:231-233
                    status |= _lib.ST_OUT_OF_DOMAIN
                    break
                stop = False
                while True:                                      # linear.py:235-241
                    self._apply_dir(+1.0)
                    lr *= 0.5
                    self.state[F_LR:F_LR + 1].fill_(lr)
                    if lr <= 1e-16:
                        status |= _lib.ST_LR_UNDERFLOW
                        stop = True
                        break
                    self._apply_dir(-1.0)
                    self._inverse(s)
                    _, _, _, info = self._pull()
                    if info == 0:
                        break
                self.state[17:].view(torch.int32)[I_HALTED:I_INFO + 1].zero_()
                if stop:
                    break
                continue
            it_done = chunk_end
            obj, score, h = self._objective(mu, s, lambda1)      # linear.py:279-280
            if log is not None:
                log.append((0, it_done, obj, score, h, lr))
            if telemetry is not None and diag is not None:
                telemetry(diag, it_done, obj, score, h, self.last_trek_val, lr)
            if np.abs((obj_prev - obj) / obj_prev) <= tol:       # linear.py:328
                break
            obj_prev = obj
        W_np[...] = self.W.cpu().numpy()
        self.last_lr = lr
        return status, it_done
