"""Device engine: owns one libb200gp context per GPU, the torch-allocated workspace and the staging of
programs / hyperparameters.  PyTorch is used for device memory, streams and copies only."""
import ctypes

import numpy as np

from . import _lib
from .program import ProgramBatch

_ENGINES = {}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.B200GPError('no CUDA device: ms_project_b200 has no CPU fallback')
    return torch


class Engine(object):
    """One per GPU.  `bind` the (standardised) training data once per search, then score batches."""

    def __init__(self, device=0, own_stream=False):
        torch = _torch()
        self.torch = torch
        self.device = torch.device('cuda', device)
        self.lib = _lib.lib()
        torch.cuda.set_device(self.device)
        # own_stream: a private non-blocking stream, so that two engines driven from two host threads overlap on the GPU
        # (fit.fit_models_overlapped); otherwise everything is enqueued on torch's current stream
        self.own_stream = bool(own_stream)
        self.stream = torch.cuda.Stream(self.device) if own_stream else torch.cuda.current_stream(self.device)
        self._peer = None
        ctx = ctypes.c_void_p()
        rc = self.lib.b200gp_create(device, ctypes.c_void_p(self.stream.cuda_stream), ctypes.byref(ctx))
        if rc != 0:
            raise _lib.B200GPError('b200gp_create failed (rc=%d): needs an sm_100 GPU' % rc)
        self.ctx = ctx
        self.N = self.D = 0
        self.slots = 0
        self.max_params = 0
        self._ws = None
        self._X = self._y = None
        self._X_host = self._y_host = None

    def on_stream(self):
        """Context manager: torch work issued inside runs on this engine's stream."""
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.own_stream else contextlib.nullcontext()

    def peer(self):
        """A second engine on the same GPU with its own stream and workspace (created on first use)."""
        if self._peer is None:
            self._peer = Engine(self.device.index, own_stream=True)
        return self._peer

    def close(self):
        if getattr(self, 'ctx', None):
            self.lib.b200gp_destroy(self.ctx)
            self.ctx = None

    # ------------------------------------------------------------------ configuration
    def set_quirks(self, mask):
        _lib.check(self.ctx, self.lib.b200gp_set_quirks(self.ctx, int(mask)), 'b200gp_set_quirks')

    def set_profiling(self, on):
        _lib.check(self.ctx, self.lib.b200gp_set_profiling(self.ctx, int(bool(on))), 'b200gp_set_profiling')

    def set_panel_tiles(self, tiles):
        _lib.check(self.ctx, self.lib.b200gp_set_panel_tiles(self.ctx, int(tiles)), 'b200gp_set_panel_tiles')

    def set_tile64(self, on):
        """CTA tile shape of the DMMA phases: 64 x 64 (default) or 128 x 64; bit-identical results."""
        _lib.check(self.ctx, self.lib.b200gp_set_tile64(self.ctx, int(bool(on))), 'b200gp_set_tile64')

    def set_potrf_la(self, on):
        """Diagonal tiles: 1 look-ahead register factorisation (default), 2 the same with a split barrier (bit-identical),
        0 the blocked one (same results up to rounding)."""
        _lib.check(self.ctx, self.lib.b200gp_set_potrf_la(self.ctx, int(on)), 'b200gp_set_potrf_la')

    def set_tma(self, mode):
        """Operand staging of the 64 x 64 tile GEMMs: 3 TMA + 32-byte-atom swizzle (default), 2 TMA + 128-byte swizzle,
        1 TMA unswizzled, 0 cp.async; bit-identical results."""
        _lib.check(self.ctx, self.lib.b200gp_set_tma(self.ctx, int(mode)), 'b200gp_set_tma')

    def set_fused_grad(self, on):
        """K^-1 fused with the gradient pass of the <= 4-leaf programs (one kernel, K^-1 never written to HBM)."""
        _lib.check(self.ctx, self.lib.b200gp_set_fused_grad(self.ctx, int(bool(on))), 'b200gp_set_fused_grad')

    def launch_count(self):
        return int(self.lib.b200gp_launch_count(self.ctx))

    def last_phase_ms(self):
        buf = (ctypes.c_float * 8)()
        n = self.lib.b200gp_last_phase_ms(self.ctx, buf, 8)
        return [float(buf[i]) for i in range(max(n, 0))]

    def slots_for(self, N, D, n_items, max_params, mem_fraction=0.8, extra_bytes=0):
        """Largest batch the free HBM holds (180 GB / GPU on B200), capped at the number of items.
        `extra_bytes` counts memory that will be released before the new workspace is allocated."""
        free, _total = self.torch.cuda.mem_get_info(self.device)
        free += int(extra_bytes)
        per_slot = self.lib.b200gp_workspace_bytes(N, D, 2, max_params) - self.lib.b200gp_workspace_bytes(N, D, 1, max_params)
        fixed = self.lib.b200gp_workspace_bytes(N, D, 1, max_params) - per_slot
        budget = int(free * mem_fraction) - fixed
        return int(max(1, min(n_items, budget // max(per_slot, 1))))

    def _bind_impl(self, X, y, slots, max_params=32):
        """X: N x D, y: N or N x 1 (host numpy, already standardised by the caller)."""
        torch = self.torch
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(-1)
        if X.ndim != 2 or X.shape[0] != y.shape[0]:
            raise ValueError('X must be N x D and y must have N entries')
        N, D = X.shape
        nbytes = self.lib.b200gp_workspace_bytes(N, D, int(slots), int(max_params))
        if nbytes == 0:
            raise ValueError('invalid workspace geometry')
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._X = torch.from_numpy(X).to(self.device)
        self._y = torch.from_numpy(y).to(self.device)
        self._X_host, self._y_host = X, y
        rc = self.lib.b200gp_bind(self.ctx, self._ws.data_ptr(), self._ws.numel(), int(slots), int(max_params),
                                  self._X.data_ptr(), N, D, self._y.data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_bind')
        self.N, self.D, self.slots, self.max_params = N, D, int(slots), int(max_params)

    def ensure_bound(self, X, y, n_items, max_params):
        """Bind (X, y) unless the engine already holds the same data with enough slots / parameter room.
        Slots are sized for `n_items` but capped by the free HBM."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        yv = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).reshape(-1)
        max_params = max(int(max_params), 1)
        same = (self._X_host is not None and self._X_host.shape == X.shape and self._y_host.shape == yv.shape
                and np.array_equal(self._X_host, X) and np.array_equal(self._y_host, yv))
        want = min(int(n_items), self.slots_for(X.shape[0], X.shape[1], int(n_items), max(max_params, self.max_params),
                                                extra_bytes=0 if self._ws is None else self._ws.numel()))
        if same and self.slots >= want and self.max_params >= max_params:
            return
        self.bind(X, yv, max(want, 1), max(max_params, self.max_params if same else 1))

    # ------------------------------------------------------------------ program staging
    def _upload_programs_impl(self, batch):
        """ProgramBatch -> device tensors (kept on the batch object)."""
        torch = self.torch
        if batch.max_params > self.max_params:
            raise ValueError('batch has kernels with %d parameters; engine bound for %d' % (batch.max_params, self.max_params))
        if batch.max_dim >= self.D:
            raise ValueError('a kernel uses input dimension %d but X has %d columns' % (batch.max_dim, self.D))
        if getattr(batch, 'lookup', None) is not None:      # OP_LOOKUP leaves: (re-)upload the builder's distance table
            builder, n_models = batch.lookup
            table = np.ascontiguousarray(np.asarray(builder.get_kernel(n_models), dtype=np.float64))
            self._lut = torch.from_numpy(table).to(self.device)
            _lib.check(self.ctx, self.lib.b200gp_set_lookup(self.ctx, self._lut.data_ptr(), int(table.shape[0])), 'b200gp_set_lookup')
        batch.d_tokens = torch.from_numpy(batch.tokens.reshape(-1)).to(self.device)
        batch.d_prog_off = torch.from_numpy(batch.prog_off).to(self.device)
        batch.d_theta_off = torch.from_numpy(batch.theta_off).to(self.device)
        n, nt = batch.n_items, int(batch.theta_off[-1])
        batch.d_theta = torch.empty(nt, dtype=torch.float64, device=self.device)
        batch.d_lml = torch.empty(n, dtype=torch.float64, device=self.device)
        batch.d_dlml = torch.empty(nt, dtype=torch.float64, device=self.device)
        batch.d_status = torch.empty(n, dtype=torch.int32, device=self.device)
        batch.d_tries = torch.empty(n, dtype=torch.int32, device=self.device)
        batch.h_theta = torch.empty(nt, dtype=torch.float64).pin_memory()
        batch.h_out = torch.empty(n + nt, dtype=torch.float64).pin_memory()
        batch.h_int = torch.empty(2 * n, dtype=torch.int32).pin_memory()
        batch.d_out = torch.empty(n + nt, dtype=torch.float64, device=self.device)
        batch.d_int = torch.empty(2 * n, dtype=torch.int32, device=self.device)
        batch.d_all_ids = torch.arange(n, dtype=torch.int32, device=self.device)
        batch.engine = self
        return batch

    # ------------------------------------------------------------------ scoring
    def _lml_grad_impl(self, batch, theta, ids=None, with_grad=True):
        """Evaluate items `ids` (default: all) of an uploaded batch at the flat host `theta`
        (device layout, see ProgramBatch.pack_theta).  Host buffers in, host results out:
        returns (lml[n_items], dlml[theta layout], status[n_items], tries[n_items]); entries of items not
        in `ids` are left untouched from the previous call."""
        torch = self.torch
        n = batch.n_items
        batch.h_theta.numpy()[:] = theta
        batch.d_theta.copy_(batch.h_theta, non_blocking=True)
        if ids is None:
            d_ids, n_act = batch.d_all_ids, n
        else:
            ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
            n_act = int(ids.size)
            d_ids = torch.from_numpy(ids).to(self.device, non_blocking=True)
        if n_act == 0:
            return (batch.h_out.numpy()[:n], batch.h_out.numpy()[n:], batch.h_int.numpy()[:n], batch.h_int.numpy()[n:])
        d_lml, d_dlml = batch.d_out[:n], batch.d_out[n:]
        d_status, d_tries = batch.d_int[:n], batch.d_int[n:]
        rc = self.lib.b200gp_lml_grad(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(),
                                      batch.d_theta.data_ptr(), batch.d_theta_off.data_ptr(), d_ids.data_ptr(), n_act,
                                      int(bool(with_grad)), d_lml.data_ptr(), d_dlml.data_ptr(), d_status.data_ptr(),
                                      d_tries.data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_lml_grad')
        batch.h_out.copy_(batch.d_out, non_blocking=True)
        batch.h_int.copy_(batch.d_int, non_blocking=True)
        self.stream.synchronize()
        out = batch.h_out.numpy()
        ints = batch.h_int.numpy()
        return out[:n], out[n:], ints[:n], ints[n:]

    def _lml_grad_device_impl(self, batch, ids=None, with_grad=True):
        """Same evaluation with theta already in `batch.d_theta`; results stay on the device
        (batch.d_out / batch.d_int).  Used by bench.py for the HBM-resident timing."""
        n = batch.n_items
        if ids is None:
            d_ids, n_act = batch.d_all_ids, n
        else:
            d_ids, n_act = ids, int(ids.numel())
        rc = self.lib.b200gp_lml_grad(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(),
                                      batch.d_theta.data_ptr(), batch.d_theta_off.data_ptr(), d_ids.data_ptr(), n_act,
                                      int(bool(with_grad)), batch.d_out[:n].data_ptr(), batch.d_out[n:].data_ptr(),
                                      batch.d_int[:n].data_ptr(), batch.d_int[n:].data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_lml_grad')

    def _posterior_impl(self, batch, theta, item, Xs, include_likelihood=True):
        """GP.predict of item `item` of an uploaded batch at the flat host `theta`: (mean[Ns], var[Ns], lml, status)."""
        torch = self.torch
        Xs = np.ascontiguousarray(np.asarray(Xs, dtype=np.float64))
        if Xs.ndim != 2 or Xs.shape[1] != self.D:
            raise ValueError('Xnew must be Ns x %d' % self.D)
        ns = Xs.shape[0]
        batch.h_theta.numpy()[:] = theta
        batch.d_theta.copy_(batch.h_theta, non_blocking=True)
        d_Xs = torch.from_numpy(Xs).to(self.device)
        d_out = torch.empty(2 * ns, dtype=torch.float64, device=self.device)
        rc = self.lib.b200gp_posterior(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                                       batch.d_theta_off.data_ptr(), batch.d_all_ids[item:].data_ptr(), int(item), d_Xs.data_ptr(),
                                       ns, int(bool(include_likelihood)), d_out[:ns].data_ptr(), d_out[ns:].data_ptr(),
                                       batch.d_lml.data_ptr(), batch.d_status.data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_posterior')
        out = d_out.cpu().numpy()
        return out[:ns].copy(), out[ns:].copy(), float(batch.d_lml[item].item()), int(batch.d_status[item].item())

    def _build_K_impl(self, batch, theta, ids=None, add_diag=False):
        """Dense kernel matrices (n x N x N, host numpy) of the selected items over the bound X."""
        torch = self.torch
        batch.h_theta.numpy()[:] = theta
        batch.d_theta.copy_(batch.h_theta, non_blocking=True)
        if ids is None:
            d_ids, n_act = batch.d_all_ids, batch.n_items
        else:
            ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
            n_act = int(ids.size)
            d_ids = torch.from_numpy(ids).to(self.device)
        out = torch.empty((n_act, self.N, self.N), dtype=torch.float64, device=self.device)
        rc = self.lib.b200gp_build_K(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(),
                                     batch.d_theta.data_ptr(), batch.d_theta_off.data_ptr(), d_ids.data_ptr(), n_act,
                                     int(bool(add_diag)), out.data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_build_K')
        return out.cpu().numpy()

    def _gram_logdet_impl(self, batch, theta, diag_mode, K_out=None, logdet=None, status=None):
        """Gram matrices (device tensor K_out [n_items, N, N]) and / or their log-determinants (device tensors) of every
        item of an uploaded batch at the flat host `theta`; diag_mode as in b200gp_gram_logdet."""
        batch.h_theta.numpy()[:] = theta
        batch.d_theta.copy_(batch.h_theta, non_blocking=True)
        rc = self.lib.b200gp_gram_logdet(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(), batch.d_theta.data_ptr(),
                                         batch.d_theta_off.data_ptr(), batch.d_all_ids.data_ptr(), batch.n_items, int(diag_mode),
                                         K_out.data_ptr() if K_out is not None else None,
                                         logdet.data_ptr() if logdet is not None else None,
                                         status.data_ptr() if status is not None else None)
        _lib.check(self.ctx, rc, 'b200gp_gram_logdet')

    def _centered_alignments_impl(self, batch, theta):
        """Centred alignment of every pair of an uploaded batch's kernels over the bound X: n x n, diagonal 1."""
        torch = self.torch
        n = batch.n_items
        batch.h_theta.numpy()[:] = theta
        batch.d_theta.copy_(batch.h_theta, non_blocking=True)
        out = torch.empty((n, n), dtype=torch.float64, device=self.device)
        rc = self.lib.b200gp_centered_alignments(self.ctx, batch.d_tokens.data_ptr(), batch.d_prog_off.data_ptr(),
                                                 batch.d_theta.data_ptr(), batch.d_theta_off.data_ptr(),
                                                 batch.d_all_ids.data_ptr(), n, out.data_ptr())
        _lib.check(self.ctx, rc, 'b200gp_centered_alignments')
        return out.cpu().numpy()

    def _potrf_impl(self, A):
        """In-place blocked Cholesky of a device tensor A (N x N float64, N % 128 == 0).
        Returns (logdet, status)."""
        torch = self.torch
        N = A.shape[0]
        wb = self.lib.b200gp_potrf_work_bytes(N)
        if wb == 0:
            raise ValueError('N must be a positive multiple of 128')
        work = torch.empty(wb, dtype=torch.uint8, device=self.device)
        ld = ctypes.c_double()
        st = ctypes.c_int32()
        rc = self.lib.b200gp_potrf(self.ctx, A.data_ptr(), N, work.data_ptr(), ctypes.byref(ld), ctypes.byref(st))
        _lib.check(self.ctx, rc, 'b200gp_potrf')
        return ld.value, st.value



    # ------------------------------------------------------------------ stream-scoped public entry points
    def bind(self, *args, **kwargs):
        with self.on_stream():
            return self._bind_impl(*args, **kwargs)

    def upload_programs(self, *args, **kwargs):
        with self.on_stream():
            return self._upload_programs_impl(*args, **kwargs)

    def lml_grad(self, *args, **kwargs):
        with self.on_stream():
            return self._lml_grad_impl(*args, **kwargs)

    def lml_grad_device(self, *args, **kwargs):
        with self.on_stream():
            return self._lml_grad_device_impl(*args, **kwargs)

    def posterior(self, *args, **kwargs):
        with self.on_stream():
            return self._posterior_impl(*args, **kwargs)

    def build_K(self, *args, **kwargs):
        with self.on_stream():
            return self._build_K_impl(*args, **kwargs)

    def gram_logdet(self, *args, **kwargs):
        with self.on_stream():
            return self._gram_logdet_impl(*args, **kwargs)

    def centered_alignments(self, *args, **kwargs):
        with self.on_stream():
            return self._centered_alignments_impl(*args, **kwargs)

    def potrf(self, *args, **kwargs):
        with self.on_stream():
            return self._potrf_impl(*args, **kwargs)


def get_engine(device=0, role='search'):
    """Engines are cached per (device, role); 'adhoc' serves one-off kern.K calls so that a search's
    bound data is never disturbed."""
    key = (device, role)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device)
    return _ENGINES[key]


def kernel_matrix(kernel, X, X2=None):
    """kern.K(X, X2) on the device (compute_kernel, src/autoks/backend/kernel.py:317-322)."""
    X = np.asarray(X, dtype=np.float64)
    if X2 is not None and (X2 is not X) and not (X2.shape == X.shape and np.array_equal(X2, X)):
        raise NotImplementedError('cross-covariances K(X, X2 != X) go through the posterior path (SURVEY.md 8f, f4)')
    eng = get_engine(0, 'adhoc')
    batch = ProgramBatch([kernel])
    eng.bind(X, np.zeros(X.shape[0]), slots=1, max_params=max(batch.max_params, 1))
    eng.upload_programs(batch)
    return eng.build_K(batch, batch.initial_theta(), add_diag=False)[0]


def kernel_diag(kernel, X):
    return np.diag(kernel_matrix(kernel, X)).copy()
