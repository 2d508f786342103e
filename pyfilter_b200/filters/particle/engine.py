"""Python owner of one ``smcb_filter`` handle (C ABI, include/smcb200.h): device buffers, views, the time loop."""
import ctypes as C

import torch

from ... import _lib
from ...timeseries import StateSpaceModel, TimeseriesState


class Engine:
    def __init__(self, model: StateSpaceModel, proposal_id: int, algorithm_id: int, resampler_id: int, particles: int,
                 batch_shape: torch.Size, ess_threshold: float, seed: int, history_rows: int, fold_lookahead: bool = True,
                 exact_weights: bool = False, column_offset: int = 0, proposal_config: dict = None):
        _lib.require_cuda()
        self.lib = model.library() if hasattr(model, "library") else _lib.load_library()   # (a user model lives in its own build)
        self.model = model
        self.batch_shape = torch.Size(batch_shape)
        self.B = int(self.batch_shape[0]) if len(self.batch_shape) else 1
        self.N = int(particles)
        params = model.parameter_matrix(self.B)
        cfg = _lib.smcb_config()
        cfg.model, cfg.proposal, cfg.algorithm, cfg.resampler = model.model_id, proposal_id, algorithm_id, resampler_id
        cfg.particles, cfg.batch = self.N, self.B
        cfg.n_raw_params, cfg.param_cols = params.shape[0], params.shape[1]
        cfg.params_host = params.data_ptr() and C.cast(params.data_ptr(), C.POINTER(C.c_float))
        cfg.ess_threshold = float(ess_threshold)
        cfg.seed = int(seed) & (2**64 - 1)
        cfg.history_rows = int(history_rows)
        cfg.fold_lookahead = int(bool(fold_lookahead))
        cfg.exact_weights = int(bool(exact_weights))
        cfg.column_offset = int(column_offset)   # global index of column 0: the Philox counters use column_offset + column
        pc = proposal_config or {}
        cfg.lin_steps, cfg.lin_alpha = int(pc.get("n_steps", 1)), float(pc.get("alpha", 1e-4))
        cfg.lin_second_order = int(bool(pc.get("use_second_order", False)))
        cfg.nested_samples = int(pc.get("num_samples", 0))
        self._params_keepalive = params
        h = C.c_void_p()
        _lib.check(self.lib.smcb_filter_create(C.byref(cfg), C.byref(h)), self.lib)
        self.handle = h
        info = self.info()
        self.ld, self.D, self.OD, self.history_rows = info.ld, info.state_dim, info.obs_dim, info.history_rows
        self.event_shape = model.hidden.event_shape
        self.stamp = 0  # bumps on every mutation of the device state; live state objects carry the stamp they were made at
        self.t = 0
        self._keep = []

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.smcb_filter_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ---- raw access
    def info(self):
        out = _lib.smcb_info()
        _lib.check(self.lib.smcb_filter_info(self.handle, C.byref(out)))
        return out

    def _ptr(self, what: int) -> int:
        p = C.c_void_p()
        _lib.check(self.lib.smcb_filter_ptr(self.handle, what, C.byref(p)))
        return p.value

    def raw(self, what: int, shape, typestr="<f4"):
        return _lib.as_tensor(self._ptr(what), shape, typestr, self)

    def _squeeze_batch(self, t: torch.Tensor, batch_dim: int):
        return t.select(batch_dim, 0) if len(self.batch_shape) == 0 else t

    def x_view(self) -> torch.Tensor:
        """Particles in the reference's layout ``(N, [B], [d])`` - a strided view of the SoA device buffer ``(D, B, ld)``."""
        v = self.raw(_lib.PTR_X, (self.D, self.B, self.ld))[:, :, : self.N].permute(2, 1, 0)
        if len(self.event_shape) == 0:
            v = v[..., 0]
        return self._squeeze_batch(v, 1)

    def logw_view(self) -> torch.Tensor:
        return self._squeeze_batch(self.raw(_lib.PTR_LOGW, (self.B, self.ld))[:, : self.N].t(), 1)

    def prev_inds(self) -> torch.Tensor:
        v = self.raw(_lib.PTR_PREV_INDS, (self.B, self.ld), "<i4")[:, : self.N].t().long()
        return self._squeeze_batch(v, 1)

    def small(self, what: int, with_dim: bool) -> torch.Tensor:
        shape = (self.B, self.D) if with_dim else (self.B,)
        return self._squeeze_batch(self.raw(what, shape).clone(), 0)

    def history(self, rows: int):
        m = self.raw(_lib.PTR_HIST_MEAN, (self.history_rows, self.B, self.D))[:rows].clone()
        v = self.raw(_lib.PTR_HIST_VAR, (self.history_rows, self.B, self.D))[:rows].clone()
        ll = self.raw(_lib.PTR_HIST_LL, (self.history_rows, self.B))[:rows].clone()
        if len(self.batch_shape) == 0:
            m, v, ll = m[:, 0], v[:, 0], ll[:, 0]
        return m, v, ll

    # ---- operations
    def initialize(self):
        _lib.check(self.lib.smcb_filter_initialize(self.handle, _lib.current_stream()))
        self.t = 0
        self.stamp += 1

    def load_state(self, x: torch.Tensor, w: torch.Tensor, prev_inds: torch.Tensor, t: int):
        """``init_state=`` / an external state for ``filter``: copy it into the device buffers and refresh the statistics."""
        self.t = int(t)
        # the handle picks the live ping-pong buffer from its move counter: set it before fetching the pointer
        _lib.check(self.lib.smcb_filter_refresh_state(self.handle, self.t, _lib.current_stream()))
        self.x_view().copy_(x.to("cuda", torch.float32))
        self.logw_view().copy_(w.to("cuda", torch.float32))
        pi = self.raw(_lib.PTR_PREV_INDS, (self.B, self.ld), "<i4")[:, : self.N].t()
        self._squeeze_batch(pi, 1).copy_(prev_inds.to("cuda", torch.int32))
        _lib.check(self.lib.smcb_filter_refresh_state(self.handle, self.t, _lib.current_stream()))
        self.stamp += 1

    def set_observations(self, y_dev: torch.Tensor, base_t: int):
        assert y_dev.is_cuda and y_dev.dtype == torch.float32 and y_dev.is_contiguous()
        self._keep = [y_dev]
        count = y_dev.shape[0]
        _lib.check(self.lib.smcb_filter_set_observations(self.handle, y_dev.data_ptr(), count, base_t, _lib.current_stream()))

    def run(self, steps: int):
        _lib.check(self.lib.smcb_filter_run(self.handle, steps, _lib.current_stream()))
        self.t += steps
        self.stamp += 1

    def run_stepwise(self, steps: int):
        """``steps`` launches of one move each (the online pattern of SMC2 / NESS), the loop itself in C."""
        _lib.check(self.lib.smcb_filter_run_stepwise(self.handle, steps, _lib.current_stream()))
        self.t += steps
        self.stamp += 1

    def set_noise(self, eps=None, u=None, U=None):
        self._noise_keep = (eps, u, U)
        _lib.check(self.lib.smcb_filter_set_noise(self.handle, eps.data_ptr() if eps is not None else None,
                                                  u.data_ptr() if u is not None else None, U.data_ptr() if U is not None else None))

    def set_nested_noise(self, z=None, e=None):
        """NestedProposal parity hook: ``z`` (num_samples, N, [B], [d]) standard normals of the inner samples, ``e`` (N, [B], num_samples)
        the Exp(1) values of ``torch.multinomial``'s single-sample draw."""
        zs = es = None
        if z is not None:
            m = z.shape[0]
            zs = torch.zeros((m, self.D, self.B, self.ld), device="cuda", dtype=torch.float32)
            zs[..., : self.N] = z.to("cuda", torch.float32).reshape(m, self.N, self.B, self.D).permute(0, 3, 2, 1)
        if e is not None:
            m = e.shape[-1]
            es = torch.ones((m, self.B, self.ld), device="cuda", dtype=torch.float32)
            es[..., : self.N] = e.to("cuda", torch.float32).reshape(self.N, self.B, m).permute(2, 1, 0)
        self._nested_keep = (zs, es)
        _lib.check(self.lib.smcb_filter_set_nested_noise(self.handle, zs.data_ptr() if zs is not None else None,
                                                         es.data_ptr() if es is not None else None), self.lib)

    def dump_noise(self, eps=None, u=None, w=None):
        self._dump_keep = (eps, u, w)
        _lib.check(self.lib.smcb_filter_dump_noise(self.handle, eps.data_ptr() if eps is not None else None,
                                                   u.data_ptr() if u is not None else None, w.data_ptr() if w is not None else None))

    def ess(self):
        _lib.check(self.lib.smcb_filter_sync_stats(self.handle, _lib.current_stream()))
        packed = self.raw(_lib.PTR_ESS, (2, self.B)).clone()
        return self._squeeze_batch(packed[0], 0), self._squeeze_batch(packed[1], 0)

    # ---- layouts: the reference's (N, [B], [d]) tensors <-> the device's SoA rows (D, B, ld)
    def to_soa(self, x: torch.Tensor) -> torch.Tensor:
        buf = torch.zeros((self.D, self.B, self.ld), device="cuda", dtype=torch.float32)
        v = x.to(device="cuda", dtype=torch.float32).reshape(self.N, self.B, self.D)
        buf[:, :, : self.N] = v.permute(2, 1, 0)
        return buf

    def from_soa(self, buf: torch.Tensor) -> torch.Tensor:
        v = buf[:, :, : self.N].permute(2, 1, 0).contiguous()   # (N, B, D)
        return v.reshape((self.N,) + tuple(self.batch_shape) + tuple(self.event_shape))

    def from_rows(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[:, : self.N].t().contiguous().reshape((self.N,) + tuple(self.batch_shape))

    def _y_dev(self, y) -> torch.Tensor:
        return torch.as_tensor(y, dtype=torch.float32).reshape(-1).to("cuda").contiguous()

    # ---- the proposal plug-in as stand-alone passes (proposals/base.py:52-85)
    def pre_weight(self, y, x: torch.Tensor = None) -> torch.Tensor:
        yd = self._y_dev(y)
        xs = self.to_soa(x) if x is not None else None
        out = torch.empty((self.B, self.ld), device="cuda", dtype=torch.float32)
        _lib.check(self.lib.smcb_filter_pre_weight(self.handle, yd.data_ptr(), xs.data_ptr() if xs is not None else None, out.data_ptr(),
                                                   _lib.current_stream()))
        return self.from_rows(out)

    def sample_and_weight(self, y, x: torch.Tensor = None, eps: torch.Tensor = None, t: int = None):
        yd = self._y_dev(y)
        xs = self.to_soa(x) if x is not None else None
        es = self.to_soa(eps) if eps is not None else None
        xo = torch.zeros((self.D, self.B, self.ld), device="cuda", dtype=torch.float32)
        wo = torch.empty((self.B, self.ld), device="cuda", dtype=torch.float32)
        _lib.check(self.lib.smcb_filter_sample_and_weight(self.handle, yd.data_ptr(), xs.data_ptr() if xs is not None else None,
                                                          es.data_ptr() if es is not None else None, int(self.t if t is None else t),
                                                          xo.data_ptr(), wo.data_ptr(), _lib.current_stream()))
        return self.from_soa(xo), self.from_rows(wo)

    def predict_path(self, steps: int, x: torch.Tensor = None):
        """``model.sample_states(steps, x_0)`` for every particle: ``(steps, N, [B], [d])`` states and ``(steps, N, [B], [obs d])`` observations."""
        xs = self.to_soa(x) if x is not None else None
        xo = torch.empty((steps, self.D, self.B, self.ld), device="cuda", dtype=torch.float32)
        yo = torch.empty((steps, self.OD, self.B, self.ld), device="cuda", dtype=torch.float32)
        _lib.check(self.lib.smcb_filter_predict_path(self.handle, int(steps), xs.data_ptr() if xs is not None else None, xo.data_ptr(),
                                                     yo.data_ptr(), _lib.current_stream()))
        xp = xo[..., : self.N].permute(0, 3, 2, 1).contiguous().reshape((steps, self.N) + tuple(self.batch_shape) + tuple(self.event_shape))
        obs_event = tuple(self.model.event_shape)
        yp = yo[..., : self.N].permute(0, 3, 2, 1).contiguous().reshape((steps, self.N) + tuple(self.batch_shape) + obs_event)
        return xp, yp

    def set_params(self, model: StateSpaceModel):
        """New parameter values for the same compiled model (SMC2 / PMMH rebuild the model per theta: filters/base.py:75-83)."""
        params = model.parameter_matrix(self.B)
        self._params_keepalive = params
        _lib.check(self.lib.smcb_filter_set_params(self.handle, C.cast(params.data_ptr(), C.POINTER(C.c_float)), params.shape[0], params.shape[1],
                                                   _lib.current_stream()))
        self.model = model
        self.stamp += 1

    def set_seed(self, seed: int):
        _lib.check(self.lib.smcb_filter_set_seed(self.handle, int(seed) & (2**64 - 1)))

    def set_ess_threshold(self, relative: float):
        _lib.check(self.lib.smcb_filter_set_ess_threshold(self.handle, float(relative)))

    # ---- theta-level operations on the resident state (SMC2 / PMMH): FilterResult.resample / exchange without leaving the device
    def resample_columns(self, indices: torch.Tensor, entire_history: bool = True):
        idx = indices.to(device="cuda", dtype=torch.int64).contiguous()
        _lib.check(self.lib.smcb_filter_resample_columns(self.handle, idx.data_ptr(), int(bool(entire_history)), _lib.current_stream()))
        self.stamp += 1

    def exchange_columns(self, other: "Engine", mask: torch.Tensor):
        m = mask.to(device="cuda", dtype=torch.uint8).contiguous()
        _lib.check(self.lib.smcb_filter_exchange_columns(self.handle, other.handle, m.data_ptr(), _lib.current_stream()))
        self.stamp += 1

    # ---- columns as records (cross-rank theta-resampling of a sharded batch)
    def export_columns(self) -> torch.Tensor:
        n = int(self.lib.smcb_filter_column_record_elems(self.handle))
        buf = torch.empty((self.B, n), device="cuda", dtype=torch.int32)
        _lib.check(self.lib.smcb_filter_export_columns(self.handle, buf.data_ptr(), _lib.current_stream()), self.lib)
        return buf

    def import_columns(self, records: torch.Tensor, indices: torch.Tensor = None, folded: int = -1):
        assert records.is_cuda and records.dtype == torch.int32 and records.is_contiguous()
        idx = indices.to(device="cuda", dtype=torch.int64).contiguous() if indices is not None else None
        _lib.check(self.lib.smcb_filter_import_columns(self.handle, records.data_ptr(), int(records.shape[0]),
                                                       idx.data_ptr() if idx is not None else None, int(folded), _lib.current_stream()), self.lib)
        self.stamp += 1

    def make_state(self):
        from .state import ParticleFilterCorrection

        x = TimeseriesState(torch.tensor(self.t), self.x_view(), self.event_shape)
        # x / log-weights are zero-copy views of the LIVE device buffers (valid until the engine moves again: `is_live`); a state that is
        # kept (FilterResult with record_states, a state handed back to filter() later) is detached at append time (filters/result.py)
        return ParticleFilterCorrection(x, self.logw_view(), self.small(_lib.PTR_LL, False), self.prev_inds,
                                        self.small(_lib.PTR_MEAN, True), self.small(_lib.PTR_VAR, True), engine=self,
                                        stamp=self.stamp)
