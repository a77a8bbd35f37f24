"""ADMMBasedOptimizer -- host side of the B200-native ADMM-LSTM sweep.

Mirrors the reference's class (reference: admm.py:22-78 / admm.no_dual_y.py:12-66): same constructor
signature, `step()`, and public attributes (model, train_x, train_y, batch_size, seq_len,
input_size, output_size, hidden_size, verbose, betas, rhos, gates, duals, summary), so that
demo.py:317-321,355 and comparison_experiment/comparison.py:167-170 drive it unchanged.

What differs is where the work happens: every N*T-sized operation is a hand-written sm_100a kernel
behind the C ABI of include/admm_lstm_b200.h (loaded through ctypes in _lib.py); this file only
owns buffers (torch is the device allocator / stream provider), sequences the launches in the
reference's update order and all-reduces the few small cross-sample sums when the samples are
sharded over several GPUs.  The theta decisions of the reference's `while` loops are replayed on
the device, so a step contains no host synchronisation.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib
from .comm import Comm
from .logging_utils import error, info, log_assert, warning
from .parameters import example_parameter_dictionary

_GATES = ("i", "f", "g", "o")
_STATE_KEYS = ("i", "f", "g", "o", "c", "h")
_VARIANTS = {"admm": _lib.VARIANT_ADMM, "no_dual_y": _lib.VARIANT_NO_DUAL_Y}


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.AdmmLibraryError(
            "admm_lstm_b200 needs a CUDA device (B200, sm_100a): there is no CPU path. "
            "For CPU experiments use the reference implementation.")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _feature_major(x: torch.Tensor, ldn: int, device) -> torch.Tensor:
    """[N, A, B] -> [A, B, ldn] (or [N, B] -> [B, ldn]) zero-padded along the sample axis."""
    x = x.to(device=device, dtype=torch.float32)
    if x.dim() == 3:
        out = torch.zeros((x.size(1), x.size(2), ldn), dtype=torch.float32, device=device)
        out[:, :, : x.size(0)] = x.permute(1, 2, 0)
    else:
        out = torch.zeros((x.size(1), ldn), dtype=torch.float32, device=device)
        out[:, : x.size(0)] = x.t()
    return out


class _StateDict:
    """Read-only mapping that presents the feature-major device buffers in the reference's shapes
    ([N, T+1, H] for i,f,g,o,c,h and their duals, [N, O] for a / y) as zero-copy views.

    Ownership: the reference's getters hand out private copies (admm.py:191-213, `.clone().detach()` on every read -- 18 244
    clones per step).  Here `opt.gates[k]` ALIASES the optimizer's state: it changes with the next step(), and writing through
    it changes the iteration (call opt.state_changed() afterwards).  `copy()` gives the reference's semantics."""

    def copy(self):
        """{key: private clone} -- what the reference's accessors return."""
        return {k: g().clone() for k, g in self._getters.items()}

    def __init__(self, getters):
        self._getters = getters

    def __getitem__(self, key):
        return self._getters[key]()

    def keys(self):
        return self._getters.keys()

    def items(self):
        return [(k, g()) for k, g in self._getters.items()]

    def __contains__(self, key):
        return key in self._getters

    def __iter__(self):
        return iter(self._getters)


class ADMMBasedOptimizer(object):
    """ADMM optimizer for the LSTM-Linear model of blocks/lstm.py.

    One `step()` = Wy, then (x2g, h2g) for g in i,f,g,o, then for t = 1..T the primal updates of
    i,f,g,o,c,h (a at t = T) followed by the dual ascent at t -- the reference's order
    (admm.py:62-78).

    Extra keyword arguments (all optional, the reference call sites do not pass them):
      variant      'admm' (admm.py as shipped) or 'no_dual_y' (admm.no_dual_y.py, "Fast").
                   Default: module attribute `admm.variant` / env ADMM_LSTM_VARIANT / 'admm'.
      with_dual_y  admm.py:12 switch (default False, as shipped).
      sharding     under torch.distributed with world_size > 1: 'slice' (every rank was given the
                   full data, as `torchrun demo.py` would; each keeps its contiguous slice) or
                   'presharded' (each rank passes only its own samples).
      use_tensor_cores  None = automatic (tcgen05 path when the shape is eligible).
      use_cuda_graph  True: replay the step as one CUDA graph (launch-bound shapes; opt-in, see _graph_wanted).
      probe        how the backtracking loop of the weight updates (admm.py:331-338) gets f(w + G/theta): 'moments'
                   (default; one pass, every theta at once from a 6th-order expansion along the probe ray, valid where
                   the perturbation of the pre-activations is <= 2^-4 -- checked on the device, exact passes follow if
                   not) or 'exact' (every candidate evaluated one by one).  Both replay the same comparison; they can
                   stop at different thetas only where the reference's own fp32 comparison is decided by rounding
                   (the "absorption exits" of SURVEY section 7), with no effect on the iterates at the parity tolerance.
    """

    def __init__(self, model, training_samples: Tuple[torch.Tensor, torch.Tensor],
                 parameter_dictionary: Optional[Dict[str, Dict[str, float]]] = None, verbose: bool = True, *,
                 variant: Optional[str] = None, with_dual_y: bool = False, sharding: str = "slice",
                 use_tensor_cores: Optional[bool] = None, scratch_bytes: Optional[int] = None,
                 comm: Optional[Comm] = None, keep_preactivations: Optional[bool] = None,
                 probe: Optional[str] = None, use_cuda_graph: Optional[bool] = None) -> None:
        self._lib = _lib.load()
        self.device = _require_cuda()
        if self._lib.admm_device_ok() <= 0:
            raise _lib.AdmmLibraryError("no sm_100 (B200) device visible to libadmm_lstm_b200.so")
        variant = variant or os.environ.get("ADMM_LSTM_VARIANT", "admm")
        log_assert(variant in _VARIANTS, f"variant must be one of {list(_VARIANTS)} (Got: {variant}).")
        log_assert(not (with_dual_y and variant == "no_dual_y"), "with_dual_y needs variant='admm'.")
        self.variant, self.with_dual_y = variant, bool(with_dual_y)
        self.probe = probe or os.environ.get("ADMM_LSTM_PROBE", "moments")
        self.moment_order = int(os.environ.get("ADMM_LSTM_MOMENT_ORDER", "4"))      # 4 (fused pass) or 6; A/B switch
        log_assert(self.probe in ("moments", "exact"), f"probe must be 'moments' or 'exact' (Got: {self.probe}).")
        log_assert(sharding in ("slice", "presharded"), f"sharding must be 'slice' or 'presharded' (Got: {sharding}).")
        self.verbose = verbose
        self.summary = None
        self.comm = comm if comm is not None else Comm()
        self.kernel_events = None         # set by enable_kernel_timing()

        self.model = model.to(self.device)
        train_x, train_y = training_samples
        self.train_x, self.train_y = train_x, train_y
        (self.batch_size, self.seq_len, self.input_size, self.output_size,
         self.hidden_size) = self.__initialize_training_constants(train_x, train_y)

        # ---- sample sharding ------------------------------------------------------------------
        if self.comm.active and sharding == "slice":
            lo, hi = self.comm.shard_range(self.batch_size)
            local_x, local_y = train_x[lo:hi], train_y[lo:hi]
            self.n_global = self.batch_size
        else:
            local_x, local_y = train_x, train_y
            self.n_global = self.comm.sum_int(self.batch_size, self.device) if self.comm.active else self.batch_size
        self.n_local = int(local_x.size(0))
        log_assert(self.n_local > 0, "Every rank needs at least one training sample.")
        self.ldn = _round_up(self.n_local, 128)

        self.betas: Dict[str, torch.Tensor] = dict()
        self.rhos: Dict[str, torch.Tensor] = dict()
        hyper = self.__initialize_hyper_parameters(parameter_dictionary)

        N, T, D, H, O = self.n_local, self.seq_len, self.input_size, self.hidden_size, self.output_size
        log_assert(O <= _lib.ADMM_MAX_O, f"output_size must be <= {_lib.ADMM_MAX_O} (Got: {O}).")
        dev, f32 = self.device, torch.float32
        # ---- device buffers (feature-major) --------------------------------------
        self._x = _feature_major(local_x, self.ldn, dev)                  # [T][D][ldn]
        self._y = _feature_major(local_y, self.ldn, dev)                  # [O][ldn]
        self._state = {k: torch.zeros((T + 1, H, self.ldn), dtype=f32, device=dev) for k in _STATE_KEYS}
        self._dual = {k: torch.zeros((T + 1, H, self.ldn), dtype=f32, device=dev) for k in ("i", "f", "g", "o", "c")}
        self._dual_h = torch.zeros((H, self.ldn), dtype=f32, device=dev)
        self._a = torch.zeros((O, self.ldn), dtype=f32, device=dev)
        self._dual_y = torch.zeros((O, self.ldn), dtype=f32, device=dev)
        self._wx = torch.empty((4, D, H), dtype=f32, device=dev)
        self._wh = torch.empty((4, H, H), dtype=f32, device=dev)
        self._wy = torch.empty((H, O), dtype=f32, device=dev)
        self._pull_weights_from_model()
        self._bind_weights_to_model()

        # ---- small accumulators, packed so that each phase needs ONE all-reduce --------------------
        kmax = max(D, H)
        self._acc_wy = torch.zeros(H * O, dtype=torch.float64, device=dev)
        self._acc_grad = torch.zeros(4 * kmax * H + 4, dtype=torch.float64, device=dev)   # [G_acc | f(w)]
        self._acc_fk = torch.zeros(4 * _lib.ADMM_FK_SLOTS, dtype=torch.float64, device=dev)
        self._acc_est = torch.zeros(4 * _lib.ADMM_EST_CAND * 2, dtype=torch.float64, device=dev)
        self._acc_last = torch.zeros(1 + 3 * 4, dtype=torch.float64, device=dev)
        self._metrics = torch.zeros(_lib.ADMM_N_METRICS, dtype=torch.float64, device=dev)
        self._grad = torch.zeros(4 * kmax * H, dtype=f32, device=dev)
        self._done = torch.zeros(12, dtype=torch.int32, device=dev)      # [0,4) decided, [4,12) diagnostics
        self._theta_w = torch.zeros(8, dtype=f32, device=dev)       # [src][gate]
        self._qmax = torch.zeros(4, dtype=f32, device=dev)          # max |Q| per gate of the current probe pass
        self._qmax_w = torch.zeros(8, dtype=f32, device=dev)        # [src][gate]: the same, kept for the next plans
        self._theta_h = torch.zeros(1, dtype=f32, device=dev)
        # ring of pinned read-backs of the chosen thetas: the host runs ahead of the device (it is throttled only
        # by the launch queue), so the hint for step s usually comes from step s-2
        self._theta_ring = [[torch.zeros(17, dtype=f32).pin_memory(), None, -1] for _ in range(4)]
        self._step_index = 0
        self._hint = None            # per src: exit exponents of step s-2 (list of 4) or None
        self._hint_q = None          # per src: max |Q| per gate of step s-2 or None

        # scratch of the weight phase: R^T for the gradient pass, Z0 and Q for the probe pass -> 8*H*ldn floats
        # per timestep of a chunk
        budget = scratch_bytes if scratch_bytes is not None else int(os.environ.get("ADMM_LSTM_SCRATCH_BYTES", 4 << 30))
        self._tc_chunk = max(1, min(T, budget // (32 * H * self.ldn)))
        self._scratch = torch.empty(8 * H * self._tc_chunk * self.ldn, dtype=f32, device=dev)

        # ---- C problem descriptor --------------------------------------------------------------------
        p = _lib.Problem()
        p.n, p.n_global, p.ldn = N, self.n_global, self.ldn
        p.T, p.D, p.H, p.O = T, D, H, O
        p.variant, p.with_dual_y = _VARIANTS[variant], int(self.with_dual_y)
        p.hp = hyper
        p.x, p.y = self._x.data_ptr(), self._y.data_ptr()
        for q, k in enumerate(_STATE_KEYS):
            p.gate[q] = self._state[k].data_ptr()
        for q, k in enumerate(("i", "f", "g", "o", "c")):
            p.dual[q] = self._dual[k].data_ptr()
        p.dual_h, p.a, p.dual_y = self._dual_h.data_ptr(), self._a.data_ptr(), self._dual_y.data_ptr()
        p.wx, p.wh, p.wy = self._wx.data_ptr(), self._wh.data_ptr(), self._wy.data_ptr()
        p.tc_ws, p.tc_ws_bytes = None, 0
        self._p = p
        self._pp = C.byref(p)
        self._tc_ws = None
        ws_bytes = int(self._lib.admm_tc_workspace_bytes(self._pp))
        want_tc = (ws_bytes > 0) if use_tensor_cores is None else bool(use_tensor_cores)
        if want_tc and ws_bytes <= 0:
            raise _lib.AdmmLibraryError(f"tensor-core path requested but shape (D={D}, H={H}) is not eligible")
        if want_tc:
            self._tc_ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
            p.tc_ws, p.tc_ws_bytes = self._tc_ws.data_ptr(), ws_bytes
            self._call("admm_tc_refresh", self._pp, _lib.TC_WEIGHTS | _lib.TC_INPUTS, _stream_ptr())
        self.uses_tensor_cores = want_tc
        # pre-activation store: 4 state tensors' worth of memory buys 3 of the 7 full-size GEMM passes per step
        self._zstore = self._wx_prev = None
        z_bytes = 16 * H * T * self.ldn
        if keep_preactivations is None:
            free, _ = torch.cuda.mem_get_info(dev)
            keep_preactivations = want_tc and z_bytes < 0.6 * free
        if keep_preactivations and want_tc:
            self._zstore = torch.empty(4 * H * T * self.ldn, dtype=f32, device=dev)
            self._wx_prev = self._wx.clone()
            p.zstore, p.wx_prev = self._zstore.data_ptr(), self._wx_prev.data_ptr()
        self.keeps_preactivations = self._zstore is not None

        self.gates = _StateDict({**{k: (lambda k=k: self._state[k].permute(2, 0, 1)[:N]) for k in _STATE_KEYS},
                                 "a": lambda: self._a.t()[:N]})
        self.duals = _StateDict({**{k: (lambda k=k: self._dual[k].permute(2, 0, 1)[:N]) for k in ("i", "f", "g", "o", "c")},
                                 "h": self._materialize_dual_h, "y": lambda: self._dual_y.t()[:N]})
        self.last_metrics: Optional[Dict[str, float]] = None
        self.phase_events = None          # set by enable_phase_timing()
        env_graph = os.environ.get("ADMM_LSTM_GRAPH")
        self.use_cuda_graph = use_cuda_graph if use_cuda_graph is not None else (None if env_graph is None else env_graph != "0")
        self._copy_stream = self._x_stage = self._y_stage = self._stage_ready = self._stage_free = None   # prefetch_inputs()
        self._graphs: Dict[tuple, tuple] = {}
        self.graph_replays = 0
        self.graph_replayed_launches = 0          # kernels launched by graph replays (the library's counter sees only the capture)
        self.__initialize_primal_gates()
        self._set_z_valid(True)           # the forward pass stored z of the initial weights and states

    # ------------------------------------------------------------------------------------ setup
    def __initialize_training_constants(self, train_x, train_y) -> Tuple[int, int, int, int, int]:
        """Shape checks of admm.py:92-108 (failure = log line + exit(1), like the reference)."""
        log_assert(train_x.dim() == 3 and train_y.dim() == 2,
                   f"train_x must be [N, T, D] and train_y [N, O] (Got: {list(train_x.shape)}, {list(train_y.shape)}).")
        train_batch, train_seq, train_feat = train_x.size()
        train_label_batch, train_label_feat = train_y.size()
        log_assert(train_batch == train_label_batch,
                   f"Batch size of samples mismatch (Got train_x: {train_batch}, train_y: {train_label_batch}).")
        log_assert(train_feat == self.model.input_size and train_label_feat == self.model.output_size,
                   f"Input and output size of samples must match that of the model "
                   f"(Got train_x: {train_feat}, train_y: {train_label_feat}, "
                   f"model: {self.model.input_size} -> {self.model.output_size}).")
        return train_batch, train_seq, train_feat, train_label_feat, self.model.hidden_size

    def __initialize_hyper_parameters(self, param_dict) -> _lib.Hyper:
        """admm.py:110-162: beta keys wy, w{g} (x2g), v{g} (h2g); rho keys i,f,g,o,c,h,y."""
        if not param_dict:
            param_dict = example_parameter_dictionary["GoogleStock"]
            warning(f"Parameter dictionary is empty, a default one will be applied: {param_dict}")
        if "beta" not in param_dict:
            error("Normalization factors missing in parameter dictionary.")
        log_assert("rho" in param_dict, "Penalties missing in parameter dictionary.")
        beta_dict, rho_dict = param_dict["beta"], param_dict["rho"]
        hyper = _lib.Hyper()

        def beta_of(key):
            if key not in beta_dict:
                error(f"Key {key} missing in normalization factors.")
            value = beta_dict[key]
            log_assert(isinstance(value, (float, int)), f"Beta {key} must be a float or integer.")
            log_assert(value >= 0, f"Beta {key} must be non-negative.")
            return float(value)

        hyper.beta_wy = beta_of("wy")
        self.betas["wy"] = torch.tensor(hyper.beta_wy, dtype=torch.float32, device=self.device)
        for q, gate in enumerate(_GATES):
            hyper.beta_x[q], hyper.beta_h[q] = beta_of("w" + gate), beta_of("v" + gate)
            self.betas["x2" + gate] = torch.tensor(hyper.beta_x[q], dtype=torch.float32, device=self.device)
            self.betas["h2" + gate] = torch.tensor(hyper.beta_h[q], dtype=torch.float32, device=self.device)
        for q, key in enumerate(("i", "f", "g", "o", "c", "h", "y")):
            log_assert(key in rho_dict, "Penalties missing in parameter dictionary.")
            log_assert(isinstance(rho_dict[key], (float, int)), "Penalties must be a float or integer.")
            hyper.rho[q] = float(rho_dict[key])
            self.rhos[key] = torch.tensor(hyper.rho[q], dtype=torch.float32, device=self.device)
        if self.verbose:
            self.summary = f"Parameters: {{'beta': {dict(beta_dict)}, 'rho': {dict(rho_dict)}}}"
        return hyper

    def __initialize_primal_gates(self) -> None:
        """admm.py:164-167 / blocks/lstm.py:65-88: forward pass fills i,f,g,o,c,h and a."""
        st = _stream_ptr()
        for t in range(1, self.seq_len + 1):
            self._call("admm_forward_t", self._pp, t, st)

    # ------------------------------------------------------------------------------------ weights
    def _set_z_valid(self, valid: bool) -> None:
        """zstore holds z = x W + h U of the current inputs, states and weights for every t (admm_problem::z_valid):
        true after the forward initialisation and after every complete sweep, false once anything else wrote them."""
        self._p.z_valid = int(bool(valid) and self._zstore is not None)

    def _param_names(self):
        return [f"{s}2{g}" for g in _GATES for s in ("x", "h")] + ["out"]

    def _pull_weights_from_model(self) -> None:
        with torch.no_grad():
            for q, g in enumerate(_GATES):
                self._wx[q].copy_(getattr(self.model, "x2" + g).detach())
                self._wh[q].copy_(getattr(self.model, "h2" + g).detach())
            self._wy.copy_(self.model.out.detach())

    def _bind_weights_to_model(self) -> None:
        """The model's parameters become views of the optimizer's weight buffers, so `model(x)` and
        `torch.save(model)` (demo.py:307,341) see every update without a copy.  The reference
        re-creates the Parameters on each set_weight (blocks/lstm.py:35,41); identity is not part of
        the contract, values are."""
        for q, g in enumerate(_GATES):
            setattr(self.model, "x2" + g, nn.Parameter(self._wx[q], requires_grad=False))
            setattr(self.model, "h2" + g, nn.Parameter(self._wh[q], requires_grad=False))
        setattr(self.model, "out", nn.Parameter(self._wy, requires_grad=False))
        self._bound_ptrs = {name: getattr(self.model, name).data_ptr() for name in self._param_names()}
        self._note_weight_versions()

    def _note_weight_versions(self) -> None:
        """torch's in-place version counters of the bound Parameters: the library's kernels write through raw pointers and
        never bump them, every torch in-place write (load_state_dict's param.copy_, p.data.mul_(), ...) does."""
        self._bound_versions = {name: getattr(self.model, name)._version for name in self._param_names()}

    def _resync_weights_if_replaced(self) -> None:
        """Weights written behind the optimizer's back since the last step: a Parameter object replaced (set_weight,
        blocks/lstm.py:35; load_state_dict on a new module) is imported; an in-place write into a bound Parameter
        (model.load_state_dict(sd), p.data.mul_()) already changed the fp32 buffers, but the tensor-core operand copies
        of the weights and the stored pre-activations are derived data and must be refreshed / dropped."""
        for name in self._param_names():
            if getattr(self.model, name).data_ptr() != self._bound_ptrs[name]:
                self._pull_weights_from_model()
                self._bind_weights_to_model()
                self._set_z_valid(False)
                if self._tc_ws is not None:
                    self._call("admm_tc_refresh", self._pp, _lib.TC_WEIGHTS, _stream_ptr())
                return
        if any(getattr(self.model, name)._version != self._bound_versions[name] for name in self._param_names()):
            self._set_z_valid(False)
            if self._tc_ws is not None:
                self._call("admm_tc_refresh", self._pp, _lib.TC_WEIGHTS, _stream_ptr())
            self._note_weight_versions()

    # ------------------------------------------------------------------------------------ plumbing
    def _call(self, name, *args) -> None:
        if self.kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(self._lib, name)(*args)
            e1.record()
            self.kernel_events.append((name, e0, e1))
        else:
            rc = getattr(self._lib, name)(*args)
        if rc != 0:
            _lib.check(rc, name)

    def enable_kernel_timing(self, enabled: bool = True) -> None:
        """Bracket every C-ABI call with CUDA events on the launching stream (bench.py roofline)."""
        self.kernel_events = [] if enabled else None

    def kernel_time_summary(self) -> Dict[str, Tuple[int, float]]:
        """{entry point: (calls, total ms)} of the recorded events; synchronises."""
        torch.cuda.synchronize()
        out: Dict[str, Tuple[int, float]] = {}
        for name, e0, e1 in self.kernel_events or []:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1))
        return out

    def prefetch_inputs(self, train_x: torch.Tensor, train_y: torch.Tensor) -> None:
        """Start the host->device copy of the NEXT inputs (this rank's samples, pinned host memory) on a side stream into a
        staging buffer, so that it overlaps the step in flight; a following `refresh_inputs()` without arguments installs
        them.  Double buffering of the inputs: the samples the running step reads are not touched."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._x_stage = torch.empty((self.n_local,) + tuple(train_x.shape[1:]), dtype=torch.float32, device=self.device)
            self._y_stage = torch.empty((self.n_local,) + tuple(train_y.shape[1:]), dtype=torch.float32, device=self.device)
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free is not None:
                self._copy_stream.wait_event(self._stage_free)          # the previous install has read the staging buffers
            self._x_stage.copy_(train_x, non_blocking=True)
            self._y_stage.copy_(train_y, non_blocking=True)
            self._stage_ready = torch.cuda.Event()
            self._stage_ready.record(self._copy_stream)

    def refresh_inputs(self, train_x: Optional[torch.Tensor] = None, train_y: Optional[torch.Tensor] = None) -> None:
        """Install new inputs for this rank's samples in the device layout: from (pinned) host memory -- the host->device leg
        of the end-to-end measurement in bench.py -- or, without arguments, from the staging buffers a `prefetch_inputs`
        call filled in the background."""
        n = self.n_local
        if train_x is None:
            log_assert(self._stage_ready is not None, "refresh_inputs() without arguments needs a prefetch_inputs() before it.")
            torch.cuda.current_stream().wait_event(self._stage_ready)
            xd, yd = self._x_stage, self._y_stage
        else:
            xd = train_x.to(self.device, non_blocking=True)
            yd = train_y.to(self.device, non_blocking=True)
        xd = xd.to(torch.float32).contiguous()
        yd = yd.to(torch.float32).contiguous()
        # one transposing kernel per tensor into the device layout; on the tensor-core path it also notes ON THE DEVICE whether
        # x changed at all -- if not, the stored pre-activations stay valid (admm_load_inputs), so z_valid is left alone
        self._call("admm_load_inputs", self._pp, xd.data_ptr(), yd.data_ptr(), _stream_ptr())
        if train_x is None:
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
            self._stage_ready = None
        if self._tc_ws is None:
            self._set_z_valid(False)

    def state_changed(self) -> None:
        """Call after writing the state or weight buffers directly (tests do): re-derives the tensor-core side
        buffers and drops the stored pre-activations."""
        self._set_z_valid(False)
        if self._tc_ws is not None:
            self._call("admm_tc_refresh", self._pp, _lib.TC_WEIGHTS | _lib.TC_STATE, _stream_ptr())

    def export_weights(self, out: Dict[str, torch.Tensor]) -> None:
        """Device->host read of the step's result (the nine weight tensors) into pinned buffers."""
        out["wx"].copy_(self._wx, non_blocking=True)
        out["wh"].copy_(self._wh, non_blocking=True)
        out["wy"].copy_(self._wy, non_blocking=True)
        out["metrics"].copy_(self._metrics, non_blocking=True)

    def _materialize_dual_h(self) -> torch.Tensor:
        """lambda_h in the reference's shape: zero except the t = T slot (admm.py:532-534)."""
        out = torch.zeros((self.n_local, self.seq_len + 1, self.hidden_size), dtype=torch.float32, device=self.device)
        out[:, self.seq_len, :] = self._dual_h.t()[: self.n_local]
        return out

    def enable_phase_timing(self, enabled: bool = True) -> None:
        self.phase_events = [] if enabled else None

    def _mark(self, label: str) -> None:
        if self.phase_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.phase_events.append((label, ev))

    def _time_chunks(self):
        T, tc = self.seq_len, self._tc_chunk
        return [(t0, min(tc, T - t0)) for t0 in range(0, T, tc)]

    def _poll_theta_hint(self) -> None:
        """Fetch the thetas chosen in step s-2: they size the first probe pass of step s.

        The schedule must be identical on every rank (the per-pass candidate sums are all-reduced), so it may
        only depend on replicated data, never on timing: the thetas are replicated, and "step s-2" is a fixed
        choice.  Waiting for that read-back costs nothing in steady state: the host is throttled by the launch
        queue to less than one step ahead of the device."""
        want = self._step_index - 2
        slot = self._theta_ring[want % len(self._theta_ring)] if want >= 0 else None
        if slot is None or slot[2] != want:
            self._hint = self._hint_q = None
            return
        slot[1].synchronize()
        th = slot[0][:8].view(2, 4)
        # theta_out = 2^(k-1) for exit index k (0.5 -> k = 0)
        self._hint = {src: [int(v).bit_length() if v >= 1.0 else 0 for v in th[src].tolist()]
                      for src in (_lib.SRC_X, _lib.SRC_H)}
        self._hint_q = {src: slot[0][9 + 4 * src: 13 + 4 * src].tolist() for src in (_lib.SRC_X, _lib.SRC_H)}

    def _probe_plans(self, src: int):
        """First pass: the moment pass (admm_probe_plan::moments) -- one activation per element gives f for every
        candidate k >= k0 at once; k0 = first exponent at which the expansion is valid with a factor 2 to spare
        (max|Q| 2^-k0 <= 2^-5 against the limit 2^-4, from the max|Q| of step s-2), everything below k0 represented by lower-bound sums.  On the
        benchmark workloads k0 is 0..2.  Then, in case the expansion turns out not to be valid or a bound fails, exact
        32-candidate passes from 0 (launched speculatively, they exit at once when all four gates are decided)."""
        full = [((0, 0, 0, 0), _lib.ADMM_MAX_CAND, 0, 0), ((32, 32, 32, 32), _lib.ADMM_MAX_CAND, 0, 0)]
        if self.probe == "exact":
            # every candidate evaluated one by one: a window of 8 around the exit of step s-2, lower-bound proofs on
            # 1/8 (5/8 next to the window) of the units below it.  Window = [k* - 5, k* + 2]: the bound proves `f > est`
            # only with a margin of 8x, which the quadratic growth of f along the step gives ~4 doublings below the exit.
            if self._hint is None:
                return full
            ks = self._hint[src]
            k0 = [max(0, min(k - 5, _lib.ADMM_MAX_CAND)) for k in ks]
            span = max(k + 3 - a for k, a in zip(ks, k0))
            ncand = min(_lib.ADMM_MAX_CAND, max(8, -(-span // 8) * 8))
            return [(tuple(k0), ncand, int(any(k0)), 0)] + full
        # fused moment pass (tensor-core path with a z store): 4th-order expansion, valid two exponents later (|Q| 2^-k <= 2^-6)
        # but a quarter fewer instructions per element in a kernel that is bound by them; otherwise 6th order (<= 2^-4)
        order = 4 if (self.uses_tensor_cores and self._zstore is not None and self.moment_order != 6) else 6
        lim = 2.0 ** -7 if order == 4 else 2.0 ** -5                  # a factor 2 inside the validity limit
        k0 = [0, 0, 0, 0]
        if self._hint_q is not None:
            for g, q in enumerate(self._hint_q[src]):
                if q == q and q > lim:                            # NaN-safe
                    if not math.isfinite(q) or q >= 2.0 ** (_lib.ADMM_MAX_CAND - 9):
                        return full          # a diverging run: the expansion's window would not fit, go straight to the exact passes
                    k0[g] = int(math.ceil(math.log2(q / lim)))     # <= ADMM_MAX_CAND - 2, so the proofs (k0 + 2) fit
        if self.use_cuda_graph:
            # the plan is baked into a captured graph: one common, even k0 for the four gates changes rarely (a larger k0 is
            # always valid, it only moves candidates from the expansion to the lower-bound proofs)
            k0 = [min(_lib.ADMM_MAX_CAND - 2, (max(k0) + 1) // 2 * 2)] * 4
        # proofs always cover two exponents more than the hint asks for (ncand = 2): max|Q| may grow by 8x between the
        # hint and this step before the exact passes are needed
        return [(tuple(k0), 2, 1, 1, order)] + full

    def _push_theta_hint(self) -> None:
        slot = self._theta_ring[self._step_index % len(self._theta_ring)]
        if slot[1] is not None:
            slot[1].synchronize()       # four steps old: long finished
        slot[0][:8].copy_(self._theta_w, non_blocking=True)
        slot[0][8:9].copy_(self._theta_h, non_blocking=True)
        slot[0][9:].copy_(self._qmax_w, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        slot[2] = self._step_index

    # ------------------------------------------------------------------------------------ step
    def step(self) -> None:
        """One ADMM iteration (admm.py:62-78)."""
        self._resync_weights_if_replaced()
        self._poll_theta_hint()
        if self._graph_wanted():
            self._step_graphed()
        else:
            self._step_body(_stream_ptr())
        self._set_z_valid(True)           # the sweep's GEMMs left z of the new weights / states in zstore
        self._push_theta_hint()
        self._step_index += 1

    def _step_body(self, st) -> None:
        """Everything a step enqueues on the stream: Wy, the weights, the sweep, the t = T tail.  No host synchronisation,
        no allocation -- which is what makes it capturable as a CUDA graph."""
        pp = self._pp
        self._mark("begin")
        self.__update_wy(st)
        self._mark("wy")
        for src in (_lib.SRC_X, _lib.SRC_H):
            self.__update_weights(src, st)
            self._mark("weights_x" if src == _lib.SRC_X else "weights_h")
        self._metrics.zero_()
        for t in range(1, self.seq_len + 1):
            self._call("admm_sweep_t", pp, t, self._metrics.data_ptr(), st)
        self._mark("sweep")
        self.__update_last(st)
        self._mark("last")

    # ------------------------------------------------------------------------------------ CUDA-graph replay (small shapes)
    def _graph_wanted(self) -> bool:
        """Launch-bound shapes (GoogleStock: ~100 launches of a few microseconds per step, admm.py:62-78 run as a Python
        loop in the reference too) can replay the step as ONE CUDA graph: `use_cuda_graph=True` / ADMM_LSTM_GRAPH=1, after
        three eager steps.  Not with sample sharding (the all-reduces stay eager) and not while per-call timing is recording."""
        if self.comm.active or self.kernel_events is not None or self.phase_events is not None or self._step_index < 3:
            return False
        # Opt-in.  Measured on one B200 (profiles/r02_f_graph_vs_eager.txt): GoogleStock 0.47 ms per replayed step against
        # 0.57 ms eager (the eager loop is bound by the host's ~45 launches, the replay by ~80 dependent nodes of a few
        # microseconds each) -- but every change of the probe plan costs a re-capture (~15 ms), so over 50 steps eager wins.
        return bool(self.use_cuda_graph)

    def _step_graphed(self) -> None:
        # what the captured launches bake in besides device pointers: the z_valid flag and the probe plans (by value)
        key = (int(self._p.z_valid), tuple(self._probe_plans(_lib.SRC_X)), tuple(self._probe_plans(_lib.SRC_H)))
        entry = self._graphs.get(key)
        if entry is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            graph = torch.cuda.CUDAGraph()
            before = int(self._lib.admm_launch_count(0))
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                self._step_body(_stream_ptr())
            entry = (graph, int(self._lib.admm_launch_count(0)) - before)
            self._graphs[key] = entry
        entry[0].replay()
        self.graph_replays += 1
        self.graph_replayed_launches += entry[1]

    def __update_wy(self, st) -> None:
        """admm.py:246-280 / admm.no_dual_y.py:226-249."""
        self._acc_wy.zero_()
        self._call("admm_wy_grad", self._pp, self._acc_wy.data_ptr(), st)
        self.comm.allreduce_sum_(self._acc_wy)
        self._call("admm_wy_apply", self._pp, self._acc_wy.data_ptr(), st)

    def __update_weights(self, src: int, st) -> None:
        """admm.py:282-343 for the four gates of one `src` at once.  The reference visits
        (x2i, h2i, x2f, ...); the gates do not interact in the weight phase, so (x2*, then h2*) gives
        the same iterates: h2g still sees the new x2g (admm.py:300)."""
        pp = self._pp
        K = self.input_size if src == _lib.SRC_X else self.hidden_size
        H = self.hidden_size
        n_g = 4 * K * H
        acc = self._acc_grad[: n_g + 4]
        acc.zero_()
        g_ptr, fw_ptr = acc.data_ptr(), acc[n_g:].data_ptr()
        chunks = self._time_chunks()
        self._call("admm_weight_begin", pp, src, st)
        for t0, tc in chunks:
            self._call("admm_weight_grad", pp, src, t0, tc, self._scratch.data_ptr(), g_ptr, fw_ptr, st)
        self.comm.allreduce_sum_(acc)
        self._call("admm_weight_finish_grad", pp, src, g_ptr, self._grad.data_ptr(), self._acc_est.data_ptr(), st)
        self._done.zero_()
        theta_ptr = self._theta_w[4 * src:].data_ptr()
        plans = self._probe_plans(src)
        for q, entry in enumerate(plans):
            k0, ncand, proof, moments = entry[:4]
            plan = _lib.ProbePlan()
            for g in range(4):
                plan.k0[g] = k0[g]
            plan.ncand, plan.proof, plan.moments = ncand, proof, moments
            plan.order = entry[4] if len(entry) > 4 else 0
            self._acc_fk.zero_()
            if moments:
                self._qmax.zero_()
            for t0, tc in chunks:
                self._call("admm_weight_probe", pp, src, t0, tc, self._scratch.data_ptr(), self._grad.data_ptr(),
                           C.byref(plan), self._done.data_ptr(), self._acc_fk.data_ptr(), self._qmax.data_ptr(), st)
            self.comm.allreduce_sum_(self._acc_fk)
            if moments:
                self.comm.allreduce_max_(self._qmax)
                self._qmax_w[4 * src: 4 * src + 4].copy_(self._qmax)
            self._call("admm_weight_select", pp, src, self._acc_est.data_ptr(), self._acc_fk.data_ptr(), self._qmax.data_ptr(),
                       C.byref(plan), int(q == len(plans) - 1), self._done.data_ptr(), theta_ptr, st)
        self._call("admm_weight_apply", pp, src, self._grad.data_ptr(), theta_ptr, st)

    def __update_last(self, st) -> None:
        """t = T tail: h_T backtracking, a, lambda_h (, lambda_y)  -- admm.py:459-502, 532-546."""
        self._acc_last.zero_()
        self._call("admm_last_probe", self._pp, self._acc_last.data_ptr(), st)
        self.comm.allreduce_sum_(self._acc_last)
        self._call("admm_last_select", self._pp, self._acc_last.data_ptr(), self._theta_h.data_ptr(), st)
        self._call("admm_last_apply", self._pp, self._theta_h.data_ptr(), self._metrics.data_ptr(), st)

    # ------------------------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> Dict[str, object]:
        """This rank's ADMM state (absent upstream): gates i,f,g,o,c,h and a, duals, in the reference's
        shapes on the CPU, plus what a resumed run needs to continue bit-identically (hyper-parameters, variant, the
        theta history that sizes the probe windows).  The weights are NOT included: they live in the model, whose file
        format (torch.save(model), demo.py:302-308) is unchanged."""
        out: Dict[str, object] = {
            "format": 1, "variant": self.variant, "with_dual_y": self.with_dual_y,
            "shape": (self.n_local, self.seq_len, self.input_size, self.hidden_size, self.output_size),
            "n_global": self.n_global, "rank": self.comm.rank, "world_size": self.comm.world_size,
            "rho": {k: float(v) for k, v in self.rhos.items()}, "beta": {k: float(v) for k, v in self.betas.items()},
            "step_index": self._step_index,
            "theta_w": self._theta_w.cpu(), "theta_h": self._theta_h.cpu(), "qmax_w": self._qmax_w.cpu(),
            "gates": {k: self.gates[k].cpu().contiguous() for k in _STATE_KEYS + ("a",)},
            "duals": {k: self.duals[k].cpu().contiguous() for k in ("i", "f", "g", "o", "c", "y")},
            "dual_h_T": self._dual_h.t()[: self.n_local].cpu().contiguous(),
        }
        return out

    def load_state_dict(self, sd: Dict[str, object]) -> None:
        log_assert(sd.get("format") == 1, "Unknown optimizer state format.")
        log_assert(tuple(sd["shape"]) == (self.n_local, self.seq_len, self.input_size, self.hidden_size, self.output_size),
                   f"Optimizer state shape mismatch (Got: {tuple(sd['shape'])}).")
        log_assert(sd["variant"] == self.variant and bool(sd["with_dual_y"]) == self.with_dual_y,
                   "Optimizer state was written by another variant.")
        n, dev = self.n_local, self.device
        for k in _STATE_KEYS:
            self._state[k][:, :, :n] = sd["gates"][k].to(dev).permute(1, 2, 0)
        for k in ("i", "f", "g", "o", "c"):
            self._dual[k][:, :, :n] = sd["duals"][k].to(dev).permute(1, 2, 0)
        self._dual_h[:, :n] = sd["dual_h_T"].to(dev).t()
        self._a[:, :n] = sd["gates"]["a"].to(dev).t()
        self._dual_y[:, :n] = sd["duals"]["y"].to(dev).t()
        self._theta_w.copy_(sd["theta_w"])
        self._theta_h.copy_(sd["theta_h"])
        self._qmax_w.copy_(sd.get("qmax_w", torch.zeros(8)))
        # the probe-window hint of the next two steps comes from the saved thetas (the ring is empty after a restart)
        self._step_index = int(sd["step_index"])
        for slot in self._theta_ring:
            slot[1], slot[2] = None, -1
        for back in (1, 2):
            idx = self._step_index - back
            if idx >= 0:
                slot = self._theta_ring[idx % len(self._theta_ring)]
                slot[0][:8].copy_(sd["theta_w"])
                slot[0][8:9].copy_(sd["theta_h"])
                slot[0][9:].copy_(sd.get("qmax_w", torch.zeros(8)))
                slot[1] = torch.cuda.Event()
                slot[1].record()
                slot[2] = idx
        self._resync_weights_if_replaced()
        self.state_changed()

    def save_state(self, path: str) -> str:
        """Write this rank's shard next to the model file: `<path>.rank<r>of<w>.pt`."""
        fn = f"{path}.rank{self.comm.rank}of{self.comm.world_size}.pt"
        torch.save(self.state_dict(), fn)
        return fn

    def load_state(self, path: str) -> None:
        fn = f"{path}.rank{self.comm.rank}of{self.comm.world_size}.pt"
        self.load_state_dict(torch.load(fn, map_location="cpu", weights_only=True))   # tensors, str, numbers, tuples only

    # ------------------------------------------------------------------------------------ observables
    def metrics(self) -> Dict[str, float]:
        """Objective / primal / dual residuals of the step just taken (the reference
        computes none of these).  Synchronises."""
        m = self._metrics.clone()
        self.comm.allreduce_sum_(m)
        reg = torch.zeros((), dtype=torch.float64, device=self.device)
        hp = self._p.hp
        for q in range(4):
            reg += 0.5 * hp.beta_x[q] * self._wx[q].double().pow(2).sum()
            reg += 0.5 * hp.beta_h[q] * self._wh[q].double().pow(2).sum()
        reg += (0.5 if self.variant == "admm" else 1.0) * hp.beta_wy * self._wy.double().pow(2).sum()
        m = m.cpu()
        loss = float(m[3]) / self.n_global
        out = {"objective": loss + float(reg) + float(m[2]), "primal_residual": float(m[0]) ** 0.5,
               "dual_residual": float(m[1]) ** 0.5, "loss_term": loss}
        if self.operand_overflow():
            warning("An |h| >= 32 did not fit the fp16-pair operand of the tensor-core path: the iterates since are invalid. "
                    "Re-create the optimizer with use_tensor_cores=False for this problem.")
        self.last_metrics = out
        return out

    def operand_overflow(self, reset: bool = False) -> bool:
        """Sticky device flag of the tensor-core path: some h_t (= (rho_h o tanh c - lambda_h)/rho_h, admm.py:455-457; o is an
        unconstrained ADMM primal) reached |h| >= 32 and was clamped in the fp16-pair GEMM operand.  Synchronises."""
        if self._tc_ws is None:
            return False
        rc = self._lib.admm_tc_overflow(self._pp, int(reset), _stream_ptr())
        if rc < 0:
            _lib.check(rc, "admm_tc_overflow")
        return rc == 1

    def training_loss(self) -> float:
        """MSE of the model on the WHOLE (global) training set -- what demo.py:341 computes every epoch with model(train_x) --
        from the samples already resident in the device layout: no re-upload, no transposition, one scalar all-reduce when
        the samples are sharded.  Synchronises."""
        lib, dev = self._lib, self.device
        out = torch.empty((self.output_size, self.ldn), dtype=torch.float32, device=dev)
        _predict_feature_major(lib, dev, self._x, self.n_local, self._wx, self._wh, self._wy, out)
        err = (out[:, : self.n_local] - self._y[:, : self.n_local]).double()
        acc = torch.stack([(err * err).sum(), torch.tensor(float(err.numel()), dtype=torch.float64, device=dev)])
        self.comm.allreduce_sum_(acc)
        return float(acc[0] / acc[1])

    def theta_trace(self) -> Dict[str, float]:
        """theta chosen by the last step for each weight and for h_T (diagnostics; synchronises)."""
        th = self._theta_w.cpu()
        out = {f"{'xh'[s]}2{g}": float(th[4 * s + q]) for s in (0, 1) for q, g in enumerate(_GATES)}
        out["h_T"] = float(self._theta_h.cpu()[0])
        return out


def _predict_feature_major(lib, dev, xt: torch.Tensor, n: int, wx, wh, wy, out: torch.Tensor) -> None:
    """out [O][ldn] <- the model's prediction for the feature-major batch xt [T][D][ldn] (admm_predict)."""
    T, D, ldn = xt.shape
    H, O = wh.shape[1], wy.shape[1]
    work = torch.empty(4 * H * ldn, dtype=torch.float32, device=dev)
    p = _lib.Problem()
    p.n, p.n_global, p.ldn = n, n, ldn
    p.T, p.D, p.H, p.O = T, D, H, O
    p.variant, p.with_dual_y = 0, 0
    p.x = xt.data_ptr()
    p.wx, p.wh, p.wy = wx.data_ptr(), wh.data_ptr(), wy.data_ptr()
    with torch.cuda.device(dev):
        _lib.check(lib.admm_predict(C.byref(p), work.data_ptr(), out.data_ptr(), _stream_ptr()), "admm_predict")


_PREDICT_CHUNK = 32768          # samples per admm_predict call: bounds the transposed copy of x and the 4*H*ldn work buffer


def predict_cuda(model, x: torch.Tensor) -> torch.Tensor:
    """model(x) for a CUDA batch through the library's forward kernel (reference: blocks/lstm.py:43-46,
    evaluated by demo.py:341-342 on the train and validation sets every epoch).  Chunked over samples, so the only
    temporaries are one chunk of x in the device layout and four [H, chunk] slabs -- the reference allocates six
    [N, T+1, H] tensors per call."""
    lib = _lib.load()
    dev = x.device
    N = x.shape[0]
    O = model.output_size
    wx = torch.stack([getattr(model, "x2" + g).detach() for g in _GATES]).contiguous().float()
    wh = torch.stack([getattr(model, "h2" + g).detach() for g in _GATES]).contiguous().float()
    wy = model.out.detach().contiguous().float()
    result = torch.empty((N, O), dtype=torch.float32, device=dev)
    for lo in range(0, N, _PREDICT_CHUNK):
        n = min(_PREDICT_CHUNK, N - lo)
        ldn = _round_up(n, 128)
        xt = _feature_major(x[lo:lo + n], ldn, dev)
        out = torch.empty((O, ldn), dtype=torch.float32, device=dev)
        _predict_feature_major(lib, dev, xt, n, wx, wh, wy, out)
        result[lo:lo + n] = out.t()[:n]
    return result


def sharded_mse_loss(model, x_local: torch.Tensor, y_local: torch.Tensor, comm: Optional[Comm] = None) -> float:
    """nn.MSELoss()(model(x), y) (demo.py:316,341-342) over samples that are sharded across ranks: every rank
    predicts its own samples with the library's forward kernel; ONE all-reduce of (sum of squared errors, element count)
    gives the global mean on every rank."""
    comm = comm if comm is not None else Comm()
    pred = predict_cuda(model, x_local)
    err = (pred - y_local.to(pred.device, torch.float32)).double()
    acc = torch.stack([(err * err).sum(), torch.tensor(float(err.numel()), dtype=torch.float64, device=pred.device)])
    comm.allreduce_sum_(acc)
    return float(acc[0] / acc[1])
