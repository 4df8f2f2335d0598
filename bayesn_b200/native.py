"""ctypes binding of ``libbayesn_b200.so`` (C ABI in ``include/bayesn_b200.h``).

There is no CPU fallback: if the library is missing or no CUDA device is present the product
path raises.  Tensors are torch CUDA tensors used purely as device buffers (``data_ptr()``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BAYESN_B200_LIB', os.path.join(_HERE, 'csrc', 'libbayesn_b200.so'))

NBINS, BPAD, NPHASE, OVERRUN = 300, 304, 105, 320
RV_FIXED, RV_POP, RV_UNIFORM = 0, 1, 2

_fp = C.POINTER(C.c_float)


class ModelDesc(C.Structure):
    _fields_ = [('L', C.c_int32), ('T', C.c_int32), ('rv_mode', C.c_int32),
                ('tau_knots', _fp), ('KD_t', _fp), ('J_l', _fp), ('hsiao', _fp), ('M_fitz', _fp), ('a_dust', _fp),
                ('W0', _fp), ('W1', _fp), ('L_Sigma', _fp),
                ('tauA', C.c_float), ('Ds_prior_sd', C.c_float),
                ('RV', C.c_float), ('mu_R', C.c_float), ('sigma_R', C.c_float), ('trunc_R', C.c_float),
                ('fix_tmax', C.c_int32), ('fix_theta', C.c_int32), ('fix_AV', C.c_int32),
                ('theta_val', C.c_float), ('AV_val', C.c_float)]


class DatasetDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('N_obs', C.c_int32), ('n_b', C.c_int32), ('wt_stride', C.c_int32),
                ('obs4', C.c_void_p), ('n_valid', C.c_void_p), ('weights', C.c_void_p), ('support', C.c_void_p),
                ('muhat', C.c_void_p), ('lconst', C.c_void_p), ('muhat_err2', C.c_void_p)]


class TilesDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('n_b', C.c_int32), ('n_bw', C.c_int32),
                ('master', C.c_void_p), ('z', C.c_void_p), ('ebv', C.c_void_p), ('fold', C.c_void_p),
                ('bscale', C.c_void_p), ('used', C.c_void_p), ('model_wave', C.c_void_p),
                ('band_spacing', C.c_double), ('RV_MW', C.c_double),
                ('f99_xk', C.c_double * 9), ('f99_yk', C.c_double * 9), ('f99_mk', C.c_double * 9)]


class NutsCfg(C.Structure):
    _fields_ = [('num_warmup', C.c_int32), ('num_samples', C.c_int32), ('num_chains', C.c_int32),
                ('max_tree_depth', C.c_int32), ('target_accept', C.c_float), ('init_step_size', C.c_float),
                ('adapt_step_size', C.c_int32), ('adapt_mass_matrix', C.c_int32),
                ('regularize_mass_matrix', C.c_int32), ('seed', C.c_uint64), ('unit_offset', C.c_uint64),
                ('warps_per_chain', C.c_int32), ('reserved', C.c_int32)]


EXPORTS = ['bayesn_abi_version', 'bayesn_ctx_create', 'bayesn_ctx_destroy', 'bayesn_last_error',
           'bayesn_set_model', 'bayesn_bind_dataset', 'bayesn_param_dim', 'bayesn_logp_grad_fit',
           'bayesn_flux_batch', 'bayesn_nuts_workspace_bytes', 'bayesn_nuts_fit', 'bayesn_launch_count',
           'bayesn_init_to_median', 'bayesn_ffma_peak', 'bayesn_train_dims', 'bayesn_train_begin',
           'bayesn_train_finish', 'bayesn_train_nuts_init', 'bayesn_train_nuts_sn', 'bayesn_train_nuts_ctl',
           'bayesn_train_nuts_status', 'bayesn_tiles_support', 'bayesn_tiles_fill', 'bayesn_train_set_exchange', 'bayesn_spectra_batch', 'bayesn_chain_flux',
           'bayesn_logp_grad_fit_team', 'bayesn_train_warps_per_unit', 'bayesn_train_fuse_reduction', 'bayesn_debug_poison_smem']

_lib = None


def load_library():
    """Load the shared library (fails loudly if it has not been built: run
    ``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C bayesn_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} not found: the CUDA library has not been built. There is no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.bayesn_abi_version.restype = C.c_int
    lib.bayesn_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.bayesn_ctx_destroy.argtypes = [C.c_void_p]
    lib.bayesn_ctx_destroy.restype = None
    lib.bayesn_last_error.argtypes = [C.c_void_p]
    lib.bayesn_last_error.restype = C.c_char_p
    lib.bayesn_set_model.argtypes = [C.c_void_p, C.POINTER(ModelDesc)]
    lib.bayesn_bind_dataset.argtypes = [C.c_void_p, C.POINTER(DatasetDesc)]
    lib.bayesn_param_dim.argtypes = [C.c_void_p]
    lib.bayesn_logp_grad_fit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _optional(lambda: setattr(lib.bayesn_logp_grad_fit_team, 'argtypes', [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]))
    lib.bayesn_flux_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 12 + [
        C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
    lib.bayesn_nuts_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(NutsCfg)]
    lib.bayesn_nuts_workspace_bytes.restype = C.c_size_t
    lib.bayesn_nuts_fit.argtypes = [C.c_void_p, C.POINTER(NutsCfg)] + [C.c_void_p] * 5 + [C.c_size_t, C.c_void_p]
    lib.bayesn_launch_count.argtypes = [C.c_void_p]
    lib.bayesn_init_to_median.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.bayesn_ffma_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.bayesn_launch_count.restype = C.c_longlong
    lib.bayesn_spectra_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 10
    _optional(lambda: setattr(lib.bayesn_chain_flux, 'argtypes', [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p] * 10 + [C.c_int, C.c_void_p, C.c_void_p]))
    lib.bayesn_tiles_support.argtypes = [C.c_void_p, C.POINTER(TilesDesc), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bayesn_tiles_fill.argtypes = [C.c_void_p, C.POINTER(TilesDesc), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bayesn_train_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    lib.bayesn_train_begin.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.bayesn_train_finish.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.bayesn_train_nuts_init.argtypes = [C.c_void_p, C.POINTER(NutsCfg), C.c_int] + [C.c_void_p] * 6
    lib.bayesn_train_nuts_sn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bayesn_train_nuts_ctl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    _optional(lambda: setattr(lib.bayesn_train_warps_per_unit, 'argtypes', [C.c_void_p]))
    _optional(lambda: setattr(lib.bayesn_train_fuse_reduction, 'argtypes', [C.c_void_p, C.c_int]))
    _optional(lambda: setattr(lib.bayesn_debug_poison_smem, 'argtypes', [C.c_void_p, C.c_void_p]))
    lib.bayesn_train_nuts_status.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
    lib.bayesn_train_set_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                              C.c_void_p]
    _lib = lib
    return lib


def _optional(fn):
    """Bind a symbol that older builds of the library do not have (A/B measurements against an earlier build through
    BAYESN_B200_LIB); the product build exports everything (tests/test_host.py)."""
    try:
        fn()
    except AttributeError:
        if 'BAYESN_B200_LIB' not in os.environ:
            raise


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Context:
    """One device context: model tables + one bound data set."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('bayesn_b200 needs a CUDA device (sm_100a); there is no CPU fallback.')
        self.lib = load_library()
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.bayesn_ctx_create(self.device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f'bayesn_ctx_create({device}) failed with {rc}')
        self.h = h
        self._keep = {}
        self.D = None

    def close(self):
        if getattr(self, 'h', None):
            self.lib.bayesn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f'{what} failed ({rc}): {self.lib.bayesn_last_error(self.h).decode()}')

    def launches(self):
        return int(self.lib.bayesn_launch_count(self.h))

    # ------------------------------------------------------------------ model
    def set_model(self, *, tau_knots, KD_t, J_l, hsiao, M_fitz, a_dust, W0, W1, L_Sigma, tauA, Ds_prior_sd,
                  rv_mode=RV_FIXED, RV=0.0, mu_R=0.0, sigma_R=1.0, trunc_R=1.2, fix_tmax=False, fix_theta=False,
                  fix_AV=False, theta_val=0.0, AV_val=0.0):
        arrs = dict(tau_knots=_f32(tau_knots), KD_t=_f32(KD_t), J_l=_f32(J_l), hsiao=_f32(hsiao),
                    M_fitz=_f32(M_fitz), a_dust=_f32(a_dust), W0=_f32(W0), W1=_f32(W1), L_Sigma=_f32(L_Sigma))
        L, T = arrs['W0'].shape
        assert arrs['J_l'].shape == (NBINS, L) and arrs['hsiao'].shape == (NBINS, NPHASE)
        assert arrs['M_fitz'].shape == (NBINS, 9) and arrs['KD_t'].shape == (T, T)
        ne = (L - 2) * T
        assert arrs['L_Sigma'].shape == (ne, ne)
        d = ModelDesc()
        d.L, d.T, d.rv_mode = L, T, int(rv_mode)
        for k, a in arrs.items():
            setattr(d, k, a.ctypes.data_as(_fp))
        d.tauA, d.Ds_prior_sd = float(tauA), float(Ds_prior_sd)
        d.RV, d.mu_R, d.sigma_R, d.trunc_R = float(RV), float(mu_R), float(sigma_R), float(trunc_R)
        d.fix_tmax, d.fix_theta, d.fix_AV = int(bool(fix_tmax)), int(bool(fix_theta)), int(bool(fix_AV))
        d.theta_val, d.AV_val = float(theta_val), float(AV_val)
        self._check(self.lib.bayesn_set_model(self.h, C.byref(d)), 'bayesn_set_model')
        self.D = int(self.lib.bayesn_param_dim(self.h))
        self.L, self.T, self.n_eps = L, T, ne
        return self.D

    # ------------------------------------------------------------------ dataset
    def bind_dataset(self, packed):
        """``packed``: dict of torch CUDA tensors from :func:`bayesn_b200.packing.pack_dataset`."""
        self._keep['data'] = packed
        d = DatasetDesc()
        d.N, d.N_obs, d.n_b = int(packed['N']), int(packed['N_obs']), int(packed['n_b'])
        d.wt_stride = int(packed['wt_stride'])
        for k in ('obs4', 'n_valid', 'weights', 'support', 'muhat', 'lconst'):
            setattr(d, k, packed[k].data_ptr())
        d.muhat_err2 = packed['muhat_err2'].data_ptr() if packed.get('muhat_err2') is not None else None
        self._check(self.lib.bayesn_bind_dataset(self.h, C.byref(d)), 'bayesn_bind_dataset')
        self.N = d.N

    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    def init_to_median(self, n_chains, seed=0, out=None, sn_offset=0):
        """Initial unconstrained positions (N, C, D) for the bound data set (device kernel)."""
        import torch
        if out is None:
            out = torch.empty((self.N, n_chains, self.D), dtype=torch.float32, device=f'cuda:{self.device}')
        self._check(self.lib.bayesn_init_to_median(self.h, int(n_chains), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                   int(sn_offset) * int(n_chains), out.data_ptr(), self._stream()),
                    'bayesn_init_to_median')
        return out

    def poison_shared_memory(self):
        """Test aid: leave NaN bit patterns in every SM's shared memory."""
        self._check(self.lib.bayesn_debug_poison_smem(self.h, self._stream()), 'bayesn_debug_poison_smem')

    def ffma_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.bayesn_ffma_peak(self.h, C.byref(v)), 'bayesn_ffma_peak')
        return float(v.value)

    def logp_grad(self, q, warps_per_chain=1):
        """q: (N, C, D) float32 CUDA tensor -> (logp (N, C), grad (N, C, D))."""
        import torch
        assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()
        N, Cn, D = q.shape
        assert N == self.N and D == self.D
        logp = torch.empty((N, Cn), dtype=torch.float32, device=q.device)
        grad = torch.empty_like(q)
        if int(warps_per_chain) <= 1:
            self._check(self.lib.bayesn_logp_grad_fit(self.h, Cn, q.data_ptr(), logp.data_ptr(), grad.data_ptr(),
                                                      self._stream()), 'bayesn_logp_grad_fit')
        else:
            self._check(self.lib.bayesn_logp_grad_fit_team(self.h, Cn, int(warps_per_chain), q.data_ptr(), logp.data_ptr(),
                                                           grad.data_ptr(), self._stream()), 'bayesn_logp_grad_fit_team')
        return logp, grad

    def resident_warps(self):
        """Warp slots of the device for the sampler kernel (16 warps per SM at 128 registers)."""
        import torch
        return 16 * torch.cuda.get_device_properties(self.device).multi_processor_count

    def auto_warps_per_chain(self, n_units):
        """Latency mode width: the largest K in (8, 4, 2, 1) whose K * n_units warps still fit the device in
        one resident wave - more warps per chain only pay while there are idle warp slots."""
        slots = self.resident_warps()
        for k in (8, 4, 2):
            if n_units * k <= slots:
                return k
        return 1

    def nuts(self, q_init, num_warmup=250, num_samples=250, max_tree_depth=10, target_accept=0.8,
             init_step_size=1.0, adapt_step_size=True, adapt_mass_matrix=True, regularize_mass_matrix=True, seed=0,
             out=None, sn_offset=0, warps_per_chain=1):
        """Run the device NUTS for every (SN, chain).  Returns dict of CUDA tensors.  ``warps_per_chain``: 1
        (throughput layout), 2 / 4 / 8 (latency mode) or 'auto'."""
        import torch
        assert q_init.is_cuda and q_init.dtype == torch.float32 and q_init.is_contiguous()
        N, Cn, D = q_init.shape
        assert N == self.N and D == self.D
        if warps_per_chain == 'auto':
            warps_per_chain = self.auto_warps_per_chain(N * Cn)
        cfg = NutsCfg(int(num_warmup), int(num_samples), int(Cn), int(max_tree_depth), float(target_accept),
                      float(init_step_size), int(adapt_step_size), int(adapt_mass_matrix),
                      int(regularize_mass_matrix), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sn_offset) * int(Cn),
                      int(warps_per_chain), 0)
        if out is None:
            out = {}
        out['warps_per_chain'] = int(warps_per_chain)
        wb = int(self.lib.bayesn_nuts_workspace_bytes(self.h, C.byref(cfg)))
        dev = q_init.device
        if out is None:
            out = {}
        ws = out.get('work')
        if ws is None or ws.numel() * 4 < wb:
            ws = out['work'] = torch.empty((wb + 3) // 4, dtype=torch.float32, device=dev)
        for name, shape in (('samples', (N, Cn, num_samples, D)), ('stats', (N, Cn, num_warmup + num_samples, 5)),
                            ('chain_info', (N, Cn, 4 + D))):
            t = out.get(name)
            if t is None or tuple(t.shape) != shape:
                out[name] = torch.empty(shape, dtype=torch.float32, device=dev)
        self._check(self.lib.bayesn_nuts_fit(self.h, C.byref(cfg), q_init.data_ptr(), out['samples'].data_ptr(),
                                             out['stats'].data_ptr(), out['chain_info'].data_ptr(), ws.data_ptr(),
                                             ws.numel() * 4, self._stream()), 'bayesn_nuts_fit')
        return out

    # ------------------------------------------------------------------ band-weight tiles on the device
    def tiles_desc(self, N, n_b, dev, builder):
        d = TilesDesc()
        d.N, d.n_b, d.n_bw = int(N), int(n_b), int(dev['master'].shape[1])
        for k in ('master', 'z', 'ebv', 'fold', 'bscale', 'used', 'model_wave'):
            setattr(d, k, dev[k].data_ptr())
        d.band_spacing, d.RV_MW = builder.band_spacing, builder.RV_MW
        for i in range(9):
            d.f99_xk[i], d.f99_yk[i], d.f99_mk[i] = float(builder.xk[i]), float(builder.yk[i]), float(builder.mk[i])
        self._keep['tiles'] = dev
        return d

    def tiles_support(self, desc, support, total):
        self._check(self.lib.bayesn_tiles_support(self.h, C.byref(desc), support.data_ptr(), total.data_ptr(),
                                                  self._stream()), 'bayesn_tiles_support')

    def tiles_fill(self, desc, support, wt_stride, weights):
        self._check(self.lib.bayesn_tiles_fill(self.h, C.byref(desc), support.data_ptr(), int(wt_stride),
                                               weights.data_ptr(), self._stream()), 'bayesn_tiles_fill')

    # ------------------------------------------------------------------ training
    def enable_train_exchange(self, group, n_chains):
        """Switch the exchange of the per-leapfrog training sums from the caller's NCCL all-reduce to
        the fused peer-memory exchange: a symmetric (peer-mapped over NVLink) buffer per rank, filled
        by every rank's reduction kernel with plain stores and consumed by the control kernel."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            self.disable_train_exchange()
            return
        R = self.train_dims()[2]
        dev = torch.device(f'cuda:{self.device}')
        n_data = 2 * world * n_chains * R
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        buf = symm.empty(n_data + 128, dtype=torch.float64, device=dev)      # data + 2 * world * C uint32 flags
        hdl = symm.rendezvous(buf, group)
        buf.zero_()
        scratch = torch.zeros(4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group)
        off = int(getattr(hdl, 'offset', 0) or 0)
        bases = [int(p) + off for p in hdl.buffer_ptrs]
        pb = (C.c_uint64 * world)(*bases)
        pf = (C.c_uint64 * world)(*[b + 8 * n_data for b in bases])
        self._check(self.lib.bayesn_train_set_exchange(self.h, world, rank, pb, pf, scratch.data_ptr()),
                    'bayesn_train_set_exchange')
        self._keep['xchg'] = (buf, hdl, scratch, n_chains)

    def disable_train_exchange(self):
        self._check(self.lib.bayesn_train_set_exchange(self.h, 1, 0, None, None, None), 'bayesn_train_set_exchange')
        self._keep.pop('xchg', None)

    def train_exchange_error(self):
        x = self._keep.get('xchg')
        return bool(x[2][2].item()) if x else False

    def train_dims(self):
        """(G, Dl, R): unconstrained global / per-SN latent dimensions and the length of the
        per-chain reduction buffer."""
        G, Dl, R = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.bayesn_train_dims(self.h, C.byref(G), C.byref(Dl), C.byref(R)), 'bayesn_train_dims')
        return G.value, Dl.value, R.value

    def train_logp_grad(self, qg, ql, group=None, out=None):
        """Log joint density of ``train_model_globalRV`` and its gradient.  ``qg`` (C, G) and ``ql``
        (N, C, Dl) are float32 CUDA tensors.  With ``group`` (a torch.distributed process group whose
        ranks hold disjoint SN shards and identical ``qg``) the per-SN sums are all-reduced between
        the two native calls - the single exchange step of the training path.
        Returns (logp (C,) float64, grad_g (C, G), grad_l (N, C, Dl))."""
        import torch
        assert qg.is_cuda and qg.dtype == torch.float32 and qg.is_contiguous()
        assert ql.is_cuda and ql.dtype == torch.float32 and ql.is_contiguous()
        if 'xchg' in self._keep:
            # with the fused peer exchange switched on the reduction kernel publishes into the peers' buffers and
            # `red` stays stale: an NCCL all-reduce of it here would be wrong
            raise RuntimeError('train_logp_grad: the peer-memory exchange of a running train_nuts is enabled on this '
                               'context; call disable_train_exchange() first')
        Cn, G = qg.shape
        N, C2, Dl = ql.shape
        Gd, Dld, R = self.train_dims()
        assert (G, Dl, C2, N) == (Gd, Dld, Cn, self.N), ((G, Dl, C2, N), (Gd, Dld, Cn, self.N))
        out = {} if out is None else out
        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != shape:
                t = out[name] = torch.empty(shape, dtype=dtype, device=qg.device)
            return t
        red = buf('red', (Cn, R), torch.float64)
        grad_l = buf('grad_l', (N, Cn, Dl), torch.float32)
        grad_g = buf('grad_g', (Cn, G), torch.float32)
        logp = buf('logp', (Cn,), torch.float64)
        st = self._stream()
        self._check(self.lib.bayesn_train_begin(self.h, Cn, qg.data_ptr(), ql.data_ptr(), grad_l.data_ptr(),
                                                red.data_ptr(), st), 'bayesn_train_begin')
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
        self._check(self.lib.bayesn_train_finish(self.h, Cn, qg.data_ptr(), red.data_ptr(), logp.data_ptr(),
                                                 grad_g.data_ptr(), st), 'bayesn_train_finish')
        return logp, grad_g, grad_l

    def train_nuts(self, qg0, ql0, num_warmup=500, num_samples=500, max_tree_depth=10, target_accept=0.8,
                   init_step_size=0.1, adapt_step_size=True, adapt_mass_matrix=True, regularize_mass_matrix=False,
                   seed=0, group=None, sn_offset=0, passes_per_graph=32, use_graph=True, max_passes=None,
                   progress=None, exchange='auto', fuse_reduction=True):
        """Device-resident NUTS over the joint training posterior (numpyro call at reference
        :1767-1770, :1913-1919).  ``qg0`` (C, G), ``ql0`` (N, C, Dl) float32 CUDA tensors.  One pass =
        one leapfrog of every chain = SN kernel + reduction -> [all-reduce over ``group``] -> control
        kernel; passes are replayed from a CUDA graph and the host only polls the chains' status.
        Returns dict(samples_g (C, S, G), samples_l (N, C, S, Dl), stats (C, W+S, 5), info (C, 8),
        passes)."""
        import torch
        assert qg0.is_cuda and qg0.dtype == torch.float32 and qg0.is_contiguous()
        assert ql0.is_cuda and ql0.dtype == torch.float32 and ql0.is_contiguous()
        Cn, G = qg0.shape
        N, C2, Dl = ql0.shape
        Gd, Dld, R = self.train_dims()
        assert (G, Dl, C2, N) == (Gd, Dld, Cn, self.N)
        dev = qg0.device
        cfg = NutsCfg(int(num_warmup), int(num_samples), int(Cn), int(max_tree_depth), float(target_accept),
                      float(init_step_size), int(adapt_step_size), int(adapt_mass_matrix),
                      int(regularize_mass_matrix), int(seed) & 0xFFFFFFFFFFFFFFFF, 0, 1, 0)
        out = dict(samples_g=torch.zeros((Cn, num_samples, G), dtype=torch.float32, device=dev),
                   samples_l=torch.zeros((N, Cn, num_samples, Dl), dtype=torch.float32, device=dev),
                   stats=torch.zeros((Cn, num_warmup + num_samples, 5), dtype=torch.float32, device=dev))
        red = torch.zeros((Cn, R), dtype=torch.float64, device=dev)
        self._keep['train_nuts'] = (out, red, qg0, ql0)
        self._check(self.lib.bayesn_train_nuts_init(self.h, C.byref(cfg), int(sn_offset), qg0.data_ptr(),
                                                    ql0.data_ptr(), out['samples_g'].data_ptr(),
                                                    out['samples_l'].data_ptr(), out['stats'].data_ptr(),
                                                    self._stream()), 'bayesn_train_nuts_init')
        if group is not None:
            import torch.distributed as dist
        world = dist.get_world_size(group) if group is not None else 1
        if exchange == 'auto':
            # measured on 8 x B200 (profiles/README.md): NCCL's LL all-reduce of the 61 KB buffer wins at two ranks,
            # the one-hop peer stores win from four ranks on (ring latency grows with the rank count)
            exchange = 'peer' if world >= 4 else 'nccl'
        peer = group is not None and exchange == 'peer' and world > 1
        if peer:
            self.enable_train_exchange(group, Cn)        # fused NVLink peer-memory exchange instead of NCCL
        else:
            self.disable_train_exchange()
        # one GPU / peer exchange: the control kernel adds the chunk partials (and exchanges them) itself - three
        # kernels per pass; with NCCL the reduced buffer has to exist between the two calls - four kernels + NCCL
        fused = (group is None or world == 1 or peer) and fuse_reduction
        self._check(self.lib.bayesn_train_fuse_reduction(self.h, int(fused)), 'bayesn_train_fuse_reduction')

        def one_pass():
            st = self._stream()
            self._check(self.lib.bayesn_train_nuts_sn(self.h, red.data_ptr(), st), 'bayesn_train_nuts_sn')
            if group is not None and not peer:
                dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
            self._check(self.lib.bayesn_train_nuts_ctl(self.h, red.data_ptr(), st), 'bayesn_train_nuts_ctl')

        status = (C.c_float * (8 * Cn))()

        def poll():
            rc = self.lib.bayesn_train_nuts_status(self.h, status, self._stream())
            if rc < 0:
                self._check(rc, 'bayesn_train_nuts_status')
            return rc == 1

        one_pass()                                   # the initial evaluation (also warms every lazy allocation)
        torch.cuda.synchronize()
        passes = 1
        graph = None
        if use_graph:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                one_pass()
                passes += 1
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(passes_per_graph):
                    one_pass()
        done = False
        import time as _time
        torch.cuda.synchronize()
        t_loop, p_loop = _time.perf_counter(), passes
        while not done:
            if graph is not None:
                graph.replay()
                passes += passes_per_graph
            else:
                for _ in range(passes_per_graph):
                    one_pass()
                passes += passes_per_graph
            done = poll()
            if progress is not None:
                progress(passes, np.ctypeslib.as_array(status).reshape(Cn, 8).copy())
            if max_passes is not None and passes >= max_passes:
                break
        torch.cuda.synchronize()
        out['loop_seconds'], out['loop_passes'] = _time.perf_counter() - t_loop, passes - p_loop
        # flush: the latent slices apply the last transition's actions (final draw) in one more SN pass
        self._check(self.lib.bayesn_train_nuts_sn(self.h, red.data_ptr(), self._stream()), 'bayesn_train_nuts_sn')
        torch.cuda.synchronize()
        poll()
        out['info'] = np.ctypeslib.as_array(status).reshape(Cn, 8).copy()
        out['passes'] = passes
        out['finished'] = bool(done)
        out['warps_per_unit'] = int(self.lib.bayesn_train_warps_per_unit(self.h))
        out['exchange'] = ('peer' if peer else 'nccl') if world > 1 else 'none'
        out['launches_per_pass'] = 3 if fused else 4
        self.lib.bayesn_train_fuse_reduction(self.h, 0)
        if peer:
            if self.train_exchange_error():
                raise RuntimeError('peer-memory exchange timed out waiting for a rank')
            dist.barrier(group)
            self.disable_train_exchange()
        return out

    def spectra_batch(self, theta, AV, W0, W1, eps, RV, J_t, hsiao_interp):
        """get_spectra with the reference's array layouts (CUDA tensors) -> (N, 300, N_obs)."""
        import torch
        f32 = lambda t: t.to(torch.float32).contiguous()
        theta, AV, W0, W1, eps, RV, J_t, hsiao_interp = map(f32, (theta, AV, W0, W1, eps, RV, J_t, hsiao_interp))
        N, N_obs = theta.shape[0], J_t.shape[2]
        out = torch.empty((N, NBINS, N_obs), dtype=torch.float32, device=theta.device)
        self._check(self.lib.bayesn_spectra_batch(self.h, N, N_obs, theta.data_ptr(), AV.data_ptr(), W0.data_ptr(),
                                                  W1.data_ptr(), eps.data_ptr(), RV.data_ptr(), J_t.data_ptr(),
                                                  hsiao_interp.data_ptr(), out.data_ptr(), self._stream()),
                    'bayesn_spectra_batch')
        return out

    def chain_flux(self, t, theta, AV, tmax, dD, eps, bands, weights, support, wt_stride, RV=None, mag=False):
        """Model photometry of posterior draws (``get_flux_from_chains``, reference :3426-3524) in one
        launch.  All arguments are CUDA tensors: ``t`` (n_t,), ``theta/AV/tmax/dD[/RV]`` (N, S), ``eps``
        (N, S, n_eps), ``bands`` (N, max_bands) int32 tile rows (-1 = padding), ``weights`` / ``support`` the
        compact tiles.  Returns (N, S, max_bands, n_t) float32."""
        import torch
        f32 = lambda a: a.to(torch.float32).contiguous()
        t, theta, AV, tmax, dD, eps = map(f32, (t, theta, AV, tmax, dD, eps))
        RVt = f32(RV) if RV is not None else None
        bands = bands.to(torch.int32).contiguous()
        N, S = theta.shape
        assert eps.shape == (N, S, self.n_eps) and bands.shape[0] == N and support.shape[0] == N
        n_t, max_bands, n_b = t.shape[0], bands.shape[1], support.shape[1]
        out = torch.empty((N, S, max_bands, n_t), dtype=torch.float32, device=theta.device)
        self._check(self.lib.bayesn_chain_flux(
            self.h, N, S, n_t, max_bands, n_b, int(bool(mag)), int(RVt is not None), t.data_ptr(), theta.data_ptr(),
            AV.data_ptr(), tmax.data_ptr(), dD.data_ptr(), RVt.data_ptr() if RVt is not None else None, eps.data_ptr(),
            bands.data_ptr(), weights.data_ptr(), support.data_ptr(), int(wt_stride), out.data_ptr(), self._stream()),
            'bayesn_chain_flux')
        return out

    def flux_batch(self, M0, theta, AV, W0, W1, eps, Ds, RV, band_indices, mask, J_t, hsiao_interp, weights,
                   band_log10_scale, mag=False):
        """Forward model with the reference's array layouts; all array arguments are CUDA tensors."""
        import torch
        N_obs, N = band_indices.shape
        n_b = weights.shape[2]
        f32 = lambda t: t.to(torch.float32).contiguous()
        theta, AV, W0, W1, eps, Ds, RV = map(f32, (theta, AV, W0, W1, eps, Ds, RV))
        J_t, hsiao_interp, weights = map(f32, (J_t, hsiao_interp, weights))
        band_indices = band_indices.to(torch.int32).contiguous()
        mask = mask.to(torch.uint8).contiguous()
        out = torch.empty((N_obs, N), dtype=torch.float32, device=theta.device)
        bs = np.ascontiguousarray(np.asarray(band_log10_scale, dtype=np.float64))
        assert bs.shape == (n_b,)
        self._check(self.lib.bayesn_flux_batch(
            self.h, N, N_obs, n_b, int(bool(mag)), float(M0), theta.data_ptr(), AV.data_ptr(), W0.data_ptr(),
            W1.data_ptr(), eps.data_ptr(), Ds.data_ptr(), RV.data_ptr(), band_indices.data_ptr(), mask.data_ptr(),
            J_t.data_ptr(), hsiao_interp.data_ptr(), weights.data_ptr(), bs.ctypes.data_as(C.POINTER(C.c_double)),
            out.data_ptr(), self._stream()), 'bayesn_flux_batch')
        return out
