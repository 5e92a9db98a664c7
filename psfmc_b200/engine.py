"""
LikelihoodEngine: the Python face of the C ABI (include/psfmc_b200.h).

It takes exactly the constants the reference's Configuration + PSFSelector prepare
once per model (/root/reference/psfMC/ModelComponents/Configuration.py:38-52,
PSFSelector.py:16-43) plus the component program, and evaluates the lnL part of
MultiComponentModel.log_posterior (/root/reference/psfMC/models.py:213-241) for a
whole batch of parameter vectors per call on the GPU(s).
"""
import ctypes

import numpy as np

from . import _lib

KIND_CODES = {'sky': _lib.SKY, 'point': _lib.POINT, 'sersic': _lib.SERSIC}
SLOT_INDEX = {
    'sky': {'adu': _lib.P_ADU},
    'point': {'x': _lib.P_X, 'y': _lib.P_Y, 'mag': _lib.P_MAG},
    'sersic': {'x': _lib.P_X, 'y': _lib.P_Y, 'mag': _lib.P_MAG,
               'reff': _lib.P_REFF, 'reff_b': _lib.P_REFF_B,
               'index': _lib.P_INDEX, 'angle': _lib.P_ANGLE},
}


def _fill_slot(slot, spec):
    where, value = spec
    if where == 'theta':
        slot.theta_index = int(value)
        slot.value = 0.0
    elif where == 'const':
        slot.theta_index = -1
        slot.value = float(value)
    else:
        raise ValueError('slot must be ("theta", index) or ("const", value)')


class LikelihoodEngine(object):
    """
    :param obs_data, obs_var, bad_px: (H, W) arrays; ``obs_var`` is 1/ivm with
        +inf at data-bad pixels, ``bad_px`` nonzero where a pixel is excluded
    :param psfs, psf_vars: sequences of K normalised PSF images / variance maps
        (all the same shape, not larger than the observation)
    :param mag_zeropoint: magnitude zeropoint
    :param program: list of ``(kind, flags, slots)`` with kind in
        {'sky','point','sersic'}, flags a dict (``angle_degrees``,
        ``shift_method``), slots a dict name -> ('theta', i) | ('const', v)
    :param psf_index_slot: slot of the PSF index (('const', 0) for one PSF)
    :param precision: 'fp32' (float32 render/FFT, float64 accumulation; default),
        'fp64' (everything float64) or 'fp64_rawf32'
    :param devices: CUDA ordinals to shard batches over (default: current device)
    :param fp64_rescue: 'fp32' only: walkers whose float32 result is non-finite are
        repeated in float64 on the GPU (default); False = raw float32 behaviour
    """

    def __init__(self, obs_data, obs_var, bad_px, psfs, psf_vars, mag_zeropoint,
                 program, psf_index_slot=('const', 0), precision='fp32',
                 devices=None, max_batch=0, library=None, fp64_rescue=True):
        self._lib = _lib.load(library)
        self._handle = ctypes.c_void_p()
        obs = np.ascontiguousarray(obs_data, dtype=np.float64)
        var = np.ascontiguousarray(obs_var, dtype=np.float64)
        bad = np.ascontiguousarray(np.asarray(bad_px) != 0, dtype=np.uint8)
        if obs.ndim != 2 or var.shape != obs.shape or bad.shape != obs.shape:
            raise ValueError('obs_data, obs_var and bad_px must share a 2-D shape')
        psf = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64)
                                             for p in psfs]))
        pvar = np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64)
                                              for p in psf_vars]))
        if psf.shape != pvar.shape or psf.ndim != 3:
            raise ValueError('psfs and psf_vars must be K images of one shape')
        if len(program) > _lib.MAX_COMPONENTS:
            raise ValueError('too many components')
        comps = (_lib.Component * max(1, len(program)))()
        n_theta = 0
        for num, (kind, flags, slots) in enumerate(program):
            comp = comps[num]
            comp.kind = KIND_CODES[kind]
            comp.flags = 0
            if flags.get('angle_degrees'):
                comp.flags |= _lib.FLAG_ANGLE_DEGREES
            method = flags.get('shift_method', 'lanczos3')
            if kind == 'point':
                if method == 'bilinear':
                    comp.flags |= _lib.FLAG_BILINEAR
                elif method != 'lanczos3':
                    raise ValueError('Unknown shift method: {}'.format(method))
            for sidx in range(_lib.NSLOTS):
                comp.slot[sidx].theta_index = -1
            for name, sidx in SLOT_INDEX[kind].items():
                _fill_slot(comp.slot[sidx], slots[name])
                if slots[name][0] == 'theta':
                    n_theta = max(n_theta, int(slots[name][1]) + 1)
        desc = _lib.Desc()
        desc.abi_version = _lib.ABI_VERSION
        desc.height, desc.width = obs.shape
        dbl_p = ctypes.POINTER(ctypes.c_double)
        desc.obs_data = obs.ctypes.data_as(dbl_p)
        desc.obs_var = var.ctypes.data_as(dbl_p)
        desc.bad_px = bad.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        desc.n_psf, desc.psf_height, desc.psf_width = psf.shape
        desc.psf = psf.ctypes.data_as(dbl_p)
        desc.psf_var = pvar.ctypes.data_as(dbl_p)
        desc.mag_zeropoint = float(mag_zeropoint)
        desc.n_components = len(program)
        desc.components = comps
        _fill_slot(desc.psf_index, psf_index_slot)
        if psf_index_slot[0] == 'theta':
            n_theta = max(n_theta, int(psf_index_slot[1]) + 1)
        desc.precision = _lib.PRECISIONS[precision]
        devs = None
        if devices is not None:
            devs = (ctypes.c_int32 * len(devices))(*[int(d) for d in devices])
            desc.n_devices = len(devices)
            desc.devices = devs
        desc.max_batch = int(max_batch)
        desc.flags = 0 if fp64_rescue else _lib.DESC_NO_FP64_RESCUE
        _lib.check(self._lib, self._lib.psfmc_engine_create(
            ctypes.byref(desc), ctypes.byref(self._handle)))
        self.shape = obs.shape
        self.num_params = n_theta
        self.precision = precision

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, '_handle', None) is not None and self._handle.value:
            self._lib.psfmc_engine_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- hot path -----------------------------------------------------------
    def lnlike(self, thetas, out=None):
        """lnL for each row of ``thetas`` (B, D); -inf for non-finite results."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if ld < self.num_params:
            raise ValueError('theta has {} columns, the program needs {}'.format(
                ld, self.num_params))
        if out is None:
            out = np.empty(n_batch, dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_lnlike_batch(
            self._handle, thetas.ctypes.data_as(dbl_p), n_batch, ld,
            out.ctypes.data_as(dbl_p)))
        return out

    def lnlike_begin(self, thetas, out=None):
        """First half of :meth:`lnlike`: copy / enqueue and return at once, so the
        caller can evaluate the priors while the GPU works. Finish with
        :meth:`lnlike_end`, which returns the lnL array."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if ld < self.num_params:
            raise ValueError('theta has {} columns, the program needs {}'.format(
                ld, self.num_params))
        if out is None:
            out = np.empty(n_batch, dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_lnlike_batch_begin(
            self._handle, thetas.ctypes.data_as(dbl_p), n_batch, ld,
            out.ctypes.data_as(dbl_p)))
        self._in_flight = (thetas, out)      # keep both buffers alive until _end

    def lnlike_end(self):
        thetas, out = self._in_flight
        try:
            _lib.check(self._lib, self._lib.psfmc_lnlike_batch_end(self._handle))
        finally:
            self._in_flight = None
        return out

    # -- the sampler's inner loop in the library (psfmc_ensemble_run) ------------
    def lnpost(self, plan, thetas, out=None):
        """lnL + log-prior (``plan``: a ``_lib.PriorPlan``) per row; -inf where either
        is not finite (psfMC/models.py:205-211, 238-243)."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if out is None:
            out = np.empty(n_batch, dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_lnpost_batch(
            self._handle, ctypes.byref(plan), thetas.ctypes.data_as(dbl_p), n_batch, ld,
            out.ctypes.data_as(dbl_p)))
        return out

    def lnpost_sharded(self, plan, thetas, out=None):
        """One process per GPU (:meth:`peer_create` / :meth:`peer_connect` first): every
        rank passes the same rows, evaluates its share and gets the values of all rows
        (lnL gathered over peer memory; host buffers). ``plan`` None: the lnL itself."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if out is None:
            out = np.empty(n_batch, dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_lnpost_batch_sharded(
            self._handle, ctypes.byref(plan) if plan is not None else None,
            thetas.ctypes.data_as(dbl_p), n_batch, ld, out.ctypes.data_as(dbl_p)))
        return out

    def ensemble_run(self, plan, pos, lnprob, mt_key, mt_pos, n_iterations, a=2.0,
                     chain=None, lnprob_chain=None, chain_start=0, thin=1,
                     n_accepted=None, sharded=False, device=False):
        """``n_iterations`` stretch-move iterations of the whole ensemble inside the
        library (emcee 2.x semantics, numpy RandomState stream; see the header).
        ``pos`` (k, D), ``lnprob`` (k,), ``mt_key`` (624,) uint32, ``n_accepted`` (k,)
        are updated in place; ``mt_pos`` is a ``ctypes.c_int32``; ``chain``
        (k, L, D) / ``lnprob_chain`` (k, L) receive every ``thin``-th iteration from
        index ``chain_start`` on. ``sharded``: one process per GPU -- every rank makes
        the same call and evaluates its share of every half-ensemble, the lnL is gathered
        over peer memory (:meth:`peer_create` / :meth:`peer_connect` first). ``device``:
        proposals, priors and acceptance on the device (``PSFMC_ENS_DEVICE``); falls back
        to the host loop where the library cannot (several devices, a prior column that is
        evaluated in Python)."""
        dbl_p = ctypes.POINTER(ctypes.c_double)
        for arr, dtype in ((pos, np.float64), (lnprob, np.float64), (mt_key, np.uint32),
                           (chain, np.float64), (lnprob_chain, np.float64),
                           (n_accepted, np.float64)):
            if arr is not None and not (isinstance(arr, np.ndarray) and arr.dtype == dtype
                                        and arr.flags.c_contiguous and arr.flags.writeable):
                raise ValueError('ensemble_run needs writable C-contiguous arrays of '
                                 'the documented dtypes (updated in place)')
        ens = _lib.Ensemble()
        ens.n_walkers, ens.n_dim = pos.shape
        if lnprob.shape != (pos.shape[0],) or mt_key.shape != (624,):
            raise ValueError('lnprob must be (k,), mt_key (624,)')
        ens.a = float(a)
        ens.pos = pos.ctypes.data_as(dbl_p)
        ens.lnprob = lnprob.ctypes.data_as(dbl_p)
        ens.mt_key = mt_key.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        ens.mt_pos = ctypes.pointer(mt_pos)
        if chain is not None:
            if chain.ndim != 3 or chain.shape[0] != pos.shape[0] or \
                    chain.shape[2] != pos.shape[1]:
                raise ValueError('chain must be (k, L, D)')
            ens.chain = chain.ctypes.data_as(dbl_p)
            ens.chain_len = chain.shape[1]
        if lnprob_chain is not None:
            if lnprob_chain.ndim != 2 or lnprob_chain.shape[0] != pos.shape[0] or \
                    (chain is not None and lnprob_chain.shape[1] != chain.shape[1]):
                raise ValueError('lnprob_chain must be (k, L)')
            ens.lnprob_chain = lnprob_chain.ctypes.data_as(dbl_p)
            ens.chain_len = lnprob_chain.shape[1]
        ens.chain_start, ens.thin = int(chain_start), int(thin)
        ens.flags = (_lib.ENS_SHARDED if sharded else 0) | (_lib.ENS_DEVICE if device else 0)
        if n_accepted is not None:
            if n_accepted.shape != (pos.shape[0],):
                raise ValueError('n_accepted must be (k,)')
            ens.n_accepted = n_accepted.ctypes.data_as(dbl_p)
        code = self._lib.psfmc_ensemble_run(
            self._handle, ctypes.byref(plan) if plan is not None else None,
            ctypes.byref(ens), int(n_iterations))
        if code == 2 and (ens.flags & _lib.ENS_DEVICE):     # PSFMC_ERR_UNSUPPORTED
            ens.flags &= ~_lib.ENS_DEVICE
            code = self._lib.psfmc_ensemble_run(
                self._handle, ctypes.byref(plan) if plan is not None else None,
                ctypes.byref(ens), int(n_iterations))
        if code != 0:
            message = self._lib.psfmc_last_error().decode('utf-8', 'replace')
            if 'parameter value was' in message or 'lnprob returned NaN' in message:
                raise ValueError(message)      # what emcee raises
            raise _lib.EngineError(code, message)

    def lnlike_device(self, theta_ptr, n_batch, ld, lnl_ptr, stream=0,
                      device_slot=0):
        """Asynchronous evaluation on device-resident buffers (raw addresses,
        e.g. ``tensor.data_ptr()``), enqueued on CUDA stream ``stream``."""
        _lib.check(self._lib, self._lib.psfmc_lnlike_batch_device(
            self._handle, device_slot, ctypes.c_void_p(theta_ptr), n_batch, ld,
            ctypes.c_void_p(lnl_ptr), ctypes.c_void_p(stream)))

    # -- lnL gather over peer memory (one process per GPU) ----------------------
    def peer_create(self, capacity):
        """Allocate this rank's mailbox for gathered vectors of up to ``capacity``
        values; returns its 64-byte CUDA IPC handle (all-gather them, then
        :meth:`peer_connect`)."""
        handle = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.check(self._lib, self._lib.psfmc_peer_create(self._handle, int(capacity), handle))
        return handle.raw

    def peer_connect(self, rank, handles):
        """Map the mailboxes of all ranks (``handles``: list of the 64-byte handles in
        rank order)."""
        blob = b''.join(bytes(h) for h in handles)
        _lib.check(self._lib, self._lib.psfmc_peer_connect(
            self._handle, int(rank), len(handles), ctypes.c_char_p(blob)))

    def lnlike_exchange(self, theta_ptr, n_rows, ld, row_offset, n_total, gathered_ptr,
                        stream=0):
        """Evaluate this rank's rows (device-resident, raw address) and gather the lnL of
        all ranks through the peer mailboxes into ``gathered_ptr`` (device address of
        ``n_total`` doubles). Asynchronous on ``stream``."""
        _lib.check(self._lib, self._lib.psfmc_lnlike_batch_exchange(
            self._handle, ctypes.c_void_p(theta_ptr), n_rows, ld, row_offset, n_total,
            ctypes.c_void_p(gathered_ptr or None), ctypes.c_void_p(stream)))

    def peer_gathered(self):
        """Device address of the gathered vector of the last exchange (in this rank's
        mailbox; valid until the exchange after the next one)."""
        ptr = ctypes.c_void_p()
        _lib.check(self._lib, self._lib.psfmc_peer_gathered(self._handle, ctypes.byref(ptr)))
        return ptr.value

    def render(self, thetas, which=('raw_model', 'convolved_model', 'residual',
                                    'composite_ivm', 'point_source_subtracted')):
        """Blob images (psfMC/models.py:222-226) as dict name -> (B, H, W)."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if ld < self.num_params:
            raise ValueError('theta has {} columns, the program needs {}'.format(
                ld, self.num_params))
        bits = 0
        for name in which:
            bits |= _lib.IMAGE_BITS[name]
        ordered = [name for name, bit in sorted(_lib.IMAGE_BITS.items(),
                                                key=lambda kv: kv[1])
                   if bits & bit]
        out = np.empty((len(ordered), n_batch) + tuple(self.shape), dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_render_batch(
            self._handle, thetas.ctypes.data_as(dbl_p), n_batch, ld, bits,
            out.ctypes.data_as(dbl_p)))
        return {name: out[num] for num, name in enumerate(ordered)}

    def accumulate(self, thetas, which=('raw_model', 'convolved_model', 'residual',
                                        'composite_ivm', 'point_source_subtracted')):
        """Per-pixel SUMS of the blob images over all rows of ``thetas`` (float64,
        added up on the device) as dict name -> (H, W); 'composite_ivm' is the sum
        of 1/ivm (variance space), as the reference's running mean uses it."""
        thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)
        n_batch, ld = thetas.shape
        if ld < self.num_params:
            raise ValueError('theta has {} columns, the program needs {}'.format(
                ld, self.num_params))
        bits = 0
        for name in which:
            bits |= _lib.IMAGE_BITS[name]
        ordered = [name for name, bit in sorted(_lib.IMAGE_BITS.items(),
                                                key=lambda kv: kv[1])
                   if bits & bit]
        out = np.zeros((len(ordered),) + tuple(self.shape), dtype=np.float64)
        dbl_p = ctypes.POINTER(ctypes.c_double)
        _lib.check(self._lib, self._lib.psfmc_accumulate_batch(
            self._handle, thetas.ctypes.data_as(dbl_p), n_batch, ld, bits,
            out.ctypes.data_as(dbl_p)))
        return {name: out[num] for num, name in enumerate(ordered)}

    def profile(self, enable=True):
        """Bracket the dominant kernel of every lnL call with CUDA events."""
        _lib.check(self._lib, self._lib.psfmc_engine_profile(self._handle,
                                                             1 if enable else 0))

    def profile_read(self):
        """(summed kernel milliseconds, launches) since the last read."""
        ms, count = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self._lib, self._lib.psfmc_engine_profile_read(
            self._handle, ctypes.byref(ms), ctypes.byref(count)))
        return ms.value, count.value

    def info(self):
        info = _lib.Info()
        _lib.check(self._lib, self._lib.psfmc_engine_info(self._handle,
                                                          ctypes.byref(info)))
        return {name: getattr(info, name) for name, _ in _lib.Info._fields_}


def fp32_peak_tflops(device=0, library=None):
    """Measured FP32 FMA throughput of one device (TFLOP/s), the FP32 roofline
    denominator."""
    lib = _lib.load(library)
    tflops, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib, lib.psfmc_fp32_peak_probe(device, ctypes.byref(tflops),
                                              ctypes.byref(ms)))
    return tflops.value
