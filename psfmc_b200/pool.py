"""
BatchPool -- the drop-in boundary.

emcee 2.x evaluates the posterior with
    results = list(pool.map(self.lnprobfn, [p[i] for i in range(len(p))]))
(EnsembleSampler._get_lnprob) and the reference's driver builds the sampler with
``lnpostfn=mc_model.log_posterior, kwargs={'model': mc_model}`` and no pool
(/root/reference/psfMC/fitting.py:56-58). A ``BatchPool`` handed over as
``pool=`` turns that per-walker map into ONE batched GPU call per (half-)ensemble:
priors column-vectorised on the host, every walker with a finite prior through
the C ABI (psfmc_lnlike_batch) in a single call.
"""
from itertools import repeat

import numpy as np


def rows_as_block(rows):
    """The (B, D) float64 array behind a list of B parameter vectors. emcee 2.x builds
    the list as ``[p[i] for i in range(len(p))]`` -- B views of consecutive rows of ONE
    array -- so when the first and last element are views of the same base, spaced
    like the rows of a C-contiguous (B, D) block, the block is re-assembled from the
    base without touching the B views (two address checks instead of a 2048-way
    concatenation); anything else is concatenated or stacked."""
    first, last = rows[0], rows[-1]
    count = len(rows)
    try:
        if (first.dtype == np.float64 and first.ndim == 1 and first.base is not None
                and first.base is last.base and first.flags.c_contiguous
                and last.shape == first.shape):
            ndim = first.shape[0]
            a0 = first.__array_interface__['data'][0]
            a1 = last.__array_interface__['data'][0]
            if a1 - a0 == (count - 1) * ndim * 8:
                mid = rows[count // 2]
                if (mid.base is first.base and mid.shape == first.shape and
                        mid.__array_interface__['data'][0] - a0 == (count // 2) * ndim * 8):
                    base = first.base
                    b0 = base.__array_interface__['data'][0]
                    if base.dtype == np.float64 and base.flags.c_contiguous and \
                            (a0 - b0) % 8 == 0:
                        start = (a0 - b0) // 8
                        flat = base.reshape(-1)[start:start + count * ndim]
                        if flat.size == count * ndim:
                            return flat.reshape(count, ndim)
        block = np.concatenate(rows).reshape(count, -1)
        if block.dtype != np.float64 or first.ndim != 1:
            raise ValueError
        return block
    except (ValueError, TypeError, AttributeError):
        return np.stack([np.asarray(p, dtype=np.float64) for p in rows])


def float32_is_enough(model, start_positions):
    """The device loop and the sharded calls have no float64 repeat: a walker whose float32
    transform comes out non-finite counts as -inf there, where the reference -- and the
    host call, which repeats it in float64 -- has a finite posterior. That is nothing for
    the odd prior-drawn walker and a stuck chain for a model whose dynamic range float32
    cannot hold (a component 10^5 times brighter than the pixel noise: DESIGN.md 4.5,
    'outside the priors'). Probe: the starting ensemble through the host call of THIS
    process's engine, which repeats and counts; more than 1 % repeated -> False, with a
    warning. (Deterministic: every rank of a sharded job comes to the same answer.)"""
    engine = model.engine
    before = engine.info()['rescued_total']
    model.log_posterior_batch(np.ascontiguousarray(start_positions, dtype=np.float64))
    repeated = engine.info()['rescued_total'] - before
    enough = repeated <= 0.01 * len(start_positions)
    if not enough:
        import warnings
        warnings.warn(
            'psfmc_b200: {} of the {} starting walkers needed the float64 repeat of the '
            'float32 likelihood; the sampler stays on the host calls (which repeat them) '
            'instead of the device loop / the sharded library calls (which would count them '
            'as -inf). precision=\'fp64\' suits this model better.'.format(
                repeated, len(start_positions)))
    return enough


class BatchPool(object):
    """
    :param model: :class:`psfmc_b200.models.MultiComponentModel`
    :param with_blobs: return the five per-walker images as blobs like the
        reference's log_posterior (slow: 5*H*W doubles per walker travel to the
        host). Default: empty blob dicts; posterior images are re-rendered from
        the chain afterwards (the reference's own alternative path,
        psfMC/analysis/images.py:74-83).
    """

    def __init__(self, model, with_blobs=False):
        self.model = model
        self.with_blobs = with_blobs
        self.calls = 0
        self.evaluations = 0
        self._fp32_enough = None        # see _float32_is_enough

    def map(self, func, iterable):
        """Order-preserving, synchronous. ``func`` is emcee's wrapper around
        ``log_posterior``; it is only inspected to make sure this pool is used for
        the model it was built for."""
        target = getattr(func, 'kwargs', {}).get('model', self.model) \
            if hasattr(func, 'kwargs') else self.model
        if target is not self.model:
            raise ValueError('BatchPool was built for a different model')
        thetas = iterable if isinstance(iterable, list) else list(iterable)
        if not thetas:
            return []
        block = rows_as_block(thetas)
        lnpost = self.model.log_posterior_batch(block)
        self.calls += 1
        self.evaluations += len(thetas)
        if not self.with_blobs:
            # one shared empty blob: emcee only stores it, accumulate_images only
            # iterates over it (psfMC/models.py:84-93)
            return list(zip(lnpost.tolist(), repeat({})))
        out = []
        alive = np.isfinite(self.model.log_priors_batch(block))
        imgs = self.model.engine.render(block[alive]) if alive.any() else {}
        cursor = 0
        for row, value in enumerate(lnpost):
            if not alive[row]:
                out.append((float(value), {}))       # models.py:209-211
                continue
            out.append((float(value), {name: arr[cursor] for name, arr in imgs.items()}))
            cursor += 1
        return out

    def map_batch(self, func, block):
        """The same evaluation without the per-walker lists of emcee's protocol:
        ``block`` (B, D) -> ``(lnpost (B,), blobs)`` with ``blobs`` None unless
        ``with_blobs``. Used by this package's own sampler (sampler.py)."""
        if not self.with_blobs:
            target = getattr(func, 'kwargs', {}).get('model', self.model) \
                if hasattr(func, 'kwargs') else self.model
            if target is not self.model:
                raise ValueError('BatchPool was built for a different model')
            block = np.ascontiguousarray(block, dtype=np.float64)
            lnpost = self.model.log_posterior_batch(block)
            self.calls += 1
            self.evaluations += len(block)
            return lnpost, None
        results = self.map(func, [block[i] for i in range(len(block))])
        return (np.array([r[0] for r in results], dtype=np.float64),
                [r[1] for r in results])

    def native_sampler(self, start_positions):
        """What this package's sampler needs to run its inner loop inside the library
        (``psfmc_ensemble_run``): ``(engine, prior plan holder)``, or None when the loop
        has to stay in Python -- blobs wanted, ``PSFMC_NATIVE_SAMPLER=0``, or no prior
        plan reproduces the Python priors on ``start_positions``
        (:meth:`MultiComponentModel.native_sampler_plan`)."""
        import os
        if self.with_blobs or os.environ.get('PSFMC_NATIVE_SAMPLER', '1') == '0':
            return None
        engine = getattr(self.model, 'engine', None)
        if engine is None or not hasattr(engine, 'ensemble_run') or \
                not hasattr(self.model, 'native_sampler_plan'):
            return None
        holder = self.model.native_sampler_plan(start_positions)
        if holder is None:
            return None
        # proposals, priors and acceptance on the device too (PSFMC_ENS_DEVICE: no host
        # round trip per half-ensemble) when every prior column is one of the library's
        # families -- for ensembles of more than 512 walkers by default: below that the
        # two loops run at the same speed (the half-step is one walker's latency on the GPU
        # either way: 57 against 60 us at 250 walkers) and the host loop keeps the float64
        # repeat; at 4096 walkers the device loop is a quarter faster.
        # PSFMC_DEVICE_LOOP=1 / 0 forces / forbids it.
        choice = os.environ.get('PSFMC_DEVICE_LOOP', 'auto')
        if not holder['python_columns'] and (
                choice == '1' or (choice != '0' and len(start_positions) > 512)):
            if choice == '1' or self._float32_is_enough(start_positions):
                holder = dict(holder)
                holder['device_loop'] = True
        return engine, holder

    def _float32_is_enough(self, start_positions):
        if self._fp32_enough is None:
            self._fp32_enough = float32_is_enough(self.model, start_positions)
        return self._fp32_enough

    # multiprocessing.Pool look-alikes some callers use
    def close(self):
        pass

    def join(self):
        pass
