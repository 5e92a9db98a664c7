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
import numpy as np


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
        try:
            # emcee hands over the rows of one (B, D) array: one concatenation is 3-5x
            # cheaper than stacking B one-row arrays
            block = np.concatenate(thetas).reshape(len(thetas), -1)
            if block.dtype != np.float64 or thetas[0].ndim != 1:
                raise ValueError
        except (ValueError, TypeError, AttributeError):
            block = np.stack([np.asarray(p, dtype=np.float64) for p in thetas])
        lnpost = self.model.log_posterior_batch(block)
        self.calls += 1
        self.evaluations += len(thetas)
        if not self.with_blobs:
            return [(v, {}) for v in lnpost.tolist()]
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

    # multiprocessing.Pool look-alikes some callers use
    def close(self):
        pass

    def join(self):
        pass
