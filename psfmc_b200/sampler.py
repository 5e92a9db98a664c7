"""
Affine-invariant ensemble sampler with emcee 2.x semantics.

emcee is the reference's sampler (/root/reference/psfMC/fitting.py:5,56-58; pinned
2.2.1 in environment.yml:25) but is not installed in this image, so the stretch
move is implemented here with the same interface and the same random-number call
order per half-step -- ``rand(Ns)`` for the stretch factors, ``randint(Nc, size=Ns)``
for the partners, ``rand(Ns)`` for the acceptance -- drawn from a
``numpy.random.RandomState``, so that seeded runs are comparable with real emcee.
The posterior is evaluated through ``pool.map`` once for the starting ensemble and
then for two sequentially dependent half-ensembles per iteration: those are the
batches the GPU engine sees.

With a pool that offers ``native_sampler`` (``BatchPool`` without blobs) the iterations
themselves run inside the library (``psfmc_ensemble_run``, include/psfmc_b200.h): same
moves, same random stream -- the RandomState's MT19937 state is handed over and taken
back -- so a seeded chain is the one this module's numpy loop produces (the tests compare
them bit for bit), at ~50 us instead of ~300 us per half-ensemble of the example's
250 walkers. ``run_mcmc`` makes ONE library call for all its steps; ``sample`` one per
iteration (it has to yield in between).
"""
import ctypes

import numpy as np

__all__ = ['EnsembleSampler', 'AutocorrError', 'integrated_time']


class AutocorrError(Exception):
    """The chain is too short to estimate the autocorrelation time."""


def autocorr_function(x, axis=0):
    """Normalised autocorrelation function along ``axis`` (FFT-based)."""
    x = np.atleast_1d(x)
    n = x.shape[axis]
    f = np.fft.fft(x - np.mean(x, axis=axis, keepdims=True), n=2 * n, axis=axis)
    index = [slice(None)] * x.ndim
    index[axis] = slice(0, n)
    acf = np.fft.ifft(f * np.conjugate(f), axis=axis)[tuple(index)].real
    first = [slice(None)] * x.ndim
    first[axis] = slice(0, 1)
    return acf / acf[tuple(first)]


def integrated_time(x, low=10, high=None, step=1, c=10, axis=0):
    """Integrated autocorrelation time with emcee 2.x's windowing rule: the
    smallest window M (low <= M < high) with M > c * tau(M)."""
    size = 0.5 * x.shape[axis]
    if int(c * low) >= size:
        raise AutocorrError('The chain is too short')
    acf = autocorr_function(x, axis=axis)
    if high is None:
        high = int(size / c)
    for window in np.arange(low, high, step).astype(int):
        index = [slice(None)] * acf.ndim
        index[axis] = slice(1, window)
        tau = 1 + 2 * np.sum(acf[tuple(index)], axis=axis)
        if np.all(tau > 1.0) and window > c * np.max(tau):
            return tau
        if c * np.max(tau) >= size:
            break
    raise AutocorrError('The chain is too short to reliably estimate the '
                        'autocorrelation time')


class _FunctionWrapper(object):
    """Picklable ``f(x, *args, **kwargs)`` (what emcee hands to ``pool.map``)."""

    def __init__(self, f, args, kwargs):
        self.f = f
        self.args = list(args or [])
        self.kwargs = dict(kwargs or {})

    def __call__(self, x):
        return self.f(x, *self.args, **self.kwargs)


class EnsembleSampler(object):
    """
    :param nwalkers: ensemble size k (even, >= 2*dim unless ``live_dangerously``)
    :param dim: number of parameters
    :param lnpostfn: ``f(theta, *args, **kwargs) -> lnpost`` or ``(lnpost, blob)``
    :param a: stretch scale (2.0)
    :param pool: object with ``map(func, iterable)``; default: the builtin ``map``
    """

    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, args=None, kwargs=None,
                 pool=None, live_dangerously=False):
        if nwalkers % 2 != 0:
            raise AssertionError('The number of walkers must be even.')
        if not live_dangerously and nwalkers < 2 * dim:
            raise AssertionError('The number of walkers needs to be more than twice '
                                 'the dimension of your parameter space.')
        self.k = int(nwalkers)
        self.dim = int(dim)
        self.a = float(a)
        self.pool = pool
        self.lnprobfn = _FunctionWrapper(lnpostfn, args, kwargs)
        self._random = np.random.mtrand.RandomState()
        self.reset()

    # -- state ------------------------------------------------------------------
    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        try:
            self._random.set_state(state)
        except Exception:
            pass

    def reset(self):
        self.iterations = 0
        self.naccepted = np.zeros(self.k)
        # storage with spare capacity: emcee concatenates a fresh block onto the chain in
        # every sample() call, which copies the whole chain each time a long run is
        # advanced in pieces (fitting.advance); here the arrays grow geometrically and
        # `chain` / `lnprobability` are views of the filled part
        self._chain_buf = np.zeros((self.k, 0, self.dim))
        self._lnprob_buf = np.zeros((self.k, 0))
        self._stored = 0
        self._blobs = []

    @property
    def _chain(self):
        return self._chain_buf[:, :self._stored]

    @property
    def _lnprob(self):
        return self._lnprob_buf[:, :self._stored]

    def _reserve(self, extra):
        """Room for ``extra`` more stored iterations; returns the index of the first."""
        start, need = self._stored, self._stored + int(extra)
        if need > self._chain_buf.shape[1]:
            capacity = max(need, 2 * self._chain_buf.shape[1])
            chain = np.zeros((self.k, capacity, self.dim))
            lnprob = np.zeros((self.k, capacity))
            chain[:, :start] = self._chain_buf[:, :start]
            lnprob[:, :start] = self._lnprob_buf[:, :start]
            self._chain_buf, self._lnprob_buf = chain, lnprob
        self._stored = need
        return start

    def clear_blobs(self):
        self._blobs = []

    @property
    def chain(self):
        """(nwalkers, iterations, dim)"""
        return self._chain

    @property
    def flatchain(self):
        shape = self._chain.shape
        return self._chain.reshape(shape[0] * shape[1], shape[2])

    @property
    def lnprobability(self):
        """(nwalkers, iterations)"""
        return self._lnprob

    @property
    def flatlnprobability(self):
        return self._lnprob.flatten()

    @property
    def blobs(self):
        return self._blobs

    @property
    def acceptance_fraction(self):
        return self.naccepted / self.iterations

    def get_autocorr_time(self, low=10, high=None, step=1, c=10):
        return integrated_time(np.mean(self.chain, axis=0), axis=0, low=low,
                               high=high, step=step, c=c)

    @property
    def acor(self):
        return self.get_autocorr_time()

    # -- evaluation -------------------------------------------------------------
    def _get_lnprob(self, pos):
        pos = np.asarray(pos)
        if np.any(np.isinf(pos)):
            raise ValueError('At least one parameter value was infinite.')
        if np.any(np.isnan(pos)):
            raise ValueError('At least one parameter value was NaN.')
        if hasattr(self.pool, 'map_batch'):
            # a pool that takes the ensemble as one array (BatchPool): no per-walker
            # lists on the way in or out
            lnprob, blob = self.pool.map_batch(self.lnprobfn, pos)
            lnprob = np.array(lnprob, dtype=np.float64)
            if np.any(np.isnan(lnprob)):
                raise ValueError('lnprob returned NaN.')
            return lnprob, blob
        mapper = self.pool.map if self.pool is not None else map
        results = list(mapper(self.lnprobfn, [pos[i] for i in range(len(pos))]))
        try:
            lnprob = np.array([float(r[0]) for r in results])
            blob = [r[1] for r in results]
        except (IndexError, TypeError):
            lnprob = np.array([float(r) for r in results])
            blob = None
        if np.any(np.isnan(lnprob)):
            raise ValueError('lnprob returned NaN.')
        return lnprob, blob

    # -- the loop inside the library ------------------------------------------------
    def _native(self, pos):
        getter = getattr(self.pool, 'native_sampler', None)
        if getter is None:
            return None
        return getter(pos)

    def _native_advance(self, native, p, lnprob, first, count, start, thin, storechain):
        """Iterations ``first`` .. ``first + count - 1`` of a sample() call in one
        library call; p / lnprob / naccepted / chain are updated in place, the random
        state is taken from and returned to ``self._random``."""
        engine, holder = native
        sharded = bool(holder.get('sharded', False))
        device = bool(holder.get('device_loop', False))
        state = self._random.get_state()
        key = np.array(state[1], dtype=np.uint32)
        mt_pos = ctypes.c_int32(int(state[2]))
        if storechain:
            # emcee stores iteration `it` at start + it // thin when it % thin == 0: skip
            # ahead to the first such iteration of this call
            skip = (-first) % thin
            if skip:
                done = min(skip, count)
                engine.ensemble_run(holder['plan'], p, lnprob, key, mt_pos, done, a=self.a,
                                    n_accepted=self.naccepted, sharded=sharded, device=device)
                first, count = first + done, count - done
            if count > 0:
                engine.ensemble_run(holder['plan'], p, lnprob, key, mt_pos, count, a=self.a,
                                    chain=self._chain_buf, lnprob_chain=self._lnprob_buf,
                                    chain_start=start + first // thin, thin=thin,
                                    n_accepted=self.naccepted, sharded=sharded, device=device)
        elif count > 0:
            engine.ensemble_run(holder['plan'], p, lnprob, key, mt_pos, count, a=self.a,
                                n_accepted=self.naccepted, sharded=sharded, device=device)
        self._random.set_state((state[0], key, int(mt_pos.value), state[3], state[4]))

    def _propose_stretch(self, active, complement, lnprob_active):
        s = np.atleast_2d(active)
        c = np.atleast_2d(complement)
        ns, nc = len(s), len(c)
        zz = ((self.a - 1.0) * self._random.rand(ns) + 1) ** 2.0 / self.a
        partner = self._random.randint(nc, size=(ns,))
        q = c[partner] - zz[:, np.newaxis] * (c[partner] - s)
        newlnprob, blob = self._get_lnprob(q)
        lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnprob - lnprob_active
        accept = lnpdiff > np.log(self._random.rand(len(lnpdiff)))
        return q, newlnprob, accept, blob

    def sample(self, p0, lnprob0=None, rstate0=None, blobs0=None, iterations=1,
               thin=1, storechain=True):
        """Generator advancing the ensemble; yields ``(pos, lnprob, rstate)`` or
        ``(pos, lnprob, rstate, blobs)`` after every iteration."""
        if rstate0 is not None:
            self.random_state = rstate0
        p = np.array(p0, dtype=np.float64)
        if p.shape != (self.k, self.dim):
            raise ValueError('p0 must have shape (nwalkers, dim)')
        halfk = self.k // 2
        lnprob, blobs = lnprob0, blobs0
        if lnprob is None:
            lnprob, blobs = self._get_lnprob(p)
        lnprob = np.array(lnprob, dtype=np.float64)
        if np.any(np.isnan(lnprob)):
            raise ValueError('The initial lnprob was NaN.')
        start = self._stored
        if storechain:
            # one slot per stored iteration (it % thin == 0). emcee 2.x reserves
            # int(iterations / thin) and fails with an IndexError at the last store when
            # iterations is not a multiple of thin; here that sample is kept.
            start = self._reserve(-(-int(iterations) // int(thin)))
        native = self._native(p) if blobs is None else None
        if native is not None:
            single = getattr(self, '_one_call', False)
            steps = [int(iterations)] if single else [1] * int(iterations)
            done = 0
            for count in steps:
                if count <= 0:
                    continue
                self._native_advance(native, p, lnprob, done, count, start, int(thin),
                                     storechain)
                done += count
                self.iterations += count
                yield p, lnprob, self.random_state
            return
        halves = (slice(0, halfk), slice(halfk, self.k))
        for it in range(int(iterations)):
            self.iterations += 1
            for s0, s1 in (halves, halves[::-1]):
                q, newlnp, acc, blob = self._propose_stretch(p[s0], p[s1], lnprob[s0])
                if np.any(acc):
                    lnprob[s0][acc] = newlnp[acc]
                    p[s0][acc] = q[acc]
                    self.naccepted[s0][acc] += 1
                    if blob is not None:
                        if blobs is None:
                            raise AssertionError(
                                'If you start sampling with a given lnprob, you '
                                'also need to provide the current list of blobs.')
                        full = np.arange(self.k)[s0][acc]
                        for src, dst in zip(np.flatnonzero(acc), full):
                            blobs[dst] = blob[src]
            if storechain and it % thin == 0:
                ind = start + it // thin
                self._chain_buf[:, ind, :] = p
                self._lnprob_buf[:, ind] = lnprob
                if blobs is not None:
                    self._blobs.append(list(blobs))
            if blobs is not None:
                yield p, lnprob, self.random_state, blobs
            else:
                yield p, lnprob, self.random_state

    def run_mcmc(self, pos0, nsteps, rstate0=None, lnprob0=None, **kwargs):
        results = None
        # (with the loop in the library: one call for all steps, one yield at the end)
        self._one_call = True
        try:
            for results in self.sample(pos0, lnprob0, rstate0, iterations=nsteps, **kwargs):
                pass
        finally:
            self._one_call = False
        return results
