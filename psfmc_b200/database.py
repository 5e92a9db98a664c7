"""
MCMC trace database in the reference's on-disk format
(/root/reference/psfMC/database.py:6-56): one FITS binary table, a column per
model parameter (vector parameters such as ``xy`` are 2-wide columns) plus
``lnprobability``, ``walker`` and ``sample``; run metadata MCITER / MCBURN /
MCCHAINS / MCCONVRG / MCACCEPT / MAPWLKR / MAPSAMP as header keywords. Written and
read with this package's minimal FITS layer (astropy is not in the image).
"""
from collections import OrderedDict

import numpy as np

from . import fitsio

_COMMENTS = {'MCITER': 'number of retained samples',
             'MCBURN': 'number of burn-in (discarded) samples',
             'MCWALKRS': 'number of walkers run',
             'MCCONVRG': 'Has MCMC sampler converged?',
             'MCACCEPT': 'Acceptance fraction (avg of all walkers)',
             'MAPWLKR': 'Walker index of maximum posterior model',
             'MAPSAMP': 'Sample index of maximum posterior model',
             'PSFIMG': 'PSF image of maximum posterior model'}


def annotate_metadata(input_dict):
    """key -> (value, FITS comment); unknown keys are taken to be model parameters
    (cf. database.py:90-109)."""
    return OrderedDict((key, (value, _COMMENTS.get(key, 'psfMC model parameter')))
                       for key, value in input_dict.items())


def save_database(sampler, model, db_name, meta_dict=None):
    """Flatten the sampler's chain (walker-major, like the reference) into a FITS
    table and return it re-loaded from disk."""
    chain = sampler.chain
    nwalkers, nsamples, ndim = chain.shape
    flat = chain.reshape(nwalkers * nsamples, ndim)
    columns = OrderedDict()
    start = 0
    for name, length in zip(model.param_names, model.param_lens):
        block = flat[:, start:start + length]
        columns[name] = block[:, 0] if length == 1 else block
        start += length
    walker_col = np.repeat(np.arange(nwalkers, dtype=int), nsamples)
    # NB the reference labels samples with repeat(arange(nsamples), nwalkers)
    # (database.py:27), which does not match its walker-major row order; kept.
    sample_col = np.repeat(np.arange(nsamples, dtype=int), nwalkers)
    columns['lnprobability'] = np.asarray(sampler.lnprobability).ravel()
    columns['walker'] = walker_col
    columns['sample'] = sample_col
    meta = OrderedDict(meta_dict or {})
    map_row = int(np.argmax(columns['lnprobability']))
    meta['MAPWLKR'] = int(walker_col[map_row])
    meta['MAPSAMP'] = int(sample_col[map_row])
    fitsio.write_table(db_name, columns, header=annotate_metadata(meta))
    return load_database(db_name)


def load_database(db_name):
    return fitsio.read_table(db_name)


def filter_lowp_walkers(database, percentile=10):
    """Drop walkers whose samples ALL lie below the given lnprobability percentile
    (cf. database.py:112-126)."""
    cut = np.percentile(database['lnprobability'], percentile)
    keep = np.unique(database['walker'][database['lnprobability'] > cut])
    return database.select(np.isin(database['walker'], keep))


def param_matrix(database, model):
    """(rows, D) parameter vectors of a database, in the model's theta order."""
    cols = [np.asarray(database[name], dtype=np.float64).reshape(len(database), -1)
            for name in model.param_names]
    return np.concatenate(cols, axis=1) if cols else np.zeros((len(database), 0))
