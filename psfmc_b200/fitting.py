"""
model_galaxy_mcmc with the reference's signature and behaviour
(/root/reference/psfMC/fitting.py:13-113): build the model from a model file, burn
in, sample with the stretch-move ensemble sampler, write the FITS trace database
and the posterior images. The one difference: the sampler gets
``pool=BatchPool(model)``, so each (half-)ensemble is one GPU batch, and posterior
images are re-rendered on the GPU from the stored chain instead of travelling to
the host as blobs at every step.
"""
import os
from collections import OrderedDict
from warnings import warn

import numpy as np

from . import fitsio
from .database import (annotate_metadata, filter_lowp_walkers, load_database,
                       param_matrix, save_database)
from .models import IMAGE_TYPES, MultiComponentModel
from .pool import BatchPool
from .sampler import AutocorrError, EnsembleSampler

default_filetypes = ('raw_model', 'convolved_model', 'composite_ivm', 'residual',
                     'point_source_subtracted')


def print_progress(step, total, label=''):
    """Percent-complete line (cf. psfMC/utils.py:167-171)."""
    if total and (step + 1) % max(1, total // 10) == 0:
        print('{}: {:d}%'.format(label, int(100 * (step + 1) / total)))


def advance(sampler, pos, lnprob, total, label='', verbose=True):
    """``total`` sampler iterations from ``pos`` in at most ten ``run_mcmc`` calls, a
    progress line after each (the reference prints from inside its per-iteration loop,
    psfMC/fitting.py:66-83; with the loop inside the library -- sampler.py -- a call
    covers a tenth of the run). Returns the final ``(pos, lnprob)``."""
    done, chunk = 0, max(1, int(total) // 10)
    while done < total:
        count = min(chunk, total - done)
        pos, lnprob = sampler.run_mcmc(pos, count, lnprob0=lnprob)[:2]
        sampler.clear_blobs()
        done += count
        if verbose:
            print_progress(done - 1, total, label)
    return pos, lnprob


def check_convergence_autocorr(sampler, min_chain_to_tau_ratio=10, verbose=0):
    """True when the chain is longer than ``min_chain_to_tau_ratio`` integrated
    autocorrelation times of every parameter (cf. analysis/statistics.py:134-155)."""
    try:
        acorr = sampler.get_autocorr_time(c=1)
    except AutocorrError:
        warn('unable to estimate the autocorrelation time, assuming chain is not '
             'converged')
        return False
    if verbose > 0:
        print('Autocorrelation times: {}'.format(acorr))
    return bool(np.all(sampler.chain.shape[1] > min_chain_to_tau_ratio * acorr))


def add_stats_to_header(header, model, database):
    """Sampler metadata and the posterior summary of every stochastic parameter as
    header cards (cf. analysis/images.py:104-143): the MC*/MAP* keywords of the trace
    database, then per parameter, under its FITS abbreviation, 'mean +/- std' (tuple form
    for vector parameters such as xy), then PSFIMG = the PSF file of the model."""
    header.commentary.append(('COMMENT', 'psfMC MCMC SAMPLER PARAMETERS'))
    for key, (value, comment) in annotate_metadata(
            OrderedDict((k, database.meta[k]) for k in database.meta
                        if k.startswith(('MC', 'MAP')))).items():
        header.set(key, value, comment)
    header.commentary.append(('COMMENT', 'psfMC POSTERIOR MODEL INFORMATION'))
    stats = OrderedDict()
    for col_name, fits_abbr in zip(model.param_names, model.param_fits_abbrs):
        column = np.asarray(database[col_name], dtype=np.float64)
        mean_post, std_post = np.mean(column, axis=0), np.std(column, axis=0)
        if np.ndim(mean_post) == 0:
            value = '{:0.4g} +/- {:0.4g}'.format(float(mean_post), float(std_post))
        else:
            value = '({}) +/- ({})'.format(
                ','.join('{:0.4g}'.format(dim) for dim in mean_post),
                ','.join('{:0.4g}'.format(dim) for dim in std_post))
        stats[fits_abbr] = value
    selector = model.config.psf_selector
    index = 0
    if len(selector.filenames) > 1 and 'PSF_Index' in getattr(database, 'colnames',
                                                               list(database.keys())):
        best_row = int(np.argmax(database['lnprobability']))
        index = int(np.rint(np.asarray(database['PSF_Index'])[best_row]))
        index = min(max(index, 0), len(selector.filenames) - 1)
    name = selector.filenames[index]
    stats['PSFIMG'] = name if isinstance(name, str) else 'array'
    for key, (value, comment) in annotate_metadata(stats).items():
        header.set(key, value, comment)


def save_posterior_images(model, database, output_name='out_{}', mode='weighted',
                          filetypes=default_filetypes, bad_px_value=0,
                          walker_min_percentile=10):
    """Posterior-mean ('weighted') or maximum-a-posteriori ('maximum'/'MAP')
    images as FITS files (cf. analysis/images.py:17-101); every database row is
    rendered on the GPU."""
    header = model.obs_header.copy()
    if '{}' not in output_name:
        output_name += '_{}'
    database = filter_lowp_walkers(database, percentile=walker_min_percentile)
    unknown = set(ftype for ftype in filetypes if ftype not in IMAGE_TYPES)
    if unknown:
        warn('Unknown filetypes requested: {} Output images will not be generated '
             'for these types.'.format(unknown))
    filetypes = [ftype for ftype in filetypes if ftype in IMAGE_TYPES]
    thetas = param_matrix(database, model)
    output = {}
    if mode in ('maximum', 'MAP'):
        best = int(np.argmax(database['lnprobability']))
        imgs = model.sample_images(thetas[best], which=filetypes)
        output = {ftype: np.array(imgs[ftype]) for ftype in filetypes}
    elif mode == 'weighted':
        if len(database) != model.accumulated_samples:
            model.reset_images()
            model.accumulate_from_chain(thetas, which=filetypes)
        output = {ftype: np.array(model.posterior_images[ftype]) for ftype in filetypes}
    else:
        warn('Unknown posterior output mode ({}). Posterior model images will not '
             'be saved.'.format(mode))
        return None
    add_stats_to_header(header, model, database)
    written = []
    for ftype in filetypes:
        img = output[ftype]
        img[~np.isfinite(img)] = bad_px_value
        header.set('OBJECT', ftype)
        fname = output_name.format(ftype) + '.fits'
        fitsio.writeto(fname, img, header=header, overwrite=True)
        written.append(fname)
    return written


def model_galaxy_mcmc(model_file, output_name=None, write_fits=default_filetypes,
                      iterations=0, burn=0, chains=None, max_iterations=1,
                      convergence_check=check_convergence_autocorr,
                      precision='fp32', devices=None, seed=None, verbose=True):
    """Same arguments as the reference plus ``precision`` / ``devices`` (engine)
    and ``seed`` (reproducible runs). Returns the trace database table."""
    if output_name is None:
        output_name = 'out_' + model_file.replace('.py', '')
    output_name += '_{}'
    mc_model = model_file if isinstance(model_file, MultiComponentModel) else \
        MultiComponentModel(components=model_file, precision=precision,
                            devices=devices)
    if chains is None:
        chains = 2 * mc_model.num_params + 2
    sampler = EnsembleSampler(nwalkers=chains, dim=mc_model.num_params,
                              lnpostfn=mc_model.log_posterior,
                              kwargs={'model': mc_model},
                              pool=BatchPool(mc_model))
    if seed is not None:
        sampler._random.seed(seed)
        np.random.seed(seed)
    db_name = output_name.format('db') + '.fits'
    if not os.path.exists(db_name):
        param_vec = mc_model.init_params_from_priors(chains)
        param_vec, lnprob = advance(sampler, param_vec, None, burn, 'Burning', verbose)
        sampler.reset()
        converged = False
        for sampling_iter in range(max_iterations):
            # (every round starts from the end of the burn-in, like the reference's)
            advance(sampler, param_vec, lnprob, iterations, 'Sampling', verbose)
            if convergence_check(sampler):
                converged = True
                break
            warn('Not yet converged after {:d} iterations:'.format(
                (sampling_iter + 1) * iterations))
        metadata = OrderedDict([
            ('MCITER', sampler.chain.shape[1]), ('MCBURN', burn),
            ('MCCHAINS', chains), ('MCCONVRG', converged),
            ('MCACCEPT', float(sampler.acceptance_fraction.mean()))])
        database = save_database(sampler, mc_model, db_name, meta_dict=metadata)
    else:
        print('Database already contains sampled chains, skipping sampling')
        database = load_database(db_name)
    if write_fits:
        save_posterior_images(mc_model, database, output_name=output_name,
                              filetypes=write_fits)
    return database
