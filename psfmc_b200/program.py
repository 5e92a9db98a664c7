"""
Flatten a psfMC component list into the engine's program.

Layout rule of the parameter vector theta (reference:
/root/reference/psfMC/models.py:174-185 and ComponentBase.py:45-97): components in
model order with the PSF selector last; inside a component the free parameters in
alphabetical order of their attribute names; ``xy`` takes two slots (x, y); fixed
values take none. The functions here are duck-typed on ``_priors`` /
``_constants`` / the class name, so they accept this package's components and the
reference's own component objects alike.
"""
import numpy as np

KINDS = {'Sky': 'sky', 'PointSource': 'point', 'Sersic': 'sersic'}
ENGINE_PARAMS = {
    'sky': ('adu',),
    'point': ('x', 'y', 'mag'),
    'sersic': ('x', 'y', 'mag', 'reff', 'reff_b', 'index', 'angle'),
}


def _layout(component, offset):
    """attribute -> (first theta slot, n slots) for the component's free params."""
    where = {}
    for attr in sorted(component._priors):
        length = int(np.size(component._priors[attr].value))
        where[attr] = (offset, length)
        offset += length
    return where, offset


def compile_program(components):
    """
    :param components: model components in model order, PSF selector included
        (last), Configuration excluded
    :return: (program, psf_index_slot, num_params) where program is a list of
        ``(kind, flags, slots)`` as :class:`psfmc_b200.engine.LikelihoodEngine`
        takes it
    """
    program = []
    psf_slot = ('const', 0)
    offset = 0
    for comp in components:
        cls = type(comp).__name__
        where, offset = _layout(comp, offset)

        def slot(attr, sub=0):
            if attr in where:
                start, length = where[attr]
                if sub >= length:
                    raise ValueError('{}.{} has no element {}'.format(cls, attr, sub))
                return ('theta', start + sub)
            return ('const', float(np.ravel(comp._constants[attr])[sub]))

        if cls == 'PSFSelector':
            psf_slot = slot('psf_index')
            continue
        if cls not in KINDS:
            if hasattr(comp, 'add_to_array'):
                raise TypeError('component type {} is not supported by the engine'
                                .format(cls))
            continue
        kind = KINDS[cls]
        slots = {}
        for name in ENGINE_PARAMS[kind]:
            if name == 'x':
                slots[name] = slot('xy', 0)
            elif name == 'y':
                slots[name] = slot('xy', 1)
            else:
                slots[name] = slot(name)
        flags = {}
        if kind == 'sersic':
            flags['angle_degrees'] = bool(getattr(comp, 'angle_degrees', False))
        if kind == 'point':
            flags['shift_method'] = getattr(comp, 'shift_method', 'lanczos3')
        program.append((kind, flags, slots))
    return program, psf_slot, offset
