"""
TEST / BENCH INFRASTRUCTURE ONLY -- recipe that places the UNMODIFIED reference
package where the CPU arm of bench.py can import it on the GPU box.

    python oracle/make_ref.py

The reference (mmechtley/psfMC) is pure Python, so there is nothing to compile:
the recipe copies the package directory /root/reference/psfMC verbatim into
oracle/_ref/psfMC and records a SHA-256 per file in oracle/_ref/MANIFEST.json.
oracle/_ref/ is git-ignored (nothing of the reference enters this repository's
history) but NOT gpurun-ignored, so the copy travels to the GPU box like the
built shared libraries do. It is imported only through oracle/refshim.py, and
only by `bench.py --impl reference` / bench.py's `cpu_baseline` leg and the CPU
test tier. __graft_entry__.build() runs this recipe whenever /root/reference is
present; on the GPU box (no /root/reference) the prebuilt copy is used as is.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get('PSFMC_REFERENCE_ROOT', '/root/reference')
TARGET = os.path.join(HERE, '_ref')


def _digest(path):
    with open(path, 'rb') as fobj:
        return hashlib.sha256(fobj.read()).hexdigest()


def manifest_of(root):
    out = {}
    for dirpath, dirnames, files in os.walk(root):
        dirnames[:] = sorted(d for d in dirnames if d != '__pycache__')
        for name in sorted(files):
            if name.endswith('.py'):
                full = os.path.join(dirpath, name)
                out[os.path.relpath(full, root)] = _digest(full)
    return out


def make_ref(verbose=True):
    """Copy SOURCE/psfMC -> oracle/_ref/psfMC (verbatim). Returns True if the copy
    exists afterwards."""
    src = os.path.join(SOURCE, 'psfMC')
    dst = os.path.join(TARGET, 'psfMC')
    if not os.path.isdir(src):
        if verbose:
            print('make_ref: {} not present; keeping {}'.format(
                src, 'the existing copy' if os.path.isdir(dst) else 'nothing'))
        return os.path.isdir(dst)
    want = manifest_of(src)
    have = manifest_of(dst) if os.path.isdir(dst) else None
    if have != want:
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(TARGET, exist_ok=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        with open(os.path.join(TARGET, 'MANIFEST.json'), 'w') as fobj:
            json.dump({'source': src, 'files': want}, fobj, indent=1, sort_keys=True)
        if verbose:
            print('make_ref: copied {} files to {}'.format(len(want), dst))
    elif verbose:
        print('make_ref: {} is up to date ({} files)'.format(dst, len(want)))
    return True


if __name__ == '__main__':
    sys.exit(0 if make_ref() else 1)
