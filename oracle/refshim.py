"""ORACLE (test infrastructure, NOT product code) -- import shim for the *real* reference.

Lets the reference's own, unmodified `crank.net.*` code (under /root/reference) run in a
container that lacks its third-party dependencies, by registering `sys.modules` stubs
*before* importing `crank.*`:

  parallel_wavegan(.models/.bin.preprocess) -> oracle.pwg / oracle.mel restatements
  librosa.filters.mel                       -> oracle.mel.mel_basis
  matplotlib, soundfile, h5py, sprocket, torch_optimizer, pytorch_lamb -> inert stubs
  numpy.long                                -> numpy.int64

/root/reference exists only in the build container, never on the GPU box: this module is
used (a) by tests/golden/make_golden.py to generate the committed golden fixtures and
(b) by CPU tests that are skipped when the reference is absent.  Nothing under `-m gpu`,
smoke() or bench.py may import it.
"""

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CRANK_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "crank"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Register stubs + put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import numpy as np

    from oracle import mel as omel
    from oracle import pwg

    if not hasattr(np, "long"):
        np.long = np.int64
    if not hasattr(np, "float"):
        np.float = np.float64

    pw = _mod("parallel_wavegan")
    pw.models = _mod(
        "parallel_wavegan.models",
        ParallelWaveGANGenerator=pwg.ParallelWaveGANGenerator,
        ParallelWaveGANDiscriminator=pwg.ParallelWaveGANDiscriminator,
        ResidualParallelWaveGANDiscriminator=pwg.ResidualParallelWaveGANDiscriminator,
    )
    pw.bin = _mod("parallel_wavegan.bin")
    pw.bin.preprocess = _mod(
        "parallel_wavegan.bin.preprocess", logmelfilterbank=omel.logmelfilterbank
    )

    def _librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
        return omel.mel_basis(sr, n_fft, n_mels, fmin, fmax)

    lib = _mod("librosa")
    lib.filters = _mod("librosa.filters", mel=_librosa_mel)

    mpl = _mod("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _mod("matplotlib.pyplot")
    _mod("soundfile")
    _mod("h5py")
    sp = _mod("sprocket")
    sp.speech = _mod("sprocket.speech", Synthesizer=object, FeatureExtractor=object)
    sp.util = _mod("sprocket.util", HDF5=object)
    _mod("torch_optimizer", RAdam=object)
    _mod("pytorch_lamb", Lamb=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def ref(name):
    """Import a module of the real reference, e.g. ref('crank.net.module.vqvae2')."""
    install()
    return importlib.import_module(name)


class NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass
