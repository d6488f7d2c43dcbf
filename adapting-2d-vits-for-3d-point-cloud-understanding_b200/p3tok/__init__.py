"""p3tok - B200-native point-patch tokenizer (FPS -> kNN -> gather/normalise -> mini-PointNet embed).

Host-side mirror of the reference's tokenizer interface (src/models/apf.py, src/models/pix4point.py,
src/data/sampler.py of Irish-77/adapting-2D-ViTs-for-3D-point-cloud-understanding) over a C-ABI
library of hand-written sm_100a kernels.  CUDA only: there is no CPU fallback.
"""
__all__ = ["synth"]


def __getattr__(name):
    # torch-dependent submodules are imported lazily so that `p3tok.synth` (numpy only) stays light
    import importlib
    if name in ("ops", "functional", "modules", "fold", "shard", "_lib", "_build", "synth"):
        return importlib.import_module("." + name, __name__)
    for mod in ("modules", "functional"):
        m = importlib.import_module("." + mod, __name__)
        if hasattr(m, name):
            return getattr(m, name)
    raise AttributeError(name)
