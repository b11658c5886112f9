"""``maua`` import surface on top of ``maua_b200``: the reference's module paths resolve to this build's
modules, so code written against the reference -- ``from maua.audiovisual import audioreactive as ar``,
``from maua.audiovisual.patches.base.stylegan3 import StyleGAN3Patch``, ``from maua.GAN.wrappers import
get_generator_class`` -- imports unchanged.  No code lives here: a meta-path finder maps every ``maua.*`` name onto the
module (or the merge of modules) that holds the same functions in ``maua_b200``; where the layouts agree the name is mapped
by prefix, the table below lists the places where this build's layout differs from the reference's.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys
import types

_B = "maua_b200"
_AR = _B + ".audiovisual.audioreactive"
_SS = "maua.audiovisual.audioreactive.selfsupervised"

# reference module -> module(s) of this build holding its functions (several: merged, first definition wins)
_TABLE = {
    # classic ar namespace: audio.py + latent.py + mir.py + signal.py + util.py star-imported (audioreactive/__init__.py:30-34)
    "maua.audiovisual.audioreactive": [_AR + ".audio", _AR + ".mir_classic", _AR + ".latent", _AR + ".signal", _AR + ".util"],
    "maua.audiovisual.audioreactive.mir": [_AR + ".mir_classic"],
    "maua.audiovisual.audioreactive.latent": [_AR + ".latent"],
    "maua.audiovisual.audioreactive.signal": [_AR + ".signal"],
    # torch-native twins (selfsupervised/**): flat modules here
    _SS: [],
    _SS + ".features": [],
    _SS + ".features.audio": [_AR + ".features", _AR + ".chroma", _AR + ".selfsupervised"],
    _SS + ".features.processing": [_AR + ".selfsupervised"],
    _SS + ".features.efficient_quantile": [_AR + ".selfsupervised"],
    _SS + ".features.rosa": [],
    _SS + ".features.rosa.segment": [_AR + ".segment"],
    _SS + ".features.rosa.beat": [_AR + ".beat", _AR + ".selfsupervised"],
    _SS + ".latent": [_AR + ".selfsupervised"],
    _SS + ".noise": [_AR + ".noise", _AR + ".selfsupervised"],
    _SS + ".patch": [_AR + ".patch"],
    _SS + ".sample": [_AR + ".sample"],
    _SS + ".mir": [_AR + ".mir", _AR + ".selfsupervised"],
    # the in-tree inference network
    "maua.GAN.wrappers.inference": [],
    "maua.GAN.wrappers.inference.stylegan2": [_B + ".GAN.networks.stylegan2"],
    "maua.GAN.wrappers.inference.ops": [_B + ".ops"],
    # maua/ops/{image,noise,video,io}.py
    "maua.ops": [],
    "maua.ops.image": [_B + ".ops"],
    "maua.ops.noise": [_B + ".ops"],
    "maua.ops.video": [_B + ".audiovisual.render.video"],
    "maua.ops.io": [_B + ".audiovisual.render.video"],
}
# modules whose functions take host tensors in the reference (signal.py, latent.py): wrapped by the host-tensor adapter
_HOST_IO = {_AR + ".latent", _AR + ".signal"}
_PACKAGES = {k for k in _TABLE if any(o != k and o.startswith(k + ".") for o in _TABLE)} | {"maua.audiovisual.audioreactive"}


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name != "maua" and not name.startswith("maua."):
            return None
        if name in _TABLE:
            return importlib.machinery.ModuleSpec(name, self, is_package=name in _PACKAGES)
        real = _B + name[4:]
        try:
            found = importlib.util.find_spec(real)
        except (ImportError, AttributeError, ValueError):
            found = None
        if found is None:
            return None
        spec = importlib.machinery.ModuleSpec(name, self, is_package=found.submodule_search_locations is not None)
        return spec

    def create_module(self, spec):
        name = spec.name
        if name not in _TABLE:
            real = importlib.import_module(_B + name[4:])
            if not spec.submodule_search_locations and not hasattr(real, "__path__"):
                return real                                        # a plain module: the very object, under a second name
            sources, package = [real.__name__], True               # a package: its own namespace, so that its sub-modules
        else:                                                      # resolve through this finder and not to maua_b200's
            sources, package = _TABLE[name], name in _PACKAGES
        merged = types.ModuleType(name, f"{name}: maua_b200's {', '.join(sources) or '(namespace)'}")
        for source in reversed(sources):
            mod = importlib.import_module(source)
            names = getattr(mod, "__all__", None) or [n for n in vars(mod) if not n.startswith("_")]
            for n in names:
                value = getattr(mod, n)
                if package and isinstance(value, types.ModuleType) and value.__name__.startswith(_B):
                    continue                                       # sub-modules are imported on demand, by their maua.* name
                if source in _HOST_IO and isinstance(value, types.FunctionType) and value.__module__ == source:
                    from maua_b200.audiovisual.audioreactive._hostio import host_io

                    value = host_io(value)
                setattr(merged, n, value)
        if package:
            merged.__path__ = []
        return merged

    def exec_module(self, module):
        pass


if not any(isinstance(f, _Alias) for f in sys.meta_path):
    sys.meta_path.append(_Alias())
