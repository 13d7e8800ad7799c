"""baseline/ref_loader.py -- imports the UNMODIFIED reference model code from baseline/_ref (see install_ref.py) for the
reference arm of bench.py and the B-REF-GPU comparator.  Only what cannot run here is substituted, at import time:
  * mcubes / trimesh / igl (mesh export, plotting, the warp's closest-point query): empty stub modules -- never called on the
    canonical render / train path;
  * the two JIT-compiled extension loaders (encoder/*/backend.py, `-std=c++14`, rejected by torch >= 2.1 headers):
      on CUDA  -> the reference's own hashencoder.cu compiled for sm_100a by oracle/build_ref.py (oracle/_ref/_ref_hash_encoder.so)
      on CPU   -> oracle/hashgrid.py, the C restatement of that kernel (the reference has no CPU hash kernel at all).
Test / measurement infrastructure only: nothing in avatarcraft_b200/ imports this."""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF, "models", "instant_nsr.py"))


def _ext(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader(name, loader))
    loader.exec_module(mod)
    return mod


def load_reference(cuda: bool):
    """The reference's `models.instant_nsr` module (NeRFNetwork, NeRFRenderer, near_far_from_bound, ...)."""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run baseline/install_ref.py in the build container")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    for name in ("mcubes", "trimesh", "igl"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if cuda:
        hash_be = _ext("_ref_hash_encoder")
        if hash_be is None:
            raise RuntimeError("oracle/_ref/_ref_hash_encoder.so missing (oracle/build_ref.py)")
        sh_be = _ext("_ref_sh_encoder") or hash_be
    else:
        from oracle import hashgrid as ohg
        hash_be = sh_be = types.SimpleNamespace(hash_encode_forward=ohg.hash_encode_forward, hash_encode_backward=ohg.hash_encode_backward)
    for pkg, be in (("encoder.hashencoder.backend", hash_be), ("encoder.shencoder.backend", sh_be)):
        m = types.ModuleType(pkg)
        m._backend = be
        sys.modules[pkg] = m
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "encoder" or
              (k.startswith("encoder.") and not k.endswith(".backend")) or k == "utils" or k.startswith("utils.") or k == "geometry" or k.startswith("geometry.")]:
        del sys.modules[k]                                   # a previous load with the other backend
    sys.path.insert(0, REF)
    try:
        warnings.simplefilter("ignore")
        return importlib.import_module("models.instant_nsr")
    finally:
        sys.path.remove(REF)
