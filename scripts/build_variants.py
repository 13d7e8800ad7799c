"""Build tuning variants of the library into avatarcraft_b200/_variants/ (git-ignored, travels with gpurun).
usage: python scripts/build_variants.py name1:DEF1,DEF2 name2:DEF3=4 ...   (name 'base' with no defines = the shipped flags)"""
import os, sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avatarcraft_b200 import _lib
out = os.path.join(os.path.dirname(_lib.__file__), "_variants")
os.makedirs(out, exist_ok=True)
def one(spec):
    name, _, defs = spec.partition(":")
    path = os.path.join(out, f"{name}.so")
    _lib.build_variant(path, [d for d in defs.split(",") if d])
    return path
with ThreadPoolExecutor(4) as ex:
    for p in ex.map(one, sys.argv[1:]):
        print("built", p)
