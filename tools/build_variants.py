#!/usr/bin/env python
"""Development: build liblidfe variants with extra -D flags next to the product library (variants/*.so, git-ignored,
shipped to the GPU box).  usage: build_variants.py name=-DLIDFE_ABL=2 name2="-DX=1 -DY=2" ..."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
out = os.path.join(ROOT, "speech-lid_b200", "variants")
os.makedirs(out, exist_ok=True)


def one(spec):
    name, flags = spec.split("=", 1)
    lib = os.path.join(out, "liblidfe_%s.so" % name)
    cmd = ["nvcc"] + g.NVCC_FLAGS + flags.split() + ["-o", lib, os.path.join(g.CSRC, "lidfe_abi.cu")]
    subprocess.check_call(cmd)
    return lib


with ThreadPoolExecutor(8) as ex:
    for lib in ex.map(one, sys.argv[1:]):
        print("built", lib, flush=True)
