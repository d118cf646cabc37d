#!/usr/bin/env python
"""Development: time every library variant under speech-lid_b200/variants/ on cfg2 shapes (one process per variant,
LIDFE_LIB_PATH), modes given on the command line.  Prints one line per (variant, mode)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
modes = sys.argv[1] if len(sys.argv) > 1 else "none,utt,global_accum"
pick = sys.argv[2].split(",") if len(sys.argv) > 2 else None
libs = sorted(glob.glob(os.path.join(ROOT, "speech-lid_b200", "variants", "liblidfe_*.so")))
for lib in libs:
    name = os.path.basename(lib)[len("liblidfe_"):-3]
    if pick and name not in pick:
        continue
    env = dict(os.environ, LIDFE_LIB_PATH=lib)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dev_perf.py"), modes], env=env, capture_output=True, text=True, timeout=120)
        out = [l for l in r.stdout.splitlines() if "us/step" in l]
        for l in out:
            print("%-10s %s" % (name, l.split("mode=")[1]), flush=True)
        if r.returncode != 0:
            print(name, "rc", r.returncode, r.stderr[-300:], flush=True)
    except subprocess.TimeoutExpired:
        print(name, "TIMEOUT", flush=True)
