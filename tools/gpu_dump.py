"""Dump GPU stage outputs for named golden points into gpurun_out/dump_<name>.npz (diagnostics)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import acropolis_b200.flags as flags
flags.verbose = False
from acropolis_b200.input import shared_input_interface
from acropolis_b200.engine import Engine
from oracle_helpers import spec_from_golden
eng = Engine.get(shared_input_interface())
os.makedirs("gpurun_out", exist_ok=True)
for name in sys.argv[1:]:
    z = np.load(os.path.join("tests", "golden", name + ".npz"))
    sp = spec_from_golden(z)
    r = eng.run([sp], want=("G", "F", "pdi", "mpdi", "mdcy", "Yf", "status"))
    np.savez_compressed("gpurun_out/dump_%s.npz" % name, **r)
print("ok")
