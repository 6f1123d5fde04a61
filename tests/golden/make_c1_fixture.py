"""Config 1 fixture: a bundled mesh of the reference (source/data/nefertiti.off: 299 V / 562 F, open disk) taken through
the CGAL-free front end (surface-remesher_b200/frontend.py).  Runs in the authoring container, where /root/reference
exists; stores only the derived arrays (parameterised points, weights, triangles, border)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from surface_remesher_b200 import frontend as FE

V, F = FE.read_off("/root/reference/source/data/nefertiti.off")
prep = FE.prepare(V, F, 1024)
np.savez_compressed(os.path.join(HERE, "c1_nefertiti.npz"), V=V, F=F, uv=prep["uv"], loop=prep["loop"], weights=prep["weights"])
print("saved", len(V), len(F), "border", len(prep["loop"]), "weights", prep["weights"].min(), prep["weights"].max())
