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

# horse.off (3088 V / 6172 F, closed, genus 0) + its seam (data/horse.selection.txt, 44 edges): long-edge split of the
# seam (0.72, main.cpp:150,165), cut to a disk, Tutte map, area-ratio weights — the mesh SURVEY §8(d) names for config 1
V, F = FE.read_off("/root/reference/source/data/horse.off")
pairs = FE.read_seam_pairs("/root/reference/source/data/horse.selection.txt")
prep = FE.prepare(V, F, 1024, seam_pairs=pairs)
np.savez_compressed(os.path.join(HERE, "c1_horse.npz"), V=prep["V"].astype(np.float64), F=prep["F"], uv=prep["uv"], loop=prep["loop"],
                    weights=prep["weights"], orig=prep["orig"], V_src=V, F_src=F, seam_pairs=np.asarray(pairs, np.int32))
print("saved horse: cut mesh", len(prep["V"]), len(prep["F"]), "border", len(prep["loop"]), "weights", prep["weights"].min(), prep["weights"].max())
