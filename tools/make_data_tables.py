"""Writes the data tables that ship next to the library (cuda-pt_b200/data/) from golden scene dumps:
   spectral_tables.bin  CIE 1931 XYZ + D65 (from a dump of a scene with a dispersion BSDF: sections cie_xyz, d65_spd)
   blackbody_table.bin  emission colour by normalised temperature (from a dump of a scene with a grid medium: section blackbody)
The dumps come from oracle/_ref/ref_dump, which reads the tables back from the reference at run time."""
import os, struct, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpts
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "cuda-pt_b200", "data")
os.makedirs(out, exist_ok=True)
sc = cpts.read_cpts(os.path.join(root, "tests", "golden", "ref", "grid_cbox_256.cpts"))
bb = np.ascontiguousarray(sc["blackbody"], dtype=np.float32).reshape(-1, 4)
with open(os.path.join(out, "blackbody_table.bin"), "wb") as f:
    f.write(b"CPTBBDY1" + struct.pack("<ii", bb.shape[0], 0) + bb.tobytes())
print("blackbody_table.bin:", bb.shape)
