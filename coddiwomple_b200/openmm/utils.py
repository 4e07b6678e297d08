"""OpenMM utilities (coddiwomple/openmm/utils.py).  `get_dummy_integrator` only exists to create an OpenMM Context in
the reference; no Context exists here, so it returns None."""
import numpy as np


def get_dummy_integrator():
    return None


def read_pdb_frames(path):
    """Minimal reader for the `*.cache.pdb` wire format the reference loads with mdtraj (openmm/coddiwomple.py:286):
    fixed-column ATOM records in Angstrom, frames delimited by ENDMDL.  Returns float32 nm (float32(A)/float32(10))."""
    frames, cur = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("ATOM") or line.startswith("HETATM"):
                cur.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            elif line.startswith("ENDMDL"):
                frames.append(cur)
                cur = []
    if cur:
        frames.append(cur)
    return (np.asarray(frames, dtype=np.float32) / np.float32(10.0)).astype(np.float32)


def load_frames(source):
    """Frames (n, A, 3) float32 nm from a PDB path, a .npy path or an array."""
    if isinstance(source, str):
        if source.endswith(".npy"):
            return np.load(source).astype(np.float32)
        return read_pdb_frames(source)
    return np.asarray(source, dtype=np.float32)
