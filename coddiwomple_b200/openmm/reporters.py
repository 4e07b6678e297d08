"""OpenMM Reporter module (coddiwomple/openmm/reporters.py:14-125).

`OpenMMReporter.record(particles, save_to_disk)` keeps, per particle, the list of recorded positions and writes one
multi-model PDB per particle (`<prefix>.<index:04>.pdb`) — the same wire format as the reference's end-state caches
(fixed-column ATOM records in Angstrom with 3 decimals, frames delimited by MODEL/ENDMDL; mdtraj is not needed).
With `trajectory_directory=None` nothing is written, which is how the reference's own SMC test uses it
(coddiwomple/tests/legacy_tests.py:83-84).  Recording is outside the hot path: a `ParticleList` is snapshotted with one
device-to-host copy per call.
"""
import os

import numpy as np

from ..particles import ParticleList
from ..reporters import Reporter


def write_pdb_trajectory(path, frames_nm, atom_names=None, elements=None):
    """frames_nm: (n_frames, A, 3) in nm."""
    frames = np.asarray(frames_nm, dtype=np.float64) * 10.0
    A = frames.shape[1]
    atom_names = atom_names or [f"A{a + 1}" for a in range(A)]
    elements = elements or [n[0] for n in atom_names]
    with open(path, "w") as fh:
        fh.write("REMARK   1 CREATED WITH coddiwomple_b200\n")
        fh.write("CRYST1   20.000   20.000   20.000  90.00  90.00  90.00 P 1           1 \n")
        for m, xyz in enumerate(frames):
            fh.write(f"MODEL     {m:4d}\n")
            for a in range(A):
                fh.write(f"ATOM  {a + 1:5d} {atom_names[a]:<4s} MOL A   1    {xyz[a, 0]:8.3f}{xyz[a, 1]:8.3f}{xyz[a, 2]:8.3f}  1.00  0.00          {elements[a]:>2s}  \n")
            fh.write(f"TER   {A + 1:5d}      MOL A   1\n")
            fh.write("ENDMDL\n")
        fh.write("END\n")


class OpenMMReporter(Reporter):
    def __init__(self, trajectory_directory=None, trajectory_prefix=None, md_topology=None, subset_indices=None, atom_names=None, **kwargs):
        self.trajectory_directory, self.trajectory_prefix = trajectory_directory, trajectory_prefix
        self.write_traj = trajectory_directory is not None and trajectory_prefix is not None
        self.atom_names = atom_names
        self.subset_indices = subset_indices
        self.hex_dict, self.hex_to_index, self.hex_counter = {}, {}, 0
        if self.write_traj:
            os.makedirs(trajectory_directory, exist_ok=True)
            self.neq_traj_filename = os.path.join(trajectory_directory, trajectory_prefix)

    def _key(self, particle, slot):
        return f"slot{slot}" if slot is not None else hex(id(particle))

    def record(self, particles, save_to_disk=False, **kwargs):
        if not self.write_traj:
            return
        if isinstance(particles, ParticleList):
            pos = particles.ensemble.positions.cpu().numpy()
            items = [(f"slot{i}", pos[i]) for i in range(len(particles))]
        else:
            items = [(hex(id(p)), np.asarray(p.state.positions)) for p in particles]
        for key, xyz in items:
            if key not in self.hex_dict:
                self.hex_dict[key] = []
                self.hex_to_index[key] = self.hex_counter
                self.hex_counter += 1
            x = xyz if self.subset_indices is None else xyz[self.subset_indices]
            self.hex_dict[key].append(np.array(x, dtype=np.float64))
        if save_to_disk:
            for key in self.hex_dict:
                self._write_trajectory(key, f"{self.neq_traj_filename}.{format(self.hex_to_index[key], '04')}.pdb")

    def _write_trajectory(self, key, filename):
        write_pdb_trajectory(filename, np.stack(self.hex_dict[key]), atom_names=self.atom_names)
