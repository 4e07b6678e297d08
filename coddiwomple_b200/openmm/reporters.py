"""OpenMM Reporter module (coddiwomple/openmm/reporters.py).  Trajectory I/O (mdtraj) is outside the hot path;
`record` is callable and a no-op when no directory is given, which is how the reference's own SMC test uses it
(coddiwomple/tests/legacy_tests.py:83-84)."""
from ..reporters import Reporter


class OpenMMReporter(Reporter):
    def __init__(self, trajectory_directory=None, trajectory_prefix=None, md_topology=None, subset_indices=None, **kwargs):
        self.trajectory_directory = trajectory_directory
        self.trajectory_prefix = trajectory_prefix
        self.write_traj = trajectory_directory is not None and trajectory_prefix is not None
        if self.write_traj:
            raise NotImplementedError("writing trajectories is out of scope for the B200 hot path; pass directory_name=None")

    def record(self, particles, save_to_disk=False, **kwargs):
        pass
