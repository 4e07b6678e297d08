"""Reporter module — abstract interface with the reference's name (coddiwomple/reporters.py:14-22).
Disk I/O is outside the hot path; `record` is a no-op hook."""


class Reporter():
    """Generalized reporter object for Particles"""

    def __init__(self, **kwargs):
        pass

    def record(self, particles, save_to_disk=False, **kwargs):
        pass
