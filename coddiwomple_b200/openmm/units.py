"""Unit handling without simtk: plain numbers are MD units (nm, ps, K, kJ/mol, amu); simtk/openmm Quantities are
converted when those packages happen to be installed."""


def strip(q, default=None):
    """Return a float / ndarray in the MD unit system."""
    if q is None:
        return default
    if hasattr(q, "value_in_unit_system"):
        try:
            from openmm import unit
        except ImportError:
            from simtk import unit
        return q.value_in_unit_system(unit.md_unit_system)
    if hasattr(q, "_value"):
        return q._value
    return q
