"""B200-native counterparts of the reference's `coddiwomple.openmm` adapters (same class and function names)."""
