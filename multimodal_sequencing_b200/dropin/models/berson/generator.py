"""models/berson/generator.py of the reference: same Beam as models/beam.py."""
from ..beam import Beam  # noqa: F401
