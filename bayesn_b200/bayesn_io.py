"""The reference's ``bayesn.bayesn_io`` name (``bayesn/bayesn_io.py:13-132``): ``write_snana_lcfile``."""
from .snana import write_snana_lcfile, read_snana_ascii  # noqa: F401
