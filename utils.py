"""Top-level `utils` module the reference's scripts import (`from utils import Logger, Losses`)."""
from py_psnode_b200.utils import Logger, Losses                                             # noqa: F401
