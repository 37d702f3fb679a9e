"""Path subpackage (mirrors ``cvpack.path``)."""

from .path import Metric, deviation, progress  # noqa: F401
