"""Subset of the reference's bev_utils needed on the hot path: camera/dataset enums and denormalisation."""
from bevgen_b200.geometry import Cameras, Dataset  # noqa: F401
from multi_view_generation.bev_utils import util  # noqa: F401
