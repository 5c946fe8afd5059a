from . import grid
from .core import contour, from_features, from_file
