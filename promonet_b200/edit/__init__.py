from . import grid
