from . import ops
